"""Oracle shim: empty pyplot."""
