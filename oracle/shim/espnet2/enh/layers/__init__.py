"""Oracle shim package (test infrastructure only): restated subset of espnet==202412, see oracle/README.md."""
