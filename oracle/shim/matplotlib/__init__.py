"""Oracle shim: empty stand-in so `import matplotlib.pyplot` in baseline_code/sampling/__init__.py:12 resolves."""
