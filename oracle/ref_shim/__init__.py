"""Test infrastructure: lets the reference's own hot-path source files run on NumPy (see paddle/__init__.py)."""
