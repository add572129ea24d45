"""No-op stand-in so the read-only reference imports in a container without matplotlib.
Only used by tests/golden/make_goldens.py (golden generation); never by the product."""
