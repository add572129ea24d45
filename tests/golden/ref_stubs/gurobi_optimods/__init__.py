"""Stand-in package for gurobi-optimods==1.1.0 (requirements.txt:5; needs a Gurobi licence,
absent here).  Only `mwis.maximum_weighted_independent_set` is called by the reference
(traceweaver_v3.py:1411)."""
