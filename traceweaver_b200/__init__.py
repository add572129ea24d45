"""traceweaver_b200 — B200-native engine for TraceWeaver's span-assignment hot path.

Scope: `TraceWeaverV3.FindAssignments` (method "MaxScoreBatchSubsetWithSkips") of
Sachin-A/TraceWeaver, nothing else.  The compute lives in
csrc/ (hand-written sm_100a CUDA behind a C ABI, include/traceweaver_b200.h); this package is the
host-side mirror of the reference's predictor interface."""
from . import _abi  # noqa: F401
from .batch import Problem, HostBatch, build_batch  # noqa: F401

__all__ = ["Problem", "HostBatch", "build_batch"]
