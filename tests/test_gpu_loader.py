"""Loader -> engine end to end on the GPU: synthetic requests written as Jaeger JSON trace files,
read back by traceweaver_b200.loader, solved by the batch API, compared with the CPU oracle on the
same arrays and scored against the traces' own parent links (the reference's AccuracyForService)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_trace_directory_to_assignments(tmp_path):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test needs a CUDA device")
    from oracle import tw_oracle
    from test_loader import _write_traces
    from traceweaver_b200 import synth
    from traceweaver_b200.api import BatchSolver
    from traceweaver_b200.loader import load_jaeger_dir, to_host_batch, accuracy
    services = []
    for q, (shape, load) in enumerate([("hotel_frontend", 120.0), ("hotel_search", 150.0), ("media_nginx", 60.0)]):
        blk = synth.make_block(shape, 1, 400, load, seed=20 + q)
        d = tmp_path / f"dir{q}"
        d.mkdir()
        _write_traces(blk, 0, str(d), [f"svc{q}_{e}" for e in range(len(blk.out_start))])
        loaded = load_jaeger_dir(str(d))
        assert len(loaded) == 1
        # ground truth + FindOrder derived on the device (row f-2) == the NumPy derivation
        from traceweaver_b200.engine import Engine
        eng = Engine(0)
        on_dev = load_jaeger_dir(str(d), engine=eng)
        eng.close()
        assert len(on_dev) == 1 and on_dev[0].out_eps == loaded[0].out_eps
        assert np.array_equal(on_dev[0].truth, loaded[0].truth)
        assert on_dev[0].graph_edges == loaded[0].graph_edges and on_dev[0].problem.preds == loaded[0].problem.preds
        for a, b in zip(on_dev[0].problem.out_start, loaded[0].problem.out_start):
            assert np.array_equal(a, b)
        services += loaded
    hb = to_host_batch(services)
    solver = BatchSolver(device=0, seed_select=10)
    out = solver.solve(hb)
    solver.close()
    ref = tw_oracle.find_assignments(hb, 10, threads=2)
    assert np.array_equal(out["assign"], ref["assign"])
    assert np.array_equal(out["topk_idx"], ref["topk_idx"])
    for p, svc in enumerate(services):
        a = out["assign"][int(hb.prob_tuple_off[p]):int(hb.prob_tuple_off[p + 1])]
        assert accuracy(svc, a) > 0.8, svc.name
