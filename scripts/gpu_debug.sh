#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -rf 2>&1 | tail -100 > gpurun_out/dbg_pytest.txt
python scripts/score_phase_profile.py 2048 hotel > gpurun_out/dbg_score_phases.txt 2>&1
python scripts/score_phase_profile.py 1020 media >> gpurun_out/dbg_score_phases.txt 2>&1
python scripts/stitch_phase_profile.py 4096 hotel > gpurun_out/dbg_stitch_phases.txt 2>&1
python scripts/stitch_phase_profile.py 1020 media >> gpurun_out/dbg_stitch_phases.txt 2>&1
python scripts/time_phases.py 8192 > gpurun_out/dbg_phases.txt 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/dbg_bench.json 2> gpurun_out/dbg_bench.err
tail -3 gpurun_out/dbg_pytest.txt
