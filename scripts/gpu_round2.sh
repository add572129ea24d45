#!/bin/bash
# GPU call: full parity suite, quick bench, score phase counters, fused-refit variant check
set -u
T=${1:-g}
mkdir -p gpurun_out
( time timeout 420 python -m pytest tests -m gpu -q --tb=short -x ) 2>&1 | tail -25 > gpurun_out/${T}_pytest.txt
timeout 200 python bench.py --steps 5 --warmup 3 --no-extra 2>gpurun_out/${T}_bench.err > gpurun_out/${T}_bench.json
timeout 120 python scripts/score_phase_profile.py 2048 hotel > gpurun_out/${T}_score_phases.txt 2>&1
timeout 120 python scripts/time_gmm.py 4096 > gpurun_out/${T}_gmm_time.txt 2>&1
TW_SO=traceweaver_b200/libtw_b200_fused.so timeout 120 python scripts/time_gmm.py 4096 >> gpurun_out/${T}_gmm_time.txt 2>&1
( TW_B200_SO=$PWD/traceweaver_b200/libtw_b200_fused.so timeout 300 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_stress.py -m gpu -q --tb=line -k "not 120k" ) 2>&1 | tail -15 > gpurun_out/${T}_fused_pytest.txt
tail -6 gpurun_out/${T}_pytest.txt; head -c 300 gpurun_out/${T}_bench.json; echo; cat gpurun_out/${T}_score_phases.txt | tail -12; cat gpurun_out/${T}_gmm_time.txt | tail -4; tail -5 gpurun_out/${T}_fused_pytest.txt
