#!/bin/bash
# one GPU call: parity suite, phase times, the driver's bench command (ours + CPU arm), ncu launch list
set -u
T=${1:-a}
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --tb=short -rf ) 2>&1 | tail -60 > gpurun_out/${T}_pytest.txt
python scripts/time_phases.py 8192 > gpurun_out/${T}_phases.txt 2>&1
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err ) 2> gpurun_out/${T}_bench.time
( time python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2>> gpurun_out/${T}_bench.err ) 2> gpurun_out/${T}_bench_ref.time
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/${T}_launches.csv \
    python scripts/profile_run.py 8192 > gpurun_out/${T}_ncu_launch.log 2>&1
tail -3 gpurun_out/${T}_pytest.txt; grep -E "whole|refit|score|fast" gpurun_out/${T}_phases.txt; head -c 1500 gpurun_out/${T}_bench.json; cat gpurun_out/${T}_bench.time
