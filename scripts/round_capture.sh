#!/bin/bash
# End-of-round capture on one B200 (run under gpurun): tests, phase times, bench (ours + reference arm),
# ncu launch list of one whole pass, ncu --set full of the heaviest kernels.  Outputs in gpurun_out/.
set -u
R=${1:-r04}
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) 2>&1 | tail -6 > gpurun_out/pytest_gpu_$R.txt
python scripts/time_phases.py 8192 > gpurun_out/phases_$R.txt 2>&1
python scripts/time_phases.py 2046 1000 media > gpurun_out/phases_media_$R.txt 2>&1
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_${R}_builder.json 2> gpurun_out/bench_${R}_err.log
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_${R}_reference.json 2>> gpurun_out/bench_${R}_err.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 52 -c 60 --csv --log-file gpurun_out/launches_$R.csv \
    python scripts/profile_run.py 8192 > gpurun_out/ncu_launch_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_score3 -s 6 -c 2 -o gpurun_out/prof_${R}_score3 -f \
    python scripts/profile_run.py 2048 > gpurun_out/ncu_full_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_stitch -s 3 -c 1 -o gpurun_out/prof_${R}_stitch -f \
    python scripts/profile_run.py 2048 >> gpurun_out/ncu_full_$R.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_gmm_bic<\(int\)5>' -s 1 -c 1 \
    -o gpurun_out/prof_${R}_gmm_bic5 -f python scripts/profile_run.py 2048 >> gpurun_out/ncu_full_$R.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_gmm_lloyd<\(int\)5>' -s 2 -c 1 \
    -o gpurun_out/prof_${R}_gmm_lloyd5 -f python scripts/profile_run.py 2048 >> gpurun_out/ncu_full_$R.log 2>&1
tail -3 gpurun_out/pytest_gpu_$R.txt; grep -E "whole|refit|score|fast" gpurun_out/phases_$R.txt; head -c 400 gpurun_out/bench_${R}_builder.json; ls -la gpurun_out/prof_${R}_*
