#!/bin/bash
# one GPU call: parity suite, memcheck of the scoring kernels on a stress case, short bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -60 > gpurun_out/r2_tests.log
if ! grep -q " passed" gpurun_out/r2_tests.log || grep -q "failed" gpurun_out/r2_tests.log; then
  timeout 900 compute-sanitizer --tool memcheck --launch-timeout 0 python -m pytest tests/test_gpu_stress.py -q -x -k "nginx_parallel or hotel_overload" 2>&1 | grep -v "^=========     at\|^=========         in\|frame #" | head -150 > gpurun_out/r2_memcheck.log
fi
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
tail -5 gpurun_out/r2_tests.log
