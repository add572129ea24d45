#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -rf 2>&1 | tail -300 > gpurun_out/dbg_pytest.txt
python scripts/compare_stages.py media 2046 1000 2 > gpurun_out/dbg_media.txt 2>&1
python scripts/compare_stages.py alibaba 2016 1250 3 > gpurun_out/dbg_alibaba.txt 2>&1
python scripts/time_phases.py 8192 > gpurun_out/dbg_phases.txt 2>&1
tail -5 gpurun_out/dbg_pytest.txt
