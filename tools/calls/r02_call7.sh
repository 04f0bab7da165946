#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/fan_variants.py 11 tools/fan_variants_b.txt > gpurun_out/r02_c7_fan_variants.jsonl 2> gpurun_out/r02_c7_fan_variants.err
tail -5 gpurun_out/r02_c7_fan_variants.err
cut -c1-330 gpurun_out/r02_c7_fan_variants.jsonl
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c7_gpu_tests.log 2>&1
tail -5 gpurun_out/r02_c7_gpu_tests.log
