#!/bin/bash
# Round 2, call 2 (1 GPU): native plan builder + tile kernel v2 on hardware
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02_c2_parity.log 2>&1
tail -15 gpurun_out/r02_c2_parity.log
timeout 900 python tools/scale_probe.py vecp2 512 1024 2048 3162 > gpurun_out/r02_c2_probe_vecp2.jsonl 2> gpurun_out/r02_c2_probe_vecp2.err
cat gpurun_out/r02_c2_probe_vecp2.jsonl; tail -3 gpurun_out/r02_c2_probe_vecp2.err
timeout 900 python tools/scale_probe.py tetp2 32 64 100 160 203 > gpurun_out/r02_c2_probe_tetp2.jsonl 2> gpurun_out/r02_c2_probe_tetp2.err
cat gpurun_out/r02_c2_probe_tetp2.jsonl; tail -3 gpurun_out/r02_c2_probe_tetp2.err
timeout 600 python tools/scale_probe.py trip2 1024 3162 > gpurun_out/r02_c2_probe_trip2.jsonl 2> gpurun_out/r02_c2_probe_trip2.err
cat gpurun_out/r02_c2_probe_trip2.jsonl; tail -3 gpurun_out/r02_c2_probe_trip2.err
