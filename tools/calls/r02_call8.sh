#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/fan_variants.py 11 tools/fan_variants_d.txt 7 > gpurun_out/r02_c9_fan_variants.jsonl 2> gpurun_out/r02_c9_fan_variants.err
tail -5 gpurun_out/r02_c9_fan_variants.err
cut -c1-400 gpurun_out/r02_c9_fan_variants.jsonl
