#!/bin/bash
# fan-kernel diagnostics: where does the time of the vertex-fan kernel go?
mkdir -p gpurun_out
timeout 900 python tools/fan_variants.py 11 tools/fan_variants_a.txt > gpurun_out/r02_c6_fan_variants.jsonl 2> gpurun_out/r02_c6_fan_variants.err
tail -5 gpurun_out/r02_c6_fan_variants.err
cut -c1-420 gpurun_out/r02_c6_fan_variants.jsonl
