#!/bin/bash
# vertex-fan kernel at refine(13) (134 M triangles): L2 prefetch distance and rows per tile
mkdir -p gpurun_out
timeout 500 python tools/fan_variants.py 13 tools/calls/fan_variants_refine13.txt 3 > gpurun_out/r02_c52_fan_refine13.jsonl 2> gpurun_out/r02_c52.err
python - <<PY
import json
for l in open('gpurun_out/r02_c52_fan_refine13.jsonl'):
    d=json.loads(l); print(d['variant'], d.get('lap_ms'), d.get('lap_frac'), d.get('mass_ms'), d.get('mass_frac'), d.get('error','')[:100])
PY
tail -n 3 gpurun_out/r02_c52.err
