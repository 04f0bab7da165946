#!/bin/bash
# tools/sweep_env.sh VAR "v1 v2 ..." "cfg1,cfg2"  -- fused tile kernel time per value of one env knob
var="$1"; vals="$2"; cfg="$3"
for v in $vals; do
  env $var=$v timeout 400 python tools/bench_configs.py --modes fused --only "$cfg" 2>&1 | grep '"config"' | \
    python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('$var=$v', d['config'], d.get('fused_ms'), d.get('fused_frac_of_hbm_peak'))"
done
