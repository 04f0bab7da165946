#!/bin/bash
# tools/sweep_env2.sh "VAR1=a VAR2=b" "cfgs" -- one fused-kernel timing under a set of env assignments
assign="$1"; cfg="$2"
env $assign timeout 400 python tools/bench_configs.py --modes fused --only "$cfg" 2>&1 | grep '"config"' | \
  python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('$assign', d['config'], d.get('fused_ms'), d.get('fused_frac_of_hbm_peak'))"
