#!/bin/bash
# Sweep of the element-tile kernel's tuning knobs (tile size, threads per
# element, staged / direct store) for one configuration of tools/bench_configs.py.
#   tools/sweep_tile.sh "vecP2" "48 64 96 128" "1 2 4" "0 1"
cfg="$1"; Es="$2"; NPs="$3"; Ds="$4"
for E in $Es; do for NP in $NPs; do for D in $Ds; do
  SPFEM_TILE_E=$E SPFEM_TILE_NPART=$NP SPFEM_TILE_DIRECT=$D timeout 300 \
    python tools/bench_configs.py --force --modes fused --only "$cfg" 2>&1 | grep '"config"' | \
    python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['config'], 'E=$E NP=$NP D=$D', d.get('fused_ms'), d.get('fused_frac_of_hbm_peak'), d.get('tile'))"
done; done; done
