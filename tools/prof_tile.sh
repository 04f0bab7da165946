#!/bin/bash
# One ncu --set full capture of the element-tile kernel for one configuration.
#   tools/prof_tile.sh "vecP2" out_name   (knobs through the SPFEM_TILE_* environment)
cfg="$1"; out="$2"
cmd="python tools/bench_configs.py --force --modes fused --only $cfg"
$cmd > gpurun_out/plain_$out.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:_tile -s 10 -c 1 \
    -o gpurun_out/$out $cmd > gpurun_out/ncu_$out.log 2>&1
