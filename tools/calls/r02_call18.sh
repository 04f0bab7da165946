#!/bin/bash
# ncu evidence of the round's final kernels: launch list of the bench command, full captures of
# the vertex-fan kernel (config 2) and the element-tile kernel (vector-P2, TetP2, TriP2)
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --configs none --no-cpu-baseline > gpurun_out/r02_c18_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_launches.csv \
    python bench.py --steps 2 --warmup 1 --configs none --no-cpu-baseline > gpurun_out/r02_c18_ncu_bench.log 2>&1
prof () { name=$1; kind=$2; size=$3; pat=$4; n=$5
  python tools/scale_probe.py $kind $size > gpurun_out/r02_c18_plain_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$pat -s 3 -c $n -f -o gpurun_out/r02_final_$name \
      python tools/scale_probe.py $kind $size > gpurun_out/r02_c18_ncu_$name.log 2>&1
  cut -c1-200 gpurun_out/r02_c18_plain_$name.log; tail -n 2 gpurun_out/r02_c18_ncu_$name.log; }
prof fan trip1 11 spfem_elem_fan 2
prof vecp2_tile vecp2 2048 spfem_elem_tile 1
prof tetp2_tile tetp2 128 spfem_elem_tile 1
prof trip2_tile trip2 2048 spfem_elem_tile 1
ls -la gpurun_out/*.ncu-rep
