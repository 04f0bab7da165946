#!/bin/bash
# full captures of the element-tile kernel on the final code: TriP2 (slots in CSR order: row records)
# and vector-P2 (count-ordered block slots)
mkdir -p gpurun_out
prof () { name=$1; kind=$2; size=$3; pat=$4; n=$5
  python tools/scale_probe.py $kind $size > gpurun_out/r02_c53_plain_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$pat -s 3 -c $n -f -o gpurun_out/r02_final3_$name \
      python tools/scale_probe.py $kind $size > gpurun_out/r02_c53_ncu_$name.log 2>&1
  cut -c1-200 gpurun_out/r02_c53_plain_$name.log; tail -n 1 gpurun_out/r02_c53_ncu_$name.log; }
prof trip2_tile trip2 2048 spfem_elem_tile 1
prof vecp2_tile vecp2 2048 spfem_elem_tile 1
ls -la gpurun_out/r02_final3*.ncu-rep
