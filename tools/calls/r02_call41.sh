#!/bin/bash
mkdir -p gpurun_out
python tools/scale_probe.py q2 10 > gpurun_out/r02_c41_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:spfem_elem_tile -s 3 -c 1 -f -o gpurun_out/r02_c41_q2_tile python tools/scale_probe.py q2 10 > gpurun_out/r02_c41_ncu.log 2>&1
cut -c1-300 gpurun_out/r02_c41_plain.log; tail -n 2 gpurun_out/r02_c41_ncu.log
