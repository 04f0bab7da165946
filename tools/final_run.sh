#!/bin/bash
# Round-end evidence: bench line, reference arm, ncu launch list and one full
# capture of the headline kernel (each ncu run after the same command exited 0).
set -x
python bench.py > gpurun_out/r01_final_bench.json 2> gpurun_out/r01_final_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01_final_reference.json 2>&1
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1"
$CMD > gpurun_out/plain_final.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv \
    --log-file gpurun_out/r01_final_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain_final2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fan8 -s 8 -c 2 \
    -o gpurun_out/r01_final_fan8 $CMD > gpurun_out/ncu_full.log 2>&1
tail -c 600 gpurun_out/r01_final_bench.json
