#!/bin/bash
# native fan plan builder on the device: parity + set-up time + bench
mkdir -p gpurun_out
timeout 300 python tools/scale_probe.py trip1 8 > gpurun_out/r02_c16_first.log 2>&1 || { tail -n 20 gpurun_out/r02_c16_first.log; exit 1; }
cut -c1-300 gpurun_out/r02_c16_first.log
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c16_gpu_tests.log 2>&1
tail -n 4 gpurun_out/r02_c16_gpu_tests.log
rm -f gpurun_out/r02_c16_sweep.jsonl
for s in 11 11 12 13; do timeout 600 python tools/scale_probe.py trip1 $s >> gpurun_out/r02_c16_sweep.jsonl 2>> gpurun_out/r02_c16_sweep.err; done
cut -c1-420 gpurun_out/r02_c16_sweep.jsonl; tail -n 3 gpurun_out/r02_c16_sweep.err
python bench.py --configs none > gpurun_out/r02_c16_bench.json 2> gpurun_out/r02_c16_bench.err
tail -n 3 gpurun_out/r02_c16_bench.err; cut -c1-1400 gpurun_out/r02_c16_bench.json
python -c "
import __graft_entry__ as g, time
t0=time.time(); g.smoke(); print('smoke s', time.time()-t0)" > gpurun_out/r02_c16_smoke.log 2>&1; tail -n 3 gpurun_out/r02_c16_smoke.log
