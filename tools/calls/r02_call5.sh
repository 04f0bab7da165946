#!/bin/bash
# Round 2, call 5 (2 GPUs): multi-GPU bench (weak-scaled configs, parity on hardware, exchange timing)
set -x
mkdir -p gpurun_out
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500 \
    bench.py --gpus 2 --steps 100 --warmup 3) > gpurun_out/r02_c5_bench2.json 2> gpurun_out/r02_c5_bench2.err
tail -c 5000 gpurun_out/r02_c5_bench2.json; tail -12 gpurun_out/r02_c5_bench2.err
run() {  # name kind size env...
  name=$1; kind=$2; size=$3; shift 3
  echo "## $name $*" >> gpurun_out/r02_c5_sweep.jsonl
  env "$@" timeout 300 python tools/scale_probe.py $kind $size >> gpurun_out/r02_c5_sweep.jsonl 2>> gpurun_out/r02_c5_sweep.err
}
(CUDA_VISIBLE_DEVICES=0 bash -c "$(declare -f run); run fan_r11 trip1 11 X=1; run fan_nocompute trip1 11 SPFEM_FAN_NOCOMPUTE=1; run fan_minb9 trip1 11 SPFEM_FAN_MINB8=9; run fan_r12 trip1 12 X=1; run fan_r12_nocompute trip1 12 SPFEM_FAN_NOCOMPUTE=1") &
(CUDA_VISIBLE_DEVICES=1 bash -c "$(declare -f run); run vec_full vecp2 3162 X=1; run vec_full_nobudget vecp2 3162 SPFEM_TILE_BUDGET=0; run vec_2048 vecp2 2048 X=1; run tet_100 tetp2 100 X=1; run tet_128 tetp2 128 X=1; run trip2_full trip2 3162 X=1") &
wait
cut -c1-330 gpurun_out/r02_c5_sweep.jsonl
