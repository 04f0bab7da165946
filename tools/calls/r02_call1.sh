#!/bin/bash
# Round 2, first GPU call (2 GPUs): NCCL interface exchange on hardware + scale probes
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_zz_nccl_exchange.py tests/test_zz_chunked_gpu.py \
    -m gpu -q > gpurun_out/r02_nccl_tests.log 2>&1
tail -5 gpurun_out/r02_nccl_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port 29611 tools/bench_exchange.py --refine 11 \
    > gpurun_out/r02_exchange_2gpu.json 2> gpurun_out/r02_exchange_2gpu.err
tail -2 gpurun_out/r02_exchange_2gpu.json
(CUDA_VISIBLE_DEVICES=0 timeout 900 python tools/scale_probe.py vecp2 512 1024 2048 > gpurun_out/r02_probe_vecp2.jsonl 2> gpurun_out/r02_probe_vecp2.err) &
(CUDA_VISIBLE_DEVICES=1 timeout 900 python tools/scale_probe.py tetp2 32 64 100 > gpurun_out/r02_probe_tetp2.jsonl 2> gpurun_out/r02_probe_tetp2.err;
 CUDA_VISIBLE_DEVICES=1 timeout 600 python tools/scale_probe.py trip1 11 12 13 > gpurun_out/r02_probe_trip1.jsonl 2> gpurun_out/r02_probe_trip1.err) &
wait
cat gpurun_out/r02_probe_*.jsonl
free -g | head -2; nproc
