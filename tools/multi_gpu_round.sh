#!/bin/bash
# Round evidence on N GPUs of one box:  gpurun --gpus N -- 'bash tools/multi_gpu_round.sh N r2'
n=$1; tag=${2:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${tag}_smi_${n}gpu.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 3 \
    > gpurun_out/${tag}_bench_${n}gpu.json 2> gpurun_out/${tag}_bench_${n}gpu.err
tail -c 600 gpurun_out/${tag}_bench_${n}gpu.json | head -c 600; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 tools/ddp_step.py 2> gpurun_out/${tag}_ddp_${n}gpu.err | tail -1 | tee gpurun_out/${tag}_ddp_${n}gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 tools/ddp_step.py --unpatched 2>> gpurun_out/${tag}_ddp_${n}gpu.err | tail -1 | tee gpurun_out/${tag}_ddp_unpatched_${n}gpu.json
