#!/bin/bash
# multi-GPU call (gpurun --gpus N): sharded-count identity, the weak-scaling bench line, the 64-image sigma=0.5 subset
# (BASELINE.json configs[2]) and the encoder-only sweep (configs[3]) at N ranks
N=${1:-2}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR scripts/dist_check.py > $O/r2_dist_check_${N}gpu.log 2>&1; tail -3 $O/r2_dist_check_${N}gpu.log
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-decode-sweep > $O/r2_bench_${N}gpu.log 2>&1; tail -c 1800 $O/r2_bench_${N}gpu.log
if [ "$N" = "8" ]; then
  timeout 900 $TR bench.py --gpus $N --sigma 0.5 --steps 8 --warmup 2 --no-cpu-baseline --no-decode-sweep > $O/r2_bench_${N}gpu_64images_sigma05.log 2>&1; tail -c 1200 $O/r2_bench_${N}gpu_64images_sigma05.log
fi
timeout 900 $TR scripts/encoder_sweep.py 64 256 1024 4096 > $O/r2_encoder_sweep_${N}gpu.log 2>&1; tail -5 $O/r2_encoder_sweep_${N}gpu.log
nvidia-smi --query-gpu=index,clocks.sm,power.draw --format=csv,noheader | head -8
