#!/bin/bash
# final multi-GPU call (gpurun --gpus N): N = 8: sharded-count identity, the weak-scaling bench line and the 64-image sigma = 0.5
# subset (BASELINE.json configs[2]); N = 4: the encoder-only sweep (configs[3]) the earlier calls lacked
N=${1:-8}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "$N" = "4" ]; then
  timeout 900 $TR scripts/encoder_sweep.py 64 256 1024 4096 > $O/r2f_encoder_sweep_${N}gpu.log 2>&1; tail -5 $O/r2f_encoder_sweep_${N}gpu.log
else
  timeout 600 $TR scripts/dist_check.py > $O/r2f_dist_check_${N}gpu.log 2>&1; tail -3 $O/r2f_dist_check_${N}gpu.log
  timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-decode-sweep > $O/r2f_bench_${N}gpu.log 2>&1; tail -c 900 $O/r2f_bench_${N}gpu.log | head -c 500; echo
  timeout 900 $TR bench.py --gpus $N --sigma 0.5 --steps 8 --warmup 2 --no-cpu-baseline --no-decode-sweep > $O/r2f_bench_${N}gpu_64images_sigma05.log 2>&1
  python - <<PY
import json
for n in ("r2f_bench_${N}gpu.log","r2f_bench_${N}gpu_64images_sigma05.log"):
    d=json.loads([l for l in open("$O/"+n) if l.startswith("{")][-1]); print(n, d["value"], d["e2e"]["value"], d.get("certified_images_per_min"), d["clocks"]["sm_mhz"], d["roofline"]["frac"])
PY
fi
