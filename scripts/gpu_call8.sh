#!/bin/bash
# round-2 GPU call 8: attn_long with independent max / sum chains + cycle stamps
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_attention_gpu.py -q -x > $O/r2_c8_kernels.log 2>&1
K=$?
tail -3 $O/r2_c8_kernels.log
timeout 120 python scripts/attn_long_probe.py > $O/r2_c8_probe_new.log 2>&1; tail -5 $O/r2_c8_probe_new.log
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv
