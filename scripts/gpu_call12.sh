#!/bin/bash
# round-2 GPU call 12: attn_vit with the elect.sync issue idiom: kernel tests, probe A/B against the previous build, encoder sweep A/B
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_attention_gpu.py -q -x > $O/r2_c12_kernels.log 2>&1; tail -2 $O/r2_c12_kernels.log
timeout 200 python scripts/attn_vit_probe.py > $O/r2_c12_probe.log 2>&1; grep -v "^$" $O/r2_c12_probe.log | tail -12
CGPT_LIB=$PWD/certifiedgpt_b200/lib/libcgpt_prev.so timeout 200 python scripts/attn_vit_probe.py > $O/r2_c12_probe_prev.log 2>&1; grep "head-major\|MMA thread" $O/r2_c12_probe_prev.log
timeout 600 python scripts/encoder_sweep.py 1024 > $O/r2_c12_sweep.log 2>&1; tail -2 $O/r2_c12_sweep.log
CGPT_LIB=$PWD/certifiedgpt_b200/lib/libcgpt_prev.so timeout 600 python scripts/encoder_sweep.py 1024 > $O/r2_c12_sweep_prev.log 2>&1; tail -2 $O/r2_c12_sweep_prev.log
