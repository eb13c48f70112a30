#!/bin/bash
# round-2 GPU call 10: online-softmax multi-tile attention: kernel tests, 448 px parity test, 448 px bench A/B against the two-pass build
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_attention_gpu.py -q -x > $O/r2_c10_kernels.log 2>&1; tail -2 $O/r2_c10_kernels.log
timeout 120 python scripts/attn_long_probe.py > $O/r2_c10_probe.log 2>&1; tail -6 $O/r2_c10_probe.log
CGPT_LIB=$PWD/certifiedgpt_b200/lib/libcgpt_prev.so timeout 120 python scripts/attn_long_probe.py 2>&1 | head -2 > $O/r2_c10_probe_prev.log; cat $O/r2_c10_probe_prev.log
timeout 600 python -m pytest tests/test_fullsize_gpu.py -q -x -k 448 > $O/r2_c10_448test.log 2>&1; tail -2 $O/r2_c10_448test.log
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-decode-sweep --img-size 448 --batch-size 275 > $O/r2_c10_bench448.log 2>&1; tail -c 700 $O/r2_c10_bench448.log
CGPT_LIB=$PWD/certifiedgpt_b200/lib/libcgpt_prev.so timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-decode-sweep --img-size 448 --batch-size 275 > $O/r2_c10_bench448_prev.log 2>&1; tail -c 700 $O/r2_c10_bench448_prev.log
