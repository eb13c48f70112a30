#!/bin/bash
# round-2 GPU call 7: pair-mode multi-tile attention: kernel tests, same-box A/B against the previous build, 448 px bench
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_attention_gpu.py -q -x > $O/r2_c7_kernels.log 2>&1
K=$?
tail -6 $O/r2_c7_kernels.log
if [ $K -ne 0 ]; then grep -n "^E  \|Error\|FAILED" $O/r2_c7_kernels.log | head -20; fi
timeout 120 python scripts/attn_long_probe.py > $O/r2_c7_probe_new.log 2>&1; tail -3 $O/r2_c7_probe_new.log
CGPT_LIB=$PWD/certifiedgpt_b200/lib/libcgpt_prev.so timeout 120 python scripts/attn_long_probe.py > $O/r2_c7_probe_prev.log 2>&1; tail -3 $O/r2_c7_probe_prev.log
if [ $K -ne 0 ]; then exit 1; fi
timeout 600 python -m pytest tests/test_fullsize_gpu.py -q -x -k 448 > $O/r2_c7_448test.log 2>&1; tail -3 $O/r2_c7_448test.log
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-decode-sweep --img-size 448 --batch-size 275 > $O/r2_c7_bench448.log 2>&1; tail -c 1200 $O/r2_c7_bench448.log
CGPT_LIB=$PWD/certifiedgpt_b200/lib/libcgpt_prev.so timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-decode-sweep --img-size 448 --batch-size 275 > $O/r2_c7_bench448_prev.log 2>&1; tail -c 1200 $O/r2_c7_bench448_prev.log
