#!/bin/bash
# round-2 GPU call 4: kernel tests, whole suite, attention probe (stamps), per-shape encoder profile A/B, bench
O=gpurun_out
mkdir -p $O
timeout 420 python -m pytest tests/test_gemm_gpu.py tests/test_attention_gpu.py -q -x > $O/r2_c4_kernels.log 2>&1
K=$?
tail -4 $O/r2_c4_kernels.log
if [ $K -ne 0 ]; then echo "NEW KERNEL TESTS FAILED rc=$K -> falling back to the row-major ViT path for the rest"; export CGPT_VIT_ROW_MAJOR=1; fi
timeout 1200 python -m pytest tests -m gpu -q -s > $O/r2_c4_tests.log 2>&1
echo "suite rc=$?"; grep -E "passed|failed|full shape:" $O/r2_c4_tests.log | tail -5; grep -n "^FAILED\|^E  " $O/r2_c4_tests.log | head -30
python scripts/attn_vit_probe.py > $O/r2_c4_attn_probe.log 2>&1; tail -14 $O/r2_c4_attn_probe.log
python scripts/encoder_profile.py 1024 > $O/r2_c4_prof_new.log 2>&1; tail -9 $O/r2_c4_prof_new.log
CGPT_GEMM_RED_NO_RMW=1 python scripts/encoder_profile.py 1024 > $O/r2_c4_prof_normw.log 2>&1; tail -9 $O/r2_c4_prof_normw.log
CGPT_VIT_ROW_MAJOR=1 CGPT_GEMM_RED_NO_RMW=1 python scripts/encoder_profile.py 1024 > $O/r2_c4_prof_r1.log 2>&1; tail -9 $O/r2_c4_prof_r1.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/r2_c4_bench.log 2>&1; tail -c 2500 $O/r2_c4_bench.log
