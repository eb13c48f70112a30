#!/bin/bash
# round-2 GPU call: new kernels in isolation first (under a timeout), then the whole suite, then A/B sweeps
O=gpurun_out
mkdir -p $O
timeout 420 python -m pytest tests/test_gemm_gpu.py tests/test_attention_gpu.py -q -x > $O/r2_c2_kernels.log 2>&1
K=$?
tail -25 $O/r2_c2_kernels.log
if [ $K -ne 0 ]; then echo "NEW KERNEL TESTS FAILED rc=$K -> falling back to the row-major ViT path for the rest"; export CGPT_VIT_ROW_MAJOR=1; fi
timeout 900 python -m pytest tests -m gpu -q > $O/r2_c2_tests.log 2>&1
echo "suite rc=$?"; tail -30 $O/r2_c2_tests.log
python scripts/encoder_sweep.py 1024 > $O/r2_c2_sweep_new.log 2>&1; tail -2 $O/r2_c2_sweep_new.log
CGPT_VIT_ROW_MAJOR=1 python scripts/encoder_sweep.py 1024 > $O/r2_c2_sweep_rowmajor.log 2>&1; tail -1 $O/r2_c2_sweep_rowmajor.log
CGPT_GEMM_RED_NO_RMW=1 python scripts/encoder_sweep.py 1024 > $O/r2_c2_sweep_normw.log 2>&1; tail -1 $O/r2_c2_sweep_normw.log
CGPT_VIT_ROW_MAJOR=1 CGPT_GEMM_RED_NO_RMW=1 python scripts/encoder_sweep.py 1024 > $O/r2_c2_sweep_r1.log 2>&1; tail -1 $O/r2_c2_sweep_r1.log
CGPT_GEMM_RMW_ALL=1 python scripts/encoder_sweep.py 1024 > $O/r2_c2_sweep_rmwall.log 2>&1; tail -1 $O/r2_c2_sweep_rmwall.log
CGPT_ATTN_PV_N128=1 python scripts/encoder_sweep.py 1024 > $O/r2_c2_sweep_pv128.log 2>&1; tail -1 $O/r2_c2_sweep_pv128.log
python scripts/attn_vit_probe.py > $O/r2_c2_attn_probe.log 2>&1; tail -12 $O/r2_c2_attn_probe.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/r2_c2_bench.log 2>&1; tail -c 1500 $O/r2_c2_bench.log
