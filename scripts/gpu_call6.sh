#!/bin/bash
# round-2 GPU call 6: kernel tests (3-buffer multi-tile attention), whole suite, attention probes, 448 px and 224 px benches
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_attention_gpu.py -q -x > $O/r2_c6_kernels.log 2>&1
K=$?
tail -6 $O/r2_c6_kernels.log
if [ $K -ne 0 ]; then echo "NEW KERNEL TESTS FAILED rc=$K -> falling back to the row-major ViT path for the rest"; export CGPT_VIT_ROW_MAJOR=1; fi
timeout 1500 python -m pytest tests -m gpu -q -s > $O/r2_c6_tests.log 2>&1
echo "suite rc=$?"; grep -E "passed|failed|full shape:" $O/r2_c6_tests.log | tail -5; grep -n "^FAILED\|^E  " $O/r2_c6_tests.log | head -30
python scripts/attn_vit_probe.py > $O/r2_c6_attn_probe.log 2>&1; tail -12 $O/r2_c6_attn_probe.log
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-decode-sweep --img-size 448 --batch-size 275 > $O/r2_c6_bench448.log 2>&1; tail -c 1500 $O/r2_c6_bench448.log
python scripts/encoder_profile.py 1024 > $O/r2_c6_prof.log 2>&1; tail -9 $O/r2_c6_prof.log
timeout 900 python bench.py --steps 5 --warmup 3 > $O/r2_c6_bench.log 2>&1; tail -c 3000 $O/r2_c6_bench.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_c6_bench_reference.log 2>&1; tail -c 800 $O/r2_c6_bench_reference.log
