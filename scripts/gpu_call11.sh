#!/bin/bash
# round-2 GPU call 11: whole suite, final 1-GPU bench line and the reference arm
O=gpurun_out
mkdir -p $O
# (compute-sanitizer is closed on this pool: the memcheck pass planned here could not run)
timeout 1500 python -m pytest tests -m gpu -q > $O/r2_c11_tests.log 2>&1
echo "suite rc=$?"; grep -E "passed|failed" $O/r2_c11_tests.log | tail -3; grep -n "^FAILED\|^E  " $O/r2_c11_tests.log | head -20
timeout 900 python bench.py --steps 5 --warmup 3 > $O/r2_c11_bench.log 2>&1; tail -c 2500 $O/r2_c11_bench.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2_c11_smoke.log 2>&1; tail -2 $O/r2_c11_smoke.log
