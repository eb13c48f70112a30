#!/bin/bash
# round-2 GPU call 16: final tree on 1 GPU: whole suite, smoke, bench (with cpu baseline + decode sweep), reference arm, 448 px
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/r2_c16_tests.log 2>&1; grep -E "passed|failed" $O/r2_c16_tests.log | tail -2; grep -n "^FAILED\|^E  " $O/r2_c16_tests.log | head
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_c16_smoke.log 2>&1; tail -1 $O/r2_c16_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > $O/r2_c16_bench.log 2>&1; tail -c 600 $O/r2_c16_bench.log | head -c 300; echo
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_c16_bench_reference.log 2>&1; tail -c 500 $O/r2_c16_bench_reference.log; echo
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-decode-sweep --img-size 448 --batch-size 275 > $O/r2_c16_bench448.log 2>&1; python - <<PY
import json
d=json.loads([l for l in open("$O/r2_c16_bench448.log") if l.startswith("{")][-1]); print("448px", d["value"], d["e2e"]["value"], d["clocks"]["sm_mhz"])
d=json.loads([l for l in open("$O/r2_c16_bench.log") if l.startswith("{")][-1]); print("224px", d["value"], d["e2e"]["value"], d["clocks"]["sm_mhz"], d["roofline"]["frac"])
PY
