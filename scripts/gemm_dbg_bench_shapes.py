"""CGPT_GEMM_DBG role counters (see gemm_dbg.py) at the bench's own shapes: batch 1100 ViT (M = 282 700) and Llama prefill (M = 79 200)."""
import runpy, sys, os
sys.argv = ["gemm_dbg.py"]
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "gemm_dbg.py")).read()
head = src[:src.index("run(65792")]
exec(compile(head, "gemm_dbg_head", "exec"))
run(282700, 6144, 1408, act=1)
run(282700, 4224, 1408)
run(282700, 1408, 1408, resid=True)
run(282700, 1408, 6144, resid=True)
run(79200, 22016, 4096, act=2)
run(79200, 12288, 4096)
run(79200, 4096, 11008, resid=True)
run(79200, 4096, 4096, resid=True)
