"""Generates tests/golden/ref_lr.json by stepping the REFERENCE's own LinearWarmupCosineLRScheduler
(graphs/models/minigpt4/common/optims.py, executed unmodified by path; shim: common.registry's decorator) over every
(epoch, step) of three settings, incl. a warm-up longer than an epoch (the reference warms up on the step inside the
epoch, not on the global step) and the warmup_start_lr = -1 default.  Build container only (needs /root/reference).

    python tests/golden/make_ref_lr_fixture.py
"""
import importlib.util
import json
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
common = types.ModuleType("common")
creg = types.ModuleType("common.registry")
creg.registry = types.SimpleNamespace(register_lr_scheduler=lambda name: (lambda cls: cls))
sys.modules.update({"common": common, "common.registry": creg})
spec = importlib.util.spec_from_file_location("ref_optims", "/root/reference/graphs/models/minigpt4/common/optims.py")
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)

SETTINGS = [
    dict(max_epoch=4, iters_per_epoch=53, min_lr=1e-6, init_lr=1e-5, warmup_steps=53, warmup_start_lr=1e-6, warmup_max_lr=1e-5),
    dict(max_epoch=3, iters_per_epoch=10, min_lr=1e-6, init_lr=1e-5, warmup_steps=25, warmup_start_lr=1e-6, warmup_max_lr=1e-5),
    dict(max_epoch=2, iters_per_epoch=7, min_lr=0.0, init_lr=3e-5, warmup_steps=3, warmup_max_lr=3e-5),
]
out = []
for kw in SETTINGS:
    opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=123.0)
    sch = mod.LinearWarmupCosineLRScheduler(opt, **kw)
    lrs = []
    for epoch in range(kw["max_epoch"]):
        for step in range(kw["iters_per_epoch"]):
            lr = sch.step(cur_epoch=epoch, cur_step=step)
            assert opt.param_groups[0]["lr"] == lr
            lrs.append(lr)
    out.append({"settings": kw, "lr": lrs})
json.dump(out, open(os.path.join(HERE, "ref_lr.json"), "w"), indent=1)
print([len(o["lr"]) for o in out], out[1]["lr"][8:13])
