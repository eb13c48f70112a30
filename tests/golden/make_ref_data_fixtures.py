"""Generates tests/golden/ref_data.json by running the REFERENCE's own VQAv2 loader and BLIP-2 processors
(datasets/datasets/vqav2_dataset.py, datasets/datasets/base_dataset.py, processors/blip_processors.py,
processors/base_processor.py) on the synthetic VQAv2-shaped directory of tests/ref_data_util.py.

Runs only in the build container (needs /root/reference, read-only).  The four files are executed unmodified, by path,
behind shims for what this image lacks: `omegaconf` (only OmegaConf.create is touched), `common.registry` (decorators and
the logger lookup), `torch_xla.core.xla_model` (master_print).  The reference draws answers and instruction templates from
the global `random`; it is seeded so that certifiedgpt_b200.data.vqav2 (random.Random(seed)) must reproduce every item.

    python tests/golden/make_ref_data_fixtures.py
"""
import importlib.util
import json
import logging
import os
import random
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import ref_data_util as U  # noqa: E402

REF = "/root/reference"


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, path))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    oc = types.ModuleType("omegaconf")
    oc.OmegaConf = types.SimpleNamespace(create=lambda *a, **k: dict(*a, **k))
    common = types.ModuleType("common")
    creg = types.ModuleType("common.registry")

    class _Registry:
        def get_configuration_class(self, name):
            return logging.getLogger("ref")

        def __getattr__(self, k):              # register_processor(...) and friends: pass-through decorators
            return lambda *a, **kw: (lambda f: f)
    creg.registry = _Registry()
    xla = types.ModuleType("torch_xla")
    xcore = types.ModuleType("torch_xla.core")
    xm = types.ModuleType("torch_xla.core.xla_model")
    xm.master_print = lambda *a, **k: None
    xla.core, xcore.xla_model = xcore, xm
    sys.modules.update({"omegaconf": oc, "common": common, "common.registry": creg, "torch_xla": xla,
                        "torch_xla.core": xcore, "torch_xla.core.xla_model": xm})
    for pkg in ("processors", "datasets", "datasets.datasets"):
        sys.modules[pkg] = types.ModuleType(pkg)
    _load("processors.base_processor", "processors/base_processor.py")
    procs = _load("processors.blip_processors", "processors/blip_processors.py")
    _load("datasets.datasets.base_dataset", "datasets/datasets/base_dataset.py")
    ds = _load("datasets.datasets.vqav2_dataset", "datasets/datasets/vqav2_dataset.py")
    return procs, ds


def main():
    procs, dsm = load_reference()
    out = {"reference": "datasets/datasets/vqav2_dataset.py + base_dataset.py, processors/blip_processors.py + base_processor.py "
                        "(executed unmodified; shims: see this script's docstring)"}
    text = procs.BlipCaptionProcessor()                     # the reference's default max_words = 50
    out["pre_caption"] = [{"in": c, "out": text(c)} for c in U.CAPTIONS]
    out["pre_caption_max3"] = [{"in": c, "out": procs.BlipCaptionProcessor(max_words=3)(c)} for c in U.CAPTIONS]
    with tempfile.TemporaryDirectory() as root:
        qp, ap, img_dir = U.build(root)
        out["image_sizes"] = {}
        for size in (56, 224):
            vis = procs.Blip2ImageTrainProcessor(image_size=size)
            ds = dsm.VQAv2Dataset(vis, text, [qp], img_dir, [ap], split="train")
            rec = {"len": len(ds), "items": []}
            for seed in (1, 2):
                random.seed(seed)
                for i in range(len(ds)):
                    it = ds[i]
                    rec["items"].append({"seed": seed, "index": i, "question_id": it["question_id"],
                                         "instruction_input": it["instruction_input"], "answer": it["answer"],
                                         "image": U.image_digest(it["image"]), "shape": list(it["image"].shape)})
            out["image_sizes"][str(size)] = rec
        # evaluation split: VQAv2TestDataset (questions only, COCO_test2015 names, raw question in the instruction)
        vis = procs.Blip2ImageTrainProcessor(image_size=56)
        tds = dsm.VQAv2TestDataset([qp], vis, os.path.join(root, "test2015"), "test")
        out["test_dataset"] = {"len": len(tds), "items": [
            {"index": i, "question": it["question"], "question_id": it["question_id"], "img_id": it["img_id"],
             "image": U.image_digest(it["image"])} for i, it in ((i, tds[i]) for i in range(3))]}
        # answer sampling frequencies of annotation 0 (confidence weights: red 14/16, dark red 2/16, blue 0)
        vis = procs.Blip2ImageTrainProcessor(image_size=56)
        ds = dsm.VQAv2Dataset(vis, text, [qp], img_dir, [ap], split="train")
        random.seed(3)
        draws = [ds.get_data(0)["answer"] for _ in range(400)]
        out["answer_draws_seed3"] = draws
    with open(os.path.join(HERE, "ref_data.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("items:", [(r["seed"], r["index"], r["answer"], r["instruction_input"][:48]) for r in out["image_sizes"]["56"]["items"]])
    print({a: draws.count(a) for a in set(draws)})


if __name__ == "__main__":
    main()
