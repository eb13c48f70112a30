"""Certification agent (the reference's agents/minigpt4_certify_agent.py is empty).

Call site of `Smooth.certify`, shaped after MiniGPT4EvalAgent.eval
(agents/minigpt4_eval_agent.py:71-124) and Cohen et al.'s certify.py loop: iterate a dataset of
(image, label) items, certify each, and log `idx label predict radius correct time`.
"""
import datetime
import time

import torch

from ..randomized_smoothing.smoothing import Smooth


class MiniGPT4CertifyAgent:
    name = "image_text_certify"

    def __init__(self, base_classifier, dataset, num_classes, sigma, n0=100, n=1000, alpha=0.001,
                 batch_size=1000, skip=1, max_items=-1, outfile=None, smooth_kwargs=None):
        self.smooth = Smooth(base_classifier, num_classes, sigma, **(smooth_kwargs or {}))
        self.dataset = dataset
        self.n0, self.n, self.alpha, self.batch_size = n0, n, alpha, batch_size
        self.skip, self.max_items = skip, max_items
        self.outfile = outfile
        self.records = []

    @classmethod
    def setup_agent(cls, **kwargs):
        return cls(**kwargs)

    def run(self):
        f = open(self.outfile, "w") if self.outfile else None
        header = "idx\tlabel\tpredict\tradius\tcorrect\ttime"
        if f:
            print(header, file=f, flush=True)
        for i in range(len(self.dataset)):
            if i % self.skip != 0:
                continue
            if i == self.max_items:
                break
            item = self.dataset[i]
            x, label = item["image"], int(item["label"])
            # every item has its own question (vqav2_dataset.py:19-166): the label answers THAT question
            if item.get("suffix_ids") is not None:
                self.smooth.base_classifier.set_question(item["suffix_ids"])
            self.smooth.image_id = i           # Philox stream = dataset index: results do not depend on order
            before = time.time()
            prediction, radius = self.smooth.certify(x.cuda(non_blocking=True), self.n0, self.n, self.alpha,
                                                     self.batch_size)
            after = time.time()
            correct = int(prediction == label)
            rec = {"idx": i, "label": label, "predict": prediction, "radius": radius, "correct": correct,
                   "time": after - before}
            self.records.append(rec)
            if f:
                elapsed = str(datetime.timedelta(seconds=(after - before)))
                print(f"{i}\t{label}\t{prediction}\t{radius:.3}\t{correct}\t{elapsed}", file=f, flush=True)
        if f:
            f.close()
        return self.records

    def certified_accuracy(self, radii):
        """fraction of items predicted correctly with certified radius >= r (README.md:97-102 table)."""
        n = max(1, len(self.records))
        return {float(r): sum(1 for t in self.records if t["correct"] and t["radius"] >= r) / n for r in radii}

    def finalize(self):
        torch.cuda.synchronize()
