"""Smoothed-prediction agent (the reference's agents/minigpt4_predict_agent.py is empty).

Call site of `Smooth.predict` (launch.py mode `smoothing_predict`), after Cohen et al.'s predict.py:
logs `idx label predict correct time` per item.
"""
import datetime
import time

import torch

from ..randomized_smoothing.smoothing import Smooth


class MiniGPT4PredictAgent:
    name = "image_text_predict"

    def __init__(self, base_classifier, dataset, num_classes, sigma, n=1000, alpha=0.001, batch_size=1000,
                 skip=1, max_items=-1, outfile=None, smooth_kwargs=None):
        self.smooth = Smooth(base_classifier, num_classes, sigma, **(smooth_kwargs or {}))
        self.dataset = dataset
        self.n, self.alpha, self.batch_size = n, alpha, batch_size
        self.skip, self.max_items = skip, max_items
        self.outfile = outfile
        self.records = []

    @classmethod
    def setup_agent(cls, **kwargs):
        return cls(**kwargs)

    def run(self):
        f = open(self.outfile, "w") if self.outfile else None
        if f:
            print("idx\tlabel\tpredict\tcorrect\ttime", file=f, flush=True)
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
            self.smooth.image_id = i
            before = time.time()
            prediction = self.smooth.predict(x.cuda(non_blocking=True), self.n, self.alpha, self.batch_size)
            after = time.time()
            correct = int(prediction == label)
            self.records.append({"idx": i, "label": label, "predict": prediction, "correct": correct,
                                 "time": after - before})
            if f:
                elapsed = str(datetime.timedelta(seconds=(after - before)))
                print(f"{i}\t{label}\t{prediction}\t{correct}\t{elapsed}", file=f, flush=True)
        if f:
            f.close()
        return self.records

    def abstain_rate(self):
        n = max(1, len(self.records))
        return sum(1 for t in self.records if t["predict"] == Smooth.ABSTAIN) / n

    def finalize(self):
        torch.cuda.synchronize()
