"""Drop-in `Smooth` (reference: randomized_smoothing/smoothing.py:13-117, Cohen et al.).

Same constructor, methods, return types and ABSTAIN convention as the reference, but the
Monte-Carlo loop runs on the device through libcgpt.so (include/cgpt.h):

  reference (smoothing.py:91-98, per batch)            here
  ---------------------------------------------------  ----------------------------------------
  x.repeat + randn_like*sigma + add  (7 fp32 passes)   one fused Philox noise kernel (K1)
  base_classifier(batch).argmax(1)                     fused MiniGPT-4 engine, or any nn.Module
  predictions.cpu().numpy()  (sync per batch)          labels stay on the device
  _count_arr Python loop                               warp-aggregated histogram kernel
  scipy/statsmodels tail on the host                   device tail kernel over a per-(n, alpha) SciPy table
                                                       (bit-identical radius), ONE D2H read

Extras (keyword-only, all optional): `seed`, `noise_space`, `noise_kind`, `process_group`
(shards the N draws across ranks; one int64 all-reduce per _sample_noise) and
`inject_noise()` for parity tests on identical noise.

There is no CPU path: `x` must be a CUDA tensor and libcgpt.so must be built.
"""
from math import ceil

import numpy as np
import torch

from .. import _lib as L
from ..dist import allreduce_counts, rank_world, shard_range

_SPACE = {"normalized": L.SPACE_NORMALIZED, "pixel": L.SPACE_PIXEL}
_KIND = {"gaussian": L.NOISE_GAUSSIAN, "uniform": L.NOISE_UNIFORM}


class Smooth(object):
    """A smoothed classifier g"""

    # to abstain, Smooth returns this int (smoothing.py:17)
    ABSTAIN = -1

    def __init__(self, base_classifier, num_classes: int, sigma: float, *, seed: int = 0,
                 noise_space: str = "normalized", noise_kind: str = "gaussian",
                 mean=L.BLIP_MEAN, std=L.BLIP_STD, process_group=None, fuse_selection: bool = True,
                 exact_tail: bool = True):
        """
        :param base_classifier: maps [batch x channel x height x width] to [batch x num_classes]
               (any torch.nn.Module), or a fused engine exposing `noisy_labels(...)`
        :param num_classes:
        :param sigma: the noise level hyperparameter
        """
        self.base_classifier = base_classifier
        self.num_classes = num_classes
        self.sigma = sigma
        self.seed = int(seed)
        self.noise_space = _SPACE[noise_space]
        self.noise_kind = _KIND[noise_kind]
        self.mean, self.std = tuple(mean), tuple(std)
        self.process_group = process_group
        self.fuse_selection = fuse_selection
        # True: pABar and Phi^-1 come from a per-(n, alpha) table built once on the host by the SciPy calls the
        # reference makes per image (smoothing.py:55,117) -> (label, radius) bit-identical to the reference;
        # False: fp64 device bisection + AS241 (1e-12 relative), no SciPy anywhere
        self.exact_tail = exact_tail
        self.image_id = 0          # Philox stream id; bumped per certify/predict call
        self._cursor = 0           # global sample index inside the current call
        self._injected = None
        self.last_counts = None    # device counts of the last _sample_noise (diagnostics)
        self.last_invalid = None
        L.load()

    # ------------------------------------------------------------------ parity hook
    def inject_noise(self, eps):
        """Use these standard draws instead of Philox: tensor [n_total, C, H, W] (fp32, CUDA) or
        callable(first, count) -> tensor.  Consumed by global sample index; None restores Philox."""
        self._injected = eps

    # ------------------------------------------------------------------ reference API
    def certify(self, x: torch.tensor, n0: int, n: int, alpha: float, batch_size: int) -> (int, float):
        """Monte Carlo certification (smoothing.py:29-56).  Returns (class, radius) or (ABSTAIN, 0.0)."""
        self._eval()
        self._cursor = 0
        if self._native() and self.fuse_selection and not callable(self._injected):
            # whole call inside libcgpt (cgpt_certify): fused selection+estimation pass, device tail, one D2H read;
            # x may live on the host (the library stages it) or on the device
            label, radius, d = self.base_classifier.certify(
                self._as_f32(x), n0, n, alpha, batch_size, self.sigma, eps=self._injected, process_group=self.process_group,
                exact_tail=self.exact_tail, **self._noise_kw())
            self.last_cAHat, self.last_pABar = d["cAHat"], d["pABar"]
            self.last_counts_selection, self.last_counts_estimation = d["counts_selection"], d["counts_estimation"]
            self.last_counts = self.last_counts_selection
            self.image_id += 1
            return (Smooth.ABSTAIN, 0.0) if label == Smooth.ABSTAIN else (label, radius)
        if self.fuse_selection:
            # The n0 selection draws and the n estimation draws are independent (fresh noise, smoothing.py:44,48)
            # and keyed by global sample index, so both phases run as ONE sharded pass over [0, n0 + n) and are
            # histogrammed into two count vectors: bit-identical counts, no small selection batch, one all-reduce.
            counts_selection, counts_estimation = self._sample_noise_device(x, n0 + n, batch_size, split=n0)
        else:
            counts_selection = self._sample_noise_device(x, n0, batch_size)
            counts_estimation = self._sample_noise_device(x, n, batch_size)
        lut = L.radius_lut(n, alpha, counts_estimation.device) if self.exact_tail else None
        lab, st = L.certify_tail(counts_selection, counts_estimation, n, alpha, self.sigma, lut)
        label = int(lab[0].item())         # the one device->host read of the call
        radius = float(st[0].item())
        self.last_cAHat = int(lab[1].item())
        self.last_pABar = float(st[1].item())
        self.last_counts_selection = counts_selection
        self.last_counts_estimation = counts_estimation
        self.image_id += 1
        if label == Smooth.ABSTAIN:
            return Smooth.ABSTAIN, 0.0
        return label, radius

    def certify_batch(self, xs, n0: int, n: int, alpha: float, batch_size: int):
        """`certify` for a list of images, sharing every pass between them (extension; native engine only): with the
        draws of an image sharded over W GPUs a rank holds only (n0 + n) / W draws of it, so K images per pass keep the
        per-rank batch large (BASELINE.json configs[2]: 64 images, N = 1000, 8 GPUs).  Image k is drawn from Philox
        stream `image_id + k`, exactly as K successive `certify` calls would: the result list equals theirs bit for
        bit.  Returns [(class or ABSTAIN, radius), ...]."""
        if not self._native():
            return [self.certify(x, n0, n, alpha, batch_size) for x in xs]
        self._eval()
        res = self.base_classifier.certify_batch([self._as_f32(x) for x in xs], n0, n, alpha, batch_size, self.sigma,
                                                 process_group=self.process_group, exact_tail=self.exact_tail,
                                                 **self._noise_kw())
        self.image_id += len(xs)
        self.last_batch_detail = [d for _, _, d in res]
        return [((Smooth.ABSTAIN, 0.0) if lab == Smooth.ABSTAIN else (lab, rad)) for lab, rad, _ in res]

    def predict(self, x: torch.tensor, n: int, alpha: float, batch_size: int) -> int:
        """Monte Carlo prediction with the top-2 binomial test (smoothing.py:58-79)."""
        self._eval()
        self._cursor = 0
        if self._native() and not callable(self._injected):
            label, self.last_pvalue, self.last_counts = self.base_classifier.predict(
                self._as_f32(x), n, alpha, batch_size, self.sigma, eps=self._injected, process_group=self.process_group,
                **self._noise_kw())
            self.image_id += 1
            return label
        counts = self._sample_noise_device(x, n, batch_size)
        lab, st = L.predict_tail(counts, alpha)
        label = int(lab[0].item())
        self.last_pvalue = float(st[0].item())
        self.image_id += 1
        return label

    def _sample_noise(self, x: torch.tensor, num: int, batch_size) -> np.ndarray:
        """Per-class counts of the base classifier under noise (smoothing.py:81-99)."""
        return self._sample_noise_device(x, num, batch_size).cpu().numpy()

    def _count_arr(self, arr, length: int) -> np.ndarray:
        """Histogram of predictions (smoothing.py:101-105), on the device."""
        t = torch.as_tensor(np.asarray(arr), dtype=torch.int32).cuda()
        counts = torch.zeros(length, dtype=torch.int64, device=t.device)
        L.label_hist(t.contiguous(), counts)
        return counts.cpu().numpy()

    def _lower_confidence_bound(self, NA: int, N: int, alpha: float) -> float:
        """Clopper-Pearson (1 - alpha) lower bound (smoothing.py:107-117), fp64 on the device."""
        dev = torch.device("cuda", torch.cuda.current_device())
        sel = torch.ones(1, dtype=torch.int64, device=dev)
        est = torch.full((1,), int(NA), dtype=torch.int64, device=dev)
        _, st = L.certify_tail(sel, est, N, alpha, 1.0, L.radius_lut(N, alpha, dev) if self.exact_tail else None)
        return float(st[1].item())

    # ------------------------------------------------------------------ device loop
    def _native(self):
        return bool(getattr(self.base_classifier, "cgpt_native", False))

    @staticmethod
    def _as_f32(x):
        return x.detach().to(torch.float32).contiguous()

    def _noise_kw(self):
        return dict(seed=self.seed, stream_id=self.image_id, noise_space=self.noise_space,
                    noise_kind=self.noise_kind, mean=self.mean, std=self.std)

    def _eval(self):
        ev = getattr(self.base_classifier, "eval", None)
        if callable(ev):
            ev()  # smoothing.py:42,71

    def _eps_for(self, first, count):
        if self._injected is None:
            return None
        if callable(self._injected):
            e = self._injected(first, count)
        else:
            e = self._injected[first:first + count]
        assert e.shape[0] == count, "injected noise exhausted"
        return e.contiguous()

    def _sample_noise_device(self, x, num: int, batch_size, split=None):
        """Counts over the global sample range [cursor, cursor + num).  With `split`, samples
        [cursor, cursor + split) are counted into a first vector and the rest into a second one."""
        if self._native() and not callable(self._injected):
            # the whole Monte-Carlo loop (chunking, sharding, histogram, NCCL all-reduce) runs inside libcgpt
            counts2 = self.base_classifier.sample_noise(
                self._as_f32(x), num, batch_size, self.sigma, base=self._cursor, split=split, eps=self._injected,
                process_group=self.process_group, **self._noise_kw())
            self._cursor += num
            self.last_counts = counts2[0]
            self.last_invalid = self.base_classifier.last_invalid
            return counts2[0] if split is None else (counts2[0], counts2[1])
        if not (isinstance(x, torch.Tensor) and x.is_cuda):
            raise L.CgptError("Smooth: x must be a CUDA tensor (no CPU path exists)")
        x = x.detach().to(torch.float32).contiguous()
        dev = x.device
        rank, world = rank_world(self.process_group)
        # this rank's contiguous slice of the global sample range [cursor, cursor + num)
        base = self._cursor
        lo, hi = shard_range(base, num, rank, world)
        self._cursor += num
        nvec = 1 if split is None else 2
        counts2 = torch.zeros(nvec, self.num_classes, dtype=torch.int64, device=dev)
        counts = counts2[0]
        boundary = base + split if split is not None else None
        invalid = torch.zeros(1, dtype=torch.int32, device=dev)
        fused = getattr(self.base_classifier, "noisy_labels", None)
        with torch.no_grad():
            first = lo
            n_chunks = ceil((hi - lo) / batch_size) if hi > lo else 0
            for ci in range(n_chunks):
                # balanced chunks (each <= batch_size): no tiny tail batch
                this_batch_size = (hi - first + (n_chunks - ci) - 1) // (n_chunks - ci)
                eps = self._eps_for(first, this_batch_size)
                kw = dict(eps=eps, seed=self.seed, stream_id=self.image_id, first_sample=first,
                          noise_space=self.noise_space, noise_kind=self.noise_kind)
                if fused is not None:
                    labels = fused(x, this_batch_size, self.sigma, mean=self.mean, std=self.std, **kw)
                else:
                    mean = self.mean if self.noise_space == L.SPACE_PIXEL else None
                    std = self.std if self.noise_space == L.SPACE_PIXEL else None
                    batch = L.noise_image(x, this_batch_size, self.sigma, mean=mean, std=std, **kw)
                    logits = self.base_classifier(batch)
                    labels = L.argmax_rows(logits.float().contiguous())
                if boundary is None or first + this_batch_size <= boundary:
                    L.label_hist(labels, counts2[0], invalid)
                elif first >= boundary:
                    L.label_hist(labels, counts2[1], invalid)
                else:
                    k = boundary - first
                    L.label_hist(labels[:k], counts2[0], invalid)
                    L.label_hist(labels[k:], counts2[1], invalid)
                first += this_batch_size
        if world > 1:
            allreduce_counts(counts2, self.process_group)   # the only collective on the path
        self.last_counts = counts2[0]
        self.last_invalid = invalid
        if split is None:
            return counts2[0]
        return counts2[0], counts2[1]
