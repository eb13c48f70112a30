"""ctypes binding of libcgpt's native engine (include/cgpt.h, "native engine").

`NativeMiniGPT4Engine` is the base classifier `Smooth` drives on the B200 path: the whole per-batch
pipeline (noise -> EVA ViT-g -> Q-Former -> llama_proj -> Llama prefill + greedy decode -> answer label),
the Monte-Carlo loop of `Smooth._sample_noise` (smoothing.py:81-99) and the certify / predict tails run
inside libcgpt.so (C++ host code + sm_100a kernels + CUDA-graph replay); Python only packs the weights
once, owns the device memory (one workspace tensor) and passes pointers.

`MiniGPT4Engine` (engine.py) is the same kernel sequence driven from Python; it stays as the
introspectable twin (`collect=` of intermediate activations for the parity tests) and the two must
produce bit-identical token ids (tests/test_native_gpu.py).
"""
import ctypes as C

import torch

from . import _lib as L
from .config import ModelConfig
from .engine import MiniGPT4Engine


class ModelConfigC(C.Structure):
    _fields_ = [
        ("img_size", C.c_int), ("vit_dim", C.c_int), ("vit_depth", C.c_int), ("vit_heads", C.c_int),
        ("vit_mlp", C.c_int), ("vit_eps", C.c_float), ("ln_vision_eps", C.c_float),
        ("qf_hidden", C.c_int), ("qf_layers", C.c_int), ("qf_heads", C.c_int), ("qf_inter", C.c_int),
        ("qf_queries", C.c_int), ("qf_cross_freq", C.c_int), ("qf_eps", C.c_float),
        ("llm_hidden", C.c_int), ("llm_layers", C.c_int), ("llm_heads", C.c_int), ("llm_inter", C.c_int),
        ("llm_vocab", C.c_int), ("llm_rms_eps", C.c_float), ("eos_id", C.c_int), ("pad_id", C.c_int),
        ("n_prefix", C.c_int), ("n_suffix", C.c_int), ("max_new_tokens", C.c_int), ("min_length", C.c_int),
        ("num_classes", C.c_int), ("early_exit", C.c_int), ("use_graphs", C.c_int),
    ]


class NoiseSpecC(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64), ("stream_id", C.c_uint32), ("sigma", C.c_float),
        ("mean", C.c_float * 3), ("std", C.c_float * 3),
        ("noise_space", C.c_int), ("noise_kind", C.c_int), ("eps", C.c_void_p),
    ]


_declared = False


def _declare(lib):
    global _declared
    if _declared:
        return
    vp, i64, i32, f64, u64 = C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_uint64
    pns = C.POINTER(NoiseSpecC)
    sigs = {
        "cgpt_create": [C.POINTER(ModelConfigC), C.POINTER(vp)],
        "cgpt_destroy": [vp],
        "cgpt_bind_weight": [vp, C.c_char_p, vp, i64, i64, i32],
        "cgpt_set_prompt": [vp, vp, vp],
        "cgpt_set_answer_table": [vp, vp, vp, i32],
        "cgpt_set_question": [vp, vp, i32],
        "cgpt_workspace_bytes": [vp, i32, i32, C.POINTER(i64)],
        "cgpt_bind_workspace": [vp, vp, i64, i32, i32, vp],
        "cgpt_vit_forward": [vp, vp, i32, vp, vp],
        "cgpt_qformer_forward": [vp, vp, i32, vp, vp, vp],
        "cgpt_llm_prefill_decode": [vp, vp, i32, vp, vp, C.POINTER(i32), vp],
        "cgpt_noisy_labels": [vp, vp, pns, u64, i32, vp, vp],
        "cgpt_lm_loss": [vp, vp, i32, vp, i32, vp, vp, vp],
        "cgpt_sample_noise": [vp, vp, pns, i64, i64, i32, i64, i32, i32, vp, vp, vp, vp],
        "cgpt_certify": [vp, vp, pns, i64, i64, f64, i32, i32, i32, vp, C.POINTER(i32), C.POINTER(f64),
                         C.POINTER(f64), vp],
        "cgpt_predict": [vp, vp, pns, i64, f64, i32, i32, i32, vp, C.POINTER(i32), C.POINTER(f64), vp],
        "cgpt_set_radius_lut": [vp, i64, f64, vp],
        "cgpt_certify_batch": [vp, vp, i32, pns, i64, i64, f64, i32, i32, i32, vp, vp, vp, vp, vp],
        "cgpt_last_counts": [vp, C.POINTER(vp)],
        "cgpt_last_decode_steps": [vp],
        "cgpt_set_option": [vp, C.c_char_p, i32],
        "cgpt_comm_unique_id": [vp],
        "cgpt_comm_init": [vp, i32, i32, C.POINTER(vp)],
        "cgpt_comm_destroy": [vp],
        "cgpt_allreduce_counts": [vp, i32, vp, vp],
        "cgpt_allreduce_f32": [vp, i64, vp, vp],
    }
    for name, args in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    _declared = True


def lib():
    h = L.load()
    _declare(h)
    return h


def config_struct(cfg: ModelConfig, n_prefix, n_suffix, max_new_tokens, min_length, num_classes, early_exit,
                  use_graphs):
    v, q, l = cfg.vit, cfg.qf, cfg.llm
    assert v.patch == 14, "the patchify kernel is written for 14x14 patches (eva_vit.py:425-437)"
    c = ModelConfigC()
    c.img_size, c.vit_dim, c.vit_depth, c.vit_heads, c.vit_mlp = v.img_size, v.dim, v.depth, v.heads, v.mlp
    c.vit_eps, c.ln_vision_eps = v.eps, cfg.ln_vision_eps
    c.qf_hidden, c.qf_layers, c.qf_heads, c.qf_inter = q.hidden, q.layers, q.heads, q.inter
    c.qf_queries, c.qf_cross_freq, c.qf_eps = q.n_query, q.cross_freq, q.eps
    c.llm_hidden, c.llm_layers, c.llm_heads, c.llm_inter, c.llm_vocab = l.hidden, l.layers, l.heads, l.inter, l.vocab
    c.llm_rms_eps, c.eos_id, c.pad_id = l.rms_eps, l.eos_id, l.pad_id
    c.n_prefix, c.n_suffix = n_prefix, n_suffix
    c.max_new_tokens, c.min_length = max_new_tokens, min_length
    c.num_classes, c.early_exit, c.use_graphs = num_classes, int(bool(early_exit)), int(bool(use_graphs))
    return c


# GEMM timing hook (lives in libcgpt, serves both engines)
gemm_profile_begin = L.gemm_profile_start
gemm_profile_end = L.gemm_profile_stop


# ------------------------------------------------------------------------------- NCCL communicator
class CountsComm:
    """NCCL communicator for the one collective on the path (int64 label counts), created inside libcgpt
    from an ncclUniqueId that rank 0 draws and `torch.distributed` broadcasts."""

    def __init__(self, process_group=True):
        import torch.distributed as dist
        g = None if process_group is True else process_group
        self.rank, self.world = dist.get_rank(g), dist.get_world_size(g)
        h = lib()
        buf = (C.c_char * 128)()
        if self.rank == 0:
            L.check(h.cgpt_comm_unique_id(C.cast(buf, C.c_void_p)))
        box = [bytes(buf.raw)]
        src = 0 if g is None else dist.get_global_rank(g, 0)
        dist.broadcast_object_list(box, src=src, group=g)
        ident = (C.c_char * 128).from_buffer_copy(box[0])
        self.handle = C.c_void_p()
        L.check(h.cgpt_comm_init(C.cast(ident, C.c_void_p), self.rank, self.world, C.byref(self.handle)))

    def allreduce(self, counts):
        assert counts.dtype == torch.int64 and counts.is_cuda and counts.is_contiguous()
        L.check(lib().cgpt_allreduce_counts(L.ptr(counts), counts.numel(), self.handle, L.stream_ptr()))
        return counts

    def allreduce_f32(self, values):
        assert values.dtype == torch.float32 and values.is_cuda and values.is_contiguous()
        L.check(lib().cgpt_allreduce_f32(L.ptr(values), values.numel(), self.handle, L.stream_ptr()))
        return values

    def close(self):
        if self.handle:
            lib().cgpt_comm_destroy(self.handle)
            self.handle = C.c_void_p()


# ------------------------------------------------------------------------------- engine
class NativeMiniGPT4Engine:
    """MiniGPT-4 noisy-sample classifier behind a libcgpt handle.  Same constructor as `MiniGPT4Engine`."""
    cgpt_fused = True
    cgpt_native = True

    def __init__(self, cfg: ModelConfig, state_dict, prefix_ids, suffix_ids, answer_table, num_classes, *,
                 max_new_tokens=20, min_length=1, device="cuda", early_exit=True, use_graphs=True,
                 max_batch=0, _packed_from=None):
        self._lib = lib()
        self.cfg = cfg
        self.dev = torch.device(device)
        self.num_classes = int(num_classes)
        self.max_new_tokens = int(max_new_tokens)
        self.min_length = int(min_length)
        if _packed_from is not None:
            src = _packed_from
        else:
            # weight packing (fused qkv, interleaved gate/up, padded patch weight, rope tables) is shared with
            # the Python twin; it allocates nothing but the packed weights and the 7-token prefix K/V
            src = MiniGPT4Engine(cfg, state_dict, prefix_ids, suffix_ids, answer_table, num_classes,
                                 max_new_tokens=max_new_tokens, min_length=min_length, device=device,
                                 early_exit=early_exit, use_graphs=False)
        self._src = src      # keeps the packed tensors alive (libcgpt holds raw pointers)
        self.w = src.w
        self.P, self.Tp, self.S = src.P, src.Tp, src.S
        self.table_keys, self.table_vals = src.table_keys, src.table_vals
        c = config_struct(cfg, src.P, len(src.suffix_ids), self.max_new_tokens, self.min_length, self.num_classes,
                          early_exit, use_graphs)
        self._h = C.c_void_p()
        L.check(self._lib.cgpt_create(C.byref(c), C.byref(self._h)))
        for name, t in self.w.items():
            rows, cols = (1, t.numel()) if t.dim() == 1 else (t.shape[0], t.shape[1])
            assert t.is_contiguous() and t.dtype in (torch.bfloat16, torch.float32)
            L.check(self._lib.cgpt_bind_weight(self._h, name.encode(), L.ptr(t), rows, cols,
                                               L.DT_F32 if t.dtype == torch.float32 else L.DT_BF16))
        L.check(self._lib.cgpt_set_prompt(self._h, L.ptr(src.prefix_ids_dev) if src.P else None,
                                          L.ptr(src.suffix_ids_dev) if len(src.suffix_ids) else None))
        L.check(self._lib.cgpt_set_answer_table(self._h, L.ptr(self.table_keys), L.ptr(self.table_vals),
                                                self.table_keys.numel()))
        self._ws = None
        self._ws_B = 0
        self._ws_encoder_only = False
        self._images = 1
        self._comm = None
        self.last_steps = 0
        if max_batch:
            self.reserve(max_batch)

    @classmethod
    def from_engine(cls, eng: MiniGPT4Engine, *, early_exit=None, use_graphs=True, max_batch=0, max_new_tokens=None):
        """Native engine over the packed weights of an existing Python engine (no second copy).  max_new_tokens may be
        lowered below the source engine's (its rotary tables cover every shorter budget)."""
        assert max_new_tokens is None or 1 <= max_new_tokens <= eng.max_new_tokens
        return cls(eng.cfg, None, eng.prefix_ids, eng.suffix_ids, None, eng.num_classes,
                   max_new_tokens=max_new_tokens or eng.max_new_tokens, min_length=eng.min_length, device=eng.dev,
                   early_exit=eng.early_exit if early_exit is None else early_exit, use_graphs=use_graphs,
                   max_batch=max_batch, _packed_from=eng)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.cgpt_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def eval(self):
        return self

    def set_question(self, suffix_ids):
        """The question of the next calls (token ids after the image, at most as many as the constructor's): one device
        buffer rewritten in place, so the graphs captured for a question length are replayed for every question of
        that length (cgpt_set_question)."""
        ids = [int(i) for i in suffix_ids]
        src = self._src
        if getattr(self, "_suffix_buf", None) is None:
            self._suffix_buf = torch.zeros(max(1, src.Tp_max - self.cfg.qf.n_query), dtype=torch.int32, device=self.dev)
        assert len(ids) <= self._suffix_buf.numel(), "question longer than the one the engine was built for"
        if ids:
            self._suffix_buf[:len(ids)] = torch.tensor(ids, dtype=torch.int32, device=self.dev)
        L.check(self._lib.cgpt_set_question(self._h, L.ptr(self._suffix_buf), len(ids)))
        self.Tp = self.cfg.qf.n_query + len(ids)
        self.S = self.P + self.Tp

    def set_answer_table(self, answer_table):
        """Replace the answer vocabulary: iterable of (token id sequence, class id); unknown answers -> num_classes - 1."""
        self.table_keys, self.table_vals = L.build_answer_table(answer_table, self.cfg.llm.eos_id, self.dev)
        L.check(self._lib.cgpt_set_answer_table(self._h, L.ptr(self.table_keys), L.ptr(self.table_vals),
                                                self.table_keys.numel()))

    def set_option(self, key, value):
        """run-time switches of the handle: "use_graphs", "early_exit"."""
        L.check(self._lib.cgpt_set_option(self._h, key.encode(), int(value)))

    # ------------------------------------------------------------------ workspace (caller-owned device memory)
    def workspace_bytes(self, B, encoder_only=False):
        n = C.c_int64(0)
        L.check(self._lib.cgpt_workspace_bytes(self._h, int(B), int(encoder_only), C.byref(n)))
        return n.value

    def reserve(self, B, encoder_only=False, images=1):
        """Make sure the workspace serves batches of B samples, drawn from up to `images` images per pass
        (re-binding drops the captured graphs)."""
        if (self._ws is not None and B <= self._ws_B and images <= self._images
                and (self._ws_encoder_only == encoder_only or not self._ws_encoder_only)):
            return
        if images > self._images:
            self.set_option("max_images", images)      # changes the layout: unbinds the workspace
            self._images, self._ws, self._ws_B = int(images), None, 0
            B = max(B, 1)
        need = self.workspace_bytes(B, encoder_only)
        self._ws = None
        torch.cuda.empty_cache()
        with torch.cuda.device(self.dev):
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.dev)
            L.check(self._lib.cgpt_bind_workspace(self._h, L.ptr(self._ws), need, int(B), int(encoder_only),
                                                  L.stream_ptr()))
        self._ws_B, self._ws_encoder_only = int(B), bool(encoder_only)

    def _ws_view(self, ptr_value, nbytes):
        off = ptr_value - self._ws.data_ptr()
        assert 0 <= off and off + nbytes <= self._ws.numel()
        return self._ws[off:off + nbytes]

    # ------------------------------------------------------------------ noise spec
    def _spec(self, sigma, seed, stream_id, noise_space, noise_kind, mean, std, eps=None, eps_first=0):
        s = NoiseSpecC()
        s.seed, s.stream_id, s.sigma = int(seed) & 0xFFFFFFFFFFFFFFFF, int(stream_id) & 0xFFFFFFFF, float(sigma)
        s.mean = (C.c_float * 3)(*[float(m) for m in mean])
        s.std = (C.c_float * 3)(*[float(m) for m in std])
        s.noise_space, s.noise_kind = int(noise_space), int(noise_kind)
        if eps is not None:
            S = self.cfg.vit.img_size
            assert eps.is_cuda and eps.dtype == torch.float32 and eps.is_contiguous() and tuple(eps.shape[1:]) == (3, S, S)
            # the library indexes injected draws by GLOBAL sample index: eps[0] belongs to sample eps_first
            s.eps = eps.data_ptr() - int(eps_first) * 3 * S * S * 4
        return s

    def _check_x(self, x):
        S = self.cfg.vit.img_size
        assert x.dtype == torch.float32 and tuple(x.shape) == (3, S, S) and x.is_contiguous()

    # ------------------------------------------------------------------ per-batch entry (Smooth's fused hook)
    @torch.no_grad()
    def noisy_labels(self, x, B, sigma, *, eps=None, seed=0, stream_id=0, first_sample=0,
                     noise_space=L.SPACE_NORMALIZED, noise_kind=L.NOISE_GAUSSIAN,
                     mean=L.BLIP_MEAN, std=L.BLIP_STD):
        """labels[b] = f(x + sigma*eps_b), b = first_sample .. +B; eps (optional) = this batch's draws [B,3,S,S]."""
        self._check_x(x)
        assert x.is_cuda
        self.reserve(B)
        spec = self._spec(sigma, seed, stream_id, noise_space, noise_kind, mean, std, eps, first_sample)
        labels = torch.empty(B, dtype=torch.int32, device=self.dev)
        L.check(self._lib.cgpt_noisy_labels(self._h, L.ptr(x), C.byref(spec), int(first_sample), int(B),
                                            L.ptr(labels), L.stream_ptr()))
        self.last_steps = self._lib.cgpt_last_decode_steps(self._h)
        return labels

    # ------------------------------------------------------------------ whole-loop entries
    def _comm_for(self, process_group):
        if process_group is None:
            return 0, 1, None
        import torch.distributed as dist
        g = None if process_group is True else process_group
        rank, world = dist.get_rank(g), dist.get_world_size(g)
        if world == 1:
            return 0, 1, None
        if self._comm is None:
            self._comm = CountsComm(process_group)
        return rank, world, self._comm.handle

    @torch.no_grad()
    def sample_noise(self, x, num, batch_size, sigma, *, base=0, split=None, eps=None, seed=0, stream_id=0,
                     noise_space=L.SPACE_NORMALIZED, noise_kind=L.NOISE_GAUSSIAN, mean=L.BLIP_MEAN, std=L.BLIP_STD,
                     process_group=None):
        """Smooth._sample_noise over the global range [base, base+num): device int64 counts [nvec, num_classes].
        eps (optional): injected draws indexed by global sample index (eps[0] = sample 0)."""
        self._check_x(x)
        self.reserve(min(int(batch_size), max(1, int(num))))
        rank, world, comm = self._comm_for(process_group)
        spec = self._spec(sigma, seed, stream_id, noise_space, noise_kind, mean, std, eps, 0)
        nvec = 1 if split is None else 2
        counts = torch.empty(nvec, self.num_classes, dtype=torch.int64, device=self.dev)
        invalid = torch.empty(1, dtype=torch.int32, device=self.dev)
        xp = C.c_void_p(x.data_ptr())
        L.check(self._lib.cgpt_sample_noise(self._h, xp, C.byref(spec), int(base), int(num), int(batch_size),
                                            -1 if split is None else int(split), rank, world, comm,
                                            L.ptr(counts), L.ptr(invalid), L.stream_ptr()))
        self.last_steps = self._lib.cgpt_last_decode_steps(self._h)
        self.last_invalid = invalid
        return counts

    def _last_counts(self):
        p = C.c_void_p()
        L.check(self._lib.cgpt_last_counts(self._h, C.byref(p)))
        return self._ws_view(p.value, 2 * self.num_classes * 8).view(torch.int64).view(2, self.num_classes)

    @torch.no_grad()
    def certify(self, x, n0, n, alpha, batch_size, sigma, *, eps=None, seed=0, stream_id=0,
                noise_space=L.SPACE_NORMALIZED, noise_kind=L.NOISE_GAUSSIAN, mean=L.BLIP_MEAN, std=L.BLIP_STD,
                process_group=None, exact_tail=True):
        """Smooth.certify in one library call.  x: [3,S,S] fp32, host (pinned or pageable) or device.
        Returns (label or -1, radius, detail dict)."""
        self._check_x(x)
        self.reserve(min(int(batch_size), int(n0 + n)))
        rank, world, comm = self._comm_for(process_group)
        spec = self._spec(sigma, seed, stream_id, noise_space, noise_kind, mean, std, eps, 0)
        # pABar / radius table of this (n, alpha), built once on the host with the reference's SciPy calls
        self._lut = L.radius_lut(n, alpha, self.dev) if exact_tail else None
        L.check(self._lib.cgpt_set_radius_lut(self._h, int(n), float(alpha), L.ptr(self._lut)))
        label, radius = C.c_int(0), C.c_double(0.0)
        detail = (C.c_double * 3)()
        with torch.cuda.device(self.dev):
            L.check(self._lib.cgpt_certify(self._h, C.c_void_p(x.data_ptr()), C.byref(spec), int(n0), int(n),
                                           float(alpha), int(batch_size), rank, world, comm, C.byref(label),
                                           C.byref(radius), detail, L.stream_ptr()))
        self.last_steps = self._lib.cgpt_last_decode_steps(self._h)
        counts = self._last_counts()
        return label.value, radius.value, {"cAHat": int(detail[0]), "pABar": detail[1], "nA": int(detail[2]),
                                           "counts_selection": counts[0], "counts_estimation": counts[1]}

    @torch.no_grad()
    def certify_batch(self, xs, n0, n, alpha, batch_size, sigma, *, seed=0, stream_id=0, noise_space=L.SPACE_NORMALIZED,
                      noise_kind=L.NOISE_GAUSSIAN, mean=L.BLIP_MEAN, std=L.BLIP_STD, process_group=None, exact_tail=True):
        """Smooth.certify of K images in shared passes (cgpt_certify_batch): every pass holds this rank's next draws
        of ALL K images.  xs: list of [3,S,S] fp32 tensors (host or device); image k uses Philox stream stream_id + k.
        Returns a list of K (label or -1, radius, detail dict); per image bit-identical to `certify`."""
        K = len(xs)
        assert 1 <= K <= 64
        for x in xs:
            self._check_x(x)
        rank, world, comm = self._comm_for(process_group)
        per_rank = -(-int(n0 + n) // world)
        self.reserve(min(int(batch_size), K * per_rank), images=K)
        spec = self._spec(sigma, seed, stream_id, noise_space, noise_kind, mean, std, None, 0)
        self._lut = L.radius_lut(n, alpha, self.dev) if exact_tail else None
        L.check(self._lib.cgpt_set_radius_lut(self._h, int(n), float(alpha), L.ptr(self._lut)))
        ptrs = (C.c_void_p * K)(*[x.data_ptr() for x in xs])
        labels, radii, detail = (C.c_int * K)(), (C.c_double * K)(), (C.c_double * (3 * K))()
        with torch.cuda.device(self.dev):
            L.check(self._lib.cgpt_certify_batch(self._h, ptrs, K, C.byref(spec), int(n0), int(n), float(alpha),
                                                 int(batch_size), rank, world, comm, labels, radii, detail, L.stream_ptr()))
        self.last_steps = self._lib.cgpt_last_decode_steps(self._h)
        p = C.c_void_p()
        L.check(self._lib.cgpt_last_counts(self._h, C.byref(p)))
        counts = self._ws_view(p.value, K * 2 * self.num_classes * 8).view(torch.int64).view(K, 2, self.num_classes).clone()
        return [(labels[k], radii[k], {"cAHat": int(detail[3 * k]), "pABar": detail[3 * k + 1], "nA": int(detail[3 * k + 2]),
                                       "counts_selection": counts[k, 0], "counts_estimation": counts[k, 1]})
                for k in range(K)]

    @torch.no_grad()
    def predict(self, x, n, alpha, batch_size, sigma, *, eps=None, seed=0, stream_id=0,
                noise_space=L.SPACE_NORMALIZED, noise_kind=L.NOISE_GAUSSIAN, mean=L.BLIP_MEAN, std=L.BLIP_STD,
                process_group=None):
        """Smooth.predict in one library call; returns (label or -1, p-value, counts)."""
        self._check_x(x)
        self.reserve(min(int(batch_size), int(n)))
        rank, world, comm = self._comm_for(process_group)
        spec = self._spec(sigma, seed, stream_id, noise_space, noise_kind, mean, std, eps, 0)
        label = C.c_int(0)
        detail = (C.c_double * 1)()
        with torch.cuda.device(self.dev):
            L.check(self._lib.cgpt_predict(self._h, C.c_void_p(x.data_ptr()), C.byref(spec), int(n), float(alpha),
                                           int(batch_size), rank, world, comm, C.byref(label), detail,
                                           L.stream_ptr()))
        self.last_steps = self._lib.cgpt_last_decode_steps(self._h)
        return label.value, detail[0], self._last_counts()[0]

    # ------------------------------------------------------------------ per-subsystem entries
    @torch.no_grad()
    def vit_forward(self, patches):
        """patches bf16 [B*G*G, 592] -> ln_vision(ViT) tokens bf16 [B, T, dim]."""
        v = self.cfg.vit
        B = patches.shape[0] // (v.grid * v.grid)
        assert patches.dtype == torch.bfloat16 and patches.is_contiguous() and patches.shape[1] == 592
        if self._ws is None or B > self._ws_B:
            self.reserve(B, encoder_only=self._ws is None)
        out = torch.empty(B, v.tokens, v.dim, dtype=torch.bfloat16, device=self.dev)
        L.check(self._lib.cgpt_vit_forward(self._h, L.ptr(patches), B, L.ptr(out), L.stream_ptr()))
        return out

    @torch.no_grad()
    def qformer_forward(self, tokens, want_llm_embeds=True):
        """tokens bf16 [B, T, dim] -> (queries [B, 32, qf.hidden], inputs_llama [B, 32, llm.hidden] or None)."""
        q, l = self.cfg.qf, self.cfg.llm
        B = tokens.shape[0]
        assert tokens.dtype == torch.bfloat16 and tokens.is_contiguous()
        if self._ws is None or B > self._ws_B:
            self.reserve(B, encoder_only=self._ws is None)
        out = torch.empty(B, q.n_query, q.hidden, dtype=torch.bfloat16, device=self.dev)
        emb = torch.empty(B, q.n_query, l.hidden, dtype=torch.bfloat16, device=self.dev) if want_llm_embeds else None
        L.check(self._lib.cgpt_qformer_forward(self._h, L.ptr(tokens), B, L.ptr(out), L.ptr(emb), L.stream_ptr()))
        return out, emb

    @torch.no_grad()
    def llm_prefill_decode(self, queries):
        """queries bf16 [B, 32, qf.hidden] -> (ids int32 [B, max_new], top-2 margins f32 [B, max_new], steps)."""
        B = queries.shape[0]
        assert queries.dtype == torch.bfloat16 and queries.is_contiguous()
        self.reserve(B)
        ids = torch.empty(B, self.max_new_tokens, dtype=torch.int32, device=self.dev)
        margin = torch.empty(B, self.max_new_tokens, dtype=torch.float32, device=self.dev)
        steps = C.c_int(0)
        L.check(self._lib.cgpt_llm_prefill_decode(self._h, L.ptr(queries), B, L.ptr(ids), L.ptr(margin),
                                                  C.byref(steps), L.stream_ptr()))
        self.last_steps = steps.value
        return ids, margin, steps.value

    @torch.no_grad()
    def lm_loss(self, images, answers, noise_level=0.0, *, noise_kind=L.NOISE_UNIFORM, noise_space=L.SPACE_NORMALIZED,
                seed=0, step=0, mean=L.BLIP_MEAN, std=L.BLIP_STD, label_smoothing=0.1):
        """Teacher-forced LM loss of the fine-tune / validation forward (MiniGPTBase.forward, minigpt_base.py:323-362;
        modeling_llama.py:101-123) for B different images [B,3,S,S] with answers [B, na] (int, -100 = padding).
        Noise as in MiniGPT4FineTuneAgent.maybe_add_noise (agents/minigpt4_finetune_agent.py:142-148:
        image + rand_like(image) * noise_level), drawn by K1 with Philox key (seed, step, image index).
        Returns (mean loss over answer tokens, per-token losses [B, na]); forward only."""
        B, na = answers.shape
        v = self.cfg.vit
        assert tuple(images.shape) == (B, 3, v.img_size, v.img_size) and images.is_cuda
        self.reserve(max(B * max(na, 2), self._ws_B))
        G2 = v.grid * v.grid
        patches = torch.empty(B * G2, 592, dtype=torch.bfloat16, device=self.dev)
        for b in range(B):      # a different image per row: one K1 launch each (training batches are small)
            L.noise_patchify(images[b].float().contiguous(), 1, float(noise_level), seed=seed, stream_id=step, first_sample=b,
                             noise_space=noise_space, noise_kind=noise_kind, mean=mean, std=std,
                             out=patches[b * G2:(b + 1) * G2])
        ans = answers.to(device=self.dev, dtype=torch.int32).contiguous()
        tok = torch.empty(B * na, dtype=torch.float32, device=self.dev)
        mc = torch.empty(2, dtype=torch.float32, device=self.dev)
        self.set_option("label_smoothing_permille", int(round(1000 * label_smoothing)))   # modeling_llama.py:107
        L.check(self._lib.cgpt_lm_loss(self._h, L.ptr(patches), B, L.ptr(ans), na, L.ptr(tok), L.ptr(mc), L.stream_ptr()))
        return mc[0], tok.view(B, na)

    @torch.no_grad()
    def encode_noisy(self, x, B, sigma, *, seed=0, stream_id=0, first_sample=0, noise_space=L.SPACE_NORMALIZED,
                     noise_kind=L.NOISE_GAUSSIAN, mean=L.BLIP_MEAN, std=L.BLIP_STD):
        """MiniGPT4.encode_img (minigpt4.py:121-149) of B noisy copies of x -> inputs_llama [B, 32, llm.hidden]."""
        patches = L.noise_patchify(x, B, sigma, seed=seed, stream_id=stream_id, first_sample=first_sample,
                                   noise_space=noise_space, noise_kind=noise_kind, mean=mean, std=std)
        if self._ws is None or B > self._ws_B:
            self.reserve(B, encoder_only=self._ws is None)
        tokens = self.vit_forward(patches)
        return self.qformer_forward(tokens)[1]
