"""GPU tests of libcgpt's native engine (cgpt_create ... cgpt_certify, include/cgpt.h):

* the C++ orchestration launches the same kernels in the same order as the Python twin (engine.py), so
  token ids, labels and counts must be BIT-IDENTICAL to it (eager and CUDA-graph replay);
* the per-subsystem entry points (cgpt_vit_forward / cgpt_qformer_forward / cgpt_llm_prefill_decode)
  match the fp32 oracle within bf16 tolerance (rel 2e-2) and the Python twin bit for bit;
* Smooth over the native engine equals the oracle on identical injected noise (counts exact wherever
  the oracle's top-2 margin exceeds 1e-2), for device and for host `x`;
* error behaviour: unbound weights, oversized batches and missing workspaces fail loudly.
"""
import numpy as np
import pytest
import torch

from certifiedgpt_b200.config import LlmConfig, ModelConfig, QFormerConfig, VitConfig
from certifiedgpt_b200.weights import random_state_dict, round_to_bf16
from oracle import model_oracle as mo
from oracle import smoothing_oracle as so

pytestmark = pytest.mark.gpu
REL = 2e-2
MARGIN = 1e-2

WIDE = ModelConfig(vit=VitConfig(img_size=56, depth=2), qf=QFormerConfig(layers=2),
                   llm=LlmConfig(hidden=512, layers=2, heads=4, inter=1024, vocab=512))


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def _setup(cfg, seed, max_new=3, n_classes=8, prefix=(1, 5, 6), suffix=(7, 8, 9, 10, 11), use_graphs=True,
           early_exit=True, oracle=False, all_pairs=False):
    from certifiedgpt_b200.engine import MiniGPT4Engine
    from certifiedgpt_b200.native import NativeMiniGPT4Engine
    sd = round_to_bf16(random_state_dict(cfg, seed=seed))
    V = cfg.llm.vocab
    table = [((t,), t % (n_classes - 1)) for t in range(3, V)]
    if all_pairs:   # every 2-token answer has its own class: label histograms of 2-token runs are not degenerate
        table += [((t, u), (t * 7 + u) % (n_classes - 1)) for t in range(3, V) for u in range(3, V)]
    else:
        table += [((t, u), (t + u) % (n_classes - 1)) for t in range(3, V, 7) for u in range(3, V, 5)]
    py = MiniGPT4Engine(cfg, sd, prefix, suffix, table, n_classes, max_new_tokens=max_new, use_graphs=False,
                        early_exit=early_exit)
    nat = NativeMiniGPT4Engine.from_engine(py, use_graphs=use_graphs)
    orc = mo.MiniGPT4ClassifierOracle(sd, cfg, prefix, suffix, table, n_classes, max_new_tokens=max_new) if oracle else None
    return sd, py, nat, orc


@pytest.mark.parametrize("name,cfg", [("tiny", ModelConfig.tiny()), ("wide", WIDE)])
@pytest.mark.parametrize("use_graphs", [False, True])
def test_noisy_labels_bit_identical_to_python_twin(name, cfg, use_graphs):
    sd, py, nat, _ = _setup(cfg, seed=5, use_graphs=use_graphs)
    S = cfg.vit.img_size
    x = torch.rand(3, S, S, generator=torch.Generator().manual_seed(3)).cuda()
    for first, B in [(0, 7), (7, 16), (100, 7)]:      # a repeated batch size replays the captured graph
        got = nat.noisy_labels(x, B, 0.5, seed=11, stream_id=2, first_sample=first)
        want = py.noisy_labels(x, B, 0.5, seed=11, stream_id=2, first_sample=first)
        torch.cuda.synchronize()
        assert torch.equal(got.cpu(), want.cpu()), (first, B)
        assert nat.last_steps == py.last_steps


def test_subsystem_entry_points_match_twin_and_oracle():
    cfg = WIDE
    sd, py, nat, orc = _setup(cfg, seed=21, use_graphs=False, oracle=True)
    from certifiedgpt_b200 import _lib as L
    g = torch.Generator().manual_seed(9)
    B, S = 5, cfg.vit.img_size
    images = torch.randn(B, 3, S, S, generator=g)
    # patches of B clean images (sigma = 0), as the twin's forward_images builds them
    patches = torch.cat([L.noise_patchify(images[b].cuda().contiguous(), 1, 0.0) for b in range(B)])
    got = {}
    py.forward_images(images.cuda(), collect=got)
    tokens = nat.vit_forward(patches)
    queries, inputs_llama = nat.qformer_forward(tokens)
    ids, margins, steps = nat.llm_prefill_decode(queries)
    torch.cuda.synchronize()
    # bit-identical to the Python-driven kernel sequence
    assert torch.equal(tokens.float().cpu(), got["image_embeds"].cpu())
    assert torch.equal(queries.float().cpu(), got["qformer"].cpu())
    assert torch.equal(ids.cpu(), got["ids"].cpu())
    assert torch.equal(margins.cpu(), got["margins"].cpu())
    assert steps == py.last_steps
    # and within bf16 tolerance of the fp32 oracle (reference eva_vit / Qformer / HF Llama restatement)
    ref = {}
    with torch.no_grad():
        img = mo.encode_img(sd, cfg, images, collect=ref)
        embeds = mo.build_prompt_embeds(sd, cfg, img, orc.prefix_ids, orc.suffix_ids)
        rids, first_logits, rmargins = mo.generate_ids(sd, cfg, embeds, orc.max_new_tokens)
    assert _rel(tokens, ref["image_embeds"]) < REL
    assert _rel(inputs_llama, img) < REL
    safe = (rmargins > MARGIN).all(dim=1)
    assert safe.float().mean() > 0.5
    assert torch.equal(ids.cpu().long()[safe], rids[safe])


@pytest.mark.parametrize("use_graphs", [False, True])
def test_smooth_native_matches_oracle_on_injected_noise(use_graphs):
    from certifiedgpt_b200.randomized_smoothing.smoothing import Smooth
    cfg = ModelConfig.tiny()
    n_classes = 6
    sd, py, nat, orc = _setup(cfg, seed=1, max_new=2, n_classes=n_classes, use_graphs=use_graphs, oracle=True)
    S = cfg.vit.img_size
    x = torch.rand(3, S, S, generator=torch.Generator().manual_seed(1000))
    n0, n, sigma, alpha = 24, 72, 0.25, 0.001
    eps = torch.randn(n0 + n, 3, S, S, generator=torch.Generator().manual_seed(1234))
    sm = Smooth(nat, n_classes, sigma)
    sm.inject_noise(eps.cuda())
    label, radius = sm.certify(x.cuda(), n0, n, alpha, 32)
    sm.last_label, sm.last_radius = label, radius
    got_sel, got_est = sm.last_counts_selection.cpu().numpy(), sm.last_counts_estimation.cpu().numpy()
    # the Python twin on the same injected noise: identical counts, label, radius
    sp = Smooth(py, n_classes, sigma)
    sp.inject_noise(eps.cuda())
    plabel, pradius = sp.certify(x.cuda(), n0, n, alpha, 32)
    assert np.array_equal(got_sel, sp.last_counts_selection.cpu().numpy())
    assert np.array_equal(got_est, sp.last_counts_estimation.cpu().numpy())
    assert (label, radius) == (plabel, pradius)
    # the oracle, PER SAMPLE: every draw whose top-2 logit margin exceeds 1e-2 at every decode step gets the oracle's label,
    # and the count vectors are exactly the histogram of the per-sample labels
    _assert_per_sample_parity(nat, orc, sm, x, eps, n0, n, sigma, alpha, n_classes, bs=32)


def _oracle_per_sample(orc, x, eps, sigma, bs=64):
    labels, margins = [], []
    for f in range(0, eps.shape[0], bs):
        orc(x[None] + eps[f:f + bs] * sigma)
        labels.append(orc.last["labels"].clone())
        margins.append(orc.last["margins"].clone())
    labels, margins = torch.cat(labels), torch.cat(margins)
    return labels, (margins > MARGIN).all(dim=1)


def _assert_per_sample_parity(nat, orc, sm, x, eps, n0, n, sigma, alpha, n_classes, bs):
    """`sm` has just run certify(x, n0, n, alpha, bs) over `nat` on the injected draws `eps` [n0+n]."""
    got_sel, got_est = sm.last_counts_selection.cpu(), sm.last_counts_estimation.cpu()
    lab = torch.cat([nat.noisy_labels(x.cuda(), min(97, n0 + n - f), sigma, eps=eps[f:f + 97].cuda(), first_sample=f)
                     for f in range(0, n0 + n, 97)]).cpu().long()          # another batching: labels are per-sample
    assert torch.equal(torch.bincount(lab[:n0], minlength=n_classes), got_sel)
    assert torch.equal(torch.bincount(lab[n0:], minlength=n_classes), got_est)
    ref, safe = _oracle_per_sample(orc, x, eps, sigma)
    assert safe.float().mean() > 0.5, "test inputs too close to ties to be informative"
    assert torch.equal(lab[safe], ref[safe])                                # bit-exact wherever the margin is safe
    # hence the counts differ from the oracle's by at most the unsafe draws, and not at all without them
    ref_sel = torch.bincount(ref[:n0], minlength=n_classes)
    ref_est = torch.bincount(ref[n0:], minlength=n_classes)
    unsafe_sel, unsafe_est = int((~safe[:n0]).sum()), int((~safe[n0:]).sum())
    assert int((got_sel - ref_sel).abs().sum()) <= 2 * unsafe_sel
    assert int((got_est - ref_est).abs().sum()) <= 2 * unsafe_est
    if torch.equal(got_sel, ref_sel) and torch.equal(got_est, ref_est):
        want = so.certify_tail(ref_sel.numpy(), ref_est.numpy(), n, alpha, sigma)
        assert (sm.last_label, sm.last_radius) == want                      # label and radius bit-exact


def test_certify_n0_100_n_1000_per_sample_parity_on_the_wide_config():
    """BASELINE configs[1]'s draw counts (N0 = 100, N = 1000) through cgpt_certify on the wide model: per-sample labels
    equal the oracle's wherever its top-2 margin exceeds 1e-2 (north_star), counts are their histogram."""
    from certifiedgpt_b200.randomized_smoothing.smoothing import Smooth
    cfg, n_classes = WIDE, 8
    # 2-token answers, every (t, u) pair a class: on the CPU oracle ~80 % of the 1100 draws are margin-safe and they
    # spread over 6 classes (122 / 102 / ... of the first 240), so the per-sample comparison is informative
    sd, py, nat, orc = _setup(cfg, seed=31, max_new=2, n_classes=n_classes, oracle=True, all_pairs=True)
    S = cfg.vit.img_size
    x = torch.rand(3, S, S, generator=torch.Generator().manual_seed(1000))
    n0, n, sigma, alpha = 100, 1000, 0.25, 0.001
    eps = torch.randn(n0 + n, 3, S, S, generator=torch.Generator().manual_seed(1234))
    sm = Smooth(nat, n_classes, sigma)
    sm.inject_noise(eps.cuda())
    sm.last_label, sm.last_radius = sm.certify(x.cuda(), n0, n, alpha, 256)
    _assert_per_sample_parity(nat, orc, sm, x, eps, n0, n, sigma, alpha, n_classes, bs=256)


def test_certify_predict_host_x_and_batch_size_invariance():
    from certifiedgpt_b200.randomized_smoothing.smoothing import Smooth
    cfg = ModelConfig.tiny()
    n_classes = 6
    sd, py, nat, _ = _setup(cfg, seed=2, max_new=2, n_classes=n_classes)
    S = cfg.vit.img_size
    x = torch.rand(3, S, S, generator=torch.Generator().manual_seed(7))
    res = []
    for xin, bs in [(x.cuda(), 40), (x, 40), (x.pin_memory(), 17), (x.cuda(), 200)]:
        sm = Smooth(nat, n_classes, 0.5, seed=3)       # Philox draws keyed by global sample index
        out = sm.certify(xin, 20, 100, 0.001, bs)
        res.append((out, sm.last_counts_selection.cpu().numpy().copy(), sm.last_counts_estimation.cpu().numpy().copy()))
    for r in res[1:]:
        assert r[0] == res[0][0]
        assert np.array_equal(r[1], res[0][1]) and np.array_equal(r[2], res[0][2])
    # Python twin, sequential (non-fused) phases: same counts
    sp = Smooth(py, n_classes, 0.5, seed=3, fuse_selection=False)
    assert sp.certify(x.cuda(), 20, 100, 0.001, 40) == res[0][0]
    assert np.array_equal(sp.last_counts_estimation.cpu().numpy(), res[0][2])
    # predict
    sm = Smooth(nat, n_classes, 0.5, seed=3)
    sp = Smooth(py, n_classes, 0.5, seed=3)
    assert sm.predict(x, 64, 0.001, 25) == sp.predict(x.cuda(), 64, 0.001, 25)
    assert np.array_equal(sm.last_counts.cpu().numpy(), sp.last_counts.cpu().numpy())
    assert abs(sm.last_pvalue - sp.last_pvalue) <= 1e-15
    # _sample_noise keeps the reference's return type
    c = Smooth(nat, n_classes, 0.5, seed=3)._sample_noise(x.cuda(), 50, 16)
    assert isinstance(c, np.ndarray) and c.dtype.kind == "i" and c.sum() == 50 and len(c) == n_classes


def test_eos_early_exit_matches_twin():
    cfg = ModelConfig.tiny()
    from certifiedgpt_b200.engine import MiniGPT4Engine
    from certifiedgpt_b200.native import NativeMiniGPT4Engine
    sd = random_state_dict(cfg, seed=33)
    sd["llama_model.lm_head.weight"][cfg.llm.eos_id] *= 6.0      # EOS strongly preferred
    sd = round_to_bf16(sd)
    table = [((t,), t % 5) for t in range(3, cfg.llm.vocab)]
    py = MiniGPT4Engine(cfg, sd, (1, 4), (6, 7, 8), table, 6, max_new_tokens=5, use_graphs=False)
    S = cfg.vit.img_size
    x = torch.rand(3, S, S, generator=torch.Generator().manual_seed(5)).cuda()
    want = py.noisy_labels(x, 9, 0.25, seed=1)
    for use_graphs in (False, True):
        nat = NativeMiniGPT4Engine.from_engine(py, use_graphs=use_graphs)
        got = nat.noisy_labels(x, 9, 0.25, seed=1)
        assert torch.equal(got.cpu(), want.cpu())
        assert nat.last_steps == py.last_steps and nat.last_steps < 5     # early exit happened


def test_native_engine_fails_loudly():
    from certifiedgpt_b200 import _lib as L
    from certifiedgpt_b200 import native as N
    cfg = ModelConfig.tiny()
    sd, py, nat, _ = _setup(cfg, seed=4)
    h = N.lib()
    # a handle with one weight missing cannot bind a workspace; the message names the weight
    import ctypes as C
    c = N.config_struct(cfg, py.P, len(py.suffix_ids), 3, 1, 8, True, True)
    hd = C.c_void_p()
    L.check(h.cgpt_create(C.byref(c), C.byref(hd)))
    for name, t in py.w.items():
        if name == "vit.1.fc2.w":
            continue
        rows, cols = (1, t.numel()) if t.dim() == 1 else tuple(t.shape)
        L.check(h.cgpt_bind_weight(hd, name.encode(), L.ptr(t), rows, cols, L.DT_F32 if t.dtype == torch.float32 else L.DT_BF16))
    ws = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
    rc = h.cgpt_bind_workspace(hd, L.ptr(ws), ws.numel(), 1, 1, L.stream_ptr())
    assert rc != 0 and b"vit.1.fc2.w" in h.cgpt_last_error()
    h.cgpt_destroy(hd)
    # too small a workspace, and a batch beyond the bound workspace
    need = nat.workspace_bytes(4)
    small = torch.empty(need // 2, dtype=torch.uint8, device="cuda")
    assert h.cgpt_bind_workspace(nat._h, L.ptr(small), small.numel(), 4, 0, L.stream_ptr()) != 0
    nat.reserve(4)
    lab = torch.empty(8, dtype=torch.int32, device="cuda")
    x = torch.rand(3, cfg.vit.img_size, cfg.vit.img_size).cuda()
    spec = nat._spec(0.25, 0, 0, 0, 0, L.BLIP_MEAN, L.BLIP_STD)
    rc = h.cgpt_noisy_labels(nat._h, L.ptr(x), C.byref(spec), 0, 8, L.ptr(lab), L.stream_ptr())
    assert rc != 0 and b"exceeds" in h.cgpt_last_error()


def test_gemm_profile_hook_counts_eager_launches():
    from certifiedgpt_b200 import native as N
    cfg = ModelConfig.tiny()
    sd, py, nat, _ = _setup(cfg, seed=6, use_graphs=False)
    x = torch.rand(3, cfg.vit.img_size, cfg.vit.img_size).cuda()
    nat.noisy_labels(x, 4, 0.25)
    N.gemm_profile_begin()
    nat.noisy_labels(x, 4, 0.25)
    rec = N.gemm_profile_end()
    v, q, l = cfg.vit, cfg.qf, cfg.llm
    n_cross = len(q.cross_layers())
    expect = (1 + 4 * v.depth) + (1 + q.layers * 4 + n_cross * 2) + 1 + nat.last_steps * (4 * l.layers + 1)
    assert len(rec) == expect
    assert all(ms >= 0 and fl > 0 for ms, fl, _ in rec)


@pytest.mark.parametrize("name,cfg", [("tiny", ModelConfig.tiny()), ("wide", WIDE)])
def test_lm_loss_forward_matches_oracle(name, cfg):
    """cgpt_lm_loss (validation / fine-tune forward, teacher forcing + shifted CE) vs the fp32 oracle, clean and with
    the fine-tune agent's uniform noise (noise regenerated on the host side from the same Philox draws)."""
    from certifiedgpt_b200 import _lib as L
    sd, py, nat, _ = _setup(cfg, seed=31, max_new=4)
    S = cfg.vit.img_size
    B = 5
    images = torch.randn(B, 3, S, S, generator=torch.Generator().manual_seed(2))
    V = cfg.llm.vocab
    answers = torch.tensor([[7, 9, 2, -100], [11, 2, -100, -100], [V - 1, 5, 6, 2], [3, 2, -100, -100], [8, 8, 8, 2]])
    loss, tok = nat.lm_loss(images.cuda(), answers, 0.0)
    ref_loss, ref_tok = mo.lm_loss(sd, cfg, images, py.prefix_ids, py.suffix_ids, answers, label_smoothing=0.1)
    assert abs(loss.item() - ref_loss.item()) < 2e-2 * max(1.0, ref_loss.item())
    assert (tok.cpu() - ref_tok).abs().max().item() < 5e-2 * max(1.0, ref_tok.max().item())
    assert (tok.cpu()[answers < 0] == 0).all()
    # uniform noise: image + U[0,1) * level, keyed by (seed, step, image index)
    lvl = 0.25
    loss_n, _ = nat.lm_loss(images.cuda(), answers, lvl, seed=3, step=5)
    noisy = torch.stack([L.noise_image(images[b].cuda().contiguous(), 1, lvl, seed=3, stream_id=5, first_sample=b,
                                       noise_kind=L.NOISE_UNIFORM)[0] for b in range(B)]).cpu()
    assert (noisy - images).min() >= 0 and (noisy - images).max() < lvl
    ref_n, _ = mo.lm_loss(sd, cfg, noisy, py.prefix_ids, py.suffix_ids, answers, label_smoothing=0.1)
    assert abs(loss_n.item() - ref_n.item()) < 2e-2 * max(1.0, ref_n.item())
    # the generate path still works afterwards (the loss pass wrote answer K/V into the cache rows decode reuses)
    x = torch.rand(3, S, S, generator=torch.Generator().manual_seed(3)).cuda()
    assert torch.equal(nat.noisy_labels(x, 4, 0.5, seed=1).cpu(), py.noisy_labels(x, 4, 0.5, seed=1).cpu())


def test_ce_loss_kernel_matches_torch():
    from certifiedgpt_b200 import _lib as L
    from certifiedgpt_b200 import native as N
    import ctypes as C
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(37, 32000, generator=g) * 4
    tg = torch.randint(0, 32000, (37,), generator=g)
    tg[5], tg[20] = -100, -100
    lg, tgd = logits.cuda(), tg.int().cuda()
    tok = torch.empty(37, device="cuda")
    mc = torch.empty(2, device="cuda")
    h = N.lib()
    h.cgpt_ce_loss.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.check(h.cgpt_ce_loss(L.ptr(lg), lg.stride(0), 37, 32000, L.ptr(tgd), L.ptr(tok), L.ptr(mc), L.stream_ptr()))
    ref = torch.nn.functional.cross_entropy(logits.double(), tg, ignore_index=-100, reduction="none")
    assert (tok.cpu().double() - ref).abs().max().item() < 1e-4
    assert abs(mc[0].item() - ref.sum().item() / 35) < 1e-4 and mc[1].item() == 35


def _raw_sample_noise(nat, x, num, batch_size, sigma, rank, world, split=None, base=0, seed=3):
    """cgpt_sample_noise for one (rank, world) slice without a communicator: this rank's own counts."""
    import ctypes as C
    from certifiedgpt_b200 import _lib as L
    nat.reserve(max(1, min(batch_size, num)))
    spec = nat._spec(sigma, seed, 0, L.SPACE_NORMALIZED, L.NOISE_GAUSSIAN, L.BLIP_MEAN, L.BLIP_STD)
    nvec = 1 if split is None else 2
    counts = torch.full((nvec, nat.num_classes), -7, dtype=torch.int64, device="cuda")   # the call must zero them
    invalid = torch.empty(1, dtype=torch.int32, device="cuda")
    L.check(nat._lib.cgpt_sample_noise(nat._h, C.c_void_p(x.data_ptr()), C.byref(spec), base, num, batch_size,
                                       -1 if split is None else split, rank, world, None, L.ptr(counts),
                                       L.ptr(invalid), L.stream_ptr()))
    torch.cuda.synchronize()
    assert int(invalid.item()) == 0
    return counts.cpu().numpy()


def test_rank_slices_sum_to_the_single_rank_counts_incl_empty_slices():
    """SURVEY 8(e) on one GPU: the union of the per-rank slices of [base, base+num) is the 1-rank run, for world
    sizes that do not divide num, for more ranks than draws (empty slices) and with the selection/estimation split."""
    cfg = ModelConfig.tiny()
    n_classes = 6
    sd, py, nat, _ = _setup(cfg, seed=9, max_new=2, n_classes=n_classes)
    S = cfg.vit.img_size
    x = torch.rand(3, S, S, generator=torch.Generator().manual_seed(11)).cuda()
    for num, split, base in [(45, None, 0), (45, 10, 0), (5, 2, 7), (1, None, 0)]:
        whole = _raw_sample_noise(nat, x, num, 16, 0.5, 0, 1, split=split, base=base)
        assert whole.sum() == num
        if split is not None:
            assert whole[0].sum() == split and whole[1].sum() == num - split
        for world in (2, 3, 8):
            parts = [_raw_sample_noise(nat, x, num, 16, 0.5, r, world, split=split, base=base) for r in range(world)]
            assert np.array_equal(sum(parts), whole), (num, split, base, world)
            sizes = [int(p.sum()) for p in parts]
            assert max(sizes) - min(sizes) <= 1                        # balanced contiguous slices
    # the same draws through a different batch size and through the Python twin's Smooth loop
    from certifiedgpt_b200.randomized_smoothing.smoothing import Smooth
    whole = _raw_sample_noise(nat, x, 45, 16, 0.5, 0, 1)
    assert np.array_equal(_raw_sample_noise(nat, x, 45, 1, 0.5, 0, 1), whole)
    assert np.array_equal(_raw_sample_noise(nat, x, 45, 4096, 0.5, 0, 1), whole)
    assert np.array_equal(Smooth(py, n_classes, 0.5, seed=3)._sample_noise(x, 45, 7), whole[0])


def test_empty_and_single_draw_calls():
    """num = 0 runs no batch and returns zero counts (smoothing.py:91: range(ceil(0 / batch_size)) is empty);
    one-draw certify / predict abstain (alpha ** 1 < 0.5; binomial test of 1 vs 0 gives p = 1)."""
    from certifiedgpt_b200 import _lib as L
    from certifiedgpt_b200.randomized_smoothing.smoothing import Smooth
    cfg = ModelConfig.tiny()
    n_classes = 6
    sd, py, nat, _ = _setup(cfg, seed=9, max_new=2, n_classes=n_classes)
    S = cfg.vit.img_size
    x = torch.rand(3, S, S, generator=torch.Generator().manual_seed(11)).cuda()
    assert _raw_sample_noise(nat, x, 0, 16, 0.5, 0, 1).sum() == 0
    assert _raw_sample_noise(nat, x, 0, 16, 0.5, 1, 2, split=0).sum() == 0
    for eng in (nat, py):
        sm = Smooth(eng, n_classes, 0.5, seed=3)
        z = sm._sample_noise(x, 0, 16)
        assert z.shape == (n_classes,) and z.sum() == 0
        assert sm.certify(x, 1, 1, 0.001, 1000) == (Smooth.ABSTAIN, 0.0)
        assert sm.predict(x, 1, 0.001, 1000) == Smooth.ABSTAIN
    # bad arguments come back as errors, not as crashes
    import ctypes as C
    spec = nat._spec(0.5, 0, 0, L.SPACE_NORMALIZED, L.NOISE_GAUSSIAN, L.BLIP_MEAN, L.BLIP_STD)
    counts = torch.zeros(2, n_classes, dtype=torch.int64, device="cuda")
    for num, bs, split, rank, world in [(-1, 8, -1, 0, 1), (8, 0, -1, 0, 1), (8, 8, 9, 0, 1), (8, 8, -1, 2, 2)]:
        rc = nat._lib.cgpt_sample_noise(nat._h, C.c_void_p(x.data_ptr()), C.byref(spec), 0, num, bs, split, rank, world,
                                        None, L.ptr(counts), None, L.stream_ptr())
        assert rc != 0 and nat._lib.cgpt_last_error()


@pytest.mark.parametrize("use_graphs", [False, True])
def test_certify_batch_equals_one_image_at_a_time(use_graphs):
    """cgpt_certify_batch: K images share every pass; per image the counts, label and radius of K successive
    Smooth.certify calls, bit for bit, whatever the batch size (image k is drawn from Philox stream image_id + k)."""
    from certifiedgpt_b200.randomized_smoothing.smoothing import Smooth
    cfg, n_classes = WIDE, 8
    sd, py, nat, _ = _setup(cfg, seed=31, max_new=2, n_classes=n_classes, use_graphs=use_graphs, all_pairs=True)
    S = cfg.vit.img_size
    xs = [torch.rand(3, S, S, generator=torch.Generator().manual_seed(50 + i)) for i in range(5)]
    n0, n, sigma, alpha = 20, 90, 0.5, 0.001
    one = Smooth(nat, n_classes, sigma, seed=9)
    want, want_counts = [], []
    for x in xs:
        want.append(one.certify(x.cuda(), n0, n, alpha, 64))
        want_counts.append((one.last_counts_selection.clone(), one.last_counts_estimation.clone()))
    for bs, host in ((550, False), (64, True), (7, False)):     # 110, 12 and 1 draw(s) of each image per pass
        many = Smooth(nat, n_classes, sigma, seed=9)
        got = many.certify_batch([x if host else x.cuda() for x in xs], n0, n, alpha, bs)
        assert got == want, bs
        for d, (sel, est) in zip(many.last_batch_detail, want_counts):
            assert torch.equal(d["counts_selection"], sel) and torch.equal(d["counts_estimation"], est)
            assert int(d["counts_selection"].sum()) == n0 and int(d["counts_estimation"].sum()) == n
        assert many.image_id == len(xs)
    # the histograms are not degenerate: the images do not all land in one class
    assert len({int(c[1].argmax()) for c in want_counts}) + sum(int((c[1] > 0).sum()) > 1 for c in want_counts) > 1


@pytest.mark.parametrize("use_graphs", [False, True])
def test_set_question_equals_an_engine_built_for_that_question(use_graphs):
    """cgpt_set_question / MiniGPT4Engine.set_question: one engine, built for the longest question, answers every
    (shorter or equal) question exactly as an engine constructed with that question does - ids, labels, decode steps."""
    from certifiedgpt_b200.engine import MiniGPT4Engine
    from certifiedgpt_b200.native import NativeMiniGPT4Engine
    cfg = WIDE
    long_q, short_q, other_q = (7, 8, 9, 10, 11, 12, 13), (20, 21, 22), (30, 31, 32, 33, 34, 35, 36)
    sd, py, nat, _ = _setup(cfg, seed=23, max_new=2, suffix=long_q, use_graphs=use_graphs, all_pairs=True)
    S = cfg.vit.img_size
    x = torch.rand(3, S, S, generator=torch.Generator().manual_seed(3)).cuda()
    want = {}
    for q in (long_q, short_q, other_q):
        ref = MiniGPT4Engine(cfg, sd, (1, 5, 6), q, [], 8, max_new_tokens=2, use_graphs=False)
        ref.table_keys, ref.table_vals = py.table_keys, py.table_vals
        want[q] = ref.noisy_labels(x, 24, 0.5, seed=5).clone().cpu()
    assert not (torch.equal(want[long_q], want[short_q]) and torch.equal(want[long_q], want[other_q])), "questions must matter"
    for q in (short_q, long_q, other_q, short_q):          # back and forth: graphs are keyed by the question length
        nat.set_question(q)
        py.set_question(q)
        got = nat.noisy_labels(x, 24, 0.5, seed=5).cpu()
        assert torch.equal(got, want[q]), q
        assert torch.equal(py.noisy_labels(x, 24, 0.5, seed=5).clone().cpu(), want[q]), q
    with pytest.raises(AssertionError):
        nat.set_question(tuple(range(3, 3 + len(long_q) + 1)))        # longer than the engine was built for
