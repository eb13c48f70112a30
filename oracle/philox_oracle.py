"""ORACLE (test infrastructure, never the product path).

numpy restatement of the random stream defined in certifiedgpt_b200/csrc/noise_patchify.cu and
of the noise + Normalize + patchify step it fuses:
  - smoothing.py:95-97   batch = x.repeat(B,1,1,1); noise = randn_like(batch) * sigma; batch + noise
  - processors/base_processor.py:17-34   transforms.Normalize(mean, std)
  - eva_vit.py:202-209   Conv2d(k=s=14) unfold order: column = c*196 + ky*14 + kx

The reference draws its noise from torch's global CUDA generator (smoothing.py:96); that stream
is not reproduced.  Philox4x32-10 follows Salmon et al. (Random123); the
known-answer vectors of that paper pin `philox4x32_10` in tests/test_oracle_cpu.py.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10; all inputs broadcastable uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(v, dtype=np.uint32) for v in (c0, c1, c2, c3)]
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            n0 = (p1 >> np.uint64(32)).astype(np.uint32) ^ c1 ^ k0
            n1 = (p1 & MASK).astype(np.uint32)
            n2 = (p0 >> np.uint64(32)).astype(np.uint32) ^ c3 ^ k1
            n3 = (p0 & MASK).astype(np.uint32)
            c0, c1, c2, c3 = n0, n1, n2, n3
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def _box_muller(r0, r1):
    f = np.float32
    u0 = r0.astype(f) * f(2.3283064365386963e-10) + f(1.1641532182693481e-10)
    u1 = r1.astype(f) * f(2.3283064365386963e-10) + f(1.1641532182693481e-10)
    rad = np.sqrt(f(-2.0) * np.log(u0)).astype(f)
    ang = (np.float64(2.0) * u1.astype(np.float64)) * np.pi
    return (rad * np.cos(ang).astype(f)).astype(f), (rad * np.sin(ang).astype(f)).astype(f)


def draws(n_elems, samples, seed=0, stream_id=0, kind="gaussian"):
    """Standard draws eps[len(samples), n_elems] (image element order) of the K1 stream."""
    assert n_elems % 4 == 0
    samples = np.asarray(samples, dtype=np.uint64)
    g = np.arange(n_elems // 4, dtype=np.uint32)[None, :]
    s_lo = (samples & MASK).astype(np.uint32)[:, None]
    s_hi = (samples >> np.uint64(32)).astype(np.uint32)[:, None]
    r = philox4x32_10(g, s_lo, np.uint32(stream_id), s_hi, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    out = np.empty((len(samples), n_elems // 4, 4), dtype=np.float32)
    if kind == "gaussian":
        out[:, :, 0], out[:, :, 1] = _box_muller(r[0], r[1])
        out[:, :, 2], out[:, :, 3] = _box_muller(r[2], r[3])
    else:
        for j in range(4):
            out[:, :, j] = (r[j] >> np.uint32(8)).astype(np.float32) * np.float32(5.9604644775390625e-08)
    return out.reshape(len(samples), n_elems)


def noisy_batch(x, eps, sigma, mean=None, std=None):
    """x [C,H,W] fp32, eps [B,C,H,W] fp32 -> (x + eps*sigma), then Normalize when mean/std given.
    fp32 throughout with the same operation order as the reference (mul, add, sub, div)."""
    f = np.float32
    x = np.asarray(x, dtype=f)
    eps = np.asarray(eps, dtype=f)
    v = x[None] + eps * f(sigma)
    if mean is not None:
        m = np.asarray(mean, dtype=f).reshape(1, -1, 1, 1)
        s = np.asarray(std, dtype=f).reshape(1, -1, 1, 1)
        v = (v - m) / s
    return v.astype(f)


def patchify(batch, patch=14, k_pad=592):
    """[B,3,S,S] -> [B*G*G, k_pad] with column c*196+ky*14+kx (Conv2d unfold order), zero pad."""
    B, C, S, _ = batch.shape
    G = S // patch
    t = batch.reshape(B, C, G, patch, G, patch).transpose(0, 2, 4, 1, 3, 5)  # B,py,px,c,ky,kx
    t = t.reshape(B * G * G, C * patch * patch)
    out = np.zeros((B * G * G, k_pad), dtype=batch.dtype)
    out[:, : t.shape[1]] = t
    return out
