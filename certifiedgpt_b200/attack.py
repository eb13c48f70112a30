"""Black-box attack inner loop (BASELINE.json configs[4]).

The reference describes this stage only in prose (README.md:62-64: "clean images are perturbed, textual and
visual features are compared across multiple backbone encoders, and attack parameters are updated over multiple
steps"; result table README.md:108-120, AttackVLM with a ViT-L/14 backbone) and ships no code for it.
This module is therefore a definition, not a port:

  per step (8 by default)
    1. Q query images x_adv + sigma_q * u_b with Philox Gaussian directions u_b: the SAME fused K1 kernel as the
       smoothing loop (noise + CLIP/BLIP Normalize + 14x14 patchify in one pass; the directions never touch HBM)
    2. CLIP ViT-L/14 image features of the Q queries (tcgen05 GEMMs with bias / QuickGELU / residual epilogues,
       LayerNorm and attention kernels of libcgpt) and their cosine to the target feature (cgpt_cosine_rows)
    3. random-gradient-free estimate g = mean_b ((s_b - s_0) / sigma_q) u_b (directions regenerated from the same
       Philox keys by cgpt_noise_image), sign step, projection onto the eps-ball around the clean image
    4. smoothed prediction of the victim at the new point: Smooth.predict(x_adv, N=100) through the MiniGPT-4 engine

`ClipVisionEngine` takes the state dict of `transformers.CLIPVisionModelWithProjection` (HF key names).
There is no CPU path: everything runs through libcgpt.so.
"""
from dataclasses import dataclass

import torch

from . import _lib as L


@dataclass
class ClipVisionConfig:
    """CLIP ViT-L/14 vision tower (openai/clip-vit-large-patch14)."""
    img_size: int = 224
    patch: int = 14
    hidden: int = 1024
    layers: int = 24
    heads: int = 16
    mlp: int = 4096
    proj: int = 768
    eps: float = 1e-5

    @property
    def grid(self):
        return self.img_size // self.patch

    @property
    def tokens(self):
        return self.grid * self.grid + 1

    @property
    def head_dim(self):
        return self.hidden // self.heads

    @staticmethod
    def tiny():
        return ClipVisionConfig(img_size=56, hidden=64, layers=2, heads=4, mlp=128, proj=32)


def _bf16(t, dev):
    return t.detach().to(device=dev, dtype=torch.bfloat16).contiguous()


def _f32(t, dev):
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


class ClipVisionEngine:
    """CLIP vision tower with projection on the libcgpt kernels: images / noisy copies -> image_embeds [B, proj] fp32."""

    def __init__(self, cfg: ClipVisionConfig, state_dict, device="cuda"):
        L.load()
        assert cfg.patch == 14, "the patchify kernel is written for 14x14 patches"
        self.cfg, self.dev = cfg, torch.device(device)
        self._pack(state_dict)
        self._buf, self._buf_B = None, 0

    def _pack(self, sd):
        c, dev, w = self.cfg, self.dev, {}
        p = "vision_model."
        pw = sd[p + "embeddings.patch_embedding.weight"].reshape(c.hidden, -1)       # [D, 588] (c, ky, kx)
        pwp = torch.zeros(c.hidden, 592, dtype=pw.dtype, device=pw.device)
        pwp[:, :588] = pw
        w["patch.w"] = _bf16(pwp, dev)
        pos = sd[p + "embeddings.position_embedding.weight"]
        assert pos.shape[0] == c.tokens
        w["pos"] = _f32(pos, dev)
        w["cls_pos"] = _f32(sd[p + "embeddings.class_embedding"] + pos[0], dev)
        w["pre.w"], w["pre.b"] = _f32(sd[p + "pre_layrnorm.weight"], dev), _f32(sd[p + "pre_layrnorm.bias"], dev)
        for i in range(c.layers):
            l, o = f"{p}encoder.layers.{i}.", f"l{i}."
            w[o + "ln1.w"], w[o + "ln1.b"] = _f32(sd[l + "layer_norm1.weight"], dev), _f32(sd[l + "layer_norm1.bias"], dev)
            w[o + "qkv.w"] = _bf16(torch.cat([sd[l + f"self_attn.{n}_proj.weight"] for n in "qkv"]), dev)
            w[o + "qkv.b"] = _f32(torch.cat([sd[l + f"self_attn.{n}_proj.bias"] for n in "qkv"]), dev)
            w[o + "o.w"], w[o + "o.b"] = _bf16(sd[l + "self_attn.out_proj.weight"], dev), _f32(sd[l + "self_attn.out_proj.bias"], dev)
            w[o + "ln2.w"], w[o + "ln2.b"] = _f32(sd[l + "layer_norm2.weight"], dev), _f32(sd[l + "layer_norm2.bias"], dev)
            w[o + "fc1.w"], w[o + "fc1.b"] = _bf16(sd[l + "mlp.fc1.weight"], dev), _f32(sd[l + "mlp.fc1.bias"], dev)
            w[o + "fc2.w"], w[o + "fc2.b"] = _bf16(sd[l + "mlp.fc2.weight"], dev), _f32(sd[l + "mlp.fc2.bias"], dev)
        w["post.w"], w["post.b"] = _f32(sd[p + "post_layernorm.weight"], dev), _f32(sd[p + "post_layernorm.bias"], dev)
        w["proj.w"] = _bf16(sd["visual_projection.weight"], dev)
        self.w = w

    def _buffers(self, B):
        if B <= self._buf_B:
            return self._buf
        c, dev = self.cfg, self.dev
        bf, f32 = torch.bfloat16, torch.float32
        M = B * c.tokens
        e = lambda *s, dtype=bf: torch.empty(*s, dtype=dtype, device=dev)
        self._buf = None
        self._buf = {"patches": e(B * c.grid * c.grid, 592), "res": e(M, c.hidden, dtype=f32), "xn": e(M, c.hidden),
                     "qkv": e(M, 3 * c.hidden), "att": e(M, c.hidden), "h": e(M, c.mlp), "pooled": e(B, c.hidden),
                     "feat": e(B, c.proj, dtype=f32)}
        self._buf_B = B
        return self._buf

    def features_from_patches(self, B, buf):
        """buf['patches'] [B*G*G, 592] -> image_embeds [B, proj] fp32 (modeling_clip.py CLIPVisionTransformer)."""
        c, w = self.cfg, self.w
        T, Pn, D = c.tokens, c.grid * c.grid, c.hidden
        M = B * T
        res, xn, qkv, att, h = (buf[k][:M] for k in ("res", "xn", "qkv", "att", "h"))
        # patch_embedding (conv14, no bias) + position_embedding; class_embedding + pos[0] into row 0 of each image
        L.gemm(buf["patches"][:B * Pn], w["patch.w"], out=res, row_add=w["pos"], row_period=Pn, row_add_offset=1,
               remap_stride=T, remap_offset=1)
        res.view(B, T, D)[:, 0] = w["cls_pos"]
        L.norm_rows(res, w["pre.w"], w["pre.b"], c.eps, res)              # pre_layrnorm, in place (rows are register-resident)
        scale = c.head_dim ** -0.5
        for i in range(c.layers):
            o = f"l{i}."
            L.norm_rows(res, w[o + "ln1.w"], w[o + "ln1.b"], c.eps, xn)
            L.gemm(xn, w[o + "qkv.w"], bias=w[o + "qkv.b"], out=qkv)
            L.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], att, B=B, H=c.heads, Tq=T, Tk=T,
                        head_dim=c.head_dim, scale=scale)
            L.gemm(att, w[o + "o.w"], bias=w[o + "o.b"], resid=res, out=res)
            L.norm_rows(res, w[o + "ln2.w"], w[o + "ln2.b"], c.eps, xn)
            L.gemm(xn, w[o + "fc1.w"], bias=w[o + "fc1.b"], act=L.ACT_QUICKGELU, out=h)
            L.gemm(h, w[o + "fc2.w"], bias=w[o + "fc2.b"], resid=res, out=res)
        pooled, feat = buf["pooled"][:B], buf["feat"][:B]
        L.norm_rows(res, w["post.w"], w["post.b"], c.eps, pooled, gather=(1, T, 0))   # post_layernorm on the class token
        L.gemm(pooled, w["proj.w"], out=feat)                                          # visual_projection (no bias)
        return feat

    @torch.no_grad()
    def encode_perturbed(self, x, B, sigma, *, seed=0, stream_id=0, first_sample=0, noise_space=L.SPACE_PIXEL,
                         mean=L.BLIP_MEAN, std=L.BLIP_STD):
        """Features of the B queries x + sigma * u_b (Philox Gaussian u_b keyed by (seed, stream_id, first_sample + b));
        sigma = 0 encodes x itself B times.  x: [3,S,S] fp32 pixel space (noise_space = PIXEL: CLIP Normalize in K1)."""
        buf = self._buffers(B)
        G2 = self.cfg.grid ** 2
        L.noise_patchify(x, B, sigma, seed=seed, stream_id=stream_id, first_sample=first_sample, noise_space=noise_space,
                         mean=mean, std=std, out=buf["patches"][:B * G2])
        return self.features_from_patches(B, buf)

    @torch.no_grad()
    def encode_images(self, images, *, normalized=True):
        """images [B,3,S,S] fp32 -> image_embeds [B, proj].  normalized: already CLIP-normalised (HF pixel_values)."""
        B = images.shape[0]
        buf = self._buffers(B)
        G2 = self.cfg.grid ** 2
        space = L.SPACE_NORMALIZED if normalized else L.SPACE_PIXEL
        for b in range(B):
            L.noise_patchify(images[b].float().contiguous(), 1, 0.0, noise_space=space, out=buf["patches"][b * G2:(b + 1) * G2])
        return self.features_from_patches(B, buf)


class BlackBoxAttack:
    """Query-based (random gradient-free) targeted attack scored by CLIP image-feature cosine, with the smoothed
    victim's prediction tracked at every step.  `smooth`: a certifiedgpt_b200 `Smooth` (noise_space='pixel')."""

    def __init__(self, clip: ClipVisionEngine, smooth=None, *, steps=8, queries=100, sigma_q=8.0 / 255, step_size=1.0 / 255,
                 eps=8.0 / 255, predict_n=100, predict_alpha=0.001, predict_batch=100, seed=0):
        self.clip, self.smooth = clip, smooth
        self.steps, self.queries = steps, queries
        self.sigma_q, self.step_size, self.eps = float(sigma_q), float(step_size), float(eps)
        self.predict_n, self.predict_alpha, self.predict_batch = predict_n, predict_alpha, predict_batch
        self.seed = int(seed)

    @torch.no_grad()
    def step(self, x_adv, x_clean, target_feat, step_idx):
        """One perturbation update; returns (x_adv_new, score_before, query_scores)."""
        Q = self.queries
        f0 = L.cosine_rows(self.clip.encode_perturbed(x_adv, 1, 0.0), target_feat)
        feats = self.clip.encode_perturbed(x_adv, Q, self.sigma_q, seed=self.seed, stream_id=step_idx)
        scores = L.cosine_rows(feats, target_feat)
        # the same directions again (same Philox keys), this time as NCHW fp32 for the estimate (0 + 1.0 * u_b)
        u = L.noise_image(torch.zeros_like(x_adv), Q, 1.0, seed=self.seed, stream_id=step_idx)
        wgt = ((scores.double() - f0.double()) / self.sigma_q).float()
        g = (wgt.view(Q, 1, 1, 1) * u).mean(0)
        x_new = x_adv + self.step_size * torch.sign(g)
        x_new = torch.max(torch.min(x_new, x_clean + self.eps), x_clean - self.eps).clamp_(0.0, 1.0)
        return x_new.contiguous(), f0, scores

    @torch.no_grad()
    def run(self, x_clean, target_image):
        """x_clean, target_image: [3,S,S] fp32 in [0,1].  Returns the adversarial image and a per-step log."""
        x_clean = x_clean.to(self.clip.dev, torch.float32).contiguous()
        target_feat = self.clip.encode_perturbed(target_image.to(self.clip.dev, torch.float32).contiguous(), 1, 0.0)[0].clone()
        x_adv, log = x_clean.clone(), []
        for s in range(self.steps):
            x_adv, f0, scores = self.step(x_adv, x_clean, target_feat, s)
            rec = {"step": s, "score_before": float(f0.item()), "query_score_mean": float(scores.mean().item())}
            if self.smooth is not None:
                self.smooth.image_id = 1000 + s
                rec["smoothed_prediction"] = self.smooth.predict(x_adv, self.predict_n, self.predict_alpha, self.predict_batch)
            log.append(rec)
        final = L.cosine_rows(self.clip.encode_perturbed(x_adv, 1, 0.0), target_feat)
        return x_adv, {"steps": log, "final_score": float(final.item()),
                       "linf": float((x_adv - x_clean).abs().max().item())}
