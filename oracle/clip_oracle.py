"""TEST INFRASTRUCTURE ONLY (imported by tests/, never by certifiedgpt_b200/).

CPU fp32 restatement of the CLIP ViT vision tower with projection, and of the black-box attack inner loop of
BASELINE.json configs[4] ("8-step perturbation update with CLIP ViT-L/14 feature cosine scoring plus smoothed
predict at N=100 per step").

The reference repository ships NO code for this stage: it is prose and a result table only
(/root/reference/README.md:62-64,108-120, "AttackVLM" black-box attack).  Parity is therefore
pinned against the third-party model the prose names: the algorithm of `transformers.CLIPVisionModelWithProjection`
(transformers 5.5.0 in this image; modeling_clip.py: CLIPVisionEmbeddings, pre_layrnorm, CLIPEncoderLayer with
quick_gelu, post_layernorm on the class token, visual_projection), restated here in plain torch and checked
against that implementation in tests/test_oracle_clip_cpu.py.  The attack update is OUR definition (random
gradient-free estimate, sign step, eps-ball projection) because the reference defines none: parity unpinned by
reference tests, pinned by this restatement.
"""
import torch
import torch.nn.functional as F


def clip_vision_features(sd, cfg, pixel_values):
    """pixel_values [B,3,S,S] (CLIP-normalised) -> image_embeds [B, proj] fp32 (not L2-normalised).
    sd: HF state_dict of CLIPVisionModelWithProjection; cfg: has hidden, heads, layers, patch, eps."""
    p = "vision_model."
    x = F.conv2d(pixel_values.float(), sd[p + "embeddings.patch_embedding.weight"].float(), stride=cfg.patch)
    B, D = x.shape[0], x.shape[1]
    x = x.flatten(2).transpose(1, 2)                                   # [B, G*G, D]
    cls = sd[p + "embeddings.class_embedding"].float().expand(B, 1, D)
    x = torch.cat([cls, x], dim=1) + sd[p + "embeddings.position_embedding.weight"].float()[None]
    x = F.layer_norm(x, (D,), sd[p + "pre_layrnorm.weight"].float(), sd[p + "pre_layrnorm.bias"].float(), cfg.eps)
    H = cfg.heads
    hd = D // H
    for i in range(cfg.layers):
        l = f"{p}encoder.layers.{i}."
        h = F.layer_norm(x, (D,), sd[l + "layer_norm1.weight"].float(), sd[l + "layer_norm1.bias"].float(), cfg.eps)
        q = F.linear(h, sd[l + "self_attn.q_proj.weight"].float(), sd[l + "self_attn.q_proj.bias"].float())
        k = F.linear(h, sd[l + "self_attn.k_proj.weight"].float(), sd[l + "self_attn.k_proj.bias"].float())
        v = F.linear(h, sd[l + "self_attn.v_proj.weight"].float(), sd[l + "self_attn.v_proj.bias"].float())
        T = x.shape[1]
        q, k, v = (t.view(B, T, H, hd).transpose(1, 2) for t in (q, k, v))
        a = ((q @ k.transpose(-1, -2)) * hd ** -0.5).softmax(-1) @ v
        a = a.transpose(1, 2).reshape(B, T, D)
        x = x + F.linear(a, sd[l + "self_attn.out_proj.weight"].float(), sd[l + "self_attn.out_proj.bias"].float())
        h = F.layer_norm(x, (D,), sd[l + "layer_norm2.weight"].float(), sd[l + "layer_norm2.bias"].float(), cfg.eps)
        h = F.linear(h, sd[l + "mlp.fc1.weight"].float(), sd[l + "mlp.fc1.bias"].float())
        h = h * torch.sigmoid(1.702 * h)                               # quick_gelu
        x = x + F.linear(h, sd[l + "mlp.fc2.weight"].float(), sd[l + "mlp.fc2.bias"].float())
    pooled = F.layer_norm(x[:, 0], (D,), sd[p + "post_layernorm.weight"].float(), sd[p + "post_layernorm.bias"].float(),
                          cfg.eps)
    return F.linear(pooled, sd["visual_projection.weight"].float())


def cosine_scores(feats, target):
    return F.cosine_similarity(feats.float(), target.float()[None], dim=1, eps=1e-12)


def rgf_step(x_adv, x_clean, directions, f0, scores, sigma_q, step_size, eps):
    """One perturbation update from Q random directions u_b [Q,3,S,S] (pixel space):
    g = mean_b ((score_b - f0) / sigma_q) * u_b ;  x <- clip(x + step_size * sign(g)) to the eps-ball and [0,1]."""
    w = ((scores.double() - float(f0)) / sigma_q).float()
    g = (w.view(-1, 1, 1, 1) * directions.float()).mean(0)
    x = x_adv + step_size * torch.sign(g)
    x = torch.max(torch.min(x, x_clean + eps), x_clean - eps)
    return x.clamp(0.0, 1.0), g
