"""CPU ORACLE, second opinion (test infrastructure, not product): an INDEPENDENT float64 numpy restatement of the Grad-TTS
score network - no torch operators, every convolution written as explicit shifted-slice sums - so that the pinned torch
oracle (oracle/gradtts_oracle.py, bit-identical to the reference because it calls the same ATen ops) is cross-checked by
arithmetic that shares nothing with ATen: padding, tap order, the ConvTranspose2d phase rule, GroupNorm's biased variance and
eps, Mish, the softmax axis of LinearAttention, the sinusoid frequencies and the Euler step are all spelled out.
Small cases only (pure numpy).  Reference lines as in gradtts_oracle.py (Grad-TTS/model/diffusion.py)."""
from __future__ import annotations

import math

import numpy as np

HEADS, GROUPS = 4, 8


def _np(sd):
    return {k: v.detach().cpu().numpy().astype(np.float64) for k, v in sd.items()}


def mish(x):                                       # :16-18  x * tanh(softplus(x)), softplus threshold 20
    sp = np.where(x > 20.0, x, np.log1p(np.exp(np.minimum(x, 20.0))))
    return x * np.tanh(sp)


def conv2d(x, w, b=None, stride=1, pad=0):
    """x [B,Ci,H,W], w [Co,Ci,kh,kw]: out[b,o,i,j] = sum_{c,r,s} x[b,c,i*stride+r-pad,j*stride+s-pad] * w[o,c,r,s]."""
    B, Ci, H, W = x.shape
    Co, _, kh, kw = w.shape
    xp = np.zeros((B, Ci, H + 2 * pad, W + 2 * pad))
    xp[:, :, pad:pad + H, pad:pad + W] = x
    Ho, Wo = (H + 2 * pad - kh) // stride + 1, (W + 2 * pad - kw) // stride + 1
    out = np.zeros((B, Co, Ho, Wo))
    for r in range(kh):
        for s in range(kw):
            patch = xp[:, :, r:r + (Ho - 1) * stride + 1:stride, s:s + (Wo - 1) * stride + 1:stride]
            out += np.einsum("bchw,oc->bohw", patch, w[:, :, r, s])
    return out if b is None else out + b[None, :, None, None]


def conv_transpose2d_k4s2p1(x, w, b):
    """ConvTranspose2d(4, 2, 1), w [Ci,Co,4,4] (:24): out[b,o,2i-1+r,2j-1+s] += x[b,c,i,j] * w[c,o,r,s]."""
    B, Ci, H, W = x.shape
    Co = w.shape[1]
    full = np.zeros((B, Co, 2 * H + 2, 2 * W + 2))          # index = 2i + r  (before removing the padding of 1)
    for r in range(4):
        for s in range(4):
            full[:, :, r:r + 2 * H:2, s:s + 2 * W:2] += np.einsum("bchw,co->bohw", x, w[:, :, r, s])
    return full[:, :, 1:2 * H + 1, 1:2 * W + 1] + b[None, :, None, None]


def group_norm(x, gamma, beta, eps=1e-5):            # :53, biased variance over (C/8, H, W)
    B, C, H, W = x.shape
    g = x.reshape(B, GROUPS, -1)
    m = g.mean(-1, keepdims=True)
    v = ((g - m) ** 2).mean(-1, keepdims=True)
    return ((g - m) / np.sqrt(v + eps)).reshape(B, C, H, W) * gamma[None, :, None, None] + beta[None, :, None, None]


def block(p, pre, x, mask):                          # :56-58
    y = conv2d(x * mask, p[f"{pre}.block.0.weight"], p[f"{pre}.block.0.bias"], pad=1)
    return mish(group_norm(y, p[f"{pre}.block.1.weight"], p[f"{pre}.block.1.bias"])) * mask


def resnet(p, pre, x, mask, temb):                   # :74-79
    h = block(p, f"{pre}.block1", x, mask)
    h = h + (mish(temb) @ p[f"{pre}.mlp.1.weight"].T + p[f"{pre}.mlp.1.bias"])[:, :, None, None]
    h = block(p, f"{pre}.block2", h, mask)
    xm = x * mask
    res = conv2d(xm, p[f"{pre}.res_conv.weight"], p[f"{pre}.res_conv.bias"]) if f"{pre}.res_conv.weight" in p else xm
    return h + res


def attention(p, pre, x):                            # :39-46, 82-110
    B, C, H, W = x.shape
    qkv = conv2d(x, p[f"{pre}.fn.fn.to_qkv.weight"]).reshape(B, 3, HEADS, 32, H * W)
    q, k, v = qkv[:, 0], qkv[:, 1], qkv[:, 2]
    k = np.exp(k - k.max(-1, keepdims=True))
    k = k / k.sum(-1, keepdims=True)                 # softmax over ALL H*W positions, no mask
    ctx = np.einsum("bhdn,bhen->bhde", k, v)
    out = np.einsum("bhde,bhdn->bhen", ctx, q).reshape(B, HEADS * 32, H, W)
    out = conv2d(out, p[f"{pre}.fn.fn.to_out.weight"], p[f"{pre}.fn.fn.to_out.bias"])
    return out * p[f"{pre}.fn.g"] + x


def estimator(sd, cfg, x, mask, mu, t, spk=None):
    """:174-216 for n_spks == 1.  x, mu [B,80,T], mask [B,1,T], t [B] (torch tensors or arrays) -> float64 array [B,80,T]."""
    assert cfg.n_spks < 2, "single-speaker cross-check only"
    p = sd if isinstance(next(iter(sd.values())), np.ndarray) else _np(sd)
    x, mask, mu, t = (np.asarray(a, dtype=np.float64) for a in (x, mask, mu, t))
    half = cfg.dim // 2
    freqs = np.exp(np.arange(half) * -(math.log(10000) / (half - 1)))            # :118-125
    e = cfg.pe_scale * t[:, None] * freqs[None, :]
    temb = np.concatenate([np.sin(e), np.cos(e)], -1)
    temb = temb @ p["estimator.mlp.0.weight"].T + p["estimator.mlp.0.bias"]
    temb = mish(temb) @ p["estimator.mlp.2.weight"].T + p["estimator.mlp.2.bias"]
    h = np.stack([mu, x], 1)
    m = mask[:, None]
    pre, skips, masks = "estimator", [], [m]
    for l in range(3):
        mk = masks[-1]
        h = resnet(p, f"{pre}.downs.{l}.0", h, mk, temb)
        h = resnet(p, f"{pre}.downs.{l}.1", h, mk, temb)
        h = attention(p, f"{pre}.downs.{l}.2", h)
        skips.append(h)
        h = conv2d(h * mk, p[f"{pre}.downs.{l}.3.conv.weight"], p[f"{pre}.downs.{l}.3.conv.bias"], stride=2, pad=1) if l < 2 else h * mk
        masks.append(mk[:, :, :, ::2])
    masks = masks[:-1]
    mk = masks[-1]
    h = resnet(p, f"{pre}.mid_block1", h, mk, temb)
    h = attention(p, f"{pre}.mid_attn", h)
    h = resnet(p, f"{pre}.mid_block2", h, mk, temb)
    for j in range(2):
        mk = masks.pop()
        h = np.concatenate([h, skips.pop()], 1)
        h = resnet(p, f"{pre}.ups.{j}.0", h, mk, temb)
        h = resnet(p, f"{pre}.ups.{j}.1", h, mk, temb)
        h = attention(p, f"{pre}.ups.{j}.2", h)
        h = conv_transpose2d_k4s2p1(h * mk, p[f"{pre}.ups.{j}.3.conv.weight"], p[f"{pre}.ups.{j}.3.conv.bias"])
    h = block(p, f"{pre}.final_block", h, m)
    out = conv2d(h * m, p[f"{pre}.final_conv.weight"], p[f"{pre}.final_conv.bias"])
    return (out * m)[:, 0]


def reverse_diffusion(sd, cfg, z, mask, mu, n_timesteps):
    """:254-275, deterministic branch: xt <- (xt - 0.5 (mu - xt - est) beta_t h) * mask, t = 1 - (i + 0.5) h."""
    p = _np(sd)
    z, mask, mu = (np.asarray(a, dtype=np.float64) for a in (z, mask, mu))
    h = 1.0 / n_timesteps
    xt = z * mask
    for i in range(n_timesteps):
        t = (1.0 - (i + 0.5) * h) * np.ones(z.shape[0])
        beta = cfg.beta_min + (cfg.beta_max - cfg.beta_min) * t[:, None, None]
        dxt = 0.5 * (mu - xt - estimator(p, cfg, xt, mask, mu, t)) * beta * h
        xt = (xt - dxt) * mask
    return xt
