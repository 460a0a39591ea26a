"""Shared helpers for the parity tests: rebuild a golden case's weights/inputs from its seeds."""
import torch

from speech_backbones_b200 import UNetConfig, synthetic_inputs, synthetic_state_dict, synthetic_noise  # noqa: F401


def case_id(c):
    keys = ["kind", "n_spks", "B", "T", "ragged"] + (["t", "scale"] if c["kind"] == "est" else ["N", "stoc"])
    return "-".join(f"{k}{c[k]}" for k in keys)


def case_inputs(golden, c):
    cfg = UNetConfig(n_spks=c["n_spks"])
    sd = synthetic_state_dict(cfg, golden["seed"])
    z, mask, mu, spk, lengths = synthetic_inputs(c["B"], c["T"], seed=golden["seed"], ragged=c["ragged"],
                                                 n_spks=cfg.n_spks)
    return cfg, sd, z, mask, mu, spk


def stoc_noise(golden, c):
    """The noise the reference drew: torch.manual_seed(noise_seed), then randn(z.shape) per step (diffusion.py:267)."""
    torch.manual_seed(golden["noise_seed"])
    return torch.stack([torch.randn(c["B"], 80, c["T"]) for _ in range(c["N"])])


def rel_l2(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def nhwc_to_nchw(flat, B, H, W, C, layout=0):
    """layout 0: [B][H][W][C]; layout 1: [B][H][C/4][W][4] (tensor-core modes); layout 2: [B][H][C/8][W][8] (bf16 operands)."""
    if layout in (1, 2):
        e = 4 if layout == 1 else 8
        return flat.view(B, H, C // e, W, e).permute(0, 2, 4, 1, 3).reshape(B, C, H, W).contiguous()
    return flat.view(B, H, W, C).permute(0, 3, 1, 2).contiguous()
