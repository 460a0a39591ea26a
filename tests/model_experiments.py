"""CPU-only what-if experiments with the operand-rounding model (oracle/precision_model.py); not collected by pytest.
They back the round-2 numbers quoted in DESIGN.md section 11.   python tests/model_experiments.py [longN]

  1. bf16 mode with Mish (a) in fp32 (shipped), (b) exp + reciprocal in bf16, (c) entirely in bf16 arithmetic
  2. bf16 mode with the raw conv outputs (GroupNorm inputs) stored as bf16 too
  3. an fp32-accurate tensor-core mode: 3xTF32 (hi/lo splits of both operands, three products, fp32 accumulation)
  4. [longN] error vs number of reverse steps, N = 50 / 200 / 1000 (several minutes)
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import case_inputs, rel_l2  # noqa: E402
from oracle import gradtts_oracle as O  # noqa: E402
from oracle.precision_model import _Shim, operand_rounding, round_bf16, round_tf32_rna  # noqa: E402

golden = torch.load(os.path.join(ROOT, "tests", "golden", "gradtts_golden.pt"), weights_only=False)
CASES = [c for c in golden["cases"] if c["kind"] == "est" and c["n_spks"] == 1 and c["scale"] == 1.0]


def run(mode, before=None):
    errs = []
    for c in CASES:
        cfg, sd, z, mask, mu, spk = case_inputs(golden, c)
        with operand_rounding(mode, sd), torch.no_grad():
            if before:
                before()
            errs.append(rel_l2(O.estimator(sd, cfg, z * mask, mask, mu, torch.tensor(c["t"]), spk), c["out"]))
    return " ".join(f"{e:.3e}" for e in errs)


def mish_mufu_bf16(x):
    n = torch.exp(torch.clamp(x, max=20.0).bfloat16()).float()
    a = n * (n + 2)
    return x * (a / (a + 2)).bfloat16().float()


def mish_all_bf16(x):
    x16 = x.bfloat16()
    n = torch.exp(torch.clamp(x16, max=20.0))
    a = n * (n + 2)
    return (x16 * (a / (a + 2))).float()


orig_mish = O.mish
print("1a bf16 mode, fp32 Mish (shipped)      ", run("bf16"))
for label, fn in (("1b bf16 mode, exp + rcp in bf16        ", mish_mufu_bf16), ("1c bf16 mode, Mish entirely in bf16    ", mish_all_bf16)):
    O.mish = fn
    try:
        print(label, run("bf16"))
    finally:
        O.mish = orig_mish


def raw_bf16():
    gn0 = O.F.group_norm
    O.F._over["group_norm"] = lambda x, *a, **k: gn0(round_bf16(x), *a, **k)


print("2  bf16 mode, raw conv outputs in bf16 ", run("bf16", raw_bf16))

F0 = O.F


def _split(x):
    hi = round_tf32_rna(x)
    return hi, round_tf32_rna(x - hi)


def _three(op, x, w, b, stride, padding, *a, **k):
    xh, xl = _split(x)
    wh, wl = _split(w)
    y = op(xh, wh, None, stride, padding, *a, **k) + op(xl, wh, None, stride, padding, *a, **k) + op(xh, wl, None, stride, padding, *a, **k)
    return y if b is None else y + b[None, :, None, None]


O.F = _Shim(F0, conv2d=lambda x, w, b=None, stride=1, padding=0, *a, **k: _three(F0.conv2d, x, w, b, stride, padding, *a, **k),
            conv_transpose2d=lambda x, w, b=None, stride=1, padding=0, *a, **k: _three(F0.conv_transpose2d, x, w, b, stride, padding, *a, **k))
try:
    errs = []
    for c in CASES:
        cfg, sd, z, mask, mu, spk = case_inputs(golden, c)
        with torch.no_grad():
            errs.append(rel_l2(O.estimator(sd, cfg, z * mask, mask, mu, torch.tensor(c["t"]), spk), c["out"]))
    print("3  3xTF32 convs, everything else fp32  ", " ".join(f"{e:.3e}" for e in errs))
finally:
    O.F = F0

if len(sys.argv) > 1 and sys.argv[1] == "longN":
    from speech_backbones_b200 import UNetConfig, synthetic_inputs, synthetic_state_dict
    cfg = UNetConfig()
    sd = synthetic_state_dict(cfg)
    z, mask, mu, _, _ = synthetic_inputs(2, 32, ragged=True)
    for n in (50, 200, 1000):
        t0 = time.time()
        with torch.no_grad():
            ref = O.reverse_diffusion(sd, cfg, z, mask, mu, n)
        row = {}
        for mode in ("tf32", "bf16"):
            with operand_rounding(mode, sd), torch.no_grad():
                row[mode] = rel_l2(O.reverse_diffusion(sd, cfg, z, mask, mu, n), ref)
        print(f"4  N={n}: tf32 {row['tf32']:.2e}  bf16 {row['bf16']:.2e}  ({time.time() - t0:.0f} s)", flush=True)
