"""CPU ORACLE (test infrastructure, not product) for the HiFi-GAN V1 generator - the step immediately AFTER the hot path
(mel -> waveform; Grad-TTS/inference.py:81 `vocoder.forward(y_dec)`).

A functional, state_dict-driven restatement of Grad-TTS/hifi-gan/models.py:77-128 (Generator) and :13-49 (ResBlock1) in
plain PyTorch CPU fp32 ops over the EFFECTIVE weights, i.e. after `remove_weight_norm()` (inference.py:63), which is how
the reference runs it.  Only tests/ may import this file.  No product kernel exists for this row yet: the oracle, its
parameter inventory and its goldens are groundwork for the next round.

Pinned: scripts/make_golden_hifigan.py builds the UNMODIFIED reference Generator (imported from /root/reference) from
Grad-TTS/checkpts/hifigan-config.json with seeded random weights, removes weight norm, and asserts this file reproduces
its output before writing tests/golden/hifigan_golden.pt; tests/test_oracle_hifigan.py re-checks on every CPU run.
Paths below are relative to /root/reference/Grad-TTS/hifi-gan/.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.1                      # models.py:10

# checkpts/hifigan-config.json (V1)
V1 = dict(upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4], upsample_initial_channel=512,
          resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]], num_mels=80)


def get_padding(kernel_size, dilation=1):
    """xutils.py:36-37."""
    return int((kernel_size * dilation - dilation) / 2)


def param_spec(h=V1):
    """[(name, shape)] of the generator's state_dict after remove_weight_norm (models.py:77-101)."""
    spec = [("conv_pre.weight", (h["upsample_initial_channel"], h["num_mels"], 7)), ("conv_pre.bias", (h["upsample_initial_channel"],))]
    ch = h["upsample_initial_channel"]
    for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        cin, cout = h["upsample_initial_channel"] // 2 ** i, h["upsample_initial_channel"] // 2 ** (i + 1)
        spec += [(f"ups.{i}.weight", (cin, cout, k)), (f"ups.{i}.bias", (cout,))]          # ConvTranspose1d: [in, out, k]
    n = 0
    for i in range(len(h["upsample_rates"])):
        ch = h["upsample_initial_channel"] // 2 ** (i + 1)
        for k, d in zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"]):
            for grp in ("convs1", "convs2"):
                for j in range(len(d)):
                    spec += [(f"resblocks.{n}.{grp}.{j}.weight", (ch, ch, k)), (f"resblocks.{n}.{grp}.{j}.bias", (ch,))]
            n += 1
    spec += [("conv_post.weight", (1, ch, 7)), ("conv_post.bias", (1,))]
    return spec


def resblock1(p, pre, x, k, dilations):
    """models.py:41-48: x += conv2(lrelu(conv1_dilated(lrelu(x)))) for each dilation."""
    for j, d in enumerate(dilations):
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = F.conv1d(xt, p[f"{pre}.convs1.{j}.weight"], p[f"{pre}.convs1.{j}.bias"], padding=get_padding(k, d), dilation=d)
        xt = F.leaky_relu(xt, LRELU_SLOPE)
        xt = F.conv1d(xt, p[f"{pre}.convs2.{j}.weight"], p[f"{pre}.convs2.{j}.bias"], padding=get_padding(k, 1))
        x = xt + x
    return x


def generator(p, x, h=V1):
    """models.py:104-119: mel [B,80,T] -> waveform [B,1,T*prod(upsample_rates)]."""
    nk = len(h["resblock_kernel_sizes"])
    x = F.conv1d(x, p["conv_pre.weight"], p["conv_pre.bias"], padding=3)
    for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv_transpose1d(x, p[f"ups.{i}.weight"], p[f"ups.{i}.bias"], stride=u, padding=(k - u) // 2)
        xs = None
        for j in range(nk):
            r = resblock1(p, f"resblocks.{i * nk + j}", x, h["resblock_kernel_sizes"][j], h["resblock_dilation_sizes"][j])
            xs = r if xs is None else xs + r
        x = xs / nk
    x = F.leaky_relu(x)                       # models.py:115: default slope 0.01 here, not LRELU_SLOPE
    x = F.conv1d(x, p["conv_post.weight"], p["conv_post.bias"], padding=3)
    return torch.tanh(x)


def macs_per_mel_frame(h=V1):
    """Algorithmic multiply-accumulates per input mel frame (for the next round's roofline): every conv's
    out_channels * in_channels * kernel * (output samples per frame)."""
    total = h["upsample_initial_channel"] * h["num_mels"] * 7
    rate = 1
    for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        cin, cout = h["upsample_initial_channel"] // 2 ** i, h["upsample_initial_channel"] // 2 ** (i + 1)
        total += cin * cout * k * rate                   # transposed conv: each INPUT sample feeds k taps of cout channels
        rate *= u
        for kk, d in zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"]):
            total += 2 * len(d) * cout * cout * kk * rate
    total += cout * 7 * rate
    return total
