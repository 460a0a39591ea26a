"""Generate tests/golden/hifigan_golden.pt from the UNMODIFIED reference HiFi-GAN generator (container only).

Imports Grad-TTS/hifi-gan/models.py from /root/reference (matplotlib, which xutils.py imports for plotting only, is
stubbed), builds Generator(h) from Grad-TTS/checkpts/hifigan-config.json, re-initialises every weight from a seeded
generator (the reference ships no vocoder checkpoint; init_weights' std 0.01 would make the output vanish), calls
remove_weight_norm() as inference.py:63 does, and stores ONLY the reference outputs; tests rebuild weights and inputs from
the seeds.  Asserts oracle/hifigan_oracle.py == reference on every case and that the oracle's parameter inventory is the
reference's state_dict (names and shapes).

    python scripts/make_golden_hifigan.py
"""
import json
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import hifigan_oracle as H  # noqa: E402
from speech_backbones_b200.spec import hifigan_param_spec, synthetic_hifigan_state_dict  # noqa: E402

CASES = [dict(B=1, T=32), dict(B=2, T=20), dict(B=1, T=5)]
SEED = 2468


def import_reference_generator():
    for n in ("matplotlib", "matplotlib.pylab"):
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["matplotlib"].use = lambda *a, **k: None
    sys.modules["matplotlib"].pylab = sys.modules["matplotlib.pylab"]
    sys.path.insert(0, "/root/reference/Grad-TTS/hifi-gan")
    from env import AttrDict
    from models import Generator
    with open("/root/reference/Grad-TTS/checkpts/hifigan-config.json") as f:
        h = AttrDict(json.load(f))
    return Generator, h


def main():
    Generator, h = import_reference_generator()
    for k in ("upsample_rates", "upsample_kernel_sizes", "upsample_initial_channel", "resblock_kernel_sizes", "resblock_dilation_sizes"):
        assert h[k] == H.V1[k], k
    ref = Generator(h).eval()
    ref.remove_weight_norm()
    sd = synthetic_hifigan_state_dict(SEED)
    ref_shapes = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    assert ref_shapes == dict(H.param_spec()) == dict(hifigan_param_spec()), "parameter inventory differs from the reference's state_dict"
    ref.load_state_dict(sd, strict=True)
    out = {"seed": SEED, "torch": torch.__version__, "cases": [], "nparams": sum(v.numel() for v in sd.values()),
           "macs_per_mel_frame": H.macs_per_mel_frame()}
    for c in CASES:
        g = torch.Generator().manual_seed(SEED + c["T"])
        mel = torch.randn(c["B"], 80, c["T"], generator=g)
        with torch.no_grad():
            y = ref(mel)
            yo = H.generator(sd, mel)
        assert y.shape == (c["B"], 1, c["T"] * 256)
        err = (yo - y).abs().max().item()
        assert err == 0.0, err
        out["cases"].append(dict(c, out=y.clone()))
        print(f"B={c['B']} T={c['T']}: |y|max={y.abs().max():.3f}, oracle == reference (max abs diff {err})")
    path = os.path.join(ROOT, "tests", "golden", "hifigan_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes; params", out["nparams"], "MAC/frame", out["macs_per_mel_frame"])


if __name__ == "__main__":
    main()
