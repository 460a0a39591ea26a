"""Generate tests/golden/diffvc_glue_golden.pt from the UNMODIFIED reference (container only): the lines of DiffVC.forward
between the encoders and the decoder (DiffVC/model/vc.py:104-127).  The reference `DiffVC` is imported from /root/reference,
its mel encoder is replaced by a stub returning seeded synthetic `mean` / `mean_ref` and its decoder by a stub that records
the (z, x_mask_new, mean_new) it is handed (compute_diffused_mean stays the reference's); everything in between runs as
shipped.  Asserts oracle/diffvc_oracle.py:prepare_decoder_inputs reproduces every tensor bit for bit.

    python scripts/make_golden_vc_glue.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from oracle import diffvc_oracle as O  # noqa: E402
from speech_backbones_b200.spec import DiffVCConfig  # noqa: E402
from _ref_import import import_diffvc  # noqa: E402

CASES = [dict(B=3, lengths=[37, 22, 30]), dict(B=1, lengths=[64]), dict(B=4, lengths=[5, 9, 1, 7])]
SEED, NOISE_SEED = 1357, 13


def synth(c):
    g = torch.Generator().manual_seed(SEED + sum(c["lengths"]))
    T = max(c["lengths"])
    x = torch.randn(c["B"], 80, T, generator=g)
    mean = torch.randn(c["B"], 80, T, generator=g)
    return x, torch.tensor(c["lengths"]), mean


def main():
    md = import_diffvc()
    from model import DiffVC
    cfg = DiffVCConfig()
    m = DiffVC(80, 192, 768, 2, 6, 3, 0.1, 4, 128, 128, True, 256, 0.05, 20.0).eval()
    real_dec = m.decoder
    out = {"seed": SEED, "noise_seed": NOISE_SEED, "torch": torch.__version__, "cases": []}
    for c in CASES:
        x, x_lengths, mean = synth(c)

        class Enc(torch.nn.Module):
            def forward(self, xx, mask):
                return mean if xx.shape == mean.shape else torch.zeros_like(xx)

        class Dec(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.anchor = torch.nn.Parameter(torch.zeros(1))
                self.cap = None

            def compute_diffused_mean(self, *a, **k):
                return real_dec.compute_diffused_mean(*a, **k)

            def forward(self, z, mask, mean_, ref, ref_mask, mean_ref, cc, n, mode):
                self.cap = dict(z=z.clone(), mask=mask.clone(), mean=mean_.clone())
                return torch.zeros_like(z)

        m.encoder, m.decoder = Enc(), Dec()
        torch.manual_seed(NOISE_SEED)
        ref_x = torch.zeros(c["B"], 80, 8)
        mean_x, y = m(x, x_lengths, ref_x, torch.full((c["B"],), 8), torch.zeros(c["B"], 256), n_timesteps=1, mode="ml")
        cap = m.decoder.cap
        torch.manual_seed(NOISE_SEED)
        noise = torch.randn(cap["z"].shape)
        o = O.prepare_decoder_inputs(cfg, x, x_lengths, mean, noise)
        assert torch.equal(o["z"], cap["z"]) and torch.equal(o["x_mask_new"], cap["mask"]) and torch.equal(o["mean_new"], cap["mean"])
        assert torch.equal(o["mean_x"], mean_x) and y.shape[-1] == o["max_length"]
        out["cases"].append(dict(c, z=cap["z"], mask=cap["mask"], mean_new=cap["mean"], mean_x=mean_x.clone()))
        print(f"lengths={c['lengths']}: T'={cap['z'].shape[-1]}  oracle == reference")
    path = os.path.join(ROOT, "tests", "golden", "diffvc_glue_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
