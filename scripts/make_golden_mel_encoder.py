"""Generate tests/golden/mel_encoder_golden.pt from the UNMODIFIED DiffVC MelEncoder (container only: needs /root/reference).

Builds `MelEncoder(80, 192, 768, 2, 6, 3, 0.1, window_size=4)` (DiffVC/params.py:16-22, DiffVC/model/vc.py:32) from the
reference tree, loads seeded weights strictly, asserts oracle/text_encoder_oracle.py:mel_encoder reproduces it, and stores
ONLY the reference outputs; tests rebuild weights and inputs from the seeds.

    python scripts/make_golden_mel_encoder.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import, text_encoder_oracle as T  # noqa: E402

CASES = [dict(B=2, T=40, lengths=[40, 23]), dict(B=1, T=256, lengths=[256]), dict(B=3, T=9, lengths=[9, 1, 5])]
SEED = 9753


def main():
    ref_import.import_model("diffvc")
    import model.encoder as renc
    ref = renc.MelEncoder(80, 192, 768, 2, 6, 3, 0.1, window_size=4).eval()
    sd = T.mel_synthetic_weights(SEED)
    assert {k: tuple(v.shape) for k, v in ref.state_dict().items()} == dict(T.mel_param_spec())
    ref.load_state_dict(sd, strict=True)
    out = {"seed": SEED, "torch": torch.__version__, "nparams": sum(v.numel() for v in sd.values()), "cases": []}
    for c in CASES:
        g = torch.Generator().manual_seed(SEED + c["T"])
        x = torch.randn(c["B"], 80, c["T"], generator=g)
        mask = (torch.arange(c["T"])[None, :] < torch.tensor(c["lengths"])[:, None]).float()[:, None]
        with torch.no_grad():
            y = ref(x, mask)
            yo = T.mel_encoder(sd, x, mask)
        err = (y - yo).abs().max().item()
        print(c, "oracle vs reference max abs", err)
        assert err <= 1e-5
        out["cases"].append(dict(c, out=y))
    torch.save(out, os.path.join(ROOT, "tests", "golden", "mel_encoder_golden.pt"))
    print("wrote tests/golden/mel_encoder_golden.pt", out["nparams"], "parameters")


if __name__ == "__main__":
    main()
