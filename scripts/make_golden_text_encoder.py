"""Generate tests/golden/text_encoder_golden.pt from the UNMODIFIED reference TextEncoder (container only):
TextEncoder(149, 80, 192, 768, 256, 2, 6, 3, 0.1, 4) (Grad-TTS/params.py) in eval mode, strict-loaded with seeded synthetic
weights (oracle/text_encoder_oracle.py:param_spec must equal its state_dict), on seeded token ids with ragged lengths.
Stores only the reference outputs; asserts the oracle reproduces them.

    python scripts/make_golden_text_encoder.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from oracle import text_encoder_oracle as T  # noqa: E402
from _ref_import import import_gradtts  # noqa: E402

SEED = 9753
CASES = [dict(B=2, Tx=37, lengths=[37, 20]), dict(B=1, Tx=221, lengths=[221]), dict(B=3, Tx=6, lengths=[6, 1, 3])]


def main():
    import_gradtts()
    from model.text_encoder import TextEncoder
    ref = TextEncoder(149, 80, 192, 768, 256, 2, 6, 3, 0.1, 4).eval()
    assert {k: tuple(v.shape) for k, v in ref.state_dict().items()} == dict(T.param_spec()), "inventory differs"
    sd = T.synthetic_weights(SEED)
    ref.load_state_dict(sd, strict=True)
    out = {"seed": SEED, "torch": torch.__version__, "cases": []}
    for c in CASES:
        g = torch.Generator().manual_seed(SEED + c["Tx"])
        x = torch.randint(0, 148, (c["B"], c["Tx"]), generator=g)
        xl = torch.tensor(c["lengths"])
        with torch.no_grad():
            mu, logw, mask = ref(x, xl)
            mu_o, logw_o, mask_o = T.text_encoder(sd, x, xl)
        e1 = ((mu_o - mu).norm() / mu.norm()).item()
        e2 = ((logw_o - logw).norm() / logw.norm()).item()
        assert torch.equal(mask_o, mask) and e1 < 2e-6 and e2 < 2e-6, (e1, e2)
        out["cases"].append(dict(c, mu=mu.clone(), logw=logw.clone()))
        print(f"B={c['B']} Tx={c['Tx']}: oracle vs reference rel-L2 mu {e1:.1e} logw {e2:.1e}")
    path = os.path.join(ROOT, "tests", "golden", "text_encoder_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
