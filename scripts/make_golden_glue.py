"""Generate tests/golden/gradtts_glue_golden.pt from the UNMODIFIED reference (container only): the lines of
GradTTS.forward between the text encoder and the decoder (Grad-TTS/model/tts.py:77-99, model/utils.py:6-39).

The reference `GradTTS` is imported from /root/reference, its text encoder is replaced by a stub that returns seeded
synthetic (mu_x, logw, x_mask) and its decoder by a stub that records the (z, y_mask, mu_y) it is called with; everything
in between runs as shipped.  Stored per case: the recorded z / mu_y / y_mask, y_lengths, and the alignment as one token
index per output frame (the script asserts that the reference's attn is exactly the one-hot matrix of those indices times
the masks, so nothing is lost).  Also asserts that oracle/gradtts_oracle.py:prior_expand reproduces every tensor bit for
bit, and that the product's noise draw (`gradtts.reference_order_noise`: randn_like on a tensor with the strides of the reference's
transposed mu_y) reproduces the reference's `randn_like(mu_y)` under the same seed.

    python scripts/make_golden_glue.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))

from oracle import gradtts_oracle as O  # noqa: E402
from speech_backbones_b200.spec import synthetic_encoder_outputs  # noqa: E402
from speech_backbones_b200.gradtts import reference_order_noise  # noqa: E402
from _ref_import import import_gradtts  # noqa: E402

CASES = [
    dict(B=3, Tx=40, x_lengths=[40, 25, 33], dur_mean=1.0, length_scale=0.91, temperature=1.5),
    dict(B=2, Tx=17, x_lengths=[17, 9], dur_mean=0.8, length_scale=1.0, temperature=1.0),
    dict(B=1, Tx=5, x_lengths=[5], dur_mean=-1.0, length_scale=0.3, temperature=2.0),
    dict(B=1, Tx=221, x_lengths=[221], dur_mean=0.9, length_scale=0.91, temperature=1.5),       # config 1's token count
    dict(B=4, Tx=12, x_lengths=[12, 1, 7, 3], dur_mean=0.5, length_scale=1.3, temperature=0.7),
]
SEED = 4321
NOISE_SEED = 11


class _Enc(torch.nn.Module):
    def __init__(self, out):
        super().__init__()
        self.out = out

    def forward(self, x, x_lengths, spk=None):
        return self.out


class _Dec(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.cap = None
        self.anchor = torch.nn.Parameter(torch.zeros(1))      # relocate_input (base.py:33) asks the module for its device

    def forward(self, z, mask, mu, n_timesteps, stoc=False, spk=None):
        self.cap = dict(z=z.clone(), mask=mask.clone(), mu=mu.clone())
        return torch.zeros_like(z)


def main():
    import_gradtts()
    from model import GradTTS
    model = GradTTS(149, 1, 64, 192, 768, 256, 2, 6, 3, 0.1, 4, 80, 64, 0.05, 20.0, 1000).eval()
    out = {"seed": SEED, "noise_seed": NOISE_SEED, "torch": torch.__version__, "cases": []}
    for c in CASES:
        mu_x, logw, x_mask = synthetic_encoder_outputs(c["B"], c["Tx"], c["x_lengths"], c["dur_mean"], seed=SEED)
        model.encoder, model.decoder = _Enc((mu_x, logw, x_mask)), _Dec()
        real_randn_like, drawn = torch.randn_like, []

        def rec(t, *a, **k):
            r = real_randn_like(t, *a, **k)
            drawn.append(r)
            return r
        torch.randn_like = rec
        try:
            torch.manual_seed(NOISE_SEED)
            x = torch.zeros((c["B"], c["Tx"]), dtype=torch.long)
            enc_out, dec_out, attn = model(x, torch.tensor(c["x_lengths"]), n_timesteps=1, temperature=c["temperature"],
                                           length_scale=c["length_scale"])
        finally:
            torch.randn_like = real_randn_like
        cap = model.decoder.cap
        z, y_mask, mu_y = cap["z"], cap["mask"], cap["mu"]
        B, Fm, Ty = z.shape
        # the product draws the same numbers as torch.randn(B, Ty, F) (memory order of the transposed mu_y)
        torch.manual_seed(NOISE_SEED)
        noise_tf = reference_order_noise(B, Fm, Ty, torch.float32, "cpu")
        assert len(drawn) == 1 and noise_tf.is_contiguous() and torch.equal(drawn[0], noise_tf.transpose(1, 2)), "randn_like(mu_y)"
        # oracle == reference, bit for bit
        o = O.prior_expand(mu_x, logw, x_mask, c["length_scale"], c["temperature"], noise_tf)
        assert torch.equal(o["z"], z) and torch.equal(o["mu_y"], mu_y) and torch.equal(o["y_mask"], y_mask)
        y_max = o["y_max_length"]
        assert torch.equal(o["attn"][:, :, :y_max], attn) and torch.equal(o["mu_y"][:, :, :y_max], enc_out)
        # compact alignment: one token per frame (-1 = none); assert it encodes the reference attn exactly
        full = o["attn"][:, 0]                                            # [B,Tx,Ty]
        assert ((full == 0) | (full == 1)).all() and (full.sum(1) <= 1).all()
        tok = torch.where(full.sum(1) > 0, full.argmax(1), torch.full((B, Ty), -1)).to(torch.int16)
        rebuilt = torch.zeros_like(full)
        for b in range(B):
            for t in range(Ty):
                if tok[b, t] >= 0:
                    rebuilt[b, int(tok[b, t]), t] = 1.0
        assert torch.equal(rebuilt, full)
        cc = dict(c)
        cc.update(Ty=Ty, y_max_length=y_max, y_lengths=o["y_lengths"].clone(), z=z.contiguous(), mu_y=mu_y.contiguous(),
                  y_mask=y_mask.contiguous(), tok=tok)
        out["cases"].append(cc)
        print(f"case B={c['B']} Tx={c['Tx']} ls={c['length_scale']}: y_lengths={o['y_lengths'].tolist()} Ty={Ty}  oracle == reference")
    path = os.path.join(ROOT, "tests", "golden", "gradtts_glue_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
