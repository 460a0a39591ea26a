"""BASELINE config 1 in its end-to-end form: the reference `GradTTS.forward` (unmodified, CPU, random-init text
encoder, 221 token ids, N=10, temperature 1.5, length_scale 0.91) recorded by scripts/make_golden_config1.py, against
  CPU: the oracle chain  prior_expand -> reverse_diffusion  (bit-level: same torch build),
  GPU: the product chain  synthesize_from_encoder = sbk_prior_expand + sbk_reverse_diffusion  fed with the reference's own
       encoder outputs and noise draws (the text encoder itself is outside the path and stays the reference's).
(Runs last in the GPU suite: it only recombines pieces the earlier files test one by one.)"""
import os

import pytest
import torch

from helpers import rel_l2
from oracle import gradtts_oracle as O
from speech_backbones_b200 import UNetConfig, synthetic_state_dict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def c1():
    return torch.load(os.path.join(ROOT, "tests", "golden", "gradtts_config1_golden.pt"), weights_only=False)


def _noise(c1):
    from speech_backbones_b200.gradtts import reference_order_noise
    torch.manual_seed(c1["noise_seed"])
    return reference_order_noise(1, 80, c1["Ty"], torch.float32, "cpu")


def test_oracle_chain_matches_reference_forward(c1):
    cfg = UNetConfig()
    sd = synthetic_state_dict(cfg, c1["seed"])
    o = O.prior_expand(c1["mu_x"], c1["logw"], c1["x_mask"], c1["length_scale"], c1["temperature"], _noise(c1))
    L = o["y_max_length"]
    assert L == c1["y_dec"].shape[-1] == 293 and o["mu_y"].shape[-1] == c1["Ty"] == 296
    assert torch.equal(o["mu_y"][:, :, :L], c1["y_enc"])
    full = o["attn"][0, 0]                        # [Tx, Ty]: the reference returns attn[:, :, :y_max_length], a no-op slice of the token axis
    tok = torch.where(full.sum(0) > 0, full.argmax(0), torch.full((full.shape[1],), -1)).to(torch.int16)
    assert torch.equal(tok, c1["tok"]) and (tok[L:] == -1).all()
    with torch.no_grad():
        y = O.reverse_diffusion(sd, cfg, o["z"], o["y_mask"], o["mu_y"], c1["N"])[:, :, :L]
    assert torch.allclose(y, c1["y_dec"], rtol=1e-5, atol=1e-5 * c1["y_dec"].abs().max().item())


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("fp32", 2e-3), ("tf32", 8e-3), ("bf16", 1e-2)])
def test_product_chain_matches_reference_forward(sbk_lib, c1, precision, tol):
    from speech_backbones_b200 import gradtts as G
    dec = G.Diffusion(80, 64, precision=precision).eval()
    dec.load_state_dict(synthetic_state_dict(UNetConfig(), c1["seed"]))
    dec = dec.cuda()
    enc, y, attn = G.synthesize_from_encoder(dec, c1["mu_x"].cuda(), c1["logw"].cuda(), c1["x_mask"].cuda(), c1["N"],
                                             c1["temperature"], False, None, c1["length_scale"], True, _noise(c1).cuda())
    L = c1["y_dec"].shape[-1]
    assert enc.shape == (1, 80, L) and torch.equal(enc.cpu(), c1["y_enc"])            # the gather is exact
    full = attn[0, 0].cpu()
    tok = torch.where(full.sum(0) > 0, full.argmax(0), torch.full((full.shape[1],), -1)).to(torch.int16)
    assert torch.equal(tok, c1["tok"])
    err = rel_l2(y.cpu(), c1["y_dec"])
    print(precision, "config 1 end to end: rel_l2 vs reference GradTTS.forward =", err)
    assert y.shape == (1, 80, L) and err <= tol
