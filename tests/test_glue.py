"""The step before the path: GradTTS.forward between the text encoder and the decoder
(Grad-TTS/model/tts.py:77-99).  Fixtures: tests/golden/gradtts_glue_golden.pt, recorded from the UNMODIFIED reference by
scripts/make_golden_glue.py (z / mu_y / y_mask as handed to the decoder, y_lengths, the alignment as one token per frame).
CPU: the oracle restatement reproduces the fixtures bit for bit.  GPU: sbk_prior_expand (through the C ABI) and the
drop-in `synthesize_from_encoder` do too - this is integer / copy / one-add work, so the bar is torch.equal."""
import os

import pytest
import torch

from oracle import gradtts_oracle as O
from speech_backbones_b200.spec import synthetic_encoder_outputs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def glue_golden():
    return torch.load(os.path.join(ROOT, "tests", "golden", "gradtts_glue_golden.pt"), weights_only=False)


def _attn_from_tok(tok, Tx):
    B, Ty = tok.shape
    a = torch.zeros(B, Tx, Ty)
    idx = tok.long().clamp_min(0)
    a.scatter_(1, idx[:, None, :], (tok >= 0).float()[:, None, :])
    return a[:, None]


def _case_inputs(g, c):
    return synthetic_encoder_outputs(c["B"], c["Tx"], c["x_lengths"], c["dur_mean"], seed=g["seed"])


def _noise(g, c, device="cpu"):
    from speech_backbones_b200.gradtts import reference_order_noise
    torch.manual_seed(g["noise_seed"])
    return reference_order_noise(c["B"], 80, c["Ty"], torch.float32, "cpu").to(device)


@pytest.mark.parametrize("idx", range(5))
def test_oracle_prior_expand_matches_reference_golden(glue_golden, idx):
    assert len(glue_golden["cases"]) == 5
    c = glue_golden["cases"][idx]
    mu_x, logw, x_mask = _case_inputs(glue_golden, c)
    o = O.prior_expand(mu_x, logw, x_mask, c["length_scale"], c["temperature"], _noise(glue_golden, c))
    assert o["y_max_length"] == c["y_max_length"] and torch.equal(o["y_lengths"], c["y_lengths"])
    assert torch.equal(o["y_mask"], c["y_mask"]) and torch.equal(o["mu_y"], c["mu_y"]) and torch.equal(o["z"], c["z"])
    assert torch.equal(o["attn"], _attn_from_tok(c["tok"], c["Tx"]))


def test_oracle_prior_expand_properties():
    """Domain properties of the alignment, independent of the fixtures: every valid frame maps to exactly one unmasked
    token, tokens appear in order with their ceil'ed durations, masked frames are zero, z - mu_y is the scaled noise."""
    mu_x, logw, x_mask = synthetic_encoder_outputs(3, 23, [23, 11, 1], 0.7, seed=99)
    noise = torch.randn(3, 1, 80)          # placeholder, replaced below once Ty is known
    o = O.prior_expand(mu_x, logw, x_mask, 1.0, 1.0, None)
    attn, y_len = o["attn"][:, 0], o["y_lengths"]
    for b in range(3):
        L = int(y_len[b])
        assert attn[b, :, :L].sum(0).eq(1).all() and attn[b, :, L:].abs().sum() == 0
        tok = attn[b, :, :L].argmax(0)
        assert (tok[1:] >= tok[:-1]).all()
        dur = torch.ceil(torch.exp(logw[b, 0]) * x_mask[b, 0])
        assert torch.equal(torch.bincount(tok, minlength=23).float(), dur)
        assert o["mu_y"][b, :, L:].abs().sum() == 0
    del noise


def test_product_glue_has_no_cpu_path(sbk_lib):
    """The product entry points refuse CPU tensors loudly (no fallback): binding.prior_expand and synthesize_from_encoder."""
    from speech_backbones_b200 import gradtts as G
    from speech_backbones_b200.binding import prior_expand
    mu_x, logw, x_mask = synthetic_encoder_outputs(1, 6, [6], 0.5, seed=1)
    with pytest.raises(RuntimeError, match="CUDA"):
        prior_expand(mu_x, torch.ones(1, 6), x_mask.reshape(1, 6), torch.tensor([6]), 8)
    with pytest.raises(RuntimeError, match="CUDA"):
        G.synthesize_from_encoder(lambda *a: None, mu_x, logw, x_mask, 2)
    assert G.fix_len_compatibility(5) == 8 and G.fix_len_compatibility(8) == 8
    n = G.reference_order_noise(2, 80, 12, torch.float32, "cpu")
    assert n.shape == (2, 12, 80) and n.is_contiguous()


# ---------------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("idx", range(5))
def test_prior_expand_kernel_matches_reference_golden(sbk_lib, glue_golden, idx):
    from speech_backbones_b200.binding import prior_expand
    c = glue_golden["cases"][idx]
    mu_x, logw, x_mask = _case_inputs(glue_golden, c)
    w_ceil = torch.ceil(torch.exp(logw) * x_mask) * c["length_scale"]                 # tts.py:77-78, the reference's own ops
    y_lengths = torch.clamp_min(torch.sum(w_ceil, [1, 2]), 1).long()
    assert torch.equal(y_lengths, c["y_lengths"])
    B, Tx, Ty = c["B"], c["Tx"], c["Ty"]
    mu_y, z, y_mask, attn = prior_expand(mu_x.cuda(), w_ceil.reshape(B, Tx).cuda(), x_mask.reshape(B, Tx).cuda(),
                                         y_lengths.cuda(), Ty, _noise(glue_golden, c, "cuda"), c["temperature"], True)
    assert torch.equal(mu_y.cpu(), c["mu_y"]), "mu_y"
    assert torch.equal(z.cpu(), c["z"]), "z"
    assert torch.equal(y_mask.cpu(), c["y_mask"]), "y_mask"
    assert torch.equal(attn.cpu(), _attn_from_tok(c["tok"], Tx)), "attn"
    # without noise / without attn
    mu2, z2, _, none = prior_expand(mu_x.cuda(), w_ceil.reshape(B, Tx).cuda(), x_mask.reshape(B, Tx).cuda(), y_lengths.cuda(), Ty,
                                    None, 1.0, False)
    assert none is None and torch.equal(mu2, z2) and torch.equal(mu2.cpu(), c["mu_y"])


@pytest.mark.gpu
def test_synthesize_from_encoder_is_the_reference_glue(sbk_lib, glue_golden):
    """The drop-in for tts.py:77-99 with a recording decoder: what the decoder is handed must be the reference's tensors.
    (The CUDA generator differs from the CPU one, so z is checked as mu_y + the drawn noise / temperature instead.)"""
    from speech_backbones_b200 import gradtts as G
    c = glue_golden["cases"][0]
    mu_x, logw, x_mask = _case_inputs(glue_golden, c)
    seen = {}

    def decoder(z, mask, mu, n_timesteps, stoc, spk):
        seen.update(z=z, mask=mask, mu=mu, n=n_timesteps)
        assert z.is_contiguous() and mu.is_contiguous() and mask.shape == (c["B"], 1, c["Ty"])
        return mu + 1.0

    torch.manual_seed(5)
    enc, dec, attn = G.synthesize_from_encoder(decoder, mu_x.cuda(), logw.cuda(), x_mask.cuda(), 7, c["temperature"], False, None,
                                               c["length_scale"])
    L = c["y_max_length"]
    assert seen["n"] == 7 and torch.equal(seen["mu"].cpu(), c["mu_y"]) and torch.equal(seen["mask"].cpu(), c["y_mask"])
    assert torch.equal(enc.cpu(), c["mu_y"][:, :, :L]) and torch.equal(dec.cpu(), c["mu_y"][:, :, :L] + 1.0)
    assert torch.equal(attn.cpu(), _attn_from_tok(c["tok"], c["Tx"])[:, :, :L])
    torch.manual_seed(5)
    noise_tf = G.reference_order_noise(c["B"], 80, c["Ty"], torch.float32, "cuda")
    # (CPU arithmetic for the expectation: torch's CUDA `tensor / python_scalar` multiplies by the reciprocal, the CPU kernel
    #  and sbk_prior_expand divide)
    assert torch.equal(seen["z"].cpu(), seen["mu"].cpu() + noise_tf.cpu().transpose(1, 2) / c["temperature"])
    with pytest.raises(RuntimeError, match="CUDA"):
        G.synthesize_from_encoder(decoder, mu_x, logw, x_mask, 2)


@pytest.mark.gpu
def test_glue_feeds_the_sampler_end_to_end(sbk_lib, glue_golden):
    """encoder outputs -> sbk_prior_expand -> sbk_reverse_diffusion: the tensors go straight into the sampler (layout,
    dtype, T % 4 == 0) and the result equals the sampler run on the reference's own (z, mask, mu)."""
    from speech_backbones_b200 import UNetConfig, synthetic_state_dict
    from speech_backbones_b200 import gradtts as G
    c = glue_golden["cases"][1]
    mu_x, logw, x_mask = _case_inputs(glue_golden, c)
    dec = G.Diffusion(80, 64, precision="tf32").eval()
    dec.load_state_dict(synthetic_state_dict(UNetConfig()))
    dec = dec.cuda()
    torch.manual_seed(3)
    enc_out, y, attn = G.synthesize_from_encoder(dec, mu_x.cuda(), logw.cuda(), x_mask.cuda(), 4, c["temperature"], False, None,
                                                 c["length_scale"])
    torch.manual_seed(3)
    noise_tf = G.reference_order_noise(c["B"], 80, c["Ty"], torch.float32, "cuda")
    z_ref = c["mu_y"] + noise_tf.cpu().transpose(1, 2) / c["temperature"]          # CPU: true division, as in the kernel
    y_ref = dec(z_ref.contiguous().cuda(), c["y_mask"].cuda(), c["mu_y"].cuda(), 4)
    assert y.shape == (c["B"], 80, c["y_max_length"]) and torch.equal(y, y_ref[:, :, :c["y_max_length"]])
