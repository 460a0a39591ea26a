"""DiffVC.forward between the encoders and the decoder (DiffVC/model/vc.py:104-127).  Fixtures recorded
from the UNMODIFIED reference by scripts/make_golden_vc_glue.py.  The drop-in `convert_from_encoder` is host logic (torch
ops, device-agnostic: masked pads instead of the reference's per-sample copy loop), so it is checked here on the CPU with a
recording decoder - bit for bit, including the generator stream."""
import os

import pytest
import torch

from oracle import diffvc_oracle as O
from speech_backbones_b200.spec import DiffVCConfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def vg():
    return torch.load(os.path.join(ROOT, "tests", "golden", "diffvc_glue_golden.pt"), weights_only=False)


def _synth(g, c):
    gen = torch.Generator().manual_seed(g["seed"] + sum(c["lengths"]))
    T = max(c["lengths"])
    return torch.randn(c["B"], 80, T, generator=gen), torch.tensor(c["lengths"]), torch.randn(c["B"], 80, T, generator=gen)


@pytest.mark.parametrize("idx", range(3))
def test_oracle_matches_reference_golden(vg, idx):
    c = vg["cases"][idx]
    x, x_lengths, mean = _synth(vg, c)
    torch.manual_seed(vg["noise_seed"])
    o = O.prepare_decoder_inputs(DiffVCConfig(), x, x_lengths, mean)
    assert torch.equal(o["z"], c["z"]) and torch.equal(o["x_mask_new"], c["mask"])
    assert torch.equal(o["mean_new"], c["mean_new"]) and torch.equal(o["mean_x"], c["mean_x"])


@pytest.mark.parametrize("idx", range(3))
def test_convert_from_encoder_is_the_reference_glue(vg, idx):
    from speech_backbones_b200.diffvc import Diffusion, convert_from_encoder
    c = vg["cases"][idx]
    x, x_lengths, mean = _synth(vg, c)
    real = Diffusion(80, 256, 128, True, 0.05, 20.0)
    seen = {}

    class Dec:
        compute_diffused_mean = staticmethod(real.compute_diffused_mean)

        def __call__(self, z, mask, mean_, ref, ref_mask, mean_ref, cc, n, mode):
            seen.update(z=z.clone(), mask=mask, mean=mean_, n=n, mode=mode)
            return z * 2.0

    torch.manual_seed(vg["noise_seed"])
    B = c["B"]
    mean_x, y = convert_from_encoder(Dec(), x, x_lengths, mean, torch.zeros(B, 80, 8), torch.ones(B, 1, 8), torch.zeros(B, 80, 8),
                                     torch.zeros(B, 256), 6, "ml")
    assert seen["n"] == 6 and seen["mode"] == "ml"
    assert torch.equal(seen["z"], c["z"]) and torch.equal(seen["mask"], c["mask"]) and torch.equal(seen["mean"], c["mean_new"])
    assert torch.equal(mean_x, c["mean_x"]) and torch.equal(y, (c["z"] * 2.0)[:, :, :max(c["lengths"])])


def test_convert_from_encoder_rejects_unpadded_batches():
    from speech_backbones_b200.diffvc import convert_from_encoder
    with pytest.raises(RuntimeError, match="frames"):
        convert_from_encoder(None, torch.zeros(1, 80, 10), torch.tensor([8]), torch.zeros(1, 80, 10), None, None, None, None, 1)
