"""GPU parity tests for the DiffVC sampler: libsbk's U-Net + pf/em/ml samplers vs the
committed outputs of the UNMODIFIED reference.  The conditioning vectors (the hoisted, xt-independent branch) are
computed by the CPU oracle here so that the C ABI is tested in isolation; the drop-in module test lets the module
compute them itself (natively, `sbk_vc_conditioning`)."""
import os

import pytest
import torch

from helpers import rel_l2
from oracle import diffvc_oracle as O
from speech_backbones_b200.spec import DiffVCConfig, diffvc_param_spec, synthetic_diffvc_inputs, synthetic_state_dict

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# bf16: the U-Net runs on bf16 operand tensors (the hoisted conditioning branch stays tf32)
# fp32x3: the fp32-class tensor-core mode (tf32 + fp16-correction splits; the modules' default); fp32: the CUDA-core FFMA path
# (DiffVC's 3x3 convs run K up to 9 x 2048: fp32x3 measures 1.05e-5 per call, the CUDA-core fp32 mode 2-3e-6)
TOL = {"fp32": (1e-4, 2e-3), "fp32x3": (2e-5, 2e-4), "tf32": (4e-3, 1e-2), "bf16": (3e-2, 4e-2)}   # (estimator call, trajectory)
COND_TOL = {"fp32": 1e-5, "fp32x3": 1e-5, "tf32": 4e-3, "bf16": 4e-3}          # the hoisted RefBlock + cond_block branch


@pytest.fixture(scope="module")
def vc_golden():
    return torch.load(os.path.join(ROOT, "tests", "golden", "diffvc_golden.pt"), weights_only=False)


@pytest.fixture(scope="module")
def vc_engines(sbk_lib):
    from speech_backbones_b200.binding import Engine
    cfg = DiffVCConfig()
    sd = synthetic_state_dict(cfg, 1234, spec=diffvc_param_spec(cfg))
    cache = {}

    def get(precision):
        if precision not in cache:
            e = Engine(80, cfg.dim_unet, model="diffvc", dim_cond=cfg.dim_spk, precision=precision)
            e.load_state_dict(sd)
            cache[precision] = e
        return cache[precision], cfg, sd
    yield get
    for e in cache.values():
        e.close()


def _inputs(g, c):
    return synthetic_diffvc_inputs(c["B"], c["T"], c["Tr"], seed=g["seed"], ragged=c["ragged"])


@pytest.mark.parametrize("precision", ["fp32", "fp32x3", "tf32", "bf16"])
def test_vc_estimator_vs_reference_golden(vc_engines, vc_golden, precision):
    eng, cfg, sd = vc_engines(precision)
    for c in [c for c in vc_golden["cases"] if c["kind"] == "est"]:
        z, mask, mean, r, rmask, mean_ref, spk = _inputs(vc_golden, c)
        t = torch.tensor(c["t"])
        g = O._gamma(cfg, 0, 0.5)
        xt_ref = ((r * g + mean_ref * (1.0 - g)) * rmask)[:, None]
        _, cond = O.conditioning(sd, cfg, xt_ref, rmask, spk, t)
        y = eng.vc_estimator((z * mask).cuda(), mask.cuda(), mean.cuda(), cond.cuda(), t.cuda()).cpu()
        err = rel_l2(y, c["out"])
        print(precision, c["B"], c["T"], "rel_l2", err)
        assert err <= TOL[precision][0]
        assert (y * (1 - mask)).abs().max().item() == 0.0


@pytest.mark.parametrize("precision", ["fp32", "fp32x3", "tf32", "bf16"])
def test_vc_samplers_vs_reference_golden(vc_engines, vc_golden, precision):
    eng, cfg, sd = vc_engines(precision)
    for c in [c for c in vc_golden["cases"] if c["kind"] == "traj"]:
        z, mask, mean, r, rmask, mean_ref, spk = _inputs(vc_golden, c)
        N, mode = c["N"], c["mode"]
        rows = []
        for i in range(N):                                  # the hoisted conditioning branch, step by step
            t, _, _, _, g0t = O.step_coefficients(cfg, N, i, mode)
            xt_ref = ((r * g0t + mean_ref * (1.0 - g0t)) * rmask)[:, None]
            rows.append(O.conditioning(sd, cfg, xt_ref, rmask, spk, t * torch.ones(c["B"]))[1])
        cond = torch.stack(rows)
        noise = None
        if mode != "pf":                                     # the reference's draws: manual_seed, then randn_like(z) per step
            torch.manual_seed(vc_golden["noise_seed"])
            noise = torch.stack([torch.randn_like(z) for _ in range(N)]).cuda()
        y = eng.vc_reverse_diffusion(z.cuda(), mask.cuda(), mean.cuda(), cond.cuda(), N, mode, noise).cpu()
        err = rel_l2(y, c["out"])
        print(precision, mode, N, "rel_l2", err)
        assert err <= TOL[precision][1]
        assert (y * (1 - mask)).abs().max().item() == 0.0


@pytest.mark.parametrize("precision", ["fp32", "fp32x3", "tf32", "bf16"])
def test_vc_native_conditioning_vs_oracle(vc_engines, vc_golden, precision):
    """RefBlock + cond_block natively in EVERY precision: sbk_vc_conditioning vs the CPU oracle,
    every step.  The fp32-class handles (fp32x3, and the CUDA-core fp32 mode) run the RefBlock convs with the tf32 + fp16-correction split."""
    eng, cfg, sd = vc_engines(precision)
    c = next(c for c in vc_golden["cases"] if c["kind"] == "traj" and c["mode"] == "ml" and c["B"] == 2)
    z, mask, mean, r, rmask, mean_ref, spk = _inputs(vc_golden, c)
    N = c["N"]
    got = eng.vc_conditioning(r.cuda(), rmask.cuda(), mean_ref.cuda(), spk.cuda(), N).cpu()
    for i in range(N):
        t, _, _, _, g0t = O.step_coefficients(cfg, N, i, "ml")
        xt_ref = ((r * g0t + mean_ref * (1.0 - g0t)) * rmask)[:, None]
        ref = O.conditioning(sd, cfg, xt_ref, rmask, spk, t * torch.ones(c["B"]))[1]
        err = rel_l2(got[i], ref)
        print(precision, "step", i, "cond rel_l2", err)
        assert err <= COND_TOL[precision]
    # and the sampler fed by the native table
    torch.manual_seed(vc_golden["noise_seed"])
    noise = torch.stack([torch.randn_like(z) for _ in range(N)]).cuda()
    y = eng.vc_reverse_diffusion(z.cuda(), mask.cuda(), mean.cuda(), got.cuda(), N, "ml", noise).cpu()
    assert rel_l2(y, c["out"]) <= TOL[precision][1]


def test_vc_dropin_module_tf32_native_conditioning(vc_golden):
    from speech_backbones_b200.diffvc import Diffusion
    cfg = DiffVCConfig()
    sd = synthetic_state_dict(cfg, vc_golden["seed"], spec=diffvc_param_spec(cfg))
    dec = Diffusion(80, 256, 128, True, 0.05, 20.0, precision="tf32").eval()
    dec.load_state_dict(sd, strict=True)
    dec = dec.cuda()
    c = next(c for c in vc_golden["cases"] if c["kind"] == "traj" and c["mode"] == "pf")
    args = [v.cuda() for v in _inputs(vc_golden, c)]
    y = dec(*args, n_timesteps=c["N"], mode="pf")
    assert rel_l2(y.cpu(), c["out"]) <= TOL["tf32"][1]


def test_vc_dropin_module(vc_golden):
    """Diffusion(...).load_state_dict(strict) -> .cuda() -> forward(...): the call DiffVC/model/vc.py:125 makes."""
    from speech_backbones_b200.diffvc import Diffusion
    cfg = DiffVCConfig()
    sd = synthetic_state_dict(cfg, vc_golden["seed"], spec=diffvc_param_spec(cfg))
    dec = Diffusion(80, 256, 128, True, 0.05, 20.0).eval()
    dec.load_state_dict(sd, strict=True)
    dec = dec.cuda()
    c = next(c for c in vc_golden["cases"] if c["kind"] == "traj" and c["mode"] == "pf")
    args = [v.cuda() for v in _inputs(vc_golden, c)]
    y = dec(*args, n_timesteps=c["N"], mode="pf")
    assert dec.precision == "fp32x3"                               # the default is the fp32-class tensor-core mode
    assert rel_l2(y.cpu(), c["out"]) <= TOL["fp32x3"][1]
    assert torch.isfinite(dec(*args, n_timesteps=3, mode="ml")).all()
    z = args[0]
    assert dec(*args, n_timesteps=3, mode="bogus") is z          # reference behaviour: print + return z
    with pytest.raises(RuntimeError, match="CUDA"):
        Diffusion(80, 256, 128, True, 0.05, 20.0)(*[v.cpu() for v in args], n_timesteps=2, mode="pf")
