"""bf16 mode bring-up on the GPU: stage-by-stage error of one estimator call against the CPU oracle (every named
intermediate, first divergent stage marked), golden estimator / trajectory cases, then speed vs tf32 at B=32 T=512.
usage: python scripts/gpu_bf16_check.py [quick]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402

ge.build()
from helpers import case_id, case_inputs, nhwc_to_nchw, rel_l2, stoc_noise  # noqa: E402
from oracle import gradtts_oracle as O  # noqa: E402
from speech_backbones_b200 import UNetConfig, synthetic_inputs, synthetic_state_dict  # noqa: E402
from speech_backbones_b200.binding import Engine  # noqa: E402

ATTN_INPUTS = {"estimator.downs.0.1.out", "estimator.downs.1.1.out", "estimator.downs.2.1.out",
               "estimator.mid_block1.out", "estimator.ups.0.1.out", "estimator.ups.1.1.out"}
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
cfg = UNetConfig()
sd = synthetic_state_dict(cfg)
engs = {p: Engine(precision=p) for p in ("bf16", "tf32")}
for e in engs.values():
    e.load_state_dict(sd)


def stagewise(eng, B, T):
    z, mask, mu, spk, _ = synthetic_inputs(B, T, ragged=True)
    t = torch.linspace(0.9, 0.2, B)
    xt = z * mask
    eng.debug_capture(True)
    y = eng.estimator(xt.cuda(), mask.cuda(), mu.cuda(), t.cuda())
    torch.cuda.synchronize()
    eng.debug_capture(False)
    taps = {}
    y_ref = O.estimator(sd, cfg, xt, mask, mu, t, None, taps=taps)
    worst = 0.0
    for name in eng.debug_names():
        if name not in taps:
            continue
        got = eng.debug_read(name)
        if got is None:
            continue
        ref = taps[name]
        if name.endswith(".ctx"):
            got = got.view(ref.shape)
        else:
            Bq, C, H, W = ref.shape
            got = nhwc_to_nchw(got, Bq, H, W, C, eng.debug_layout(name))
            if not name.endswith(".raw") and name not in ATTN_INPUTS:
                ref = ref * mask[:, None, :, ::mask.shape[-1] // W]
        e = rel_l2(got, ref)
        worst = max(worst, e)
        print(f"  {name:44s} layout={eng.debug_layout(name)} rel_l2={e:.3e} |ref|max={ref.abs().max().item():.3g}"
              + ("   <== LARGE" if not e < 5e-2 else ""))
    e = rel_l2(y.cpu(), y_ref)
    print(f"  estimator.out rel_l2={e:.3e}; worst stage {worst:.3e}; padded frames max |y| = "
          f"{(y.cpu() * (1 - mask)).abs().max().item():.3g}", flush=True)


for prec in ("bf16",) if quick else ("bf16", "tf32"):
    for B, T in ((2, 32), (3, 100), (1, 256), (1, 4)):
        print(f"== stagewise {prec} B={B} T={T}", flush=True)
        try:
            stagewise(engs[prec], B, T)
        except Exception as ex:      # keep going: one run should tell as much as possible
            print("  FAILED:", ex, flush=True)

golden = torch.load(os.path.join(ROOT, "tests", "golden", "gradtts_golden.pt"), weights_only=False)
for prec in ("bf16", "tf32"):
    eng = engs[prec]
    for c in golden["cases"]:
        if c["n_spks"] != 1:
            continue
        _, _, z, mask, mu, spk = case_inputs(golden, c)
        try:
            if c["kind"] == "est":
                y = eng.estimator((z * mask * c["scale"]).cuda(), mask.cuda(), mu.cuda(), torch.tensor(c["t"]).cuda()).cpu()
            else:
                noise = stoc_noise(golden, c).cuda() if c["stoc"] else None
                y = eng.reverse_diffusion(z.cuda(), mask.cuda(), mu.cuda(), c["N"], c["stoc"], None, noise).cpu()
            print(f"golden {prec} {case_id(c)}: rel_l2 = {rel_l2(y, c['out']):.3e}", flush=True)
        except Exception as ex:
            print(f"golden {prec} {case_id(c)}: FAILED {ex}", flush=True)

# ---- speed: one reverse step at config 2's shape, bf16 vs tf32 (graph replay + per-launch events)
B, T = 32, 512
z, mask, mu, _, _ = synthetic_inputs(B, T)
zd, md, mud = z.cuda(), mask.cuda(), mu.cuda()
outs = {}
for prec in ("tf32", "bf16"):
    eng = engs[prec]
    try:
        eng.reverse_diffusion(zd, md, mud, 3)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        outs[prec] = eng.reverse_diffusion(zd, md, mud, 50)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"# {prec} B={B} T={T} N=50: {ms:.1f} ms/call = {ms / 50:.3f} ms/step = {B * T / ms * 1e3:.0f} mel-frames/s", flush=True)
        prof = eng.profile_ops()
        with open(os.path.join(ROOT, "gpurun_out", f"ops_{prec}_check.txt"), "w") as f:
            f.write(f"# B={B} T={T} {prec}: {ms / 50:.3f} ms/step (graph replay)\n# sum of per-launch event times: {sum(p[1] for p in prof):.3f} ms\n")
            for n, m, fl, by in prof:
                f.write(f"{n:44s} {m:8.4f} ms {fl / (m * 1e-3) / 1e12 if m > 0 else 0:9.1f} TFLOP/s {by / (m * 1e-3) / 1e9 if m > 0 else 0:9.1f} GB/s\n")
    except Exception as ex:
        print(f"speed {prec}: FAILED {ex}", flush=True)
if len(outs) == 2:
    print("bf16 vs tf32 N=50 trajectory rel_l2:", rel_l2(outs["bf16"].cpu(), outs["tf32"].cpu()))
