#!/usr/bin/env python
"""Headline benchmark: mel-frames/sec of the Grad-TTS reverse-diffusion sampler at N=50 steps.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One bench "step" = one full `Diffusion.forward(z, mask, mu, n_timesteps=50)` call on the workload
BASELINE.json quotes the metric on (config 2: B=32 utterances x T=512 frames, fp32 in/out, per GPU;
weak scaling: every rank samples its own 32 utterances, outputs all-gathered).  Prints ONE JSON line.

  value        frames/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e          same metric through the host-buffer entry point (pinned host tensors in/out, copies timed)
  roofline     the dominant kernel class (3x3 conv implicit GEMMs), timed per launch with CUDA events
  cpu_baseline the CPU oracle (a port of the reference's PyTorch path) on a bounded sample, this box's cores
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: Grad-TTS batch=32, T~512, N=50, fp32, 1xB200
    "gradtts_b32_t512_n50": dict(B=32, T=512, N=50, n_spks=1),
    # BASELINE.json configs[4]'s per-GPU share (2048 utterances over 8 GPUs = 256 per GPU); not the default bench line:
    #   torchrun --nproc-per-node 8 bench.py --gpus 8 --workload gradtts_b256_t512_n50 --steps 2 --warmup 3 --no-fp32-leg
    "gradtts_b256_t512_n50": dict(B=256, T=512, N=50, n_spks=1),
}
FLOP_PER_FRAME_STEP = 134.15e6      # 67,077,120 MAC per mel frame per reverse step
IDEAL_BYTES_PER_FRAME_STEP = 713280.0


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm_gbs=6650.0, bf16=1400.0, src="fallback")


def measure_tf32_matmul_tflops(torch, dev, seconds=1.0):
    """Sustained cuBLAS TF32 rate on this GPU (torch.matmul 8192^3, fp32 tensors, allow_tf32), back to back for
    `seconds`: MEASURED_PEAKS.json only carries the bf16 rate, and tf32 is not exactly half of it in practice."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn(n, n, device=dev)
        b = torch.randn(n, n, device=dev)
        for _ in range(3):
            a @ b
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters, t0 = 0, time.perf_counter()
        e0.record()
        while True:
            for _ in range(10):
                a @ b
            iters += 10
            torch.cuda.synchronize()
            if time.perf_counter() - t0 > seconds:
                break
        e1.record()
        torch.cuda.synchronize()
        return 2.0 * n ** 3 * iters / (e0.elapsed_time(e1) * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, dev):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(dev)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.th.join(timeout=2)
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def _ncpu():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count()


_THREADS = None


def pick_cpu_threads(torch, O, cfg, sd):
    """The reference arm gets all the host threads it can USE: sweep the thread count on a tiny problem and keep
    the fastest (on many-core boxes PyTorch's CPU convs slow down badly when oversubscribed)."""
    global _THREADS
    if _THREADS is not None:
        torch.set_num_threads(_THREADS)
        return _THREADS
    from speech_backbones_b200 import synthetic_inputs
    z, mask, mu, spk, _ = synthetic_inputs(1, 128, n_spks=cfg.n_spks)
    n = _ncpu()
    cands = sorted({c for c in (4, 8, 16, 32, 64, n) if c <= n})
    best, best_t = cands[0], float("inf")
    for c in cands:
        torch.set_num_threads(c)
        O.reverse_diffusion(sd, cfg, z, mask, mu, 1, False, spk)
        t0 = time.perf_counter()
        O.reverse_diffusion(sd, cfg, z, mask, mu, 1, False, spk)
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
    _THREADS = best
    torch.set_num_threads(best)
    return best


def cpu_reference_sample(wl, torch, n_steps=2, b_sample=2, repeats=1):
    """Time the CPU oracle (port of the reference PyTorch path) on a bounded sample of the workload:
    `b_sample` utterances at the workload's T for `n_steps` Euler steps, after one warm-up step.
    Cost is linear in B and in N (the loop body is step-independent, diffusion.py:258-274)."""
    from oracle import gradtts_oracle as O
    from speech_backbones_b200 import UNetConfig, synthetic_inputs, synthetic_state_dict
    cfg = UNetConfig(n_spks=wl["n_spks"])
    sd = synthetic_state_dict(cfg)
    threads = pick_cpu_threads(torch, O, cfg, sd)
    z, mask, mu, spk, _ = synthetic_inputs(b_sample, wl["T"], n_spks=cfg.n_spks)
    O.reverse_diffusion(sd, cfg, z, mask, mu, 1, False, spk)                    # warm-up
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        O.reverse_diffusion(sd, cfg, z, mask, mu, n_steps, False, spk)
        best = min(best, time.perf_counter() - t0)
    sec_per_frame_step = best / (b_sample * wl["T"] * n_steps)
    frames_per_sec = 1.0 / (sec_per_frame_step * wl["N"])
    sample = (f"oracle port (PyTorch CPU fp32, {threads} threads = fastest of a sweep up to {_ncpu()} usable cores), "
              f"B={b_sample} x T={wl['T']} for {n_steps} of N={wl['N']} Euler steps ({best:.2f} s), "
              f"scaled linearly in B and N")
    return frames_per_sec, sec_per_frame_step, sample, threads


def run_reference(args, wl):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    for i in range(args.warmup + args.steps):
        fps, spfs, sample, threads = cpu_reference_sample(wl, torch, n_steps=2, b_sample=2)
        if i >= args.warmup:
            vals.append((fps, spfs))
    fps = statistics.mean(v[0] for v in vals)
    ms_full = statistics.mean(v[1] for v in vals) * wl["B"] * wl["T"] * wl["N"] * 1e3
    out = {
        "impl": "reference", "metric": "mel-frames/sec at N=50 reverse-diffusion steps", "value": fps,
        "unit": "mel-frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_full, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "batch_per_gpu": wl["B"], "frames": wl["T"], "n_timesteps": wl["N"],
                   "note": "CPU reference arm: each step is a bounded sample, ms_per_step is the extrapolated full step"},
        "cpu_baseline": {"value": fps, "unit": "mel-frames/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "mel-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def run_ours(args, wl):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the sampler has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    from speech_backbones_b200 import UNetConfig, estimator_param_spec, synthetic_inputs, synthetic_state_dict
    from speech_backbones_b200.gradtts import Diffusion
    from speech_backbones_b200.sharded import broadcast_state_dict

    B, T, N = wl["B"], wl["T"], wl["N"]
    cfg = UNetConfig(n_spks=wl["n_spks"])
    # weights: rank 0 owns them, NCCL broadcast to the other ranks (north_star: weight broadcast + mel gather only)
    t0 = time.perf_counter()
    sd = synthetic_state_dict(cfg) if rank == 0 else None
    if world > 1:
        sd = broadcast_state_dict(sd, estimator_param_spec(cfg), dev)
        torch.cuda.synchronize()
    bcast_s = time.perf_counter() - t0
    dec = Diffusion(cfg.n_feats, cfg.dim, n_spks=cfg.n_spks, precision=args.precision).eval()
    dec.load_state_dict(sd)
    dec = dec.to(dev)
    eng = dec.engine()

    z, mask, mu, spk, _ = synthetic_inputs(B, T, seed=1234 + rank, n_spks=cfg.n_spks)
    zd, md, mud = z.to(dev), mask.to(dev), mu.to(dev)
    spd = None if spk is None else spk.to(dev)
    gathered = torch.empty((world * B, cfg.n_feats, T), dtype=torch.float32, device=dev) if world > 1 else None

    def one_call():
        y = dec(zd, md, mud, N, False, spd)
        if world > 1:
            dist.all_gather_into_tensor(gathered, y)       # output mel gather over NVLink
        return y

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_call()
    fence()
    clocks = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    launches = 0
    for _ in range(args.steps):
        one_call()
        launches += eng.last_launch_count()
    e1.record()
    fence()
    ms_total = e0.elapsed_time(e1)
    clk = clocks.stop() if clocks else None
    tms = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_step = tms.item() / args.steps
    value = world * B * T / (ms_step * 1e-3)

    # ---- end to end through the host-buffer entry point (pinned host memory in/out, copies inside the timed region)
    zh, mh, muh = z.pin_memory(), mask.pin_memory(), mu.pin_memory()
    outh = torch.empty_like(z).pin_memory()
    sph = None if spk is None else spk.pin_memory()
    eng.reverse_diffusion_host(zh, mh, muh, N, False, sph, None, outh)          # warm-up
    fence()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        eng.reverse_diffusion_host(zh, mh, muh, N, False, sph, None, outh)
        _ = float(outh[0, 0, 0])                                               # host read of the result
    fence()
    te = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * B * T / te.item()
    h2d = (zh.numel() + mh.numel() + muh.numel() + (0 if sph is None else sph.numel())) * 4
    d2h = outh.numel() * 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel class, one CUDA event pair per launch
    peaks = load_peaks()
    prof = eng.profile_ops()
    conv = [(n, ms, fl, by) for n, ms, fl, by in prof if n.endswith(".raw")]
    conv_ms, conv_fl, conv_by = (sum(x[i] for x in conv) for i in (1, 2, 3))
    all_ms = sum(x[1] for x in prof)
    # bf16 operands: the measured cuBLAS bf16 rate of MEASURED_PEAKS.json.  tf32: that file has no tf32 figure, so the
    # denominator is the larger of half the bf16 rate and a cuBLAS TF32 matmul timed here (sustained, ~1 s)
    tf32_here = measure_tf32_matmul_tflops(torch, dev) if args.precision != "bf16" else None
    tensor_peak = peaks["bf16"] if args.precision == "bf16" else max(peaks["bf16"] * 0.5, tf32_here)
    achieved = conv_fl / (conv_ms * 1e-3) / 1e12
    by_kind = {}
    for n, ms, fl, by in prof:
        k = ("conv3x3" if n.endswith(".raw") else "gn_mish_act" if n.endswith(".act") else
             "attention" if (".2." in n or "mid_attn" in n) else "resample" if ".3." in n else
             "final_euler" if n == "estimator.out" else "resblock_tail")
        by_kind[k] = by_kind.get(k, 0.0) + ms
    # DRAM bytes per launch of the same kernel class from the committed ncu capture (scripts/ncu_traffic.py), if present
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r1_traffic_conv3x3.json")
    if os.path.exists(tpath) and args.precision == "tf32" and (B, T) == (32, 512):
        tj = json.load(open(tpath))
        traffic, traffic_src = tj["dram_bytes_per_launch"], "profiles/r1_traffic_conv3x3.json (ncu dram__bytes_read+write, %d launches)" % tj["launches"]
    roofline = {
        "kernel": "conv3x3 implicit GEMM (25 launches/step)", "bound": "tensor", "achieved": achieved, "peak": tensor_peak,
        "unit": "TFLOP/s", "frac": achieved / tensor_peak, "traffic": traffic, "traffic_source": traffic_src,
        "algorithmic_bytes_per_launch": conv_by / max(1, len(conv)),
        "peak_note": (f"{peaks['src']} cuBLAS bf16 sustained (MEASURED_PEAKS.json)" if args.precision == "bf16" else
                      f"max(0.5 x {peaks['src']} cuBLAS bf16 sustained = {peaks['bf16'] * 0.5:.1f}, cuBLAS TF32 matmul 8192^3 "
                      f"sustained measured in this run = {tf32_here:.1f})"),
        "launches": len(conv), "avg_launch_ms": conv_ms / max(1, len(conv)),
        "flop_per_launch_avg": conv_fl / max(1, len(conv)), "share_of_step": conv_ms / all_ms,
        "hbm": {"achieved_gbs": conv_by / (conv_ms * 1e-3) / 1e9, "peak_gbs": peaks["hbm_gbs"],
                "frac": conv_by / (conv_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]},
        "step_ms_by_kind": {k: round(v, 4) for k, v in by_kind.items()},
        "whole_step": {"tflops": FLOP_PER_FRAME_STEP * B * T / (ms_step / N * 1e-3) / 1e12,
                       "ideal_hbm_gbs": IDEAL_BYTES_PER_FRAME_STEP * B * T / (ms_step / N * 1e-3) / 1e9},
    }
    fps_cpu, _, sample, threads = cpu_reference_sample(wl, torch, n_steps=3, b_sample=2) if world == 1 else (None,) * 4
    # the exact-fp32 CUDA-core mode of the same engine, one timed call (context for the tf32 headline)
    fp32_leg = None
    if world == 1 and args.precision != "fp32" and not args.no_fp32_leg:
        dec32 = Diffusion(cfg.n_feats, cfg.dim, n_spks=cfg.n_spks, precision="fp32").eval()
        dec32.load_state_dict(sd)
        dec32 = dec32.to(dev)
        dec32(zd, md, mud, N, False, spd)
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        y32 = dec32(zd, md, mud, N, False, spd)
        f1.record()
        torch.cuda.synchronize()
        ytc = dec(zd, md, mud, N, False, spd)
        rel = ((ytc - y32).double().norm() / y32.double().norm()).item()
        fp32_leg = {"value": B * T / (f0.elapsed_time(f1) * 1e-3), "unit": "mel-frames/s", "dtype": "f32",
                    "note": "same engine, precision=fp32 (CUDA-core FFMA convs), 1 timed call",
                    "rel_l2_of_headline_output_vs_this": rel}
    # the bf16-operand mode of the same engine (BASELINE config 3's arithmetic), one timed call, for context
    bf16_leg = None
    if world == 1 and args.precision == "tf32" and not args.no_fp32_leg:
        dec16 = Diffusion(cfg.n_feats, cfg.dim, n_spks=cfg.n_spks, precision="bf16").eval()
        dec16.load_state_dict(sd)
        dec16 = dec16.to(dev)
        dec16(zd, md, mud, N, False, spd)
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        y16 = dec16(zd, md, mud, N, False, spd)
        f1.record()
        torch.cuda.synchronize()
        ref = y32 if fp32_leg is not None else dec(zd, md, mud, N, False, spd)
        bf16_leg = {"value": B * T / (f0.elapsed_time(f1) * 1e-3), "unit": "mel-frames/s", "dtype": "bf16",
                    "note": "same engine, precision=bf16 (bf16 operand tensors + weights, fp32 accumulate/GN/state), 1 timed call",
                    "rel_l2_vs_fp32_mode": ((y16 - ref).double().norm() / ref.double().norm()).item()}
        del dec16
    out = {
        "metric": "mel-frames/sec at N=50 reverse-diffusion steps", "value": value, "unit": "mel-frames/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32", "tf32": "tf32", "bf16": "bf16"}[args.precision], "data": "synthetic",
        "config": {"workload": args.workload, "batch_per_gpu": B, "global_batch": B * world, "frames": T,
                   "n_timesteps": N, "stoc": False, "parallelism": f"dp{world}",
                   "l2": f"per-step working set ({eng.workspace_bytes(B, T) / 1e9:.1f} GB of activations) exceeds the 126 MB L2; no flush needed",
                   "weights": "synthetic seeded (no checkpoints ship with the reference)",
                   "weight_broadcast_s": round(bcast_s, 4)},
        "frame_steps_per_s": value * N,
        "e2e": {"value": e2e_value, "unit": "mel-frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "clocks": clk,
        "roofline": roofline,
    }
    if fp32_leg is not None:
        out["fp32_mode"] = fp32_leg
    if bf16_leg is not None:
        out["bf16_mode"] = bf16_leg
    if fps_cpu is not None:
        out["cpu_baseline"] = {"value": fps_cpu, "unit": "mel-frames/s", "cores": threads, "kind": "port",
                               "sample": sample}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="gradtts_b32_t512_n50", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="tf32", choices=["fp32", "tf32", "bf16"],
                    help="tf32: tcgen05 tensor cores, fp32 accumulate/IO (PyTorch's default GPU conv arithmetic); "
                         "fp32: CUDA-core FFMA path; bf16: bf16 operand tensors (BASELINE config 3's arithmetic; not the "
                         "headline, which is quoted on the fp32 config)")
    ap.add_argument("--no-fp32-leg", action="store_true", help="skip the extra exact-fp32 timing call")
    ap.add_argument("--batch", type=int, default=None, help="override B (debug only; not a valid bench line)")
    ap.add_argument("--frames", type=int, default=None, help="override T (debug only)")
    ap.add_argument("--n-timesteps", type=int, default=None, help="override N (debug only)")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.batch: wl["B"] = args.batch
    if args.frames: wl["T"] = args.frames
    if args.n_timesteps: wl["N"] = args.n_timesteps
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
