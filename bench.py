#!/usr/bin/env python
"""Critic-VAE hot-path benchmark (BASELINE.json: "VAE train frames/s (fwd+bwd+loss)").

  python bench.py --gpus 1 --steps 20 --warmup 5            # this repo's B200 path
  python bench.py --impl reference --steps 3 --warmup 1     # the reference algorithm on host cores
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = the reference's training iteration (vae.py:47-58) on one synthetic batch: critic value
-> encoder -> reparametrise -> decoder -> MS-SSIM + KLD -> backward -> (NCCL gradient all-reduce)
-> Adam.  Workload: BASELINE.json configs[1], batch 256 per GPU, synthetic 64x64 frames, random-init
VAE (tests/synth.py seed 0), the shipped critic checkpoint.  Prints ONE JSON line (rank 0).
"""
import argparse
import functools
import json
import operator
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in ("tests", "critic-vae_b200"):
    sys.path.insert(0, os.path.join(ROOT, _p))

import numpy as np
import torch

CONV_FLOPS = 484_966_400          # conv layers, forward, per frame
TRAIN_FLOPS = 1_437_622_272       # fwd + dgrad + wgrad incl. linear layers, per frame
E0_FLOPS = 19_660_800             # encoder conv 0 needs no data-gradient
CRITIC_CKPT = os.path.join(ROOT, "critic-vae_b200", "saved-networks",
                           "critic-rewidx=1-cepochs=15-datamode=trunk-datasize=99999-shift=12-chfak=1-dropout=0.3.pt")
METRIC = "VAE train frames/s (fwd+bwd+loss+Adam)"


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel family, from the committed
    summary of the CURRENT round's `ncu --set full` capture (profiles/r02_ncu_traffic.json, written by
    tools/ncu_summary.py); None when there is no capture of the kernels as they are now."""
    path = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    if not os.path.exists(path):
        return {}, None
    with open(path) as f:
        d = json.load(f)
    return d.get("bytes_per_launch", {}), d.get("source")


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p.get("bf16_tflops_sustained", 1400.0), p.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json, sustained)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, threading.Event(), []

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.05)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        mhz = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows if len(r) >= 6)]
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


# --------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's algorithm (oracle port) on the host cores
# --------------------------------------------------------------------------------------------------
def cpu_train_step_rate(batch, steps, warmup, threads, budget_s=None):
    """Training steps of the oracle port on the host cores.  With `budget_s` the loop keeps going (at least `steps`
    steps) until that much wall time has been spent, so the baseline is a 10-30 s sample whatever the core count."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import critic_vae_oracle as O
    import synth
    torch.set_num_threads(threads)
    enc, dec = synth.make_vae_state(0)
    crit = torch.load(CRITIC_CKPT, map_location="cpu")
    keys = [f"encoder.{k}" for k in O.PARAM_KEYS_ENC] + [f"decoder.{k}" for k in O.PARAM_KEYS_DEC]
    m, v = {}, {}
    x, eps = synth.make_frames(batch, seed=100), synth.make_eps(batch, seed=101)
    times = []
    t_begin, s = time.perf_counter(), -1
    while True:
        s += 1
        if s >= warmup + steps and (budget_s is None or time.perf_counter() - t_begin >= budget_s or s >= warmup + 400):
            break
        t0 = time.perf_counter()
        pred = O.critic_forward(crit, x)
        _, _, _, _, grads = O.loss_and_grads(enc, dec, x, pred, eps)
        for k in keys:
            sd, kk = (enc, k[8:]) if k.startswith("encoder.") else (dec, k[8:])
            if k not in m:
                m[k], v[k] = torch.zeros_like(sd[kk]), torch.zeros_like(sd[kk])
            O.adam_step(sd[kk], grads[k], m[k], v[k], s + 1)
        if s >= warmup:
            times.append(time.perf_counter() - t0)
    return batch / (sum(times) / len(times)), sum(times) / len(times), len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    batch = args.batch            # the same per-step batch as the B200 arm (256): same config on both arms
    rate, sec, _ = cpu_train_step_rate(batch, args.steps, args.warmup, threads)
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE.json configs[1]: Critic-VAE training step (critic+fwd+MS-SSIM/KLD+bwd+Adam), "
                                   f"batch {batch} per GPU, 64x64x3 synthetic frames, random-init VAE (seed 0), shipped critic",
                       "per_gpu_batch": batch, "global_batch": batch, "parallelism": "host cores",
                       "note": "the reference's own algorithm (oracle port, torch CPU fp32) on the host cores; every step is one full batch"},
            "cpu_baseline": {"value": rate, "unit": "frames/s", "cores": threads, "kind": "port",
                             "sample": f"{args.steps} training steps of batch {batch} (oracle/critic_vae_oracle.py, torch CPU fp32)"},
            "e2e": {"value": rate, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# this repo's arm
# --------------------------------------------------------------------------------------------------
def build_modules(device):
    import synth
    import vae_nets
    from critic_net import Critic
    vae = vae_nets.VariationalAutoencoder().to(device)
    enc, dec = synth.make_vae_state(0)
    vae.encoder.load_state_dict(enc)
    vae.decoder.load_state_dict(dec)
    critic = Critic()
    critic.load_state_dict(torch.load(CRITIC_CKPT, map_location="cpu"))
    critic.eval().to(device)
    vae.train()
    return vae, critic


def mask_iou_secondary(device, hbm_peak):
    """BASELINE.json's second metric, "mask+IoU frames/s": (a) the one-pass difference-map -> clamp -> uint8 ->
    threshold -> IoU kernel (vae_utility.py:279-284,153-157,57-59 for the 13-threshold sweep of vae.py:121) on a
    scaled synthetic N (so it is HBM- not launch-bound; 45,056 algorithmic bytes per frame) and (b) the
    whole `-video -thresh` path of configs[2] (1200 uint8 frames in host memory -> critic -> encoder -> 2 decodes ->
    difference map -> all thresholds) through vae_utility.eval_threshold_sweep."""
    import synth
    import vae_utility as U
    from cvae_native import binding as L
    N = 1 << 15
    g = torch.Generator(device=device).manual_seed(5)
    diff = torch.rand(N, 64, 64, dtype=torch.float64, device=device, generator=g)
    gt = (torch.rand(N, 64, 64, device=device, generator=g) > 0.7).to(torch.uint8)
    thr = list(range(0, 130, 10))
    for _ in range(3):
        U._mask_iou(diff, gt, 0.5, 2.0, thr[0], thr)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        U._mask_iou(diff, gt, 0.5, 2.0, thr[0], thr)
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3 / reps
    gbs = N * 45056 / sec / 1e9
    out = {"kernel_frames_per_s": N / sec, "frames": N, "algorithmic_bytes_per_frame": 45056, "achieved_GBps": gbs,
           "hbm_peak_GBps": hbm_peak, "frac": gbs / hbm_peak, "thresholds": len(thr),
           "note": "timed region includes the small allocations of the host wrapper; input 1.2 GB > 126 MB L2"}
    del diff, gt
    # configs[2]: the -video -thresh path end to end on a synthetic 1200-frame episode
    vae, critic = build_modules(device)
    vae.eval()
    frames_u8 = (synth.make_frames(64, seed=40).permute(0, 2, 3, 1) * 255).round().to(torch.uint8).numpy()
    frames_u8 = np.tile(frames_u8, (19, 1, 1, 1))[:1200]
    gtm = np.tile(synth.make_gt_masks(64, seed=41), (19, 1, 1))[:1200]
    U.eval_threshold_sweep(frames_u8, vae, critic, gtm)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sweep = U.eval_threshold_sweep(frames_u8, vae, critic, gtm)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out["video_thresh_e2e"] = {"frames_per_s": 1200 / dt, "frames": 1200, "thresholds": len(sweep),
                               "workload": "BASELINE.json configs[2] on a synthetic stand-in episode (X.npy / Y.npy are not shipped): "
                                           "host uint8 frames -> critic -> encoder -> 2 decodes -> difference map -> 13 thresholds, IoU"}
    return out


def latent_secondary(device, hbm_peak):
    """HBM fraction of the fused reparameterise + critic-concat kernel (vae_nets.py:48-51,143) on a scaled synthetic
    N: at batch 256 it moves 133 KB and is launch-bound, so the bandwidth claim needs 2^21 rows (1.1 GB; 520
    algorithmic bytes per row: mu, logvar, eps, pred in, z|pred out -- SURVEY.md 8d)."""
    from cvae_native import binding as L
    N = 1 << 21
    g = torch.Generator(device=device).manual_seed(9)
    ml = torch.randn(N, 64, device=device, generator=g)
    eps = torch.randn(N, 32, device=device, generator=g)
    pred = torch.rand(N, device=device, generator=g)
    zc = torch.empty(N, 33, device=device)
    parts = torch.empty(L.lib.cvae_latent_kld_partials(N), dtype=torch.float64, device=device)
    call = lambda: L.check(L.lib.cvae_latent_fwd(N, 1, ml.data_ptr(), eps.data_ptr(), pred.data_ptr(), zc.data_ptr(), parts.data_ptr(),
                                                 L.stream_ptr()))
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        call()
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3 / reps
    gbs = N * 520 / sec / 1e9
    return {"rows": N, "algorithmic_bytes_per_row": 520, "achieved_GBps": gbs, "hbm_peak_GBps": hbm_peak, "frac": gbs / hbm_peak}


def gpu_baseline(device, batch, steps=20, warmup=5):
    """The kernel to beat: the reference's algorithm through STOCK PyTorch / cuDNN / cuBLAS on
    the same B200 -- the oracle's functional restatement (F.conv2d, F.batch_norm, F.max_pool2d, F.interpolate, the
    MS-SSIM pyramid, autograd, torch.optim.Adam) with its tensors on the GPU.  Variants: fp32 (TF32 off), TF32,
    bf16 autocast + channels_last; each eager and as one CUDA graph.  None of this repo's kernels run here."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import critic_vae_oracle as O
    import synth
    out = {"batch": batch, "what": "oracle/critic_vae_oracle.py (torch functional ops + autograd) on cuda, torch.optim.Adam(capturable)",
           "variants": {}}
    crit = {k: v.to(device) for k, v in torch.load(CRITIC_CKPT, map_location="cpu").items()}
    x = synth.make_frames(64, seed=100).repeat((batch + 63) // 64, 1, 1, 1)[:batch].to(device)
    eps = torch.randn(batch, 32, device=device)

    def make_step(mode):
        enc, dec = synth.make_vae_state(0)
        enc = {k: v.to(device) for k, v in enc.items()}
        dec = {k: v.to(device) for k, v in dec.items()}
        params = [enc[k].requires_grad_(True) for k in O.PARAM_KEYS_ENC] + [dec[k].requires_grad_(True) for k in O.PARAM_KEYS_DEC]
        opt = torch.optim.Adam(params, lr=5e-5, capturable=True)
        xin = x.contiguous(memory_format=torch.channels_last) if mode == "bf16" else x

        def step():
            opt.zero_grad(set_to_none=False)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16")):
                with torch.no_grad():
                    pred = O.critic_forward(crit, xin)
                _, mu, logvar, recon = O.vae_forward(enc, dec, xin, pred, eps, training=True, update_stats=True)
            losses = O.vae_loss(x, mu.float(), logvar.float(), recon.float())
            losses["total_loss"].backward()
            opt.step()
            return losses["total_loss"]
        return step

    def timed(fn, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    old_tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    old_bench = torch.backends.cudnn.benchmark
    torch.backends.cudnn.benchmark = True
    try:
        for mode in ("fp32", "tf32", "bf16"):
            torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = (mode != "fp32")
            for graphed in (False, True):
                name = f"{mode}_{'graph' if graphed else 'eager'}"
                try:
                    # parameters, warm-up and capture all live on ONE side stream: autograd's AccumulateGrad nodes run on the
                    # stream their leaf was created on, and a node on the default stream invalidates the capture
                    side = torch.cuda.Stream()
                    side.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(side):
                        step = make_step(mode)
                        for _ in range(warmup):
                            loss = step()
                        fn = step
                        if graphed:
                            side.synchronize()
                            g = torch.cuda.CUDAGraph()
                            # torch.prod's backward looks for zeros on the host (a sync, illegal during capture): for the graphed
                            # variants the 4-element product of the MS-SSIM levels is spelled out as multiplications
                            orig_prod = torch.prod
                            torch.prod = lambda t, *a, **k: functools.reduce(operator.mul, t.unbind(0)) if not a and not k and t.dim() == 1 else orig_prod(t, *a, **k)
                            try:
                                with torch.cuda.graph(g, stream=side):
                                    loss = step()
                            finally:
                                torch.prod = orig_prod
                            fn = g.replay
                            fn()
                        side.synchronize()
                        ms = timed(fn, steps)
                    out["variants"][name] = {"ms_per_step": ms, "frames_per_s": batch / (ms * 1e-3), "loss": float(loss.detach())}
                except Exception as exc:      # a variant that stock PyTorch cannot run (e.g. graph capture) is reported, not fatal
                    out["variants"][name] = {"error": repr(exc)[:160]}
                torch.cuda.synchronize()
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old_tf32
        torch.backends.cudnn.benchmark = old_bench
    ok = [v for v in out["variants"].values() if "ms_per_step" in v]
    if ok:
        best = min(ok, key=lambda v: v["ms_per_step"])
        out["best_ms_per_step"] = best["ms_per_step"]
        out["best_frames_per_s"] = best["frames_per_s"]
    return out


def dataset_secondary(device):
    """BASELINE.json configs[4]: the `-dataset` path (vae_utility.py:416-443) on 8192 host frames per call: critic value of
    every frame, balanced selection, reconstructions with the critic value and with 0 of the selection."""
    import synth
    import vae_utility as U
    vae, critic = build_modules(device)
    vae.eval()
    n = 8192
    povs = (synth.make_frames(256, seed=50).permute(0, 2, 3, 1) * 255).round().to(torch.uint8).numpy()
    povs = np.tile(povs, (n // 256, 1, 1, 1))
    U.dataset_from_trajectory(povs, critic, recon_dset=True, vae=vae)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        ds = U.dataset_from_trajectory(povs, critic, recon_dset=True, vae=vae)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    return {"frames_per_s": n / dt, "frames_per_call": n, "reconstructions_per_call": len(ds),
            "workload": "BASELINE.json configs[4]: host uint8 frames -> critic scores (8192) -> balanced selection -> recon(pred), recon(0) "
                        "of the selection -> host float32 arrays (vae_utility.dataset_from_trajectory)"}


def run_b200(args):
    import torch.distributed as dist
    import synth
    from cvae_native.trainer import TrainStep
    from cvae_native import binding as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU path for the product arm")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    pg = None
    if world > 1:
        # one box, NVLink / NVSwitch only: never fall back to the host network
        os.environ.setdefault("NCCL_P2P_LEVEL", "NVL")
        os.environ.setdefault("NCCL_IB_DISABLE", "1")
        dist.init_process_group("nccl", device_id=device)
        pg = dist.group.WORLD
    B, R = args.batch, 4
    vae, critic = build_modules(device)
    step = TrainStep(vae, critic, B, lr=5e-5, process_group=pg)

    # R rotating batches: device-resident fp32 (value) and pinned-host uint8 (e2e)
    frames = synth.make_frames(R * 64, seed=1000 + rank)                       # 64 distinct frames per slot, tiled to B
    reps = (B + 63) // 64
    dev_batches, host_batches = [], []
    for r in range(R):
        f = frames[r * 64:(r + 1) * 64].repeat(reps, 1, 1, 1)[:B].contiguous()
        dev_batches.append(f.to(device))
        host_batches.append((f.permute(0, 2, 3, 1) * 255).round().to(torch.uint8).contiguous().pin_memory())
    gen = torch.Generator(device=device).manual_seed(7 + rank)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    def one_step(i):
        step.load(frames=dev_batches[i % R])
        step.eps.normal_(generator=gen)
        return step.run()

    for i in range(max(args.warmup, 3)):
        one_step(i)
    vae._engine.check_fault()

    # ---- timed: inputs resident in HBM -------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        losses = one_step(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    final_loss = [float(v) for v in losses.cpu()]
    vae._engine.check_fault()

    # ---- timed: end to end through the product's loader (pinned host uint8 frames -> H2D -> step -> D2H loss) ------
    from cvae_native.loader import FrameStager
    stager = FrameStager(step, eps_generator=gen)

    def e2e_loop(n):
        seen = 0.0
        for loss in stager.run(host_batches[i % R] for i in range(n)):
            seen += float(loss[0])                      # the host reads every step's loss
        return seen

    e2e_loop(max(args.warmup, 3))
    barrier()
    t0 = time.perf_counter()
    e2e_loop(args.steps)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=device)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = e2e_s.item()
    clocks = sampler.summary() if rank == 0 else None

    launches_per_step = int(step.launches_per_step)
    # ---- per-kernel-family CUDA-event pass (eager, same stream) for the roofline -----------------------
    eng = vae._engine
    fam = {}
    for it in range(6):
        eng.profile = []
        step.load(frames=dev_batches[it % R])
        step._eager(False)
        torch.cuda.synchronize()
        if it >= 2:
            for name, a, b in eng.profile:
                fam.setdefault(name, []).append(a.elapsed_time(b))
    eng.profile = None
    n_it = 4
    fam_ms = {k: sum(v) / n_it for k, v in fam.items()}              # ms per step per family
    fam_calls = {k: len(v) // n_it for k, v in fam.items()}
    tf_peak, hbm_peak, peak_src = peaks()
    traffic, traffic_src = ncu_traffic()
    flops = {"conv_gemm": B * (2 * CONV_FLOPS - E0_FLOPS), "conv_wgrad": B * CONV_FLOPS}
    dominant = max(fam_ms, key=fam_ms.get)
    ach = flops[dominant] / (fam_ms[dominant] * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": dominant, "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak,
                "traffic": traffic.get(dominant) if B == 256 else None, "traffic_unit": f"bytes per launch (ncu --set full, {traffic_src})",
                "peak_source": peak_src,
                "launches_per_step": fam_calls[dominant], "ms_per_step_in_kernel": fam_ms[dominant],
                "families": {k: {"ms_per_step": fam_ms[k], "launches": fam_calls[k],
                                 "achieved_tflops": flops[k] / (fam_ms[k] * 1e-3) / 1e12} for k in fam_ms}}

    # ---- BASELINE.json configs[3]: global batch 4096 split over the N GPUs (strong scaling), N > 1 only ----------------
    cfg4 = None
    if world > 1 and 4096 % world == 0 and not args.no_secondary:
        Bg = 4096 // world
        del step
        torch.cuda.empty_cache()
        step4 = TrainStep(vae, critic, Bg, lr=5e-5, process_group=pg)
        f4 = frames[:64].repeat((Bg + 63) // 64, 1, 1, 1)[:Bg].contiguous().to(device)
        step4.load(frames=f4)
        for _ in range(3):
            step4.eps.normal_(generator=gen)
            step4.run()
        barrier()
        n4 = max(5, min(args.steps, 25))       # 25 steps of 4096 = the 100k-frame epoch of configs[3]
        e0.record()
        for _ in range(n4):
            step4.eps.normal_(generator=gen)
            step4.run()
        e1.record()
        barrier()
        t4 = torch.tensor([e0.elapsed_time(e1)], device=device)
        dist.all_reduce(t4, op=dist.ReduceOp.MAX)
        cfg4 = {"global_batch": 4096, "per_gpu_batch": Bg, "steps": n4, "ms_per_step": t4.item() / n4,
                "frames_per_s": 4096 * n4 / (t4.item() * 1e-3), "scaling": "strong",
                "workload": "BASELINE.json configs[3]: data-parallel training step, global batch 4096 (25 steps = one pass over 100k "
                            "synthetic frames), NCCL gradient all-reduce"}
        vae._engine.check_fault()
        step = step4
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu_threads = os.cpu_count() or 1
    cpu_rate, cpu_sec, cpu_steps = cpu_train_step_rate(B, 2, 1, cpu_threads, budget_s=12.0)
    step_ms = ms / args.steps
    value = world * B * args.steps / (ms * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": "BASELINE.json configs[1]: Critic-VAE training step (critic+fwd+MS-SSIM/KLD+bwd+Adam), "
                               f"batch {B} per GPU, 64x64x3 synthetic frames, random-init VAE (seed 0), shipped critic",
                   "per_gpu_batch": B, "global_batch": world * B, "parallelism": f"dp{world}" if world > 1 else "single",
                   "l2": f"{R} rotating input batches; per-step activation+gradient working set ~{B * 1.6:.0f} MB > 126 MB L2",
                   "step_flops_algorithmic": B * TRAIN_FLOPS, "final_loss": final_loss},
        "clocks": clocks,
        "e2e": {"value": world * B * args.steps / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": stager.h2d_bytes_per_step,
                "d2h_bytes_per_step": stager.d2h_bytes_per_step,
                "input": "uint8 HWC frames in pinned host memory through cvae_native.loader.FrameStager (double-buffered H2D on a copy "
                         "stream, uint8 -> fp32 inside the step graph), every step's loss read on the host"},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": roofline,
        "step_tensor_frac": (B * TRAIN_FLOPS / (step_ms * 1e-3) / 1e12) / tf_peak,
        "cpu_baseline": {"value": cpu_rate, "unit": "frames/s", "cores": cpu_threads, "kind": "port",
                         "sample": f"{cpu_steps} training steps of batch {B} in ~{cpu_steps * cpu_sec:.0f} s (oracle/critic_vae_oracle.py, torch CPU fp32), same step definition"},
    }
    if cfg4 is not None:
        line["cfg4"] = cfg4
    if world == 1 and not args.no_secondary:
        try:
            line["gpu_baseline"] = gpu_baseline(device, B)
            if "best_frames_per_s" in line["gpu_baseline"]:
                line["gpu_baseline"]["speedup_vs_best_stock_pytorch"] = value / line["gpu_baseline"]["best_frames_per_s"]
        except Exception as exc:
            line["gpu_baseline"] = {"error": repr(exc)[:200]}
        try:
            line["dataset_path"] = dataset_secondary(device)
        except Exception as exc:
            line["dataset_path"] = {"error": repr(exc)[:200]}
        try:
            del step
            torch.cuda.empty_cache()
            line["mask_iou"] = mask_iou_secondary(device, hbm_peak)
        except Exception as exc:   # the secondary metrics must never cost the headline line
            line["mask_iou"] = {"error": repr(exc)[:200]}
        try:
            torch.cuda.empty_cache()
            line["latent_kernel"] = latent_secondary(device, hbm_peak)
        except Exception as exc:
            line["latent_kernel"] = {"error": repr(exc)[:200]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="frames per GPU per step")
    ap.add_argument("--no-secondary", action="store_true", help="skip the mask+IoU secondary measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
