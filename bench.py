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
import json
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
# dram__bytes_read.sum + dram__bytes_write.sum per launch (mean over the family's launches of one step) from the
# ncu --set full capture summarised in profiles/r01_gemm_full_v2.md (batch 256)
NCU_TRAFFIC = {"conv_gemm": 13.24e6, "conv_wgrad": 27.80e6}


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
            self.stop_flag.wait(0.2)

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
    batch = 64
    rate, sec, _ = cpu_train_step_rate(batch, args.steps, args.warmup, threads)
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "Critic-VAE training step (critic+fwd+MS-SSIM/KLD+bwd+Adam), 64x64x3 synthetic frames, "
                                   "random-init VAE, shipped critic; bounded sample: batch 64 per step on host cores"},
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

    # ---- timed: end to end (pinned host uint8 frames -> H2D -> step -> D2H loss), double buffered ------
    copy_stream = torch.cuda.Stream()
    stage = [torch.empty(B, 64, 64, 3, dtype=torch.uint8, device=device) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    loss_host = [torch.empty(3, pin_memory=True) for _ in range(2)]
    loss_done = [torch.cuda.Event() for _ in range(2)]

    def issue_copy(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i % 2])
            stage[i % 2].copy_(host_batches[i % R], non_blocking=True)
            ready[i % 2].record(copy_stream)

    def e2e_loop(n):
        cur = torch.cuda.current_stream()
        for b in range(2):
            consumed[b].record(cur)
        issue_copy(0)
        seen = 0.0
        for i in range(n):
            if i + 1 < n:
                issue_copy(i + 1)
            cur.wait_event(ready[i % 2])
            step.load(frames_u8=stage[i % 2])
            consumed[i % 2].record(cur)
            step.eps.normal_(generator=gen)
            out = step.run(from_u8=True)
            loss_host[i % 2].copy_(out, non_blocking=True)
            loss_done[i % 2].record(cur)
            if i > 0:                                   # read the previous step's loss on the host
                loss_done[(i - 1) % 2].synchronize()
                seen += float(loss_host[(i - 1) % 2][0])
        loss_done[(n - 1) % 2].synchronize()
        return seen + float(loss_host[(n - 1) % 2][0])

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
    flops = {"conv_gemm": B * (2 * CONV_FLOPS - E0_FLOPS), "conv_wgrad": B * CONV_FLOPS}
    dominant = max(fam_ms, key=fam_ms.get)
    ach = flops[dominant] / (fam_ms[dominant] * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": dominant, "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak,
                "traffic": NCU_TRAFFIC.get(dominant) if B == 256 else None, "traffic_unit": "bytes per launch (ncu, profiles/r01_gemm_full_v2.md)",
                "peak_source": peak_src,
                "launches_per_step": fam_calls[dominant], "ms_per_step_in_kernel": fam_ms[dominant],
                "families": {k: {"ms_per_step": fam_ms[k], "launches": fam_calls[k],
                                 "achieved_tflops": flops[k] / (fam_ms[k] * 1e-3) / 1e12} for k in fam_ms}}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu_threads = os.cpu_count() or 1
    cpu_rate, cpu_sec, cpu_steps = cpu_train_step_rate(64, 4, 1, cpu_threads, budget_s=12.0)
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
        "e2e": {"value": world * B * args.steps / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": B * 64 * 64 * 3,
                "d2h_bytes_per_step": 12, "input": "uint8 HWC frames in pinned host memory, double-buffered H2D on a copy stream"},
        "gpu_launches": int(step.launches_per_step) * args.steps,
        "roofline": roofline,
        "step_tensor_frac": (B * TRAIN_FLOPS / (step_ms * 1e-3) / 1e12) / tf_peak,
        "cpu_baseline": {"value": cpu_rate, "unit": "frames/s", "cores": cpu_threads, "kind": "port",
                         "sample": f"{cpu_steps} training steps of batch 64 in ~12 s (oracle/critic_vae_oracle.py, torch CPU fp32), same step definition"},
    }
    if world == 1 and not args.no_secondary:
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
