"""Time every conv_gemm configuration of one training step in isolation (random operands).
usage: python tools/conv_bench.py [batch] [--sweep] [--only NAME]
--sweep tries every (n_block, tm) the kernel accepts and prints the best per layer."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tests", "critic-vae_b200"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch
from cvae_native import binding as L

args = [a for a in sys.argv[1:] if not a.startswith("--")]
B = int(args[0]) if args and args[0].isdigit() else 256
SWEEP = "--sweep" in sys.argv
WA = "--wa" in sys.argv          # run the layers the weights-as-A kernel covers through it (CVAE_KTAB_BLOCK64)
WA_LAYERS = {"E1f": (32, 2), "E2f": (64, 1), "E3f": (64, 1), "D0f": (64, 1), "D1f": (64, 1), "D2f": (64, 1), "D3f": (32, 1),
             "E1g": (64, 4), "E2g": (64, 2), "E3g": (64, 1), "D0g": (64, 1), "D1g": (64, 1), "D2g": (32, 2), "D3g": (32, 4)}
if os.environ.get("CVAE_WA_ONLY") is not None:
    WA_LAYERS = {k: v for k, v in WA_LAYERS.items() if k in os.environ["CVAE_WA_ONLY"].split(",")}
_WS = {}
ONLY = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None
dev = "cuda"
bf = torch.bfloat16

# name, H, ksize, src_channels, n_total, loader, epilogue, real MACs per output pixel row (for TFLOP/s)
LAYERS = [
    ("E0f", 64, 5, 8, 32, L.LOAD_NCHW3, L.EPI_STATS),
    ("D4g", 32, 3, 16, 32, L.LOAD_S2D_NCHW3_DTANH, L.EPI_MASK),
    ("E1f", 32, 5, 32, 64, L.LOAD_NHWC, L.EPI_STATS),
    ("E2f", 16, 5, 64, 128, L.LOAD_NHWC, L.EPI_STATS),
    ("E3f", 8, 5, 128, 256, L.LOAD_NHWC, L.EPI_STATS),
    ("D0f", 4, 5, 256, 128, L.LOAD_NHWC, L.EPI_BIAS_RELU),
    ("D1f", 4, 3, 128, 256, L.LOAD_NHWC, L.EPI_PHASE_BIAS_RELU),
    ("D2f", 8, 3, 64, 128, L.LOAD_NHWC, L.EPI_PHASE_BIAS_RELU),
    ("D3f", 16, 3, 32, 128, L.LOAD_NHWC, L.EPI_PHASE_BIAS_RELU),
    ("D4f", 32, 3, 32, 16, L.LOAD_NHWC, L.EPI_PHASE_BIAS_TANH),
    ("D3g", 16, 3, 128, 32, L.LOAD_S2D, L.EPI_MASK),
    ("D2g", 8, 3, 128, 64, L.LOAD_S2D, L.EPI_MASK),
    ("D1g", 4, 3, 256, 128, L.LOAD_S2D, L.EPI_MASK),
    ("D0g", 4, 5, 128, 256, L.LOAD_NHWC, L.EPI_PLAIN),
    ("E3g", 8, 5, 256, 128, L.LOAD_NHWC, L.EPI_PLAIN),
    ("E2g", 16, 5, 128, 64, L.LOAD_NHWC, L.EPI_PLAIN),
    ("E1g", 32, 5, 64, 32, L.LOAD_NHWC, L.EPI_PLAIN),
]


def bench(name, H, k, C, N, loader, epi, tm=0, nblk=0, iters=20):
    ktab = L.KTAB_PAIR8 if loader == L.LOAD_NCHW3 else L.KTAB_GENERIC
    wa = WA and name in WA_LAYERS
    ksteps = L.lib.cvae_conv_ksteps(k, C, ktab)
    stack = 0
    if wa:
        ktab = L.KTAB_BLOCK64
        kb, stack = WA_LAYERS[name]
        groups = k * k if stack == 1 else {(5, 2): 15, (5, 4): 10, (3, 2): 6, (3, 4): 3}[(k, stack)]
        ksteps = (C // kb) * groups * (kb // 16) * stack      # x stack: the packed GEMM has N * stack rows
    src2 = None
    if loader == L.LOAD_NCHW3:
        src = torch.rand(B, 3, H, H, device=dev)
    elif loader == L.LOAD_S2D_NCHW3_DTANH:
        src = torch.randn(B, 3, 2 * H, 2 * H, device=dev)
        src2 = torch.rand(B, 3, 2 * H, 2 * H, device=dev)
    elif loader == L.LOAD_S2D:
        src = torch.randn(B, 2 * H, 2 * H, C // 4, device=dev).to(bf)
    else:
        src = torch.randn(B, H, H, C, device=dev).to(bf)
    wp = (torch.randn(N * ksteps * 16, device=dev) * 0.05).to(bf)
    if epi == L.EPI_PHASE_BIAS_TANH:
        out = torch.empty(B, 3, 2 * H, 2 * H, device=dev)
    elif epi == L.EPI_PHASE_BIAS_RELU:
        out = torch.empty(B, 2 * H, 2 * H, N // 4, device=dev, dtype=bf)
    else:
        out = torch.empty(B, H, H, N, device=dev, dtype=bf)
    bias = torch.zeros(256, device=dev)
    act = torch.randn(B, H, H, N, device=dev).to(bf) if epi == L.EPI_MASK else None
    stats = torch.zeros(2 * N, dtype=torch.float64, device=dev) if epi == L.EPI_STATS else None
    d = L.ConvDesc(batch=B, height=H, width=H, ksize=k, src_channels=C, n_total=N, loader=loader, epilogue=epi,
                   ktab=ktab, tm=tm, n_block=nblk, stack=stack, src=src.data_ptr(), src2=src2.data_ptr() if src2 is not None else None,
                   wpack=wp.data_ptr(), bias=bias.data_ptr(),
                   act=act.data_ptr() if act is not None else None, out=out.data_ptr(),
                   stats=stats.data_ptr() if stats is not None else None)
    s = L.stream_ptr()
    need = int(L.lib.cvae_conv_gemm_workspace_bytes(ctypes.byref(d)))
    if need > 0:
        if "b" not in _WS or _WS["b"].numel() < need:
            _WS["b"] = torch.zeros(need, dtype=torch.uint8, device=dev)
        d.workspace, d.workspace_bytes = _WS["b"].data_ptr(), _WS["b"].numel()
    rc = L.lib.cvae_conv_gemm(ctypes.byref(d), s)
    if rc != 0:
        return None
    torch.cuda.synchronize()
    L.check(L.lib.cvae_check_device_fault(s))
    # time GPU work only: capture `iters` launches in a CUDA graph (the eager loop is bound by the host launch path)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for _ in range(iters):
                L.lib.cvae_conv_gemm(ctypes.byref(d), L.stream_ptr())
        g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(side)
        g.replay()
        b.record(side)
    torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / iters
    if os.environ.get("CVAE_COUNTERS") and wa:
        buf = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
        L.lib.cvae_conv_wa_debug_counters(buf.data_ptr())
        L.lib.cvae_conv_gemm(ctypes.byref(d), s)
        torch.cuda.synchronize()
        L.lib.cvae_conv_wa_debug_counters(None)
        c = buf.view(148, 16).cpu().double()
        c = c[c[:, 0] > 0]
        m = c.mean(0)
        t0 = c[:, 8].min()
        ph = lambda k: (c[:, k] - t0).mean().item() * 1e-3
        print(f"   {name} [wa phases, us from first CTA start, mean over CTAs] entry {ph(8):.1f}  prologue done {ph(9):.1f}  first MMA {ph(10):.1f}  "
              f"last MMA issued {ph(11):.1f}  epilogue done {ph(12):.1f}  exit {ph(13):.1f} (last CTA exit {(c[:, 13].max() - t0) * 1e-3:.1f})")
        print(f"   {name} [wa] ctas={c.shape[0]} MMA thread: total {m[0]:.0f} cyc (max {c[:, 0].max():.0f}), wait acc {m[1]:.0f}, pixels {m[2]:.0f}, "
              f"weights {m[3]:.0f}, items {m[4]:.1f}, SM clock {m[0] / max(m[5], 1) * 1e3:.0f} MHz | epilogue {m[6]:.0f} (split-K reduce max {c[:, 7].max():.0f})")
    elif os.environ.get("CVAE_COUNTERS"):
        buf = torch.zeros(148 * 8, dtype=torch.int64, device=dev)
        L.lib.cvae_conv_debug_counters(buf.data_ptr())
        L.lib.cvae_conv_gemm(ctypes.byref(d), s)
        torch.cuda.synchronize()
        L.lib.cvae_conv_debug_counters(None)
        c = buf.view(148, 8).cpu().double()
        c = c[c[:, 0] > 0]
        m = c.mean(0)
        mmas = m[4] * (ksteps // 1) * d.tm if d.tm else 0
        print(f"   {name} ctas={c.shape[0]} issue cycles/MMA {(m[0]-m[1]-m[2]-m[3])/max(mmas,1):.1f} | MMA thread: total {m[0]:.0f} cyc, wait acc {m[1]:.0f}, planes {m[2]:.0f}, weights {m[3]:.0f}, "
              f"items {m[4]:.1f} | producer fill {m[5]:.0f} | elected-issue region {m[6]:.0f} | epilogue {m[7]:.0f}")
    flops = 2.0 * B * H * H * N * k * k * C      # issued-on-real-pixels FLOPs of this GEMM form
    return us, flops / us * 1e-6


if os.environ.get("WA_TUNE"):
    L.lib.cvae_conv_wa_tune(*[int(v) for v in os.environ["WA_TUNE"].split(",")])
tot = 0.0
for (name, H, k, C, N, loader, epi) in LAYERS:
    if ONLY and name != ONLY:
        continue
    r = bench(name, H, k, C, N, loader, epi, tm=int(os.environ.get("CVAE_TM", "0")), nblk=int(os.environ.get("CVAE_NBLK", "0")),
              iters=int(os.environ.get("CVAE_ITERS", "20")))
    if r is None:
        print(name, "FAILED:", L.lib.cvae_last_error().decode(), flush=True)
        continue
    line = f"{name}: auto {r[0]:7.1f} us {r[1]:7.1f} TF/s"
    best = r
    if SWEEP:
        res = []
        for nblk in (16, 32, 64, 128):
            for tm in range(1, 9):
                if os.environ.get("CVAE_TRACE"):
                    print("try", name, nblk, tm, flush=True)
                rr = bench(name, H, k, C, N, loader, epi, tm=tm, nblk=nblk, iters=10)
                if rr:
                    res.append((rr[0], nblk, tm))
        res.sort()
        line += "  | best " + ", ".join(f"n{n}/tm{t}:{u:.1f}" for u, n, t in res[:5])
        if res and res[0][0] < best[0]:
            best = (res[0][0], 0)
    if WA and name in WA_LAYERS and "--wa-sweep" in sys.argv:
        res = []
        for ksp in (1, 2, 4):
            for ups in (1, 2, 4):
                for grid in (0, 148, 144, 132, 112, 96, 74, 72, 64, 48, 36):
                    if grid % ksp:
                        continue
                    L.lib.cvae_conv_wa_tune(1, grid, ups, 0, ksp, int(os.environ.get("WA_WLOAD", "0")))
                    rr = bench(name, H, k, C, N, loader, epi, iters=10)
                    if rr:
                        res.append((rr[0], ksp, grid, ups))
        L.lib.cvae_conv_wa_tune(0, 0, 0, 0, 0, 0)
        res.sort()
        line += "  | best " + ", ".join(f"k{c}/g{g}/u{u}:{t:.1f}" for t, c, g, u in res[:6])
        if res and res[0][0] < best[0]:
            best = (res[0][0], 0)
    tot += best[0]
    print(line, flush=True)
print(f"sum {tot:.1f} us")
