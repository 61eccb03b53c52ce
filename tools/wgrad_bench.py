"""Time every conv_wgrad configuration of one training step in isolation (random operands).
usage: python tools/wgrad_bench.py [batch]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tests", "critic-vae_b200"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch
from cvae_native import binding as L

B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 256
ONLY = sys.argv[sys.argv.index("--only") + 1].split(",") if "--only" in sys.argv else None
dev, bf = "cuda", torch.bfloat16
# name, kind, H (grid of the GEMM's K pixels), cin, cout
LAYERS = [("E0w", L.WGRAD_SHIFT_FRAMES, 64, 3, 32), ("E1w", L.WGRAD_5X5, 32, 32, 64), ("E2w", L.WGRAD_5X5, 16, 64, 128),
          ("E3w", L.WGRAD_5X5, 8, 128, 256), ("D0w", L.WGRAD_5X5, 4, 256, 128), ("D1w", L.WGRAD_PHASE, 4, 128, 64),
          ("D2w", L.WGRAD_PHASE, 8, 64, 32), ("D3w", L.WGRAD_PHASE, 16, 32, 32), ("D4w", L.WGRAD_SHIFT_PHASE12, 32, 32, 3)]
ws = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
tot = 0.0
for name, kind, H, cin, cout in LAYERS:
    if ONLY and name not in ONLY:
        continue
    if kind == L.WGRAD_SHIFT_FRAMES:
        x = torch.rand(B, 3, H, H, device=dev); dy = torch.randn(B, H, H, cout, device=dev).to(bf); dy2 = None
    elif kind == L.WGRAD_5X5:
        x = torch.randn(B, H, H, cin, device=dev).to(bf); dy = torch.randn(B, H, H, cout, device=dev).to(bf); dy2 = None
    elif kind == L.WGRAD_PHASE:
        x = torch.randn(B, H, H, cin, device=dev).to(bf); dy = torch.randn(B, 2 * H, 2 * H, cout, device=dev).to(bf); dy2 = None
    else:
        x = torch.randn(B, H, H, cin, device=dev).to(bf); dy = torch.randn(B, 3, 2 * H, 2 * H, device=dev)
        dy2 = torch.rand(B, 3, 2 * H, 2 * H, device=dev)
    dw = torch.empty(cout, cin, 5, 5, device=dev); db = torch.empty(cout, device=dev)
    d = L.WgradDesc(kind=kind, batch=B, height=H, width=H, cout=cout, cin=cin, splits=int(os.environ.get("CVAE_SPLITS", "0")),
                    x=x.data_ptr(), dy=dy.data_ptr(), dy2=dy2.data_ptr() if dy2 is not None else None, dw=dw.data_ptr(),
                    dbias=None if name.startswith("E") else db.data_ptr(),      # (encoder convs sit in front of BatchNorm: no bias gradient, as in the engine)
                    workspace=ws.data_ptr())
    assert L.lib.cvae_conv_wgrad_workspace_bytes(ctypes.byref(d)) <= ws.numel(), L.lib.cvae_conv_wgrad_workspace_bytes(ctypes.byref(d))
    L.check(L.lib.cvae_conv_wgrad(ctypes.byref(d), L.stream_ptr()))
    torch.cuda.synchronize()
    L.check(L.lib.cvae_check_device_fault(L.stream_ptr()))
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g, iters = torch.cuda.CUDAGraph(), 10
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for _ in range(iters):
                L.lib.cvae_conv_wgrad(ctypes.byref(d), L.stream_ptr())
        g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(side); g.replay(); b.record(side)
    torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / iters
    if os.environ.get("CVAE_COUNTERS"):
        buf = torch.zeros(4096 * 8, dtype=torch.int64, device=dev)
        L.lib.cvae_wgrad_debug_counters(buf.data_ptr())
        L.lib.cvae_conv_wgrad(ctypes.byref(d), L.stream_ptr())
        torch.cuda.synchronize()
        L.lib.cvae_wgrad_debug_counters(None)
        c = buf.view(4096, 8).cpu().double()
        c = c[c[:, 0] > 0]
        m = c.mean(0)
        print(f"   {name} ctas={c.shape[0]} MMA thread total {m[0]:.0f} cyc (max {c[:,0].max():.0f}), wait data {m[1]:.0f} | loaders: wait buffer {m[3]:.0f}, "
              f"issue {m[4]:.0f}, landing {m[5]:.0f}, chunks {m[6]:.1f} | epilogue {m[7]:.0f}")
    flops = 2.0 * B * H * H * cout * cin * 25 * (4 if kind in (L.WGRAD_PHASE, L.WGRAD_SHIFT_PHASE12) else 1)
    tot += us
    print(f"{name}: {us:7.1f} us (kernel + fold)  {flops / us * 1e-6:7.1f} TF/s useful", flush=True)
print(f"sum {tot:.1f} us")
