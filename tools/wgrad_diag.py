"""Diagnose the order-dependent weight-gradient mismatch: run the 5x5 cases in test order, report where the errors sit."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tests", "critic-vae_b200"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch
from test_conv_gemm import _native, _rand, nhwc_bf16, rb
from test_conv_wgrad import run_wgrad
L = _native()
cases = [(3, 32, 64, 32, 0), (4, 64, 128, 16, 0), (5, 128, 256, 8, 0), (6, 256, 128, 4, 0), (3, 32, 64, 32, 1), (7, 64, 128, 16, 3), (41, 128, 256, 8, 0),
         (41, 128, 256, 8, 0)]
skip = int(os.environ.get("DIAG_SKIP", "0"))
cases = cases[skip:]
if os.environ.get("DIAG_PRED"):        # one predecessor, then the failing case
    cases = [cases[int(os.environ["DIAG_PRED"])], (41, 128, 256, 8, 0)]
for (B, Cin, Cout, HW, splits) in cases:
    x, dy = rb(_rand((B, Cin, HW, HW), 31)), rb(_rand((B, Cout, HW, HW), 32))
    ref_w = torch.nn.grad.conv2d_weight(x.double(), (Cout, Cin, 5, 5), dy.double(), padding=2).float()
    ref_b = dy.double().sum((0, 2, 3)).float()
    dw, db = run_wgrad(L, L.WGRAD_5X5, B, HW, HW, Cout, Cin, nhwc_bf16(x), nhwc_bf16(dy), splits=splits)
    bad = (dw - ref_w).abs() > 1e-3 * ref_w.abs().max()
    badb = (db - ref_b).abs() > 1e-3 * ref_b.abs().max()
    print(f"case {(B, Cin, Cout, HW, splits)}: bad dw {int(bad.sum())} / {bad.numel()}, bad db {int(badb.sum())} / {badb.numel()}, nan {int(torch.isnan(dw).sum())}")
    if bad.any():
        print("  by tap (ky,kx):", bad.sum((0, 1)).tolist())
        co = bad.sum((1, 2, 3)); ci = bad.sum((0, 2, 3))
        print("  co with errors:", int((co > 0).sum()), "first", torch.nonzero(co > 0).flatten()[:12].tolist(), " ci with errors:", int((ci > 0).sum()),
              "first", torch.nonzero(ci > 0).flatten()[:12].tolist())
        print("  bad db idx:", torch.nonzero(badb).flatten()[:16].tolist())

# --- deeper look at the last (failing) case: is the wrong part a sum over the wrong set of images?
if bad.any():
    xs, dys = x.double(), dy.double()
    per = torch.stack([torch.nn.grad.conv2d_weight(xs[n:n + 1], (Cout, Cin, 5, 5), dys[n:n + 1], padding=2)[:, :, 4, 4].reshape(-1) for n in range(B)], 1)  # [co*ci][B]
    got = dw[:, :, 4, 4].reshape(-1).double()
    for name, sel in (("mb0 (co<128)", slice(0, 128 * Cin)), ("mb1 (co>=128)", slice(128 * Cin, 256 * Cin))):
        sol = torch.linalg.lstsq(per[sel], got[sel, None]).solution.flatten()
        res = (per[sel] @ sol - got[sel]).norm() / got[sel].norm()
        print(f"  tap (4,4) {name}: image coefficients {[round(v, 2) for v in sol.tolist()]} residual {res:.2e}")
    b4 = bad[:, :, 4, 4]
    print("  tap (4,4): co rows wrong:", [int(v) for v in torch.nonzero(b4.any(1)).flatten()[:8]], "... count", int(b4.any(1).sum()),
          "; ci cols wrong count", int(b4.any(0).sum()))
    perb = dys.sum((2, 3)).t()      # [co][B]
    solb = torch.linalg.lstsq(perb, db.double()[:, None]).solution.flatten()
    print("  bias: image coefficients", [round(v, 2) for v in solb.tolist()])
