"""Phase timestamps of the two fused bottleneck kernels (CTA 0, clock64): where a 30-50 us latency chain spends its time."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tests", "critic-vae_b200"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch
from cvae_native import binding as L
B = 256
dev = "cuda"
ptr = lambda t: t.data_ptr()
wfc, wdec = torch.randn(4096, 64, device=dev) * 0.02, torch.randn(34, 4096, device=dev) * 0.1
a = torch.randn(B, 4096, device=dev).to(torch.bfloat16)
bmu, bvar = torch.zeros(32, device=dev), torch.zeros(32, device=dev)
eps, pred = torch.randn(B, 32, device=dev), torch.rand(B, device=dev)
ml, zc, h = torch.empty(B, 64, device=dev), torch.empty(B, 33, device=dev), torch.empty(B, 4096, dtype=torch.bfloat16, device=dev)
dh = torch.randn(B, 4096, device=dev).to(torch.bfloat16)
dml, da = torch.empty(B, 64, device=dev), torch.empty(B, 4096, dtype=torch.bfloat16, device=dev)
dbg = torch.zeros(16 + 4 * 128, dtype=torch.int64, device=dev)
s = L.stream_ptr()
def fwd():
    L.check(L.lib.cvae_bottleneck_fwd(B, ptr(a), ptr(wfc), ptr(bmu), ptr(bvar), ptr(eps), ptr(pred), ptr(wdec), ptr(ml), ptr(zc), ptr(h), s))
def bwd():
    L.check(L.lib.cvae_bottleneck_bwd(B, ptr(dh), ptr(wdec), ptr(ml), ptr(eps), 1e-6, ptr(wfc), None, ptr(dml), ptr(da), s))
for _ in range(3):
    fwd(); bwd()
torch.cuda.synchronize()
for name, fn in (("fwd", fwd), ("bwd", bwd)):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) * 50:.1f} us per launch (back to back, warm)")
L.check(L.lib.cvae_bottleneck_debug(ptr(dbg)))
fwd(); bwd()
torch.cuda.synchronize()
L.check(L.lib.cvae_bottleneck_debug(None))
t = dbg.cpu().tolist()
names_f = ["start", "operands landed", "fc partial done", "cluster sync 1", "reduce + latent + sync 2 + gather", "wdec landed", "decoder_input slice done", "final sync"]
names_b = ["start", "operands landed", "decin partial done", "cluster sync 1", "reduce + latent + sync 2 + gather", "fc^T slice done", "final sync"]
print("forward (cycles since start of CTA 0):", [(n, t[i] - t[0]) for i, n in enumerate(names_f)])
print("backward:", [(n, t[8 + i] - t[8]) for i, n in enumerate(names_b)])
w = dbg[16:].view(128, 4).cpu().double()
w = w[w[:, 0] > 0]                       # CTAs that ran (13 clusters of 20 rows at batch 256)
print("CTAs:", w.shape[0])
for name, a, b in (("forward", 0, 1), ("backward", 2, 3)):
    t0 = w[:, a].min()
    print(f"{name}: CTA entry {((w[:, a] - t0).min() / 1e3):.1f} .. {((w[:, a] - t0).max() / 1e3):.1f} us, exit {((w[:, b] - t0).min() / 1e3):.1f} .. "
          f"{((w[:, b] - t0).max() / 1e3):.1f} us, CTA 0 lives {(w[0, b] - w[0, a]) / 1e3:.1f} us")
for which, name in ((0, "forward"), (1, "backward")):
    print(name, "max active 8-CTA clusters:", {kb: L.lib.cvae_bottleneck_max_clusters(which, kb * 1024) for kb in (0, 48, 100, 112, 150, 200)})
