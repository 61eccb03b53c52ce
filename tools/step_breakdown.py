"""Per-C-ABI-call GPU time of one eager training step (CUDA events around every libcvae call).
usage: python tools/step_breakdown.py [batch] [iters]"""
import ctypes, os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
for p in ("tests", "critic-vae_b200"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch
import bench, synth
from cvae_native import binding as L
from cvae_native.trainer import TrainStep

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
records, enabled = [], [False]


class Timed:
    def __init__(self, name, fn):
        self.name, self.fn = name, fn

    def __call__(self, *args):
        if not enabled[0]:
            return self.fn(*args)
        label = self.name
        if self.name == "cvae_conv_gemm":
            d = args[0]._obj
            label += f" L{d.loader} E{d.epilogue} {d.height}x{d.width} C{d.src_channels}->N{d.n_total} k{d.ksize}"
        elif self.name == "cvae_conv_wgrad":
            d = args[0]._obj
            label += f" kind{d.kind} {d.height}x{d.width} cin{d.cin} cout{d.cout}"
        elif self.name in ("cvae_bn_pool_act_fwd", "cvae_bn_pool_act_bwd"):
            label += f" {args[1]}x{args[2]} C{args[3]}"
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = self.fn(*args)
        b.record()
        records.append((label, a, b))
        return r


class LibProxy:
    def __init__(self, lib):
        object.__setattr__(self, "_lib", lib)
        object.__setattr__(self, "_cache", {})

    def __getattr__(self, name):
        c = self._cache
        if name not in c:
            fn = getattr(self._lib, name)
            c[name] = Timed(name, fn) if name.startswith("cvae_") and name not in ("cvae_last_error", "cvae_launch_count") else fn
        return c[name]


L.lib = LibProxy(L.lib)
vae, critic = bench.build_modules(torch.device("cuda", 0))
st = TrainStep(vae, critic, B, use_graph=False)
x = synth.make_frames(64, seed=1).repeat((B + 63) // 64, 1, 1, 1)[:B].cuda()
agg = collections.OrderedDict()
for i in range(iters):
    st.load(frames=x, eps=synth.make_eps(B, seed=i).cuda())
    enabled[0] = i >= 2
    records.clear()
    st.run()
    torch.cuda.synchronize()
    for label, a, b in records:
        agg.setdefault(label, []).append(a.elapsed_time(b) * 1e3)
tot = 0.0
n = iters - 2
for label, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    us = sum(v) / n
    tot += us
    print(f"{us:9.1f} us  x{len(v) // n:2d}  {label}")
print(f"total {tot:.1f} us per step (eager, event-timed per call; includes launch gaps inside a call)")
