"""Per-call timeline of ONE training step, warm caches: the step is run eagerly on a single stream (no graph, no side
streams) with a CUDA event pair around every libcvae call, so each kernel's duration is seen with the working set its
predecessors left in L2 -- what the ncu launch list (cold, serialised) cannot show.
usage: python tools/step_timeline.py [batch]      (prints one line per call, then totals per kernel family)"""
import ctypes, os, sys
os.environ["CVAE_NO_SIDE_STREAM"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tests", "oracle", "critic-vae_b200"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch
import synth
from cvae_native import binding as L
from cvae_native.trainer import TrainStep
from test_vae_module import _modules

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
crit = torch.load(os.path.join(ROOT, "critic-vae_b200", "saved-networks",
                               "critic-rewidx=1-cepochs=15-datamode=trunk-datasize=99999-shift=12-chfak=1-dropout=0.3.pt"), map_location="cpu")
vae, critic = _modules(crit, seed=0)
vae.train()
st = TrainStep(vae, critic, B, use_graph=False)
st.load(frames=synth.make_frames(64, seed=1).repeat((B + 63) // 64, 1, 1, 1)[:B].cuda(), eps=torch.randn(B, 32).cuda())
for _ in range(3):
    st.run()
torch.cuda.synchronize()

records = []
def wrap(name, fn):
    def call(*args):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        tag = ""
        if name in ("cvae_conv_gemm", "cvae_conv_wgrad") and args and hasattr(args[0], "_obj"):
            d = args[0]._obj
            tag = " ".join(f"{f}={getattr(d, f)}" for f in ("height", "ksize", "src_channels", "n_total", "cout", "cin", "kind", "epilogue", "ktab") if hasattr(d, f))
        elif args and isinstance(args[0], int):
            tag = " ".join(str(a) for a in args[:5] if isinstance(a, int))
        records.append((name, tag, e0, e1))
        return rc
    return call

for name in list(L.EXPORTS):
    fn = getattr(L.lib, name)
    if name in ("cvae_last_error", "cvae_version", "cvae_check_device_fault") or "workspace" in name or "debug" in name or "tune" in name or "ksteps" in name or "count" in name or "partials" in name:
        continue
    setattr(L.lib, name, wrap(name, fn))

t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
st.run()
t1.record()
torch.cuda.synchronize()
fam = {}
prev = None
for name, tag, e0, e1 in records:
    us = e0.elapsed_time(e1) * 1e3
    gap = prev.elapsed_time(e0) * 1e3 if prev is not None else 0.0
    prev = e1
    fam[name] = fam.get(name, 0.0) + us
    print(f"{us:8.1f} us  (gap {gap:5.1f})  {name:28s} {tag}")
print("--- per entry point")
for name, us in sorted(fam.items(), key=lambda kv: -kv[1]):
    print(f"{us:8.1f} us  {name}")
print(f"sum of calls {sum(fam.values()):.1f} us; whole eager step {t0.elapsed_time(t1) * 1e3:.1f} us (single stream, includes host launch gaps)")
