"""A few eager (non-graph) training steps at batch 256, for ncu captures of individual kernels."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
for p in ("tests", "critic-vae_b200"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch
import bench, synth
from cvae_native.trainer import TrainStep

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
vae, critic = bench.build_modules(torch.device("cuda", 0))
st = TrainStep(vae, critic, B, use_graph=False)
x = synth.make_frames(64, seed=1).repeat((B + 63) // 64, 1, 1, 1)[:B].cuda()
for i in range(steps):
    st.load(frames=x, eps=synth.make_eps(B, seed=i).cuda())
    st.run()
torch.cuda.synchronize()
vae._engine.check_fault()
print("losses", st.losses.tolist())
