"""Minimal data-parallel smoke run (torchrun): NCCL init, TrainStep with a process group, a few graph-replayed
steps, loss printed by every rank.  Progress goes to stderr so a hang shows where it stopped."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
for p in ("tests", "critic-vae_b200"):
    sys.path.insert(0, os.path.join(ROOT, p))
t00 = time.time()


def log(msg):
    print(f"[rank {os.environ.get('RANK', '0')} +{time.time() - t00:5.1f}s] {msg}", file=sys.stderr, flush=True)


import torch
import torch.distributed as dist
log("torch imported")
import bench, synth
from cvae_native.trainer import TrainStep

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
log("process group up")
vae, critic = bench.build_modules(dev)
st = TrainStep(vae, critic, 64, process_group=dist.group.WORLD)
x = synth.make_frames(64, seed=1 + rank).to(dev)
for i in range(4):
    st.load(frames=x, eps=synth.make_eps(64, seed=i).to(dev))
    out = st.run()
    torch.cuda.synchronize()
    log(f"step {i} loss {out[0].item():.5f}")
w = vae._engine.flat[:1000].clone()
ref = w.clone()
dist.broadcast(ref, 0)
assert torch.equal(w, ref), "ranks diverged"
log("weights identical across ranks")
dist.barrier(device_ids=[local])
dist.destroy_process_group()
log("done")
