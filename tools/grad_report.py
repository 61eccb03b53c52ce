"""Diagnostic: per-parameter gradient error of the drop-in module vs the golden reference fixtures."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tests", "oracle", "critic-vae_b200"):
    sys.path.insert(0, os.path.join(ROOT, p))
import synth, vae_nets
from critic_net import Critic

g = np.load(os.path.join(ROOT, "tests/golden/train_step.npz"))
cs = torch.load(os.path.join(ROOT, "critic-vae_b200/saved-networks/critic-rewidx=1-cepochs=15-datamode=trunk-datasize=99999-shift=12-chfak=1-dropout=0.3.pt"))
vae = vae_nets.VariationalAutoencoder().to("cuda")
enc, dec = synth.make_vae_state(0)
vae.encoder.load_state_dict(enc); vae.decoder.load_state_dict(dec)
critic = Critic(); critic.load_state_dict(cs); critic.eval().to("cuda")
vae.train()
B = int(g["B"])
x, eps = synth.make_frames(B, seed=10).cuda(), synth.make_eps(B, seed=20).cuda()
out = vae(x, critic.evaluate(x), eps=eps)
losses = vae.vae_loss(*out)
losses["total_loss"].backward()
print("loss", losses["total_loss"].item(), g["s0_losses"][0])
for name, prm in vae.named_parameters():
    norm = g[f"s0_grad_norm/{name}"][0]
    gr = prm.grad.detach().flatten().double().cpu()
    idx = synth.sample_indices(gr.numel())
    ref = g[f"s0_grad_smp/{name}"]
    rel = np.linalg.norm(gr[idx].numpy() - ref) / max(np.linalg.norm(ref), 1e-30)
    print(f"{name:36s} norm got {gr.norm().item():.4e} ref {norm:.4e}  probe rel {rel:.3e}")
