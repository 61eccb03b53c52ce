"""Diagnostic: forward error of the golden training batch over repeated fresh runs."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tests", "oracle", "critic-vae_b200"):
    sys.path.insert(0, os.path.join(ROOT, p))
import synth, vae_nets
from critic_net import Critic
torch.set_num_threads(8)
g = np.load(os.path.join(ROOT, "tests/golden/train_step.npz"))
cs = torch.load(os.path.join(ROOT, "critic-vae_b200/saved-networks/critic-rewidx=1-cepochs=15-datamode=trunk-datasize=99999-shift=12-chfak=1-dropout=0.3.pt"))
critic = Critic(); critic.load_state_dict(cs); critic.eval().to("cuda")
for rep in range(12):
    t = time.time()
    vae = vae_nets.VariationalAutoencoder().to("cuda")
    enc, dec = synth.make_vae_state(0)
    vae.encoder.load_state_dict(enc); vae.decoder.load_state_dict(dec)
    vae.train()
    x, eps = synth.make_frames(4, seed=10).cuda(), synth.make_eps(4, seed=20).cuda()
    out = vae(x, critic.evaluate(x), eps=eps)
    l = vae.vae_loss(*out)
    torch.cuda.synchronize()
    vae._engine.check_fault()
    print(rep, f"mu err {np.abs(out[1].detach().cpu().numpy()-g['s0_mu']).max():.4f} lv err {np.abs(out[2].detach().cpu().numpy()-g['s0_logvar']).max():.4f}"
          f" recon err {np.abs(out[3].detach().cpu().numpy()-g['s0_recon']).max():.2e} loss {l['total_loss'].item():.6f} ref {g['s0_losses'][0]:.6f} t={time.time()-t:.2f}s", flush=True)
