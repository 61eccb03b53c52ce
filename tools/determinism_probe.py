"""Diagnostic: repeatability and timing of forward / loss / backward through the drop-in module."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tests", "oracle", "critic-vae_b200"):
    sys.path.insert(0, os.path.join(ROOT, p))
import synth, vae_nets

vae = vae_nets.VariationalAutoencoder().to("cuda")
enc, dec = synth.make_vae_state(0)
vae.encoder.load_state_dict(enc); vae.decoder.load_state_dict(dec)
for B in (4, 64):
    x, eps = synth.make_frames(B, seed=10).cuda(), synth.make_eps(B, seed=20).cuda()
    pred = torch.full((B,), 0.5, device="cuda")
    for mode in ("eval", "train"):
        vae.train(mode == "train")
        outs, grads, times = [], [], []
        for it in range(8):
            for bi in (1, 5, 9, 13):
                bn = vae.encoder.model[bi]
                bn.running_mean.copy_(enc[f"model.{bi}.running_mean"]); bn.running_var.copy_(enc[f"model.{bi}.running_var"])
            torch.cuda.synchronize(); t = time.time()
            if mode == "eval":
                r = vae.evaluate(x, pred)
                ml = vae._engine.workspace(B, True).ml.clone()
            else:
                vae.zero_grad()
                o = vae(x, pred, eps=eps)
                l = vae.vae_loss(*o)["total_loss"]; l.backward()
                r, ml = o[3].detach(), torch.cat((o[1], o[2]), 1).detach()
                grads.append(vae._engine.gflat.clone())
            torch.cuda.synchronize(); times.append(time.time() - t)
            vae._engine.check_fault()
            outs.append((ml.clone(), r.clone()))
        dml = max((o[0] - outs[0][0]).abs().max().item() for o in outs)
        dr = max((o[1] - outs[0][1]).abs().max().item() for o in outs)
        dg = max(((g - grads[0]).norm() / grads[0].norm()).item() for g in grads) if grads else 0.0
        print(f"B={B} {mode}: max run-to-run diff ml {dml:.3e} recon {dr:.3e} grad rel {dg:.3e}; "
              f"time first {times[0]*1e3:.1f} ms, median {np.median(times)*1e3:.2f} ms", flush=True)
