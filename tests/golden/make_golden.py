"""Generate the committed golden fixtures by running the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

The reference classes (vae_nets.VariationalAutoencoder, critic_net.Critic) and utilities
(vae_utility.get_diff_image / get_diff_and_thr_masks / get_iou / preprocess_observation) are
imported from /root/reference as they are; `minerl` and `denseCRF` are stubbed in sys.modules and the
hard-coded Ubuntu font falls back to PIL's default.  Weights and inputs come from
tests/synth.py so the GPU box can rebuild them bit-identically; torch.randn_like is patched for the
duration of `reparametrize` so both sides see the same eps (vae_nets.py:50).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, REF)

import synth  # noqa: E402

sys.modules.setdefault("minerl", types.ModuleType("minerl"))
sys.modules.setdefault("denseCRF", types.ModuleType("denseCRF"))
from PIL import ImageFont  # noqa: E402

_tt = ImageFont.truetype
_default_font = ImageFont.load_default()
ImageFont.truetype = lambda *a, **k: _default_font
os.chdir(REF)
import vae_nets  # noqa: E402
import vae_utility  # noqa: E402
import vae_parameters  # noqa: E402
from critic_net import Critic  # noqa: E402

ImageFont.truetype = _tt
torch.set_num_threads(8)


def ref_vae(seed=0):
    vae = vae_nets.VariationalAutoencoder()
    enc, dec = synth.make_vae_state(seed)
    vae.encoder.load_state_dict(enc)
    vae.decoder.load_state_dict(dec)
    return vae


def ref_critic():
    c = Critic()
    c.load_state_dict(torch.load(os.path.join(REF, vae_parameters.CRITIC_PATH), map_location="cpu"))
    c.eval()
    return c


class patched_eps:
    def __init__(self, eps):
        self.eps = eps

    def __enter__(self):
        self.orig = torch.randn_like
        torch.randn_like = lambda t, *a, **k: self.eps.to(t.dtype)

    def __exit__(self, *a):
        torch.randn_like = self.orig


def summarize(t: torch.Tensor):
    f = t.detach().flatten().double()
    idx = synth.sample_indices(f.numel())
    return np.array([f.norm().item(), f.sum().item()]), f[idx].float().numpy()


def golden_train(B=4, steps=2):
    """vae.py:47-58, `steps` consecutive optimizer steps on fixed batches."""
    vae, critic = ref_vae(0), ref_critic()
    vae.train()
    opt = torch.optim.Adam(vae.parameters(), lr=vae_parameters.lr)
    out = {}
    for s in range(steps):
        x = synth.make_frames(B, seed=10 + s)
        eps = synth.make_eps(B, seed=20 + s)
        preds = critic.evaluate(x)
        opt.zero_grad()
        with patched_eps(eps):
            o = vae(x, preds)
        losses = vae.vae_loss(o[0], o[1], o[2], o[3])
        losses["total_loss"].backward()
        out[f"s{s}_pred"] = preds.numpy()
        out[f"s{s}_mu"] = o[1].detach().numpy()
        out[f"s{s}_logvar"] = o[2].detach().numpy()
        out[f"s{s}_recon"] = o[3].detach().numpy()
        out[f"s{s}_losses"] = np.array([losses["total_loss"].item(), losses["recon_loss"].item(), losses["KLD"].item()])
        for pref, mod in (("encoder", vae.encoder), ("decoder", vae.decoder)):
            for k, p in mod.named_parameters():
                n, smp = summarize(p.grad)
                out[f"s{s}_grad_norm/{pref}.{k}"] = n
                out[f"s{s}_grad_smp/{pref}.{k}"] = smp
        opt.step()
        for pref, mod in (("encoder", vae.encoder), ("decoder", vae.decoder)):
            for k, p in mod.state_dict().items():
                n, smp = summarize(p.float())
                out[f"s{s}_post_norm/{pref}.{k}"] = n
                out[f"s{s}_post_smp/{pref}.{k}"] = smp
    # NaN semantics: with weight seed 1 a coarse-level cs mean is negative and `negative ** w`
    # (vae_nets.py:243) turns the reference loss into NaN; the drop-in must do the same.
    vae = ref_vae(1)
    vae.train()
    x, eps = synth.make_frames(B, seed=10), synth.make_eps(B, seed=20)
    with patched_eps(eps):
        o = vae(x, critic.evaluate(x))
    nl = vae.vae_loss(o[0], o[1], o[2], o[3])
    out["nan_losses"] = np.array([nl["total_loss"].item(), nl["recon_loss"].item(), nl["KLD"].item()])
    np.savez_compressed(os.path.join(HERE, "train_step.npz"), B=B, steps=steps, **out)
    print("train_step: losses", [out[f"s{s}_losses"] for s in range(steps)], "nan case", out["nan_losses"])


def golden_eval(N=4):
    """vae_nets.py:42-46 (evaluate), :31-40 (inject), critic_net.py:66-69 on real JPEG frames."""
    from PIL import Image
    vae, critic = ref_vae(0), ref_critic()
    vae.eval()
    files = sorted(os.listdir(os.path.join(REF, "source-images")))[:N]
    out = {"files": np.array(files)}
    xs, preds, r1, r0, inj = [], [], [], [], []
    for f in files:
        img = np.array(Image.open(os.path.join(REF, "source-images", f)))
        x = vae_utility.preprocess_observation(img)          # vae_utility.py:337-343
        p = critic.evaluate(x)
        with torch.no_grad():
            r1.append(vae.evaluate(x, torch.zeros(1) + p[0]).numpy())
            r0.append(vae.evaluate(x, torch.zeros(1)).numpy())
            inj.append(np.stack([t.numpy() for t in vae.inject(x)]))
        xs.append(x.numpy())
        preds.append(p.numpy())
    out.update(x=np.concatenate(xs), pred=np.concatenate(preds), recon_pred=np.concatenate(r1),
               recon_zero=np.concatenate(r0), inject=np.stack(inj))
    np.savez_compressed(os.path.join(HERE, "eval.npz"), **out)
    print("eval: preds", out["pred"].ravel())


def golden_critic(N=16):
    critic = ref_critic()
    x = synth.make_frames(N, seed=30)
    np.savez_compressed(os.path.join(HERE, "critic.npz"), pred=critic.evaluate(x).numpy())


def golden_mask(N=12):
    """vae_utility.py:256-277 (get_diff_image), :148-160, :56-68 on N frames, thr sweep of vae.py:121."""
    vae, critic = ref_vae(0), ref_critic()
    vae.eval()
    frames = synth.make_frames(N, seed=40)
    gt = synth.make_gt_masks(N, seed=41)
    ones, zeros, diffs, maxes, preds = [], [], [], [], []
    for i in range(N):
        x = frames[i:i + 1]
        p = critic.evaluate(x)
        ro, rz, d, m = vae_utility.get_diff_image(vae, x, p[0])
        ones.append(ro); zeros.append(rz); diffs.append(d); maxes.append(m); preds.append(p.numpy())
    out = dict(recon_one=np.stack(ones), recon_zero=np.stack(zeros), diff=np.stack(diffs),
               max_values=np.array(maxes), pred=np.concatenate(preds), gt=gt)
    ious = []
    for thr in range(0, 130, 10):
        dm, tm = vae_utility.get_diff_and_thr_masks([d.copy() for d in diffs], list(maxes), thr=thr)
        ious.append(vae_utility.get_iou(gt, tm))
        if thr == 50:
            out["diff_u8"] = dm
            out["thr_mask_50"] = tm
    out["iou_sweep"] = np.array(ious)
    # a second, harsher diff set (values straddling mean_max, zeros, exact ties) straight into the
    # numpy stage, independent of any network
    rng = np.random.Generator(np.random.PCG64(42))
    d2 = rng.uniform(0, 1, (N, 64, 64)) ** 3
    d2[0] = 0.0
    d2[1, :8] = d2[1].max()
    m2 = [np.amax(d) for d in d2]
    dm2, tm2 = vae_utility.get_diff_and_thr_masks([d.copy() for d in d2], m2, thr=50)
    out.update(diff2=d2, diff2_u8=dm2, thr2_mask_50=tm2, iou2=vae_utility.get_iou(gt, tm2))
    np.savez_compressed(os.path.join(HERE, "mask_pipeline.npz"), **out)
    print("mask: iou sweep", out["iou_sweep"], "iou2", out["iou2"])


def golden_msssim():
    """MSSIM.forward (vae_nets.py:217-247) value and gradient w.r.t. img1."""
    m = vae_nets.MSSIM()
    out = {}
    rng = np.random.Generator(np.random.PCG64(50))
    for tag, B in (("a", 2), ("b", 3)):
        x = synth.make_frames(B, seed=51 + B)
        r = torch.from_numpy(np.clip(x.numpy() + rng.normal(0, 0.15, x.shape), -1, 1).astype(np.float32))
        r.requires_grad_(True)
        loss = m(r, x)
        loss.backward()
        out[f"{tag}_recon"] = r.detach().numpy()
        out[f"{tag}_loss"] = np.array(loss.item())
        out[f"{tag}_grad"] = r.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "msssim.npz"), **out)
    print("msssim:", out["a_loss"], out["b_loss"])


if __name__ == "__main__":
    golden_train()
    golden_eval()
    golden_critic()
    golden_mask()
    golden_msssim()
