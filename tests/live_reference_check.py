"""Run in a CLEAN interpreter by tests/test_oracle_live_reference.py (build container only: needs /root/reference).

Imports the UNMODIFIED reference modules next to oracle/critic_vae_oracle.py and compares them on inputs that are
NOT in the committed fixtures (other seeds, other batch sizes): training-step loss / gradients, evaluate / inject,
the critic, and the numpy mask pipeline.  A subprocess is needed because the drop-in package uses the reference's
module names (vae_nets, critic_net, ...)."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, REF)
import synth  # noqa: E402
import critic_vae_oracle as O  # noqa: E402

sys.modules.setdefault("minerl", types.ModuleType("minerl"))
sys.modules.setdefault("denseCRF", types.ModuleType("denseCRF"))
from PIL import ImageFont  # noqa: E402

_tt, _default_font = ImageFont.truetype, ImageFont.load_default()
ImageFont.truetype = lambda *a, **k: _default_font
os.chdir(REF)
import vae_nets  # noqa: E402
import vae_utility  # noqa: E402
import vae_parameters  # noqa: E402
from critic_net import Critic  # noqa: E402

ImageFont.truetype = _tt
assert os.path.realpath(vae_nets.__file__).startswith(os.path.realpath(REF)), vae_nets.__file__
torch.set_num_threads(4)

crit_sd = torch.load(os.path.join(REF, vae_parameters.CRITIC_PATH), map_location="cpu")
critic = Critic()
critic.load_state_dict(crit_sd)
critic.eval()


def ref_vae(seed):
    vae = vae_nets.VariationalAutoencoder()
    enc, dec = synth.make_vae_state(seed)
    vae.encoder.load_state_dict(enc)
    vae.decoder.load_state_dict(dec)
    return vae, enc, dec


# ---- training step: loss and every gradient (vae.py:47-57) ---------------------------------------------------------
for seed, B in ((0, 5), (3, 2)):
    vae, enc, dec = ref_vae(seed)
    vae.train()
    x, eps = synth.make_frames(B, seed=300 + seed), synth.make_eps(B, seed=400 + seed)
    preds = critic.evaluate(x)
    assert torch.allclose(preds, O.critic_forward(crit_sd, x), atol=1e-6)
    orig = torch.randn_like
    torch.randn_like = lambda t, *a, **k: eps.to(t.dtype)
    try:
        o = vae(x, preds)
    finally:
        torch.randn_like = orig
    losses = vae.vae_loss(o[0], o[1], o[2], o[3])
    l_or, recon_or, mu_or, lv_or, g_or = O.loss_and_grads(enc, dec, x, preds, eps, update_stats=False)
    assert torch.allclose(o[1], mu_or, atol=1e-5) and torch.allclose(o[2], lv_or, atol=1e-5)
    assert torch.allclose(o[3], recon_or, atol=1e-5)
    for k in ("total_loss", "recon_loss", "KLD"):
        a, b = losses[k].item(), l_or[k].item()
        assert (np.isnan(a) and np.isnan(b)) or abs(a - b) <= 1e-5 * max(1.0, abs(b)), (k, a, b)
    if not np.isnan(losses["total_loss"].item()):
        losses["total_loss"].backward()
        for pref, mod in (("encoder", vae.encoder), ("decoder", vae.decoder)):
            for k, p in mod.named_parameters():
                ref, got = p.grad.double(), g_or[f"{pref}.{k}"].double()
                err = (ref - got).norm().item() / max(ref.norm().item(), 1e-12)
                assert err < 1e-3 or ref.norm().item() < 1e-7, (pref, k, err)

# ---- evaluate / inject (vae_nets.py:42-46, :31-40) -----------------------------------------------------------------
vae, enc, dec = ref_vae(0)
vae.eval()
x = synth.make_frames(3, seed=500)
for i in range(3):
    p = critic.evaluate(x[i:i + 1])
    with torch.no_grad():
        r1 = vae.evaluate(x[i:i + 1], torch.zeros(1) + p[0])
        inj = torch.cat(vae.inject(x[i:i + 1]))
    assert torch.allclose(r1, O.vae_evaluate(enc, dec, x[i:i + 1], torch.zeros(1) + p[0]), atol=1e-5)
    assert torch.allclose(inj, torch.cat(O.vae_inject(enc, dec, x[i:i + 1])), atol=1e-5)

# ---- numpy mask pipeline (vae_utility.py:148-160, 279-284, 56-68, 106-110) ------------------------------------------
rng = np.random.Generator(np.random.PCG64(600))
d = rng.uniform(0, 1, (9, 64, 64)) ** 2
d[2] = 0.0
mx = [np.amax(v) for v in d]
gt = synth.make_gt_masks(9, seed=601)
for thr in (0, 50, 120):
    dm, tm = vae_utility.get_diff_and_thr_masks([v.copy() for v in d], list(mx), thr=thr)
    dm_o, tm_o = O.diff_and_thr_masks([v.copy() for v in d], list(mx), thr=thr)
    assert np.array_equal(dm, dm_o) and np.array_equal(tm, tm_o)
    assert vae_utility.get_iou(gt, tm) == O.iou(gt, tm_o)
assert vae_utility.get_diff_factor(list(mx)) == O.diff_factor(list(mx))
print("LIVE REFERENCE CHECK OK")
