"""Pin the CPU oracle (oracle/critic_vae_oracle.py) against fixtures produced by the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import torch

import critic_vae_oracle as O
import synth



def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_critic_matches_reference(golden_dir, critic_state):
    g = _load(golden_dir, "critic.npz")
    pred = O.critic_forward(critic_state, synth.make_frames(16, seed=30))
    np.testing.assert_allclose(pred.numpy(), g["pred"], rtol=0, atol=2e-6)


def test_train_steps_match_reference(golden_dir, critic_state):
    g = _load(golden_dir, "train_step.npz")
    B, steps = int(g["B"]), int(g["steps"])
    enc, dec = synth.make_vae_state(0)
    keys = [f"encoder.{k}" for k in O.PARAM_KEYS_ENC] + [f"decoder.{k}" for k in O.PARAM_KEYS_DEC]
    m = {k: None for k in keys}
    v = {k: None for k in keys}
    for s in range(steps):
        x, eps = synth.make_frames(B, seed=10 + s), synth.make_eps(B, seed=20 + s)
        pred = O.critic_forward(critic_state, x)
        np.testing.assert_allclose(pred.numpy(), g[f"s{s}_pred"], atol=2e-6)
        losses, recon, mu, logvar, grads = O.loss_and_grads(enc, dec, x, pred, eps)
        np.testing.assert_allclose(mu.numpy(), g[f"s{s}_mu"], atol=2e-5)
        np.testing.assert_allclose(logvar.numpy(), g[f"s{s}_logvar"], atol=2e-5)
        np.testing.assert_allclose(recon.numpy(), g[f"s{s}_recon"], atol=2e-6)
        got = [losses["total_loss"].item(), losses["recon_loss"].item(), losses["KLD"].item()]
        np.testing.assert_allclose(got, g[f"s{s}_losses"], rtol=2e-5)
        for k in keys:
            gr = grads[k].flatten().double()
            norm = g[f"s{s}_grad_norm/{k}"][0]
            # conv biases in front of BatchNorm have an exactly-zero true gradient: pure rounding noise
            if norm > 1e-7:
                np.testing.assert_allclose(gr.norm().item(), norm, rtol=2e-3)
                idx = synth.sample_indices(gr.numel())
                np.testing.assert_allclose(gr[idx].numpy(), g[f"s{s}_grad_smp/{k}"], atol=2e-3 * norm / np.sqrt(gr.numel()) + 1e-9, rtol=2e-2)
            sd, kk = (enc, k[len("encoder."):]) if k.startswith("encoder.") else (dec, k[len("decoder."):])
            if m[k] is None:
                m[k], v[k] = torch.zeros_like(sd[kk]), torch.zeros_like(sd[kk])
            O.adam_step(sd[kk], grads[k], m[k], v[k], s + 1)
        for pref, sd in (("encoder", enc), ("decoder", dec)):
            for kk, t in sd.items():
                f = t.flatten().double()
                if "conv-bias-before-bn" and kk in ("model.0.bias", "model.4.bias", "model.8.bias", "model.12.bias") and pref == "encoder":
                    continue  # Adam amplifies the rounding-noise gradient of these biases (|update| <= lr)
                idx = synth.sample_indices(f.numel())
                np.testing.assert_allclose(f[idx].numpy(), g[f"s{s}_post_smp/{pref}.{kk}"], atol=3e-5, rtol=1e-5,
                                           err_msg=f"step {s} {pref}.{kk}")


def test_nan_semantics(golden_dir, critic_state):
    g = _load(golden_dir, "train_step.npz")
    enc, dec = synth.make_vae_state(1)
    x, eps = synth.make_frames(int(g["B"]), seed=10), synth.make_eps(int(g["B"]), seed=20)
    _, mu, logvar, recon = O.vae_forward(enc, dec, x, O.critic_forward(critic_state, x), eps)
    losses = O.vae_loss(x, mu, logvar, recon)
    assert np.isnan(g["nan_losses"][0]) and torch.isnan(losses["total_loss"])
    np.testing.assert_allclose(losses["KLD"].item(), g["nan_losses"][2], rtol=2e-5)


def test_eval_and_inject_match_reference(golden_dir, critic_state):
    g = _load(golden_dir, "eval.npz")
    enc, dec = synth.make_vae_state(0)
    x = torch.from_numpy(g["x"])
    pred = O.critic_forward(critic_state, x)
    np.testing.assert_allclose(pred.numpy(), g["pred"], atol=2e-6)
    for i in range(x.shape[0]):
        with torch.no_grad():
            r1 = O.vae_evaluate(enc, dec, x[i:i + 1], torch.zeros(1) + pred[i])
            r0 = O.vae_evaluate(enc, dec, x[i:i + 1], torch.zeros(1))
            inj = torch.stack(O.vae_inject(enc, dec, x[i:i + 1]))
        np.testing.assert_allclose(r1.numpy()[0], g["recon_pred"][i], atol=2e-6)
        np.testing.assert_allclose(r0.numpy()[0], g["recon_zero"][i], atol=2e-6)
        np.testing.assert_allclose(inj.numpy(), g["inject"][i], atol=2e-6)


def test_mask_pipeline_bit_exact(golden_dir):
    g = _load(golden_dir, "mask_pipeline.npz")
    # diff map: np.dot order is BLAS-defined, so compare with 1-ulp slack; everything after is exact
    for i in range(g["diff"].shape[0]):
        d, m = O.diff_grey(g["recon_one"][i], g["recon_zero"][i])
        assert np.array_equal(d, g["diff"][i]) and m == g["max_values"][i]
        d2, _ = O.diff_grey_ordered(g["recon_one"][i], g["recon_zero"][i])
        np.testing.assert_allclose(d2, g["diff"][i], rtol=4e-16, atol=1e-18)
    dm, tm = O.diff_and_thr_masks(list(g["diff"]), list(g["max_values"]), thr=50)
    assert np.array_equal(dm, g["diff_u8"]) and np.array_equal(tm, g["thr_mask_50"])
    sweep = [O.iou(g["gt"], O.diff_and_thr_masks(list(g["diff"]), list(g["max_values"]), thr=t)[1]) for t in range(0, 130, 10)]
    assert sweep == list(g["iou_sweep"])
    m2 = [np.amax(d) for d in g["diff2"]]
    dm2, tm2 = O.diff_and_thr_masks(list(g["diff2"]), m2, thr=50)
    assert np.array_equal(dm2, g["diff2_u8"]) and np.array_equal(tm2, g["thr2_mask_50"])
    assert O.iou(g["gt"], tm2) == float(g["iou2"])
    assert O.iou(np.zeros((2, 4, 4), bool), np.zeros((2, 4, 4), bool)) == 1     # vae_utility.py:61-62


def test_msssim_value_and_grad(golden_dir):
    g = _load(golden_dir, "msssim.npz")
    for tag, B in (("a", 2), ("b", 3)):
        x = synth.make_frames(B, seed=51 + B)
        r = torch.from_numpy(g[f"{tag}_recon"]).requires_grad_(True)
        loss = O.msssim_loss(r, x)
        loss.backward()
        np.testing.assert_allclose(loss.item(), float(g[f"{tag}_loss"]), rtol=1e-5)
        np.testing.assert_allclose(r.grad.numpy(), g[f"{tag}_grad"], atol=1e-7, rtol=1e-3)
