"""The drop-in modules (vae_nets.VariationalAutoencoder, critic_net.Critic) driven exactly like the
reference's train() / evaluate() / inject() and compared with fixtures produced by the unmodified
reference (tests/golden/).  Tolerances are BASELINE.json's: loss rel 1e-3, pixels max-abs 1e-2
(the network runs bf16 operands with fp32 accumulation; the reference is fp32)."""
import os

import numpy as np
import pytest
import torch

import critic_vae_oracle as O
import synth

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-3      # north_star: rel 1e-3 on loss
PIXEL_ATOL = 1e-2     # north_star: 1e-2 max-abs on pixels
LATENT_ATOL = 2e-2    # mu / logvar (bf16 network, values O(1))
GRAD_BUDGET_FACTOR = 2.0   # allowed multiple of the bf16-storage error budget; the budget is ONE realisation of the rounding noise (measured ratios: 0.3 .. 1.55)
GRAD_REL_FLOOR = 5e-3      # ... or this relative L2 error outright, whichever is larger


def _modules(critic_state, seed=0):
    import vae_nets
    from critic_net import Critic
    vae = vae_nets.VariationalAutoencoder().to("cuda")
    enc, dec = synth.make_vae_state(seed)
    vae.encoder.load_state_dict(enc)
    vae.decoder.load_state_dict(dec)
    critic = Critic()
    critic.load_state_dict(critic_state)
    critic.eval().to("cuda")
    return vae, critic


def _rel(got, ref):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    return np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-30)


def test_state_dict_roundtrip_and_flat_binding(critic_state):
    vae, _ = _modules(critic_state)
    enc, dec = synth.make_vae_state(0)
    x = synth.make_frames(2, seed=10).cuda()
    vae.eval()
    vae.evaluate(x, torch.zeros(2, device="cuda"))              # binds parameters to the flat buffer
    for k, v in enc.items():
        assert torch.equal(vae.encoder.state_dict()[k].cpu(), v), k
    for k, v in dec.items():
        assert torch.equal(vae.decoder.state_dict()[k].cpu(), v), k
    eng = vae._engine
    assert eng.n_params == 2583971
    assert all(p.data_ptr() == eng.view(n).data_ptr() for n, p in vae._named_flat())
    # in-place edits and load_state_dict stay visible to the kernels
    r0 = vae.evaluate(x, torch.zeros(2, device="cuda"))
    with torch.no_grad():
        vae.decoder.model[12].bias.add_(0.25)
    r1 = vae.evaluate(x, torch.zeros(2, device="cuda"))
    assert (r1 - r0).abs().max().item() > 1e-2
    vae.decoder.load_state_dict(dec)
    assert torch.equal(vae.evaluate(x, torch.zeros(2, device="cuda")), r0)


def test_train_steps_match_reference(golden_dir, critic_state):
    """vae.py:47-58 for two optimizer steps: critic -> forward -> vae_loss -> backward -> Adam."""
    g = np.load(os.path.join(golden_dir, "train_step.npz"))
    B, steps = int(g["B"]), int(g["steps"])
    vae, critic = _modules(critic_state)
    vae.train()
    opt = torch.optim.Adam(vae.parameters(), lr=5e-5)
    report = []
    for s in range(steps):
        x, eps = synth.make_frames(B, seed=10 + s).cuda(), synth.make_eps(B, seed=20 + s).cuda()
        preds = critic.evaluate(x)
        np.testing.assert_allclose(preds.cpu().numpy(), g[f"s{s}_pred"], atol=5e-6)
        # CPU copy of the weights this step starts from (the oracle differentiates the same point)
        enc_ref = {k_: v_.detach().cpu().clone() for k_, v_ in vae.encoder.state_dict().items()}
        dec_ref = {k_: v_.detach().cpu().clone() for k_, v_ in vae.decoder.state_dict().items()}
        opt.zero_grad()
        out = vae(x, preds, eps=eps)
        losses = vae.vae_loss(out[0], out[1], out[2], out[3])
        losses["total_loss"].backward()
        vae._engine.check_fault()
        # Reference point: the CPU oracle evaluated at the weights THIS step started from.  At step 0
        # these are the golden run's weights, so the golden fixtures apply at full tolerance; from step 1
        # on the two trajectories have taken one Adam step each on slightly different gradients (every
        # weight moves by +-lr on the first step), so the golden values are only a loose sanity check.
        x_c, eps_c = x.cpu(), eps.cpu()
        l_ref, recon_ref, mu_ref, lv_ref, g_ref = O.loss_and_grads(enc_ref, dec_ref, x_c, preds.cpu(), eps_c, update_stats=False)
        mu_g, lv_g, rec_g = (out[i].detach().cpu().numpy() for i in (1, 2, 3))
        np.testing.assert_allclose(mu_g, mu_ref.numpy(), atol=LATENT_ATOL)
        np.testing.assert_allclose(lv_g, lv_ref.numpy(), atol=LATENT_ATOL)
        np.testing.assert_allclose(rec_g, recon_ref.numpy(), atol=PIXEL_ATOL)
        got = [losses["total_loss"].item(), losses["recon_loss"].item(), losses["KLD"].item()]
        np.testing.assert_allclose(got[0], l_ref["total_loss"].item(), rtol=LOSS_RTOL)
        np.testing.assert_allclose(got[1], l_ref["recon_loss"].item(), rtol=LOSS_RTOL)
        np.testing.assert_allclose(got[2], l_ref["KLD"].item(), rtol=2e-2)
        slack = 1 if s == 0 else 5
        np.testing.assert_allclose(mu_g, g[f"s{s}_mu"], atol=slack * LATENT_ATOL)
        np.testing.assert_allclose(lv_g, g[f"s{s}_logvar"], atol=slack * LATENT_ATOL)
        np.testing.assert_allclose(rec_g, g[f"s{s}_recon"], atol=slack * PIXEL_ATOL)
        np.testing.assert_allclose(got[0], g[f"s{s}_losses"][0], rtol=slack * LOSS_RTOL)
        np.testing.assert_allclose(got[1], g[f"s{s}_losses"][1], rtol=slack * LOSS_RTOL)
        report.append(f"step {s}: loss {got[0]:.6f} vs {g[f's{s}_losses'][0]:.6f}  "
                      f"recon maxerr {np.abs(out[3].detach().cpu().numpy() - g[f's{s}_recon']).max():.2e}")
        # Gradients.  ReLU / max-pool gates flip under the 2^-9 relative rounding of bf16 activation
        # storage, so ANY bf16-operand implementation deviates from the fp32 reference by an amount that
        # grows towards the first layers.  The oracle's precision model measures that budget on this very
        # batch; the kernels must stay within GRAD_BUDGET_FACTOR of it for every parameter tensor, and
        # within GRAD_REL outright where the budget is small (decoder tail).
        _, _, _, _, g_mod = O.loss_and_grads_bf16_storage(enc_ref, dec_ref, x_c, preds.cpu(), eps_c)
        worst = ("", 0.0, 0.0)
        for name, prm in vae.named_parameters():
            norm = g[f"s{s}_grad_norm/{name}"][0]
            gr = prm.grad.detach().double().cpu()
            if norm < 1e-7:          # conv biases feeding BatchNorm: the true gradient is exactly zero
                assert gr.abs().max().item() == 0.0, name
                continue
            ref = g_ref[name].double()
            np.testing.assert_allclose(ref.norm().item(), norm, rtol=2e-3 if s == 0 else 5e-2, err_msg=f"oracle vs golden {name}")
            rel = _rel(gr.numpy(), ref.numpy())
            budget = _rel(g_mod[name].double().numpy(), ref.numpy())
            if rel / max(budget, 1e-9) > worst[1] / max(worst[2], 1e-9):
                worst = (name, rel, budget)
            assert rel < max(GRAD_BUDGET_FACTOR * budget, GRAD_REL_FLOOR), \
                f"step {s} grad {name}: rel err {rel:.3e}, bf16-storage budget {budget:.3e}"
            idx = synth.sample_indices(gr.numel())
            probe = _rel(gr.flatten()[idx].numpy(), g[f"s{s}_grad_smp/{name}"])
            assert probe < max(2 * GRAD_BUDGET_FACTOR * budget, 4 * GRAD_REL_FLOOR), f"step {s} grad probe {name}: {probe:.3e}"
        report.append(f"step {s}: worst grad err/budget {worst[1]:.3e}/{worst[2]:.3e} ({worst[0]})")
        opt.step()
        for pref, mod in (("encoder", vae.encoder), ("decoder", vae.decoder)):
            for kk, t in mod.state_dict().items():
                if pref == "encoder" and kk in ("model.0.bias", "model.4.bias", "model.8.bias", "model.12.bias"):
                    continue      # Adam turns the rounding-noise gradient of these biases into +-lr steps
                f = t.flatten().double().cpu()
                idx = synth.sample_indices(f.numel())
                ref = g[f"s{s}_post_smp/{pref}.{kk}"]
                # one Adam step moves a weight by at most ~lr; agreement to a fraction of that step
                # (running BN buffers: 0.1 x the batch statistic of bf16-stored activations)
                atol = 2e-3 if "running" in kk else (1.5e-4 if s == 0 else 3e-4)
                np.testing.assert_allclose(f[idx].numpy(), ref, atol=atol, rtol=1e-3, err_msg=f"step {s} {pref}.{kk}")
    print("\n".join(report))


def test_nan_loss_like_reference(golden_dir, critic_state):
    g = np.load(os.path.join(golden_dir, "train_step.npz"))
    vae, critic = _modules(critic_state, seed=1)
    vae.train()
    B = int(g["B"])
    x, eps = synth.make_frames(B, seed=10).cuda(), synth.make_eps(B, seed=20).cuda()
    out = vae(x, critic.evaluate(x), eps=eps)
    losses = vae.vae_loss(*out)
    assert np.isnan(g["nan_losses"][0]) and torch.isnan(losses["total_loss"]) and torch.isnan(losses["recon_loss"])
    np.testing.assert_allclose(losses["KLD"].item(), g["nan_losses"][2], rtol=2e-2)


def test_evaluate_and_inject_match_reference(golden_dir, critic_state):
    """vae_nets.py:42-46 / :31-40 on real JPEG frames, batch of one like the reference and batched."""
    g = np.load(os.path.join(golden_dir, "eval.npz"))
    vae, critic = _modules(critic_state)
    vae.eval()
    x = torch.from_numpy(g["x"]).cuda()
    pred = critic.evaluate(x)
    np.testing.assert_allclose(pred.cpu().numpy(), g["pred"], atol=5e-6)
    for i in range(x.shape[0]):
        r1 = vae.evaluate(x[i:i + 1], torch.zeros(1, device="cuda") + pred[i])
        r0 = vae.evaluate(x[i:i + 1], torch.zeros(1, device="cuda"))
        assert r1.shape == (1, 3, 64, 64)
        np.testing.assert_allclose(r1.cpu().numpy()[0], g["recon_pred"][i], atol=PIXEL_ATOL)
        np.testing.assert_allclose(r0.cpu().numpy()[0], g["recon_zero"][i], atol=PIXEL_ATOL)
        inj = vae.inject(x[i:i + 1])
        assert len(inj) == 6 and inj[0].shape == (1, 3, 64, 64)
        np.testing.assert_allclose(torch.cat(inj).cpu().numpy(), g["inject"][i][:, 0], atol=PIXEL_ATOL)
    rb = vae.evaluate(x, pred)                                   # batched extension == per-frame calls
    np.testing.assert_allclose(rb.cpu().numpy(), g["recon_pred"], atol=PIXEL_ATOL)
    vae._engine.check_fault()


def test_gradient_accumulation_without_zero_grad(critic_state):
    vae, critic = _modules(critic_state)
    vae.train()
    x, eps = synth.make_frames(4, seed=10).cuda(), synth.make_eps(4, seed=20).cuda()
    preds = critic.evaluate(x)
    rm = vae.encoder.model[1].running_mean.clone()
    vae.vae_loss(*vae(x, preds, eps=eps))["total_loss"].backward()
    g1 = vae.decoder.model[0].weight.grad.clone()
    vae.encoder.model[1].running_mean.copy_(rm)
    vae.vae_loss(*vae(x, preds, eps=eps))["total_loss"].backward()      # no zero_grad: must add
    g2 = vae.decoder.model[0].weight.grad
    assert _rel(g2.cpu().numpy(), 2 * g1.cpu().numpy()) < 1e-2
