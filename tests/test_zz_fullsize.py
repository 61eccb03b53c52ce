"""Full-size checks of the CUDA-graph training step (cvae_native.trainer.TrainStep, the path bench.py times) at
BASELINE.json's batch of 256, through properties that do not need a 256-frame CPU reference run:

  * tiling invariance: a batch made of 4 copies of 64 frames (and of their eps) has the same BatchNorm statistics,
    the same batch-global MS-SSIM means, hence the same loss and the same (batch-mean) gradients as the 64-frame batch;
  * the 64-frame step itself is checked against the CPU oracle (loss rel 1e-3 as in BASELINE.json's north_star).

The file name keeps these tests last in the run."""
import numpy as np
import pytest
import torch

import critic_vae_oracle as O
import synth
from test_vae_module import _modules, _rel, LOSS_RTOL

pytestmark = pytest.mark.gpu


def _graph_step(critic_state, batch, x64, eps64):
    """One optimizer step of a fresh, identically initialised model through TrainStep (CUDA graph, side streams).
    Returns (losses [total, recon, KLD], flat gradient, engine) with everything copied to the host."""
    from cvae_native.trainer import TrainStep
    vae, critic = _modules(critic_state)
    vae.train()
    st = TrainStep(vae, critic, batch)
    reps = batch // 64
    st.load(frames=x64.repeat(reps, 1, 1, 1).cuda(), eps=eps64.repeat(reps, 1).cuda())
    losses = st.run().clone()
    torch.cuda.synchronize()
    st.eng.check_fault()
    return losses.cpu().numpy().astype(np.float64), st.eng.gflat.clone().cpu(), st


def test_batch_256_equals_four_copies_of_batch_64(critic_state):
    x64, eps64 = synth.make_frames(64, seed=70), synth.make_eps(64, seed=71)
    l64, g64, _ = _graph_step(critic_state, 64, x64, eps64)
    l256, g256, st = _graph_step(critic_state, 256, x64, eps64)
    assert np.all(np.isfinite(l256)) and torch.isfinite(g256).all()
    np.testing.assert_allclose(l256, l64, rtol=5e-4, err_msg="loss of the tiled 256 batch vs the 64 batch")
    # same gradient up to fp32 summation order (split-K partitions differ with the batch size)
    assert _rel(g256.double().numpy(), g64.double().numpy()) < 1e-2
    for name in ("decoder.model.12.weight", "decoder.model.0.weight", "encoder.model.8.weight", "encoder.model.0.weight",
                 "encoder.fc_var.weight", "decoder.decoder_input.weight"):
        a = st.eng.view(name, g256).double().numpy()
        b = st.eng.view(name, g64).double().numpy()
        assert _rel(a, b) < 2e-2, name


def test_graph_step_matches_oracle_at_64(critic_state):
    x64, eps64 = synth.make_frames(64, seed=72), synth.make_eps(64, seed=73)
    losses, g, st = _graph_step(critic_state, 64, x64, eps64)
    enc, dec = synth.make_vae_state(0)
    pred = O.critic_forward(critic_state, x64)
    l_ref, _, _, _, g_ref = O.loss_and_grads(enc, dec, x64, pred, eps64, update_stats=False)
    np.testing.assert_allclose(losses[0], l_ref["total_loss"].item(), rtol=LOSS_RTOL)
    np.testing.assert_allclose(losses[1], l_ref["recon_loss"].item(), rtol=LOSS_RTOL)
    np.testing.assert_allclose(losses[2], l_ref["KLD"].item(), rtol=2e-2)
    got = st.eng.view("decoder.model.12.weight", g).double().numpy()
    assert _rel(got, g_ref["decoder.model.12.weight"].double().numpy()) < 2e-2
