"""Data-parallel training on real GPUs: two ranks, NCCL gradient all-reduce, one optimizer step.
Needs >= 2 GPUs (skipped otherwise): run with `gpurun --gpus 2 -- python -m pytest tests/test_dp_gpu.py -m gpu`.

Checked: the all-reduced gradient / world equals the mean of the per-shard ORACLE gradients (within the bf16-storage
budget), the post-step parameters equal torch.optim.Adam's rule applied to that mean gradient, and both ranks hold
bit-identical parameters afterwards (PyTorch-DDP semantics, vae.py:47-58 per shard)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = 32


def _worker(rank, world, port, out_dir, comm):
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "critic-vae_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    import synth
    from test_vae_module import _modules
    from cvae_native.trainer import TrainStep
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), NCCL_IB_DISABLE="1", NCCL_P2P_LEVEL="NVL", CVAE_COMM=comm)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    crit = torch.load(os.path.join(ROOT, "critic-vae_b200", "saved-networks",
                                   "critic-rewidx=1-cepochs=15-datamode=trunk-datasize=99999-shift=12-chfak=1-dropout=0.3.pt"), map_location="cpu")
    vae, critic = _modules(crit, seed=0)
    if rank == 1:      # a rank that starts from different weights must be overwritten by rank 0's (TrainStep broadcasts)
        with torch.no_grad():
            vae.decoder.model[12].bias.add_(1.0)
    vae.train()
    st = TrainStep(vae, critic, B, process_group=dist.group.WORLD)
    flat0 = st.eng.flat.clone()
    st.load(frames=synth.make_frames(B, seed=900 + rank).cuda(), eps=synth.make_eps(B, seed=910 + rank).cuda())
    losses = st.run().clone()
    torch.cuda.synchronize()
    st.eng.check_fault()
    torch.save({"flat0": flat0.cpu(), "flat": st.eng.flat.cpu(), "gsum": st.eng.gflat.cpu(), "losses": losses.cpu()},
               os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    if comm == "native":
        from cvae_native import binding as L
        assert L.lib.cvae_comm_world() == world, "CVAE_COMM=native must have built libcvae's own communicator"
        L.check(L.lib.cvae_comm_destroy())
    dist.destroy_process_group()


@pytest.mark.parametrize("comm", ["torch", "native"])      # torch.distributed's all-reduce / libcvae's cvae_comm_allreduce_sum
def test_two_rank_step_equals_oracle_on_the_shards(tmp_path, critic_state, comm):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    import critic_vae_oracle as O
    import synth
    from cvae_native.engine import param_layout
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    world = 2
    mp.spawn(_worker, args=(world, port, str(tmp_path), comm), nprocs=world, join=True)
    r = [torch.load(os.path.join(tmp_path, f"rank{k}.pt")) for k in range(world)]
    assert torch.equal(r[0]["flat0"], r[1]["flat0"]), "TrainStep must broadcast rank 0's parameters"
    assert torch.equal(r[0]["flat"], r[1]["flat"]), "ranks diverged after the step"
    assert torch.equal(r[0]["gsum"], r[1]["gsum"])
    # oracle: per-shard gradients at the common starting point, averaged
    enc, dec = synth.make_vae_state(0)
    layout = param_layout()
    offs, off = {}, 0
    for name, shape in layout:
        n = int(np.prod(shape))
        offs[name] = (off, n, shape)
        off += n
    g_mean, g_budget = {}, {}
    for k in range(world):
        x, eps = synth.make_frames(B, seed=900 + k), synth.make_eps(B, seed=910 + k)
        pred = O.critic_forward(critic_state, x)
        l_ref, _, _, _, g = O.loss_and_grads(enc, dec, x, pred, eps, update_stats=False)
        _, _, _, _, gm = O.loss_and_grads_bf16_storage(enc, dec, x, pred, eps)
        np.testing.assert_allclose(r[k]["losses"][0].item(), l_ref["total_loss"].item(), rtol=1e-3)
        for name in g:
            g_mean[name] = g_mean.get(name, 0) + g[name].double() / world
            g_budget[name] = g_budget.get(name, 0) + gm[name].double() / world
    dev_mean = r[0]["gsum"].double() / world
    rel = lambda a, b: float((a - b).norm() / b.norm().clamp_min(1e-30))
    for name, (o, n, shape) in offs.items():
        ref = g_mean[name].reshape(-1)
        if ref.norm() < 1e-7:
            continue
        got = dev_mean[o:o + n]
        budget = rel(g_budget[name].reshape(-1), ref)
        assert rel(got, ref) < max(2.0 * budget, 5e-3), (name, rel(got, ref), budget)
    # Adam on the mean gradient (fused kernel with grad_scale = 1 / world)
    p = r[0]["flat0"].double()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    O.adam_step(p, dev_mean, m, v, 1)
    np.testing.assert_allclose(r[0]["flat"].double().numpy(), p.numpy(), atol=2e-7, rtol=0)
