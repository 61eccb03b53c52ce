"""Parity of the non-GEMM kernels (pack, BatchNorm/pool/act, Linear, latent, loss, Adam, critic, mask)
against the CPU oracle / plain torch fp32 on identical inputs, through the C ABI."""
import ctypes
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import critic_vae_oracle as O
import packref
import synth
from test_conv_gemm import _native, _rand, nhwc_bf16, from_nhwc, rb

pytestmark = pytest.mark.gpu


def ptr(t):
    return t.data_ptr() if t is not None else None


def sync(L):
    torch.cuda.synchronize()
    L.check(L.lib.cvae_check_device_fault(L.stream_ptr()))


# ------------------------------------------------------------------------------------------------
def test_pack_kernel_matches_reference_packer():
    L = _native()
    cases = []
    W = _rand((64, 32, 5, 5), 1, 0.05)
    cases.append((L.PACK_FWD5, 64, 25 * 2, 32, 64, 32, W, None, packref.pack_generic(packref.gemm_weights_fwd5(W))))
    cases.append((L.PACK_DGRAD5, 32, 25 * 4, 64, 64, 32, W, None, packref.pack_generic(packref.gemm_weights_dgrad5(W))))
    W3 = _rand((256, 128, 5, 5), 2, 0.05)
    cases.append((L.PACK_FWD5, 256, 25 * 8, 128, 256, 128, W3, None, packref.pack_generic(packref.gemm_weights_fwd5(W3))))
    W0 = _rand((32, 3, 5, 5), 3, 0.1)
    cases.append((L.PACK_PAIR8, 32, 13, 8, 32, 3, W0, None, packref.pack_pair8_e0(W0)))
    Wd = _rand((64, 128, 5, 5), 4, 0.05)
    cases.append((L.PACK_PHASE_FWD, 256, 9 * 8, 128, 64, 128, Wd, None, packref.pack_generic(packref.gemm_weights_phase_fwd(Wd))))
    cases.append((L.PACK_PHASE_DGRAD, 128, 9 * 16, 256, 64, 128, Wd, None, packref.pack_generic(packref.gemm_weights_phase_dgrad(Wd))))
    W4 = _rand((3, 32, 5, 5), 5, 0.05)
    cases.append((L.PACK_PHASE_FWD, 16, 9 * 2, 32, 3, 32, W4, None, packref.pack_generic(packref.gemm_weights_phase_fwd(W4, n_pad=16))))
    cases.append((L.PACK_PHASE_DGRAD, 32, 9 * 1, 16, 3, 32, W4, None, packref.pack_generic(packref.gemm_weights_phase_dgrad(W4, k_pad=16))))
    jobs = (L.PackJob * len(cases))()
    keep, outs = [], []
    for i, (kind, n, ks, kch, co, ci, src, src2, ref) in enumerate(cases):
        s = src.contiguous().cuda()
        dst = torch.zeros(ref.numel(), dtype=torch.bfloat16, device="cuda")
        keep.append(s)
        outs.append(dst)
        jobs[i] = L.PackJob(kind=kind, n=n, ksteps=ks, k_channels=kch, cout=co, cin=ci, src=ptr(s), src2=None, dst=ptr(dst))
        assert L.lib.cvae_pack_elems(ctypes.byref(jobs[i])) == ref.numel()
    L.check(L.lib.cvae_pack_weights(jobs, len(cases), L.stream_ptr()))
    sync(L)
    for (kind, *_, ref), dst in zip(cases, outs):
        assert torch.equal(dst.cpu().view(torch.int16), ref.view(torch.int16)), f"pack kind {kind}"


def test_pack_linear_layers():
    L = _native()
    wmu, wvar = _rand((32, 4096), 6), _rand((32, 4096), 7)
    wd, bd = _rand((4096, 33), 8), _rand((4096,), 9)
    d_fc = torch.zeros(4096, 64, device="cuda")
    d_dec = torch.zeros(34, 4096, device="cuda")
    a, b, c, d = wmu.cuda(), wvar.cuda(), wd.cuda(), bd.cuda()
    jobs = (L.PackJob * 2)(L.PackJob(kind=L.PACK_FC, src=ptr(a), src2=ptr(b), dst=ptr(d_fc)),
                           L.PackJob(kind=L.PACK_DECIN, src=ptr(c), src2=ptr(d), dst=ptr(d_dec)))
    L.check(L.lib.cvae_pack_weights(jobs, 2, L.stream_ptr()))
    sync(L)
    # k' = p*256 + c  <->  k = c*16 + p
    perm = torch.arange(4096).reshape(256, 16).t().reshape(-1)      # perm[k'] = k
    ref_fc = torch.cat([wmu, wvar])[:, perm].t()
    assert torch.equal(d_fc.cpu(), ref_fc.contiguous())
    ref_dec = torch.cat([wd[perm].t(), bd[perm][None]])
    assert torch.equal(d_dec.cpu(), ref_dec.contiguous())


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,C,HW,act,training", [(3, 32, 64, 0, True), (4, 64, 32, 0, True), (5, 256, 8, 1, True),
                                                (2, 128, 16, 0, False)])
def test_bn_pool_act_forward_and_backward(B, C, HW, act, training):
    L = _native()
    xc = rb(_rand((B, C, HW, HW), 10) * 1.5 + 0.2)                   # bias-free conv output (bf16 values)
    gamma, beta, cbias = 1 + _rand((C,), 11, 0.3), _rand((C,), 12, 0.3), _rand((C,), 13, 0.2)
    rmean, rvar = _rand((C,), 14, 0.1), 0.5 + _rand((C,), 15, 0.2).abs()
    # oracle: the reference module sees conv output WITH bias
    xr = (xc + cbias[None, :, None, None]).double().requires_grad_(True)
    rm_ref, rv_ref = rmean.double().clone(), rvar.double().clone()
    g64, b64 = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    bn = F.batch_norm(xr, rm_ref, rv_ref, g64, b64, training=training, momentum=0.1, eps=1e-5)
    pooled = F.max_pool2d(bn, 2)
    y_ref = torch.tanh(pooled) if act else torch.relu(pooled)

    x_dev = nhwc_bf16(xc)
    stats = torch.stack([xc.double().sum((0, 2, 3)), (xc.double() ** 2).sum((0, 2, 3))]).cuda()
    ss = torch.zeros(4, C, device="cuda")
    rm_d, rv_d = rmean.cuda(), rvar.cuda()
    nbt = torch.zeros(1, dtype=torch.int64, device="cuda")
    gam_d, bet_d, cb_d = gamma.cuda(), beta.cuda(), cbias.cuda()
    L.check(L.lib.cvae_bn_finalize(C, B * HW * HW, int(training), ptr(stats), ptr(gam_d), ptr(bet_d), ptr(cb_d),
                                   ptr(rm_d), ptr(rv_d), ptr(nbt), 0.1, 1e-5, ptr(ss), L.stream_ptr()))
    y = torch.zeros(B, HW // 2, HW // 2, C, dtype=torch.bfloat16, device="cuda")
    xh = torch.zeros(B, HW // 2, HW // 2, C, dtype=torch.bfloat16, device="cuda")
    am = torch.zeros(B, HW // 2, HW // 2, C // 8, dtype=torch.int16, device="cuda")
    L.check(L.lib.cvae_bn_pool_act_fwd(B, HW, HW, C, act, ptr(x_dev), ptr(ss), ptr(y), ptr(xh) if training else None,
                                       ptr(am) if training else None, L.stream_ptr()))
    sync(L)
    np.testing.assert_allclose(from_nhwc(y).numpy(), y_ref.detach().float().numpy(), rtol=2 ** -7, atol=2e-3)
    # the fused finalize + forward launch (the one the engine uses) is bit-identical to the two-launch form
    ss2, rm2, rv2, nbt2 = torch.zeros(4, C, device="cuda"), rmean.cuda(), rvar.cuda(), torch.zeros(1, dtype=torch.int64, device="cuda")
    y2, xh2, am2 = torch.zeros_like(y), torch.zeros_like(xh), torch.zeros_like(am)
    L.check(L.lib.cvae_bn_fwd(B, HW, HW, C, act, int(training), ptr(x_dev), ptr(stats), ptr(gam_d), ptr(bet_d), ptr(cb_d), ptr(rm2), ptr(rv2),
                              ptr(nbt2), 0.1, 1e-5, ptr(ss2), ptr(y2), ptr(xh2) if training else None, ptr(am2) if training else None, L.stream_ptr()))
    sync(L)
    assert torch.equal(y2, y) and torch.equal(ss2, ss) and torch.equal(rm2, rm_d) and torch.equal(rv2, rv_d) and torch.equal(nbt2, nbt)
    if training:
        assert torch.equal(xh2, xh) and torch.equal(am2, am)
    if training:
        np.testing.assert_allclose(rm_d.cpu().numpy(), rm_ref.float().numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(rv_d.cpu().numpy(), rv_ref.float().numpy(), rtol=1e-5, atol=1e-6)
        assert int(nbt.item()) == 1
        # backward
        dy = rb(_rand((B, C, HW // 2, HW // 2), 16))
        y_ref.backward(dy.double())
        sums = torch.zeros(2, C, dtype=torch.float64, device="cuda")
        dconv = torch.zeros(B, HW, HW, C, dtype=torch.bfloat16, device="cuda")
        dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
        # the kernel differentiates through ITS OWN forward output (bf16 y)
        dy_d = nhwc_bf16(dy)
        L.check(L.lib.cvae_bn_pool_act_bwd(B, HW, HW, C, act, ptr(x_dev), ptr(y), ptr(dy_d), ptr(xh), ptr(am), ptr(ss), ptr(gam_d),
                                           ptr(sums), ptr(dconv), ptr(dg), ptr(db), L.stream_ptr()))
        sync(L)
        ref_dx = xr.grad.float()
        scale = ref_dx.abs().max().item()
        np.testing.assert_allclose(from_nhwc(dconv).numpy(), ref_dx.numpy(), rtol=2e-2, atol=1e-2 * scale)
        np.testing.assert_allclose(dg.cpu().numpy(), g64.grad.float().numpy(), rtol=2e-2, atol=2e-2 * g64.grad.abs().max().item())
        np.testing.assert_allclose(db.cpu().numpy(), b64.grad.float().numpy(), rtol=2e-2, atol=2e-2 * b64.grad.abs().max().item())


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B", [1, 5, 19])
def test_linear_layers(B):
    L = _native()
    wmu, wvar, bmu, bvar = _rand((32, 4096), 20, 0.02), _rand((32, 4096), 21, 0.02), _rand((32,), 22, 0.1), _rand((32,), 23, 0.1)
    wd, bd = _rand((4096, 33), 24, 0.2), _rand((4096,), 25, 0.1)
    a = rb(torch.tanh(_rand((B, 256, 4, 4), 26, 2.0)))                 # encoder output, NCHW view
    a_req = a.double().requires_grad_(True)
    W = [t.double().requires_grad_(True) for t in (wmu, wvar, bmu, bvar, wd, bd)]
    flat = torch.flatten(a_req, 1)
    mu, lv = F.linear(flat, W[0], W[2]), F.linear(flat, W[1], W[3])

    dev = [t.cuda() for t in (wmu, wvar, bmu, bvar, wd, bd)]
    wfc, wdec = torch.zeros(4096, 64, device="cuda"), torch.zeros(34, 4096, device="cuda")
    jobs = (L.PackJob * 2)(L.PackJob(kind=L.PACK_FC, src=ptr(dev[0]), src2=ptr(dev[1]), dst=ptr(wfc)),
                           L.PackJob(kind=L.PACK_DECIN, src=ptr(dev[4]), src2=ptr(dev[5]), dst=ptr(wdec)))
    L.check(L.lib.cvae_pack_weights(jobs, 2, L.stream_ptr()))
    a_dev = nhwc_bf16(a)
    ml = torch.empty(B, 64, device="cuda")
    L.check(L.lib.cvae_fc_fwd(B, ptr(a_dev), ptr(wfc), ptr(dev[2]), ptr(dev[3]), ptr(ml), L.stream_ptr()))
    sync(L)
    np.testing.assert_allclose(ml.cpu().numpy(), torch.cat([mu, lv], 1).detach().float().numpy(), rtol=1e-4, atol=1e-4)

    # latent + decoder_input forward
    eps, pred = _rand((B, 32), 27), torch.rand(B, 1, generator=torch.Generator().manual_seed(28))
    z = O.reparametrize(mu, lv, eps.double())
    h = F.linear(torch.cat((z, pred.double()), 1), W[4], W[5]).view(-1, 256, 4, 4)
    zc = torch.empty(B, 33, device="cuda")
    eps_d, pred_d = eps.cuda(), pred.cuda()      # keep device copies alive across the async launches
    L.check(L.lib.cvae_latent_fwd(B, 1, ptr(ml), ptr(eps_d), ptr(pred_d), ptr(zc), None, L.stream_ptr()))
    h_dev = torch.empty(B, 4, 4, 256, dtype=torch.bfloat16, device="cuda")
    L.check(L.lib.cvae_decin_fwd(B, ptr(zc), ptr(wdec), ptr(h_dev), L.stream_ptr()))
    sync(L)
    np.testing.assert_allclose(zc.cpu().numpy(), torch.cat((z, pred.double()), 1).detach().float().numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(from_nhwc(h_dev).numpy(), h.detach().float().numpy(), rtol=2 ** -7, atol=2e-3)

    # backward: upstream dh (bf16) and direct gradients on mu / logvar
    dh = rb(_rand((B, 256, 4, 4), 29))
    dmu_e, dlv_e = _rand((B, 32), 30), _rand((B, 32), 31)
    (h * dh.double()).sum().backward(retain_graph=True)
    (mu * dmu_e.double()).sum().backward(retain_graph=True)
    (lv * dlv_e.double()).sum().backward()
    dzc = torch.empty(B, 33, device="cuda")
    dwd, dbd = torch.empty(4096, 33, device="cuda"), torch.empty(4096, device="cuda")
    dh_d, dmu_d, dlv_d = nhwc_bf16(dh), dmu_e.cuda(), dlv_e.cuda()
    L.check(L.lib.cvae_decin_bwd(B, ptr(dh_d), ptr(zc), ptr(wdec), ptr(dzc), ptr(dwd), ptr(dbd), L.stream_ptr()))
    dml = torch.empty(B, 64, device="cuda")
    L.check(L.lib.cvae_latent_bwd(B, ptr(ml), ptr(eps_d), ptr(dzc), ptr(dmu_d), ptr(dlv_d), 0.0, ptr(dml), L.stream_ptr()))
    da = torch.empty(B, 4, 4, 256, dtype=torch.bfloat16, device="cuda")
    dwmu, dwvar = torch.empty(32, 4096, device="cuda"), torch.empty(32, 4096, device="cuda")
    dbmu, dbvar = torch.empty(32, device="cuda"), torch.empty(32, device="cuda")
    L.check(L.lib.cvae_fc_bwd(B, ptr(dml), ptr(a_dev), ptr(wfc), ptr(da), ptr(dwmu), ptr(dwvar), ptr(dbmu), ptr(dbvar), L.stream_ptr()))
    sync(L)

    def close(got, ref, tol=2e-3):
        ref = ref.float()
        np.testing.assert_allclose(got.cpu().numpy(), ref.numpy(), rtol=tol, atol=tol * ref.abs().max().item())
    close(dwd, W[4].grad); close(dbd, W[5].grad)
    close(dwmu, W[0].grad); close(dwvar, W[1].grad); close(dbmu, W[2].grad); close(dbvar, W[3].grad)
    close(from_nhwc(da), a_req.grad, tol=1e-2)


@pytest.mark.parametrize("B", [1, 19, 64, 256, 290])      # 256, 290: 20 rows per cluster (16 would need a 16th+ cluster, a second wave)
def test_bottleneck_fused_equals_the_three_separate_kernels(B):
    """cvae_bottleneck_fwd / _bwd (one launch per direction, the training path) against fc -> latent -> decoder_input and
    their data gradients run one by one (checked against torch above): bit for bit, including ragged row blocks."""
    L = _native()
    wmu, wvar, bmu, bvar = _rand((32, 4096), 20, 0.02), _rand((32, 4096), 21, 0.02), _rand((32,), 22, 0.1), _rand((32,), 23, 0.1)
    wd, bd = _rand((4096, 33), 24, 0.2), _rand((4096,), 25, 0.1)
    dev = [t.cuda() for t in (wmu, wvar, bmu, bvar, wd, bd)]
    wfc, wdec = torch.zeros(4096, 64, device="cuda"), torch.zeros(34, 4096, device="cuda")
    jobs = (L.PackJob * 2)(L.PackJob(kind=L.PACK_FC, src=ptr(dev[0]), src2=ptr(dev[1]), dst=ptr(wfc)),
                           L.PackJob(kind=L.PACK_DECIN, src=ptr(dev[4]), src2=ptr(dev[5]), dst=ptr(wdec)))
    L.check(L.lib.cvae_pack_weights(jobs, 2, L.stream_ptr()))
    a_dev = nhwc_bf16(rb(torch.tanh(_rand((B, 256, 4, 4), 26, 2.0))))
    eps_d, pred_d = _rand((B, 32), 27).cuda(), torch.rand(B, generator=torch.Generator().manual_seed(28)).cuda()
    s = L.stream_ptr()
    new = lambda *shape, dt=torch.float32: torch.full(shape, float("nan"), dtype=dt, device="cuda")
    ml1, zc1, h1 = new(B, 64), new(B, 33), new(B, 4096, dt=torch.bfloat16)
    L.check(L.lib.cvae_fc_fwd(B, ptr(a_dev), ptr(wfc), ptr(dev[2]), ptr(dev[3]), ptr(ml1), s))
    L.check(L.lib.cvae_latent_fwd(B, 1, ptr(ml1), ptr(eps_d), ptr(pred_d), ptr(zc1), None, s))
    L.check(L.lib.cvae_decin_fwd(B, ptr(zc1), ptr(wdec), ptr(h1), s))
    ml2, zc2, h2 = new(B, 64), new(B, 33), new(B, 4096, dt=torch.bfloat16)
    L.check(L.lib.cvae_bottleneck_fwd(B, ptr(a_dev), ptr(wfc), ptr(dev[2]), ptr(dev[3]), ptr(eps_d), ptr(pred_d), ptr(wdec), ptr(ml2), ptr(zc2),
                                      ptr(h2), s))
    sync(L)
    assert torch.equal(ml1, ml2) and torch.equal(zc1, zc2) and torch.equal(h1, h2)
    # backward
    dh = nhwc_bf16(rb(_rand((B, 256, 4, 4), 29))).reshape(B, 4096)
    k = 0.0001 / B
    dzc1, dml1, da1 = new(B, 33), new(B, 64), new(B, 4096, dt=torch.bfloat16)
    L.check(L.lib.cvae_decin_bwd(B, ptr(dh), None, ptr(wdec), ptr(dzc1), None, None, s))
    L.check(L.lib.cvae_latent_bwd(B, ptr(ml1), ptr(eps_d), ptr(dzc1), None, None, k, ptr(dml1), s))
    L.check(L.lib.cvae_fc_bwd(B, ptr(dml1), None, ptr(wfc), ptr(da1), None, None, None, None, s))
    dzc2, dml2, da2 = new(B, 33), new(B, 64), new(B, 4096, dt=torch.bfloat16)
    L.check(L.lib.cvae_bottleneck_bwd(B, ptr(dh), ptr(wdec), ptr(ml1), ptr(eps_d), k, ptr(wfc), ptr(dzc2), ptr(dml2), ptr(da2), s))
    sync(L)
    assert torch.equal(dzc1, dzc2) and torch.equal(dml1, dml2) and torch.equal(da1, da2)
    L.check(L.lib.cvae_bottleneck_bwd(B, ptr(dh), ptr(wdec), ptr(ml1), ptr(eps_d), k, ptr(wfc), None, ptr(dml2), ptr(da2), s))   # d_z_pred is optional
    sync(L)
    assert torch.equal(da1, da2)


# ------------------------------------------------------------------------------------------------
def _window():
    return (ctypes.c_float * 11)(*[float(v) for v in O.msssim_window_1d()])


@pytest.mark.parametrize("tag,B", [("a", 2), ("b", 3)])
def test_loss_forward_backward_vs_golden(golden_dir, tag, B):
    """MS-SSIM + KLD through cvae_loss_fwd/bwd against the reference's MSSIM (golden) and the oracle."""
    L = _native()
    g = np.load(os.path.join(golden_dir, "msssim.npz"))
    x = synth.make_frames(B, seed=51 + B)
    r = torch.from_numpy(g[f"{tag}_recon"])
    mu, lv = _rand((B, 32), 40), _rand((B, 32), 41)
    ml = torch.cat([mu, lv], 1).cuda()
    sums = torch.zeros(10, dtype=torch.float64, device="cuda")
    coef, losses = torch.zeros(8, device="cuda"), torch.zeros(3, device="cuda")
    rd, xd = r.cuda(), x.cuda()
    L.check(L.lib.cvae_loss_fwd(B, ptr(rd), ptr(xd), ptr(ml), None, _window(), 0.001, ptr(sums), ptr(coef), ptr(losses), L.stream_ptr()))
    dr = torch.zeros_like(rd)
    dmu, dlv = torch.zeros(B, 32, device="cuda"), torch.zeros(B, 32, device="cuda")
    L.check(L.lib.cvae_loss_bwd(B, ptr(rd), ptr(xd), ptr(ml), _window(), 0.001, ptr(coef), None, ptr(dr), ptr(dmu), ptr(dlv), L.stream_ptr()))
    sync(L)
    # the split form (level sums | scalars | gradient straight from the sums) gives the same bits
    sums2 = torch.full((10,), float("nan"), dtype=torch.float64, device="cuda")
    coef2, losses2, dr2 = torch.zeros_like(coef), torch.zeros_like(losses), torch.zeros_like(rd)
    L.check(L.lib.cvae_loss_sums(B, ptr(rd), ptr(xd), _window(), ptr(sums2), L.stream_ptr()))
    sync(L)
    np.testing.assert_allclose(sums2.cpu().numpy(), sums.cpu().numpy(), rtol=1e-12)      # (double atomics: the order of the CTAs is free)
    sums2.copy_(sums)
    L.check(L.lib.cvae_loss_finalize(B, ptr(ml), None, ptr(sums2), 0.001, ptr(coef2), ptr(losses2), L.stream_ptr()))
    L.check(L.lib.cvae_loss_bwd_sums(B, ptr(rd), ptr(xd), _window(), ptr(sums2), None, ptr(dr2), L.stream_ptr()))
    sync(L)
    assert torch.equal(losses, losses2) and torch.equal(coef, coef2) and torch.equal(dr, dr2)
    means = O.msssim_level_means(r, x)
    for l in range(5):
        n = B * 3 * (64 >> l) ** 2
        np.testing.assert_allclose(sums[l].item() / n, means[l][1].item(), rtol=2e-5, atol=1e-7)
        np.testing.assert_allclose(sums[5 + l].item() / n, means[l][0].item(), rtol=2e-5, atol=1e-7)
    mu_r, lv_r = mu.clone().requires_grad_(True), lv.clone().requires_grad_(True)
    kld = O.kld_loss(mu_r, lv_r)
    kld.backward()
    np.testing.assert_allclose(losses[1].item(), float(g[f"{tag}_loss"]), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(losses[2].item(), kld.item(), rtol=1e-5)
    np.testing.assert_allclose(losses[0].item(), float(g[f"{tag}_loss"]) + kld.item(), rtol=1e-4)
    gref = g[f"{tag}_grad"]
    np.testing.assert_allclose(dr.cpu().numpy(), gref, rtol=2e-3, atol=2e-3 * np.abs(gref).max())
    np.testing.assert_allclose(dmu.cpu().numpy(), mu_r.grad.numpy(), rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(dlv.cpu().numpy(), lv_r.grad.numpy(), rtol=1e-5, atol=1e-9)


def test_loss_nan_semantics():
    """A negative cs mean must give NaN, as `mcs ** weights` does in the reference (vae_nets.py:243)."""
    L = _native()
    B = 2
    x = synth.make_frames(B, seed=60)
    r = (1.0 - x).contiguous()          # anti-correlated: negative covariance at every level
    assert torch.isnan(O.msssim_loss(r, x))
    ml = torch.zeros(B, 64, device="cuda")
    sums = torch.zeros(10, dtype=torch.float64, device="cuda")
    coef, losses = torch.zeros(8, device="cuda"), torch.zeros(3, device="cuda")
    rd, xd = r.cuda(), x.cuda()
    L.check(L.lib.cvae_loss_fwd(B, ptr(rd), ptr(xd), ptr(ml), None, _window(), 0.001, ptr(sums), ptr(coef), ptr(losses), L.stream_ptr()))
    sync(L)
    assert torch.isnan(losses[0]) and torch.isnan(losses[1]) and losses[2].item() == 0.0


@pytest.mark.parametrize("B", [1, 5, 64, 65, 200, 4099])
def test_fused_latent_kld_kernels(B):
    """The vectorised latent kernel: z | pred bit-identical to fma(eps, exp(0.5 logvar), mu) per element (the same
    expression the scalar round-1 kernel evaluated), KL partial sums per 64 rows equal to the fp64 sum of the
    per-element fp32 terms, loss_fwd fed with the partials == loss_fwd reducing mu / logvar itself, and the
    backward with the KL term folded in == latent_bwd + kld_bwd (vae_nets.py:48-51,57-58,143)."""
    L = _native()
    g = torch.Generator().manual_seed(B)
    ml = torch.randn(B, 64, generator=g)
    eps, pred = torch.randn(B, 32, generator=g), torch.rand(B, generator=g)
    ml_d, eps_d, pred_d = ml.cuda(), eps.cuda(), pred.cuda()
    zc = torch.full((B, 33), float("nan"), device="cuda")
    nparts = L.lib.cvae_latent_kld_partials(B)
    assert nparts == (B + 63) // 64
    parts = torch.full((nparts,), float("nan"), dtype=torch.float64, device="cuda")
    L.check(L.lib.cvae_latent_fwd(B, 1, ptr(ml_d), ptr(eps_d), ptr(pred_d), ptr(zc), ptr(parts), L.stream_ptr()))
    zc0 = torch.empty(B, 33, device="cuda")
    L.check(L.lib.cvae_latent_fwd(B, 0, ptr(ml_d), None, ptr(pred_d), ptr(zc0), None, L.stream_ptr()))
    sync(L)
    # device expression on the device (torch's exp / fma differ from expf in the last bit, so compare to 2 ulp and
    # pin the exact bits against a device-side evaluation of the same formula)
    mu, lv = ml_d[:, :32], ml_d[:, 32:]
    ref = torch.cat((torch.addcmul(mu, eps_d, torch.exp(0.5 * lv)), pred_d[:, None]), 1)
    np.testing.assert_allclose(zc.cpu().numpy(), ref.cpu().numpy(), rtol=3e-7, atol=1e-7)
    assert torch.equal(zc[:, 32], pred_d) and torch.equal(zc0[:, :32], mu) and torch.equal(zc0[:, 32], pred_d)
    terms = (1.0 + lv - mu * mu - torch.exp(lv)).double().cpu()
    want = torch.stack([terms[i * 64:(i + 1) * 64].sum() for i in range(nparts)])
    np.testing.assert_allclose(parts.cpu().numpy(), want.numpy(), rtol=1e-6, atol=1e-4)
    # loss_fwd with the partials == loss_fwd without
    x = synth.make_frames(min(B, 4), seed=61)
    if B <= 4:
        r = (x * 0.9 + 0.05).contiguous()
        sums = torch.zeros(10, dtype=torch.float64, device="cuda")
        coef, l0, l1 = torch.zeros(8, device="cuda"), torch.zeros(3, device="cuda"), torch.zeros(3, device="cuda")
        rd, xd = r.cuda(), x.cuda()
        L.check(L.lib.cvae_loss_fwd(B, ptr(rd), ptr(xd), ptr(ml_d), None, _window(), 0.001, ptr(sums), ptr(coef), ptr(l0), L.stream_ptr()))
        L.check(L.lib.cvae_loss_fwd(B, ptr(rd), ptr(xd), None, ptr(parts), _window(), 0.001, ptr(sums), ptr(coef), ptr(l1), L.stream_ptr()))
        sync(L)
        np.testing.assert_allclose(l1.cpu().numpy(), l0.cpu().numpy(), rtol=1e-6)
    # backward: folded KL term == separate kld_bwd + latent_bwd
    dzc = torch.randn(B, 33, generator=g).cuda()
    k = 0.001 / B
    dmu, dlv = k * mu, k * 0.5 * (torch.exp(lv) - 1.0)
    d_sep, d_fused = torch.empty(B, 64, device="cuda"), torch.empty(B, 64, device="cuda")
    L.check(L.lib.cvae_latent_bwd(B, ptr(ml_d), ptr(eps_d), ptr(dzc), ptr(dmu.contiguous()), ptr(dlv.contiguous()), 0.0, ptr(d_sep), L.stream_ptr()))
    L.check(L.lib.cvae_latent_bwd(B, ptr(ml_d), ptr(eps_d), ptr(dzc), None, None, k, ptr(d_fused), L.stream_ptr()))
    sync(L)
    want_mu = dzc[:, :32] + dmu
    want_lv = dzc[:, :32] * eps_d * 0.5 * torch.exp(0.5 * lv) + dlv
    np.testing.assert_allclose(d_sep.cpu().numpy(), torch.cat((want_mu, want_lv), 1).cpu().numpy(), rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(d_fused.cpu().numpy(), d_sep.cpu().numpy(), rtol=2e-6, atol=1e-7)


def test_adam_matches_torch():
    L = _native()
    n = 100003
    p0, g1, g2 = _rand((n,), 50), _rand((n,), 51, 0.01), _rand((n,), 52, 0.01)
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=5e-5)
    p, m, v = p0.cuda(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    step = torch.zeros(1, dtype=torch.int64, device="cuda")
    for gr in (g1, g2):
        ref.grad = gr.clone()
        opt.step()
        gd = (gr * 4).cuda()            # pretend a 4-rank sum, undone by grad_scale
        L.check(L.lib.cvae_adam_step(n, ptr(p), ptr(gd), ptr(m), ptr(v), ptr(step), 5e-5, 0.9, 0.999, 1e-8, 0.25, L.stream_ptr()))
    sync(L)
    assert int(step.item()) == 2
    np.testing.assert_allclose(p.cpu().numpy(), ref.detach().numpy(), rtol=0, atol=2e-7)


def test_critic_matches_reference(golden_dir, critic_state):
    L = _native()
    g = np.load(os.path.join(golden_dir, "critic.npz"))
    x = synth.make_frames(16, seed=30)
    w = torch.cat([critic_state[k].flatten() for k in critic_state]).cuda()
    assert w.numel() == L.lib.cvae_critic_param_count()
    pred = torch.zeros(16, device="cuda")
    xd = x.cuda()
    L.check(L.lib.cvae_critic_fwd(16, ptr(xd), ptr(w), ptr(pred), L.stream_ptr()))
    sync(L)
    np.testing.assert_allclose(pred.cpu().numpy(), g["pred"].ravel(), rtol=0, atol=5e-6)
    np.testing.assert_allclose(pred.cpu().numpy(), O.critic_forward(critic_state, x).numpy().ravel(), rtol=0, atol=5e-6)


def test_mask_pipeline_bit_exact(golden_dir):
    L = _native()
    g = np.load(os.path.join(golden_dir, "mask_pipeline.npz"))
    N = g["diff"].shape[0]
    hi, lo = torch.from_numpy(g["recon_one"]).cuda(), torch.from_numpy(g["recon_zero"]).cuda()
    diff = torch.zeros(N, 64, 64, dtype=torch.float64, device="cuda")
    mx = torch.zeros(N, dtype=torch.float64, device="cuda")
    L.check(L.lib.cvae_diff_grey(N, ptr(hi), ptr(lo), ptr(diff), ptr(mx), L.stream_ptr()))
    sync(L)
    ord_d = np.stack([O.diff_grey_ordered(g["recon_one"][i], g["recon_zero"][i])[0] for i in range(N)])
    assert np.array_equal(diff.cpu().numpy(), ord_d)                       # bit-exact vs the ordered fp64 form
    np.testing.assert_allclose(diff.cpu().numpy(), g["diff"], rtol=4e-16, atol=1e-18)   # <= 1 ulp vs BLAS np.dot
    assert np.array_equal(mx.cpu().numpy(), ord_d.reshape(N, -1).max(1))

    thr_list = list(range(0, 130, 10))
    for dkey, u8key, mkey, ious in (("diff", "diff_u8", "thr_mask_50", list(g["iou_sweep"])),
                                   ("diff2", "diff2_u8", "thr2_mask_50", None)):
        d = g[dkey]
        factor, mean_max = O.diff_factor([np.amax(x) for x in d] if dkey == "diff2" else list(g["max_values"]))
        dd = torch.from_numpy(d).cuda()
        gt = torch.from_numpy(g["gt"].astype(np.uint8)).cuda()
        u8 = torch.zeros(N, 64, 64, dtype=torch.uint8, device="cuda")
        mk = torch.zeros(N, 64, 64, dtype=torch.uint8, device="cuda")
        hist = torch.zeros(512, dtype=torch.int64, device="cuda")
        thr_d = torch.tensor(thr_list, dtype=torch.int32, device="cuda")
        counts = torch.zeros(len(thr_list), 3, dtype=torch.int64, device="cuda")
        L.check(L.lib.cvae_mask_iou(N, ptr(dd), ptr(gt), float(mean_max), float(factor), 50, len(thr_list), ptr(thr_d),
                                    ptr(u8), ptr(mk), ptr(hist), ptr(counts), L.stream_ptr()))
        sync(L)
        assert np.array_equal(u8.cpu().numpy(), g[u8key])
        assert np.array_equal(mk.cpu().numpy().astype(bool), g[mkey])
        for t, c in zip(thr_list, counts.cpu().numpy()):
            _, tm = O.diff_and_thr_masks(list(d), [np.amax(x) for x in d] if dkey == "diff2" else list(g["max_values"]), thr=t)
            assert tuple(c) == O.iou_counts(g["gt"], tm)
        if ious is not None:
            got = [round(1 if c.sum() == 0 else c[0] / c.sum(), 3) for c in counts.cpu().numpy()]
            assert got == ious
    # empty input
    L.check(L.lib.cvae_mask_iou(0, None, None, 1.0, 1.0, 50, 0, None, None, None, ptr(hist), None, L.stream_ptr()))
    sync(L)
    assert int(hist.sum().item()) == 0
