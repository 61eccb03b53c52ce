"""tcgen05 implicit-GEMM convolution (csrc/conv_gemm.cu) against torch CPU convolutions on the same
bf16-rounded operands.  Weights are packed by the CPU reference packer (tests/packref.py) so these
tests isolate the kernel from the device packing kernel."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import packref

pytestmark = pytest.mark.gpu


def _native():
    import cvae_native.binding as L
    return L


def rb(t):
    """round to bf16 and back (the operand values the tensor core sees)"""
    return t.to(torch.bfloat16).float()


def nhwc_bf16(t):
    return t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()


def from_nhwc(t):
    return t.float().cpu().permute(0, 3, 1, 2).contiguous()


_WS = {}


def run_conv(L, *, B, H, W, ksize, src_channels, n_total, loader, epilogue, ktab, src, wpack, out,
             src2=None, bias=None, act=None, stats=None, tm=0, stack=0, repeat=1):
    d = L.ConvDesc(batch=B, height=H, width=W, ksize=ksize, src_channels=src_channels, n_total=n_total,
                   loader=loader, epilogue=epilogue, ktab=ktab, tm=tm, stack=stack,
                   src=src.data_ptr(), src2=src2.data_ptr() if src2 is not None else None,
                   wpack=wpack.data_ptr(), bias=bias.data_ptr() if bias is not None else None,
                   act=act.data_ptr() if act is not None else None, out=out.data_ptr(),
                   stats=stats.data_ptr() if stats is not None else None)
    need = int(L.lib.cvae_conv_gemm_workspace_bytes(ctypes.byref(d)))
    assert need >= 0, L.lib.cvae_last_error()
    if need:        # split-K scratch: zeroed ONCE (the kernel must leave its arrival counters at zero for the next call)
        if "buf" not in _WS or _WS["buf"].numel() < need:
            _WS["buf"] = torch.zeros(need, dtype=torch.uint8, device="cuda")
        d.workspace, d.workspace_bytes = _WS["buf"].data_ptr(), _WS["buf"].numel()
    for _ in range(repeat):
        L.check(L.lib.cvae_conv_gemm(ctypes.byref(d), L.stream_ptr()))
    torch.cuda.synchronize()
    L.check(L.lib.cvae_check_device_fault(L.stream_ptr()))
    return need


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(shape, generator=g) * 2 - 1) * scale


@pytest.mark.parametrize("B,Cin,Cout,HW,tm", [(3, 32, 64, 32, 0), (5, 64, 128, 16, 0), (7, 128, 256, 8, 0),
                                             (2, 32, 64, 32, 1), (9, 64, 128, 16, 4)])
def test_encoder_conv_with_stats(B, Cin, Cout, HW, tm):
    L = _native()
    x, Wt = rb(_rand((B, Cin, HW, HW), 1)), rb(_rand((Cout, Cin, 5, 5), 2, 0.05))
    ref = F.conv2d(x.double(), Wt.double(), padding=2).float()
    wp = packref.pack_generic(packref.gemm_weights_fwd5(Wt)).cuda()
    out = torch.zeros(B, HW, HW, Cout, dtype=torch.bfloat16, device="cuda")
    stats = torch.zeros(2, Cout, dtype=torch.float64, device="cuda")
    run_conv(L, B=B, H=HW, W=HW, ksize=5, src_channels=Cin, n_total=Cout, loader=L.LOAD_NHWC,
             epilogue=L.EPI_STATS, ktab=L.KTAB_GENERIC, src=nhwc_bf16(x), wpack=wp, out=out, stats=stats, tm=tm)
    got = from_nhwc(out)
    assert torch.allclose(got, ref, rtol=1e-2, atol=2e-2 * ref.abs().max().item() / 8)
    np.testing.assert_allclose(got.numpy(), rb(ref).numpy(), rtol=2 ** -7, atol=1e-3)
    g64 = got.double()
    np.testing.assert_allclose(stats[0].cpu().numpy(), g64.sum((0, 2, 3)).numpy(), rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(stats[1].cpu().numpy(), (g64 * g64).sum((0, 2, 3)).numpy(), rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("B", [1, 3])
def test_encoder_conv0_from_nchw_frames(B):
    L = _native()
    x = torch.rand(B, 3, 64, 64, generator=torch.Generator().manual_seed(3))
    Wt = rb(_rand((32, 3, 5, 5), 4, 0.1))
    ref = F.conv2d(rb(x).double(), Wt.double(), padding=2).float()
    out = torch.zeros(B, 64, 64, 32, dtype=torch.bfloat16, device="cuda")
    stats = torch.zeros(2, 32, dtype=torch.float64, device="cuda")
    run_conv(L, B=B, H=64, W=64, ksize=5, src_channels=8, n_total=32, loader=L.LOAD_NCHW3,
             epilogue=L.EPI_STATS, ktab=L.KTAB_PAIR8, src=x.cuda(), wpack=packref.pack_pair8_e0(Wt).cuda(),
             out=out, stats=stats)
    got = from_nhwc(out)
    np.testing.assert_allclose(got.numpy(), rb(ref).numpy(), rtol=2 ** -7, atol=1e-3)
    np.testing.assert_allclose(stats[0].cpu().numpy(), got.double().sum((0, 2, 3)).numpy(), rtol=1e-5, atol=1e-3)


def test_decoder_conv0_bias_relu():
    L = _native()
    B = 6
    x, Wt, b = rb(_rand((B, 256, 4, 4), 5)), rb(_rand((128, 256, 5, 5), 6, 0.02)), _rand((128,), 7, 0.1)
    ref = torch.relu(F.conv2d(x.double(), Wt.double(), b.double(), padding=2)).float()
    out = torch.zeros(B, 4, 4, 128, dtype=torch.bfloat16, device="cuda")
    run_conv(L, B=B, H=4, W=4, ksize=5, src_channels=256, n_total=128, loader=L.LOAD_NHWC,
             epilogue=L.EPI_BIAS_RELU, ktab=L.KTAB_GENERIC, src=nhwc_bf16(x),
             wpack=packref.pack_generic(packref.gemm_weights_fwd5(Wt)).cuda(), out=out, bias=b.cuda())
    np.testing.assert_allclose(from_nhwc(out).numpy(), rb(ref).numpy(), rtol=2 ** -7, atol=2e-3)


@pytest.mark.parametrize("B,Cin,Cout,HW", [(3, 128, 64, 4), (3, 64, 32, 8), (2, 32, 32, 16)])
def test_decoder_upsample_folded_conv(B, Cin, Cout, HW):
    """conv5x5(upsample2(x)) + bias + ReLU computed as a 3x3 conv with 4*Cout phase channels."""
    L = _native()
    x, Wt, b = rb(_rand((B, Cin, HW, HW), 8)), _rand((Cout, Cin, 5, 5), 9, 0.05), _rand((Cout,), 10, 0.1)
    Wg = packref.gemm_weights_phase_fwd(Wt)
    # the kernel sees bf16-rounded *effective* weights; build the reference from the same values
    Weff = rb(Wg).reshape(2, 2, Cout, 3, 3, Cin)
    ref = torch.zeros(B, Cout, 2 * HW, 2 * HW, dtype=torch.float64)
    for a in (0, 1):
        for bb in (0, 1):
            w3 = Weff[a, bb].permute(0, 3, 1, 2).double()      # [Co][Ci][3][3]
            ref[:, :, a::2, bb::2] = F.conv2d(x.double(), w3, b.double(), padding=1)
    ref = torch.relu(ref).float()
    # and the folded form equals the reference op up to weight rounding
    direct = torch.relu(F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), Wt, b, padding=2))
    assert torch.allclose(direct, ref, atol=3e-2 * direct.abs().max().item())
    out = torch.zeros(B, 2 * HW, 2 * HW, Cout, dtype=torch.bfloat16, device="cuda")
    run_conv(L, B=B, H=HW, W=HW, ksize=3, src_channels=Cin, n_total=4 * Cout, loader=L.LOAD_NHWC,
             epilogue=L.EPI_PHASE_BIAS_RELU, ktab=L.KTAB_GENERIC, src=nhwc_bf16(x),
             wpack=packref.pack_generic(Wg).cuda(), out=out, bias=b.cuda())
    np.testing.assert_allclose(from_nhwc(out).numpy(), rb(ref).numpy(), rtol=2 ** -7, atol=2e-3)


def test_decoder_last_conv_tanh_nchw():
    L = _native()
    B = 3
    x, Wt, b = rb(_rand((B, 32, 32, 32), 11)), _rand((3, 32, 5, 5), 12, 0.05), _rand((3,), 13, 0.1)
    Wg = packref.gemm_weights_phase_fwd(Wt, n_pad=16)
    Weff = rb(Wg)[:12].reshape(2, 2, 3, 3, 3, 32)
    ref = torch.zeros(B, 3, 64, 64, dtype=torch.float64)
    for a in (0, 1):
        for bb in (0, 1):
            ref[:, :, a::2, bb::2] = F.conv2d(x.double(), Weff[a, bb].permute(0, 3, 1, 2).double(), b.double(), padding=1)
    ref = torch.tanh(ref).float()
    out = torch.zeros(B, 3, 64, 64, dtype=torch.float32, device="cuda")
    run_conv(L, B=B, H=32, W=32, ksize=3, src_channels=32, n_total=16, loader=L.LOAD_NHWC,
             epilogue=L.EPI_PHASE_BIAS_TANH, ktab=L.KTAB_GENERIC, src=nhwc_bf16(x),
             wpack=packref.pack_generic(Wg).cuda(), out=out, bias=b.cuda())
    np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), rtol=0, atol=2e-5)


@pytest.mark.parametrize("B,Cin,Cout,HW", [(3, 32, 64, 32), (4, 128, 256, 8), (5, 256, 128, 4)])
def test_dgrad_5x5(B, Cin, Cout, HW):
    """dX = conv_transpose(dY, W): same kernel, flipped/transposed packed weights."""
    L = _native()
    dy, Wt = rb(_rand((B, Cout, HW, HW), 14)), rb(_rand((Cout, Cin, 5, 5), 15, 0.05))
    ref = F.conv_transpose2d(dy.double(), Wt.double(), padding=2).float()
    out = torch.zeros(B, HW, HW, Cin, dtype=torch.bfloat16, device="cuda")
    run_conv(L, B=B, H=HW, W=HW, ksize=5, src_channels=Cout, n_total=Cin, loader=L.LOAD_NHWC,
             epilogue=L.EPI_PLAIN, ktab=L.KTAB_GENERIC, src=nhwc_bf16(dy),
             wpack=packref.pack_generic(packref.gemm_weights_dgrad5(Wt)).cuda(), out=out)
    np.testing.assert_allclose(from_nhwc(out).numpy(), rb(ref).numpy(), rtol=2 ** -7, atol=2e-3 * ref.abs().max().item())


@pytest.mark.parametrize("B,Cin,Cout,HW", [(3, 128, 64, 4), (2, 64, 32, 8), (2, 32, 32, 16)])
def test_dgrad_upsample_folded_with_relu_mask(B, Cin, Cout, HW):
    """Gradient w.r.t. the pre-upsample ReLU output: transposed phase conv + (act > 0) mask."""
    L = _native()
    act = rb(torch.relu(_rand((B, Cin, HW, HW), 16)))
    dy, Wt = rb(_rand((B, Cout, 2 * HW, 2 * HW), 17)), _rand((Cout, Cin, 5, 5), 18, 0.05)
    Wg = packref.gemm_weights_phase_dgrad(Wt)                       # [Cin][9][4Cout]
    # reference from the same rounded effective weights
    Weff = rb(packref.phase_weights(Wt)).reshape(2, 2, Cout, 3, 3, Cin)
    xs = act.double().clone().requires_grad_(True)
    y = torch.zeros(B, Cout, 2 * HW, 2 * HW, dtype=torch.float64)
    ys = []
    for a in (0, 1):
        for bb in (0, 1):
            ys.append((a, bb, F.conv2d(xs, Weff[a, bb].permute(0, 3, 1, 2).double(), padding=1)))
    loss = sum((yy * dy.double()[:, :, a::2, bb::2]).sum() for a, bb, yy in ys)
    loss.backward()
    ref = (xs.grad * (act > 0)).float()
    out = torch.zeros(B, HW, HW, Cin, dtype=torch.bfloat16, device="cuda")
    run_conv(L, B=B, H=HW, W=HW, ksize=3, src_channels=4 * Cout, n_total=Cin, loader=L.LOAD_S2D,
             epilogue=L.EPI_MASK, ktab=L.KTAB_GENERIC, src=nhwc_bf16(dy),
             wpack=packref.pack_generic(rb(Wg)).cuda(), out=out, act=nhwc_bf16(act))
    np.testing.assert_allclose(from_nhwc(out).numpy(), rb(ref).numpy(), rtol=2 ** -7, atol=2e-3 * ref.abs().max().item())


def test_dgrad_last_conv_from_nchw_grad():
    """D4 data-gradient: source is d_recon (fp32 NCHW) * (1 - recon^2), 12 phase channels."""
    L = _native()
    B = 2
    act = rb(torch.relu(_rand((B, 32, 32, 32), 19)))
    g, recon, Wt = _rand((B, 3, 64, 64), 20), torch.tanh(_rand((B, 3, 64, 64), 21)), _rand((3, 32, 5, 5), 22, 0.05)
    dy = rb(g * (1 - recon * recon))
    Wg = packref.gemm_weights_phase_dgrad(Wt, k_pad=16)             # [32][9][16]
    Weff = rb(packref.phase_weights(Wt)).reshape(2, 2, 3, 3, 3, 32)
    xs = act.double().clone().requires_grad_(True)
    loss = 0
    for a in (0, 1):
        for bb in (0, 1):
            loss = loss + (F.conv2d(xs, Weff[a, bb].permute(0, 3, 1, 2).double(), padding=1) * dy.double()[:, :, a::2, bb::2]).sum()
    loss.backward()
    ref = (xs.grad * (act > 0)).float()
    out = torch.zeros(B, 32, 32, 32, dtype=torch.bfloat16, device="cuda")
    run_conv(L, B=B, H=32, W=32, ksize=3, src_channels=16, n_total=32, loader=L.LOAD_S2D_NCHW3_DTANH,
             epilogue=L.EPI_MASK, ktab=L.KTAB_GENERIC, src=g.cuda(), src2=recon.cuda(),
             wpack=packref.pack_generic(rb(Wg)).cuda(), out=out, act=nhwc_bf16(act))
    np.testing.assert_allclose(from_nhwc(out).numpy(), rb(ref).numpy(), rtol=2 ** -7, atol=2e-3 * ref.abs().max().item())


# ------------------------------------------------------------------------------------------------------------
# weights-as-A kernel (csrc/conv_wa.cu, CVAE_KTAB_BLOCK64): same operations, block-major packed weights.
# Every case runs with automatic tiling and with forced cluster sizes / grid sizes / tile counts, including
# ragged ones (a grid that does not divide the rows, more tiles than rows).
# ------------------------------------------------------------------------------------------------------------
# (cluster, grid, units per stage, tiles per CTA, K split, weight fetch)
WA_TUNES = [(0, 0, 0, 0, 0, 0), (1, 0, 1, 0, 1, 2), (2, 0, 2, 2, 1, 0), (4, 0, 0, 0, 1, 0), (1, 3, 0, 3, 1, 1), (2, 6, 1, 0, 1, 0)]
WA_SPLITS = [(0, 0, 0, 0, 2, 0), (0, 0, 1, 2, 2, 2), (0, 0, 0, 0, 4, 0), (0, 8, 0, 0, 4, 1)]     # split-K variants


@pytest.fixture
def wa_tune():
    L = _native()
    yield lambda t: L.lib.cvae_conv_wa_tune(*t)
    L.lib.cvae_conv_wa_tune(0, 0, 0, 0, 0, 0)


@pytest.mark.parametrize("tune", WA_TUNES)
@pytest.mark.parametrize("B,Cin,Cout,HW", [(5, 64, 128, 16), (7, 128, 256, 8), (50, 64, 128, 16)])
def test_wa_encoder_conv_with_stats(B, Cin, Cout, HW, tune, wa_tune):
    L = _native()
    wa_tune(tune)
    x, Wt = rb(_rand((B, Cin, HW, HW), 1)), rb(_rand((Cout, Cin, 5, 5), 2, 0.05))
    ref = F.conv2d(x.double(), Wt.double(), padding=2).float()
    wp = packref.pack_block64(packref.gemm_weights_fwd5(Wt)).cuda()
    out = torch.zeros(B, HW, HW, Cout, dtype=torch.bfloat16, device="cuda")
    stats = torch.zeros(2, Cout, dtype=torch.float64, device="cuda")
    run_conv(L, B=B, H=HW, W=HW, ksize=5, src_channels=Cin, n_total=Cout, loader=L.LOAD_NHWC,
             epilogue=L.EPI_STATS, ktab=L.KTAB_BLOCK64, src=nhwc_bf16(x), wpack=wp, out=out, stats=stats)
    got = from_nhwc(out)
    np.testing.assert_allclose(got.numpy(), rb(ref).numpy(), rtol=2 ** -7, atol=1e-3)
    g64 = got.double()
    np.testing.assert_allclose(stats[0].cpu().numpy(), g64.sum((0, 2, 3)).numpy(), rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(stats[1].cpu().numpy(), (g64 * g64).sum((0, 2, 3)).numpy(), rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("tune", WA_TUNES + WA_SPLITS)
def test_wa_decoder_conv0_bias_relu(tune, wa_tune):
    L = _native()
    wa_tune(tune)
    B = 6
    x, Wt, b = rb(_rand((B, 256, 4, 4), 5)), rb(_rand((128, 256, 5, 5), 6, 0.02)), _rand((128,), 7, 0.1)
    ref = torch.relu(F.conv2d(x.double(), Wt.double(), b.double(), padding=2)).float()
    out = torch.zeros(B, 4, 4, 128, dtype=torch.bfloat16, device="cuda")
    run_conv(L, B=B, H=4, W=4, ksize=5, src_channels=256, n_total=128, loader=L.LOAD_NHWC,
             epilogue=L.EPI_BIAS_RELU, ktab=L.KTAB_BLOCK64, src=nhwc_bf16(x),
             wpack=packref.pack_block64(packref.gemm_weights_fwd5(Wt)).cuda(), out=out, bias=b.cuda(), repeat=2)
    np.testing.assert_allclose(from_nhwc(out).numpy(), rb(ref).numpy(), rtol=2 ** -7, atol=2e-3)


@pytest.mark.parametrize("tune", WA_TUNES[:4] + WA_SPLITS[:2])
@pytest.mark.parametrize("B,Cin,Cout,HW", [(3, 128, 64, 4), (3, 64, 32, 8), (33, 128, 64, 4), (3, 32, 32, 16)])
def test_wa_decoder_upsample_folded_conv(B, Cin, Cout, HW, tune, wa_tune):
    L = _native()
    wa_tune(tune)
    x, Wt, b = rb(_rand((B, Cin, HW, HW), 8)), _rand((Cout, Cin, 5, 5), 9, 0.05), _rand((Cout,), 10, 0.1)
    Wg = packref.gemm_weights_phase_fwd(Wt)
    Weff = rb(Wg).reshape(2, 2, Cout, 3, 3, Cin)
    ref = torch.zeros(B, Cout, 2 * HW, 2 * HW, dtype=torch.float64)
    for a in (0, 1):
        for bb in (0, 1):
            ref[:, :, a::2, bb::2] = F.conv2d(x.double(), Weff[a, bb].permute(0, 3, 1, 2).double(), b.double(), padding=1)
    ref = torch.relu(ref).float()
    out = torch.zeros(B, 2 * HW, 2 * HW, Cout, dtype=torch.bfloat16, device="cuda")
    run_conv(L, B=B, H=HW, W=HW, ksize=3, src_channels=Cin, n_total=4 * Cout, loader=L.LOAD_NHWC,
             epilogue=L.EPI_PHASE_BIAS_RELU, ktab=L.KTAB_BLOCK64, src=nhwc_bf16(x),
             wpack=packref.pack_wa(Wg, 64 if Cin % 64 == 0 else 32, 1).cuda(), out=out, bias=b.cuda())
    np.testing.assert_allclose(from_nhwc(out).numpy(), rb(ref).numpy(), rtol=2 ** -7, atol=2e-3)


@pytest.mark.parametrize("tune", WA_TUNES[:4] + WA_SPLITS)
@pytest.mark.parametrize("B,Cin,Cout,HW", [(4, 128, 256, 8), (5, 256, 128, 4), (40, 128, 256, 8)])
def test_wa_dgrad_5x5(B, Cin, Cout, HW, tune, wa_tune):
    L = _native()
    wa_tune(tune)
    dy, Wt = rb(_rand((B, Cout, HW, HW), 14)), rb(_rand((Cout, Cin, 5, 5), 15, 0.05))
    ref = F.conv_transpose2d(dy.double(), Wt.double(), padding=2).float()
    out = torch.zeros(B, HW, HW, Cin, dtype=torch.bfloat16, device="cuda")
    run_conv(L, B=B, H=HW, W=HW, ksize=5, src_channels=Cout, n_total=Cin, loader=L.LOAD_NHWC,
             epilogue=L.EPI_PLAIN, ktab=L.KTAB_BLOCK64, src=nhwc_bf16(dy),
             wpack=packref.pack_block64(packref.gemm_weights_dgrad5(Wt)).cuda(), out=out, repeat=2)
    np.testing.assert_allclose(from_nhwc(out).numpy(), rb(ref).numpy(), rtol=2 ** -7, atol=2e-3 * ref.abs().max().item())


@pytest.mark.parametrize("tune", WA_TUNES[:4] + WA_SPLITS)
@pytest.mark.parametrize("B", [3, 21])
def test_wa_dgrad_upsample_folded_with_relu_mask(B, tune, wa_tune):
    """D1's data gradient: space-to-depth source with 4 x 64 channels (one TMA view per phase) + ReLU mask."""
    L = _native()
    wa_tune(tune)
    Cin, Cout, HW = 128, 64, 4
    act = rb(torch.relu(_rand((B, Cin, HW, HW), 16)))
    dy, Wt = rb(_rand((B, Cout, 2 * HW, 2 * HW), 17)), _rand((Cout, Cin, 5, 5), 18, 0.05)
    Wg = packref.gemm_weights_phase_dgrad(Wt)                       # [Cin][9][4Cout]
    Weff = rb(packref.phase_weights(Wt)).reshape(2, 2, Cout, 3, 3, Cin)
    xs = act.double().clone().requires_grad_(True)
    ys = []
    for a in (0, 1):
        for bb in (0, 1):
            ys.append((a, bb, F.conv2d(xs, Weff[a, bb].permute(0, 3, 1, 2).double(), padding=1)))
    loss = sum((yy * dy.double()[:, :, a::2, bb::2]).sum() for a, bb, yy in ys)
    loss.backward()
    ref = (xs.grad * (act > 0)).float()
    out = torch.zeros(B, HW, HW, Cin, dtype=torch.bfloat16, device="cuda")
    run_conv(L, B=B, H=HW, W=HW, ksize=3, src_channels=4 * Cout, n_total=Cin, loader=L.LOAD_S2D,
             epilogue=L.EPI_MASK, ktab=L.KTAB_BLOCK64, src=nhwc_bf16(dy),
             wpack=packref.pack_block64(rb(Wg)).cuda(), out=out, act=nhwc_bf16(act))
    np.testing.assert_allclose(from_nhwc(out).numpy(), rb(ref).numpy(), rtol=2 ** -7, atol=2e-3 * ref.abs().max().item())


WA_STACK_TUNES = [(0, 0, 0, 0, 0, 0), (1, 0, 1, 3, 1, 2), (1, 5, 3, 0, 1, 1)]


@pytest.mark.parametrize("tune", WA_STACK_TUNES)
@pytest.mark.parametrize("B,Cin,Cout,HW", [(3, 64, 128, 16), (2, 32, 64, 32), (19, 64, 128, 16), (6, 32, 64, 32)])
def test_wa_stacked_dgrad_5x5(B, Cin, Cout, HW, tune, wa_tune):
    """E2 / E1 data gradients: 64 / 32 GEMM rows, two / four horizontally adjacent taps stacked into the 128 MMA rows."""
    L = _native()
    wa_tune(tune)
    J = 128 // Cin
    dy, Wt = rb(_rand((B, Cout, HW, HW), 14)), rb(_rand((Cout, Cin, 5, 5), 15, 0.05))
    ref = F.conv_transpose2d(dy.double(), Wt.double(), padding=2).float()
    out = torch.zeros(B, HW, HW, Cin, dtype=torch.bfloat16, device="cuda")
    run_conv(L, B=B, H=HW, W=HW, ksize=5, src_channels=Cout, n_total=Cin, loader=L.LOAD_NHWC,
             epilogue=L.EPI_PLAIN, ktab=L.KTAB_BLOCK64, src=nhwc_bf16(dy), stack=J,
             wpack=packref.pack_wa(packref.gemm_weights_dgrad5(Wt), 64, J).cuda(), out=out)
    np.testing.assert_allclose(from_nhwc(out).numpy(), rb(ref).numpy(), rtol=2 ** -7, atol=2e-3 * ref.abs().max().item())


@pytest.mark.parametrize("tune", WA_STACK_TUNES)
@pytest.mark.parametrize("B", [3, 10])
def test_wa_stacked_encoder_conv1_with_stats(B, tune, wa_tune):
    """E1 forward: 32-channel source (64-byte rows, SWIZZLE_64B tiles), 64 output channels with two taps stacked."""
    L = _native()
    wa_tune(tune)
    Cin, Cout, HW = 32, 64, 32
    x, Wt = rb(_rand((B, Cin, HW, HW), 1)), rb(_rand((Cout, Cin, 5, 5), 2, 0.05))
    ref = F.conv2d(x.double(), Wt.double(), padding=2).float()
    out = torch.zeros(B, HW, HW, Cout, dtype=torch.bfloat16, device="cuda")
    stats = torch.zeros(2, Cout, dtype=torch.float64, device="cuda")
    run_conv(L, B=B, H=HW, W=HW, ksize=5, src_channels=Cin, n_total=Cout, loader=L.LOAD_NHWC,
             epilogue=L.EPI_STATS, ktab=L.KTAB_BLOCK64, src=nhwc_bf16(x), stack=2,
             wpack=packref.pack_wa(packref.gemm_weights_fwd5(Wt), 32, 2).cuda(), out=out, stats=stats)
    got = from_nhwc(out)
    np.testing.assert_allclose(got.numpy(), rb(ref).numpy(), rtol=2 ** -7, atol=1e-3)
    g64 = got.double()
    np.testing.assert_allclose(stats[0].cpu().numpy(), g64.sum((0, 2, 3)).numpy(), rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(stats[1].cpu().numpy(), (g64 * g64).sum((0, 2, 3)).numpy(), rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("tune", WA_STACK_TUNES)
@pytest.mark.parametrize("B,Cin,Cout,HW", [(2, 64, 32, 8), (2, 32, 32, 16), (11, 64, 32, 8), (7, 32, 32, 16)])
def test_wa_stacked_dgrad_upsample_folded_with_relu_mask(B, Cin, Cout, HW, tune, wa_tune):
    """D2 / D3 data gradients: space-to-depth source with 4 x 32 channels (64-byte rows), 64 / 32 GEMM rows stacked."""
    L = _native()
    wa_tune(tune)
    J = 128 // Cin
    act = rb(torch.relu(_rand((B, Cin, HW, HW), 16)))
    dy, Wt = rb(_rand((B, Cout, 2 * HW, 2 * HW), 17)), _rand((Cout, Cin, 5, 5), 18, 0.05)
    Wg = packref.gemm_weights_phase_dgrad(Wt)                       # [Cin][9][4Cout]
    Weff = rb(packref.phase_weights(Wt)).reshape(2, 2, Cout, 3, 3, Cin)
    xs = act.double().clone().requires_grad_(True)
    ys = []
    for a in (0, 1):
        for bb in (0, 1):
            ys.append((a, bb, F.conv2d(xs, Weff[a, bb].permute(0, 3, 1, 2).double(), padding=1)))
    loss = sum((yy * dy.double()[:, :, a::2, bb::2]).sum() for a, bb, yy in ys)
    loss.backward()
    ref = (xs.grad * (act > 0)).float()
    out = torch.zeros(B, HW, HW, Cin, dtype=torch.bfloat16, device="cuda")
    run_conv(L, B=B, H=HW, W=HW, ksize=3, src_channels=4 * Cout, n_total=Cin, loader=L.LOAD_S2D,
             epilogue=L.EPI_MASK, ktab=L.KTAB_BLOCK64, src=nhwc_bf16(dy), stack=J,
             wpack=packref.pack_wa(rb(Wg), 32, J).cuda(), out=out, act=nhwc_bf16(act))
    np.testing.assert_allclose(from_nhwc(out).numpy(), rb(ref).numpy(), rtol=2 ** -7, atol=2e-3 * ref.abs().max().item())


def test_wa_device_packing_matches_host_packing():
    """cvae_pack_weights with CVAE_PACK_KORDER_BLOCK64 == tests/packref.pack_block64, bit for bit."""
    L = _native()
    W1 = _rand((256, 128, 5, 5), 30, 0.05)       # E3: forward n = 256, data gradient n = 128
    W2 = _rand((64, 128, 5, 5), 31, 0.05)        # D1: phase forward n = 256, phase data gradient n = 128
    cases = [(L.PACK_FWD5, W1, 256, 25 * 128 // 16, 128, packref.pack_block64(packref.gemm_weights_fwd5(W1))),
             (L.PACK_DGRAD5, W1, 128, 25 * 256 // 16, 256, packref.pack_block64(packref.gemm_weights_dgrad5(W1))),
             (L.PACK_PHASE_FWD, W2, 256, 9 * 128 // 16, 128, packref.pack_block64(packref.gemm_weights_phase_fwd(W2))),
             (L.PACK_PHASE_DGRAD, W2, 128, 9 * 256 // 16, 256, packref.pack_block64(packref.gemm_weights_phase_dgrad(W2)))]
    cases = [(k | L.PACK_KORDER_BLOCK64, W, n, ks, kch, ref) for k, W, n, ks, kch, ref in cases]
    W3 = _rand((64, 32, 5, 5), 32, 0.05)         # E1: forward (32-channel blocks, two taps stacked), data gradient (four taps stacked)
    W4 = _rand((32, 64, 5, 5), 33, 0.05)         # D2: phase data gradient n = 64 (two taps stacked), 4 x 32 source channels
    cases += [(L.PACK_FWD5 | L.PACK_KORDER_BLOCK32 | L.PACK_STACK2, W3, 128, 1 * 15 * 2, 32, packref.pack_wa(packref.gemm_weights_fwd5(W3), 32, 2)),
              (L.PACK_DGRAD5 | L.PACK_KORDER_BLOCK64 | L.PACK_STACK4, W3, 128, 1 * 10 * 4, 64, packref.pack_wa(packref.gemm_weights_dgrad5(W3), 64, 4)),
              (L.PACK_PHASE_DGRAD | L.PACK_KORDER_BLOCK32 | L.PACK_STACK2, W4, 128, 4 * 6 * 2, 128,
               packref.pack_wa(packref.gemm_weights_phase_dgrad(W4), 32, 2))]
    for kind, W, n, ksteps, kch, ref in cases:
        dst = torch.zeros(n * ksteps * 16, dtype=torch.bfloat16, device="cuda")
        Wd = W.cuda()
        job = L.PackJob(kind=kind, n=n, ksteps=ksteps, k_channels=kch, cout=W.shape[0], cin=W.shape[1],
                        src=Wd.data_ptr(), src2=None, dst=dst.data_ptr())
        L.check(L.lib.cvae_pack_weights((L.PackJob * 1)(job), 1, L.stream_ptr()))
        torch.cuda.synchronize()
        assert torch.equal(dst.cpu().view(torch.int16), ref.view(torch.int16)), kind
