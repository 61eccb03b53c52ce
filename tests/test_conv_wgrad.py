"""tcgen05 weight-gradient GEMM (csrc/conv_wgrad.cu) against torch CPU autograd on bf16-rounded operands."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from test_conv_gemm import _native, _rand, nhwc_bf16, rb

pytestmark = pytest.mark.gpu


def run_wgrad(L, kind, B, H, W, cout, cin, x, dy, dy2=None, splits=0, bias=True):
    d = L.WgradDesc(kind=kind, batch=B, height=H, width=W, cout=cout, cin=cin, splits=splits,
                    x=x.data_ptr(), dy=dy.data_ptr(), dy2=dy2.data_ptr() if dy2 is not None else None)
    dw = torch.full((cout, cin, 5, 5), float("nan"), device="cuda")
    db = torch.full((cout,), float("nan"), device="cuda")
    d.dw, d.dbias = dw.data_ptr(), (db.data_ptr() if bias else None)
    # the split-K partials live in the workspace: poison it (a partial the fold reads but no CTA wrote shows up as NaN) and tell
    # the library its size (a call that would write past it fails instead)
    need = int(L.lib.cvae_conv_wgrad_workspace_bytes(ctypes.byref(d)))
    ws = torch.full((need // 4,), float("nan"), device="cuda")
    d.workspace, d.workspace_bytes = ws.data_ptr(), need
    L.check(L.lib.cvae_conv_wgrad(ctypes.byref(d), L.stream_ptr()))
    torch.cuda.synchronize()
    L.check(L.lib.cvae_check_device_fault(L.stream_ptr()))
    return dw.cpu(), db.cpu()


def _check(dw, db, ref_w, ref_b):
    sw, sb = ref_w.abs().max().item(), ref_b.abs().max().item()
    np.testing.assert_allclose(dw.numpy(), ref_w.numpy(), rtol=1e-4, atol=2e-5 * sw)
    np.testing.assert_allclose(db.numpy(), ref_b.numpy(), rtol=1e-4, atol=2e-5 * sb)


@pytest.mark.parametrize("B,Cin,Cout,HW,splits", [(3, 32, 64, 32, 0), (4, 64, 128, 16, 0), (5, 128, 256, 8, 0),
                                                 (6, 256, 128, 4, 0), (3, 32, 64, 32, 1), (7, 64, 128, 16, 3),
                                                 # TMA-fed variant: several chunks per CTA, image boxes running past the batch
                                                 (41, 128, 256, 8, 0), (33, 64, 128, 16, 0), (50, 256, 128, 4, 0),
                                                 # tap-stacked variant (32 -> 64 channels): many chunks per CTA, other map sizes, forced splits
                                                 (37, 32, 64, 32, 0), (5, 32, 64, 16, 0), (9, 32, 64, 8, 7), (150, 32, 64, 4, 0)])
def test_wgrad_5x5(B, Cin, Cout, HW, splits):
    L = _native()
    x, dy = rb(_rand((B, Cin, HW, HW), 31)), rb(_rand((B, Cout, HW, HW), 32))
    ref_w = torch.nn.grad.conv2d_weight(x.double(), (Cout, Cin, 5, 5), dy.double(), padding=2).float()
    dw, db = run_wgrad(L, L.WGRAD_5X5, B, HW, HW, Cout, Cin, nhwc_bf16(x), nhwc_bf16(dy), splits=splits)
    _check(dw, db, ref_w, dy.double().sum((0, 2, 3)).float())


@pytest.mark.parametrize("B,Cin,Cout,HW", [(3, 128, 64, 4), (3, 64, 32, 8), (2, 32, 32, 16), (45, 128, 64, 4),
                                           (21, 32, 32, 16), (5, 32, 32, 8), (40, 32, 32, 32)])      # tap-stacked variant (32 -> 4 x 32)
def test_wgrad_upsample_folded(B, Cin, Cout, HW):
    L = _native()
    x, dy = rb(_rand((B, Cin, HW, HW), 33)), rb(_rand((B, Cout, 2 * HW, 2 * HW), 34))
    Wt = torch.zeros(Cout, Cin, 5, 5, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(F.interpolate(x.double(), scale_factor=2, mode="nearest"), Wt, padding=2)
    (y * dy.double()).sum().backward()
    dw, db = run_wgrad(L, L.WGRAD_PHASE, B, HW, HW, Cout, Cin, nhwc_bf16(x), nhwc_bf16(dy))
    _check(dw, db, Wt.grad.float(), dy.double().sum((0, 2, 3)).float())


@pytest.mark.parametrize("B", [1, 3])
def test_wgrad_encoder_conv0_frames(B):
    L = _native()
    x = torch.rand(B, 3, 64, 64, generator=torch.Generator().manual_seed(35))
    dy = rb(_rand((B, 32, 64, 64), 36))
    ref_w = torch.nn.grad.conv2d_weight(rb(x).double(), (32, 3, 5, 5), dy.double(), padding=2).float()
    dw, db = run_wgrad(L, L.WGRAD_SHIFT_FRAMES, B, 64, 64, 32, 3, x.cuda(), nhwc_bf16(dy))
    _check(dw, db, ref_w, dy.double().sum((0, 2, 3)).float())


@pytest.mark.parametrize("B,splits", [(1, 0), (3, 0), (19, 0), (40, 5), (5, 1)])
def test_wgrad_encoder_conv0_frames_without_bias(B, splits):
    """No bias gradient (the conv sits in front of BatchNorm: what the training step asks for) takes the operand-stacked TMA
    variant (launch_wgrad_frames): dY rows as A with four row-shifted blocks, the frames as one plane filled by the epilogue
    warps; several chunks per CTA, chunks that run past the image and the batch, forced split counts."""
    L = _native()
    x = torch.rand(B, 3, 64, 64, generator=torch.Generator().manual_seed(135))
    dy = rb(_rand((B, 32, 64, 64), 136))
    ref_w = torch.nn.grad.conv2d_weight(rb(x).double(), (32, 3, 5, 5), dy.double(), padding=2).float()
    dw, _ = run_wgrad(L, L.WGRAD_SHIFT_FRAMES, B, 64, 64, 32, 3, x.cuda(), nhwc_bf16(dy), splits=splits, bias=False)
    sw = ref_w.abs().max().item()
    np.testing.assert_allclose(dw.numpy(), ref_w.numpy(), rtol=1e-4, atol=2e-5 * sw)


@pytest.mark.parametrize("B", [1, 3])
def test_wgrad_decoder_last_conv(B):
    L = _native()
    x = rb(torch.relu(_rand((B, 32, 32, 32), 37)))
    g, recon = _rand((B, 3, 64, 64), 38), torch.tanh(_rand((B, 3, 64, 64), 39))
    dy = rb(g * (1 - recon * recon))
    Wt = torch.zeros(3, 32, 5, 5, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(F.interpolate(x.double(), scale_factor=2, mode="nearest"), Wt, padding=2)
    (y * dy.double()).sum().backward()
    dw, db = run_wgrad(L, L.WGRAD_SHIFT_PHASE12, B, 32, 32, 3, 32, nhwc_bf16(x), g.cuda(), dy2=recon.cuda())
    _check(dw, db, Wt.grad.float(), dy.double().sum((0, 2, 3)).float())


def test_wgrad_refuses_a_workspace_that_is_too_small():
    """The engine once sized the workspace before it filled in dbias (the bias pseudo-group makes the partials larger): the
    descriptor now carries the size and the call fails instead of writing past the end; the query covers both forms."""
    L = _native()
    B, Cin, Cout, HW = 5, 128, 256, 8
    x, dy = nhwc_bf16(_rand((B, Cin, HW, HW), 1)), nhwc_bf16(_rand((B, Cout, HW, HW), 2))
    dw, db = torch.zeros(Cout, Cin, 5, 5, device="cuda"), torch.zeros(Cout, device="cuda")
    d = L.WgradDesc(kind=L.WGRAD_5X5, batch=B, height=HW, width=HW, cout=Cout, cin=Cin, splits=0, x=x.data_ptr(), dy=dy.data_ptr(), dw=dw.data_ptr())
    without_bias = int(L.lib.cvae_conv_wgrad_workspace_bytes(ctypes.byref(d)))
    d.dbias = db.data_ptr()
    assert int(L.lib.cvae_conv_wgrad_workspace_bytes(ctypes.byref(d))) == without_bias      # one answer for both forms
    ws = torch.empty(without_bias, dtype=torch.uint8, device="cuda")
    d.workspace, d.workspace_bytes = ws.data_ptr(), 25 * Cout * Cin * 4      # one split without the bias block: too small
    assert L.lib.cvae_conv_wgrad(ctypes.byref(d), L.stream_ptr()) == -1 and b"workspace" in L.lib.cvae_last_error()
    d.workspace_bytes = without_bias
    L.check(L.lib.cvae_conv_wgrad(ctypes.byref(d), L.stream_ptr()))
    torch.cuda.synchronize()
