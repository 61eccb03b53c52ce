"""Out-of-bounds write detection without compute-sanitizer (the tool is closed on this GPU pool): every output tensor of
the launch is a window inside a larger buffer filled with a canary pattern; after the kernel the 64 KiB on either side
of the window must be untouched.  Covers the kernels whose store addresses are computed from tile geometry (both
convolution kernels in all their epilogue forms, both weight-gradient kernels and their folds, the latent kernels, the
uint8 front end)."""
import ctypes

import numpy as np
import pytest
import torch

import packref
from test_conv_gemm import _native, _rand, nhwc_bf16, rb

pytestmark = pytest.mark.gpu

GUARD = 64 * 1024
CANARY = 0x5A


class Guarded:
    def __init__(self, shape, dtype):
        n = int(np.prod(shape)) * torch.empty(0, dtype=dtype).element_size()
        self.raw = torch.full((GUARD + n + GUARD,), CANARY, dtype=torch.uint8, device="cuda")
        self.t = self.raw[GUARD:GUARD + n].view(dtype).view(*shape)
        self.t.zero_()

    def check(self, what):
        torch.cuda.synchronize()
        lo, hi = self.raw[:GUARD], self.raw[-GUARD:]
        assert bool((lo == CANARY).all()) and bool((hi == CANARY).all()), f"{what}: wrote outside its output tensor"


def _conv(L, **kw):
    ws = None
    d = L.ConvDesc(**kw)
    need = int(L.lib.cvae_conv_gemm_workspace_bytes(ctypes.byref(d)))
    if need > 0:
        ws = Guarded((need,), torch.uint8)
        d.workspace, d.workspace_bytes = ws.t.data_ptr(), need
    L.check(L.lib.cvae_conv_gemm(ctypes.byref(d), L.stream_ptr()))
    torch.cuda.synchronize()
    L.check(L.lib.cvae_check_device_fault(L.stream_ptr()))
    if ws is not None:
        ws.check("split-K workspace")


@pytest.mark.parametrize("B", [1, 3, 50])
def test_conv_kernels_stay_inside_their_outputs(B):
    L = _native()
    cases = [  # name, H, ksize, C, N, loader, epilogue, weights-as-A (kb, J) or None
        ("E2f", 16, 5, 64, 128, L.LOAD_NHWC, L.EPI_STATS, (64, 1)), ("E3g", 8, 5, 256, 128, L.LOAD_NHWC, L.EPI_PLAIN, (64, 1)),
        ("E2g", 16, 5, 128, 64, L.LOAD_NHWC, L.EPI_PLAIN, (64, 2)), ("E1g", 32, 5, 64, 32, L.LOAD_NHWC, L.EPI_PLAIN, (64, 4)),
        ("D0f", 4, 5, 256, 128, L.LOAD_NHWC, L.EPI_BIAS_RELU, (64, 1)), ("D2f", 8, 3, 64, 128, L.LOAD_NHWC, L.EPI_PHASE_BIAS_RELU, (64, 1)),
        ("D2g", 8, 3, 128, 64, L.LOAD_S2D, L.EPI_MASK, (32, 2)), ("D1g", 4, 3, 256, 128, L.LOAD_S2D, L.EPI_MASK, (64, 1)),
        ("E1f", 32, 5, 32, 64, L.LOAD_NHWC, L.EPI_STATS, None), ("D3f", 16, 3, 32, 128, L.LOAD_NHWC, L.EPI_PHASE_BIAS_RELU, None),
        ("D3g", 16, 3, 128, 32, L.LOAD_S2D, L.EPI_MASK, None), ("D4f", 32, 3, 32, 16, L.LOAD_NHWC, L.EPI_PHASE_BIAS_TANH, None),
        ("E0f", 64, 5, 8, 32, L.LOAD_NCHW3, L.EPI_STATS, None)]
    for ksplit in (0, 4):
        L.lib.cvae_conv_wa_tune(0, 0, 0, 0, ksplit, 0)
        for name, H, k, C, N, loader, epi, wa in cases:
            if ksplit and (wa is None or wa[1] != 1):
                continue
            bf = torch.bfloat16
            if loader == L.LOAD_NCHW3:
                src = torch.rand(B, 3, H, H, device="cuda")
            elif loader == L.LOAD_S2D:
                src = torch.randn(B, 2 * H, 2 * H, C // 4, device="cuda").to(bf)
            else:
                src = torch.randn(B, H, H, C, device="cuda").to(bf)
            J = wa[1] if wa else 1
            ktab = L.KTAB_PAIR8 if loader == L.LOAD_NCHW3 else (L.KTAB_BLOCK64 if wa else L.KTAB_GENERIC)
            if wa:
                groups = k * k if J == 1 else {(5, 2): 15, (5, 4): 10, (3, 2): 6, (3, 4): 3}[(k, J)]
                ksteps = (C // wa[0]) * groups * (wa[0] // 16) * J
            else:
                ksteps = L.lib.cvae_conv_ksteps(k, C, ktab)
            wp = (torch.randn(N * ksteps * 16, device="cuda") * 0.05).to(bf)
            if epi == L.EPI_PHASE_BIAS_TANH:
                out = Guarded((B, 3, 2 * H, 2 * H), torch.float32)
            elif epi == L.EPI_PHASE_BIAS_RELU:
                out = Guarded((B, 2 * H, 2 * H, N // 4), bf)
            else:
                out = Guarded((B, H, H, N), bf)
            stats = Guarded((2, N), torch.float64)
            bias = torch.zeros(256, device="cuda")
            act = torch.randn(B, H, H, N, device="cuda").to(bf)
            _conv(L, batch=B, height=H, width=H, ksize=k, src_channels=C, n_total=N, loader=loader, epilogue=epi, ktab=ktab, stack=J if wa else 0,
                  src=src.data_ptr(), src2=None, wpack=wp.data_ptr(), bias=bias.data_ptr(), act=act.data_ptr(), out=out.t.data_ptr(),
                  stats=stats.t.data_ptr() if epi == L.EPI_STATS else None)
            out.check(f"{name} B={B} ksplit={ksplit} output")
            stats.check(f"{name} statistics")
    L.lib.cvae_conv_wa_tune(0, 0, 0, 0, 0, 0)


@pytest.mark.parametrize("B", [1, 5, 40])
def test_wgrad_kernels_stay_inside_their_outputs(B):
    L = _native()
    bf = torch.bfloat16
    for kind, h, co, ci in ((L.WGRAD_5X5, 4, 128, 256), (L.WGRAD_5X5, 8, 256, 128), (L.WGRAD_5X5, 32, 64, 32), (L.WGRAD_PHASE, 4, 64, 128),
                            (L.WGRAD_PHASE, 16, 32, 32), (L.WGRAD_SHIFT_FRAMES, 64, 32, 3), (L.WGRAD_SHIFT_PHASE12, 32, 3, 32)):
        if kind == L.WGRAD_SHIFT_FRAMES:
            x, dy, dy2 = torch.rand(B, 3, h, h, device="cuda"), torch.randn(B, h, h, co, device="cuda").to(bf), None
        elif kind == L.WGRAD_SHIFT_PHASE12:
            x = torch.randn(B, h, h, ci, device="cuda").to(bf)
            dy, dy2 = torch.randn(B, 3, 2 * h, 2 * h, device="cuda"), torch.rand(B, 3, 2 * h, 2 * h, device="cuda")
        elif kind == L.WGRAD_PHASE:
            x, dy, dy2 = torch.randn(B, h, h, ci, device="cuda").to(bf), torch.randn(B, 2 * h, 2 * h, co, device="cuda").to(bf), None
        else:
            x, dy, dy2 = torch.randn(B, h, h, ci, device="cuda").to(bf), torch.randn(B, h, h, co, device="cuda").to(bf), None
        dw, db = Guarded((co, ci, 5, 5), torch.float32), Guarded((co,), torch.float32)
        d = L.WgradDesc(kind=kind, batch=B, height=h, width=h, cout=co, cin=ci, splits=0, x=x.data_ptr(), dy=dy.data_ptr(),
                        dy2=dy2.data_ptr() if dy2 is not None else None, dw=dw.t.data_ptr(), dbias=db.t.data_ptr(), workspace=None, fold_stream=None)
        need = int(L.lib.cvae_conv_wgrad_workspace_bytes(ctypes.byref(d)))
        ws = Guarded((need,), torch.uint8)
        d.workspace = ws.t.data_ptr()
        L.check(L.lib.cvae_conv_wgrad(ctypes.byref(d), L.stream_ptr()))
        torch.cuda.synchronize()
        L.check(L.lib.cvae_check_device_fault(L.stream_ptr()))
        for g, what in ((dw, "dw"), (db, "dbias"), (ws, "split-K workspace")):
            g.check(f"wgrad kind {kind} {h}x{h} {ci}->{co} B={B} {what}")


@pytest.mark.parametrize("B", [1, 63, 64, 65, 257])
def test_latent_and_frame_kernels_stay_inside_their_outputs(B):
    L = _native()
    ml, eps, pred = torch.randn(B, 64, device="cuda"), torch.randn(B, 32, device="cuda"), torch.rand(B, device="cuda")
    zc, parts = Guarded((B, 33), torch.float32), Guarded((L.lib.cvae_latent_kld_partials(B),), torch.float64)
    L.check(L.lib.cvae_latent_fwd(B, 1, ml.data_ptr(), eps.data_ptr(), pred.data_ptr(), zc.t.data_ptr(), parts.t.data_ptr(), L.stream_ptr()))
    dml = Guarded((B, 64), torch.float32)
    L.check(L.lib.cvae_latent_bwd(B, ml.data_ptr(), eps.data_ptr(), zc.t.data_ptr(), None, None, 0.001 / B, dml.t.data_ptr(), L.stream_ptr()))
    u8 = torch.randint(0, 256, (B, 64, 64, 3), dtype=torch.uint8, device="cuda")
    x = Guarded((B, 3, 64, 64), torch.float32)
    L.check(L.lib.cvae_frames_u8_to_f32(B, u8.data_ptr(), x.t.data_ptr(), L.stream_ptr()))
    for g, what in ((zc, "z|pred"), (parts, "KL partials"), (dml, "d mu|logvar"), (x, "frames")):
        g.check(f"B={B} {what}")
