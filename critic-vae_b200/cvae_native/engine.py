"""Host-side step driver: owns the flat parameter / gradient / optimizer buffers and the per-batch
activation workspace, and sequences the libcvae.so kernels for encode, decode, loss, backward and
Adam.  torch is used for device memory, streams and (optionally) CUDA-graph capture only; every
number is produced by the sm_100a kernels behind include/cvae.h.

Layer table (reference lines in vae_nets.py):
  encoder  E0..E3  Conv2d 5x5 -> BatchNorm2d -> MaxPool2d(2) -> ReLU/Tanh            :68-88
  heads    fc_mu, fc_var Linear(4096, 32)                                            :98-99
  decoder  decoder_input Linear(33, 4096); D0 conv; D1..D4 conv on 2x up-sampled map :116-137
"""
from __future__ import annotations

import ctypes
import math
import os

import torch

from . import binding as L

ENC = [(3, 32, 64), (32, 64, 32), (64, 128, 16), (128, 256, 8)]          # (cin, cout, H=W of the conv)
DEC = [(256, 128, 4), (128, 64, 4), (64, 32, 8), (32, 32, 16), (32, 3, 32)]  # (cin, cout, H=W of the conv INPUT grid)
ENC_CONV_IDX, ENC_BN_IDX, DEC_CONV_IDX = (0, 4, 8, 12), (1, 5, 9, 13), (0, 3, 6, 9, 12)
KLD_WEIGHT = 0.001          # vae_parameters.py:17
BN_MOMENTUM, BN_EPS = 0.1, 1e-5


def param_layout():
    """(name, shape) in the order of reference `VariationalAutoencoder.parameters()`."""
    out = []
    for i, (ci, co, _) in enumerate(ENC):
        out += [(f"encoder.model.{ENC_CONV_IDX[i]}.weight", (co, ci, 5, 5)), (f"encoder.model.{ENC_CONV_IDX[i]}.bias", (co,)),
                (f"encoder.model.{ENC_BN_IDX[i]}.weight", (co,)), (f"encoder.model.{ENC_BN_IDX[i]}.bias", (co,))]
    out += [("encoder.fc_mu.weight", (32, 4096)), ("encoder.fc_mu.bias", (32,)),
            ("encoder.fc_var.weight", (32, 4096)), ("encoder.fc_var.bias", (32,))]
    for i, (ci, co, _) in enumerate(DEC):
        out += [(f"decoder.model.{DEC_CONV_IDX[i]}.weight", (co, ci, 5, 5)), (f"decoder.model.{DEC_CONV_IDX[i]}.bias", (co,))]
    out += [("decoder.decoder_input.weight", (4096, 33)), ("decoder.decoder_input.bias", (4096,))]
    return out


def msssim_window():
    """The 11 normalised window weights exactly as the reference builds them (vae_nets.py:170-173):
    python floats -> torch.tensor (fp32) -> divide by the fp32 sum.  Note the upstream sign."""
    k = torch.tensor([math.exp((i - 11 // 2) ** 2 / (2 * 1.5 ** 2)) for i in range(11)])
    k = k / k.sum()
    return (ctypes.c_float * 11)(*[float(v) for v in k])


def _ptr(t):
    return None if t is None else t.data_ptr()


_NVTX = os.environ.get("CVAE_NVTX") == "1"


def _nvtx(name):
    """NVTX range around a phase of the step (encode / decode / loss / backward / adam) for nsys / ncu timelines
    (SURVEY.md section 5).  Off unless CVAE_NVTX=1: the push/pop pair costs host time on the eager path."""
    def deco(fn):
        if not _NVTX:
            return fn

        def wrapped(*a, **kw):
            torch.cuda.nvtx.range_push("cvae." + name)
            try:
                return fn(*a, **kw)
            finally:
                torch.cuda.nvtx.range_pop()
        wrapped.__doc__ = fn.__doc__
        return wrapped
    return deco


class Workspace:
    """Activation / gradient buffers for one batch size."""

    def __init__(self, B, dev, with_grad):
        bf, f32, f64 = torch.bfloat16, torch.float32, torch.float64
        e = lambda *s, dtype=bf: torch.empty(*s, dtype=dtype, device=dev)
        self.B = B
        self.c = [e(B, h, h, co) for (_, co, h) in ENC]                 # bias-free conv outputs
        self.a = [e(B, h // 2, h // 2, co) for (_, co, h) in ENC]       # pooled + activated
        if with_grad:   # saved by the BatchNorm/pool forward for its backward
            self.xh = [e(B, h // 2, h // 2, co) for (_, co, h) in ENC]
            self.am = [e(B, h // 2, h // 2, co // 8, dtype=torch.int16) for (_, co, h) in ENC]
        self.ml = e(B, 64, dtype=f32)
        self.zc = e(B, 33, dtype=f32)
        self.kld_partial = torch.zeros((B + 63) // 64, dtype=f64, device=dev)   # per-64-row KL sums from the latent kernel
        self.h0 = e(B, 4, 4, 256)
        self.d = [e(B, 4, 4, 128), e(B, 8, 8, 64), e(B, 16, 16, 32), e(B, 32, 32, 32)]
        self.recon = e(B, 3, 64, 64, dtype=f32)
        self.stats = torch.zeros(2 * sum(co for _, co, _ in ENC), dtype=f64, device=dev)
        self.ss = [e(4, co, dtype=f32) for (_, co, _) in ENC]
        self.loss_sums = torch.zeros(10, dtype=f64, device=dev)
        self.coef = torch.zeros(8, dtype=f32, device=dev)
        self.losses = torch.zeros(3, dtype=f32, device=dev)
        if with_grad:
            self.d_recon = e(B, 3, 64, 64, dtype=f32)
            self.d_mu, self.d_lv = e(B, 32, dtype=f32), e(B, 32, dtype=f32)
            self.g_d = [torch.empty_like(t) for t in self.d]            # grads w.r.t. pre-ReLU decoder conv outputs
            self.g_h0 = torch.empty_like(self.h0)
            self.dzc, self.dml = e(B, 33, dtype=f32), e(B, 64, dtype=f32)
            self.g_a = [torch.empty_like(t) for t in self.a]
            self.g_c = [torch.empty_like(t) for t in self.c]
            self.bn_sums = torch.zeros(2 * 256, dtype=f64, device=dev)


class VAEEngine:
    def __init__(self, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise L.CvaeError("the Critic-VAE hot path only runs on CUDA (sm_100a); there is no CPU fallback")
        self.layout = param_layout()
        self.offsets, off = {}, 0
        for name, shape in self.layout:
            n = math.prod(shape)
            self.offsets[name] = (off, n, shape)
            off += n
        self.n_params = off                                             # 2,583,971
        dev = self.device
        self.flat = torch.zeros(off, device=dev)
        self.gflat = torch.zeros(off, device=dev)
        self.gflat_alt = None
        self.exp_avg = self.exp_avg_sq = None
        self.step = torch.zeros(1, dtype=torch.int64, device=dev)
        # BatchNorm buffers
        self.running_mean = [torch.zeros(co, device=dev) for _, co, _ in ENC]
        self.running_var = [torch.ones(co, device=dev) for _, co, _ in ENC]
        self.nbt = [torch.zeros((), dtype=torch.int64, device=dev) for _ in ENC]
        self.window = msssim_window()
        self._build_pack_jobs()
        self._ws = {}
        self._wgrad_ws = None
        self.profile = None
        # Weight gradients are leaves of the backward pass: with a side stream they run concurrently with the
        # data-gradient chain (fork / join with events, so the pair is CUDA-graph capturable).
        self.side_stream = None
        self.fold_stream = None     # third stream: split-K folds beside the next weight-gradient GEMM
        self.early_event = None     # optional torch.cuda.Event(external=True) recorded when the early gradient bucket is final
        self.mid_event = None       # the same for the middle bucket (encoder convs 2 and 3): mid_bucket_offset() .. early_bucket_offset()
        # the three layers around the latent as one launch per direction in the training step (CVAE_NO_FUSED_BOTTLENECK=1: six launches)
        self.fused_bottleneck = os.environ.get("CVAE_NO_FUSED_BOTTLENECK") is None
        self.early_adam = None      # dict(lr=, grad_scale=): update all parameters but the first conv block inside backward()
        self.adam_stream = None
        self._adam_used = False
        self._loss_on_side = False

    # ---- parameters ---------------------------------------------------------------------------
    def view(self, name, buf=None):
        off, n, shape = self.offsets[name]
        return (self.flat if buf is None else buf)[off:off + n].view(shape)

    def _build_pack_jobs(self):
        dev = self.device
        jobs, self.packed = [], {}

        # Layers whose GEMM has >= 128 output columns and a source in multiples of 64 channels run the weights-as-A
        # kernel (csrc/conv_wa.cu, CVAE_KTAB_BLOCK64); it wants its K steps packed block-major.  CVAE_NO_WA=1 keeps
        # every layer on the pixels-as-M kernel (A/B comparisons).
        # key -> (channels per K block, taps stacked into the 128 MMA rows).  Everything but the 3-channel ends of the
        # network (E0 forward, D4 forward / data gradient) qualifies; the default set is the layers where the
        # weights-as-A kernel measured faster at batch 256 (profiles/r02_conv_layers.md).  CVAE_WA_ONLY=<keys> picks
        # another subset, CVAE_WA_ONLY=all takes every eligible layer (A/B comparisons).
        wa_all = {"E1f": (32, 2), "E2f": (64, 1), "E3f": (64, 1), "D0f": (64, 1), "D1f": (64, 1), "D2f": (64, 1), "D3f": (32, 1),
                  "E1g": (64, 4), "E2g": (64, 2), "E3g": (64, 1), "D0g": (64, 1), "D1g": (64, 1), "D2g": (32, 2), "D3g": (32, 4)}
        only = os.environ.get("CVAE_WA_ONLY", "E2f,E3f,E3g,E2g,D0f,D2f,D2g")
        wa = dict(wa_all) if only == "all" else {k: v for k, v in wa_all.items() if k in only.split(",")}
        self.wa_keys = {} if os.environ.get("CVAE_NO_WA") else wa
        self._conv_ws = {}

        def add(key, kind, n, ksteps, kch, cout, cin, src, src2=None, dtype=torch.bfloat16, elems=None):
            if key in self.wa_keys:
                kb, J = self.wa_keys[key]
                assert kch % kb == 0 and (n * J) % 128 == 0, key
                taps = 25 if kind in (L.PACK_FWD5, L.PACK_DGRAD5) else 9
                groups = (25 if taps == 25 else 9) if J == 1 else {(25, 2): 15, (25, 4): 10, (9, 2): 6, (9, 4): 3}[(taps, J)]
                ksteps = (kch // kb) * groups * (kb // 16)
                n = n * J
                kind |= (L.PACK_KORDER_BLOCK64 if kb == 64 else L.PACK_KORDER_BLOCK32) | {1: 0, 2: L.PACK_STACK2, 4: L.PACK_STACK4}[J]
            numel = elems if elems is not None else n * ksteps * 16
            dst = torch.zeros(numel, dtype=dtype, device=dev)
            self.packed[key] = dst
            jobs.append(L.PackJob(kind=kind, n=n, ksteps=ksteps, k_channels=kch, cout=cout, cin=cin,
                                  src=_ptr(src), src2=_ptr(src2), dst=_ptr(dst)))

        W = lambda n: self.view(n)
        e, d = "encoder.model.", "decoder.model."
        add("E0f", L.PACK_PAIR8, 32, 13, 8, 32, 3, W(e + "0.weight"))
        for i in (1, 2, 3):
            ci, co, _ = ENC[i]
            w = W(f"{e}{ENC_CONV_IDX[i]}.weight")
            add(f"E{i}f", L.PACK_FWD5, co, 25 * ci // 16, ci, co, ci, w)
            add(f"E{i}g", L.PACK_DGRAD5, ci, 25 * co // 16, co, co, ci, w)
        w = W(d + "0.weight")
        add("D0f", L.PACK_FWD5, 128, 25 * 16, 256, 128, 256, w)
        add("D0g", L.PACK_DGRAD5, 256, 25 * 8, 128, 128, 256, w)
        for i in (1, 2, 3, 4):
            ci, co, _ = DEC[i]
            w = W(f"{d}{DEC_CONV_IDX[i]}.weight")
            n4 = max(16, 4 * co)
            add(f"D{i}f", L.PACK_PHASE_FWD, n4, 9 * ci // 16, ci, co, ci, w)
            add(f"D{i}g", L.PACK_PHASE_DGRAD, ci, 9 * n4 // 16, n4, co, ci, w)
        add("fc", L.PACK_FC, 0, 0, 0, 0, 0, W("encoder.fc_mu.weight"), W("encoder.fc_var.weight"), torch.float32, 4096 * 64)
        add("decin", L.PACK_DECIN, 0, 0, 0, 0, 0, W("decoder.decoder_input.weight"), W("decoder.decoder_input.bias"),
            torch.float32, 34 * 4096)
        self._jobs = (L.PackJob * len(jobs))(*jobs)
        # data-gradient forms (keys ending in "g") are first needed by the backward pass
        keys = list(self.packed.keys())
        # with a side stream only encoder conv 0's 13 KB form is packed in front of the forward pass; the other forward
        # forms are packed beside conv 0 / its BatchNorm, the data-gradient forms beside the rest of the forward pass
        first = [j for j, k in zip(jobs, keys) if k == "E0f"]
        fwd = [j for j, k in zip(jobs, keys) if not k.endswith("g") and k != "E0f"]
        bwd = [j for j, k in zip(jobs, keys) if k.endswith("g")]
        self._jobs_first = (L.PackJob * len(first))(*first)
        self._jobs_fwd = (L.PackJob * len(fwd))(*fwd)
        self._jobs_bwd = (L.PackJob * len(bwd))(*bwd)
        self._bwd_packed = self._fwd_packed = None

    def pack(self):
        """Repack every operand form from the fp32 masters.  With a side stream only encoder conv 0's form is packed on
        the calling stream; encode() waits for the other forward forms before encoder conv 1, backward() for the
        data-gradient forms."""
        if self.side_stream is None:
            L.check(L.lib.cvae_pack_weights(self._jobs, len(self._jobs), L.stream_ptr()))
            self._bwd_packed = self._fwd_packed = None
            return
        fork = torch.cuda.Event()
        fork.record()
        self.side_stream.wait_event(fork)
        with torch.cuda.stream(self.side_stream):
            L.check(L.lib.cvae_pack_weights(self._jobs_fwd, len(self._jobs_fwd), L.stream_ptr()))
            self._fwd_packed = torch.cuda.Event()
            self._fwd_packed.record()
            L.check(L.lib.cvae_pack_weights(self._jobs_bwd, len(self._jobs_bwd), L.stream_ptr()))
            self._bwd_packed = torch.cuda.Event()
            self._bwd_packed.record()
        L.check(L.lib.cvae_pack_weights(self._jobs_first, len(self._jobs_first), L.stream_ptr()))

    def _await_fwd_pack(self):
        if self._fwd_packed is not None:
            torch.cuda.current_stream().wait_event(self._fwd_packed)
            self._fwd_packed = None

    def workspace(self, B, with_grad):
        key = (B, with_grad)
        ws = self._ws.get(key)
        if ws is None:
            if with_grad and (B, False) in self._ws:
                del self._ws[(B, False)]
            ws = self._ws[key] = Workspace(B, self.device, with_grad)
        return ws

    # ---- forward ------------------------------------------------------------------------------
    def _timed(self, family, fn):
        """Optional per-kernel-family CUDA-event timing (bench.py's roofline pass); off by default."""
        if self.profile is None:
            return fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        self.profile.append((family, a, b))

    def _ktab(self, key):
        return L.KTAB_BLOCK64 if key in self.wa_keys else L.KTAB_GENERIC

    def _conv(self, key=None, **kw):
        d = L.ConvDesc(**{k: (_ptr(v) if isinstance(v, torch.Tensor) else v) for k, v in kw.items()})
        if key in self.wa_keys:
            d.ktab, d.stack = L.KTAB_BLOCK64, self.wa_keys[key][1]
            need = int(L.lib.cvae_conv_gemm_workspace_bytes(ctypes.byref(d)))
            if need < 0:
                L.check(need)
            if need:   # split-K scratch, zero-initialised once: the kernel leaves its arrival counters at zero
                ws = self._conv_ws.get((key, d.batch))
                if ws is None or ws.numel() < need:
                    ws = self._conv_ws[(key, d.batch)] = torch.zeros(need, dtype=torch.uint8, device=self.device)
                d.workspace, d.workspace_bytes = ws.data_ptr(), ws.numel()
        self._timed("conv_gemm", lambda: L.check(L.lib.cvae_conv_gemm(ctypes.byref(d), L.stream_ptr())))

    @_nvtx("encode")
    def encode(self, x, training, ws, pack=True, fc=True):
        """x fp32 NCHW [B,3,64,64] -> ws.ml = mu | logvar.  vae_nets.py:101-111.
        fc=False stops in front of fc_mu / fc_var (the training step runs them inside bottleneck_forward)."""
        B, s = ws.B, L.stream_ptr()
        if pack:
            self.pack()
        if training:
            ws.stats.zero_()
        so = 0
        for i, (ci, co, h) in enumerate(ENC):
            stats = ws.stats[so:so + 2 * co]
            so += 2 * co
            if i == 0:
                self._conv(batch=B, height=h, width=h, ksize=5, src_channels=8, n_total=co, loader=L.LOAD_NCHW3,
                           epilogue=L.EPI_STATS, ktab=L.KTAB_PAIR8, src=x, wpack=self.packed["E0f"], out=ws.c[0], stats=stats)
            else:
                self._await_fwd_pack()
                self._conv(f"E{i}f", batch=B, height=h, width=h, ksize=5, src_channels=ci, n_total=co, loader=L.LOAD_NHWC,
                           epilogue=L.EPI_STATS, ktab=L.KTAB_GENERIC, src=ws.a[i - 1], wpack=self.packed[f"E{i}f"],
                           out=ws.c[i], stats=stats)
            cname, bname = f"encoder.model.{ENC_CONV_IDX[i]}", f"encoder.model.{ENC_BN_IDX[i]}"
            save = training and hasattr(ws, "xh")
            # BatchNorm finalize + normalise + 2x2 max-pool + activation in one launch
            L.check(L.lib.cvae_bn_fwd(B, h, h, co, L.ACT_TANH if i == 3 else L.ACT_RELU, int(training), _ptr(ws.c[i]), _ptr(stats),
                                      _ptr(self.view(bname + ".weight")), _ptr(self.view(bname + ".bias")), _ptr(self.view(cname + ".bias")),
                                      _ptr(self.running_mean[i]), _ptr(self.running_var[i]), _ptr(self.nbt[i]), BN_MOMENTUM, BN_EPS,
                                      _ptr(ws.ss[i]), _ptr(ws.a[i]), _ptr(ws.xh[i]) if save else None, _ptr(ws.am[i]) if save else None, s))
        if fc:
            L.check(L.lib.cvae_fc_fwd(B, _ptr(ws.a[3]), _ptr(self.packed["fc"]), _ptr(self.view("encoder.fc_mu.bias")),
                                      _ptr(self.view("encoder.fc_var.bias")), _ptr(ws.ml), s))
        return ws.ml

    @_nvtx("bottleneck_forward")
    def bottleneck_forward(self, pred, eps, ws):
        """fc_mu || fc_var -> reparametrise + critic concat -> decoder_input in one launch (vae_nets.py:108-111, 48-51,
        143-144): ws.a[3] -> ws.ml, ws.zc, ws.h0.  Training path (z is sampled); bit-identical to the three kernels."""
        self._await_fwd_pack()
        L.check(L.lib.cvae_bottleneck_fwd(ws.B, _ptr(ws.a[3]), _ptr(self.packed["fc"]), _ptr(self.view("encoder.fc_mu.bias")),
                                          _ptr(self.view("encoder.fc_var.bias")), _ptr(eps), _ptr(pred), _ptr(self.packed["decin"]),
                                          _ptr(ws.ml), _ptr(ws.zc), _ptr(ws.h0), L.stream_ptr()))

    @_nvtx("decode")
    def decode(self, pred, eps, sample, ws, pack=False, head=True):
        """ws.ml, pred fp32 [B] (+ eps fp32 [B,32]) -> ws.recon fp32 NCHW.  vae_nets.py:48-51,139-147.
        head=False starts behind decoder_input (ws.h0 comes from bottleneck_forward)."""
        B, s = ws.B, L.stream_ptr()
        if pack:
            self.pack()
        self._await_fwd_pack()       # (a decode() that packs itself, e.g. Decoder.forward: "fc" / "decin" are forward forms)
        if head:
            # fused reparametrise + critic concat + KL partial sums (consumed by loss_forward when it is given ws.ml itself)
            L.check(L.lib.cvae_latent_fwd(B, int(sample), _ptr(ws.ml), _ptr(eps), _ptr(pred), _ptr(ws.zc),
                                          _ptr(ws.kld_partial) if sample else None, s))
            L.check(L.lib.cvae_decin_fwd(B, _ptr(ws.zc), _ptr(self.packed["decin"]), _ptr(ws.h0), s))
        bias = lambda i: self.view(f"decoder.model.{DEC_CONV_IDX[i]}.bias")
        self._conv("D0f", batch=B, height=4, width=4, ksize=5, src_channels=256, n_total=128, loader=L.LOAD_NHWC,
                   epilogue=L.EPI_BIAS_RELU, ktab=L.KTAB_GENERIC, src=ws.h0, wpack=self.packed["D0f"], out=ws.d[0], bias=bias(0))
        for i in (1, 2, 3):
            ci, co, h = DEC[i]
            self._conv(f"D{i}f", batch=B, height=h, width=h, ksize=3, src_channels=ci, n_total=4 * co, loader=L.LOAD_NHWC,
                       epilogue=L.EPI_PHASE_BIAS_RELU, ktab=L.KTAB_GENERIC, src=ws.d[i - 1], wpack=self.packed[f"D{i}f"],
                       out=ws.d[i], bias=bias(i))
        self._conv(batch=B, height=32, width=32, ksize=3, src_channels=32, n_total=16, loader=L.LOAD_NHWC,
                   epilogue=L.EPI_PHASE_BIAS_TANH, ktab=L.KTAB_GENERIC, src=ws.d[3], wpack=self.packed["D4f"],
                   out=ws.recon, bias=bias(4))
        return ws.recon

    # ---- loss ---------------------------------------------------------------------------------
    @_nvtx("loss_forward")
    def loss_forward(self, recon, x, ml, ws, kld_weight=KLD_WEIGHT, fused_kld=False, split=False):
        """`fused_kld`: ml is ws.ml of the decode() that just ran with sample=True, so the KL partial sums the latent
        kernel left in ws.kld_partial are used instead of re-reading mu / logvar (a decode() with its head; after
        bottleneck_forward there are no partials and the loss kernel reduces mu / logvar itself)."""
        if split and self.side_stream is not None and self.profile is None:
            # level sums on this stream; the scalars the host reads (one block: KL reduction, five powf) on the side stream,
            # off the critical path -- loss_backward(split=True) derives its coefficients from the sums itself
            L.check(L.lib.cvae_loss_sums(ws.B, _ptr(recon), _ptr(x), self.window, _ptr(ws.loss_sums), L.stream_ptr()))
            ev = torch.cuda.Event()
            ev.record()
            self.side_stream.wait_event(ev)
            with torch.cuda.stream(self.side_stream):
                L.check(L.lib.cvae_loss_finalize(ws.B, _ptr(ml), _ptr(ws.kld_partial) if fused_kld else None, _ptr(ws.loss_sums), kld_weight,
                                                 _ptr(ws.coef), _ptr(ws.losses), L.stream_ptr()))
            self._loss_on_side = True
            return ws.losses
        L.check(L.lib.cvae_loss_fwd(ws.B, _ptr(recon), _ptr(x), _ptr(ml), _ptr(ws.kld_partial) if fused_kld else None, self.window,
                                    kld_weight, _ptr(ws.loss_sums), _ptr(ws.coef), _ptr(ws.losses), L.stream_ptr()))
        return ws.losses

    @_nvtx("loss_backward")
    def loss_backward(self, recon, x, ml, ws, grad_out=None, kld_weight=KLD_WEIGHT, fused_kld=False, split=False):
        """`fused_kld`: skip the KL term's backward here; backward(..., kld_grad_scale=kld_weight / B) adds it inside the
        latent backward kernel (only valid with grad_out None, i.e. an upstream gradient of 1).
        `split` (with fused_kld, after loss_forward(split=True)): the kernel computes its coefficients from ws.loss_sums."""
        if split and fused_kld:
            L.check(L.lib.cvae_loss_bwd_sums(ws.B, _ptr(recon), _ptr(x), self.window, _ptr(ws.loss_sums), _ptr(grad_out), _ptr(ws.d_recon),
                                             L.stream_ptr()))
            return ws.d_recon, ws.d_mu, ws.d_lv
        L.check(L.lib.cvae_loss_bwd(ws.B, _ptr(recon), _ptr(x), _ptr(ml), self.window, kld_weight, _ptr(ws.coef),
                                    _ptr(grad_out), _ptr(ws.d_recon), None if fused_kld else _ptr(ws.d_mu),
                                    None if fused_kld else _ptr(ws.d_lv), L.stream_ptr()))
        return ws.d_recon, ws.d_mu, ws.d_lv

    # ---- backward -----------------------------------------------------------------------------
    def _wgrad(self, g, name, **kw):
        d = L.WgradDesc(**{k: (_ptr(v) if isinstance(v, torch.Tensor) else v) for k, v in kw.items()})
        # A conv bias in front of BatchNorm has an exactly-zero gradient (the batch mean removes it): leave
        # the zero-initialised slot of the flat buffer untouched instead of writing rounding noise.
        d.dbias = None if name.startswith("encoder.") else _ptr(self.view(name + ".bias", g))
        need = int(L.lib.cvae_conv_wgrad_workspace_bytes(ctypes.byref(d)))
        # one split-K workspace per layer: the fold of one layer overlaps the GEMM of the next
        ws = self._wgrad_ws.get(name) if isinstance(self._wgrad_ws, dict) else None
        if ws is None or ws.numel() < need:
            if not isinstance(self._wgrad_ws, dict):
                self._wgrad_ws = {}
            ws = self._wgrad_ws[name] = torch.empty(need, dtype=torch.uint8, device=self.device)
        d.dw, d.workspace, d.workspace_bytes = _ptr(self.view(name + ".weight", g)), _ptr(ws), ws.numel()
        if self.side_stream is None or self.profile is not None:
            self._timed("conv_wgrad", lambda: L.check(L.lib.cvae_conv_wgrad(ctypes.byref(d), L.stream_ptr())))
            return
        ev = torch.cuda.Event()
        ev.record()
        self.side_stream.wait_event(ev)
        if self.fold_stream is not None:
            d.fold_stream = self.fold_stream.cuda_stream
            self._fold_used = True
        with torch.cuda.stream(self.side_stream):
            L.check(L.lib.cvae_conv_wgrad(ctypes.byref(d), L.stream_ptr()))
        self._side_used = True

    def _leaf(self, launch):
        """Run a parameter-gradient launch (a leaf of the backward pass) on the side stream when there is one."""
        if self.side_stream is None or self.profile is not None:
            launch(L.stream_ptr())
            return
        ev = torch.cuda.Event()
        ev.record()
        self.side_stream.wait_event(ev)
        with torch.cuda.stream(self.side_stream):
            launch(L.stream_ptr())
        self._side_used = True

    def _early_adam_launch(self):
        """Every gradient except encoder conv 0's is final once the main, side and fold streams reach this point: update
        those parameters on a stream of their own, beside the last weight-gradient GEMM (which runs alone at the end of
        the step otherwise, with the whole optimizer waiting behind it).  Single-GPU training step only (TrainStep sets
        self.early_adam = dict(lr=, grad_scale=) around backward(); the caller finishes with adam_step(hi=first_block_end()))."""
        if self.adam_stream is None:
            self.adam_stream = torch.cuda.Stream()
        st = self.adam_stream
        st.wait_stream(torch.cuda.current_stream())
        if self._side_used:
            st.wait_stream(self.side_stream)
        if self._fold_used:
            st.wait_stream(self.fold_stream)
        with torch.cuda.stream(st):
            self.adam_step(lo=self.first_block_end(), tick=False, **self.early_adam)
        self._adam_used = True

    def _join_leaves(self):
        """Wait (on the current stream) for the parameter-gradient launches that ran on the side / fold streams."""
        if self._side_used:
            torch.cuda.current_stream().wait_stream(self.side_stream)
        if self._fold_used:
            torch.cuda.current_stream().wait_stream(self.fold_stream)
        if self._adam_used:
            torch.cuda.current_stream().wait_stream(self.adam_stream)
        self._side_used = self._fold_used = self._adam_used = False

    def early_bucket_offset(self):
        """The flat gradient splits into [encoder convolutions + BatchNorm | everything else]: the second part (the two
        Linear heads, the decoder, decoder_input: 58 % of the floats) is complete once the backward pass has walked the
        decoder and the heads, long before the encoder's own gradients -- the data-parallel step all-reduces it beside them."""
        return self.offsets["encoder.fc_mu.weight"][0]

    def mid_bucket_offset(self):
        """Inside the encoder part, convs 2 and 3 (with their BatchNorm) hold 97 % of the floats and are final two layers
        before the end of the backward pass: [mid_bucket_offset(), early_bucket_offset()) is a bucket of its own."""
        return self.offsets[f"encoder.model.{ENC_CONV_IDX[2]}.weight"][0]

    @_nvtx("backward")
    def backward(self, x, eps, ws, d_recon, d_mu, d_lv, g=None, kld_grad_scale=0.0, stage="all"):
        """Gradients of every parameter into the flat buffer `g` (default self.gflat), given the
        gradients of the loss w.r.t. recon / mu / logvar.  Mirrors autograd through vae_nets.py:14-19.
        `stage`: "all", or "decoder" followed by "encoder" (two calls; after the first one the gradients from
        early_bucket_offset() on are final)."""
        g = self.gflat if g is None else g
        B, s = ws.B, L.stream_ptr()
        G = lambda n: self.view(n, g)
        if stage != "encoder":
            self._side_used = self._loss_on_side      # (a loss_finalize launched on the side stream by loss_forward(split=True))
            self._fold_used = False
            self._loss_on_side = False
            self._backward_decoder(x, eps, ws, d_recon, d_mu, d_lv, g, kld_grad_scale)
            if stage == "decoder":
                self._join_leaves()
                return g
            if self.early_event is not None and self.side_stream is not None and self.profile is None:
                # the gradients from early_bucket_offset() on are final once the side and fold streams reach this point:
                # mark it with an event another stream (the data-parallel all-reduce) can wait on, also from outside a
                # CUDA graph this call is captured into (external event record node)
                if self._fold_used:
                    self.side_stream.wait_stream(self.fold_stream)
                self.early_event.record(self.side_stream)
        self._backward_encoder(x, ws, g)
        self._join_leaves()
        return g

    def _backward_decoder(self, x, eps, ws, d_recon, d_mu, d_lv, g, kld_grad_scale):
        B, s = ws.B, L.stream_ptr()
        G = lambda n: self.view(n, g)
        if self._bwd_packed is not None:
            torch.cuda.current_stream().wait_event(self._bwd_packed)
            self._bwd_packed = None
        dm = "decoder.model."
        # D4 .. D1: up-sample-folded convs
        self._wgrad(g, dm + "12", kind=L.WGRAD_SHIFT_PHASE12, batch=B, height=32, width=32, cout=3, cin=32,
                    x=ws.d[3], dy=d_recon, dy2=ws.recon)
        self._conv(batch=B, height=32, width=32, ksize=3, src_channels=16, n_total=32, loader=L.LOAD_S2D_NCHW3_DTANH,
                   epilogue=L.EPI_MASK, ktab=L.KTAB_GENERIC, src=d_recon, src2=ws.recon, wpack=self.packed["D4g"],
                   out=ws.g_d[3], act=ws.d[3])
        for i in (3, 2, 1):
            ci, co, h = DEC[i]
            self._wgrad(g, f"{dm}{DEC_CONV_IDX[i]}", kind=L.WGRAD_PHASE, batch=B, height=h, width=h, cout=co, cin=ci,
                        x=ws.d[i - 1], dy=ws.g_d[i])
            self._conv(f"D{i}g", batch=B, height=h, width=h, ksize=3, src_channels=4 * co, n_total=ci, loader=L.LOAD_S2D,
                       epilogue=L.EPI_MASK, ktab=L.KTAB_GENERIC, src=ws.g_d[i], wpack=self.packed[f"D{i}g"],
                       out=ws.g_d[i - 1], act=ws.d[i - 1])
        self._wgrad(g, dm + "0", kind=L.WGRAD_5X5, batch=B, height=4, width=4, cout=128, cin=256, x=ws.h0, dy=ws.g_d[0])
        self._conv("D0g", batch=B, height=4, width=4, ksize=5, src_channels=128, n_total=256, loader=L.LOAD_NHWC,
                   epilogue=L.EPI_PLAIN, ktab=L.KTAB_GENERIC, src=ws.g_d[0], wpack=self.packed["D0g"], out=ws.g_h0)
        self._leaf(lambda st: L.check(L.lib.cvae_decin_bwd(B, _ptr(ws.g_h0), _ptr(ws.zc), None, None,
                                                           _ptr(G("decoder.decoder_input.weight")),
                                                           _ptr(G("decoder.decoder_input.bias")), st)))
        if d_mu is None and d_lv is None and self.fused_bottleneck:
            # decoder_input^T -> reparametrise backward (+ KL gradient) -> (fc_mu || fc_var)^T in one launch
            L.check(L.lib.cvae_bottleneck_bwd(B, _ptr(ws.g_h0), _ptr(self.packed["decin"]), _ptr(ws.ml), _ptr(eps), float(kld_grad_scale),
                                              _ptr(self.packed["fc"]), None, _ptr(ws.dml), _ptr(ws.g_a[3]), s))
            self._leaf(lambda st: L.check(L.lib.cvae_fc_bwd(B, _ptr(ws.dml), _ptr(ws.a[3]), None, None,
                                                            _ptr(G("encoder.fc_mu.weight")), _ptr(G("encoder.fc_var.weight")),
                                                            _ptr(G("encoder.fc_mu.bias")), _ptr(G("encoder.fc_var.bias")), st)))
            return
        L.check(L.lib.cvae_decin_bwd(B, _ptr(ws.g_h0), None, _ptr(self.packed["decin"]), _ptr(ws.dzc), None, None, s))
        L.check(L.lib.cvae_latent_bwd(B, _ptr(ws.ml), _ptr(eps), _ptr(ws.dzc), _ptr(d_mu), _ptr(d_lv), float(kld_grad_scale),
                                      _ptr(ws.dml), s))
        self._leaf(lambda st: L.check(L.lib.cvae_fc_bwd(B, _ptr(ws.dml), _ptr(ws.a[3]), None, None,
                                                        _ptr(G("encoder.fc_mu.weight")), _ptr(G("encoder.fc_var.weight")),
                                                        _ptr(G("encoder.fc_mu.bias")), _ptr(G("encoder.fc_var.bias")), st)))
        L.check(L.lib.cvae_fc_bwd(B, _ptr(ws.dml), None, _ptr(self.packed["fc"]), _ptr(ws.g_a[3]), None, None, None, None, s))

    def _backward_encoder(self, x, ws, g):
        B, s = ws.B, L.stream_ptr()
        G = lambda n: self.view(n, g)
        em = "encoder.model."
        for i in (3, 2, 1, 0):
            ci, co, h = ENC[i]
            bname, cname = f"{em}{ENC_BN_IDX[i]}", f"{em}{ENC_CONV_IDX[i]}"
            L.check(L.lib.cvae_bn_pool_act_bwd(B, h, h, co, L.ACT_TANH if i == 3 else L.ACT_RELU, _ptr(ws.c[i]), _ptr(ws.a[i]),
                                               _ptr(ws.g_a[i]), _ptr(ws.xh[i]), _ptr(ws.am[i]), _ptr(ws.ss[i]), _ptr(self.view(bname + ".weight")),
                                               _ptr(ws.bn_sums), _ptr(ws.g_c[i]), _ptr(G(bname + ".weight")),
                                               _ptr(G(bname + ".bias")), s))
            if i == 0:
                if self.early_adam is not None and self.side_stream is not None and self.profile is None and g is self.gflat:
                    self._early_adam_launch()
                self._wgrad(g, cname, kind=L.WGRAD_SHIFT_FRAMES, batch=B, height=64, width=64, cout=32, cin=3, x=x, dy=ws.g_c[0])
            else:
                self._wgrad(g, cname, kind=L.WGRAD_5X5, batch=B, height=h, width=h, cout=co, cin=ci, x=ws.a[i - 1], dy=ws.g_c[i])
                if i == 2 and self.mid_event is not None and self.side_stream is not None and self.profile is None:
                    # encoder convs 3 and 2 (+ their BatchNorm) are final once the side and fold streams reach this point:
                    # 97 % of the encoder's gradient floats, ready while convs 1 and 0 are still being walked
                    if self._fold_used:
                        self.side_stream.wait_stream(self.fold_stream)
                    self.mid_event.record(self.side_stream)
                self._conv(f"E{i}g", batch=B, height=h, width=h, ksize=5, src_channels=co, n_total=ci, loader=L.LOAD_NHWC,
                           epilogue=L.EPI_PLAIN, ktab=L.KTAB_GENERIC, src=ws.g_c[i], wpack=self.packed[f"E{i}g"], out=ws.g_a[i - 1])

    # ---- optimizer ----------------------------------------------------------------------------
    @_nvtx("adam")
    def adam_step(self, lr, grad_scale=1.0, betas=(0.9, 0.999), eps=1e-8, g=None, lo=0, hi=None, tick=True):
        """torch.optim.Adam's update (vae.py:36, :56) of the flat parameters [lo, hi) (default: all of them); `tick` advances
        the device step counter -- a step applied in several ranges ticks with the last one only."""
        if self.exp_avg is None:
            self.exp_avg, self.exp_avg_sq = torch.zeros_like(self.flat), torch.zeros_like(self.flat)
        g = self.gflat if g is None else g
        hi = self.n_params if hi is None else hi
        L.check(L.lib.cvae_adam_update(hi - lo, _ptr(self.flat[lo:hi]), _ptr(g[lo:hi]), _ptr(self.exp_avg[lo:hi]), _ptr(self.exp_avg_sq[lo:hi]),
                                       _ptr(self.step), lr, betas[0], betas[1], eps, grad_scale, int(tick), L.stream_ptr()))

    def first_block_end(self):
        """Flat offset behind encoder conv 0 and its BatchNorm: the last gradient of the backward pass to arrive (a multiple
        of 4 floats, so both sides of the split stay 16-byte aligned)."""
        return self.offsets[f"encoder.model.{ENC_CONV_IDX[1]}.weight"][0]

    @_nvtx("critic")
    def critic(self, x, weights, out=None):
        N = x.shape[0]
        out = torch.empty(N, 1, device=self.device) if out is None else out
        L.check(L.lib.cvae_critic_fwd(N, _ptr(x), _ptr(weights), _ptr(out), L.stream_ptr()))
        return out

    def check_fault(self):
        L.check(L.lib.cvae_check_device_fault(L.stream_ptr()))
