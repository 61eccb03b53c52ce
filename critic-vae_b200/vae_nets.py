"""Drop-in for the reference's vae_nets.py: same classes, attributes, call signatures and state_dict
keys (encoder `model.{0,1,4,5,8,9,12,13}.*`, `fc_mu.*`, `fc_var.*`; decoder `model.{0,3,6,9,12}.*`,
`decoder_input.*`), but every tensor op of the hot path runs in libcvae.so (sm_100a kernels behind
include/cvae.h).  The torch modules below only *hold* parameters -- their own forward() is never used.

Reference behaviour followed (file:line in the reference checkout):
  VariationalAutoencoder.forward / evaluate / inject / recon_samples / reparametrize / vae_loss
      vae_nets.py:14-62
  VariationalEncoder vae_nets.py:64-111, Decoder :113-147, MSSIM :150-247.
Additive extension: `forward(x, pred, eps=None)` / `set_eps()` let the host supply the reparameterisation
noise (the reference draws it inside reparametrize, :50) so results can be compared bit-for-bit.
`evaluate` also accepts batches (the reference's evalu=True path is batch-1 only, :140-142).
"""
import weakref

import torch
from torch import nn, Tensor

from vae_parameters import *  # noqa: F401,F403  (device, ch, k, step, p, bottleneck, latent_dim, inject_n, ...)
from cvae_native import binding as _L
from cvae_native.engine import VAEEngine, ENC_CONV_IDX, ENC_BN_IDX


def _container_encoder(dims):
    layers, cin = [], ch
    for i, cout in enumerate(dims):
        layers += [nn.Conv2d(cin, cout, k, step, p), nn.BatchNorm2d(cout), nn.MaxPool2d(2),
                   nn.Tanh() if i == len(dims) - 1 else nn.ReLU()]
        cin = cout
    return nn.Sequential(*layers)


def _container_decoder(dims):
    chans = [dims[3], dims[2], dims[1], dims[0], dims[0], ch]
    layers = []
    for i in range(5):
        layers.append(nn.Conv2d(chans[i], chans[i + 1], k, step, p))
        layers += [nn.ReLU(), nn.Upsample(scale_factor=2)] if i < 4 else [nn.Tanh()]
    return nn.Sequential(*layers)


class _Bound:
    """Mixin: finds the owning VariationalAutoencoder (which owns the engine)."""
    _owner = None

    def _vae(self):
        owner = self._owner() if self._owner is not None else None
        if owner is None:
            raise _L.CvaeError("encoder/decoder must live inside a VariationalAutoencoder to run (it owns the CUDA engine)")
        return owner


class VariationalEncoder(nn.Module, _Bound):
    def __init__(self, dims):
        super().__init__()
        self.model = _container_encoder(dims)
        self.fc_mu = nn.Linear(bottleneck, latent_dim)
        self.fc_var = nn.Linear(bottleneck, latent_dim)

    def forward(self, x):
        """mu, log_var of a batch of frames (inference path; training goes through the fused
        VariationalAutoencoder.forward so that one autograd node covers the whole network)."""
        vae = self._vae()
        eng, ws = vae._prepare(x.shape[0])
        ml = eng.encode(vae._frames(x), self.training, ws).clone()
        return ml[:, :latent_dim], ml[:, latent_dim:]


class Decoder(nn.Module, _Bound):
    def __init__(self, dims):
        super().__init__()
        self.model = _container_decoder(dims)
        self.decoder_input = nn.Linear(latent_dim + 1, bottleneck)

    def forward(self, z, pred, evalu=False, dim=1):
        vae = self._vae()
        if evalu:                      # reference: z = z[0], concat along dim 0 (batch of one)
            z = z[:1]
        B = z.shape[0]
        eng, ws = vae._prepare(B)
        ws.ml[:, :latent_dim].copy_(z)
        pred = pred.reshape(-1).to(z.device, torch.float32).expand(B).contiguous()
        return eng.decode(pred, None, False, ws, pack=True).clone()


class _VAEFunction(torch.autograd.Function):
    """encoder -> reparametrize -> decoder as ONE autograd node (vae_nets.py:14-19)."""

    @staticmethod
    def forward(ctx, vae, x, pred, eps, *params):
        eng, ws = vae._prepare(x.shape[0])
        eng.encode(x, vae.encoder.training, ws)
        eng.decode(pred, eps, True, ws)
        ws.generation = getattr(ws, "generation", 0) + 1
        ctx.vae, ctx.ws, ctx.x, ctx.eps, ctx.generation = vae, ws, x, eps, ws.generation
        return ws.ml[:, :latent_dim].clone(), ws.ml[:, latent_dim:].clone(), ws.recon.clone()

    @staticmethod
    def backward(ctx, d_mu, d_logvar, d_recon):
        vae, ws = ctx.vae, ctx.ws
        eng = vae._engine
        if getattr(ws, "generation", 0) != ctx.generation:
            raise _L.CvaeError("backward() after another forward() of the same batch size: activations were overwritten")
        B = ws.B
        zeros = lambda *s: torch.zeros(*s, device=ctx.x.device)
        d_mu = zeros(B, latent_dim) if d_mu is None else d_mu.contiguous()
        d_logvar = zeros(B, latent_dim) if d_logvar is None else d_logvar.contiguous()
        d_recon = zeros(B, ch, w, w) if d_recon is None else d_recon.contiguous()
        # write into the persistent flat gradient buffer unless the caller is accumulating
        accumulating = any(q.grad is not None for q in vae._flat_params)
        if accumulating and eng.gflat_alt is None:
            eng.gflat_alt = torch.zeros_like(eng.gflat)
        g = eng.backward(ctx.x, ctx.eps, ws, d_recon, d_mu, d_logvar, eng.gflat_alt if accumulating else eng.gflat)
        grads = tuple(eng.view(name, g) for name in vae._flat_names)
        return (None, None, None, None) + grads


class _LossFunction(torch.autograd.Function):
    """MS-SSIM + KLD (vae_nets.py:53-62) with the backward of both in one node."""

    @staticmethod
    def forward(ctx, vae, x, mu, logvar, recon):
        eng, ws = vae._prepare(x.shape[0])
        ml = torch.cat((mu, logvar), dim=1).contiguous()
        recon, x = recon.contiguous(), x.contiguous()
        losses = eng.loss_forward(recon, x, ml, ws, kld_weight).clone()
        ctx.vae, ctx.ws, ctx.saved = vae, ws, (recon, x, ml, ws.coef.clone())
        return losses[0], losses[1], losses[2]

    @staticmethod
    def backward(ctx, g_total, g_recon, g_kld):
        recon, x, ml, coef = ctx.saved
        ws, eng = ctx.ws, ctx.vae._engine
        ws.coef.copy_(coef)
        if g_total is None:
            g_total = g_recon        # MSSIM used on its own: only the reconstruction term is live
        g = g_total.reshape(1).float().contiguous() if g_total is not None else None
        d_recon, d_mu, d_lv = eng.loss_backward(recon, x, ml, ws, g, kld_weight)
        return None, None, d_mu, d_lv, d_recon


class VariationalAutoencoder(nn.Module):
    def __init__(self, dims=[32, 64, 128, 256]):
        super(VariationalAutoencoder, self).__init__()
        if list(dims) != [32, 64, 128, 256]:
            raise _L.CvaeError("the sm_100a kernels are specialised for dims=[32, 64, 128, 256]")
        self.encoder = VariationalEncoder(dims)
        self.decoder = Decoder(dims)
        self.mssim_loss = MSSIM()
        self.encoder._owner = self.decoder._owner = self.mssim_loss._owner = weakref.ref(self)
        self._engine = None
        self._eps = None

    # ---- engine plumbing ------------------------------------------------------------------------
    def _named_flat(self):
        sd = dict(self.named_parameters())
        return [(n, sd[n]) for n, _ in self._engine.layout]

    def _bind(self):
        """Make every parameter a view into the engine's flat buffer (and keep it that way after
        .to(), load_state_dict() or any re-assignment of .data), and hand the BatchNorm buffers over."""
        any_p = next(self.parameters())
        if any_p.device.type != "cuda":
            raise _L.CvaeError("VariationalAutoencoder must be on a CUDA device: call .to(device) first; there is no CPU path")
        if self._engine is None or self._engine.device != any_p.device:
            self._engine = VAEEngine(any_p.device)
        eng = self._engine
        base = eng.flat.data_ptr()
        pairs = self._named_flat()
        for name, prm in pairs:
            off = eng.offsets[name][0]
            if prm.data_ptr() != base + 4 * off:
                v = eng.view(name)
                v.copy_(prm.data)
                prm.data = v
        self._flat_params = [q for _, q in pairs]
        self._flat_names = [n for n, _ in pairs]
        for i, bi in enumerate(ENC_BN_IDX):
            bn = self.encoder.model[bi]
            eng.running_mean[i], eng.running_var[i], eng.nbt[i] = bn.running_mean, bn.running_var, bn.num_batches_tracked
        return eng

    def _prepare(self, B):
        eng = self._bind()
        return eng, eng.workspace(B, True)

    @staticmethod
    def _frames(x):
        return x.to(torch.float32).contiguous()

    def set_eps(self, eps):
        """Supply the N(0,1) noise the next forward() / recon_samples() call(s) will use."""
        self._eps = eps

    def _take_eps(self, like, eps=None):
        if eps is None:
            eps, self._eps = self._eps, None
        if eps is None:
            eps = torch.randn(like, latent_dim, device=next(self.parameters()).device)
        return eps.to(torch.float32).contiguous()

    # ---- reference API --------------------------------------------------------------------------
    def forward(self, x, pred, eps=None):
        x = self._frames(x)
        B = x.shape[0]
        self._bind()
        pred = pred.reshape(-1).to(x.device, torch.float32).contiguous()
        eps = self._take_eps(B, eps)
        mu, logvar, recon = _VAEFunction.apply(self, x, pred, eps, *self._flat_params)
        return x, mu, logvar, recon

    def recon_samples(self, x, reward):
        mu, logvar = self.encoder(x)
        recons = []
        for _ in range(6):
            sample = self.reparametrize(mu, logvar)
            recons.append(self.decoder(sample, reward))
        return recons

    def inject(self, x, reward=Tensor([0, 0.2, 0.4, 0.6, 0.8, 1])):
        reward = reward.to(next(self.parameters()).device)
        mu, _ = self.encoder(x)
        return [self.decoder(mu, reward[i].view(1), evalu=True) for i in range(inject_n)]

    def evaluate(self, x, pred):
        """Decode the mean with the given critic value(s).  x: (B,3,64,64); pred: B values (or one)."""
        x = self._frames(x)
        B = x.shape[0]
        eng, ws = self._prepare(B)
        eng.encode(x, self.encoder.training, ws)
        pred = pred.reshape(-1).to(x.device, torch.float32).expand(B).contiguous()
        return eng.decode(pred, None, False, ws).clone()

    def reparametrize(self, mu, logvar, eps=None):
        eps = self._take_eps(mu.shape[0], eps)
        return mu + eps * torch.exp(0.5 * logvar)

    def vae_loss(self, x, mu, logvar, recon):
        total, recon_loss, kld = _LossFunction.apply(self, x, mu, logvar, recon)
        return {'total_loss': total, 'recon_loss': recon_loss.detach(), 'KLD': kld.detach()}


class MSSIM(nn.Module, _Bound):
    """Same constructor as the reference (vae_nets.py:152-167); forward = 1 - MS-SSIM(img1, img2)."""

    def __init__(self, in_channels: int = 3, window_size: int = 11, size_average: bool = True) -> None:
        super(MSSIM, self).__init__()
        if (in_channels, window_size, size_average) != (3, 11, True):
            raise _L.CvaeError("the MS-SSIM kernel is specialised for 3 channels, window 11, size_average=True")
        self.in_channels, self.window_size, self.size_average = in_channels, window_size, size_average

    def forward(self, img1: Tensor, img2: Tensor) -> Tensor:
        vae = self._vae()
        B = img1.shape[0]
        zero = torch.zeros(B, latent_dim, device=img1.device)      # KLD of (0, 0) is exactly 0
        _, recon_loss, _ = _LossFunction.apply(vae, img2, zero, zero, img1)
        return recon_loss
