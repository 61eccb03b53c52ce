"""CPU oracle for the Critic-VAE hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain functional torch-CPU fp32 (network, losses) and numpy fp64 (mask
pipeline), the algorithm of the reference's hot path.  Only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s cpu_baseline / `--impl reference` legs may import it; the product
(`critic-vae_b200/`) never does and fails loudly when its CUDA library is missing.

Parity pin: the reference has no tests or golden vectors of its own, so this
restatement is pinned against outputs of the *unmodified reference modules* run in the build
container by `tests/golden/make_golden.py` (fixtures committed under `tests/golden/`); see
`tests/test_oracle_vs_golden.py`.

Every function cites the reference lines it follows (paths relative to the reference checkout).
State dicts use the reference's key names, so reference checkpoints drop in.
"""
from __future__ import annotations

import math
import statistics

import numpy as np
import torch
import torch.nn.functional as F

# vae_parameters.py:4-21
CH, WIDTH, KSIZE, PAD, LATENT, BOTTLENECK = 3, 64, 5, 2, 32, 4096
KLD_WEIGHT = 0.001
DIMS = (32, 64, 128, 256)
ENC_CONV = (0, 4, 8, 12)       # vae_nets.py:69,74,79,84  (indices inside encoder.model)
ENC_BN = (1, 5, 9, 13)         # vae_nets.py:70,75,80,85
DEC_CONV = (0, 3, 6, 9, 12)    # vae_nets.py:117,121,125,129,133
BN_EPS, BN_MOMENTUM = 1e-5, 0.1  # torch.nn.BatchNorm2d defaults used at vae_nets.py:70
MSSSIM_WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)  # vae_nets.py:219
GREY = (0.2989, 0.5870, 0.1140)  # vae_utility.py:273


# ----------------------------------------------------------------------------------------------
# Network
# ----------------------------------------------------------------------------------------------
def encoder_forward(enc: dict, x: torch.Tensor, training: bool, update_stats: bool = True):
    """vae_nets.py:101-111 over the Sequential of vae_nets.py:68-88.

    conv5x5(pad 2) -> BatchNorm2d -> MaxPool2d(2) -> ReLU (Tanh after the 4th block), then the two
    Linear heads on the NCHW-flattened 256x4x4 map.  In training mode batch statistics normalise and
    the running buffers are updated in place when `update_stats`.
    """
    h = x
    for i, (ci, bi) in enumerate(zip(ENC_CONV, ENC_BN)):
        h = F.conv2d(h, enc[f"model.{ci}.weight"], enc[f"model.{ci}.bias"], stride=1, padding=PAD)
        rm, rv = enc[f"model.{bi}.running_mean"], enc[f"model.{bi}.running_var"]
        if training and not update_stats:
            rm, rv = rm.clone(), rv.clone()
        h = F.batch_norm(h, rm, rv, enc[f"model.{bi}.weight"], enc[f"model.{bi}.bias"],
                         training=training, momentum=BN_MOMENTUM, eps=BN_EPS)
        if training and update_stats:
            enc[f"model.{bi}.num_batches_tracked"] += 1
        h = F.max_pool2d(h, 2)
        h = torch.tanh(h) if i == 3 else torch.relu(h)
    flat = torch.flatten(h, start_dim=1)                                   # vae_nets.py:105
    mu = F.linear(flat, enc["fc_mu.weight"], enc["fc_mu.bias"])            # vae_nets.py:108
    logvar = F.linear(flat, enc["fc_var.weight"], enc["fc_var.bias"])      # vae_nets.py:109
    return mu, logvar


def decoder_forward(dec: dict, z: torch.Tensor, pred: torch.Tensor, evalu: bool = False):
    """vae_nets.py:139-147: concat critic value, Linear(33->4096), view (-1,256,4,4), then
    4 x [conv5x5 -> ReLU -> nearest Upsample x2] and conv5x5 -> Tanh (vae_nets.py:116-135)."""
    if evalu:                       # vae_nets.py:140-142: batch of one, concat along dim 0
        zc = torch.cat((z[0], pred), dim=0)
    else:
        zc = torch.cat((z, pred), dim=1)
    h = F.linear(zc, dec["decoder_input.weight"], dec["decoder_input.bias"]).view(-1, 256, 4, 4)
    for i, ci in enumerate(DEC_CONV):
        h = F.conv2d(h, dec[f"model.{ci}.weight"], dec[f"model.{ci}.bias"], stride=1, padding=PAD)
        if i < 4:
            h = F.interpolate(torch.relu(h), scale_factor=2, mode="nearest")
        else:
            h = torch.tanh(h)
    return h


def reparametrize(mu, logvar, eps):
    """vae_nets.py:48-51 with the noise supplied by the caller instead of randn_like."""
    return mu + eps * torch.exp(0.5 * logvar)


def vae_forward(enc, dec, x, pred, eps, training=True, update_stats=True):
    """vae_nets.py:14-19."""
    mu, logvar = encoder_forward(enc, x, training, update_stats)
    recon = decoder_forward(dec, reparametrize(mu, logvar, eps), pred)
    return x, mu, logvar, recon


def vae_evaluate(enc, dec, x, pred):
    """vae_nets.py:42-46 (eval-mode BN, decodes the mean, batch of one)."""
    mu, _ = encoder_forward(enc, x, training=False)
    return decoder_forward(dec, mu, pred.view(1), evalu=True)


def vae_inject(enc, dec, x, rewards=(0.0, 0.2, 0.4, 0.6, 0.8, 1.0)):
    """vae_nets.py:31-40."""
    mu, _ = encoder_forward(enc, x, training=False)
    return [decoder_forward(dec, mu, torch.tensor([r], dtype=torch.float32), evalu=True) for r in rewards]


# ----------------------------------------------------------------------------------------------
# Loss
# ----------------------------------------------------------------------------------------------
def msssim_window_1d() -> torch.Tensor:
    """vae_nets.py:170-173.  NOTE the upstream sign: exp(+d^2/(2 sigma^2)), edge-heavy weights."""
    k = torch.tensor([math.exp((i - 11 // 2) ** 2 / (2 * 1.5 ** 2)) for i in range(11)])
    return k / k.sum()


def msssim_window_2d(channels=3) -> torch.Tensor:
    """vae_nets.py:175-179."""
    g = msssim_window_1d().unsqueeze(1)
    return g.mm(g.t()).float()[None, None].expand(channels, 1, 11, 11).contiguous()


_MSSSIM_CONST = {}


def _msssim_constants(channels, device):
    """(window, level weights) on `device`, built once per device: bench.py's gpu_baseline leg captures the step in a
    CUDA graph, where a host-to-device copy of a fresh constant is not allowed."""
    key = (channels, str(device))
    if key not in _MSSSIM_CONST:
        _MSSSIM_CONST[key] = (msssim_window_2d(channels).to(device), torch.tensor(MSSSIM_WEIGHTS, dtype=torch.float32).to(device))
    return _MSSSIM_CONST[key]


def ssim_level(a, b, win):
    """vae_nets.py:181-215 with size_average=True: returns (ssim mean, cs mean) over all of B,3,H,W."""
    c = a.shape[1]
    blur = lambda t: F.conv2d(t, win, padding=5, groups=c)
    mu_a, mu_b = blur(a), blur(b)
    var_a = blur(a * a) - mu_a * mu_a
    var_b = blur(b * b) - mu_b * mu_b
    cov = blur(a * b) - mu_a * mu_b
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    v1 = 2.0 * cov + c2
    v2 = var_a + var_b + c2
    cs = torch.mean(v1 / v2)
    ssim = torch.mean(((2 * mu_a * mu_b + c1) * v1) / ((mu_a * mu_a + mu_b * mu_b + c1) * v2))
    return ssim, cs


def msssim_loss(recon, x):
    """vae_nets.py:217-247: 1 - prod_{l<4}( cs_l^w_l * ssim_4^w_4 )."""
    win, w = _msssim_constants(recon.shape[1], recon.device)
    a, b = recon, x
    ss, cs = [], []
    for _ in range(5):
        s, c = ssim_level(a, b, win)
        ss.append(s)
        cs.append(c)
        a, b = F.avg_pool2d(a, (2, 2)), F.avg_pool2d(b, (2, 2))
    ss, cs = torch.stack(ss), torch.stack(cs)
    return 1 - torch.prod((cs ** w)[:-1] * (ss ** w)[-1])


def msssim_level_means(recon, x):
    """The ten batch-global means (ssim_l, cs_l) of vae_nets.py:224-236, for kernel-level checks."""
    win, _ = _msssim_constants(recon.shape[1], recon.device)
    a, b = recon, x
    out = []
    for _ in range(5):
        out.append(ssim_level(a, b, win))
        a, b = F.avg_pool2d(a, (2, 2)), F.avg_pool2d(b, (2, 2))
    return out


def kld_loss(mu, logvar):
    """vae_nets.py:57-58."""
    return torch.mean(-0.5 * torch.sum(1 + logvar - mu ** 2 - logvar.exp(), dim=1), dim=0) * KLD_WEIGHT


def vae_loss(x, mu, logvar, recon):
    """vae_nets.py:53-62."""
    r = msssim_loss(recon, x)
    k = kld_loss(mu, logvar)
    return {"total_loss": r + k, "recon_loss": r.detach(), "KLD": k.detach()}


# ----------------------------------------------------------------------------------------------
# Critic (critic_net.py:15-42, evaluate at :66-69; dropout inactive in eval mode)
# ----------------------------------------------------------------------------------------------
def critic_forward(crit: dict, x: torch.Tensor) -> torch.Tensor:
    h = x
    for idx in (0, 3, 6, 10):                                   # 3x3 pad 1 convs + ReLU + MaxPool2
        h = F.max_pool2d(torch.relu(F.conv2d(h, crit[f"features.{idx}.weight"],
                                             crit[f"features.{idx}.bias"], padding=1)), 2)
    h = torch.relu(F.conv2d(h, crit["features.14.weight"], crit["features.14.bias"]))  # 4x4, no pad
    h = torch.flatten(h, 1)
    h = torch.relu(F.linear(h, crit["crit.1.weight"], crit["crit.1.bias"]))
    return torch.sigmoid(F.linear(h, crit["crit.4.weight"], crit["crit.4.bias"]))


# ----------------------------------------------------------------------------------------------
# Training step (vae.py:47-58) with Adam restated from torch.optim.Adam defaults (vae.py:36)
# ----------------------------------------------------------------------------------------------
PARAM_KEYS_ENC = [f"model.{i}.{p}" for i in (0, 1, 4, 5, 8, 9, 12, 13) for p in ("weight", "bias")] + \
    ["fc_mu.weight", "fc_mu.bias", "fc_var.weight", "fc_var.bias"]
PARAM_KEYS_DEC = [f"model.{i}.{p}" for i in DEC_CONV for p in ("weight", "bias")] + \
    ["decoder_input.weight", "decoder_input.bias"]


def loss_and_grads(enc, dec, x, pred, eps, update_stats=True):
    """Forward (training-mode BN), vae_loss, backward.  Returns (loss dict, recon, mu, logvar, grads)
    with grads keyed 'encoder.<k>' / 'decoder.<k>'."""
    enc_l = {k: (v.detach().clone().requires_grad_(True) if k in PARAM_KEYS_ENC else v) for k, v in enc.items()}
    dec_l = {k: v.detach().clone().requires_grad_(True) for k, v in dec.items()}
    _, mu, logvar, recon = vae_forward(enc_l, dec_l, x, pred, eps, training=True, update_stats=update_stats)
    if update_stats:
        for k in enc:
            if "running" in k or "tracked" in k:
                enc[k] = enc_l[k]
    losses = vae_loss(x, mu, logvar, recon)
    losses["total_loss"].backward()
    grads = {f"encoder.{k}": enc_l[k].grad for k in PARAM_KEYS_ENC}
    grads.update({f"decoder.{k}": dec_l[k].grad for k in PARAM_KEYS_DEC})
    return losses, recon.detach(), mu.detach(), logvar.detach(), grads


def adam_step(param, grad, m, v, step, lr=5e-5, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam (no amsgrad, no weight decay) single-tensor update, step counted from 1."""
    m.mul_(b1).add_(grad, alpha=1 - b1)
    v.mul_(b2).addcmul_(grad, grad, value=1 - b2)
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    param.addcdiv_(m, denom, value=-lr / bc1)


# ----------------------------------------------------------------------------------------------
# Mask pipeline (numpy, fp64) -- vae_utility.py
# ----------------------------------------------------------------------------------------------
def diff_grey(recon_one: np.ndarray, recon_zero: np.ndarray):
    """vae_utility.py:270-275 on (3,64,64) fp32 arrays -> ((64,64) float64 diff, float64 max)."""
    d = np.abs(np.subtract(recon_zero, recon_one))
    d = np.transpose(d, (1, 2, 0))
    d = np.dot(d[..., :3], list(GREY))
    return d, np.amax(d)


def diff_grey_ordered(recon_one: np.ndarray, recon_zero: np.ndarray):
    """Same quantity with the accumulation order fixed as ((r*wr + g*wg) + b*wb) in fp64 -- the
    order the CUDA kernel uses.  np.dot's BLAS order may differ from this by <= 1 ulp,
    which is why the bit-exact contract starts from the difference map."""
    d = np.abs(recon_zero.astype(np.float32) - recon_one.astype(np.float32)).astype(np.float64)
    g = (d[0] * GREY[0] + d[1] * GREY[1]) + d[2] * GREY[2]
    return g, g.max()


def diff_factor(max_values):
    """vae_utility.py:106-110."""
    mean_max = statistics.mean(max_values)
    return (1.0 / mean_max if mean_max != 0 else 0), mean_max


def diff_and_thr_masks(diffs, max_values, thr=50):
    """vae_utility.py:148-160 + prepare_diff :279-284.  diffs: sequence of (64,64) float64."""
    factor, mean_max = diff_factor(max_values)
    out_d, out_t = [], []
    for d in diffs:
        d = np.array(d, dtype=np.float64, copy=True)
        d[d > mean_max] = mean_max
        q = (d * factor * 255).astype(np.uint8)
        out_d.append(q)
        out_t.append(q > thr)
    return np.array(out_d), np.array(out_t)


def iou_counts(G: np.ndarray, T: np.ndarray):
    """Integer tp/fn/fp of vae_utility.py:57-59."""
    tp = int(np.sum(G & T))
    fn = int(np.sum(G & np.logical_not(T)))
    fp = int(np.sum(np.logical_not(G) & T))
    return tp, fn, fp


def iou(G, T):
    """vae_utility.py:56-68."""
    tp, fn, fp = iou_counts(G, T)
    val = 1 if tp + fn + fp == 0 else tp / (tp + fn + fp)
    return round(val, 3)


# ----------------------------------------------------------------------------------------------
# Precision model of the device path (test calibration only)
# ----------------------------------------------------------------------------------------------
class _StoreBF16(torch.autograd.Function):
    """Round a tensor to bf16 where the device path stores it in bf16 (activations forward,
    activation gradients backward)."""

    @staticmethod
    def forward(ctx, t):
        return t.to(torch.bfloat16).float()

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).float()


def loss_and_grads_bf16_storage(enc, dec, x, pred, eps):
    """Same computation as loss_and_grads (fp32 math, training-mode BN, no buffer update) but with every
    inter-layer activation, its gradient and the conv weights rounded to bf16 -- i.e. what ANY
    implementation with bf16 tensor-core operands computes at best.  The gap between this and the
    fp32 result is the error budget the bf16 design choice itself costs (ReLU / max-pool gates flip
    under 2^-9 relative perturbations); the GPU tests require the kernels to stay within a small
    factor of it."""
    r = _StoreBF16.apply
    wq = lambda t: t + (t.to(torch.bfloat16).float() - t).detach()
    enc_l = {k: (v.detach().clone().requires_grad_(True) if k in PARAM_KEYS_ENC else v.clone()) for k, v in enc.items()}
    dec_l = {k: v.detach().clone().requires_grad_(True) for k, v in dec.items()}
    h = r(x)
    for i, (ci, bi) in enumerate(zip(ENC_CONV, ENC_BN)):
        h = r(F.conv2d(h, wq(enc_l[f"model.{ci}.weight"]), enc_l[f"model.{ci}.bias"], padding=PAD))
        h = F.batch_norm(h, None, None, enc_l[f"model.{bi}.weight"], enc_l[f"model.{bi}.bias"], training=True, eps=BN_EPS)
        h = F.max_pool2d(h, 2)
        h = r(torch.tanh(h) if i == 3 else torch.relu(h))
    flat = torch.flatten(h, 1)
    mu = F.linear(flat, enc_l["fc_mu.weight"], enc_l["fc_mu.bias"])
    logvar = F.linear(flat, enc_l["fc_var.weight"], enc_l["fc_var.bias"])
    z = reparametrize(mu, logvar, eps)
    h = r(F.linear(torch.cat((z, pred), 1), dec_l["decoder_input.weight"], dec_l["decoder_input.bias"]).view(-1, 256, 4, 4))
    for i, ci in enumerate(DEC_CONV):
        h = F.conv2d(h, wq(dec_l[f"model.{ci}.weight"]), dec_l[f"model.{ci}.bias"], padding=PAD)
        h = F.interpolate(r(torch.relu(h)), scale_factor=2, mode="nearest") if i < 4 else torch.tanh(h)
    losses = vae_loss(x, mu, logvar, h)
    losses["total_loss"].backward()
    grads = {f"encoder.{k}": enc_l[k].grad for k in PARAM_KEYS_ENC}
    grads.update({f"decoder.{k}": dec_l[k].grad for k in PARAM_KEYS_DEC})
    return losses, h.detach(), mu.detach(), logvar.detach(), grads
