"""Deterministic synthetic weights / frames shared by the golden generator, the tests, smoke() and
bench.py.  numpy's PCG64 stream is stable across numpy versions, so nothing but the seed has to be
committed for the weights: the container that makes the golden vectors and the GPU box that checks
them rebuild identical tensors."""
from __future__ import annotations

import numpy as np
import torch

DIMS = (32, 64, 128, 256)
ENC_CONV = {0: (32, 3), 4: (64, 32), 8: (128, 64), 12: (256, 128)}       # idx -> (Cout, Cin)
ENC_BN = {1: 32, 5: 64, 9: 128, 13: 256}
DEC_CONV = {0: (128, 256), 3: (64, 128), 6: (32, 64), 9: (32, 32), 12: (3, 32)}


def _uniform(rng, shape, bound):
    return torch.from_numpy(rng.uniform(-bound, bound, size=shape).astype(np.float32))


def make_vae_state(seed: int = 0, nontrivial_bn: bool = True):
    """(encoder_state_dict, decoder_state_dict) with the reference's keys and PyTorch-default-like
    scales (U(+-1/sqrt(fan_in))).  With `nontrivial_bn` the BN affine and running buffers are
    perturbed so that eval-mode folding and the affine gradients are actually exercised."""
    rng = np.random.Generator(np.random.PCG64(seed))
    bn_rng = np.random.Generator(np.random.PCG64(seed + 1000))   # separate stream: conv weights
    enc, dec = {}, {}                                             # do not depend on the BN flag
    for idx, (co, ci) in ENC_CONV.items():
        b = 1.0 / np.sqrt(ci * 25)
        enc[f"model.{idx}.weight"] = _uniform(rng, (co, ci, 5, 5), b)
        enc[f"model.{idx}.bias"] = _uniform(rng, (co,), b)
        c = ENC_BN[idx + 1]
        if nontrivial_bn:
            enc[f"model.{idx + 1}.weight"] = 1.0 + _uniform(bn_rng, (c,), 0.2)
            enc[f"model.{idx + 1}.bias"] = _uniform(bn_rng, (c,), 0.2)
            enc[f"model.{idx + 1}.running_mean"] = _uniform(bn_rng, (c,), 0.1)
            enc[f"model.{idx + 1}.running_var"] = 0.25 + _uniform(bn_rng, (c,), 0.1).abs()
        else:
            enc[f"model.{idx + 1}.weight"] = torch.ones(c)
            enc[f"model.{idx + 1}.bias"] = torch.zeros(c)
            enc[f"model.{idx + 1}.running_mean"] = torch.zeros(c)
            enc[f"model.{idx + 1}.running_var"] = torch.ones(c)
        enc[f"model.{idx + 1}.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    for name in ("fc_mu", "fc_var"):
        b = 1.0 / np.sqrt(4096)
        enc[f"{name}.weight"] = _uniform(rng, (32, 4096), b)
        enc[f"{name}.bias"] = _uniform(rng, (32,), b)
    for idx, (co, ci) in DEC_CONV.items():
        b = 1.0 / np.sqrt(ci * 25)
        dec[f"model.{idx}.weight"] = _uniform(rng, (co, ci, 5, 5), b)
        dec[f"model.{idx}.bias"] = _uniform(rng, (co,), b)
    b = 1.0 / np.sqrt(33)
    dec["decoder_input.weight"] = _uniform(rng, (4096, 33), b)
    dec["decoder_input.bias"] = _uniform(rng, (4096,), b)
    return enc, dec


def make_frames(n: int, seed: int = 1, quantize: bool = True) -> torch.Tensor:
    """(n,3,64,64) fp32 in [0,1]: smooth blobs + noise, optionally snapped to k/255 like the
    reference's uint8/255 frames (vae_utility.py:324-328)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    yy, xx = np.mgrid[0:64, 0:64].astype(np.float32) / 64.0
    out = np.empty((n, 3, 64, 64), dtype=np.float32)
    for i in range(n):
        for c in range(3):
            fx, fy, ph = rng.uniform(0.5, 4.0), rng.uniform(0.5, 4.0), rng.uniform(0, 6.28)
            base = 0.5 + 0.35 * np.sin(6.28 * (fx * xx + fy * yy) + ph)
            out[i, c] = np.clip(base + rng.normal(0, 0.08, (64, 64)), 0, 1)
    if quantize:
        out = np.round(out * 255.0) / 255.0
    return torch.from_numpy(out.astype(np.float32))


def make_eps(n: int, seed: int = 2) -> torch.Tensor:
    rng = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy(rng.standard_normal((n, 32)).astype(np.float32))


def make_gt_masks(n: int, seed: int = 3) -> np.ndarray:
    """(n,64,64) bool ground truth: one random rectangle per frame (stand-in for Y.npy)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    g = np.zeros((n, 64, 64), dtype=bool)
    for i in range(n):
        y0, x0 = rng.integers(0, 40, 2)
        h, w = rng.integers(8, 24, 2)
        g[i, y0:y0 + h, x0:x0 + w] = True
    return g


def sample_indices(numel: int, k: int = 48) -> np.ndarray:
    """Fixed probe positions inside a flattened tensor (first few + an even stride)."""
    if numel <= k:
        return np.arange(numel)
    head = np.arange(8)
    stride = np.linspace(8, numel - 1, k - 8).astype(np.int64)
    return np.unique(np.concatenate([head, stride]))
