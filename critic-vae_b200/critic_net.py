"""Drop-in for the reference's critic_net.py: same constructor, parameter containers and state_dict
keys (`features.{0,3,6,10,14}.*`, `crit.{1,4}.*`), so the shipped checkpoints load unchanged.  The
forward pass (critic_net.py:44-59, evaluate :66-69) runs in one fused sm_100a kernel per frame tile
(csrc/critic.cu); the torch layers below only hold the parameters."""
import numpy as np
import torch
from torch import nn, Tensor

from cvae_native import binding as _L


class Critic(nn.Module):
    def __init__(self, width=64, dims=[8, 8, 8, 16], bottleneck=32, colorchs=3,
                 chfak=1, activation=nn.ReLU, pool='max', dropout=0.5):
        super().__init__()
        if (width, list(dims), bottleneck, colorchs, chfak, activation, pool) != (64, [8, 8, 8, 16], 32, 3, 1, nn.ReLU, 'max'):
            raise _L.CvaeError("the critic kernel is specialised for the reference's default architecture")
        self.width = width
        self.pool = nn.MaxPool2d(2)
        d = list(dims)
        feats = []
        for i, (ci, co) in enumerate(zip([colorchs] + d[:-1], d)):
            feats += [nn.Conv2d(ci, co, 3, 1, 1), activation(), self.pool]
            if i >= 2:
                feats.append(nn.Dropout(dropout))
        feats += [nn.Conv2d(d[-1], bottleneck, 4), activation()]
        self.features = nn.Sequential(*feats)
        self.crit = nn.Sequential(nn.Flatten(), nn.Linear(bottleneck, bottleneck), activation(), nn.Dropout(dropout),
                                  nn.Linear(bottleneck, 1), nn.Sigmoid())
        self._flat, self._flat_key = None, None

    def _weights(self):
        prms = list(self.parameters())
        key = tuple((q.data_ptr(), q._version) for q in prms)
        if key != self._flat_key:
            self._flat = torch.cat([q.detach().reshape(-1).float() for q in prms]).contiguous()
            self._flat_key = key
        return self._flat

    def forward(self, X, collect=False):
        if collect:
            raise _L.CvaeError("collect=True (intermediate embeddings) is not part of the Critic-VAE hot path")
        if self.training:
            raise _L.CvaeError("the critic is a frozen value network: call .eval() (dropout is not implemented)")
        X = X.to(torch.float32).contiguous()
        if X.device.type != "cuda":
            raise _L.CvaeError("Critic must run on a CUDA device; there is no CPU path")
        out = torch.empty(X.shape[0], 1, device=X.device)
        _L.check(_L.lib.cvae_critic_fwd(X.shape[0], X.data_ptr(), self._weights().data_ptr(), out.data_ptr(), _L.stream_ptr()))
        return out

    def preprocess(self, X: Tensor):
        return (X / 255.0).permute(0, 3, 1, 2).float()

    def evaluate(self, X):
        with torch.no_grad():
            return self.forward(X, collect=False)
