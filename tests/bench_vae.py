#!/usr/bin/env python
"""BASELINE configs[3] (SURVEY 8d "config 4"): end-to-end SO(3) VAE with the action decoder on synthetic 64x64 RGB images,
batch 1024, random-init weights -- one training step (forward, backward, Adam) with the SO(3) hot path

  kernels : this package's modules (SO3reparameterize with fused encoder heads, group_matrix_to_eazyz, ActionNet)
  kernels_fused : the same with ActionNet.fuse_consumer: action + first ConvTranspose2d as L2-resident chunks of
            [Wigner forward kernel -> tcgen05 TF32 GEMM]
  eager   : the reference algorithm as eager PyTorch on the same GPU (the oracle's device-agnostic restatement, J cached per
            device as the reference's lru_cache does) -- what `python main.py` runs today
  none    : the hot path replaced by a Linear stand-in (encoder + deconv only), to read off the hot path's share

The conv encoder / deconv decoder are plain PyTorch with the reference's layer layout (experiments/nets.py:33-76,
experiments/vae.py:56-120; degrees 6, rep_copies 10, group_reparam_in_dims 10, k = 10, deconv_hidden 200 = main.py:167).  Measurement tooling that lives under tests/
because its "before" arm executes the oracle (test infrastructure); not collected by pytest, not product code.
"""
import argparse
import json
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lie_vae_b200.lie_tools as lt  # noqa: E402
import lie_vae_b200.reparameterize as rp  # noqa: E402
import lie_vae_b200.decoders as dc  # noqa: E402
from oracle import so3_oracle as O  # noqa: E402  (the "before" arm only)

_jcache = {}
_j_orig = O.j_matrix


def _j_cached(l, dtype=torch.float64, device=None):
    key = (l, dtype, str(device))
    if key not in _jcache:
        _jcache[key] = _j_orig(l, dtype, device)
    return _jcache[key]


O.j_matrix = _j_cached


def conv_encoder(out_dims, hidden=50):                       # nets.py:33-57 (ConvNetBN, rgb)
    layers, c = [], 3
    for mult in (1, 2, 4, 8):
        layers += [nn.Conv2d(c, hidden * mult, 4, 2, 1), nn.BatchNorm2d(hidden * mult), nn.LeakyReLU(0.2, inplace=True)]
        c = hidden * mult
    return nn.Sequential(*layers, nn.Conv2d(c, out_dims, 4, 1, 0), nn.Flatten())


class View(nn.Module):                                       # experiments/utils.py:36-42
    def __init__(self, *v):
        super().__init__()
        self.v = v

    def forward(self, x):
        return x.view(*self.v)


def deconv_decoder(in_dims, hidden):                         # nets.py:60-76 (DeconvNet, rgb)
    layers = [View(-1, in_dims, 1, 1), nn.ConvTranspose2d(in_dims, hidden, 4, 1, 0), nn.ReLU()]
    for _ in range(3):
        layers += [nn.ConvTranspose2d(hidden, hidden, 4, 2, 1), nn.ReLU()]
    return nn.Sequential(*layers, nn.ConvTranspose2d(hidden, 3, 4, 2, 1))


class EagerLatent(nn.Module):
    """SO3reparameterize (reparameterize.py:200-278) + mean map + N0reparameterize as eager torch ops."""

    def __init__(self, din, mean_mode, k):
        super().__init__()
        self.mean_mode, self.k = mean_mode, k
        self.map = nn.Linear(din, 3 if mean_mode == "alg" else 6)
        self.sigma_linear = nn.Linear(din, 3)

    def forward(self, h, n=1):
        if self.mean_mode == "alg":
            mu = O.rodrigues(self.map(h))
        else:
            v = self.map(h).double().view(-1, 2, 3)
            mu = O.s2s2_gram_schmidt(v[:, 0], v[:, 1]).float()
        sigma = F.softplus(self.sigma_linear(h))
        eps = torch.randn(n, h.shape[0], 3, device=h.device)
        z, log_q = O.so3_reparameterize(mu, sigma, eps, self.k)
        self._kl = (log_q - O.so3_log_prior(z)).mean(0)
        return z

    def kl(self):
        return self._kl


class EagerAction(nn.Module):
    def __init__(self, degrees, copies):
        super().__init__()
        self.degrees = degrees
        self.item_rep = nn.Parameter(torch.randn((degrees + 1) ** 2, copies))

    def forward(self, angles):
        return O.action_net_forward(angles, self.item_rep, self.degrees)


class VAE(nn.Module):
    def __init__(self, hot, mean_mode, degrees=6, copies=10, din=10, k=10, deconv_hidden=200):
        super().__init__()
        self.hot = hot
        M = (degrees + 1) ** 2
        self.encoder = conv_encoder(din)
        self.deconv = deconv_decoder(M * copies, deconv_hidden)
        if hot in ("kernels", "kernels_fused"):
            mean = rp.AlgebraMean(din) if mean_mode == "alg" else rp.S2S2Mean(din)
            self.latent = rp.SO3reparameterize(rp.N0reparameterize(din, 3), mean, k=k)
            self.action = dc.ActionNet(degrees, self.deconv, rep_copies=copies)      # decoders.py:61: the decoder owns the deconv stack
            self.action.fuse_consumer = hot == "kernels_fused"
            self.deconv = nn.Sequential()
            self.eazyz = lt.group_matrix_to_eazyz
        elif hot == "eager":
            self.latent = EagerLatent(din, mean_mode, k)
            self.action = EagerAction(degrees, copies)
            self.eazyz = O.group_matrix_to_eazyz
        else:
            self.stand_in = nn.Linear(din, M * copies)

    def loss(self, x):
        h = self.encoder(x)
        if self.hot == "none":
            rec, kl = self.deconv(self.stand_in(h)), 0.0
        else:
            z = self.latent(h, 1)                                         # (1,B,3,3)            vae.py:134-143
            kl = self.latent.kl()
            rec = self.deconv(self.action(self.eazyz(z.view(-1, 3, 3))))  #                      vae.py:173-190
        return (((rec - x) ** 2).sum((-1, -2, -3)) + kl).mean()           #                      vae.py:199-204


def run(hot, mean_mode, batch, steps, warm, deconv_hidden=200):
    torch.manual_seed(0)
    model = VAE(hot, mean_mode, deconv_hidden=deconv_hidden).cuda()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    x = torch.rand(batch, 3, 64, 64, device="cuda")

    def step():
        opt.zero_grad(set_to_none=True)
        loss = model.loss(x)
        loss.backward()
        opt.step()
        return loss
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        loss = step()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps, float(loss.detach())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--deconv-hidden", type=int, default=200, help="main.py:167 default; the VAE class default is 50")
    args = ap.parse_args()
    for mean_mode in ("s2s2", "alg"):                                     # s2s2 is the CLI default (main.py:155)
        t = {hot: run(hot, mean_mode, args.batch, args.steps, args.warmup, args.deconv_hidden) for hot in ("none", "eager", "kernels", "kernels_fused")}
        nets = t["none"][0]
        print(json.dumps({
            "config": "configs[3]: SO(3) VAE, action decoder, 64x64x3 images, batch %d, degrees 6, rep_copies 10, k 10, deconv_hidden %d, mean_mode %s"
                      % (args.batch, args.deconv_hidden, mean_mode),
            "ms_per_step": {k: round(v[0], 3) for k, v in t.items()},
            "hot_path_ms": {"eager": round(t["eager"][0] - nets, 3), "kernels": round(t["kernels"][0] - nets, 3),
                            "kernels_fused (also replaces the nets' first deconv layer)": round(t["kernels_fused"][0] - nets, 3)},
            "hot_path_share_of_step": {"eager": round(1 - nets / t["eager"][0], 3), "kernels": round(1 - nets / t["kernels"][0], 3)},
            "step_speedup": {"kernels": round(t["eager"][0] / t["kernels"][0], 2), "kernels_fused": round(t["eager"][0] / t["kernels_fused"][0], 2)},
            "images_per_s": {k: round(args.batch / v[0] * 1e3) for k, v in t.items()},
            "loss": {k: round(v[1], 2) for k, v in t.items()}}), flush=True)


if __name__ == "__main__":
    main()
