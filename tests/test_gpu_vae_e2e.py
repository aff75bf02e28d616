"""End-to-end training-step parity (BASELINE configs[0] / configs[3] in miniature).

A VAE step composed exactly as the reference's ``VAE.elbo`` (``experiments/vae.py:134-215``: encoder -> SO(3)
reparameterize -> pose -> decoder -> squared-error reconstruction + KL, ``(recon + kl).mean().backward()``) is run
twice on the same weights, inputs and noise:

  * with this package's modules on the GPU (``SO3reparameterize`` with every mean map, ``ActionNet`` / ``MLPNet``),
  * with the CPU oracle in float64 (functional restatement of the same composition),

and the loss and the gradient of EVERY parameter must agree.  The encoder / decoder nets are plain PyTorch (callers of
the hot path, ``experiments/nets.py:78-91``); what is being checked is that the kernels compose into the reference's
training step as drop-ins.  ``-m gpu``.
"""
import math

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import so3_oracle as O

pytestmark = pytest.mark.gpu

DEG, COPIES, DIN, B = 3, 3, 10, 64          # configs[0]: ToyDataset(degrees=3, rep_copies=3), batch 64, reparam in-dims 10


def build(mean_mode, decoder_mode, k, dtype):
    import lie_vae_b200.reparameterize as rp
    import lie_vae_b200.decoders as dc
    torch.manual_seed(7)
    M = (DEG + 1) ** 2
    encoder = nn.Sequential(nn.Flatten(), dc.MLP(M * COPIES, DIN, 100, 2))                      # vae.py:66-70
    mean = {"alg": rp.AlgebraMean, "q": rp.QuaternionMean, "s2s1": rp.S2S1Mean, "s2s2": rp.S2S2Mean}[mean_mode](DIN)
    if mean_mode == "s2s2":       # the reference's U(-10,10) init gives near-degenerate frames; keep the test well conditioned
        mean.map.weight.data.uniform_(-1, 1)
        mean.map.bias.data.uniform_(-1, 1)
    rep = rp.SO3reparameterize(rp.N0reparameterize(DIN, 3), mean, k=k)                           # vae.py:73-88
    if decoder_mode == "action":
        decoder = dc.ActionNet(DEG, nn.Sequential(), rep_copies=COPIES)                          # vae.py:113-120
    else:
        decoder = dc.MLPNet(DEG, nn.Sequential(), in_dims=9, rep_copies=COPIES)                  # vae.py:121-130
    return nn.ModuleDict(dict(encoder=encoder, rep=rep, decoder=decoder)).to(dtype)


def elbo_ours(model, x, eps, n, decoder_mode):
    import lie_vae_b200.lie_tools as lt
    rep = model["rep"]
    rep.reparameterize.sample_noise = lambda n_=1: eps           # the same noise on both sides
    z = rep(model["encoder"](x), n)                               # (n,B,3,3)          vae.py:134-143
    kl = rep.kl()                                                 # (B,)               vae.py:145-147
    zp = z.view(-1, 3, 3)                                         #                    vae.py:173-176
    if decoder_mode == "action":
        recon = model["decoder"](lt.group_matrix_to_eazyz(zp))    #                    vae.py:182-190
    else:
        recon = model["decoder"](zp)
    recon = recon.reshape(n, x.shape[0], *x.shape[1:])
    loss_rec = ((recon - x.expand_as(recon)) ** 2).sum(-1).sum(-1)         # (n,B)     vae.py:199-204
    return (loss_rec + kl).mean()                                 # main.py: (recon + kl).mean()


def mlp64(sd, prefix, h):
    idx = sorted({int(k_[len(prefix):].split(".")[0]) for k_ in sd if k_.startswith(prefix)})
    for j, i in enumerate(idx):
        h = F.linear(h, sd["%s%d.weight" % (prefix, i)], sd["%s%d.bias" % (prefix, i)])
        if j + 1 < len(idx):
            h = F.relu(h)
    return h


def elbo_oracle(sd, x, eps, n, mean_mode, decoder_mode, k):
    h = mlp64(sd, "encoder.1.", x.flatten(1))
    lin = lambda name: F.linear(h, sd["rep.mean_module.%s.weight" % name], sd["rep.mean_module.%s.bias" % name])   # noqa: E731
    if mean_mode == "alg":
        mu = O.rodrigues(lin("map"))
    elif mean_mode == "q":
        mu = O.quaternions_to_group_matrix(lin("map"))
    elif mean_mode == "s2s1":
        s2, s1 = lin("s2_map"), lin("s1_map")
        mu = O.s2s1rodrigues(s2 / s2.norm(dim=-1, keepdim=True), s1 / s1.norm(dim=-1, keepdim=True))
    else:
        v = lin("map").view(-1, 2, 3)
        mu = O.s2s2_gram_schmidt(v[:, 0], v[:, 1])
    sigma = F.softplus(F.linear(h, sd["rep.reparameterize.sigma_linear.weight"], sd["rep.reparameterize.sigma_linear.bias"]))
    z, log_q = O.so3_reparameterize(mu, sigma, eps, k)
    kl = (log_q - O.so3_log_prior(z)).mean(0)
    zp = z.reshape(-1, 3, 3)
    if decoder_mode == "action":
        recon = O.action_net_forward(O.group_matrix_to_eazyz(zp), sd["decoder.item_rep"], DEG)
    else:
        recon = mlp64(sd, "decoder.mlp.", zp.reshape(-1, 9))
    recon = recon.reshape(n, x.shape[0], *x.shape[1:])
    return (((recon - x.expand_as(recon)) ** 2).sum(-1).sum(-1) + kl).mean()


@pytest.mark.parametrize("mean_mode,decoder_mode,n,k,dtype", [
    ("alg", "mlp", 1, 10, torch.float32),          # configs[0]: toy VAE, SO3 reparameterize + MLP decoder
    ("alg", "action", 1, 10, torch.float32),       # configs[3] in miniature: action decoder
    ("q", "action", 3, 3, torch.float32),
    ("s2s1", "action", 2, 10, torch.float32),
    ("s2s2", "action", 1, 10, torch.float32),      # the CLI default mean map (main.py:155)
    ("alg", "action", 2, 10, torch.float64),       # the float64 instantiations of every kernel on the path
    ("q", "mlp", 1, 5, torch.float64),
])
def test_vae_step_matches_oracle(mean_mode, decoder_mode, n, k, dtype):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    model = build(mean_mode, decoder_mode, k, dtype).cuda()
    g = torch.Generator().manual_seed(11)
    x64 = torch.randn(B, (DEG + 1) ** 2, COPIES, generator=g, dtype=torch.float64)
    eps64 = torch.randn(n, B, 3, generator=g, dtype=torch.float64)
    loss = elbo_ours(model, x64.to(dtype).cuda(), eps64.to(dtype).cuda(), n, decoder_mode)
    loss.backward()
    sd = {k_: v.detach().double().cpu().requires_grad_(True) for k_, v in model.state_dict().items()}
    ref = elbo_oracle(sd, x64, eps64, n, mean_mode, decoder_mode, k)
    ref.backward()
    tol = 1e-9 if dtype == torch.float64 else 2e-5
    assert abs(loss.item() - ref.item()) <= tol * max(1.0, abs(ref.item())), (loss.item(), ref.item())
    checked = 0
    for name, p in model.named_parameters():
        gref = sd[name].grad
        assert p.grad is not None and gref is not None, name
        scale = max(1.0, gref.abs().max().item())
        gtol = (1e-8 if dtype == torch.float64 else 5e-4) * scale
        err = (p.grad.double().cpu() - gref).abs().max().item()
        assert err <= gtol, "%s: gradient differs by %.3g (scale %.3g)" % (name, err, scale)
        checked += 1
    assert checked >= 8


def test_toy_dataset_tensors():
    """ToyDataset.generate (experiments/datasets.py:142-158) on the GPU kernels against the oracle on the same poses."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from lie_vae_b200.toy import toy_tensors
    q, h, x = toy_tensors(n=1000, degrees=6, rep_copies=10)
    assert tuple(q.shape) == (1000, 4) and tuple(x.shape) == (1000, 49, 10) and h.stride(0) == 0
    assert abs(h[0].norm().item() - 10.0) < 1e-4
    assert torch.allclose(q.norm(dim=-1), torch.ones(1000, device="cuda"), atol=1e-5)
    ref = O.block_wigner_matrix_multiply(O.quaternions_to_eazyz(q.double().cpu()), h.double().cpu(), 6)
    assert (x.double().cpu() - ref).abs().max().item() < 2e-5 * 10
    # rotations preserve the signal's norm degree by degree
    for l in range(7):
        a, b = x[:, l * l:(l + 1) ** 2].norm(dim=1), h[:, l * l:(l + 1) ** 2].norm(dim=1)
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-4)
    q2, _, x2 = toy_tensors(n=1000, degrees=6, rep_copies=10)
    assert torch.equal(q, q2) and torch.equal(x, x2)               # seeded like the reference
