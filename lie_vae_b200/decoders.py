"""Decoders -- drop-in for the reference's ``lie_vae.decoders`` (``ActionNet`` on the sm_100a Wigner kernels;
``MLPNet``, the group-agnostic baseline, as host-side PyTorch)."""
import torch
from torch import nn as nn

from . import _ops
from .lie_tools import MAX_DEGREE


class MLP(nn.Sequential):
    """Same layer layout (hence ``state_dict`` keys) as the reference's ``experiments/nets.py:78-91``."""

    def __init__(self, input_dims, output_dims, hidden_dims, num_layers=1, activation=nn.ReLU):
        if num_layers == 0:
            layers = [nn.Linear(input_dims, output_dims)]
        else:
            layers = [nn.Linear(input_dims, hidden_dims), activation()]
            for _ in range(num_layers - 1):
                layers += [nn.Linear(hidden_dims, hidden_dims), activation()]
            layers.append(nn.Linear(hidden_dims, output_dims))
        super().__init__(*layers)


class ActionNet(nn.Module):
    """Learned harmonics acted on by the block Wigner-D of the pose (``decoders.py:9-61``).

    ``forward(angles (N,3))``: item_rep ((degrees+1)^2, rep_copies) is rotated by D(angles)
    (or D^T with ``transpose``) in one sm_100a kernel -- the (N,M,C) expand of the reference
    (``decoders.py:53``) is never materialised and the gradient w.r.t. ``item_rep`` is reduced
    over the batch inside the backward kernel -- then optionally an MLP, then ``deconv``.
    """

    def __init__(self, degrees, deconv, rep_copies=10, with_mlp=False, item_rep=None, transpose=False):
        super().__init__()
        if degrees > MAX_DEGREE:
            raise NotImplementedError("degrees > %d not supported by the sm_100a Wigner kernels" % MAX_DEGREE)
        self.degrees = degrees
        self.rep_copies = rep_copies
        self.matrix_dims = (degrees + 1) ** 2
        self.transpose = transpose

        if item_rep is None:
            self.item_rep = nn.Parameter(torch.randn((self.matrix_dims, rep_copies)))
        else:
            self.register_buffer('item_rep', item_rep)

        if with_mlp:
            self.mlp = MLP(self.matrix_dims * rep_copies, self.matrix_dims * rep_copies, 50, 3)
        else:
            self.mlp = None

        self.deconv = deconv

    def forward(self, angles):
        """Input is ZYZ Euler angles."""
        n, d = angles.shape
        assert d == 3, 'Input should be Euler angles.'
        item = _ops.wigner_apply(angles, self.item_rep, 0, self.degrees, self.transpose) \
            .view(-1, self.matrix_dims * self.rep_copies)
        if self.mlp:
            item = self.mlp(item)
        return self.deconv(item)


class MLPNet(nn.Module):
    """Baseline decoder of the reference (``decoders.py:64-87``): the flattened group element ((N,9) matrix or (N,3)
    angles) goes through an MLP to a ((degrees+1)^2 * rep_copies)-vector and then through ``deconv``.  No group
    action, hence no kernel of this package: plain PyTorch with the reference's layer layout."""

    def __init__(self, degrees, deconv, in_dims=9, rep_copies=10, layers=3, hidden_dims=50, activation=nn.ReLU):
        super().__init__()
        matrix_dims = (degrees + 1) ** 2
        self.mlp = MLP(in_dims, matrix_dims * rep_copies, hidden_dims, layers, activation)
        self.deconv = deconv

    def forward(self, x, content_data=None):
        return self.deconv(self.mlp(x.view(x.size(0), -1)))
