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
        # fuse_consumer: run the action together with the layer that consumes it -- the first Linear of the MLP or the first
        # ConvTranspose2d(M*C -> hidden, 4, 1, 0) of a DeconvNet-shaped ``deconv`` (experiments/nets.py:63-66) -- as chunks of
        # [Wigner forward kernel -> tcgen05 TF32 GEMM] whose intermediate y stays in L2.  TF32 operands (what
        # cuDNN uses for the reference's FP32 convolutions by default); off by default: the FP32-exact path is the parity path.
        self.fuse_consumer = False

    def _consumer(self):
        """(weight (M*C, Nout), bias, bias_div, view shape after it, layers that follow, then ``deconv``?) of the layer that
        consumes the action output, or None if its shape is not one the fused op covers."""
        mc = self.matrix_dims * self.rep_copies
        if self.mlp is not None:
            lin = self.mlp[0]
            if isinstance(lin, nn.Linear) and lin.in_features == mc:
                return lin.weight.t(), lin.bias, 1, None, list(self.mlp)[1:], True
            return None
        layers = list(self.deconv) if isinstance(self.deconv, nn.Sequential) else []
        if len(layers) >= 2 and isinstance(layers[1], nn.ConvTranspose2d) and type(layers[0]).__name__ == "View":
            ct = layers[1]
            if (ct.in_channels == mc and tuple(ct.kernel_size) == (4, 4) and tuple(ct.stride) == (1, 1) and tuple(ct.padding) == (0, 0)
                    and tuple(ct.output_padding) == (0, 0) and tuple(ct.dilation) == (1, 1) and ct.groups == 1):
                return ct.weight.reshape(mc, ct.out_channels * 16), ct.bias, 16, (ct.out_channels, 4, 4), layers[2:], False
        return None

    def forward(self, angles):
        """Input is ZYZ Euler angles."""
        n, d = angles.shape
        assert d == 3, 'Input should be Euler angles.'
        cons = self._consumer() if (self.fuse_consumer and angles.is_cuda and angles.dtype == torch.float32
                                    and self.degrees <= _ops.FAST_MAX_DEGREE) else None
        if cons is not None:
            weight, bias, bias_div, shape, rest, then_deconv = cons
            x = _ops.ActionGemm.apply(angles, self.item_rep, weight, bias, bias_div, self.degrees, self.transpose, _ops.ACTION_GEMM_CHUNK)
            if shape is not None:
                x = x.view(-1, *shape)
            for layer in rest:
                x = layer(x)
            return self.deconv(x) if then_deconv else x
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
