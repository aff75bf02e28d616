"""On-GPU synthetic pose data -- the tensors of the reference's ``ToyDataset.generate``
(``experiments/datasets.py:142-158``): uniform random quaternions, one random harmonic signal scaled to norm 10,
and the signal rotated by every pose with the block Wigner-D action.

The reference builds the set in batches of 64 through ~75 ATen launches per batch; here it is one elementwise kernel
(quaternion -> Euler) and one Wigner forward launch over all n samples (the (n,M,C) expand of the signal is a
stride-0 view and is never materialised).
"""
import torch

from .lie_tools import random_quaternions, quaternions_to_eazyz, block_wigner_matrix_multiply

__all__ = ["toy_tensors"]


def toy_tensors(n=1000, degrees=6, rep_copies=10, device="cuda", seed=0, quaternions=None, harmonics=None):
    """Returns ``(q (n,4), harmonics expanded to (n,M,C), x (n,M,C))`` -- the ``tensors`` of ``ToyDataset``.

    ``seed`` reproduces the reference's ``torch.manual_seed(0)`` convention on this device (its values differ from a
    CPU run: different generator); pass ``quaternions`` / ``harmonics`` to rotate given data instead.
    """
    dev = torch.device(device)
    if harmonics is None or quaternions is None:
        gen_state = torch.random.get_rng_state(), (torch.cuda.get_rng_state(dev) if dev.type == "cuda" else None)
        torch.manual_seed(seed)
        torch.cuda.manual_seed(seed)
        if harmonics is None:
            harmonics = torch.randn((degrees + 1) ** 2, rep_copies, device=dev)
            harmonics = harmonics / harmonics.norm() * 10
        if quaternions is None:
            quaternions = random_quaternions(n, device=dev)
        torch.random.set_rng_state(gen_state[0])
        if gen_state[1] is not None:
            torch.cuda.set_rng_state(gen_state[1], dev)
    n = quaternions.shape[0]
    expanded = harmonics.expand(n, -1, -1)
    x = block_wigner_matrix_multiply(quaternions_to_eazyz(quaternions), expanded, degrees)
    return quaternions, expanded, x
