"""Batch-sharded data parallelism for the SO(3) hot path (host-side plumbing over torch.distributed).

The reference is single-device (``experiments/main.py:17``); the path shards naturally by batch:
every sample is independent in the forward, per-sample gradients (mu, sigma, angles) stay on the rank
that owns the sample, and the only cross-sample quantities are the scalar loss and the gradient of
the decoder's shared spectrum ``item_rep``.  One all-reduce(sum) of the packed buffer
``[loss, grad item_rep.flatten()]`` per step is therefore the only collective (NCCL over NVLink on
GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist

__all__ = ["shard_bounds", "pack_reduction", "unpack_reduction", "allreduce_packed", "allreduce_loss_and_grad"]


def shard_bounds(total, world_size, rank, multiple=1):
    """Contiguous slice [lo, hi) of ``total`` samples owned by ``rank``.

    Shards differ by at most ``multiple`` samples and are multiples of ``multiple`` (the micro-batch)
    except possibly the last one, which takes the remainder.  All n samples of a datapoint live on
    one rank, so grad mu / grad sigma reduce locally.
    """
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank %d of %d" % (rank, world_size))
    if total < 0 or multiple < 1:
        raise ValueError("bad sizes")
    units = total // multiple
    base, extra = divmod(units, world_size)
    lo_u = rank * base + min(rank, extra)
    hi_u = lo_u + base + (1 if rank < extra else 0)
    lo, hi = lo_u * multiple, hi_u * multiple
    if rank == world_size - 1:
        hi = total
    return lo, hi


def pack_reduction(loss, grad_item_rep, out=None):
    """[loss, grad item_rep.flatten()] as one flat float32 buffer (one collective, one launch)."""
    n = 1 + grad_item_rep.numel()
    if out is None:
        out = torch.empty(n, dtype=torch.float32, device=grad_item_rep.device)
    elif out.numel() != n:
        raise ValueError("reduction buffer has %d elements, need %d" % (out.numel(), n))
    out[0] = loss
    out[1:] = grad_item_rep.reshape(-1)
    return out


def unpack_reduction(buf, shape):
    """Inverse of ``pack_reduction``: (loss scalar tensor, grad item_rep of ``shape``)."""
    return buf[0], buf[1:].reshape(shape)


def allreduce_packed(buf, group=None):
    """The path's one collective: all-reduce(sum) of an already packed ``[loss, grad item_rep]`` buffer, in place
    (a step captured in a CUDA graph packs inside the graph and reduces outside it).  A no-op without a process group."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf


def allreduce_loss_and_grad(loss, grad_item_rep, group=None, out=None):
    """Sum the local loss and item_rep gradient over all ranks; a no-op without a process group."""
    buf = allreduce_packed(pack_reduction(loss, grad_item_rep, out), group)
    return unpack_reduction(buf, grad_item_rep.shape)
