"""Multi-GPU plumbing: one process per GPU, pages sharded by image, no data-path collective.

The only coupling between ranks is the batch-global normaliser of both losses
(reference ``model/losses.py:40-44``, ``:88-90``; with ``keras.utils.multi_gpu_model`` the loss sees the
merged batch, ``RetinaNet.py:106-112``): the per-rank positive counts K1 returns are summed with ONE
``all_reduce`` before K2 scales its gradients, and the two loss sums are reduced afterwards.  Works with
any ``torch.distributed`` backend (NCCL over NVLink on the GPU box; gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_pages(num_pages, rank=None, world_size=None):
    """Contiguous page range ``[lo, hi)`` of this rank; the first ``num_pages % world`` ranks take one
    extra page, so every page is owned exactly once."""
    if rank is None or world_size is None:
        rank, world_size = world()
    base, extra = divmod(int(num_pages), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def global_positive_count(npos_per_page, group=None):
    """Sum of the per-page positive-anchor counts over all pages of all ranks, as a 1-element float32
    tensor on the same device (the ``normalizer`` argument of the loss functors)."""
    total = npos_per_page.to(torch.float32).sum().reshape(1)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    return total


def reduce_losses(losses, group=None):
    """Each rank's loss is already divided by the GLOBAL normaliser, so the batch loss is the plain sum."""
    out = losses.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out
