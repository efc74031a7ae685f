"""Multi-GPU plumbing: one process per GPU, pages sharded by image, no data-path collective.

The only coupling between ranks is the batch-global normaliser of both losses
(reference ``model/losses.py:40-44``, ``:88-90``; with ``keras.utils.multi_gpu_model`` the loss sees the
merged batch, ``RetinaNet.py:106-112``): the per-rank positive counts K1 returns are summed with ONE
``all_reduce`` before K2 scales its gradients, and the two loss sums are reduced afterwards.  Works with
any ``torch.distributed`` backend (NCCL over NVLink on the GPU box; gloo in the CPU tests).
"""
import ctypes
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib


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


def balanced_shards(weights, world_size):
    """Load-aware page sharding: K1's cost grows with the number of GT tables on a page, and K2 of every rank waits
    for the slowest rank's K1 (the normaliser is global), so pages are dealt out by descending weight in snake
    order -- every rank gets the same number of pages (+-1) and nearly the same total weight.  ``weights``: one
    number per page of the GLOBAL batch (e.g. its GT count).  Returns ``world_size`` ascending page-index lists;
    deterministic, so every rank computes the same assignment from the annotations alone (no communication)."""
    weights = [float(w) for w in weights]
    order = sorted(range(len(weights)), key=lambda i: (-weights[i], i))
    shards = [[] for _ in range(int(world_size))]
    for k, page in enumerate(order):
        rnd, pos = divmod(k, int(world_size))
        shards[pos if rnd % 2 == 0 else int(world_size) - 1 - pos].append(page)
    return [sorted(s) for s in shards]


def page_cost(annotations_group, image_hw, anchor_params=None, pyramid_levels=None, fixed=None):
    """K1's cost of each page, for :func:`balanced_shards`: the kernel does a fixed amount of work per anchor plus one
    IoU per (anchor, GT table) pair that overlaps, so the estimate is ``fixed + pairs`` with
    ``pairs = sum over tables, levels, anchor types of (w + aw)(h + ah) / stride^2`` (the number of anchor centres whose
    box can touch the table; closed form, no anchors are generated).  ``fixed`` defaults to the measured ratio
    (``profiles/k1_page_cost.py``): the per-anchor work of a page equals ~PAIRS_PER_FIXED overlapping pairs.
    Deterministic from the annotations alone, so every rank derives the same sharding without communication."""
    from . import anchors as _anchors
    ap = anchor_params or _anchors.AnchorParameters_default
    levels = pyramid_levels or [3, 4, 5, 6, 7]
    ratios, scales = np.asarray(ap.ratios, dtype=np.float64), np.asarray(ap.scales, dtype=np.float64)
    aw, ah, inv_s2 = [], [], []
    for size, stride in zip(ap.sizes, ap.strides):
        base = _anchors.generate_anchors(size, ap.ratios, ap.scales)
        aw.append(base[:, 2] - base[:, 0])
        ah.append(base[:, 3] - base[:, 1])
        inv_s2.append(np.full(base.shape[0], 1.0 / (float(stride) * float(stride))))
    aw, ah, inv_s2 = np.concatenate(aw), np.concatenate(ah), np.concatenate(inv_s2)
    H, W = float(image_hw[0]), float(image_hw[1])
    if fixed is None:
        n_anchors = sum(len(ratios) * len(scales) * (-(-int(H) // s)) * (-(-int(W) // s)) for s in ap.strides[:len(levels)])
        fixed = PAIRS_PER_FIXED * n_anchors
    out = []
    for ann in annotations_group:
        bb = np.asarray(ann['bboxes'], dtype=np.float64).reshape(-1, 4) if len(ann['bboxes']) else np.zeros((0, 4))
        w = np.clip(bb[:, 2] - bb[:, 0], 0, None)[:, None]
        h = np.clip(bb[:, 3] - bb[:, 1], 0, None)[:, None]
        pairs = (np.minimum(w + aw, W + aw) * np.minimum(h + ah, H + ah) * inv_s2).sum()
        out.append(float(fixed + pairs))
    return out


# K1 on 16 copies of one 800x1333 page: 35.9 us + 3.68e-5 us per overlapping (anchor, table) pair of the page (32 pages,
# rms error 1.4 us; profiles/k1_page_cost.py, round-2 kernel: profiles/r2/r2k_k1_page_cost.log), i.e. the fixed per-anchor work
# of a page equals 4.87 pairs per anchor (round-1 kernel: 43.3 us, 5.9)
PAIRS_PER_FIXED = 4.87


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


class PeerCounter(object):
    """The positive-count exchange over NVLink peer memory (C-ABI ``rn_peer_*``, ``csrc/peer_box.cu``): every
    rank of the node owns a mailbox, mapped into the others through CUDA IPC; ``publish`` (one tiny kernel after
    K1) stores this rank's count into every mailbox, and K2 -- called with ``box`` as its normaliser and the
    ``RN_LOSS_NPOS_PEER_BOX`` flag -- waits in its prologue, on local memory, for all counts of the step and adds
    them in rank order.  Replaces the NCCL all-reduce between K1 and K2 (one launch + ~10-20 us per step) and
    keeps the whole step capturable in CUDA graphs.

    ``PeerCounter.create()`` returns ``None`` when there is one rank, when the ranks are not all on one node /
    GPU peer access is unavailable, or when ``RN_B200_PEER_BOX=0``; callers then fall back to
    :func:`global_positive_count` (``all_reduce``)."""

    def __init__(self, rank, world, box, peers):
        self.rank, self.world, self.box, self.peers = rank, world, box, peers
        self._arr = (ctypes.c_void_p * world)(*peers)
        self.bound = None
        # the device-side sends of a loss launch (loss sums, fused count publish) need the peers' pointers in the mailbox
        _lib.check(_lib.load().rn_peer_box_connect(ctypes.c_void_p(self.box), self._arr, rank, world), "rn_peer_box_connect")
        # how long a device-side wait for a peer may last before it gives up (sticky error flag + NaN losses instead of a
        # hung GPU).  Rank skew of seconds is normal (checkpointing, evaluation callbacks, a stalled loader): default 30 s.
        self.set_timeout(float(os.environ.get("RN_B200_PEER_TIMEOUT_S", "30")))

    def set_timeout(self, seconds):
        _lib.check(_lib.load().rn_peer_box_set_timeout(ctypes.c_void_p(self.box), float(seconds)), "rn_peer_box_set_timeout")
        self.timeout_s = float(seconds)

    def timed_out(self, clear=True):
        """True when a device-side wait of this rank has timed out since the flag was last cleared (synchronous read)."""
        flag = ctypes.c_int(0)
        _lib.check(_lib.load().rn_peer_box_status(ctypes.c_void_p(self.box), ctypes.byref(flag), 1 if clear else 0), "rn_peer_box_status")
        return bool(flag.value)

    def check(self):
        """Raises :class:`RnError` if a peer did not publish within the timeout (the losses / gradients of that step are NaN
        and must be discarded).  Synchronous; call it wherever the step's losses are read on the host."""
        if self.timed_out(clear=True):
            raise _lib.RnError("peer mailbox: a rank did not publish its count / loss sums within %.1f s "
                               "(RN_B200_PEER_TIMEOUT_S); the step's losses and gradients are invalid" % self.timeout_s)

    @classmethod
    def create(cls, group=None):
        rank, world = (dist.get_rank(group), dist.get_world_size(group)) if (dist.is_available() and dist.is_initialized()) else (0, 1)
        mode = os.environ.get("RN_B200_PEER_BOX", "1")
        if not torch.cuda.is_available() or world > _lib.RN_MAX_WORLD or mode == "0":
            return None
        lib = _lib.load()
        if world < 2:
            if mode != "force":                             # "force": a one-rank mailbox, for profiling the path on one GPU
                return None
            solo, handle = ctypes.c_void_p(), ctypes.create_string_buffer(64)
            _lib.check(lib.rn_peer_box_create(1, ctypes.byref(solo), handle), "rn_peer_box_create")
            return cls(0, 1, solo.value, [solo.value])
        box = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        ok = lib.rn_peer_box_create(world, ctypes.byref(box), handle) == 0
        infos = [None] * world
        dist.all_gather_object(infos, (ok, bytes(handle.raw), os.uname().nodename), group=group)
        usable = all(i[0] for i in infos) and len(set(i[2] for i in infos)) == 1
        peers = [None] * world
        if usable:
            for r in range(world):
                if r == rank:
                    peers[r] = box.value
                    continue
                ptr = ctypes.c_void_p()
                hb = ctypes.create_string_buffer(infos[r][1], 64)
                if lib.rn_peer_box_open(hb, ctypes.byref(ptr)) != 0:
                    usable = False
                    break
                peers[r] = ptr.value
        flags = [None] * world
        dist.all_gather_object(flags, bool(usable), group=group)
        if not all(flags):                                  # every rank must take the same path
            for r, ptr in enumerate(peers):
                if ptr is not None and r != rank:
                    lib.rn_peer_box_close(ctypes.c_void_p(ptr))
            if ok:
                lib.rn_peer_box_destroy(box)
            return None
        return cls(rank, world, box.value, peers)

    def bind(self, value, value_odd=None):
        """Prepare the FUSED publish: record the peers' mailboxes, this rank and the 1-float device tensor ``value``
        (K1's positive count) in the local mailbox, so that the loss kernel launched with ``peer_publish=True`` sends
        the count itself -- no publish launch between K1 and K2.  With ``value_odd`` the launch completing an odd
        step sends that tensor instead (double-buffered targets, see ``pipeline``).  Synchronous; the tensors must stay
        allocated and the stream using the mailbox must be idle."""
        odd = value if value_odd is None else value_odd
        _lib.check(_lib.load().rn_peer_box_bind(ctypes.c_void_p(self.box), self._arr, self.rank, self.world, _lib.ptr(value),
                                                _lib.ptr(odd)), "rn_peer_box_bind")
        self.bound = (value, odd)

    def steps(self):
        """Steps this rank has completed / published so far (synchronous read of the device counter)."""
        out = ctypes.c_ulonglong(0)
        _lib.check(_lib.load().rn_peer_box_step(ctypes.c_void_p(self.box), ctypes.byref(out)), "rn_peer_box_step")
        return int(out.value)

    def publish(self, value, device=None):
        """Enqueue the publication of the 1-float device tensor ``value`` (this rank's positive count)."""
        _lib.check(_lib.load().rn_peer_publish(_lib.ptr(value), ctypes.c_void_p(self.box), self._arr, self.rank, self.world,
                                               _lib.stream_ptr(device)), "rn_peer_publish")

    def close(self):
        lib = _lib.load()
        for r, ptr in enumerate(self.peers):
            if ptr is not None and r != self.rank:
                lib.rn_peer_box_close(ctypes.c_void_p(ptr))
        if self.box is not None:
            lib.rn_peer_box_destroy(ctypes.c_void_p(self.box))
        self.peers, self.box = [], None
