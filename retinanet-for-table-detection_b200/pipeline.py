"""The training-target step as one replayable unit: K1 (targets) -> count exchange -> K2 (losses fwd+bwd).

``TargetLossStep`` owns static device buffers for one batch shape and captures the kernel launches into CUDA
graphs (the kernels are tens of microseconds; launch overhead would otherwise dominate):

* ``run()``            the whole step as ONE graph launch (K1 [+ publish] + K2); with ``events`` the two halves are
                       replayed separately so that they can be timed;
* ``run_from_host()``  inputs in (pinned) host memory: the classification tensor is copied on a copy stream while the
                       GT list is packed and K1 runs; the regression rows of the positive anchors can be read in
                       place from the host buffer (``gather_reg_from_host``);
* ``run_pipelined()``  K1 of the NEXT batch ahead of (or, ``overlap=True``, concurrently with) K2 of the current one,
                       double-buffered targets.

With several ranks the positive count is exchanged through the NVLink peer mailbox
(``distributed.PeerCounter``) -- graph-capturable, no NCCL call -- or, where peers cannot map each other's memory,
all-reduced between the two halves.

This is host plumbing around the C-ABI; it adds no arithmetic of its own.
"""
import os

import numpy as np
import torch

from . import _lib
from . import anchors as _anchors
from . import distributed as _dist
from . import losses as _losses


class TargetLossStep(object):
    def __init__(self, image_shape, batch, gmax, num_classes, anchor_params=None, pyramid_levels=None,
                 negative_overlap=0.4, positive_overlap=0.5, alpha=0.25, gamma=2.0, sigma=3.0, bce="tf2",
                 use_graph=True, device=None, shared_state=True, peer_box=True, sparse_targets=False):
        _lib.require_cuda()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.spec = _anchors.make_spec(image_shape, pyramid_levels, anchor_params, None)
        self.B, self.G, self.C, self.N = int(batch), max(1, int(gmax)), int(num_classes), self.spec.num_anchors
        self.neg, self.pos = negative_overlap, positive_overlap
        # sparse_targets (extension): K1 writes the regression rows of positive anchors only -- the only ones K2 reads when it
        # takes the anchor state from the label tensor (shared_state); y_reg is then NOT the reference's full tensor
        self.sparse_targets = bool(sparse_targets)
        if self.sparse_targets and not (shared_state and int(num_classes) == 1):
            raise ValueError("sparse_targets needs shared_state=True and one class")
        # both target tensors come from K1 in the same step, so their state columns are identical
        self.loss_kw = dict(alpha=alpha, gamma=gamma, sigma=sigma, bce=bce, shared_state=shared_state)
        d, B, G, N, C = self.device, self.B, self.G, self.N, self.C
        # one staging block: boxes f64 | labels i32 | counts i32 | img_hw i32  (same layout as upload_annotations) | page order i32
        self.nb, self.nl, self.nc, self.ni = B * G * 32, B * G * 4, B * 4, B * 8
        total = self.nb + self.nl + self.nc + self.ni + B * 4
        self.gt_host = torch.zeros(total, dtype=torch.uint8, pin_memory=True)
        self.gt_dev = torch.zeros(total, dtype=torch.uint8, device=d)
        hv = self.gt_host.numpy()
        self._gt_views = (hv[:self.nb].view(np.float64).reshape(B, G, 4),
                          hv[self.nb:self.nb + self.nl].view(np.int32).reshape(B, G),
                          hv[self.nb + self.nl:self.nb + self.nl + self.nc].view(np.int32),
                          hv[self.nb + self.nl + self.nc:self.nb + self.nl + self.nc + self.ni].view(np.int32).reshape(B, 2))
        self._order_view = hv[self.nb + self.nl + self.nc + self.ni:].view(np.int32)
        self._order_view[:] = np.arange(B, dtype=np.int32)
        o = 0
        self.d_boxes = self.gt_dev[o:o + self.nb].view(torch.float64).view(B, G, 4); o += self.nb
        self.d_labels = self.gt_dev[o:o + self.nl].view(torch.int32).view(B, G); o += self.nl
        self.d_counts = self.gt_dev[o:o + self.nc].view(torch.int32); o += self.nc
        self.d_hw = self.gt_dev[o:o + self.ni].view(torch.int32).view(B, 2); o += self.ni
        self.d_order = self.gt_dev[o:o + B * 4].view(torch.int32)      # K1 starts the heaviest pages first (anchors.page_launch_order)
        self.d_order.copy_(torch.arange(B, dtype=torch.int32))
        self.cls_pred = torch.zeros((B, N, C), dtype=torch.float32, device=d)
        self.reg_pred = torch.zeros((B, N, 4), dtype=torch.float32, device=d)
        self.y_reg = (torch.zeros if self.sparse_targets else torch.empty)((B, N, 5), dtype=torch.float32, device=d)
        self.y_cls = torch.empty((B, N, C + 1), dtype=torch.float32, device=d)
        # per-page counts (B int32) and the batch total (1 float32) in one allocation (cleared by the reset kernel K1 is launched behind)
        self._counts = torch.zeros(B + 1, dtype=torch.int32, device=d)
        self.npos = self._counts[:B]
        self.npos_total = self._counts[B:].view(torch.float32)
        self.losses = torch.zeros(3, dtype=torch.float32, device=d)
        self.grad_cls = torch.empty_like(self.cls_pred)
        self.grad_reg = torch.empty_like(self.reg_pred)
        self.loss_ws = torch.zeros(int(_lib.load().rn_loss_workspace_bytes()), dtype=torch.uint8, device=d)
        self.use_graph = use_graph
        self._graphs = None
        self._fused = None
        self.kernel_launches_per_step = 3      # the counter reset K1 is launched behind, K1, K2 (NCCL is not ours)
        # run_from_host(): copy stream, per-chunk events, per-chunk loss rows
        rank, world = _dist.world()
        self.peer = _dist.PeerCounter.create() if peer_box else None   # NVLink mailbox; None with one rank / no peer access
        # fused publish: K2 sends this rank's count itself (CTA 0, P2P stores) before waiting for the others' --
        # no publish launch between K1 and K2 (RN_B200_PEER_FUSED=0: the separate rn_peer_publish kernel)
        self.peer_fused = self.peer is not None and os.environ.get("RN_B200_PEER_FUSED", "1") != "0"
        self._peer_mode = None                 # what the mailbox's value pointers are bound for: 'inorder' / 'pipe'
        if self.peer_fused:
            self._bind_inorder()
        else:
            self.kernel_launches_per_step += 1 if self.peer is not None else 0
        # run_pipelined(): second set of target buffers, graphs per buffer
        self._pipe = None
        self._copy_stream = None
        self._chunk_losses = None
        self._chunk_events = None
        self._losses_host = None
        self._done_event = None
        self._gt_event = None                  # the last H2D copy of the pinned GT block (guards its reuse by the host)

    def _bind_inorder(self):
        """Fused publish, in-order schedule: every step's loss launch sends ``npos_total``."""
        if self._peer_mode != 'inorder':
            torch.cuda.synchronize(self.device)
            self.peer.bind(self.npos_total)
            self._peer_mode = 'inorder'

    # ---- inputs ---------------------------------------------------------------------------------
    def load_annotations(self, image_group, annotations_group):
        """Pack the ragged GT list (reference format) and copy it to the static device block (async).  The steps are
        asynchronous graph replays, so the host may run ahead: the pinned block is only rewritten once the previous copy
        out of it has executed."""
        if self._gt_event is not None:
            self._gt_event.synchronize()
        _anchors.pack_annotations(image_group, annotations_group, self.C, out=self._gt_views)   # straight into the pinned block
        _anchors.page_launch_order(self._gt_views[0], out=self._order_view)
        self.gt_dev.copy_(self.gt_host, non_blocking=True)
        if self._gt_event is None:
            self._gt_event = torch.cuda.Event()
        self._gt_event.record(torch.cuda.current_stream(self.device))
        return self.gt_host.numel()

    def load_predictions(self, cls_pred, reg_pred):
        """Head outputs (host pinned or device tensors) into the static buffers (async)."""
        self.cls_pred.copy_(cls_pred, non_blocking=True)
        self.reg_pred.copy_(reg_pred, non_blocking=True)

    # ---- the two halves ----------------------------------------------------------------------------
    def _targets(self):
        _anchors.anchor_targets_device(self.spec, self.d_boxes, self.d_labels, self.d_counts, self.d_hw, self.C,
                                       self.neg, self.pos, out=(self.y_reg, self.y_cls), npos_total=self.npos_total,
                                       npos_out=self.npos, page_order=self.d_order, sparse_regression=self.sparse_targets)
        if self.peer is not None and not self.peer_fused:
            self.peer.publish(self.npos_total, self.device)     # this rank's count -> every rank's mailbox

    def _exchange(self, npos_total=None):
        """Several ranks without the peer mailbox: the count is all-reduced between the two kernels."""
        if self.peer is None and _dist.world()[1] > 1:
            torch.distributed.all_reduce(self.npos_total if npos_total is None else npos_total)

    def _reduce_losses(self):
        """Several ranks without the peer mailbox: every rank's sums are already divided by the global normaliser, so
        the merged batch's losses are their plain sum (with the mailbox K2's last CTA does this itself)."""
        if self.peer is None and _dist.world()[1] > 1:
            torch.distributed.all_reduce(self.losses[:2])

    def _losses(self):
        _losses.detection_losses(self.y_reg, self.y_cls, self.reg_pred, self.cls_pred, normalizer=self.npos_total,
                                 out=(self.losses, self.grad_cls, self.grad_reg), workspace=self.loss_ws,
                                 peer_box=self.peer, peer_publish=self.peer_fused, peer_losses=self.peer is not None,
                                 **self.loss_kw)

    def _capture(self, fn):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        return g

    def _build_graphs(self):
        # warm up on a side stream (allocations, lazy module loading) before capture
        s = torch.cuda.Stream(self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            self._targets()
            self._losses()
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        self._graphs = (self._capture(self._targets), self._capture(self._losses))
        # the whole step as ONE graph (one launch) when nothing has to happen between the halves on the host:
        # one rank, or several ranks exchanging the count through the peer mailbox
        self._fused = None
        if self.peer is not None or _dist.world()[1] == 1:
            def whole():
                self._targets()
                self._losses()
            self._fused = self._capture(whole)

    # ---- pipelined schedule for several ranks ------------------------------------------------------------
    def _pipe_setup(self):
        """Targets are double-buffered: K1 (+ publish) of the NEXT batch is enqueued ahead of -- and, in the overlapped
        form, concurrently with -- K2 of the current one (the reference's generator threads likewise prepare the next
        batch's targets during a train step).  With several ranks K2 reads the mailbox with lag 1, so the count
        exchange and the skew between ranks leave the critical path."""
        d = self.device
        second = torch.zeros(self.B + 1, dtype=torch.int32, device=d)
        bufs = [(self.y_reg, self.y_cls, self.npos, self.npos_total),
                ((torch.zeros_like if self.sparse_targets else torch.empty_like)(self.y_reg), torch.empty_like(self.y_cls), second[:self.B], second[self.B:].view(torch.float32))]

        no_box = self.peer is None and _dist.world()[1] > 1    # several ranks, no mailbox: all_reduce between the graphs

        def targets(i):
            y_reg, y_cls, npos, npos_total = bufs[i]
            _anchors.anchor_targets_device(self.spec, self.d_boxes, self.d_labels, self.d_counts, self.d_hw, self.C,
                                           self.neg, self.pos, out=(y_reg, y_cls), npos_total=npos_total, npos_out=npos,
                                           page_order=self.d_order, sparse_regression=self.sparse_targets)
            if self.peer is not None and not self.peer_fused:
                self.peer.publish(npos_total, self.device)

        def losses(i):
            y_reg, y_cls, npos, npos_total = bufs[i]
            # fused publish: the loss launch of batch s sends the count of batch s itself (from the buffer bound for the
            # step's parity) and waits for the other ranks' -- while K1 of batch s+1 runs beside it, so the exchange and
            # the skew between ranks are hidden behind the longer kernel.  Separate publish kernel: mailbox lag 1.
            _losses.detection_losses(y_reg, y_cls, self.reg_pred, self.cls_pred, normalizer=npos_total,
                                     out=(self.losses, self.grad_cls, self.grad_reg), workspace=self.loss_ws,
                                     peer_box=self.peer, peer_lag=0 if self.peer_fused else 1,
                                     peer_publish=self.peer_fused, peer_losses=self.peer is not None, **self.loss_kw)
        if self.peer_fused:
            # the loss launch completing step t sends value[t & 1]: bind the two count buffers so that the warm-up's
            # losses(0) (step t0 = steps() + 1) reads buffer 0 and the following steps alternate 1, 0, 1, ...
            torch.cuda.synchronize(d)
            t0 = self.peer.steps() + 1
            pair = (bufs[0][3], bufs[1][3]) if t0 % 2 == 0 else (bufs[1][3], bufs[0][3])
            self.peer.bind(*pair)
            self._peer_mode = 'pipe'
        # warm-up outside capture keeps every rank's publish count equal: one full in-order-equivalent round
        s = torch.cuda.Stream(d)
        s.wait_stream(torch.cuda.current_stream(d))
        with torch.cuda.stream(s):
            targets(0)
            self._exchange(bufs[0][3])
            if self.peer_fused:
                losses(0)
                targets(1)
            else:
                targets(1)
                self._exchange(bufs[1][3])
                losses(0)
            self._reduce_losses()
        torch.cuda.current_stream(d).wait_stream(s)
        torch.cuda.synchronize(d)
        # K2's branch gets the higher stream priority: its CTAs are placed first whenever an SM has room, so the short
        # HBM-bound kernel (and, with several ranks, the exchange it carries) finishes inside the long issue-bound one
        # (60.9 -> 59.7 us per step; profiles/r2/r2v_overlap_share_experiment.log also holds what did NOT help: K2 on
        # 1-4 CTAs per SM, K1 capped at 64 / 56 registers so that a K2 CTA fits beside three of K1's)
        side = (torch.cuda.Stream(d), torch.cuda.Stream(d, priority=-1 if os.environ.get("RN_B200_K2_PRIORITY", "1") != "0" else 0))

        def both(nxt, cur):
            """K1 (+ publish) of the next batch and K2 of the current one on two streams: captured, they become two
            parallel branches of one graph -- K1 is instruction-issue bound, K2 HBM-bound, so they share an SM well."""
            main = torch.cuda.current_stream(d)
            for st in side:
                st.wait_stream(main)
            with torch.cuda.stream(side[0]):
                targets(nxt)
            with torch.cuda.stream(side[1]):
                losses(cur)
            for st in side:
                main.wait_stream(st)

        if self.use_graph:
            ga = [self._capture(lambda i=i: targets(i)) for i in range(2)]
            gb = [self._capture(lambda i=i: losses(i)) for i in range(2)]
            # the two-branch graph is only legal when nothing has to happen between K1 and K2 on the host
            gc = None if no_box else [self._capture(lambda i=i: both(i, 1 - i)) for i in range(2)]   # index = the NEXT batch's buffer
            run_a, run_b = (lambda i: ga[i].replay()), (lambda i: gb[i].replay())
            run_ab = None if no_box else (lambda i: gc[i].replay())
        else:
            run_a, run_b, run_ab = targets, losses, (None if no_box else (lambda i: both(i, 1 - i)))
        if no_box:
            # the count of the NEXT batch is all-reduced right behind its K1, one call before the K2 that divides by it
            plain_a, plain_b = run_a, run_b
            run_a = lambda i: (plain_a(i), self._exchange(bufs[i][3]))
            run_b = lambda i: (plain_b(i), self._reduce_losses())
        # after the warm-up the latest published batch sits in buffer 1 (lag 0); prime: it becomes "current"
        self._pipe = dict(bufs=bufs, run_a=run_a, run_b=run_b, run_ab=run_ab, cur=1)

    def run_pipelined(self, events=None, overlap=False):
        """One step of the pipelined schedule: enqueue K1 (+ publish) for the NEXT batch (the annotations currently
        loaded), then K2 for the batch whose targets were produced by the previous call.  ``losses`` / ``grad_*``
        then refer to that previous batch; ``targets_of_losses()`` returns its target tensors.
        ``overlap=True``: the two kernels run concurrently on two streams (one graph launch per step)."""
        if overlap and self.peer is not None and not self.peer_fused:
            raise ValueError("overlap=True with several ranks needs the fused publish (the separate publish kernel would "
                             "bump the step counter while K2 of the previous batch reads it)")
        if overlap and self.peer is None and _dist.world()[1] > 1:
            raise ValueError("overlap=True with several ranks needs the peer mailbox: without it the positive count is "
                             "all-reduced on the host side between K1 and K2, which a two-branch graph cannot contain")
        if self._pipe is None:
            self._pipe_setup()
        pp = self._pipe
        if self.peer_fused and self._peer_mode != 'pipe':
            # in-order steps ran in between: bind the count buffers again so that the next loss launch (step t) reads
            # the buffer of the batch it consumes
            torch.cuda.synchronize(self.device)
            t = self.peer.steps() + 1
            mine, other = pp['bufs'][pp['cur']][3], pp['bufs'][1 - pp['cur']][3]
            self.peer.bind(*((mine, other) if t % 2 == 0 else (other, mine)))
            self._peer_mode = 'pipe'
        cur, nxt = pp['cur'], 1 - pp['cur']
        if overlap:
            if events is not None:
                events[0].record()
            pp['run_ab'](nxt)                               # K1(s+1) [+ publish(s+1)]  ||  K2(s)
            if events is not None:
                events[1].record()
                events[2].record()
        else:
            if events is not None:
                events[0].record()
            pp['run_a'](nxt)                                # K1(s+1) + publish(s+1)
            if events is not None:
                events[1].record()
            pp['run_b'](cur)                                # K2(s), mailbox lag 1
            if events is not None:
                events[2].record()
        pp['cur'] = nxt
        pp['done'] = cur
        return self.losses

    def targets_of_losses(self):
        """(y_reg, y_cls) of the batch the last run_pipelined() computed the losses for."""
        y_reg, y_cls = self._pipe['bufs'][self._pipe['done']][:2]
        return y_reg, y_cls

    # ---- host inputs, copies overlapped with the kernels -----------------------------------------------
    def run_from_host(self, image_group, annotations_group, cls_host, reg_host, chunks=4, gather_reg_from_host=False):
        """One step whose inputs live in (pinned) HOST memory: the head outputs are copied page-chunk by
        page-chunk on a copy stream while K1 runs on the compute stream (K1 needs only the GT block), and K2
        is launched per chunk as soon as that chunk's predictions have landed.  Every K2 launch uses the
        batch-global normaliser, so per-chunk losses add up to the batch loss (the loss is a sum over anchors)
        and the gradients are identical to a single launch.  Returns the three floats
        ``[focal, smooth_l1, normaliser]`` on the host (one D2H copy of ``chunks`` x 12 bytes); the device tensor
        ``losses`` of :meth:`run` is not touched.

        ``gather_reg_from_host=True`` (``reg_host`` must be PINNED): the (B, N, 4) regression predictions are not
        copied at all.  The smooth-L1 loss only ever reads the rows of positive anchors (the reference gathers
        exactly those, ``model/losses.py:72-74``; ~0.1 % of the rows), so K2 fetches those 16-byte rows straight
        from the pinned host buffer over PCIe (unified addressing): 51 MB of the step's 64 MB never cross the
        bus.  Losses and gradients are bit-identical to the copying path."""
        self._enqueue_from_host(image_group, annotations_group, cls_host, reg_host, chunks, gather_reg_from_host)
        return self._finish_from_host()

    def _enqueue_from_host(self, image_group, annotations_group, cls_host, reg_host, chunks=4, gather_reg_from_host=False,
                           wait_compute=True):
        """Everything of :meth:`run_from_host` except the final wait: copies, K1, K2 per chunk and the D2H copy of the
        loss rows are enqueued, ``_done_event`` is recorded behind them.  (``HostStepPipeline`` keeps several steps
        in flight this way.)"""
        dev = self.device
        if self.peer_fused:
            self._bind_inorder()
        chunks = max(1, min(int(chunks), self.B))
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(dev)
        if self._chunk_losses is None or self._chunk_losses.shape[0] != chunks:
            self._chunk_losses = torch.zeros((chunks, 3), dtype=torch.float32, device=dev)
            self._losses_host = torch.zeros((chunks, 3), dtype=torch.float32).pin_memory()
            self._chunk_events = [torch.cuda.Event() for _ in range(chunks)]
            self._done_event = torch.cuda.Event()
        if self.use_graph and self._graphs is None:
            self._build_graphs()
        if gather_reg_from_host and not (reg_host.is_pinned() and self.loss_kw.get("shared_state") and self.C == 1):
            raise ValueError("gather_reg_from_host needs a pinned reg_host and the C == 1 shared-state loss path "
                             "(the only one that reads regression rows of positive anchors only)")
        compute = torch.cuda.current_stream(dev)
        if wait_compute:                                    # HostStepPipeline has already waited for this slot's last step
            self._copy_stream.wait_stream(compute)          # the previous step's K2 has consumed the buffers
        # the 64 MB of head outputs go first: the PCIe link is the bottleneck of this step, so it starts before
        # the (Python) GT packing, which then runs on the CPU while the copies are in flight
        bounds = [(self.B * i) // chunks for i in range(chunks + 1)]
        with torch.cuda.stream(self._copy_stream):
            for i in range(chunks):
                lo, hi = bounds[i], bounds[i + 1]
                sl = (lambda t: t) if chunks == 1 else (lambda t: t[lo:hi])
                sl(self.cls_pred).copy_(sl(cls_host), non_blocking=True)
                if not gather_reg_from_host:
                    sl(self.reg_pred).copy_(sl(reg_host), non_blocking=True)
                self._chunk_events[i].record(self._copy_stream)
        self.load_annotations(image_group, annotations_group)
        if self.use_graph:
            self._graphs[0].replay()
        else:
            self._targets()
        self._exchange()
        reg_src = reg_host if gather_reg_from_host else self.reg_pred
        for i in range(chunks):
            lo, hi = bounds[i], bounds[i + 1]
            compute.wait_event(self._chunk_events[i])
            sl = (lambda t: t) if chunks == 1 else (lambda t: t[lo:hi])     # one chunk: no views to build
            _losses.detection_losses(sl(self.y_reg), sl(self.y_cls), sl(reg_src), sl(self.cls_pred),
                                     normalizer=self.npos_total,
                                     out=(self._chunk_losses[i], sl(self.grad_cls), sl(self.grad_reg)),
                                     workspace=self.loss_ws, peer_box=self.peer,
                                     peer_publish=self.peer_fused and i == 0,     # the step's first launch sends the count
                                     peer_losses=self.peer is not None and chunks == 1,   # one launch per step: sums via the mailbox
                                     **self.loss_kw)
        if _dist.world()[1] > 1 and not (self.peer is not None and chunks == 1):
            merged = self._chunk_losses[:, :2].contiguous()                # page chunks / no mailbox: the merged batch's losses
            torch.distributed.all_reduce(merged)
            self._chunk_losses[:, :2] = merged
        self._losses_host.copy_(self._chunk_losses, non_blocking=True)
        self._done_event.record(compute)

    def _finish_from_host(self):
        self._done_event.synchronize()
        out = self._losses_host.sum(dim=0)
        out[2] = self._losses_host[0, 2]                    # the normaliser is the same in every row
        if self.peer is not None and bool(torch.isnan(out[2])):
            self.peer.check()                               # a peer never published: raise instead of returning NaN losses
        return out

    def run(self, events=None):
        """One step on the current stream.  Results: ``losses`` [focal, smooth_l1, normaliser],
        ``grad_cls``, ``grad_reg``, ``y_reg``, ``y_cls`` (static tensors, overwritten every step).
        ``events``: optional 3 CUDA events recorded before K1, between K1 and K2 (after the all-reduce
        when there are several ranks) and after K2 -- used by the benchmark's roofline accounting."""
        if self.peer_fused:
            self._bind_inorder()
        if self.use_graph and self._graphs is None:
            self._build_graphs()
        if events is None and self.use_graph and self._fused is not None:
            self._fused.replay()                            # K1 (+ publish) + K2: one graph launch
            return self.losses
        if events is not None:
            events[0].record()
        if self.use_graph:
            self._graphs[0].replay()
        else:
            self._targets()
        self._exchange()
        if events is not None:
            events[1].record()
        if self.use_graph:
            self._graphs[1].replay()
        else:
            self._losses()
        self._reduce_losses()
        if events is not None:
            events[2].record()
        return self.losses

    def check(self):
        """Raises when a peer did not deliver its count / loss sums within the mailbox timeout (the step's losses are NaN).
        Synchronous."""
        if self.peer is not None:
            torch.cuda.synchronize(self.device)
            self.peer.check()


class HostStepPipeline(object):
    """Several training-target steps with HOST inputs in flight at once (what a training loop with a prefetching
    generator does): ``depth`` independent ``TargetLossStep`` slots -- own GT block, head-output buffers, targets,
    gradients and loss rows each -- share ONE copy stream (the PCIe link is the bottleneck; copies stay in order).

    ``submit()`` enqueues a step into the next slot and returns at once; ``result()`` waits for that step's loss
    rows.  While step s runs K1/K2, the copy engine is already moving the classification tensor of step s+1, so the
    step rate is the PCIe copy rate instead of copy + kernels + launch latency.  The arithmetic is the one of
    ``TargetLossStep.run_from_host`` (same kernels, same order per step): losses and gradients are bit-identical.
    With several ranks the kernels of all steps stay in order on one compute stream, so the count exchange behaves
    exactly as in the unpipelined schedule."""

    def __init__(self, image_shape, batch, gmax, num_classes, depth=2, device=None, **step_kw):
        _lib.require_cuda()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.depth = max(1, int(depth))
        self.slots = [TargetLossStep(image_shape, batch, gmax, num_classes, device=self.device, **step_kw)
                      for _ in range(self.depth)]
        # One rank: every slot has its own compute stream, so K1 of step s+1 is not held behind K2 of step s (which,
        # with ``gather_reg_from_host``, spends most of its time waiting for PCIe reads that compete with the bulk
        # copy).  Several ranks: one compute stream for all slots -- kernels of different steps stay in order, so a
        # K2 waiting in its prologue for the peers' counts can never occupy the SMs ahead of an earlier step's K1.
        one = _dist.world()[1] > 1
        self.compute = torch.cuda.Stream(self.device)
        self.streams = [self.compute if (one or k == 0) else torch.cuda.Stream(self.device) for k in range(self.depth)]
        copy = torch.cuda.Stream(self.device)
        for s in self.slots:
            s._copy_stream = copy
            if s.use_graph:
                s._build_graphs()
        self._busy = [False] * self.depth
        self._next = 0

    def submit(self, image_group, annotations_group, cls_host, reg_host, chunks=1, gather_reg_from_host=False):
        """Enqueue one step; returns the slot index to pass to :meth:`result`.  ``cls_host`` / ``reg_host`` must stay
        untouched until the step's result has been taken.  If the slot still holds an unfinished step, that step is
        waited for first (its result is dropped)."""
        k = self._next
        self._next = (k + 1) % self.depth
        slot = self.slots[k]
        if self._busy[k]:
            slot._finish_from_host()
        with torch.cuda.stream(self.streams[k]):
            # no device-side wait of the copy stream for the compute stream: the slot's previous step has finished (host
            # wait above), and waiting for the OTHER slots' kernels would serialise the copy behind them
            slot._enqueue_from_host(image_group, annotations_group, cls_host, reg_host, chunks, gather_reg_from_host,
                                    wait_compute=False)
        self._busy[k] = True
        return k

    def result(self, k):
        """``[focal, smooth_l1, normaliser]`` of the step submitted into slot ``k`` (waits for it); the slot's
        gradients / targets are ``slots[k].grad_cls`` / ``grad_reg`` / ``y_reg`` / ``y_cls`` until it is reused."""
        if not self._busy[k]:
            raise RuntimeError("slot %d holds no submitted step" % k)
        out = self.slots[k]._finish_from_host()
        self._busy[k] = False
        return out

    def drain(self):
        """Wait for every step in flight; returns their results oldest first."""
        order = [(self._next + i) % self.depth for i in range(self.depth)]
        return [self.result(k) for k in order if self._busy[k]]


class DetectionStep(object):
    """The inference tail (``layers.DetectionHead``: decode + clip + threshold + sort + NMS + merge) for one batch shape
    as a replayable unit, like ``TargetLossStep`` for the training half: static device buffers for the head outputs
    (``cls_pred`` (B,N,C), ``reg_pred`` (B,N,4)), the results (``boxes`` (B,M,4), ``scores`` (B,M), ``labels`` (B,M),
    ``indices`` (B,M), ``status`` (B)) and the workspace, and the three kernel launches captured into ONE CUDA graph.
    ``run()`` replays it on the current stream; ``load_predictions`` copies new head outputs in (async).  Results are
    bit-identical to calling the head directly."""

    def __init__(self, image_hw, batch, num_classes=1, head=None, use_graph=True, device=None, **head_kw):
        from . import layers as _layers
        _lib.require_cuda()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.head = head if head is not None else _layers.DetectionHead(**head_kw)
        self.hw = (int(image_hw[0]), int(image_hw[1]))
        self.B, self.C = int(batch), int(num_classes)
        self.N, M, d = self.head.spec_for(self.hw).num_anchors, int(self.head.max_detections), self.device
        self.cls_pred = torch.zeros((self.B, self.N, self.C), dtype=torch.float32, device=d)
        self.reg_pred = torch.zeros((self.B, self.N, 4), dtype=torch.float32, device=d)
        self.out = _layers.filter_outputs(self.B, M, d)
        self.boxes, self.scores, self.labels, self.indices, self.status = self.out
        self.workspace = torch.empty(max(256, self.head.workspace_bytes(self.B, self.hw, self.C)), dtype=torch.uint8, device=d)
        self.use_graph = use_graph
        self._graph = None
        self.kernel_launches_per_step = 4       # the workspace reset, k_threshold_keys, k_segment_nms, k_merge_topk

    def load_predictions(self, cls_pred, reg_pred):
        self.cls_pred.copy_(cls_pred, non_blocking=True)
        self.reg_pred.copy_(reg_pred, non_blocking=True)

    def _launch(self):
        self.head([(self.B,) + self.hw + (3,), self.reg_pred, self.cls_pred], check=False, out=self.out, workspace=self.workspace)

    def run(self):
        """One batch on the current stream; returns ``[boxes, scores, labels]`` (static tensors, overwritten per call)."""
        if not self.use_graph:
            self._launch()
        else:
            if self._graph is None:
                s = torch.cuda.Stream(self.device)
                s.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(s):
                    self._launch()                          # warm-up outside capture (lazy module loading, attributes)
                torch.cuda.current_stream(self.device).wait_stream(s)
                torch.cuda.synchronize(self.device)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._launch()
                self._graph = g
            self._graph.replay()
        return [self.boxes, self.scores, self.labels]

    def check(self):
        """Raises when a candidate slab overflowed in the last batch (only possible with ``cand_cap``); synchronises."""
        from . import layers as _layers
        _layers._raise_on_overflow(self.status, self.head.cand_cap)


class HostDetectionPipeline(object):
    """The inference tail (``layers.DetectionHead``) as a serving loop over HOST head outputs with several batches in
    flight: every slot has its own stream, device classification buffer and pinned result buffers.

    ``submit(regression_host, classification_host)`` (both PINNED float32) copies the classification tensor, runs the
    fused decode + filter kernels -- the regression rows of the candidates are read in place from the pinned buffer,
    see ``DetectionHead`` -- and copies the (B, max_detections, .) results back, all asynchronously; ``result(k)``
    waits and returns ``[boxes, scores, labels]`` as pinned host tensors (valid until the slot is reused).  While
    batch s is in its NMS kernel the copy of batch s+1 is under way, and two NMS kernels (one CTA per page and class)
    can share the GPU.  Results are bit-identical to calling the head directly."""

    def __init__(self, head, batch, image_hw, num_classes=1, depth=2, device=None):
        _lib.require_cuda()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.head, self.depth = head, max(1, int(depth))
        self.shape = (int(batch), int(image_hw[0]), int(image_hw[1]), 3)
        N, M, d = head.spec_for(image_hw).num_anchors, int(head.max_detections), self.device
        self.streams = [torch.cuda.Stream(d) for _ in range(self.depth)]
        self.cls_dev = [torch.empty((batch, N, num_classes), dtype=torch.float32, device=d) for _ in range(self.depth)]
        self.out = [[torch.empty((batch, M, 4), dtype=torch.float32).pin_memory(),
                     torch.empty((batch, M), dtype=torch.float32).pin_memory(),
                     torch.empty((batch, M), dtype=torch.int32).pin_memory()] for _ in range(self.depth)]
        self.events = [torch.cuda.Event() for _ in range(self.depth)]
        self._busy = [False] * self.depth
        self._next = 0

    def submit(self, regression_host, classification_host):
        k = self._next
        self._next = (k + 1) % self.depth
        if self._busy[k]:
            self.events[k].synchronize()
        if not (regression_host.is_pinned() and classification_host.is_pinned()):
            raise ValueError("HostDetectionPipeline needs pinned host tensors")
        with torch.cuda.stream(self.streams[k]):
            self.cls_dev[k].copy_(classification_host, non_blocking=True)
            res = self.head([self.shape, regression_host, self.cls_dev[k]])
            for dst, src in zip(self.out[k], res):
                dst.copy_(src, non_blocking=True)
            self.events[k].record(self.streams[k])
        self._busy[k] = True
        return k

    def result(self, k):
        if not self._busy[k]:
            raise RuntimeError("slot %d holds no submitted batch" % k)
        self.events[k].synchronize()
        self._busy[k] = False
        return self.out[k]

    def drain(self):
        order = [(self._next + i) % self.depth for i in range(self.depth)]
        return [[t.clone() for t in self.result(k)] for k in order if self._busy[k]]
