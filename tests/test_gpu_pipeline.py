"""GPU: TargetLossStep -- the graph-replayed resident step and the overlapped host-input step give the same
targets, gradients (bit for bit) and losses (the per-chunk loss sums are added on the host: 1e-6 relative)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("chunks", [1, 3, 4])
def test_run_from_host_matches_resident_run(rn, chunks):
    import synthetic
    from oracle import anchors_np as OA
    from oracle import losses_np as OL
    hw, B = (256, 320), 5
    anchors = rn.anchors_for_shape(hw + (3,))
    N = anchors.shape[0]
    imgs = [synthetic.PageShape(hw + (3,)) for _ in range(B)]
    anns = [synthetic.gt_for_page(2, i, hw=hw, gmax=6) for i in range(B)]
    cls, reg = synthetic.training_predictions(2, B, N, classes=1)
    cls_h, reg_h = torch.from_numpy(cls).pin_memory(), torch.from_numpy(reg).pin_memory()

    step = rn.pipeline.TargetLossStep(hw + (3,), B, 8, 1)
    step.load_annotations(imgs, anns)
    step.load_predictions(cls_h, reg_h)
    want = step.run().cpu().numpy().copy()
    gc, gr = step.grad_cls.clone(), step.grad_reg.clone()
    yr, yc = step.y_reg.clone(), step.y_cls.clone()

    step2 = rn.pipeline.TargetLossStep(hw + (3,), B, 8, 1)
    for _ in range(2):                                   # twice: buffers are reused across steps
        got = step2.run_from_host(imgs, anns, cls_h, reg_h, chunks=chunks).numpy()
    # regression rows of positive anchors fetched straight from the pinned host buffer: nothing changes
    step3 = rn.pipeline.TargetLossStep(hw + (3,), B, 8, 1)
    got3 = step3.run_from_host(imgs, anns, cls_h, reg_h, chunks=chunks, gather_reg_from_host=True).numpy()
    assert np.array_equal(got3, got) and torch.equal(step3.grad_cls, step2.grad_cls) and torch.equal(step3.grad_reg, step2.grad_reg)
    assert not step3.reg_pred.any()                      # the device copy of the regression tensor was never filled
    with pytest.raises(ValueError):
        step3.run_from_host(imgs, anns, cls_h, torch.from_numpy(reg), chunks=chunks, gather_reg_from_host=True)   # not pinned
    assert torch.equal(step2.y_reg, yr) and torch.equal(step2.y_cls, yc)
    assert torch.equal(step2.grad_cls, gc) and torch.equal(step2.grad_reg, gr)
    assert got[2] == want[2] and np.allclose(got[:2], want[:2], rtol=1e-6, atol=0)

    # and both agree with the oracle
    oreg, olab = OA.anchor_targets_bbox(OA.anchors_for_shape(hw + (3,)), imgs, anns, 1)
    assert yr.cpu().numpy().tobytes() == oreg.tobytes() and yc.cpu().numpy().tobytes() == olab.tobytes()
    wf, ws = OL.focal()(olab, cls), OL.smooth_l1()(oreg, reg)
    assert abs(got[0] - wf) <= 1e-5 * abs(wf) and abs(got[1] - ws) <= 1e-5 * abs(ws)


@pytest.mark.parametrize("gather", [False, True])
def test_host_step_pipeline_matches_run_from_host(rn, gather):
    """Two steps in flight (different batches in the two slots): every step's losses, gradients and targets are
    bit-identical to the unpipelined host-input step on the same batch."""
    import synthetic
    hw, B = (256, 320), 4
    N = rn.anchors_for_shape(hw + (3,)).shape[0]
    imgs = [synthetic.PageShape(hw + (3,)) for _ in range(B)]
    batches = []
    for s in range(5):
        anns = [synthetic.gt_for_page(2, 10 * s + i, hw=hw, gmax=6) for i in range(B)]
        cls, reg = synthetic.training_predictions(20 + s, B, N, classes=1)
        batches.append((anns, torch.from_numpy(cls).pin_memory(), torch.from_numpy(reg).pin_memory()))
    ref = rn.pipeline.TargetLossStep(hw + (3,), B, 8, 1)
    want = []
    for anns, c, r in batches:
        l = ref.run_from_host(imgs, anns, c, r, chunks=1, gather_reg_from_host=gather).numpy().copy()
        want.append((l, ref.grad_cls.clone(), ref.grad_reg.clone(), ref.y_reg.clone(), ref.y_cls.clone()))

    pipe = rn.pipeline.HostStepPipeline(hw + (3,), B, 8, 1, depth=2)
    with pytest.raises(RuntimeError):
        pipe.result(0)                                    # nothing submitted yet
    pending = []
    for s, (anns, c, r) in enumerate(batches):
        pending.append((s, pipe.submit(imgs, anns, c, r, chunks=1, gather_reg_from_host=gather)))
        if len(pending) == 2:                             # take the older step's result while the newer one runs
            t, k = pending.pop(0)
            got = pipe.result(k).numpy()
            slot = pipe.slots[k]
            assert np.array_equal(got, want[t][0])
            assert torch.equal(slot.grad_cls, want[t][1]) and torch.equal(slot.grad_reg, want[t][2])
            assert torch.equal(slot.y_reg, want[t][3]) and torch.equal(slot.y_cls, want[t][4])
    rest = pipe.drain()
    assert len(rest) == 1 and np.array_equal(rest[0].numpy(), want[-1][0])
    assert pipe.drain() == []


def test_host_detection_pipeline_matches_direct_head(rn):
    """Batches in flight on two streams give the detections of the direct DetectionHead call, bit for bit."""
    import synthetic
    hw, B = (256, 320), 3
    anchors = np.asarray(rn.anchors_for_shape(hw + (3,)))
    head = rn.DetectionHead()
    batches, want = [], []
    for s in range(4):
        anns = [synthetic.gt_for_page(3, 7 * s + i, hw=hw, gmax=5) for i in range(B)]
        cls, reg = synthetic.inference_predictions(3, B, anchors, anns, classes=1, first_page=5 * s)
        cls_h, reg_h = torch.from_numpy(cls).pin_memory(), torch.from_numpy(reg).pin_memory()
        batches.append((reg_h, cls_h))
        want.append([t.cpu() for t in head([(B,) + hw + (3,), reg_h.cuda(), cls_h.cuda()])])
    pipe = rn.pipeline.HostDetectionPipeline(head, B, hw, depth=2)
    pend = []
    for s, (reg_h, cls_h) in enumerate(batches):
        pend.append((s, pipe.submit(reg_h, cls_h)))
        if len(pend) == 2:
            t, k = pend.pop(0)
            got = pipe.result(k)
            assert all(torch.equal(g, w) for g, w in zip(got, want[t]))
            assert int((got[1] >= 0).sum()) > 0
    rest = pipe.drain()
    assert len(rest) == 1 and all(torch.equal(g, w) for g, w in zip(rest[0], want[-1]))
    with pytest.raises(ValueError):
        pipe.submit(torch.zeros(1), batches[0][1])           # not pinned


@pytest.mark.parametrize("fused", ["1", "0"])
def test_peer_mailbox_one_rank(rn, monkeypatch, fused):
    """The count exchange through the peer mailbox on ONE GPU (a one-rank mailbox: publish to self, wait on self), with
    the publish fused into K2 and as a separate kernel: same losses / gradients as without a mailbox, over several
    steps (the slots rotate), through the one-graph step, the split graphs and the chunked host-input step."""
    import synthetic
    hw, B = (256, 320), 4
    N = rn.anchors_for_shape(hw + (3,)).shape[0]
    imgs = [synthetic.PageShape(hw + (3,)) for _ in range(B)]
    cls, reg = synthetic.training_predictions(5, B, N, classes=1)
    cls_h, reg_h = torch.from_numpy(cls).pin_memory(), torch.from_numpy(reg).pin_memory()
    plain = rn.pipeline.TargetLossStep(hw + (3,), B, 8, 1)
    assert plain.peer is None
    monkeypatch.setenv("RN_B200_PEER_BOX", "force")
    monkeypatch.setenv("RN_B200_PEER_FUSED", fused)
    boxed = rn.pipeline.TargetLossStep(hw + (3,), B, 8, 1)
    assert boxed.peer is not None and boxed.peer_fused == (fused == "1")
    assert boxed.kernel_launches_per_step == (3 if fused == "1" else 4)     # reset + K1 + K2 (+ the separate publish)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for s in range(7):                                    # > RN_PEER_SLOTS steps
        anns = [synthetic.gt_for_page(2, 10 * s + i, hw=hw, gmax=6) for i in range(B)]
        for st in (plain, boxed):
            st.load_annotations(imgs, anns)
            st.load_predictions(cls_h, reg_h)
        want = plain.run().cpu().numpy().copy()
        if s % 3 == 0:
            got = boxed.run().cpu().numpy()
        elif s % 3 == 1:
            got = boxed.run(events=evs).cpu().numpy()
        else:
            got = boxed.run_from_host(imgs, anns, cls_h, reg_h, chunks=3).numpy()
        assert got[2] == want[2] and want[2] >= 1
        assert np.allclose(got[:2], want[:2], rtol=1e-6, atol=0)
        assert torch.equal(boxed.grad_cls, plain.grad_cls) and torch.equal(boxed.grad_reg, plain.grad_reg)


@pytest.mark.parametrize("fused", ["1", "0"])
def test_peer_mailbox_pipelined_schedule(rn, monkeypatch, fused):
    """run_pipelined (K1 of the next batch ahead of / beside K2 of the current one) with the one-rank mailbox: call s
    returns the losses of batch s-1, bit-identical gradients; switching between the in-order and the pipelined schedule
    re-binds the mailbox's count buffers (fused publish sends value[step & 1])."""
    import synthetic
    hw, B = (256, 320), 4
    N = rn.anchors_for_shape(hw + (3,)).shape[0]
    imgs = [synthetic.PageShape(hw + (3,)) for _ in range(B)]
    cls, reg = synthetic.training_predictions(6, B, N, classes=1)
    cls_h, reg_h = torch.from_numpy(cls).pin_memory(), torch.from_numpy(reg).pin_memory()
    batches = [[synthetic.gt_for_page(2, 10 * s + i, hw=hw, gmax=6) for i in range(B)] for s in range(9)]
    plain = rn.pipeline.TargetLossStep(hw + (3,), B, 8, 1)
    plain.load_predictions(cls_h, reg_h)
    want = []
    for anns in batches:
        plain.load_annotations(imgs, anns)
        l = plain.run().cpu().numpy().copy()
        want.append((l, plain.grad_cls.clone(), plain.grad_reg.clone()))
    assert len(set(float(w[0][2]) for w in want)) > 3        # the batches have different positive counts

    monkeypatch.setenv("RN_B200_PEER_BOX", "force")
    monkeypatch.setenv("RN_B200_PEER_FUSED", fused)
    boxed = rn.pipeline.TargetLossStep(hw + (3,), B, 8, 1)
    boxed.load_predictions(cls_h, reg_h)

    def check(got, s):
        got = got.cpu().numpy()
        assert got[2] == want[s][0][2], (s, got, want[s][0])
        assert np.allclose(got[:2], want[s][0][:2], rtol=1e-6, atol=0)
        assert torch.equal(boxed.grad_cls, want[s][1]) and torch.equal(boxed.grad_reg, want[s][2])

    overlap_ok = fused == "1"
    boxed.load_annotations(imgs, batches[0])
    check(boxed.run_pipelined(), 0)                           # the warm-up's targets are batch 0's
    for s in (1, 2, 3):
        boxed.load_annotations(imgs, batches[s])
        check(boxed.run_pipelined(overlap=overlap_ok and s % 2 == 1), s - 1)
    if not overlap_ok:
        # separate publish kernel: K2 takes "the step before the latest" (lag 1), so the two schedules cannot be mixed
        # on one step object, and the overlapped form is refused
        with pytest.raises(ValueError):
            boxed.run_pipelined(overlap=True)
        return
    for s in (4, 5, 6):                                       # in order again (odd number of steps: the parity flips)
        boxed.load_annotations(imgs, batches[s])
        check(boxed.run(), s)
    boxed.load_annotations(imgs, batches[7])
    boxed.run_pipelined(overlap=overlap_ok)                   # consumes the targets left in the pipeline (batch 3)
    check(boxed.losses, 3)
    boxed.load_annotations(imgs, batches[8])
    check(boxed.run_pipelined(overlap=overlap_ok), 7)


def test_detection_step_graph_matches_direct_head(rn):
    """DetectionStep (static buffers, the three kernels of the inference tail replayed as one CUDA graph) gives the
    detections of the direct DetectionHead call, bit for bit, batch after batch."""
    import synthetic
    hw, B = (256, 320), 3
    anchors = np.asarray(rn.anchors_for_shape(hw + (3,)))
    head = rn.DetectionHead()
    step = rn.pipeline.DetectionStep(hw, B, 1)
    eager = rn.pipeline.DetectionStep(hw, B, 1, use_graph=False)
    for s in range(3):
        anns = [synthetic.gt_for_page(3, 11 * s + i, hw=hw, gmax=5) for i in range(B)]
        cls, reg = synthetic.inference_predictions(3, B, anchors, anns, classes=1, first_page=3 * s)
        cls_d, reg_d = torch.from_numpy(cls).cuda(), torch.from_numpy(reg).cuda()
        want = head([(B,) + hw + (3,), reg_d, cls_d])
        for st in (step, eager):
            st.load_predictions(cls_d, reg_d)
            got = st.run()
            torch.cuda.synchronize()
            assert all(torch.equal(g, w) for g, w in zip(got, want))
            assert torch.equal(st.indices, head.last_indices)
            st.check()
        assert int((want[1] >= 0).sum()) > 0
