"""Multi-GPU checks that need real GPUs (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 profiles/multigpu_check.py

1. peer mailbox: run(), run_pipelined(overlap=True) and run_from_host give the MERGED batch's losses, identical on every rank
   and equal (1e-6) to one process computing the whole batch; the normaliser is the global positive count;
2. RN_B200_PEER_BOX=0 (no mailbox: NCCL all-reduce of the count and of the loss sums): run_pipelined() equals run(),
   overlap=True is refused;
3. a mailbox timeout raises instead of hanging or returning NaN silently (rank 1 skips a step on purpose).
`--two-devices` (single process, >= 2 GPUs visible): the library's per-device attribute caches -- K1 / NMS / K2 are first
launched on cuda:1, then on cuda:0.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import synthetic  # noqa: E402


def two_devices():
    import retinanet_b200 as rn
    assert torch.cuda.device_count() >= 2
    hw, B = (256, 320), 3
    anchors = np.asarray(rn.anchors_for_shape(hw + (3,)))
    images = [synthetic.PageShape(hw + (3,)) for _ in range(B)]
    anns = [synthetic.gt_for_page(2, i, hw=hw, gmax=6) for i in range(B)]
    cls, reg = synthetic.inference_predictions(3, B, anchors, anns, classes=1)
    out = {}
    for dev in (1, 0):                                      # device 1 FIRST: a process-wide "done" flag would skip device 0's opt-in
        torch.cuda.set_device(dev)
        step = rn.pipeline.TargetLossStep(hw + (3,), B, 8, 1, use_graph=False)
        step.load_annotations(images, anns)
        step.load_predictions(torch.from_numpy(cls), torch.from_numpy(reg))
        losses = step.run().cpu().numpy().copy()
        det = rn.DetectionHead()([(B,) + hw + (3,), torch.from_numpy(reg).cuda(), torch.from_numpy(cls).cuda()])
        torch.cuda.synchronize()
        out[dev] = (losses, det[1].cpu().numpy().copy(), step.y_reg.cpu().numpy().copy())
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][2], out[1][2])
    print("two devices in one process: K1 / K2 / K3 / NMS launched on cuda:1 first, then cuda:0 -- identical results, OK")


def main():
    if "--two-devices" in sys.argv:
        return two_devices()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    import retinanet_b200 as rn
    hw, B = (512, 640), 4
    anchors = np.asarray(rn.anchors_for_shape(hw + (3,)))
    N = anchors.shape[0]
    g_images = [synthetic.PageShape(hw + (3,)) for _ in range(world * B)]
    g_anns = [synthetic.gt_for_page(2, 50 + i, hw=hw, gmax=9) for i in range(world * B)]
    g_cls, g_reg = synthetic.training_predictions(2, world * B, N, classes=1)
    mine = slice(rank * B, (rank + 1) * B)
    cls_h, reg_h = torch.from_numpy(g_cls[mine]).pin_memory(), torch.from_numpy(g_reg[mine]).pin_memory()
    # the whole batch in one process (no exchange): the reference for the merged losses
    solo = rn.pipeline.TargetLossStep(hw + (3,), world * B, 12, 1, peer_box=False, use_graph=False)
    solo.peer = None
    solo.load_annotations(g_images, g_anns)
    solo.load_predictions(torch.from_numpy(g_cls), torch.from_numpy(g_reg))
    from retinanet_b200 import anchors as _a, losses as _l
    solo._targets()
    _l.detection_losses(solo.y_reg, solo.y_cls, solo.reg_pred, solo.cls_pred, normalizer=solo.npos_total,
                        out=(solo.losses, solo.grad_cls, solo.grad_reg), workspace=solo.loss_ws, **solo.loss_kw)
    want = solo.losses.cpu().numpy().copy()
    want_gc = solo.grad_cls[mine].clone()

    def same_on_all_ranks(t, what):
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), (what, lo, hi)

    def check(step, losses, what):
        l = losses.cpu().numpy() if isinstance(losses, torch.Tensor) else np.asarray(losses)
        assert l[2] == want[2], (what, l, want)
        assert np.allclose(l[:2], want[:2], rtol=2e-6, atol=0), (what, l, want)
        same_on_all_ranks(torch.as_tensor(l, device=device), what)

    # ---- 1. mailbox ------------------------------------------------------------------------------------------
    step = rn.pipeline.TargetLossStep(hw + (3,), B, 12, 1)
    assert step.peer is not None and step.peer_fused, "peer mailbox unavailable on this box"
    step.load_annotations(g_images[mine], g_anns[mine])
    step.load_predictions(cls_h, reg_h)
    for _ in range(6):
        check(step, step.run(), "mailbox run()")
    assert torch.equal(step.grad_cls, want_gc), "gradients differ from the one-process batch"
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    check(step, step.run(events=evs), "mailbox run(events)")
    for _ in range(5):
        l = step.run_pipelined(overlap=True)
    check(step, l, "mailbox run_pipelined(overlap)")
    check(step, step.run_from_host(g_images[mine], g_anns[mine], cls_h, reg_h, chunks=1, gather_reg_from_host=True), "run_from_host chunks=1")
    check(step, step.run_from_host(g_images[mine], g_anns[mine], cls_h, reg_h, chunks=2), "run_from_host chunks=2")
    check(step, step.run(), "mailbox run() again")
    step.check()
    # ---- 3. timeout: rank 1 sits out one step ------------------------------------------------------------------
    step.peer.set_timeout(0.25)
    if rank != 1:
        step.run()
        torch.cuda.synchronize()
        try:
            step.check()
            raise AssertionError("a missing peer must raise")
        except rn._lib.RnError:
            pass
        assert bool(torch.isnan(step.losses[2])) or bool(torch.isnan(step.losses[0]))
    dist.barrier()
    del step
    # ---- 2. no mailbox -------------------------------------------------------------------------------------------
    os.environ["RN_B200_PEER_BOX"] = "0"
    nb = rn.pipeline.TargetLossStep(hw + (3,), B, 12, 1)
    assert nb.peer is None
    nb.load_annotations(g_images[mine], g_anns[mine])
    nb.load_predictions(cls_h, reg_h)
    for _ in range(3):
        check(nb, nb.run(), "all_reduce run()")
    for _ in range(4):
        l = nb.run_pipelined()
    check(nb, l, "all_reduce run_pipelined()")
    assert torch.equal(nb.grad_cls, want_gc)
    try:
        nb.run_pipelined(overlap=True)
        raise AssertionError("overlap without the mailbox must be refused")
    except ValueError:
        pass
    check(nb, nb.run_from_host(g_images[mine], g_anns[mine], cls_h, reg_h, chunks=1), "all_reduce run_from_host")
    dist.barrier()
    if rank == 0:
        print("multi-GPU checks on %d ranks: merged-batch losses %s identical on every rank, mailbox / all-reduce / pipelined / "
              "host-input paths agree, timeout raises -- OK" % (world, want))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
