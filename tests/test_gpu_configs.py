"""GPU parity at the FULL per-GPU batch of BASELINE configs[3] (1600x2400, 4 pages) and configs[4] (80 classes, <= 100 GT,
16 pages): the whole batch runs through the product path (TargetLossStep + DetectionStep, the objects bench.py times), the
oracle is evaluated on two pages of each (it needs seconds per page at these sizes), and size-independent properties cover
the rest (state consistency, positive counts, loss linearity over pages)."""
import numpy as np
import pytest
import torch

import synthetic
from oracle import anchors_np as OA
from oracle import layers_np as L
from oracle import losses_np as OL

pytestmark = pytest.mark.gpu


def same(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = np.asarray(b)
    return a.shape == b.shape and a.tobytes() == b.astype(a.dtype).tobytes()


def close(a, b, rtol=1e-5):
    return np.allclose(np.asarray(a, np.float64), np.asarray(b, np.float64), rtol=rtol, atol=0)


@pytest.mark.parametrize("cfg,B,pages", [(4, 4, (0, 3)), (5, 16, (0, 15))])
def test_full_batch_of_config(rn, cfg, B, pages):
    c = synthetic.CONFIGS[cfg]
    hw, C, gmax = c['hw'], c['classes'], c['gmax'] + 2
    anchors = OA.anchors_for_shape(hw + (3,))
    N = anchors.shape[0]
    images, anns = synthetic.training_batch(cfg, batch=B, anchors=anchors)
    cls_d, reg_d = synthetic.inference_predictions_torch(cfg, B, anchors, anns, classes=C, device="cuda")
    step = rn.pipeline.TargetLossStep(hw + (3,), B, gmax, C)
    step.load_annotations(images, anns)
    step.cls_pred, step.reg_pred = cls_d, reg_d
    step.grad_cls, step.grad_reg = torch.empty_like(cls_d), torch.empty_like(reg_d)
    det = rn.pipeline.DetectionStep(hw, B, C)
    det.cls_pred, det.reg_pred = cls_d, reg_d
    for _ in range(2):                                      # the second pass replays the captured graphs
        step.run()
        det.run()
    torch.cuda.synchronize()
    det.check()
    losses = step.losses.cpu().numpy()
    # ---- properties over the whole batch -------------------------------------------------------------------
    st = step.y_cls[:, :, -1]
    assert torch.equal(st, step.y_reg[:, :, -1])
    assert set(torch.unique(st).cpu().numpy().tolist()) <= {-1.0, 0.0, 1.0}
    npos = (st == 1).sum(dim=1).to(torch.int32)
    assert torch.equal(npos, step.npos) and float(step.npos_total) == float(npos.sum()) == float(losses[2])
    onehot = step.y_cls[:, :, :C].sum(dim=2)
    assert bool(((onehot == 1) | (onehot == 0)).all()) and bool((onehot[st == 1] == 1).all()) and not bool(onehot[st == 0].any())
    # the loss is a sum over pages: per-page launches with the batch normaliser add up, gradients are bit-equal
    norm = torch.tensor([losses[2]], device="cuda")
    acc = np.zeros(2)
    page = lambda t, b: t[b:b + 1].clone()                  # (page slices of an odd-sized page are not 16-byte aligned)
    for b in range(B):
        lb, gcb, grb = rn.detection_losses(page(step.y_reg, b), page(step.y_cls, b), page(reg_d, b), page(cls_d, b),
                                           normalizer=norm, shared_state=True)
        acc += lb.cpu().numpy()[:2].astype(np.float64)
        if b in pages:
            assert torch.equal(gcb[0], step.grad_cls[b]) and torch.equal(grb[0], step.grad_reg[b])
    assert close(losses[:2], acc)
    scores = det.scores
    assert bool((scores[:, :-1] >= scores[:, 1:]).all())                     # detections sorted, padding (-1) last
    # ---- the oracle on two pages ---------------------------------------------------------------------------
    for b in pages:
        oreg, olab = OA.anchor_targets_bbox(anchors, [images[b]], [anns[b]], C)
        assert same(step.y_reg[b], oreg[0]) and same(step.y_cls[b], olab[0])
        cls_b, reg_b = cls_d[b:b + 1].cpu().numpy(), reg_d[b:b + 1].cpu().numpy()
        wf, wgf = OL.focal()(olab, cls_b, return_grad=True, normalizer=float(losses[2]))
        ws, wgs = OL.smooth_l1()(oreg, reg_b, return_grad=True, normalizer=float(losses[2]))
        lb = rn.detection_losses(page(step.y_reg, b), page(step.y_cls, b), page(reg_d, b), page(cls_d, b), normalizer=norm)[0].cpu().numpy()
        assert close(lb[0], wf) and close(lb[1], ws)
        assert np.allclose(step.grad_cls[b].cpu().numpy(), wgf[0], rtol=1e-5, atol=1e-7 * float(np.abs(wgf).max()))
        assert np.allclose(step.grad_reg[b].cpu().numpy(), wgs[0], rtol=1e-5, atol=1e-9)
        want = L.detect(hw, reg_b, cls_b)
        assert same(det.indices[b], want[3][0]) and same(det.labels[b], want[2][0])
        assert same(det.scores[b], want[1][0]) and same(det.boxes[b], want[0][0])
