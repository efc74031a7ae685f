"""GPU parity, K3/K4/K5 and the detection-head layers vs the numpy oracle (restated TensorFlow
semantics, oracle/layers_np.py).  Bar: selected anchor indices, labels and order BIT-EXACT; boxes and
scores bit-exact too (fp32, same operation order, no FMA contraction) -- stronger than the 1e-5 asked."""
import numpy as np
import pytest
import torch

import synthetic
from oracle import anchors_np as OA
from oracle import layers_np as L

pytestmark = pytest.mark.gpu


def same(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = np.asarray(b)
    return a.shape == b.shape and a.tobytes() == b.astype(a.dtype).tobytes()


def test_layers_bit_exact(rn, golden_tf):
    g = golden_tf
    ap = OA.AnchorParameters_default
    feats = [(2, h, w, 256) for h, w in synthetic.level_shapes((67, 93))]
    per = [rn.Anchors(size=ap.sizes[i], stride=ap.strides[i], ratios=ap.ratios, scales=ap.scales,
                      name='anchors_%d' % i)(torch.empty(f, device="cuda")) for i, f in enumerate(feats)]
    anchors = torch.cat(per, dim=1)
    assert same(anchors[0], g['det_anchors_f32']) and same(anchors[1], g['det_anchors_f32'])
    boxes = rn.RegressBoxes(name='boxes')([anchors, torch.tensor(g['det_reg'], device="cuda")])
    assert same(boxes, g['det_boxes'])
    clipped = rn.ClipBoxes(name='clipped_boxes')([torch.empty((2, 67, 93, 3)), boxes])
    assert same(clipped, g['det_clipped'])
    assert same(rn.bbox_transform_inv(g['det_anchors_f32'][None], g['det_reg'][:1]), g['det_boxes'][:1])
    lvl = rn.utils.shift((9, 12), 8, OA.generate_anchors(32))
    assert same(lvl, L.shift_f32((9, 12), 8, OA.generate_anchors(32).astype(np.float32)))
    # configs round-trip like the Keras layers
    assert rn.RegressBoxes().get_config()['std'] == [0.2, 0.2, 0.2, 0.2]
    cfg = rn.FilterDetections(name='filtered_detections').get_config()
    assert cfg['max_detections'] == 300 and cfg['score_threshold'] == 0.05 and cfg['nms_threshold'] == 0.5
    assert rn.Anchors(32, 8, ratios=[0.5, 1, 2], scales=[1, 1.2]).get_config()['ratios'] == [0.5, 1, 2]
    with pytest.raises(ValueError):
        rn.RegressBoxes(mean=0)


@pytest.mark.parametrize("tag,kw", [('default', {}), ('agnostic', dict(class_specific_filter=False)),
                                    ('nonms', dict(nms=False)),
                                    ('small', dict(max_detections=20, nms_threshold=0.3, score_threshold=0.2))])
def test_filter_detections_golden(rn, golden_tf, tag, kw):
    g = golden_tf
    clipped = torch.tensor(g['det_clipped'], device="cuda")
    cls = torch.tensor(g['det_cls'], device="cuda")
    layer = rn.FilterDetections(**kw)
    other = torch.arange(2 * clipped.shape[1] * 2, dtype=torch.float32, device="cuda").view(2, -1, 2)
    boxes, scores, labels, oth = layer([clipped, cls, other])
    assert same(layer.last_indices, g['det_%s_idx' % tag])
    assert same(labels, g['det_%s_labels' % tag]) and labels.dtype == torch.int32
    assert same(boxes, g['det_%s_boxes' % tag]) and same(scores, g['det_%s_scores' % tag])
    idx = g['det_%s_idx' % tag]
    want_other = np.where(idx[:, :, None] >= 0, other.cpu().numpy()[np.arange(2)[:, None], np.maximum(idx, 0)], -1)
    assert same(oth, want_other)
    # fused head: anchors generated in-kernel + decode + clip + threshold in one pass
    head = rn.DetectionHead(applyNms=kw.get('nms', True), class_specific_filter=kw.get('class_specific_filter', True),
                            **{k: v for k, v in kw.items() if k not in ('nms', 'class_specific_filter')})
    b2, s2, l2 = head([(2, 67, 93, 3), torch.tensor(g['det_reg'], device="cuda"), cls])
    assert same(head.last_indices, idx) and same(b2, g['det_%s_boxes' % tag]) and same(s2, g['det_%s_scores' % tag]) and same(l2, g['det_%s_labels' % tag])
    # single-image functional form
    r = rn.filter_detections(clipped[1], cls[1], **kw)
    assert same(r[0], g['det_%s_boxes' % tag][1]) and same(r[1], g['det_%s_scores' % tag][1]) and same(r[2], g['det_%s_labels' % tag][1])


def _random_boxes(rs, n, extent=800, lo=5, hi=300):
    c = rs.uniform(0, extent, (n, 2))
    wh = rs.uniform(lo, hi, (n, 2))
    return np.concatenate([c, c + wh], 1).astype(np.float32)


@pytest.mark.parametrize("n,lo,hi,max_out", [(3000, 5, 300, 300), (9000, 2, 30, 300), (5000, 200, 400, 50),
                                              (1, 5, 50, 10), (0, 5, 50, 10), (2049, 3, 40, 1000)])
def test_nms_vs_oracle_and_torchvision(rn, n, lo, hi, max_out):
    """rn_nms = tf.image.non_max_suppression restated.  (9000, small boxes) and (2049) force several
    radix-select rounds; equal scores exercise the index tie-break; zero-area boxes are never suppressed."""
    import torchvision
    rs = np.random.RandomState(n + max_out)
    b = _random_boxes(rs, n, lo=lo, hi=hi)
    s = rs.uniform(0, 1, n).astype(np.float32)
    if n > 100:
        s[rs.randint(0, n, n // 6)] = s[0]
        b[5, 2:] = b[5, :2]                              # zero-area box
        b[7] = b[7][[2, 3, 0, 1]]                        # corners given in the other order
    want = L.non_max_suppression(b, s, max_out, 0.5)
    got = rn.layers.non_max_suppression(b, s, max_out, 0.5).cpu().numpy()
    assert np.array_equal(got, want)
    if n > 0:
        tv = torchvision.ops.nms(torch.from_numpy(np.stack([np.minimum(b[:, 0], b[:, 2]), np.minimum(b[:, 1], b[:, 3]),
                                                           np.maximum(b[:, 0], b[:, 2]), np.maximum(b[:, 1], b[:, 3])], 1)),
                                 torch.from_numpy(s), 0.5)[:max_out].numpy()
        assert np.array_equal(got, tv)


@pytest.mark.parametrize("cfg,B,C", [(3, 3, 1), (5, 1, 80)])
def test_synthetic_inference_vs_oracle(rn, cfg, B, C):
    """BASELINE config 3 (C=1) and 5 (C=80) pages at full anchor count, fused path vs oracle."""
    hw = synthetic.CONFIGS[cfg]['hw']
    anchors = OA.anchors_for_shape(hw + (3,))
    _, anns = synthetic.training_batch(cfg, batch=B)
    cls, reg = synthetic.inference_predictions(cfg, B, anchors, anns, classes=C)
    want = L.detect(hw, reg, cls)
    head = rn.DetectionHead()
    b, s, l = head([(B,) + hw + (3,), torch.tensor(reg, device="cuda"), torch.tensor(cls, device="cuda")])
    assert same(head.last_indices, want[3]) and same(l, want[2]) and same(b, want[0]) and same(s, want[1])
    # layer-by-layer path gives the same answer
    ap = OA.AnchorParameters_default
    per = [rn.Anchors(ap.sizes[i], ap.strides[i], ap.ratios, ap.scales)((B, h, w, 8)) for i, (h, w) in enumerate(synthetic.level_shapes(hw))]
    boxes = rn.ClipBoxes()([(B,) + hw + (3,), rn.RegressBoxes()([torch.cat(per, 1), torch.tensor(reg, device="cuda")])])
    fd = rn.FilterDetections()
    b2, s2, l2 = fd([boxes, torch.tensor(cls, device="cuda")])
    assert same(fd.last_indices, want[3]) and same(b2, want[0])


def test_dense_scores_many_rounds_and_extensions(rn):
    """Every anchor above the threshold and heavy overlap (all boxes near one table): NMS must walk far
    beyond the first 2048 candidates.  Also: pre_nms_top_k (extension) == reference when it cannot matter,
    and cand_cap overflow raises instead of returning inexact results."""
    rs = np.random.RandomState(77)
    hw = (200, 300)
    anchors = L.all_anchors_f32(hw)[0]
    n = anchors.shape[0]
    cls = rs.uniform(0.06, 0.99, (1, n, 1)).astype(np.float32)
    reg = rs.normal(0, 0.3, (1, n, 4)).astype(np.float32)
    want = L.detect(hw, reg, cls)
    head = rn.DetectionHead()
    b, s, l = head([(1,) + hw + (3,), reg, cls])
    assert same(head.last_indices, want[3]) and same(b, want[0])
    big = rn.DetectionHead(pre_nms_top_k=n)
    big([(1,) + hw + (3,), reg, cls])
    assert same(big.last_indices, want[3])
    topk = rn.DetectionHead(pre_nms_top_k=1000)
    topk([(1,) + hw + (3,), reg, cls])
    order = np.argsort(-cls[0, :, 0], kind='stable')[:1000]
    cls_k = np.zeros_like(cls)
    cls_k[0, order, 0] = cls[0, order, 0]
    assert same(topk.last_indices, L.detect(hw, reg, cls_k)[3])
    with pytest.raises(rn._lib.RnError):
        rn.DetectionHead(cand_cap=64)([(1,) + hw + (3,), reg, cls])


def test_empty_and_all_below_threshold(rn):
    hw = (64, 64)
    n = L.all_anchors_f32(hw).shape[1]
    cls = np.full((2, n, 2), 0.01, np.float32)
    reg = np.zeros((2, n, 4), np.float32)
    b, s, l = rn.DetectionHead()([(2,) + hw + (3,), reg, cls])
    assert (b.cpu().numpy() == -1).all() and (s.cpu().numpy() == -1).all() and (l.cpu().numpy() == -1).all()
    cls[1, 17, 1] = 0.05                         # strict >: exactly the threshold is not a candidate
    cls[1, 18, 1] = np.nextafter(np.float32(0.05), np.float32(1))
    head = rn.DetectionHead()
    b, s, l = head([(2,) + hw + (3,), reg, cls])
    assert head.last_indices.cpu().numpy()[1, 0] == 18 and (head.last_indices.cpu().numpy()[1, 1:] == -1).all()
    assert l.cpu().numpy()[1, 0] == 1


@pytest.mark.parametrize("seed", list(range(12)))
def test_randomized_filter_detections(rn, seed):
    """Random box sets with tight clusters (many IoUs near the threshold), quantised scores (many exact ties: order
    must fall back to the anchor index), random class counts, thresholds, max_detections, class-specific or not,
    NMS on or off -- boxes, scores, labels and selected indices bit for bit against the oracle."""
    rs = np.random.RandomState(4000 + seed)
    B = int(rs.randint(1, 4))
    N = int(rs.choice([1, 37, 700, 3001, 9000]))
    C = int(rs.choice([1, 2, 7]))
    centers = rs.uniform(50, 950, (B, max(1, N // 40), 2))
    which = rs.randint(0, centers.shape[1], (B, N))
    cxy = centers[np.arange(B)[:, None], which] + rs.normal(0, 6, (B, N, 2))
    wh = rs.uniform(20, 90, (B, N, 2)) * np.where(rs.uniform(size=(B, N, 1)) < 0.05, 0.0, 1.0)    # 5 % zero-area boxes
    boxes = np.concatenate([cxy - wh / 2, cxy + wh / 2], axis=2).astype(np.float32)
    flip = rs.uniform(size=(B, N)) < 0.1                                                    # corners in the other order
    boxes[flip] = boxes[flip][:, [2, 3, 0, 1]]
    cls = (np.round(rs.uniform(0, 1, (B, N, C)) ** 3 * 50) / 50).astype(np.float32)         # quantised: ties everywhere
    kw = dict(class_specific_filter=bool(seed % 2 == 0), nms=bool(seed % 5 != 4),
              score_threshold=float(rs.choice([0.05, 0.3, 0.0])), max_detections=int(rs.choice([300, 17, 1000])),
              nms_threshold=float(rs.choice([0.5, 0.3, 0.75])))
    want = L.filter_detections_batch(boxes, cls, **kw)
    layer = rn.FilterDetections(**kw)
    b, s, l = layer([torch.tensor(boxes, device="cuda"), torch.tensor(cls, device="cuda")])
    assert same(layer.last_indices, want[3]) and same(l, want[2]) and same(s, want[1]) and same(b, want[0])


@pytest.mark.parametrize("members,bg,max_out", [(500, 1000, 300), (700, 1100, 1000), (760, 400, 40)])
def test_heavy_clusters_sweep_between_rounds(rn, members, bg, max_out):
    """Ten tight clusters of high-scoring boxes (thousands of candidates suppressed by ten selections) in front of a
    background of small boxes: the first sorted chunk yields only a few selections, so the kernel strikes out the
    suppressed part of what follows in its registers before the next round (the sweep between rounds) -- the result must
    still be the greedy order of the oracle and of torchvision, bit for bit."""
    import torchvision
    rs = np.random.RandomState(members + bg)
    centers = rs.uniform(200, 1800, (10, 2))
    size = rs.uniform(150, 300, (10, 2))
    which = rs.randint(0, 10, 10 * members)
    c = centers[which] + rs.normal(0, 4, (which.size, 2))
    wh = size[which] * rs.uniform(0.93, 1.07, (which.size, 2))
    cl = np.concatenate([c - wh / 2, c + wh / 2], 1)
    bc = rs.uniform(0, 2000, (bg, 2))
    bwh = rs.uniform(8, 30, (bg, 2))
    bb = np.concatenate([bc - bwh / 2, bc + bwh / 2], 1)
    b = np.concatenate([cl, bb], 0).astype(np.float32)
    s = np.concatenate([rs.uniform(0.5, 0.99, cl.shape[0]), rs.uniform(0.05, 0.3, bg)]).astype(np.float32)
    perm = rs.permutation(b.shape[0])
    b, s = b[perm], s[perm]
    assert b.shape[0] <= 8192                               # keys held in registers: the sweep is possible
    want = L.non_max_suppression(b, s, max_out, 0.5)
    got = rn.layers.non_max_suppression(b, s, max_out, 0.5).cpu().numpy()
    assert np.array_equal(got, want)
    tv = torchvision.ops.nms(torch.from_numpy(b), torch.from_numpy(s), 0.5)[:max_out].numpy()
    assert np.array_equal(got, tv)
    # the same boxes through the layer (threshold + per-class slabs; scores of the clusters in class 1, background in class 0)
    cls = np.zeros((1, b.shape[0], 2), np.float32)
    hi = s >= 0.5
    cls[0, hi, 1] = s[hi]
    cls[0, ~hi, 0] = s[~hi]
    wantf = L.filter_detections_batch(b[None], cls, max_detections=max_out)
    layer = rn.FilterDetections(max_detections=max_out)
    fb, fs, fl = layer([torch.tensor(b[None], device="cuda"), torch.tensor(cls, device="cuda")])
    assert same(layer.last_indices, wantf[3]) and same(fl, wantf[2]) and same(fs, wantf[1]) and same(fb, wantf[0])


def test_stages_launched_one_at_a_time_equal_the_full_call(rn):
    """The measurement hook rn_debug_filter_stages (bench.py times every kernel of the filter call alone): K3, the NMS
    kernel and the merge launched as three separate calls on one workspace -- eagerly and as re-captured CUDA graphs --
    leave exactly the detections of the normal call; the default mask is restored."""
    hw, B = (256, 320), 3
    anchors = OA.anchors_for_shape(hw + (3,))
    _, anns = synthetic.training_batch(3, batch=B)
    anns = [synthetic.gt_for_page(2, i, hw=hw, gmax=6) for i in range(B)]
    cls, reg = synthetic.inference_predictions(3, B, anchors, anns, classes=1)
    lib = rn._lib.load()
    for use_graph in (False, True):
        det = rn.pipeline.DetectionStep(hw, B, 1, use_graph=use_graph)
        det.load_predictions(torch.from_numpy(cls), torch.from_numpy(reg))
        want = [t.clone() for t in det.run()]
        torch.cuda.synchronize()
        for t in det.out:
            t.fill_(-7)
        try:
            for mask in (1, 2, 4):
                lib.rn_debug_filter_stages(mask)
                det._graph = None
                det.run()
                det.run()                                   # a stage may be repeated on what the earlier stages left
        finally:
            lib.rn_debug_filter_stages(7)
            det._graph = None
        torch.cuda.synchronize()
        got = [det.boxes, det.scores, det.labels]
        assert all(torch.equal(a, b) for a, b in zip(got, want))
        assert all(torch.equal(a, b) for a, b in zip(det.run(), want))


@pytest.mark.parametrize("C,N,hot", [(2, 2000, False), (33, 1200, True), (80, 2048, True), (160, 800, False), (161, 800, False)])
def test_several_classes_key_lists(rn, C, N, hot):
    """k_threshold_keys_classes (class-specific filtering with several classes: keys collected per class in shared memory,
    flushed in runs): one, two and five 32-list words, the largest C it takes and the first one it does not (C = 161: the
    one-class stream kernel's general branch), a HOT class whose list overflows inside a round (its surplus goes straight to
    the slab), lists that are only flushed after the last round, and a slab that is too small (raises)."""
    rs = np.random.RandomState(9100 + C)
    B = 2
    cxy = rs.uniform(60, 900, (B, N, 2))
    wh = rs.uniform(20, 90, (B, N, 2))
    boxes = np.concatenate([cxy - wh / 2, cxy + wh / 2], axis=2).astype(np.float32)
    cls = (rs.uniform(0, 1, (B, N, C)) ** 6).astype(np.float32)            # ~ 60 % below the 0.05 threshold
    if hot:
        cls[:, :, 1] = rs.uniform(0.06, 0.99, (B, N)).astype(np.float32)  # every anchor fires for class 1
        cls[1, :, C - 1] = rs.uniform(0.5, 0.99, N).astype(np.float32)    # and, on page 1, for the last class
    kw = dict(class_specific_filter=True, nms=True, score_threshold=0.05, max_detections=300, nms_threshold=0.5)
    want = L.filter_detections_batch(boxes, cls, **kw)
    layer = rn.FilterDetections(**kw)
    b, s, l = layer([torch.tensor(boxes, device="cuda"), torch.tensor(cls, device="cuda")])
    assert same(layer.last_indices, want[3]) and same(l, want[2]) and same(s, want[1]) and same(b, want[0])
    if hot:
        with pytest.raises(rn._lib.RnError):
            rn.FilterDetections(cand_cap=64, **kw)([torch.tensor(boxes, device="cuda"), torch.tensor(cls, device="cuda")])
