"""GPU parity, K1: CUDA path (through the C-ABI) vs the numpy oracle and the golden fixtures generated
from the reference's own code.  Bar: BIT-EXACT (labels, states, argmax indices AND the float32 regression
targets -- the fp64 pipeline is reproduced operation by operation)."""
import hashlib

import numpy as np
import pytest
import torch

import synthetic
from oracle import anchors_np as O

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and a.tobytes() == b.tobytes()


@pytest.mark.parametrize("shape", [(800, 1333, 3), (1600, 2400, 3), (67, 93, 3), (1, 1, 3), (511, 513, 3)])
def test_anchors_for_shape_bit_exact(rn, shape):
    got = rn.anchors_for_shape(shape)
    want = O.anchors_for_shape(shape)
    assert got.dtype == np.float64 and same(np.asarray(got), want)


def test_anchors_golden(rn, golden_numpy):
    g = golden_numpy
    a = np.asarray(rn.anchors_for_shape((800, 1333, 3)))
    assert list(a.shape) == list(g['a800_shape']) and sha(a) == str(g['a800_sha'])
    assert same(a[g['a800_rows_idx']], g['a800_rows'])
    assert sha(np.asarray(rn.anchors_for_shape((1600, 2400, 3)))) == str(g['a1600_sha'])
    custom = rn.AnchorParameters([24, 48, 100, 200, 400], [8, 16, 32, 64, 128],
                                 np.array([0.3, 1, 2.5], np.float32), np.array([1, 1.3], np.float32))
    assert same(np.asarray(rn.anchors_for_shape((70, 90, 3), anchor_params=custom)), g['custom_anchors'])
    for size in (32, 64, 128, 256, 512):
        assert same(rn.generate_anchors(size), g['base_%d' % size])


def test_shift_and_shapes(rn):
    base = O.generate_anchors(64)
    assert same(rn.anchors.shift((5, 7), 16, base), O.shift((5, 7), 16, base))
    got = rn.guess_shapes((800, 1333, 3), [3, 4, 5, 6, 7])
    assert [tuple(int(v) for v in s) for s in got] == [(100, 167), (50, 84), (25, 42), (13, 21), (7, 11)]


def test_compute_overlap_bit_exact(rn, golden_numpy):
    g = golden_numpy
    got = rn.compute_overlap(g['small_anchors'], g['overlap_gt'])
    assert got.dtype == np.float32 and same(got, g['overlap_small'])
    rs = np.random.RandomState(3)
    a = O.anchors_for_shape((200, 300, 3))
    gt = np.stack([rs.uniform(0, 150, 33), rs.uniform(0, 100, 33), rs.uniform(150, 300, 33), rs.uniform(100, 200, 33)], 1)
    assert same(rn.compute_overlap(a, gt), O.compute_overlap(a, gt))
    assert rn.compute_overlap(a, np.zeros((0, 4))).shape == (a.shape[0], 0)


def test_small_batch_golden_full(rn, golden_numpy):
    g = golden_numpy
    anchors = rn.anchors_for_shape((67, 93, 3))
    imgs = [synthetic.PageShape(s) for s in g['small_shapes']]
    anns = [{'bboxes': g['small_gt_%d' % i], 'labels': g['small_gl_%d' % i]} for i in range(3)]
    reg, lab, npos = rn.anchor_targets_bbox(anchors, imgs, anns, 3, return_npos=True)
    assert same(reg, g['small_reg']) and same(lab, g['small_lab'])
    assert list(npos) == [int((g['small_lab'][b, :, -1] == 1).sum()) for b in range(3)]
    # explicit-anchor path (plain ndarray, no generation spec) gives the same bits
    reg2, lab2 = rn.anchor_targets_bbox(np.array(anchors), imgs, anns, 3)
    assert same(reg2, reg) and same(lab2, lab)


def test_survey_pages_golden(rn, golden_numpy):
    g = golden_numpy
    anchors = rn.anchors_for_shape((800, 1333, 3))
    page3 = {'bboxes': np.array([[100, 100, 600, 400], [50, 500, 700, 760], [800, 100, 1300, 300]], dtype=np.float64),
             'labels': np.zeros(3)}
    reg, lab = rn.anchor_targets_bbox(anchors, [synthetic.PageShape((800, 1333, 3))], [page3], 1)
    st = lab[0, :, -1]
    assert (st == 1).sum() == 66 and (st == -1).sum() == 943
    assert np.array_equal(np.nonzero(st == 1)[0], g['page3_pos']) and np.array_equal(np.nonzero(st == -1)[0], g['page3_ign'])
    assert sha(reg) == str(g['page3_reg_sha']) and sha(lab) == str(g['page3_lab_sha'])
    pos, ign, amax = rn.compute_gt_annotations(anchors, page3['bboxes'])
    assert amax.dtype == np.int64 and sha(amax.astype(np.int32)) == str(g['page3_argmax_sha'])
    empty = {'bboxes': np.zeros((0, 4)), 'labels': np.zeros((0,))}
    reg, lab = rn.anchor_targets_bbox(anchors, [synthetic.PageShape((800, 1333, 3))], [empty], 1)
    assert (lab[0, :, -1] == -1).sum() == 792 and not reg[0, :, :4].any()
    assert sha(reg) == str(g['empty_reg_sha']) and sha(lab) == str(g['empty_lab_sha'])


@pytest.mark.parametrize("name", ["exact", "dup", "f32tie", "degenerate"])
def test_tie_cases_golden(rn, golden_numpy, name):
    """Threshold / tie semantics: IoU == 0.5 positive, IoU == 0.4 background, equal maxima
    (also equal only after fp32 rounding) -> lowest GT index, zero-area GT."""
    g = golden_numpy
    anc, gt, gl = g['tie_%s_anchors' % name], g['tie_%s_gt' % name], g['tie_%s_gl' % name]
    reg, lab = rn.anchor_targets_bbox(anc, [synthetic.PageShape((2000, 2000, 3))], [{'bboxes': gt, 'labels': gl}], 3)
    assert same(reg, g['tie_%s_reg' % name]) and same(lab, g['tie_%s_lab' % name])
    _, _, amax = rn.compute_gt_annotations(anc, gt)
    assert np.array_equal(amax, g['tie_%s_argmax' % name])


@pytest.mark.parametrize("cfg,batch,mixed", [(1, 1, False), (2, 3, True), (5, 1, False), (4, 1, False)])
def test_synthetic_configs_golden_and_oracle(rn, golden_numpy, cfg, batch, mixed):
    g = golden_numpy
    c = synthetic.CONFIGS[cfg]
    anchors = rn.anchors_for_shape(c['hw'] + (3,))
    imgs, anns = synthetic.training_batch(cfg, batch=batch, mixed_widths=mixed, anchors=np.asarray(anchors))
    reg, lab, npos = rn.anchor_targets_bbox(anchors, imgs, anns, c['classes'], return_npos=True)
    assert sha(reg) == str(g['cfg%d_reg_sha' % cfg]) and sha(lab) == str(g['cfg%d_lab_sha' % cfg])
    assert list(npos) == list(g['cfg%d_npos' % cfg])
    assert same(reg[0, g['cfg%d_pos0' % cfg]], g['cfg%d_reg_pos0' % cfg])
    if cfg in (1, 2):
        oreg, olab = O.anchor_targets_bbox(np.asarray(anchors), imgs, anns, c['classes'])
        assert same(reg, oreg) and same(lab, olab)


def test_full_size_batch_vs_oracle_and_properties(rn):
    """Config 2 at full size (16 pages, <= 20 GT): equality with the oracle on 4 pages and the
    size-independent properties on all: state in {-1,0,1}, one-hot only on positives (or overridden
    positives), npos == count(state == 1), device output == host output."""
    c = synthetic.CONFIGS[2]
    anchors = rn.anchors_for_shape(c['hw'] + (3,))
    imgs, anns = synthetic.training_batch(2, anchors=np.asarray(anchors))
    reg_t, lab_t, npos_t = rn.anchor_targets_bbox(anchors, imgs, anns, 1, output="torch", return_npos=True)
    reg, lab = rn.anchor_targets_bbox(anchors, imgs, anns, 1)
    assert same(reg_t.cpu().numpy(), reg) and same(lab_t.cpu().numpy(), lab)
    st = lab[:, :, -1]
    assert set(np.unique(st)) <= {-1.0, 0.0, 1.0} and same(st, reg[:, :, -1])
    assert np.array_equal(npos_t.cpu().numpy(), (st == 1).sum(axis=1))
    assert not lab[:, :, 0][(st == 0)].any()
    oreg, olab = O.anchor_targets_bbox(np.asarray(anchors), imgs[:4], anns[:4], 1)
    assert same(reg[:4], oreg) and same(lab[:4], olab)


def test_many_gt_chunks_and_large_labels(rn):
    """More GT boxes than one shared-memory chunk (256) and C = 7: exercises chunked culling + generic label rows."""
    rs = np.random.RandomState(5)
    anchors = rn.anchors_for_shape((300, 400, 3))
    G = 700
    x1 = rs.uniform(0, 350, G); y1 = rs.uniform(0, 250, G)
    gt = np.stack([x1, y1, x1 + rs.uniform(5, 120, G), y1 + rs.uniform(5, 120, G)], 1)
    ann = {'bboxes': gt, 'labels': rs.randint(0, 7, G).astype(np.float64)}
    img = synthetic.PageShape((300, 390, 3))
    reg, lab = rn.anchor_targets_bbox(anchors, [img], [ann], 7)
    oreg, olab = O.anchor_targets_bbox(np.asarray(anchors), [img], [ann], 7)
    assert same(reg, oreg) and same(lab, olab)


def test_bbox_transform_and_errors(rn):
    rs = np.random.RandomState(9)
    a = O.anchors_for_shape((64, 64, 3))
    g = a + rs.normal(0, 3, a.shape)
    assert same(rn.bbox_transform(a, g), O.bbox_transform(a, g))
    assert same(rn.bbox_transform(a, g, mean=[0.1, 0, 0, 0], std=(0.1, 0.1, 0.2, 0.2)),
                O.bbox_transform(a, g, mean=[0.1, 0, 0, 0], std=(0.1, 0.1, 0.2, 0.2)))
    with pytest.raises(ValueError):
        rn.bbox_transform(a, g, mean=0.0)
    with pytest.raises(ValueError):
        rn.bbox_transform(a, g, std="x")
    with pytest.raises(AssertionError):
        rn.anchor_targets_bbox(a, [synthetic.PageShape((64, 64, 3))], [], 1)
    with pytest.raises(AssertionError):
        rn.anchor_targets_bbox(a, [synthetic.PageShape((64, 64, 3))], [{'bboxes': np.zeros((0, 4))}], 1)


@pytest.mark.parametrize("seed", list(range(14)))
def test_randomized_shapes_and_parameters(rn, seed):
    """Random page shapes, pyramid-level subsets, anchor parameter sets (1 to 30 anchors per cell: the 9-anchor
    tile kernel, the generic tile kernel and the one-anchor-per-thread kernel), class counts and GT sets (boxes on
    and across the page border, empty pages) -- every output bit for bit against the oracle.  Exercises tile
    edges, every 16-byte phase of the bulk-store write-out and the exact-division fallback."""
    rs = np.random.RandomState(1000 + seed)
    H, W = int(rs.randint(17, 420)), int(rs.randint(17, 520))
    all_levels = [3, 4, 5, 6, 7]
    k = int(rs.randint(1, 6))
    first = int(rs.randint(0, 6 - k))
    levels = all_levels[first:first + k]
    nr, ns = [(3, 3), (1, 1), (2, 2), (4, 3), (5, 6), (3, 1)][seed % 6]
    ratios = np.sort(rs.uniform(0.3, 3.0, nr)).astype(np.float32)
    scales = np.sort(rs.uniform(0.7, 1.9, ns)).astype(np.float32)
    params = rn.AnchorParameters(sizes=[int(2 ** (l + 2) * rs.uniform(0.8, 1.3)) for l in levels],
                                 strides=[2 ** l for l in levels], ratios=ratios, scales=scales)
    oparams = O.AnchorParameters(params.sizes, params.strides, ratios, scales)
    C = [1, 2, 5][seed % 3]
    B = int(rs.randint(1, 5))
    images, anns = [], []
    for b in range(B):
        h, w = int(H * rs.uniform(0.6, 1.0)) if b else H, int(W * rs.uniform(0.6, 1.0)) if b else W
        images.append(synthetic.PageShape((h, w, 3)))
        G = 0 if (b == 1 and seed % 2) else int(rs.randint(0, 41))
        x1, y1 = rs.uniform(-5, w * 0.8, G), rs.uniform(-5, h * 0.8, G)
        bw, bh = rs.uniform(2, w * 0.7, G), rs.uniform(2, h * 0.7, G)
        boxes = np.stack([x1, y1, x1 + bw, y1 + bh], axis=1).reshape(-1, 4)
        if G > 2:
            boxes[1] = boxes[0]                                   # duplicate table: first index must win
            boxes[2, 2:] = boxes[2, :2]                           # empty table
        anns.append({'bboxes': boxes, 'labels': rs.randint(0, C, G).astype(np.float64)})
    anchors = rn.anchors_for_shape((H, W, 3), pyramid_levels=levels, anchor_params=params)
    oanchors = O.anchors_for_shape((H, W, 3), pyramid_levels=levels, anchor_params=oparams)
    assert same(np.asarray(anchors), oanchors)
    want_reg, want_lab = O.anchor_targets_bbox(oanchors, images, anns, C)
    reg, lab = rn.anchor_targets_bbox(anchors, images, anns, C)
    assert same(reg, want_reg) and same(lab, want_lab)
    # the same through an explicit (N,4) array (no generation spec): one anchor per thread
    reg2, lab2 = rn.anchor_targets_bbox(np.array(oanchors), images, anns, C)
    assert same(reg2, want_reg) and same(lab2, want_lab)


@pytest.mark.parametrize("C", [1, 3])
@pytest.mark.parametrize("off_reg,off_lab", [(1, 2), (3, 0), (2, 2)])
def test_unaligned_output_tensors(rn, C, off_reg, off_lab):
    """Output tensors that are not 16-byte aligned (a view into a larger buffer) take the plain-store write-out
    instead of the TMA bulk stores: same bits."""
    hw = (131, 203)
    anchors = rn.anchors_for_shape(hw + (3,))
    oanchors = O.anchors_for_shape(hw + (3,))
    N = oanchors.shape[0]
    imgs = [synthetic.PageShape(hw + (3,)), synthetic.PageShape((120, 180, 3))]
    anns = [synthetic.gt_for_page(2, 40 + i, hw=hw, gmax=7, classes=C) for i in range(2)]
    want_reg, want_lab = O.anchor_targets_bbox(oanchors, imgs, anns, C)
    boxes, labels, counts, img_hw = rn.anchors.pack_annotations(imgs, anns, C)
    d = rn.anchors.upload_annotations(boxes, labels, counts, img_hw, torch.device("cuda"))
    big_r = torch.full((2 * N * 5 + 8,), 7.0, dtype=torch.float32, device="cuda")
    big_l = torch.full((2 * N * (C + 1) + 8,), 7.0, dtype=torch.float32, device="cuda")
    y_reg = big_r[off_reg:off_reg + 2 * N * 5].view(2, N, 5)
    y_cls = big_l[off_lab:off_lab + 2 * N * (C + 1)].view(2, N, C + 1)
    assert y_reg.data_ptr() % 16 != 0 or y_cls.data_ptr() % 16 != 0
    rn.anchors.anchor_targets_device(anchors.spec, *d, C, out=(y_reg, y_cls))
    assert same(y_reg.cpu().numpy(), want_reg) and same(y_cls.cpu().numpy(), want_lab)
    # nothing outside the views was touched
    assert float(big_r[:off_reg].sum()) == 7.0 * off_reg and float(big_r[off_reg + 2 * N * 5:].sum()) == 7.0 * (8 - off_reg)
    assert float(big_l[:off_lab].sum()) == 7.0 * off_lab and float(big_l[off_lab + 2 * N * (C + 1):].sum()) == 7.0 * (8 - off_lab)


@pytest.mark.parametrize("C,hw", [(1, (400, 650)), (3, (131, 203))])
def test_page_launch_order_does_not_change_results(rn, C, hw):
    """rn_anchor_targets_ordered: the pages may be started in any order (heaviest first in the pipeline); targets, labels,
    argmax indices and the per-page / total positive counts are the same bits as with the identity order and as the oracle's."""
    anchors = rn.anchors_for_shape(hw + (3,))
    oanchors = O.anchors_for_shape(hw + (3,))
    B = 7
    imgs = [synthetic.PageShape(hw + (3,))] * B
    anns = [synthetic.gt_for_page(2, 60 + i, hw=hw, gmax=1 + 3 * i, classes=C) for i in range(B)]
    want_reg, want_lab = O.anchor_targets_bbox(oanchors, imgs, anns, C)
    boxes, labels, counts, img_hw = rn.anchors.pack_annotations(imgs, anns, C)
    d = rn.anchors.upload_annotations(boxes, labels, counts, img_hw, torch.device("cuda"))
    heavy_first = rn.anchors.page_launch_order(boxes)
    assert sorted(heavy_first.tolist()) == list(range(B))
    area = ((boxes[:, :, 2] - boxes[:, :, 0]) * (boxes[:, :, 3] - boxes[:, :, 1])).sum(axis=1)
    assert all(area[heavy_first[i]] >= area[heavy_first[i + 1]] for i in range(B - 1))
    results = []
    for order in (None, heavy_first, heavy_first[::-1].copy(), np.roll(np.arange(B, dtype=np.int32), 3)):
        d_order = None if order is None else torch.from_numpy(np.ascontiguousarray(order)).cuda()
        total = torch.zeros(1, dtype=torch.float32, device="cuda")
        reg, lab, npos, argmax = rn.anchors.anchor_targets_device(anchors.spec, *d, C, want_argmax=True, npos_total=total,
                                                                 page_order=d_order)
        results.append((reg.cpu().numpy(), lab.cpu().numpy(), npos.cpu().numpy(), argmax.cpu().numpy(), float(total.item())))
    assert same(results[0][0], want_reg) and same(results[0][1], want_lab)
    for r in results[1:]:
        assert same(r[0], results[0][0]) and same(r[1], results[0][1])
        assert (r[2] == results[0][2]).all() and (r[3] == results[0][3]).all() and r[4] == results[0][4]


@pytest.mark.parametrize("hw,B,mixed", [((800, 1333), 6, False), ((256, 320), 5, True), ((1600, 2400), 2, False)])
def test_sparse_regression_targets(rn, hw, B, mixed):
    """rn_anchor_targets_sparse (extension for the training step): labels / counts are those of the dense call bit for
    bit, the regression rows of state == 1 anchors are the dense call's rows bit for bit (the fast quotients of the dense
    path are certified to equal the exact IEEE expression the sparse path evaluates), every other row keeps what the
    tensor held -- and the losses and gradients computed from the two forms are identical."""
    from retinanet_b200 import anchors as A
    anchors = rn.anchors_for_shape(hw + (3,))
    cfg = 2 if hw == (800, 1333) else 4
    if mixed:
        images = [synthetic.PageShape((hw[0] - 8 * (i % 3), hw[1] - 20 * (i % 2), 3)) for i in range(B)]   # border rule on
        anns = [synthetic.gt_for_page(2, i, hw=hw, gmax=6) for i in range(B)]
        anns[1] = {'bboxes': np.zeros((0, 4)), 'labels': np.zeros((0,))}                                   # an empty page
    else:
        images, anns = synthetic.training_batch(cfg, batch=B, anchors=np.asarray(anchors))
    boxes, labels, counts, img_hw = A.pack_annotations(images, anns, 1)
    d = A.upload_annotations(boxes, labels, counts, img_hw, torch.device("cuda"))
    order = torch.from_numpy(A.page_launch_order(boxes)).cuda()
    reg_d, lab_d, npos_d, _ = A.anchor_targets_device(anchors, *d, 1, page_order=order)
    N = reg_d.shape[1]
    sentinel = torch.full((B, N, 5), -123.0, dtype=torch.float32, device="cuda")
    lab_s = torch.empty_like(lab_d)
    tot = torch.zeros(1, dtype=torch.float32, device="cuda")
    reg_s, lab_s, npos_s, _ = A.anchor_targets_device(anchors, *d, 1, out=(sentinel, lab_s), npos_total=tot, page_order=order,
                                                      sparse_regression=True)
    assert torch.equal(lab_s, lab_d) and torch.equal(npos_s, npos_d) and float(tot) == float(npos_d.sum())
    fg = lab_d[:, :, 1] == 1.0
    assert int(fg.sum()) == int(npos_d.sum()) and (int(fg.sum()) > 0 or mixed)
    assert torch.equal(reg_s[fg], reg_d[fg])                          # positives: the dense rows, bit for bit
    assert bool((reg_s[~fg] == -123.0).all())                         # everything else untouched
    # the loss kernels read nothing else
    cls, reg = synthetic.training_predictions(cfg, B, N, classes=1)
    cls, reg = torch.tensor(cls, device="cuda"), torch.tensor(reg, device="cuda")
    a = rn.detection_losses(reg_d, lab_d, reg, cls, normalizer=tot, shared_state=True)
    b = rn.detection_losses(reg_s, lab_s, reg, cls, normalizer=tot, shared_state=True)
    assert all(torch.equal(x, y) for x, y in zip(a, b))


def test_sparse_targets_step_equals_dense_step(rn):
    """TargetLossStep(sparse_targets=True): in-order and overlapped schedules give the dense step's losses and gradients."""
    hw, B = (800, 1333), 4
    anchors = rn.anchors_for_shape(hw + (3,))
    images, anns = synthetic.training_batch(2, batch=B, anchors=np.asarray(anchors))
    cls, reg = synthetic.training_predictions(2, B, anchors.shape[0], classes=1)
    out = []
    for sparse in (False, True):
        step = rn.pipeline.TargetLossStep(hw + (3,), B, 22, 1, sparse_targets=sparse, peer_box=False)
        step.load_annotations(images, anns)
        step.load_predictions(torch.from_numpy(cls), torch.from_numpy(reg))
        l0 = step.run().clone()
        g0 = (step.grad_cls.clone(), step.grad_reg.clone())
        for _ in range(3):
            step.run_pipelined(overlap=True)
        out.append((l0, g0, step.losses.clone(), step.grad_cls.clone(), step.grad_reg.clone()))
    d, s = out
    assert torch.equal(d[0], s[0]) and torch.equal(d[1][0], s[1][0]) and torch.equal(d[1][1], s[1][1])
    assert torch.equal(d[2], s[2]) and torch.equal(d[3], s[3]) and torch.equal(d[4], s[4])
    with pytest.raises(ValueError):
        rn.pipeline.TargetLossStep(hw + (3,), B, 22, 3, sparse_targets=True, peer_box=False)
