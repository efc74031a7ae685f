"""N1 / N3: the batching steps before K1 and the post-step after K5.

CPU part: the oracle restatement (oracle/generator_np.py) and the product's host functions against the golden
vectors produced by the reference's own Generator (tests/golden/make_golden_generator.py), the CSV round trip.
GPU part: compute_targets (K1 behind the generator seam) and the rescale + score-cut epilogue, bit for bit."""
import os
import warnings

import numpy as np
import pytest

from oracle import generator_np as OG

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "generator_half.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLDEN)


class Page(object):
    def __init__(self, shape):
        self.shape = tuple(int(v) for v in shape)


def _filter_case(gold):
    pages = [Page(s) for s in gold['flt_shapes']]
    anns = [{'bboxes': gold['flt_boxes'][i].copy(), 'labels': gold['flt_labels'][i].copy()} for i in range(len(pages))]
    return pages, anns


@pytest.mark.parametrize("impl", ["oracle", "product"])
def test_filter_annotations_golden(gold, impl):
    import retinanet_b200 as rn
    fn = OG.filter_annotations if impl == "oracle" else rn.generator.filter_annotations
    pages, anns = _filter_case(gold)
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        _, kept = fn(pages, anns, list(range(len(pages))))
    assert len(caught) == int(gold['flt_warnings'])
    for i, a in enumerate(kept):
        assert np.array_equal(a['bboxes'], gold['flt_out_boxes%d' % i]) and np.array_equal(a['labels'], gold['flt_out_labels%d' % i])
        assert kept is anns and a['bboxes'].shape[0] == a['labels'].shape[0]
    # nothing to drop -> untouched, no warning; an image without boxes is fine
    ok = [{'bboxes': np.array([[1., 1., 5., 5.]]), 'labels': np.array([0.])}, {'bboxes': np.zeros((0, 4)), 'labels': np.zeros((0,))}]
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        fn([Page((10, 10, 3)), Page((10, 10, 3))], ok, None)
    assert not caught and ok[0]['bboxes'].shape == (1, 4)


@pytest.mark.parametrize("impl", ["oracle", "product"])
def test_compute_inputs_golden(gold, impl):
    import retinanet_b200 as rn
    fn = OG.compute_inputs if impl == "oracle" else rn.generator.compute_inputs
    imgs = [gold['inp_img%d' % i] for i in range(len(gold['inp_shapes']))]
    got = fn(imgs)
    assert got.dtype == np.float32 and got.tobytes() == gold['inp_out'].tobytes() and got.shape == gold['inp_out'].shape


def _targets_case(gold):
    pages = [Page(s) for s in gold['tgt_shapes']]
    anns = [{'bboxes': gold['tgt_boxes%d' % i], 'labels': gold['tgt_labels%d' % i]} for i in range(len(pages))]
    return pages, anns


def test_compute_targets_oracle_golden(gold):
    pages, anns = _targets_case(gold)
    neg, pos = gold['tgt_overlaps']
    reg, lab = OG.compute_targets(pages, anns, 2, negative_overlap=float(neg), positive_overlap=float(pos))
    assert reg.tobytes() == gold['tgt_out_reg'].tobytes() and lab.tobytes() == gold['tgt_out_lab'].tobytes()
    # the narrower pages ignore anchors beyond their own border although the anchors cover the batch-max shape
    assert (lab[2, :, -1] == -1).sum() > (lab[1, :, -1] == -1).sum()


def test_oracle_vs_live_reference_generator():
    from oracle import ref_loader
    if not ref_loader.reference_available():
        pytest.skip("reference tree not present (GPU box)")
    gen_mod = ref_loader.load_reference_generator()
    g = gen_mod.Generator.__new__(gen_mod.Generator)
    rs = np.random.RandomState(11)
    pages = [Page((int(rs.randint(40, 90)), int(rs.randint(40, 90)), 3)) for _ in range(4)]
    mk = lambda: [{'bboxes': rs.uniform(-5, 95, (12, 4)).round(1), 'labels': rs.randint(0, 3, 12).astype(np.float64)} for _ in pages]
    a = mk()
    b = [{k: v.copy() for k, v in d.items()} for d in a]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        g.filter_annotations(pages, a, list(range(4)))
        OG.filter_annotations(pages, b, list(range(4)))
    for x, y in zip(a, b):
        assert np.array_equal(x['bboxes'], y['bboxes']) and np.array_equal(x['labels'], y['labels'])


def test_rescale_and_cut_oracle():
    boxes = np.arange(2 * 5 * 4, dtype=np.float32).reshape(2, 5, 4) * 1.37
    scores = np.array([[0.9, 0.7, 0.6, 0.59, -1], [0.95, 0.9, 0.8, 0.7, 0.65]], np.float32)
    out, counts = OG.rescale_and_cut(boxes, scores, [0.8, 1.6662])
    assert counts.tolist() == [3, 5] and out.dtype == np.float32
    assert np.array_equal(out[0], boxes[0] / np.float32(0.8)) and np.array_equal(out[1], (boxes[1] / 1.6662).astype(np.float32))


def test_csv_round_trip(tmp_path):
    import retinanet_b200 as rn
    path = tmp_path / "ann.csv"
    path.write_text("image_id,xmin,ymin,xmax,ymax,label\n"
                    "a.png,1,2,30,40,table\nb.png,5,6,7,8,0\na.png,10.5,20,300,400,table\n")
    pages = rn.postprocess.read_annotations_csv(str(path), class_ids={"table": 0})
    assert list(pages) == ["a.png", "b.png"]
    assert pages["a.png"]['bboxes'].tolist() == [[1, 2, 30, 40], [10.5, 20, 300, 400]] and pages["a.png"]['labels'].tolist() == [0, 0]
    assert pages["b.png"]['bboxes'].dtype == np.float64 and pages["b.png"]['labels'].tolist() == [0]
    with pytest.raises(ValueError):
        bad = tmp_path / "bad.csv"
        bad.write_text("h\nx.png,1,2,3\n")
        rn.postprocess.read_annotations_csv(str(bad))
    out = tmp_path / "det.csv"
    boxes = np.array([[[1.9, 2.2, 30.7, 40.1], [0, 0, 1, 1]], [[5, 6, 7, 8], [0, 0, 0, 0]]], np.float32)
    rn.postprocess.write_detections_csv(str(out), ["a.png", "b.png"], boxes, np.array([[0.9, 0.1], [0.8, -1]], np.float32),
                                        np.array([[0, 0], [0, -1]]), np.array([1, 1]), {0: "table"})
    back = rn.postprocess.read_annotations_csv(str(out), class_ids={"table": 0})
    assert back["a.png"]['bboxes'].tolist() == [[1, 2, 30, 40]] and back["b.png"]['bboxes'].tolist() == [[5, 6, 7, 8]]


# ---------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("output", ["numpy", "torch"])
def test_compute_targets_gpu_golden(rn, gold, output):
    pages, anns = _targets_case(gold)
    neg, pos = gold['tgt_overlaps']
    reg, lab = rn.generator.compute_targets(pages, anns, 2, negative_overlap=float(neg), positive_overlap=float(pos), output=output)
    if output == "torch":
        assert reg.is_cuda and lab.is_cuda
        reg, lab = reg.cpu().numpy(), lab.cpu().numpy()
    assert reg.tobytes() == gold['tgt_out_reg'].tobytes() and lab.tobytes() == gold['tgt_out_lab'].tobytes()


@pytest.mark.gpu
def test_rescale_and_cut_gpu(rn):
    import torch
    rs = np.random.RandomState(3)
    B, M = 7, 300
    boxes = rs.uniform(0, 1333, (B, M, 4)).astype(np.float32)
    scores = -np.sort(-rs.uniform(0, 1, (B, M)).astype(np.float32), axis=1)
    scores[0, :] = 0.99                 # nothing below the cut -> M
    scores[1, :] = 0.1                  # everything below -> 0
    scores[2, 5:] = -1.0                # padding
    scores[3, 10] = np.float32(0.6)     # exactly the cut is kept (strict <)
    scores[3, 11:] = np.float32(0.59999996)
    scale = rs.uniform(0.4, 2.5, B)
    want_b, want_c = OG.rescale_and_cut(boxes, scores, scale)
    got_b, got_c = rn.postprocess.rescale_and_cut(torch.tensor(boxes).cuda(), torch.tensor(scores).cuda(), scale)
    assert got_c.cpu().numpy().tolist() == want_c.tolist() and want_c[0] == M and want_c[1] == 0 and want_c[3] == 11
    assert got_b.cpu().numpy().tobytes() == want_b.tobytes()
    # scalar scale, numpy in
    got_b2, got_c2 = rn.postprocess.rescale_and_cut(boxes, scores, 0.8)
    w2, c2 = OG.rescale_and_cut(boxes, scores, 0.8)
    assert got_b2.cpu().numpy().tobytes() == w2.tobytes() and got_c2.cpu().numpy().tolist() == c2.tolist()
