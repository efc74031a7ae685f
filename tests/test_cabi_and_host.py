"""CPU: the C-ABI library loads and exports every symbol include/rn_b200.h declares (no compute calls
without a GPU); host-side logic (GT packing, sharding, argument validation, config round trips)."""
import ctypes
import os
import re

import numpy as np
import pytest

import synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pkg():
    import retinanet_b200
    return retinanet_b200


@pytest.fixture(scope="module")
def built_lib(pkg):
    import importlib.util
    spec = importlib.util.spec_from_file_location("rn_build", os.path.join(ROOT, "retinanet-for-table-detection_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()


def declared_symbols():
    with open(os.path.join(ROOT, "include", "rn_b200.h")) as fh:
        text = re.sub(r"/\*.*?\*/", "", fh.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(rn_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(built_lib, pkg):
    lib = ctypes.CDLL(built_lib)
    names = declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), "librn_b200.so does not export %s" % n
    # the Python binding table covers exactly the declared entry points
    assert sorted(pkg._lib.SIGNATURES) == names
    lib.rn_version.restype = ctypes.c_int
    assert lib.rn_version() == 100
    assert pkg._lib.load().rn_loss_workspace_bytes() > 0
    assert pkg._lib.load().rn_filter_workspace_bytes(2, 1000, 3, 1, 1000, 300) > 2 * 3 * 1000 * 8    # key slab: 8 bytes per (page, class, anchor)


def test_bad_arguments_return_error_codes(built_lib, pkg):
    """Argument validation happens before any CUDA call, so it is testable without a GPU."""
    lib = pkg._lib.load()
    rc = lib.rn_clip_boxes(None, 10, 1.0, 1.0, None, None)
    assert rc == -1 and b"NULL" in lib.rn_last_error()
    hw = (ctypes.c_int * 2)(4, 4); st = (ctypes.c_int * 1)(8)
    rc = lib.rn_anchors_f64(ctypes.c_void_p(16), hw, st, 99, 9, ctypes.c_void_p(16), None)
    assert rc == -1 and b"num_levels" in lib.rn_last_error()
    rc = lib.rn_nms(None, None, 5, 5000, 0.5, ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.c_void_p(16), 1 << 20, None)
    assert rc == -1 and b"max_output" in lib.rn_last_error()


def test_product_path_fails_loudly_without_gpu(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg._lib.RnError):
        pkg.anchors_for_shape((64, 64, 3))
    with pytest.raises(pkg._lib.RnError):
        pkg.focal()(np.zeros((1, 4, 2), np.float32), np.zeros((1, 4, 1), np.float32))
    with pytest.raises(pkg._lib.RnError):
        pkg.RegressBoxes()([np.zeros((1, 4, 4), np.float32), np.zeros((1, 4, 4), np.float32)])


def test_product_does_not_import_oracle():
    pkg_dir = os.path.join(ROOT, "retinanet-for-table-detection_b200")
    for fn in os.listdir(pkg_dir):
        if fn.endswith(".py"):
            with open(os.path.join(pkg_dir, fn)) as fh:
                src = fh.read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn


def test_pack_annotations(pkg):
    imgs = [synthetic.PageShape((800, 1333, 3)), synthetic.PageShape((700, 1000, 3)), synthetic.PageShape((800, 1333, 3))]
    anns = [{'bboxes': np.array([[1., 2, 3, 4], [5, 6, 7, 8]]), 'labels': np.array([0., 2.9])},
            {'bboxes': np.zeros((0, 4)), 'labels': np.zeros((0,))},
            {'bboxes': np.array([[9., 9, 19, 19]]), 'labels': np.array([-1.])}]
    boxes, labels, counts, hw = pkg.anchors.pack_annotations(imgs, anns, 3)
    assert boxes.shape == (3, 2, 4) and boxes.dtype == np.float64 and list(counts) == [2, 0, 1]
    assert list(labels[0]) == [0, 2] and labels[2, 0] == 3          # astype(int) truncation; -1 wraps to the state column
    assert hw.tolist() == [[800, 1333], [700, 1000], [800, 1333]]
    with pytest.raises(IndexError):
        pkg.anchors.pack_annotations(imgs[:1], [{'bboxes': np.ones((1, 4)), 'labels': np.array([4.])}], 3)
    with pytest.raises(AssertionError):
        pkg.anchors.pack_annotations(imgs, anns[:2], 3)
    with pytest.raises(AssertionError):
        pkg.anchors.pack_annotations(imgs[:1], [{'labels': np.zeros(0)}], 3)


def test_host_constants_and_configs(pkg):
    from oracle import anchors_np as O
    assert pkg.AnchorParameters_default.ratios.dtype == np.float32
    assert pkg.AnchorParameters_default.num_anchors() == 9
    for size in (16, 32, 100, 512):
        assert pkg.generate_anchors(size).tobytes() == O.generate_anchors(size).tobytes()
    spec = pkg.anchors.make_spec((800, 1333, 3))
    assert spec.num_anchors == 200700 and spec.level_hw.tolist() == [[100, 167], [50, 84], [25, 42], [13, 21], [7, 11]]
    assert pkg.anchors.make_spec((1600, 2400, 3)).num_anchors == 719523
    fd = pkg.FilterDetections(nms=False, max_detections=100, name='filtered_detections')
    cfg = fd.get_config()
    assert cfg['nms'] is False and cfg['max_detections'] == 100 and cfg['parallel_iterations'] == 32
    assert pkg.FilterDetections(**{k: v for k, v in cfg.items()}).get_config() == cfg
    assert fd.compute_output_shape([(4, 1000, 4), (4, 1000, 2), (4, 1000, 7)]) == [(4, 100, 4), (4, 100), (4, 100), (4, 100, 7)]
    assert pkg.Anchors(32, 8, ratios=[0.5, 1, 2], scales=[1, 1.5]).compute_output_shape((2, 10, 12, 256)) == (2, 10 * 12 * 6, 4)
    assert set(pkg.custom_objects) == {'RegressBoxes', 'FilterDetections', 'Anchors', 'ClipBoxes'}
    with pytest.raises(ValueError):
        pkg.RegressBoxes(std=0.2)
    with pytest.raises(ValueError):
        pkg.focal(bce="nope")


def test_shard_pages(pkg):
    for pages, world in ((64, 8), (16, 3), (5, 8), (1, 1)):
        spans = [pkg.distributed.shard_pages(pages, r, world) for r in range(world)]
        covered = [p for lo, hi in spans for p in range(lo, hi)]
        assert covered == list(range(pages))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def test_balanced_shards(pkg):
    """Load-aware sharding: every page exactly once, equal page counts, near-equal total weight."""
    rs = np.random.RandomState(9)
    for world, pages in ((8, 128), (4, 64), (2, 33), (3, 7)):
        w = rs.randint(1, 23, pages)
        shards = pkg.distributed.balanced_shards(w, world)
        assert sorted(sum(shards, [])) == list(range(pages)) and all(s == sorted(s) for s in shards)
        assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
        totals = [int(w[s].sum()) for s in shards]
        contiguous = [int(w[lo:hi].sum()) for lo, hi in (pkg.distributed.shard_pages(pages, r, world) for r in range(world))]
        assert max(totals) - min(totals) <= max(w) and max(totals) <= max(contiguous)
    assert pkg.distributed.balanced_shards([5, 1], 1) == [[0, 1]]


def test_page_cost_tracks_exact_overlap_count(pkg):
    """distributed.page_cost: fixed per-anchor work + the closed-form number of overlapping (anchor, table) pairs.
    The pair estimate must stay within a few per cent of the exact count (oracle anchors, numpy); an empty page costs
    exactly the fixed part."""
    import synthetic
    from oracle import anchors_np as OA
    hw = (800, 1333)
    anchors = OA.anchors_for_shape(hw + (3,))
    _, anns = synthetic.training_batch(2, batch=6, anchors=anchors)
    est = pkg.distributed.page_cost(anns, hw, fixed=0.0)
    exact = []
    for a in anns:
        n = 0
        for g in np.asarray(a['bboxes']):
            iw = np.minimum(anchors[:, 2], g[2]) - np.maximum(anchors[:, 0], g[0])
            ih = np.minimum(anchors[:, 3], g[3]) - np.maximum(anchors[:, 1], g[1])
            n += int(((iw > 0) & (ih > 0)).sum())
        exact.append(n)
    for e, x in zip(est, exact):
        assert abs(e - x) <= 0.06 * x + 2000, (e, x)
    assert np.corrcoef(est, exact)[0, 1] > 0.995
    empty = {'bboxes': np.zeros((0, 4)), 'labels': np.zeros((0,))}
    full = pkg.distributed.page_cost(anns[:1] + [empty], hw)
    assert full[1] == pytest.approx(pkg.distributed.PAIRS_PER_FIXED * anchors.shape[0]) and full[0] > full[1]
    shards = pkg.distributed.balanced_shards(pkg.distributed.page_cost(anns, hw), 2)
    assert sorted(sum(shards, [])) == list(range(6))


def test_pack_annotations_fast_and_ragged_paths_agree(pkg):
    """anchors.pack_annotations: the batch-wide fast path (every page a (g,4) / (g,) array pair), the per-page path
    (lists, extra columns, scalar label broadcast, 1-D empty boxes) and the `out=` form give what a plain loop over the
    reference's conversions gives; errors keep their types."""
    A = pkg.anchors

    class Shape(object):
        def __init__(self, shape):
            self.shape = shape

    rs = np.random.RandomState(3)
    pages, images = [], []
    for g in (3, 0, 7, 1, 5):
        pages.append({'bboxes': rs.rand(g, 4) * 100, 'labels': rs.randint(-2, 2, g).astype(np.float64)})
        images.append(Shape((60 + g, 80 + g, 3)))

    def loop(pages, images, C):
        G = max(1, max(len(p['labels']) if np.ndim(p['labels']) else 1 for p in pages), max(np.asarray(p['bboxes']).shape[0] for p in pages))
        boxes, labels = np.zeros((len(pages), G, 4)), np.zeros((len(pages), G), np.int32)
        counts = np.array([np.asarray(p['bboxes']).shape[0] for p in pages], np.int32)
        for b, p in enumerate(pages):
            g = int(counts[b])
            if g:
                boxes[b, :g] = np.asarray(p['bboxes'], dtype=np.float64).reshape(g, -1)[:, :4]
                lab = np.asarray(p['labels']).astype(int).reshape(-1)[:g]
                labels[b, :g] = np.where(lab < 0, lab + C + 1, lab)
        hw = np.array([[im.shape[0], im.shape[1]] if im.shape else [np.iinfo(np.int32).max] * 2 for im in images], np.int32)
        return boxes[:, :max(1, counts.max())], labels[:, :max(1, counts.max())], counts, hw

    def same(got, want):
        return all(np.array_equal(g, w) and g.dtype == w.dtype for g, w in zip(got, want))

    assert same(A.pack_annotations(images, pages, 2), loop(pages, images, 2))                      # fast path
    ragged = [dict(p) for p in pages]
    ragged[0] = {'bboxes': np.hstack([pages[0]['bboxes'], np.ones((3, 1))]).tolist(), 'labels': pages[0]['labels'].tolist()}
    ragged[1] = {'bboxes': np.array([]), 'labels': np.array([])}                                   # 1-D empty
    ragged[3] = {'bboxes': pages[3]['bboxes'], 'labels': np.array(pages[3]['labels'][0])}          # 0-d label
    assert same(A.pack_annotations(images + [Shape(())], ragged + [ragged[1]], 2),
                loop(pages + [pages[1]], images + [Shape(())], 2))
    out = (np.full((5, 9, 4), 7.0), np.full((5, 9), 7, np.int32), np.zeros(5, np.int32), np.zeros((5, 2), np.int32))
    got = A.pack_annotations(images, pages, 2, out=out)
    want = loop(pages, images, 2)
    assert got[0] is out[0] and np.array_equal(out[0][:, :7], want[0]) and not out[0][:, 7:].any() and not out[1][:, 7:].any()
    assert np.array_equal(out[1][:, :7], want[1]) and np.array_equal(out[2], want[2]) and np.array_equal(out[3], want[3])
    with pytest.raises(ValueError):
        A.pack_annotations(images, pages, 2, out=(np.zeros((5, 4, 4)), np.zeros((5, 4), np.int32), out[2], out[3]))   # 7 GT do not fit
    with pytest.raises(IndexError):
        A.pack_annotations(images[:1], [{'bboxes': np.ones((1, 4)), 'labels': np.array([3.0])}], 2)
    with pytest.raises(AssertionError):
        A.pack_annotations(images[:2], pages[:1], 2)
    with pytest.raises(AssertionError):
        A.pack_annotations(images[:1], [{'bboxes': np.ones((1, 4))}], 2)
