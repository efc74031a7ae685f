"""Seeded synthetic pages for the tests, the golden-vector generator and ``bench.py``.

Distributions follow SURVEY.md §8(d): only page *shapes* matter (no pixels are touched by the
path), GT tables are wide boxes, training predictions sit at the prior (p ~ 0.01), inference
predictions are a low-score background plus clusters of high-score anchors planted around each
GT table.  Seeds: ``1234 + 1000*config + page_index`` with ``np.random.RandomState``.

This module is pure numpy host code; it never touches the oracle or the CUDA library.
"""
import numpy as np

# (H, W), max GT per page, classes, pages per batch  -- BASELINE.json "configs"
CONFIGS = {
    1: dict(hw=(800, 1333), gmax=3, classes=1, batch=1),
    2: dict(hw=(800, 1333), gmax=20, classes=1, batch=16),
    3: dict(hw=(800, 1333), gmax=20, classes=1, batch=64),
    4: dict(hw=(1600, 2400), gmax=20, classes=1, batch=4),
    5: dict(hw=(800, 1333), gmax=100, classes=80, batch=16),
}


class PageShape(object):
    """Stand-in for an image array: the path only ever reads ``.shape``
    (reference ``model/anchors.py:85-87``)."""

    def __init__(self, shape):
        self.shape = tuple(int(s) for s in shape)


def page_seed(config, page_index):
    return 1234 + 1000 * int(config) + int(page_index)


def level_shapes(hw, levels=(3, 4, 5, 6, 7)):
    return [((hw[0] + 2 ** l - 1) // 2 ** l, (hw[1] + 2 ** l - 1) // 2 ** l) for l in levels]


def num_anchors(hw, levels=(3, 4, 5, 6, 7), per_cell=9):
    return int(sum(h * w for h, w in level_shapes(hw, levels)) * per_cell)


def gt_for_page(config, page_index, hw=None, gmax=None, classes=None, anchors=None):
    """One page's annotations ``{'bboxes': (G,4) f64, 'labels': (G,) f64}``."""
    cfg = CONFIGS[config]
    H, W = hw or cfg['hw']
    gmax = gmax or cfg['gmax']
    classes = classes or cfg['classes']
    rs = np.random.RandomState(page_seed(config, page_index))
    G = int(rs.randint(1, gmax + 1))
    scale = 800.0 / 1700.0                      # resize factor applied by the data generator
    H0, W0 = H / scale, W / scale
    w = rs.uniform(0.2 * W0, 0.9 * W0, G)
    h = rs.uniform(0.05 * H0, 0.5 * H0, G)
    x1 = rs.uniform(0, 1, G) * (W0 - w)
    y1 = rs.uniform(0, 1, G) * (H0 - h)
    boxes = np.stack([x1, y1, x1 + w, y1 + h], axis=1) * scale
    labels = rs.randint(0, classes, G).astype(np.float64)
    if anchors is not None and rs.uniform() < 0.05:
        # adversarial: snap one GT onto the top half of an anchor (IoU ~ exactly 0.5) and
        # duplicate it so that two GT boxes tie
        a = anchors[int(rs.randint(0, anchors.shape[0]))]
        snapped = np.array([a[0], a[1], a[2], a[1] + (a[3] - a[1]) / 2])
        boxes = np.concatenate([boxes, snapped[None], snapped[None]], axis=0)
        labels = np.concatenate([labels, [0.0, float(classes - 1)]])
    return {'bboxes': boxes, 'labels': labels}


def training_batch(config, batch=None, hw=None, mixed_widths=False, anchors=None, first_page=0):
    """``(image_group, annotations_group)`` for the training-target path."""
    cfg = CONFIGS[config]
    batch = batch or cfg['batch']
    H, W = hw or cfg['hw']
    images, anns = [], []
    for i in range(batch):
        pw = W if (not mixed_widths or i % 2 == 0) else int(round(W * 0.8))
        images.append(PageShape((H, pw, 3)))
        anns.append(gt_for_page(config, first_page + i, hw=(H, pw), anchors=anchors))
    return images, anns


def training_predictions(config, batch, n_anchors, classes=None, first_page=0):
    """Head outputs at initialisation: ``cls = sigmoid(N(-4.595, 1))`` (prior 0.01,
    reference ``model/defineModel.py:78``), ``reg ~ N(0, 1)``; float32."""
    classes = classes or CONFIGS[config]['classes']
    rs = np.random.RandomState(page_seed(config, first_page) + 500)
    logits = rs.normal(-4.595, 1.0, (batch, n_anchors, classes))
    cls = (1.0 / (1.0 + np.exp(-logits))).astype(np.float32)
    reg = rs.normal(0.0, 1.0, (batch, n_anchors, 4)).astype(np.float32)
    return cls, reg


def _iou_f64(a, g):
    iw = np.maximum(0, np.minimum(a[:, 2], g[2]) - np.maximum(a[:, 0], g[0]))
    ih = np.maximum(0, np.minimum(a[:, 3], g[3]) - np.maximum(a[:, 1], g[1]))
    inter = iw * ih
    return inter / ((a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1]) + (g[2] - g[0]) * (g[3] - g[1]) - inter)


def inference_predictions(config, batch, anchors, annotations_group, classes=None, first_page=0,
                          planted_per_gt=200):
    """Head outputs of a trained model, synthetic: background ``cls = sigmoid(N(-6, 1.5))``
    (~2 % of anchors above 0.05), ``reg ~ N(0, 0.5)``; around each GT table ``planted_per_gt``
    overlapping anchors get scores U(0.3, 0.99) and regressions that decode close to the table."""
    classes = classes or CONFIGS[config]['classes']
    n = anchors.shape[0]
    rs = np.random.RandomState(page_seed(config, first_page) + 700)
    cls = (1.0 / (1.0 + np.exp(-rs.normal(-6.0, 1.5, (batch, n, classes))))).astype(np.float32)
    reg = rs.normal(0.0, 0.5, (batch, n, 4)).astype(np.float32)
    aw = anchors[:, 2] - anchors[:, 0]
    ah = anchors[:, 3] - anchors[:, 1]
    for b in range(batch):
        ann = annotations_group[b]
        for g, lab in zip(ann['bboxes'], ann['labels']):
            near = np.nonzero(_iou_f64(anchors, g) > 0.3)[0]
            if near.size == 0:
                continue
            pick = rs.choice(near, size=min(planted_per_gt, near.size), replace=False)
            cls[b, pick, int(lab)] = rs.uniform(0.3, 0.99, pick.size).astype(np.float32)
            t = np.stack([(g[0] - anchors[pick, 0]) / aw[pick], (g[1] - anchors[pick, 1]) / ah[pick],
                          (g[2] - anchors[pick, 2]) / aw[pick], (g[3] - anchors[pick, 3]) / ah[pick]], axis=1) / 0.2
            reg[b, pick] = (t + rs.normal(0, 0.15, t.shape)).astype(np.float32)
    return cls, reg


def inference_predictions_torch(config, batch, anchors, annotations_group, classes=None, first_page=0,
                                planted_per_gt=200, device="cuda"):
    """The distributions of :func:`inference_predictions` generated on the GPU with torch (seeded ``torch.Generator``;
    not the same samples as the numpy version): used by ``bench.py`` for the large configurations, where the numpy
    generator would take longer than the benchmark (C = 80: 257 M samples per 16 pages).  Returns device tensors
    ``cls (B, N, C)``, ``reg (B, N, 4)`` float32."""
    import torch
    classes = classes or CONFIGS[config]['classes']
    gen = torch.Generator(device=device)
    gen.manual_seed(page_seed(config, first_page) + 700)
    a = torch.as_tensor(np.asarray(anchors), dtype=torch.float64, device=device)
    n = a.shape[0]
    cls = torch.empty((batch, n, classes), dtype=torch.float32, device=device)
    for b in range(batch):                                # page by page: the temporaries stay small
        cls[b] = torch.sigmoid(torch.randn((n, classes), generator=gen, device=device, dtype=torch.float32) * 1.5 - 6.0)
    reg = torch.randn((batch, n, 4), generator=gen, device=device, dtype=torch.float32) * 0.5
    aw, ah = a[:, 2] - a[:, 0], a[:, 3] - a[:, 1]
    area = aw * ah
    for b in range(batch):
        ann = annotations_group[b]
        g = torch.as_tensor(np.asarray(ann['bboxes'], dtype=np.float64).reshape(-1, 4), device=device)
        if g.shape[0] == 0:
            continue
        labs = torch.as_tensor(np.asarray(ann['labels']).astype(np.int64), device=device)
        iw = (torch.minimum(a[:, None, 2], g[None, :, 2]) - torch.maximum(a[:, None, 0], g[None, :, 0])).clamp_(min=0)
        ih = (torch.minimum(a[:, None, 3], g[None, :, 3]) - torch.maximum(a[:, None, 1], g[None, :, 1])).clamp_(min=0)
        inter = iw * ih
        iou = inter / (area[:, None] + ((g[:, 2] - g[:, 0]) * (g[:, 3] - g[:, 1]))[None, :] - inter)
        r = torch.rand(iou.shape, generator=gen, device=device)
        r = torch.where(iou > 0.3, r, torch.full_like(r, -1.0))
        k = min(planted_per_gt, n)
        val, pick = torch.topk(r, k, dim=0)               # (k, G): up to k random near anchors per table
        ok = val > 0
        gi = torch.arange(g.shape[0], device=device)[None, :].expand_as(pick)[ok]
        pi = pick[ok]
        score = (torch.rand(pi.shape, generator=gen, device=device) * 0.69 + 0.3).to(torch.float32)
        cls[b, pi, labs[gi]] = score
        t = torch.stack([(g[gi, 0] - a[pi, 0]) / aw[pi], (g[gi, 1] - a[pi, 1]) / ah[pi],
                         (g[gi, 2] - a[pi, 2]) / aw[pi], (g[gi, 3] - a[pi, 3]) / ah[pi]], dim=1) / 0.2
        reg[b, pi] = (t + torch.randn(t.shape, generator=gen, device=device, dtype=torch.float64) * 0.15).to(torch.float32)
    return cls, reg


def document_page(seed, H, W):
    """A synthetic scanned page (uint8 BGR): paper-coloured noise, dark text lines, ruled table boxes, a grey photo block --
    the input of the page preprocessing (DetectTablesUtils.py:183-262); the generator of tests/test_oracle_preprocess.py."""
    rs = np.random.RandomState(seed)
    img = np.clip(rs.normal(235, 6, (H, W, 3)), 0, 255)
    for y in range(20, H - 20, 14):
        if rs.uniform() < 0.8:
            x0, x1 = int(rs.randint(10, W // 3)), int(rs.randint(W // 2, W - 10))
            for x in range(x0, x1, 7):
                if rs.uniform() < 0.75:
                    img[y:y + int(rs.randint(4, 9)), x:x + int(rs.randint(2, 6))] = rs.uniform(10, 90)
    for _ in range(3):
        y0, x0 = int(rs.randint(0, H - 60)), int(rs.randint(0, W - 80))
        h, w = int(rs.randint(30, 60)), int(rs.randint(40, 80))
        img[y0:y0 + h, x0:x0 + 2] = 30; img[y0:y0 + h, x0 + w:x0 + w + 2] = 30
        img[y0:y0 + 2, x0:x0 + w] = 30; img[y0 + h:y0 + h + 2, x0:x0 + w + 2] = 30
    y0, x0 = int(rs.randint(0, H - 50)), int(rs.randint(0, W - 50))
    img[y0:y0 + 48, x0:x0 + 48] = np.clip(rs.normal(128, 25, (48, 48, 3)), 0, 255)
    return img.astype(np.uint8)
