"""GPU parity of N4 (rn_preprocess_pages): byte-exact against the oracle (pinned to OpenCV in tests/test_oracle_preprocess.py)
and against OpenCV itself, which the image provides."""
import numpy as np
import pytest
import torch

from oracle import preprocess_np as P
from tests.test_oracle_preprocess import cv2_pipeline, document_page

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")


@pytest.mark.parametrize("B,H,W", [(1, 240, 320), (3, 333, 256), (2, 96, 1712), (1, 57, 8), (1, 700, 1000)])
def test_pages_vs_oracle_and_opencv(rn, B, H, W):
    imgs = np.stack([document_page(10 * B + i, H, W) if W >= 64 else
                     np.random.RandomState(i).randint(0, 256, (H, W, 3)).astype(np.uint8) for i in range(B)])
    out, binary = rn.preprocess.preprocess_pages(imgs, return_binary=True)
    out, binary = out.cpu().numpy(), binary.cpu().numpy()
    for b in range(B):
        gray, th, want = cv2_pipeline(imgs[b])
        assert np.array_equal(binary[b], th) and np.array_equal(binary[b], P.adaptive_threshold(P.bgr_to_gray(imgs[b])))
        assert np.array_equal(out[b], want)
        if H * W <= 120000:
            assert np.array_equal(out[b], P.preprocess_page(imgs[b]))
    assert np.array_equal(rn.preprocess.preprocess_page(imgs[0]), out[0])


def test_blank_and_sparse_pages(rn):
    """No dark pixel at all (every distance saturates), a single one, and a page whose only ink is one far corner."""
    H, W = 300, 640
    white = np.full((H, W, 3), 255, np.uint8)
    one = white.copy(); one[150, 320] = 0
    corner = white.copy(); corner[0:3, 0:3] = 0
    for img in (white, one, corner):
        gray, th, want = cv2_pipeline(img)
        got = rn.preprocess.preprocess_page(img)
        assert np.array_equal(got, want)


def test_smooth_pages_rounding_boundaries(rn):
    """Grey ramps put the blurred mean near x.5 at many pixels: the evaluation order of the float blur decides them."""
    rs = np.random.RandomState(3)
    base = cv2.GaussianBlur(rs.uniform(0, 255, (480, 640)).astype(np.float32), (31, 31), 0)
    gray = np.clip(base + rs.normal(0, 3, base.shape), 0, 255).astype(np.uint8)
    img = np.repeat(gray[:, :, None], 3, axis=2)            # B = G = R: the grey conversion is the identity
    _, th, want = cv2_pipeline(img)
    out, binary = rn.preprocess.preprocess_pages(img, return_binary=True)
    assert np.array_equal(binary.cpu().numpy(), th) and np.array_equal(out.cpu().numpy(), want)


def test_full_size_page_properties(rn):
    """The reference's page size (2200 x 1712): GPU == OpenCV, device output == host output, batch == single pages."""
    imgs = np.stack([document_page(77 + i, 2200, 1712) for i in range(2)])
    out = rn.preprocess.preprocess_pages(torch.from_numpy(imgs).cuda()).cpu().numpy()
    for b in range(2):
        assert np.array_equal(out[b], cv2_pipeline(imgs[b])[2])
        assert np.array_equal(out[b], rn.preprocess.preprocess_page(imgs[b]))
    with pytest.raises(ValueError):
        rn.preprocess.preprocess_pages(np.zeros((4, 4), np.uint8))
