"""CPU: the N4 oracle (oracle/preprocess_np.py) pinned against OpenCV itself, stage by stage -- the arithmetic the reference
calls in DetectTablesUtils.py:183-262 lives in cv2, which is installed in this image (it is not part of /root/reference)."""
import os

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from oracle import preprocess_np as P  # noqa: E402

SAMPLE = "/root/reference/data/orig/sample_0717_023_orig.jpg"


def document_page(seed, H, W):
    """A synthetic scanned page: paper-coloured noise, dark text lines, ruled table boxes, a grey photo block."""
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


def cv2_pipeline(img):
    gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    th = cv2.adaptiveThreshold(gray, 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, cv2.THRESH_BINARY, 11, 2)
    chans = [cv2.distanceTransform(th, distanceType=t, maskSize=5) for t in (cv2.DIST_L2, cv2.DIST_L1, cv2.DIST_C)]
    merged = cv2.merge(chans)
    ok, buf = cv2.imencode(".png", merged)                  # imwrite's conversion of a float image: 8 bit, rounded, saturated
    return gray, th, cv2.imdecode(buf, cv2.IMREAD_UNCHANGED)


def test_constants_match_opencv():
    k = cv2.getGaussianKernel(11, -1, cv2.CV_32F).ravel()
    assert k.dtype == np.float32 and k.tobytes() == P.GAUSS11.tobytes()


def test_gray_exhaustive_sample():
    rs = np.random.RandomState(0)
    img = rs.randint(0, 256, (512, 512, 3)).astype(np.uint8)
    assert np.array_equal(P.bgr_to_gray(img), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))


@pytest.mark.parametrize("seed,H,W", [(1, 240, 320), (2, 333, 256), (3, 96, 1712), (4, 57, 8)])
def test_stages_against_opencv(seed, H, W):
    img = document_page(seed, H, W) if W >= 64 else np.random.RandomState(seed).randint(0, 256, (H, W, 3)).astype(np.uint8)
    gray, th, out = cv2_pipeline(img)
    assert np.array_equal(P.bgr_to_gray(img), gray)
    assert np.array_equal(P.adaptive_threshold(gray), th)   # widths are multiples of 8: OpenCV's vector path everywhere
    assert np.array_equal(P.distance_transforms_u8(th), out)
    assert np.array_equal(P.preprocess_page(img), out)


def test_smooth_images_hit_rounding_boundaries():
    """Slowly varying grey levels put many blurred means near x.5: the float evaluation order matters here."""
    rs = np.random.RandomState(7)
    for _ in range(3):
        base = cv2.GaussianBlur(rs.uniform(0, 255, (400, 640)).astype(np.float32), (31, 31), 0)
        gray = np.clip(base + rs.normal(0, 3, base.shape), 0, 255).astype(np.uint8)
        th = cv2.adaptiveThreshold(gray, 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, cv2.THRESH_BINARY, 11, 2)
        assert np.array_equal(P.adaptive_threshold(gray), th)
        ref = cv2.GaussianBlur(gray.astype(np.float32), (11, 11), sigmaX=0, sigmaY=0, borderType=cv2.BORDER_REPLICATE | cv2.BORDER_ISOLATED)
        assert np.array_equal(P.gaussian_blur11(gray), ref)  # bit for bit


@pytest.mark.parametrize("zeros", [1, 2, 5, 60, 4000])
def test_distance_transforms_sparse_and_dense(zeros):
    rs = np.random.RandomState(zeros)
    H, W = int(rs.randint(40, 420)), int(rs.randint(40, 420))
    th = np.full((H, W), 255, np.uint8)
    th[rs.randint(0, H, zeros), rs.randint(0, W, zeros)] = 0
    want = cv2_pipeline_from_binary(th)
    assert np.array_equal(P.distance_transforms_u8(th), want)


def cv2_pipeline_from_binary(th):
    chans = [cv2.distanceTransform(th, distanceType=t, maskSize=5) for t in (cv2.DIST_L2, cv2.DIST_L1, cv2.DIST_C)]
    return np.stack([np.clip(np.rint(c), 0, 255).astype(np.uint8) for c in chans], axis=-1)


def test_chamfer_rounding_margin():
    """Every reachable 5x5 chamfer distance below 256 is >= 7e-4 away from x.5: the 8-bit result cannot depend on the order
    in which OpenCV's raster scan adds the float move lengths (its error is < 1e-4 at these magnitudes)."""
    dx, dy = np.meshgrid(np.arange(0, 400), np.arange(0, 400))
    d = P.chamfer5(dx, dy)
    frac = np.abs(d % 1.0 - 0.5)
    assert frac[d < 256].min() > 7e-4


@pytest.mark.skipif(not os.path.isfile(SAMPLE), reason="the reference tree is not present")
def test_reference_sample_page():
    img = cv2.imread(SAMPLE)
    assert img.shape == (2200, 1712, 3)
    gray, th, out = cv2_pipeline(img)
    assert np.array_equal(P.adaptive_threshold(P.bgr_to_gray(img)), th)
    assert np.array_equal(P.distance_transforms_u8(th, reach=int(out.max()) + 2), out)


def test_bench_page_generator_is_this_one():
    """bench.py's N4 leg draws its pages from synthetic.document_page: the same generator as the parity tests'."""
    import synthetic
    assert np.array_equal(synthetic.document_page(5, 120, 96), document_page(5, 120, 96))
