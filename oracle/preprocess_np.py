"""CPU oracle for N4, the page-image producer of the reference (TEST INFRASTRUCTURE, see oracle/__init__.py):

    DetectTablesUtils.py:183-262  preProcessTrainValImages / preProcessSampleImages
        img = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)                                            (:208, :246)
        img = cv2.adaptiveThreshold(img, 255, ADAPTIVE_THRESH_GAUSSIAN_C, THRESH_BINARY, 11, 2)  (:209, :247)
        b, g, r = cv2.distanceTransform(img, DIST_L2 | DIST_L1 | DIST_C, maskSize=5)            (:212-214, :250-252)
        cv2.imwrite(target, cv2.merge((b, g, r)))                                               (:217-218, :255-256)

The arithmetic lives in a third-party dependency (OpenCV, un-pinned by the reference; 4.13.0 in this image), restated here
from its published algorithms and PINNED against cv2 itself: ``tests/test_oracle_preprocess.py`` compares every stage with
the library on document-like and random images and, when /root/reference is present, on the reference's sample page.

* BGR2GRAY, 8-bit: ``(B * 3735 + G * 19235 + R * 9798 + 2^14) >> 15``  (OpenCV's 15-bit fixed-point coefficients).
* adaptiveThreshold, Gaussian: the image goes to float32, is blurred with the separable 11-tap kernel
  ``getGaussianKernel(11, -1, CV_32F)`` (sigma 2.0), BORDER_REPLICATE; ``mean = saturate_u8(round_half_even(blur))``;
  ``dst = 255 if gray - mean > -2 else 0``.  The float evaluation ORDER is part of the result (a mean that lands on x.5
  flips a pixel): OpenCV's vector path accumulates the row pass left to right with fused multiply-adds
  (``acc = k0 * p0; acc = fma(p_j, k_j, acc)``) and the column pass symmetrically
  (``acc = k5 * c; acc = fma(p_{+j} + p_{-j}, k_{5+j}, acc)``).  That is what is restated.  (Its scalar tail -- the last
  ``W mod 8`` columns when the width is not a multiple of 8 -- rounds differently in the build of this image; parity is
  pinned for widths that are a multiple of 8, like the reference's 2200x1712 pages.  fma is emulated in float64: the
  product of two float32 is exact there; the double rounding this leaves has probability ~2^-29 per operation.)
* distanceTransform with maskSize 5: DIST_L2 is the 5x5 CHAMFER distance (moves 1, 1.4f, 2.1969f), not the Euclidean
  one; DIST_L1 and DIST_C use the exact 3x3 masks.  Without obstacles the two-pass raster scan yields the closed forms
  ``L1 = dx + dy``, ``C = max(dx, dy)`` and, with M = max(dx, dy), m = min(dx, dy):
  ``L2c = 2.1969f m + (M - 2m)`` if M >= 2m else ``2.1969f (M - m) + 1.4f (2m - M)``, minimised over the zero pixels.
  OpenCV sums the float moves along a raster path, so its float output differs from the closed form in the last bits --
  but ``imwrite`` converts to 8 bit (round half to even, saturate), and every reachable distance below 256 is at least
  7e-4 away from x.5, so the 8-bit result does not depend on the summation order (checked exhaustively in the tests).
  For every metric the distance is monotone in dx, so the minimum over a row of zeros is attained at the horizontally
  nearest one: ``DT(x, y) = min over y' of D(g(x, y'), |y - y'|)`` with g the in-row distance to the nearest zero.
* output: uint8 (H, W, 3) in cv2.merge order (channel 0 = L2 chamfer, 1 = L1, 2 = C) -- what ``imwrite`` encodes.
"""
import numpy as np

GAUSS11 = np.array([0x3c10612b, 0x3cde5c35, 0x3d855a85, 0x3df92326, 0x3e353f0f, 0x3e4d6105,
                    0x3e353f0f, 0x3df92326, 0x3d855a85, 0x3cde5c35, 0x3c10612b], dtype=np.uint32).view(np.float32)
CHAMFER_B = np.float32(1.4)
CHAMFER_C = np.float32(2.1969)
NO_ZERO = 0xFFFF                      # in-row distance of a row without a zero


def bgr_to_gray(img):
    """cv2.cvtColor(img, COLOR_BGR2GRAY) for uint8 (H, W, 3)."""
    b, g, r = (img[..., i].astype(np.int64) for i in range(3))
    return ((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15).astype(np.uint8)


def _fma(a, b, c):
    return (a.astype(np.float64) * np.float64(b) + c.astype(np.float64)).astype(np.float32)


def gaussian_blur11(gray):
    """The float32 blur inside cv2.adaptiveThreshold(..., ADAPTIVE_THRESH_GAUSSIAN_C, ..., 11, ...) (vector-path order)."""
    k = GAUSS11
    src = gray.astype(np.float32)
    H, W = src.shape
    p = np.pad(src, ((0, 0), (5, 5)), mode='edge')
    acc = (k[0] * p[:, 0:W]).astype(np.float32)
    for j in range(1, 11):
        acc = _fma(p[:, j:j + W], k[j], acc)
    q = np.pad(acc, ((5, 5), (0, 0)), mode='edge')
    out = (k[5] * q[5:5 + H]).astype(np.float32)
    for j in range(1, 6):
        out = _fma((q[5 + j:5 + j + H] + q[5 - j:5 - j + H]).astype(np.float32), k[5 + j], out)
    return out


def adaptive_threshold(gray):
    """cv2.adaptiveThreshold(gray, 255, ADAPTIVE_THRESH_GAUSSIAN_C, THRESH_BINARY, 11, 2)."""
    mean = np.clip(np.rint(gaussian_blur11(gray)), 0, 255).astype(np.int32)
    return np.where(gray.astype(np.int32) - mean > -2, 255, 0).astype(np.uint8)


def row_distance(binary):
    """g(x, y): distance from x to the nearest zero pixel of row y (NO_ZERO when the row has none), uint16-ranged ints."""
    H, W = binary.shape
    x = np.arange(W, dtype=np.int64)[None, :]
    zero = binary == 0
    left = np.maximum.accumulate(np.where(zero, x, -10 ** 9), axis=1)                  # last zero at or before x
    right = np.minimum.accumulate(np.where(zero, x, 10 ** 9)[:, ::-1], axis=1)[:, ::-1]  # first zero at or after x
    return np.minimum(np.minimum(x - left, right - x), NO_ZERO).astype(np.int64)


def chamfer5(dx, dy):
    """The 5x5 chamfer distance of an offset (float64 evaluation of the float32 move lengths)."""
    M, m = np.maximum(dx, dy).astype(np.float64), np.minimum(dx, dy).astype(np.float64)
    b, c = float(CHAMFER_B), float(CHAMFER_C)
    return np.where(M >= 2 * m, c * m + (M - 2 * m), c * (M - m) + b * (2 * m - M))


def distance_transforms_u8(binary, reach=256):
    """saturate_u8(round(cv2.distanceTransform(binary, DIST_L2 / DIST_L1 / DIST_C, 5))) as (H, W, 3) uint8.
    Rows further than `reach` away cannot lower a distance below 256 (every metric is >= dy), i.e. cannot change the
    8-bit result."""
    H, W = binary.shape
    if not (binary == 0).any():
        # no dark pixel at all: OpenCV leaves FLT_MAX everywhere, and the 8-bit conversion of imwrite (cvRound overflows to
        # INT_MIN, which saturates to 0) turns that into a BLACK page, not a white one
        return np.zeros((H, W, 3), np.uint8)
    g = row_distance(binary)
    big = np.float64(1e9)
    best = [np.full((H, W), big), np.full((H, W), big), np.full((H, W), big)]
    for dy in range(0, min(reach, H - 1) + 1):
        for sgn in ((1,) if dy == 0 else (1, -1)):
            lo, hi = (dy, H) if sgn > 0 else (0, H - dy)        # rows y that have a row y - sgn*dy inside the image
            gy = g[lo - sgn * dy:hi - sgn * dy]
            has = gy < NO_ZERO
            gx = np.where(has, gy, 0)
            cand = (np.where(has, chamfer5(gx, dy), big), np.where(has, (gx + dy).astype(np.float64), big),
                    np.where(has, np.maximum(gx, dy).astype(np.float64), big))
            for k in range(3):
                best[k][lo:hi] = np.minimum(best[k][lo:hi], cand[k])
    return np.stack([np.clip(np.rint(b), 0, 255).astype(np.uint8) for b in best], axis=-1)


def preprocess_page(img_bgr):
    """DetectTablesUtils.py:246-256 for one uint8 BGR page -> the uint8 (H, W, 3) image imwrite encodes."""
    return distance_transforms_u8(adaptive_threshold(bgr_to_gray(img_bgr)))
