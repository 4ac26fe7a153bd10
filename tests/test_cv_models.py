"""CPU: the NumPy models that the CUDA kernels follow are bit-compatible with OpenCV (tests/cvmodels.py)."""
import os

import cv2
import numpy as np

from oracle import imageops
from tests import cvmodels as M


def test_resize_model_is_bit_exact():
    rng = np.random.default_rng(0)
    for (h, w, dh, dw) in [(120, 160, 240, 320), (300, 500, 480, 800), (77, 93, 231, 279), (100, 37, 171, 64), (64, 64, 64, 64)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(cv2.resize(img, dsize=(dw, dh)), M.resize_model(img, dw, dh))


def test_warp_model_is_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "warp.npz"))
    gray = g["warp_gray"]
    bad = total = 0
    for q, ref in list(zip(g["warp_quads"], g["warp_crops"]))[:6]:
        box = imageops.order_corners(q)
        w, h = imageops.rotated_width_height(box)
        scale = min(200 / w, 31 / h)
        dst = np.array([[0, 0], [scale * w, 0], [scale * w, scale * h], [0, scale * h]]).astype("float32")
        Mx = M.get_persp(box, dst)
        np.testing.assert_allclose(Mx, cv2.getPerspectiveTransform(src=box, dst=dst), rtol=1e-9, atol=1e-9)
        dw, dh = int(scale * w), int(scale * h)
        full = np.zeros((31, 200), np.uint8)
        full[:dh, :dw] = M.warp_model(gray, Mx, dw, dh)
        bad += int((full != ref).sum())
        total += full.size
    assert bad == 0, f"{bad}/{total} pixels differ"


def _blobs(rng, count):
    for t in range(count):
        m = np.zeros((40, 60), np.uint8)
        if t % 4 == 0:
            x0, y0 = rng.integers(0, 30, 2)
            m[y0:y0 + int(rng.integers(1, 9)), x0:x0 + int(rng.integers(1, 25))] = 1
        else:
            for _ in range(int(rng.integers(1, 4))):
                c = (int(rng.integers(10, 50)), int(rng.integers(8, 32)))
                ax = (int(rng.integers(1, 20)), int(rng.integers(1, 8)))
                cv2.ellipse(m, c, ax, float(rng.uniform(0, 180)), 0, 360, 1, -1)
        n, lab = cv2.connectedComponents(m, connectivity=8)
        if n >= 2:
            yield lab == 1


def test_hull_and_min_area_rect_models():
    rng = np.random.default_rng(1)
    worst = 0.0
    for blob in _blobs(rng, 400):
        ys = np.where(blob.any(1))[0]
        xmin = [int(np.where(blob[y])[0].min()) for y in ys]
        xmax = [int(np.where(blob[y])[0].max()) for y in ys]
        cnt = cv2.findContours(blob.astype(np.uint8), cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)[-2][0]
        ref_hull = cv2.convexHull(cnt, clockwise=False).reshape(-1, 2)
        hull = M.hull_from_rows([int(y) for y in ys], xmin, xmax)
        assert np.array_equal(ref_hull, hull)                      # same vertices, same order, same start
        if len(hull) < 3:
            continue
        ref = cv2.boxPoints(cv2.minAreaRect(cnt))
        mine = M.box_points(M.min_area_rect(hull))
        # same rectangle up to the starting corner (getBoxes re-rolls it, detection.py:284)
        d = min(np.abs(np.roll(mine, s, 0) - ref).max() for s in range(4))
        worst = max(worst, d)
    assert worst < 1e-3, worst
