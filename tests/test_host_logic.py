"""CPU: host-side logic of the drop-in (no GPU): tools.fit vs the reference goldens, label decoding,
reference-API argument checks that do not need the device."""
import os

import numpy as np
import pytest

from keras_ocr_b200 import tools
from keras_ocr_b200.recognition import labels_to_text, DEFAULT_ALPHABET


@pytest.mark.parametrize("tag", ["wide", "tall", "crop", "exact"])
def test_fit_matches_reference_golden(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "inputs.npz"))
    cval, is_crop = g[f"fit_{tag}_params"]
    out = tools.fit(g[f"fit_{tag}_src"], 200, 31, cval=int(cval), mode="crop" if is_crop else "letterbox")
    assert np.array_equal(out, g[f"fit_{tag}_dst"])          # bit-exact (reference tools.py:402-452)
    with pytest.raises(NotImplementedError):
        tools.fit(g[f"fit_{tag}_src"][:20, :50], 200, 31, mode="stretch")


def test_labels_to_text_matches_reference_rule():
    rng = np.random.default_rng(0)
    rows = np.full((64, 48), -1, np.int32)
    for r in rows:
        k = rng.integers(0, 30)
        r[:k] = rng.integers(0, 36, k)
    ref = ["".join(DEFAULT_ALPHABET[i] for i in row if i not in (36, -1)) for row in rows]
    assert labels_to_text(rows) == ref
    rows[3, :6] = [1, 36, 2, -1, 3, 36]                       # blanks / padding in the middle: generic path
    ref = ["".join(DEFAULT_ALPHABET[i] for i in row if i not in (36, -1)) for row in rows]
    assert labels_to_text(rows) == ref
    assert labels_to_text(np.zeros((0, 48), np.int32)) == []
    for alphabet in ("ab", "".join(chr(c) for c in range(32, 127)), "αβγδ漢字"):   # custom alphabets (ascii LUT and generic path)
        blank = len(alphabet)
        rows = np.full((16, 48), -1, np.int32)
        for r in rows:
            k = rng.integers(0, 30)
            r[:k] = rng.integers(0, blank, k)
        ref = ["".join(alphabet[i] for i in row if i not in (blank, -1)) for row in rows]
        assert labels_to_text(rows, alphabet) == ref


def test_adjust_boxes_and_read_contracts(tmp_path):
    boxes = np.arange(16, dtype=np.float32).reshape(2, 4, 2)
    assert tools.adjust_boxes(boxes, scale=1) is boxes
    assert np.array_equal(tools.adjust_boxes(boxes, scale=0.5), boxes * 0.5)
    with pytest.raises(NotImplementedError):
        tools.adjust_boxes(boxes, scale=2, boxes_format="nope")
    with pytest.raises(AssertionError):                      # reference tools.py:34-36
        tools.read(str(tmp_path / "missing.png"))
    import cv2
    img = np.random.default_rng(1).integers(0, 256, (20, 30, 3), dtype=np.uint8)
    path = str(tmp_path / "x.png")
    cv2.imwrite(path, img[..., ::-1])
    assert np.array_equal(tools.read(path), img)             # BGR file -> RGB array
    assert tools.read(img) is img


def test_keras_h5_dataset_mapping_round_trip():
    """weights.map_keras_datasets: Keras save_weights paths (named layers by name, the auto-named localisation
    model by shape) -> this package's keys; round trip of a full synthetic CRNN checkpoint."""
    from keras_ocr_b200 import weights as W
    w = W.synthetic_crnn_weights(4)
    flat = {}
    stn_names = {"stn.conv_a": "model_1/conv2d_8", "stn.conv_b": "model_1/conv2d_9",
                 "stn.dense_a": "model_1/dense_3", "stn.dense_b": "model_1/dense_4"}
    for key, arr in w.items():
        layer, kind = key.rsplit(".", 1)
        if layer in stn_names:
            flat[f"{stn_names[layer]}/{kind}:0"] = arr
        elif layer.startswith("lstm"):
            flat[f"{layer}/{layer}/lstm_cell_7/{kind}:0"] = arr
        else:
            flat[f"{layer}/{layer}/{kind}:0"] = arr
    flat["optimizer_weights/iter:0"] = np.zeros(())           # ignored
    back = W.map_keras_datasets(flat)
    assert set(back) == set(w)
    assert all(np.array_equal(back[k], w[k]) for k in w)
    with pytest.raises(ValueError):
        W.map_keras_datasets({"model_1/conv2d_8/kernel:0": np.zeros((3, 3, 8, 8), np.float32)})
    with pytest.raises(FileNotFoundError):
        W.load_keras_h5("/nonexistent/crnn_kurapan.h5")


# ------------------------------------------------------------------- box geometry of warpBox (host side, tools.py:41-57, 533-581)
def _rect_area(r):
    return float(np.linalg.norm(r[1] - r[0]) * np.linalg.norm(r[2] - r[1]))


def test_minimum_rotated_rectangle_properties_and_cv2_area():
    """shapely is not installable offline, so the restated ``minimum_rotated_rectangle`` is checked through what defines it:
    a rectangle, containing every point, one side on a hull edge, and of the least area -- the area against
    cv2.minAreaRect (rotating calipers, an independent implementation) and against the oracle's own restatement."""
    import cv2
    from keras_ocr_b200 import tools
    from oracle import imageops
    rng = np.random.default_rng(11)
    for trial in range(200):
        n = int(rng.integers(3, 9))
        pts = rng.uniform(0, 300, (n, 2)) if trial % 3 else rng.integers(0, 40, (n, 2)).astype(np.float64)
        rect = tools.minimum_rotated_rectangle(pts)
        hull = tools._convex_hull(pts)
        if len(hull) < 3:
            assert rect is None and imageops.min_rotated_rectangle(pts) is None
            continue
        sides = np.roll(rect, -1, 0) - rect
        for i in range(4):                                   # right angles, opposite sides equal
            assert abs(np.dot(sides[i], sides[(i + 1) % 4])) <= 1e-6 * max(1.0, _rect_area(rect))
        assert np.allclose(sides[0], -sides[2], atol=1e-8) and np.allclose(sides[1], -sides[3], atol=1e-8)
        u, v = sides[0] / np.linalg.norm(sides[0]), sides[1] / np.linalg.norm(sides[1])
        a, b = (pts - rect[0]) @ u, (pts - rect[0]) @ v      # every point inside
        assert a.min() >= -1e-7 and a.max() <= np.linalg.norm(sides[0]) + 1e-7
        assert b.min() >= -1e-7 and b.max() <= np.linalg.norm(sides[1]) + 1e-7
        edges = np.roll(hull, -1, 0) - hull                  # one side lies along a hull edge
        cosines = np.abs(edges @ u) / np.linalg.norm(edges, axis=1)
        assert (np.minimum(np.abs(cosines - 1.0), np.abs(cosines)) <= 1e-9).any()
        (_, (cw, ch), _) = cv2.minAreaRect(pts.astype(np.float32))
        assert abs(_rect_area(rect) - cw * ch) <= 2e-4 * max(cw * ch, 1.0), (trial, _rect_area(rect), cw * ch)
        np.testing.assert_allclose(rect, imageops.min_rotated_rectangle(pts), atol=1e-9)


def test_get_rotated_box_orders_corners_like_the_reference():
    """tools.get_rotated_box (reference tools.py:533-581): top-left, top-right, bottom-right, bottom-left for an upright and
    a tilted rectangle whatever the order of the input corners; width / height as get_rotated_width_height (41-57)."""
    from keras_ocr_b200 import tools
    from oracle import imageops
    upright = np.array([[10, 20], [110, 20], [110, 50], [10, 50]], np.float32)
    c, s = np.cos(0.3), np.sin(0.3)
    tilted = (upright - upright.mean(0)) @ np.array([[c, s], [-s, c]], np.float32) + upright.mean(0)
    rng = np.random.default_rng(3)
    for quad in (upright, tilted.astype(np.float32)):
        for _ in range(8):
            box, _rot = tools.get_rotated_box(quad[rng.permutation(4)])
            np.testing.assert_allclose(box, quad, atol=2e-4)
            np.testing.assert_array_equal(box, imageops.order_corners(box))
            wh = tools.get_rotated_width_height(box)
            assert wh == imageops.rotated_width_height(box)
            assert wh[0] in (99, 100) and wh[1] in (29, 30)      # int() truncates the fp32-rounded side lengths


def test_rectify_boxes_keeps_rectangles_replaces_quads_and_raises_on_degenerate_boxes():
    from keras_ocr_b200 import tools
    rect = np.array([[10, 20], [110, 20], [110, 50], [10, 50]], np.float32)
    c, s = np.cos(-0.2), np.sin(-0.2)
    tilted = ((rect - rect.mean(0)) @ np.array([[c, s], [-s, c]], np.float32) + rect.mean(0)).astype(np.float32)
    trapezoid = np.array([[10, 20], [110, 25], [100, 60], [20, 50]], np.float32)
    out = tools.rectify_boxes(np.stack([rect, tilted, trapezoid]))
    np.testing.assert_array_equal(out[0], rect)              # rectangles: bit for bit (everything getBoxes emits)
    np.testing.assert_array_equal(out[1], tilted)
    assert np.abs(out[2] - trapezoid).max() > 1.0            # a general quad becomes its minimum rotated rectangle ...
    np.testing.assert_allclose(out[2], tools.get_rotated_box(trapezoid)[0])
    sides = np.roll(out[2], -1, 0) - out[2]
    assert abs(float(np.dot(sides[0], sides[1]))) <= 1e-2    # ... which is a rectangle
    with pytest.raises(ZeroDivisionError):                   # tools.py:95: scale = min(target_width / w, target_height / h)
        tools.rectify_boxes(np.array([[[5, 5], [5.4, 5], [5.4, 30], [5, 30]]], np.float32))
