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
