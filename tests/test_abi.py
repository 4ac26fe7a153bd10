"""CPU: the C-ABI library builds, loads and exports exactly what include/b2ocr.h declares; the
product refuses to run without a GPU (no CPU fallback)."""
import os
import re

import numpy as np
import pytest
import torch

from keras_ocr_b200 import _lib, tools, weights as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "b2ocr.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2o_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_header_symbol():
    lib = _lib.load_library()
    names = _header_functions()
    assert len(names) >= 18
    for name in names:
        assert hasattr(lib, name), name
    assert sorted(_lib.SIGNATURES) == names          # the ctypes table covers the header one to one
    assert lib.b2o_version() == 1


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from keras_ocr_b200 import detection, recognition
    with pytest.raises(_lib.B2OError):
        detection.Detector(weights=W.synthetic_craft_weights(0))
    with pytest.raises(_lib.B2OError):
        recognition.Recognizer(weights=W.synthetic_crnn_weights(0))
    with pytest.raises(_lib.B2OError):
        _lib.Context(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "keras-ocr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_weight_dictionaries_have_reference_names_and_counts():
    c = W.synthetic_craft_weights(0)
    assert c["basenet.slice1.0.weight"].shape == (64, 3, 3, 3)
    assert c["basenet.slice5.1.weight"].shape == (1024, 512, 3, 3)
    assert c["conv_cls.8.weight"].shape == (2, 16, 1, 1)
    n_conv = sum(v.size for k, v in c.items() if k.endswith(".weight") and v.ndim == 4)
    n_all = sum(v.size for k, v in c.items() if "running" not in k)
    assert n_all == 20770466                      # parameter count of the reference model
    assert n_conv * 1 > 0
    macs = sum(co * ci * k * k / (1 if "slice1.0" in n or "slice1.3" in n else 1) for n, ci, co, k, *_ in W.CRAFT_CONVS) * 0
    r = W.synthetic_crnn_weights(0)
    assert r["conv_7.kernel"].shape == (3, 3, 512, 512) and r["fc_12.kernel"].shape == (256, 37)
    assert r["stn.dense_a.kernel"].shape == (11200, 64) and r["lstm_11_back.recurrent_kernel"].shape == (128, 512)


def test_craft_mac_count_matches_survey():
    """355,720 MAC per detector-input pixel; used by bench.py's roofline."""
    res = {"basenet.slice1.0": 1, "basenet.slice1.3": 1, "basenet.slice1.7": 4, "basenet.slice1.10": 4}
    total = 0.0
    for name, cin, cout, k, *_ in W.CRAFT_CONVS:
        if name in res:
            div = res[name]
        elif name.startswith("basenet.slice2") or name == "basenet.slice3.20" or name.startswith("upconv3"):
            div = 16
        elif name.startswith("basenet.slice3") or name == "basenet.slice4.30" or name.startswith("upconv2"):
            div = 64
        elif name.startswith("basenet.slice4") or name.startswith("basenet.slice5") or name.startswith("upconv1"):
            div = 256
        else:
            div = 4                               # upconv4.*, conv_cls.* at half resolution
        total += cin * cout * k * k / div
    assert round(total) == W.CRAFT_MAC_PER_PIXEL


def test_resize_plan_matches_reference_rule():
    from oracle import imageops
    rng = np.random.default_rng(0)
    for shape, scale, max_size in [((120, 160, 3), 2, 2048), ((300, 500, 3), 2, 800), ((77, 93, 3), 3, 2048), ((2000, 900, 3), 2, 2048)]:
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        out, s = imageops.resize_image(img, scale, max_size)
        ps, ph, pw = tools.resize_plan(shape, scale, max_size)
        assert ps == s and (ph, pw) == out.shape[:2]
