"""CPU: the oracle reproduces the golden vectors generated from the reference itself
(oracle/validate_against_reference.py).  These fixtures ARE the reference's outputs."""
import os

import numpy as np
import pytest
import torch

from oracle import craft, crnn, imageops
from keras_ocr_b200 import weights as W


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def test_craft_matches_reference_torch_twin(golden_dir):
    g = _load(golden_dir, "craft")
    wts = W.synthetic_craft_weights(seed=3)
    for tag in ("even", "odd"):
        img = g[f"craft_{tag}_image"]
        x = torch.from_numpy(np.stack([imageops.compute_input(i) for i in img])).permute(0, 3, 1, 2).contiguous()
        with torch.no_grad():
            mine = craft.craft_forward(wts, x).numpy()
        # bar = the reference's own Keras-vs-torch tolerance (tests/test_pytorch_keras.py:49, decimal=4)
        np.testing.assert_allclose(mine, g[f"craft_{tag}_scores"], atol=1e-4)


@pytest.mark.parametrize("tag", ["grid32", "rot12", "dense", "blank", "refmaps"])
def test_get_boxes_matches_reference(golden_dir, tag):
    g = _load(golden_dir, "boxes")
    boxes = imageops.get_boxes(g[f"boxes_{tag}_scores"])
    assert [len(b) for b in boxes] == list(g[f"boxes_{tag}_counts"])
    flat = [b.reshape(-1, 4, 2) for b in boxes if b.size]
    if flat:
        assert np.array_equal(np.concatenate(flat), g[f"boxes_{tag}_quads"])     # bit-exact
    else:
        assert g[f"boxes_{tag}_quads"].shape[0] == 0
        assert boxes[0].shape == (0,)          # np.array([]) like reference detection.py:286


def test_warp_box_matches_reference(golden_dir):
    g = _load(golden_dir, "warp")
    for q, ref in zip(g["warp_quads"], g["warp_crops"]):
        assert np.array_equal(imageops.warp_box(g["warp_gray"], q), ref)


@pytest.mark.parametrize("tag", ["x2", "capped", "x3"])
def test_resize_matches_reference(golden_dir, tag):
    g = _load(golden_dir, "inputs")
    scale, max_size, s_ref = g[f"resize_{tag}_params"]
    out, s = imageops.resize_image(g[f"resize_{tag}_src"], scale, max_size)
    assert s == s_ref and np.array_equal(out, g[f"resize_{tag}_dst"])


def test_blank_image_has_no_boxes():
    """reference tests/test_pipeline.py:9-12 at the getBoxes level: flat maps -> zero predictions."""
    assert imageops.get_boxes(np.zeros((1, 128, 128, 2), np.float32))[0].shape == (0,)


def test_ctc_greedy_semantics():
    probs = np.full((1, 6, 37), 1e-3, np.float32)
    for t, c in enumerate([5, 5, 36, 5, 7, 7]):       # repeat, blank, repeat-after-blank, new, repeat
        probs[0, t, c] = 0.9
    out = crnn.ctc_greedy(probs)
    assert out[0].tolist() == [5, 5, 7, -1, -1, -1]
    assert crnn.labels_to_text(out) == ["557"]


def test_crnn_oracle_shapes_and_backward_order():
    wts = W.synthetic_crnn_weights(seed=2)
    rng = np.random.default_rng(0)
    crops = rng.integers(0, 256, (2, 31, 200)).astype(np.float32) / 255
    with torch.no_grad():
        probs, inter = crnn.crnn_logits(wts, crops, return_intermediates=True)
    assert probs.shape == (2, 48, 37) and inter["warped"].shape == (2, 50, 7, 512)
    np.testing.assert_allclose(probs.sum(-1).numpy(), 1.0, atol=1e-5)
    # go_backwards keeps processing order: step 0 of the backward LSTM only saw the LAST input step
    w = crnn._t(wts)
    x = inter["fc_9"]
    back = crnn.lstm(w, x, "lstm_10_back", go_backwards=True)
    x2 = x.clone(); x2[:, :-1] = 0
    back2 = crnn.lstm(w, x2, "lstm_10_back", go_backwards=True)
    np.testing.assert_allclose(back[:, 0].numpy(), back2[:, 0].numpy(), atol=1e-6)


def test_stn_sampler_quirk_zero_last_row_and_column():
    """with identity theta the last row / column of the output is exactly 0."""
    feat = torch.ones(1, 50, 7, 4)
    theta = torch.tensor([[1.0, 0, 0, 0, 1.0, 0]])
    out = crnn.stn_sample(feat, theta)
    assert float(out[0, :, 6].abs().max()) == 0.0 and float(out[0, 49].abs().max()) == 0.0
    np.testing.assert_allclose(out[0, :49, :6].numpy(), 1.0, atol=1e-6)


def test_crnn_matches_reference_source_on_the_keras_shim(golden_dir):
    """tests/golden/crnn.npz = outputs of the reference's OWN ``build_model`` / ``_transform`` / ``CTCDecoder`` source
    (recognition.py:54-350) executed on oracle/keras_shim.py (written by oracle/validate_against_reference.py).
    Pins the recognizer's wiring, the STN sampler and the CTC padding of the oracle; the per-layer arithmetic is the
    documented Keras semantics (TensorFlow is not installable offline)."""
    g = np.load(os.path.join(golden_dir, "crnn.npz"))
    wts = W.synthetic_crnn_weights(seed=2, decisive=True)
    out = crnn.stn_sample(torch.from_numpy(g["stn_features"]), torch.from_numpy(g["stn_theta"]))
    assert float(np.abs(out.numpy() - g["stn_out"]).max()) < 1e-4
    with torch.no_grad():
        probs = crnn.crnn_logits(wts, g["crnn_crops"].astype(np.float32) / 255)
    assert probs.shape == g["crnn_probs"].shape and probs.shape[1:] == (48, 37) and probs.shape[0] >= 9
    assert float(np.abs(probs.numpy() - g["crnn_probs"]).max()) < 1e-4
    assert np.array_equal(crnn.ctc_greedy(probs), g["crnn_labels"])
    words = crnn.labels_to_text(g["crnn_labels"][int(g["crnn_n_noise"]):])
    assert all(3 <= len(w) <= 10 for w in words), words                # the word crops decode to whole words


def test_keras_shim_lstm_is_independent_of_the_oracle_loop():
    """The shim evaluates LSTM layers with torch.nn.LSTM; the oracle with an explicit gate loop.  Same numbers, both
    directions, so the gate order [i, f, c, o] / single bias / reversed-output conventions are cross-checked."""
    from oracle import keras_shim as shim
    wts = W.synthetic_crnn_weights(seed=4)
    x = torch.from_numpy(np.random.default_rng(1).standard_normal((2, 50, 128)).astype(np.float32))
    for name, backwards in (("lstm_10", False), ("lstm_10_back", True)):
        layer = shim.LSTM(128, go_backwards=backwards, return_sequences=True, name=name)
        layer.weights = {k: wts[f"{name}.{k}"] for k in ("kernel", "recurrent_kernel", "bias")}
        with torch.no_grad():
            ref = layer.forward(x)
            mine = crnn.lstm(crnn._t(wts), x, name, go_backwards=backwards)
        assert float((ref - mine).abs().max()) < 1e-5


def test_craft_matches_reference_keras_source_on_the_shim(golden_dir):
    """tests/golden/craft_keras.npz = the reference's KERAS CRAFT (``build_keras_model`` + ``load_torch_weights``,
    detection.py:65-103, 290-468 -- the graph ``Detector()`` builds by default) executed on oracle/keras_shim.py for the
    images of craft.npz; the oracle and the reference's torch twin both agree with it to the reference's own
    Keras-vs-torch bar (tests/test_pytorch_keras.py:49: 1e-4)."""
    g, gk = _load(golden_dir, "craft"), _load(golden_dir, "craft_keras")
    wts = W.synthetic_craft_weights(seed=3)
    for tag in ("even", "odd"):
        x = np.stack([imageops.compute_input(i) for i in g[f"craft_{tag}_image"]])
        with torch.no_grad():
            mine = craft.craft_forward(wts, torch.from_numpy(x).permute(0, 3, 1, 2).contiguous()).numpy()
        assert float(np.abs(mine - gk[f"craft_keras_{tag}_scores"]).max()) < 1e-4
        assert float(np.abs(g[f"craft_{tag}_scores"] - gk[f"craft_keras_{tag}_scores"]).max()) < 1e-4
