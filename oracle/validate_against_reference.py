"""Pin the oracle against the reference itself and (re)generate ``tests/golden/``.

Runs ONLY in the authoring container (needs /root/reference).  The reference package cannot
be imported (TensorFlow, shapely, imgaug ... are absent), so individual functions are
AST-lifted out of the reference sources *at run time* and executed with numpy / cv2 / scipy
in scope.  Nothing is copied into this repository: the goldens hold inputs and the
reference's outputs only.

    python oracle/validate_against_reference.py            # check + write tests/golden/*.npz
"""
import ast
import os
import sys
import types

import cv2
import numpy as np
import torch
from scipy import spatial

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("KERAS_OCR_REFERENCE", "/root/reference")
GOLDEN = os.path.join(ROOT, "tests", "golden")

from oracle import craft as o_craft, imageops as o_img, synth  # noqa: E402
from keras_ocr_b200 import weights as W  # noqa: E402


def lift(path, names, scope):
    """exec the named top-level functions of ``path`` inside ``scope``."""
    with open(path) as f:
        tree = ast.parse(f.read())
    wanted = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
    missing = set(names) - {n.name for n in wanted}
    assert not missing, missing
    mod = ast.Module(body=wanted, type_ignores=[])
    exec(compile(mod, path, "exec"), scope)
    return scope


class _NoShapely:
    """Drives get_rotated_box down its own AttributeError branch (tools.py:548-550)."""

    @staticmethod
    def MultiPoint(points):
        raise AttributeError("shapely absent")


def reference_tools():
    scope = {"np": np, "cv2": cv2, "spatial": spatial, "geometry": _NoShapely, "typing": __import__("typing"),
             "tx": types.SimpleNamespace(Literal={"boxes": None, "predictions": None, "lines": None}.__class__),
             "io": __import__("io"), "os": os}
    # annotations in adjust_boxes reference tx.Literal[...]; strip annotations instead of stubbing
    with open(os.path.join(REF, "keras_ocr", "tools.py")) as f:
        tree = ast.parse(f.read())
    names = {"get_rotated_width_height", "warpBox", "get_rotated_box", "pad", "resize_image", "adjust_boxes", "fix_line", "fit", "drawBoxes"}
    body = []
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            node.returns = None
            for a in node.args.args + node.args.kwonlyargs:
                a.annotation = None
            body.append(node)
    exec(compile(ast.Module(body=body, type_ignores=[]), "tools.py", "exec"), scope)
    return types.SimpleNamespace(**scope)


def reference_detection(tools_ns):
    scope = {"np": np, "cv2": cv2, "tools": tools_ns, "typing": __import__("typing")}
    lift(os.path.join(REF, "keras_ocr", "detection.py"),
         ["compute_input", "getBoxes", "get_gaussian_heatmap", "compute_maps", "build_torch_model"], scope)
    return types.SimpleNamespace(**scope)


def check_craft(det, out):
    torch.manual_seed(0)
    model = det.build_torch_model(None)
    wts = W.synthetic_craft_weights(seed=3)
    state = {k: torch.from_numpy(v) for k, v in wts.items()}
    missing, unexpected = model.load_state_dict(state, strict=False)
    assert not unexpected, unexpected
    assert all(k.endswith("num_batches_tracked") or "slice4.40" in k or "slice4.41" in k for k in missing), missing
    rng = np.random.default_rng(5)
    worst = 0.0
    for tag, (h, w) in {"even": (96, 128), "odd": (90, 114)}.items():
        img = rng.integers(0, 256, (2, h, w, 3), dtype=np.uint8)
        x = np.stack([det.compute_input(i) for i in img])
        assert np.array_equal(x, np.stack([o_img.compute_input(i) for i in img]))
        xt = torch.from_numpy(x).permute(0, 3, 1, 2).contiguous()
        with torch.no_grad():
            ref, _ = model(xt)
            mine = o_craft.craft_forward(wts, xt)
        err = float((ref - mine).abs().max())
        worst = max(worst, err)
        print(f"  CRAFT {tag} {h}x{w}: max|ref-oracle| = {err:.2e}  (out {tuple(ref.shape)})")
        out[f"craft_{tag}_image"] = img
        out[f"craft_{tag}_scores"] = ref.numpy()
    assert worst < 1e-4, worst            # the reference's own Keras-vs-torch bar (tests/test_pytorch_keras.py:49)


def check_craft_c2(det, out):
    """BASELINE.json configs[1] at its own size: CRAFT on 8 x 768 x 768 (seeded uint8 noise, weights seed 3) through
    the reference's own PyTorch CRAFT (``build_torch_model``, detection.py:472-644).  The fp32 score maps (8,384,384,2)
    are committed as float16 (4.7 MB; quantisation 5e-4 of the range, the GPU comparison allows 2e-2); the input is
    regenerated from its seed by the test.  The oracle restatement is checked on the first two images."""
    torch.manual_seed(0)
    model = det.build_torch_model(None)
    wts = W.synthetic_craft_weights(seed=3)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in wts.items()}, strict=False)
    seed = 4
    img = np.random.default_rng(seed).integers(0, 256, (8, 768, 768, 3), dtype=np.uint8)
    scores = []
    with torch.no_grad():
        for i in range(0, 8, 2):                          # two images at a time bounds the fp32 activations (~6 GB)
            x = np.stack([det.compute_input(im) for im in img[i:i + 2]])
            xt = torch.from_numpy(x).permute(0, 3, 1, 2).contiguous()
            ref, _ = model(xt)
            scores.append(ref.numpy())
            if i == 0:
                mine = o_craft.craft_forward(wts, xt).numpy()
                err = float(np.abs(mine - scores[0]).max())
                print(f"  CRAFT 768x768: max|ref-oracle| = {err:.2e} on images 0-1")
                assert err < 1e-4, err
    scores = np.concatenate(scores)
    print(f"  CRAFT C2: scores {scores.shape}, range [{scores.min():.3f}, {scores.max():.3f}]")
    out["craft_c2_seed"] = np.array(seed)
    out["craft_c2_scores_f16"] = scores.astype(np.float16)
    out["craft_c2_checksum"] = np.array([float(scores.astype(np.float64).sum()), float(np.abs(scores).astype(np.float64).sum())])


def check_boxes(det, out):
    cases = {
        "grid32": synth.score_maps(11, 2, 384, 384, 32),
        "rot12": synth.score_maps(12, 2, 256, 320, 12),
        "dense": synth.score_maps(13, 1, 200, 300, 60),
        "blank": np.zeros((1, 64, 64, 2), np.float32),
    }
    # reference label generator as an extra, independent input (detection.py:55-62,106-198)
    heat = det.get_gaussian_heatmap(size=128, distanceRatio=1.5)
    lines = [[(np.array([[20 + 18 * i, 30], [36 + 18 * i, 30], [36 + 18 * i, 60], [20 + 18 * i, 60]], "float32"), "a")
              for i in range(6)],
             [(np.array([[60 + 22 * i, 150 + 4 * i], [80 + 22 * i, 154 + 4 * i], [76 + 22 * i, 190 + 4 * i], [56 + 22 * i, 186 + 4 * i]], "float32"), "b")
              for i in range(5)]]
    cases["refmaps"] = det.compute_maps(heat, 256, 320, lines)[np.newaxis].astype(np.float32)
    for tag, maps in cases.items():
        ref = det.getBoxes(maps)
        mine = o_img.get_boxes(maps)
        for r, m in zip(ref, mine):
            assert r.shape == m.shape, (tag, r.shape, m.shape)
            if r.size:
                assert np.array_equal(r, m), (tag, np.abs(r - m).max())
        counts = [len(r) for r in ref]
        print(f"  getBoxes {tag}: boxes per image {counts} -- identical")
        out[f"boxes_{tag}_scores"] = maps.astype(np.float32)
        out[f"boxes_{tag}_counts"] = np.array(counts, np.int32)
        flat = [r.reshape(-1, 4, 2) for r in ref if r.size]
        out[f"boxes_{tag}_quads"] = np.concatenate(flat).astype(np.float32) if flat else np.zeros((0, 4, 2), np.float32)


def check_warp(tools, out):
    rng = np.random.default_rng(21)
    gray = synth.noise_gray(rng, 480, 640)
    quads = synth.random_quads(rng, 48, 480, 640)
    crops = []
    for q in quads:
        ref = tools.warpBox(image=gray, box=q, target_height=31, target_width=200)
        mine = o_img.warp_box(gray, q)
        assert np.array_equal(ref, mine)
        rb, _ = tools.get_rotated_box(q)
        assert np.array_equal(rb, o_img.order_corners(q))
        assert tools.get_rotated_width_height(rb) == o_img.rotated_width_height(rb)
        crops.append(ref)
    print(f"  warpBox: {len(quads)} quads identical")
    out["warp_gray"] = gray
    out["warp_quads"] = quads
    out["warp_crops"] = np.stack(crops)


def check_inputs(tools, out):
    rng = np.random.default_rng(31)
    for tag, (h, w, scale, max_size) in {"x2": (120, 160, 2, 2048), "capped": (300, 500, 2, 800), "x3": (77, 93, 3, 2048)}.items():
        img = cv2.GaussianBlur(rng.integers(0, 256, (h, w, 3)).astype(np.float32), (0, 0), 1.2).clip(0, 255).astype(np.uint8)
        ref, s_ref = tools.resize_image(img, max_scale=scale, max_size=max_size)
        mine, s_mine = o_img.resize_image(img, scale, max_size)
        assert s_ref == s_mine and np.array_equal(ref, mine)
        rp = tools.pad(ref, width=ref.shape[1] + 7, height=ref.shape[0] + 3)
        assert np.array_equal(rp, o_img.pad(mine, ref.shape[1] + 7, ref.shape[0] + 3))
        gray = cv2.cvtColor(ref, code=cv2.COLOR_RGB2GRAY)
        r64 = ref.astype(np.int64)
        formula = ((9798 * r64[..., 0] + 19235 * r64[..., 1] + 3735 * r64[..., 2] + 16384) >> 15).astype(np.uint8)
        assert np.array_equal(gray, formula)
        out[f"resize_{tag}_src"] = img
        out[f"resize_{tag}_dst"] = ref
        out[f"resize_{tag}_params"] = np.array([scale, max_size, s_ref], np.float64)
        print(f"  resize/pad/gray {tag}: {img.shape} -> {ref.shape} scale {s_ref:.4f} identical")
    from keras_ocr_b200 import tools as p_tools
    for tag, (h, w, mode, cval) in {"wide": (40, 300, "letterbox", 0), "tall": (90, 120, "letterbox", 0), "crop": (50, 180, "crop", 255),
                                     "exact": (31, 200, "letterbox", 0)}.items():
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        ref = tools.fit(image=img, width=200, height=31, cval=cval, mode=mode)
        assert np.array_equal(ref, p_tools.fit(img, 200, 31, cval=cval, mode=mode))
        out[f"fit_{tag}_src"] = img
        out[f"fit_{tag}_dst"] = ref
        out[f"fit_{tag}_params"] = np.array([cval, 1 if mode == "crop" else 0])
    print("  fit (letterbox / crop): identical")
    boxes = rng.uniform(0, 100, (5, 4, 2)).astype(np.float32)
    assert np.array_equal(tools.adjust_boxes(boxes=boxes, boxes_format="boxes", scale=0.5), boxes * 0.5)
    canvas = rng.integers(0, 256, (120, 120, 3), dtype=np.uint8)
    assert np.array_equal(tools.drawBoxes(image=canvas, boxes=boxes), p_tools.drawBoxes(canvas, boxes))
    preds = [("w", b) for b in boxes]
    assert np.array_equal(tools.drawBoxes(image=canvas, boxes=preds, boxes_format="predictions", thickness=2),
                          p_tools.drawBoxes(canvas, preds, thickness=2, boxes_format="predictions"))
    print("  drawBoxes: identical")


def check_craft_keras(out):
    """The KERAS flavour of CRAFT -- the graph ``Detector()`` builds by default -- from the reference's own source
    (``build_keras_model`` + ``load_torch_weights``, detection.py:65-103, 290-468) executed on ``oracle/keras_shim.py``,
    against the oracle and against the torch-twin scores already in craft.npz."""
    import tempfile
    from oracle import keras_shim as shim

    scope = lift(os.path.join(REF, "keras_ocr", "detection.py"),
                 ["upconv", "make_vgg_block", "UpsampleLike", "build_vgg_backbone", "build_keras_model", "load_torch_weights"],
                 {"tf": shim.tf, "keras": shim.keras, "np": np, "typing": __import__("typing")})
    wts = W.synthetic_craft_weights(seed=3)
    state = {"module." + k: torch.from_numpy(np.asarray(v)) for k, v in wts.items()}
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "craft_synthetic.pth")
        torch.save(state, path)
        shim.reset()
        model = scope["build_keras_model"](weights_path=path, backbone_name="vgg")
    golden = np.load(os.path.join(GOLDEN, "craft.npz"))
    worst = 0.0
    for tag in ("even", "odd"):
        img = golden[f"craft_{tag}_image"]
        x = np.stack([o_img.compute_input(i) for i in img])
        got = model.predict(x).numpy()
        with torch.no_grad():
            mine = o_craft.craft_forward(wts, torch.from_numpy(x).permute(0, 3, 1, 2).contiguous()).numpy()
        e_oracle = float(np.abs(got - mine).max())
        e_twin = float(np.abs(got - golden[f"craft_{tag}_scores"]).max())
        worst = max(worst, e_oracle, e_twin)
        print(f"  Keras CRAFT (reference source on the shim) {tag}: vs oracle {e_oracle:.2e}, vs the reference's torch twin {e_twin:.2e}")
        out[f"craft_keras_{tag}_scores"] = got.astype(np.float32)
    assert worst < 1e-4, worst            # the reference's own Keras-vs-torch bar (tests/test_pytorch_keras.py:49)


def check_crnn(out):
    """The recognizer: the reference's own ``build_model`` / ``_transform`` / ``CTCDecoder`` source
    (recognition.py:54-350) executed on ``oracle/keras_shim.py`` (numpy ``tf`` ops, torch-backed Keras layers)
    against ``oracle/crnn.py``.  Pins the graph wiring, the STN sampler line by line and the CTC padding; the
    per-layer arithmetic is the documented Keras semantics (TensorFlow itself is not installable here)."""
    from oracle import crnn as o_crnn, keras_shim as shim

    path = os.path.join(REF, "keras_ocr", "recognition.py")
    with open(path) as f:
        tree = ast.parse(f.read())
    params = next(ast.literal_eval(n.value) for n in tree.body if isinstance(n, ast.Assign)
                  and getattr(n.targets[0], "id", "") == "DEFAULT_BUILD_PARAMS")
    scope = lift(path, ["_repeat", "_meshgrid", "_transform", "CTCDecoder", "build_model"],
                 {"tf": shim.tf, "keras": shim.keras, "np": np})
    alphabet = "0123456789abcdefghijklmnopqrstuvwxyz"
    shim.reset()
    backbone, model, training_model, prediction_model = scope["build_model"](alphabet=alphabet, **params)
    wts = W.synthetic_crnn_weights(seed=2, decisive=True)      # seeded backbone + the head fitted on rendered words
    used = shim.load_weights(wts)
    assert used == set(wts), sorted(set(wts) - used)          # every tensor found its layer, by the reference's names

    # (1) the STN sampler alone, on transforms far from the identity (clipping quirks exercised)
    rng = np.random.default_rng(7)
    feat = rng.standard_normal((3, 50, 7, 8)).astype(np.float32)
    theta = np.array([[1, 0, 0, 0, 1, 0], [0.8, 0.15, 0.1, -0.1, 1.1, -0.05], [1.3, -0.2, 0.4, 0.25, 0.7, 0.3]], np.float32)
    ref = scope["_transform"]([feat, theta])
    mine = o_crnn.stn_sample(torch.from_numpy(feat), torch.from_numpy(theta)).numpy()
    err = float(np.abs(ref - mine).max())
    print(f"  _transform (reference source on the numpy shim) vs oracle.stn_sample: max|diff| = {err:.2e}")
    assert err < 1e-4, err             # coordinates reach 50 in float32: one ulp of the grid moves a sample by ~4e-6
    assert np.abs(ref[0, -1]).max() < 1e-5 and np.abs(ref[0, :, -1]).max() < 1e-5   # identity theta: last row / column cancel to 0
    out["stn_features"], out["stn_theta"], out["stn_out"] = feat, theta, ref.astype(np.float32)

    # (2) the whole recognizer on noise crops + the word crops the oracle pipeline cuts out of two rendered pages
    from oracle.pipeline import OraclePipeline
    gray = synth.noise_gray(rng, 31, 200 * 3)
    crops = np.stack([gray[:, i * 200:(i + 1) * 200] for i in range(3)])
    pages, _ = synth.text_images(seed=21, n=2, h=192, w=384, n_words=4)
    chain = OraclePipeline(W.synthetic_craft_weights(3, textlike=True), wts, scale=2)
    batch, _ = chain.prepare(pages)
    word_crops = np.array(chain.crops(batch, chain.detect(batch)))
    assert len(word_crops) >= 6
    crops = np.concatenate([crops, word_crops]).astype(np.uint8)
    out["crnn_n_noise"] = np.array(3)
    x = (crops.astype("float32") / 255)[..., np.newaxis]                  # recognition.py:524-526
    probs_ref = model.predict(x).numpy()
    labels_ref = prediction_model.predict(x).numpy()
    feats_ref = backbone.predict(x).numpy()
    probs, taps = o_crnn.crnn_logits(wts, x, return_intermediates=True)
    labels = o_crnn.ctc_greedy(probs)
    e_l2 = float((taps["l2"] - torch.from_numpy(feats_ref)).abs().max())
    e_p = float((probs - torch.from_numpy(probs_ref)).abs().max())
    print(f"  build_model (reference source on the shim) vs oracle: backbone max|diff| = {e_l2:.2e}, softmax {e_p:.2e}, "
          f"labels equal: {np.array_equal(labels, labels_ref)}")
    assert probs_ref.shape == (len(crops), 48, 37) and labels_ref.shape == (len(crops), 48)
    # fp32 on both sides, different summation orders (torch.nn.LSTM vs the explicit gate loop); the fitted head's larger
    # weights amplify that to ~1e-4 on the LSTM outputs in [-1, 1]
    assert e_l2 < 5e-4 and e_p < 1e-4, (e_l2, e_p)
    assert np.array_equal(labels, labels_ref)
    out["crnn_crops"] = crops
    out["crnn_probs"] = probs_ref.astype(np.float32)
    out["crnn_labels"] = labels_ref.astype(np.int64)


def main():
    assert os.path.isdir(REF), f"reference not found at {REF}"
    os.makedirs(GOLDEN, exist_ok=True)
    tools = reference_tools()
    det = reference_detection(tools)
    groups = {}
    only = set(sys.argv[1:])                             # e.g. `validate_against_reference.py crnn`
    for name, fn, arg in [("craft", check_craft, det), ("boxes", check_boxes, det),
                          ("warp", check_warp, tools), ("inputs", check_inputs, tools), ("crnn", check_crnn, None),
                          ("craft_keras", check_craft_keras, None), ("craft_c2", check_craft_c2, det)]:
        if (only and name not in only) or (not only and name == "craft_c2"):      # craft_c2 (~2 min) only on request
            continue
        print(f"[{name}]")
        out = {}
        fn(*([arg, out] if arg is not None else [out]))
        groups[name] = out
    for name, out in groups.items():
        path = os.path.join(GOLDEN, f"{name}.npz")
        np.savez_compressed(path, **out)
        print(f"wrote {path} ({os.path.getsize(path) / 1e3:.0f} kB)")
    print("oracle == reference on every case")


if __name__ == "__main__":
    main()
