"""Scoring of (text, box) predictions against ground truth: drop-in for ``keras_ocr.evaluation`` (reference
evaluation.py:13-147), consuming what ``Pipeline.recognize`` returns.  Host-side Python like the
reference; its two third-party helpers are absent offline and restated here:

* ``pyclipper`` (polygon intersection / union of ``iou_score``): restated for SIMPLE polygons -- both polygons are
  triangulated by ear clipping, every triangle pair is intersected by Sutherland-Hodgman clipping, and
  ``union = area1 + area2 - intersection``.  Vertices are truncated to int32 first, as the reference does
  (evaluation.py:33-38).  Clipper additionally rounds the intersection's vertices to integers before the areas are
  taken; that sub-pixel effect is not reproduced (PARITY UNPINNED beyond the reference's own test cases, which pass).
* ``editdistance.eval``: the Levenshtein distance, by the textbook two-row dynamic programme.
"""
import copy
import typing
import warnings

import numpy as np


def _area2(poly):
    """Twice the signed area (shoelace) of an (n,2) polygon."""
    x, y = poly[:, 0], poly[:, 1]
    return float(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1)))


def _triangulate(poly):
    """Ear clipping of a simple polygon (n,2) float64 -> list of counter-clockwise (3,2) triangles."""
    pts = [tuple(p) for p in poly]
    if _area2(np.array(pts)) < 0:
        pts.reverse()
    out = []

    def cross(o, a, b):
        return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])

    guard = 0
    while len(pts) > 3 and guard < 10000:
        guard += 1
        n = len(pts)
        for i in range(n):
            a, b, c = pts[i - 1], pts[i], pts[(i + 1) % n]
            if cross(a, b, c) <= 0:                               # reflex or degenerate corner: not an ear
                continue
            inside = any(p not in (a, b, c) and cross(a, b, p) >= 0 and cross(b, c, p) >= 0 and cross(c, a, p) >= 0 for p in pts)
            if not inside:
                out.append(np.array([a, b, c], dtype=np.float64))
                del pts[i]
                break
        else:                                                      # numerically degenerate: drop a collinear vertex
            del pts[0]
    if len(pts) == 3 and abs(cross(*pts)) > 0:
        out.append(np.array(pts, dtype=np.float64))
    return out


def _clip_convex(subject, clip):
    """Sutherland-Hodgman: the part of convex polygon ``subject`` inside counter-clockwise convex polygon ``clip``."""
    out = [tuple(p) for p in subject]
    for i in range(len(clip)):
        a, b = clip[i], clip[(i + 1) % len(clip)]
        src, out = out, []
        if not src:
            break

        def side(p):
            return (b[0] - a[0]) * (p[1] - a[1]) - (b[1] - a[1]) * (p[0] - a[0])

        for j, cur in enumerate(src):
            prev = src[j - 1]
            sc, sp = side(cur), side(prev)
            if (sc >= 0) != (sp >= 0):                             # the edge crosses the clip line
                t = sp / (sp - sc)
                out.append((prev[0] + t * (cur[0] - prev[0]), prev[1] + t * (cur[1] - prev[1])))
            if sc >= 0:
                out.append(cur)
    return np.array(out, dtype=np.float64).reshape(-1, 2)


def _as_polygon(box):
    if len(box) == 2:                                              # two corners -> axis-aligned rectangle (evaluation.py:23-30)
        (x1, y1), (x2, y2) = box
        box = [[x1, y1], [x2, y1], [x2, y2], [x1, y2]]
    return np.array(box, dtype="int32").astype(np.float64)       # int32 truncation as in the reference


def iou_score(box1, box2):
    """Intersection over union of two polygons given as lists of (x, y) vertices (or two opposite corners).
    Reference evaluation.py:13-57."""
    p1, p2 = _as_polygon(box1), _as_polygon(box2)
    a1, a2 = abs(_area2(p1)) / 2, abs(_area2(p2)) / 2
    if a1 == 0 or a2 == 0:
        warnings.warn("A box with zero area was detected.")
        return 0
    intersection = 0.0
    for t1 in _triangulate(p1):
        for t2 in _triangulate(p2):
            clipped = _clip_convex(t1, t2)
            if len(clipped) >= 3:
                intersection += abs(_area2(clipped)) / 2
    union = a1 + a2 - intersection
    return intersection / union


def edit_distance(a, b):
    """Levenshtein distance (insert / delete / substitute, unit costs) = ``editdistance.eval``."""
    if len(a) < len(b):
        a, b = b, a
    previous = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        current = [i]
        for j, cb in enumerate(b, 1):
            current.append(min(previous[j] + 1, current[j - 1] + 1, previous[j - 1] + (ca != cb)))
        previous = current
    return previous[-1]


def score(true, pred, iou_threshold=0.5, similarity_threshold=0.5, translator=None):
    """Precision / recall of predicted annotations, reference evaluation.py:60-147 (same arguments, same result
    dictionary, same matching rule: a ground-truth box is matched by EVERY prediction with IoU >= ``iou_threshold``;
    a match counts as a true positive when the normalised edit similarity of the texts reaches ``similarity_threshold``,
    else as a near true positive; ``ignore``d ground truth absorbs predictions without counting).

    ``true`` / ``pred``: {image_id: [{"text": str, "vertices": [(x, y), ...]}, ...]}.  ``predictions_to_annotations``
    converts the output of ``Pipeline.recognize``."""
    true_ids = sorted(true)
    pred_ids = sorted(pred)
    assert all(true_id == pred_id for true_id, pred_id in zip(true_ids, pred_ids)), \
        "true and pred dictionaries must have the same keys"
    results: typing.Dict[str, typing.List[dict]] = {
        "true_positives": [], "false_positives": [], "near_true_positives": [], "false_negatives": []}
    for image_id in true_ids:
        true_anns = true[image_id]
        pred_anns = copy.deepcopy(pred[image_id])
        pred_matched = set()
        for true_index, true_ann in enumerate(true_anns):
            match = None
            for pred_index, pred_ann in enumerate(pred_anns):
                if iou_score(true_ann["vertices"], pred_ann["vertices"]) < iou_threshold:
                    continue
                match = {"true_idx": true_index, "pred_idx": pred_index, "image_id": image_id}
                pred_matched.add(pred_index)
                if true_ann.get("ignore", False):
                    continue
                true_text, pred_text = true_ann["text"], pred_ann["text"]
                if translator is not None:
                    true_text, pred_text = true_text.translate(translator), pred_text.translate(translator)
                norm = max(len(true_text), len(pred_text))
                similarity = 1 if norm == 0 else 1 - edit_distance(true_text, pred_text) / norm
                results["true_positives" if similarity >= similarity_threshold else "near_true_positives"].append(match)
            if match is None and not true_ann.get("ignore", False):
                results["false_negatives"].append({"image_id": image_id, "true_idx": true_index})
        results["false_positives"].extend({"pred_index": pred_index, "image_id": image_id}
                                          for pred_index, _ in enumerate(pred_anns) if pred_index not in pred_matched)
    fns, fps = len(results["false_negatives"]), len(results["false_positives"])
    tps = len({(tp["image_id"], tp["true_idx"]) for tp in results["true_positives"]})
    precision = tps / (tps + fps)
    recall = tps / (tps + fns)
    return results, (precision, recall)


def predictions_to_annotations(prediction_groups, image_ids=None):
    """``Pipeline.recognize`` output (one list of (text, (4,2) box) per image) -> the {image_id: annotations} form
    ``score`` takes."""
    ids = range(len(prediction_groups)) if image_ids is None else image_ids
    return {image_id: [{"text": text, "vertices": np.asarray(box).tolist()} for text, box in group]
            for image_id, group in zip(ids, prediction_groups)}
