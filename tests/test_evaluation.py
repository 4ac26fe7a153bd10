"""CPU: keras_ocr_b200.evaluation (reference evaluation.py:13-147) -- the reference's own test cases
(tests/test_evaluation.py:4-11), a rasterised cross-check of the polygon IoU, the edit distance, and the score() rules."""
import numpy as np
import pytest

from keras_ocr_b200 import evaluation as E, tools


def test_iou_score_reference_cases():
    box1 = [(0, 0), (100, 0), (100, 100), (0, 100)]
    assert E.iou_score(box1, [(50, 50), (100, 50), (100, 100), (50, 100)]) == 0.25          # reference test, exact
    assert E.iou_score(box1, [(100, 100), (200, 100), (200, 200), (100, 200)]) == 0.0
    assert E.iou_score([(0, 0), (10, 10)], [(5, 5), (15, 15)]) == pytest.approx(25 / 175)   # two-corner form
    with pytest.warns(UserWarning):
        assert E.iou_score(box1, [(5, 5), (5, 5), (5, 5), (5, 5)]) == 0


def test_iou_score_matches_rasterised_polygons():
    import cv2
    rng = np.random.default_rng(0)

    def quad(concave=False):
        c, a = rng.uniform(60, 140, 2), rng.uniform(0, np.pi)
        w, h = rng.uniform(20, 80, 2)
        pts = np.array([[-w, -h], [w, -h], [w, h], [-w, h]]) / 2
        if concave:
            pts = np.array([[-w, -h], [0, -h / 4], [w, -h], [w, h], [-w, h]]) / 2           # notch in the top edge
        rot = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])
        return (pts @ rot.T + c).astype(np.int32)

    for k in range(120):
        q1, q2 = quad(k % 3 == 0), quad(k % 5 == 0)
        s = 8                                                     # supersampled even-odd fill as the reference area
        m1, m2 = np.zeros((200 * s, 200 * s), np.uint8), np.zeros((200 * s, 200 * s), np.uint8)
        cv2.fillPoly(m1, [q1 * s], 1); cv2.fillPoly(m2, [q2 * s], 1)
        ref = (m1 & m2).sum() / (m1 | m2).sum()
        assert abs(E.iou_score(q1, q2) - ref) <= 5e-3, (q1.tolist(), q2.tolist())


def test_edit_distance():
    assert E.edit_distance("kitten", "sitting") == 3 and E.edit_distance("", "abc") == 3 and E.edit_distance("abc", "abc") == 0
    assert E.edit_distance("flaw", "lawn") == 2 and E.edit_distance("a", "") == 1


def test_score_rules_on_pipeline_output():
    box = lambda x, y: np.array([[x, y], [x + 100, y], [x + 100, y + 30], [x, y + 30]], np.float32)
    predictions = [[("hello", box(0, 0)), ("w0rld", box(0, 50)), ("spurious", box(300, 300)), ("skipme", box(0, 200))], []]
    pred = E.predictions_to_annotations(predictions)
    true = {0: [{"text": "hello", "vertices": box(2, 1).tolist()},                       # matched, same text
                {"text": "world", "vertices": box(0, 52).tolist()},                      # matched, 1 edit of 5 -> similar
                {"text": "missed", "vertices": box(500, 500).tolist()},                  # false negative
                {"text": "whatever", "vertices": box(0, 200).tolist(), "ignore": True}], # absorbs "skipme", counts for nothing
            1: []}
    results, (precision, recall) = E.score(true, pred)
    assert [m["true_idx"] for m in results["true_positives"]] == [0, 1]
    assert results["false_negatives"] == [{"image_id": 0, "true_idx": 2}]
    assert results["false_positives"] == [{"pred_index": 2, "image_id": 0}]
    assert precision == pytest.approx(2 / 3) and recall == pytest.approx(2 / 3)
    _, (p2, r2) = E.score(true, pred, similarity_threshold=0.9)    # "w0rld" vs "world": 0.8 < 0.9 -> near true positive
    assert p2 == pytest.approx(1 / 2) and r2 == pytest.approx(1 / 2)
    import string
    upper = E.predictions_to_annotations([[("HELLO!", box(0, 0))]])
    tr = str.maketrans(string.ascii_uppercase, string.ascii_lowercase, string.punctuation)
    _, (p3, r3) = E.score({0: [{"text": "hello", "vertices": box(0, 0).tolist()}]}, upper, translator=tr)
    assert (p3, r3) == (1.0, 1.0)
    with pytest.raises(AssertionError):
        E.score({0: []}, {1: []})


def test_draw_boxes_formats():
    image = np.zeros((60, 80, 3), np.uint8)
    quad = np.array([[10, 10], [60, 10], [60, 40], [10, 40]], np.float32)
    a = tools.drawBoxes(image, quad[None], thickness=1)
    b = tools.drawBoxes(image, [("word", quad)], thickness=1, boxes_format="predictions")
    c = tools.drawBoxes(image, [[(quad, "w")]], thickness=1, boxes_format="lines")
    assert a.sum() > 0 and np.array_equal(a, b) and np.array_equal(a, c) and image.sum() == 0
    assert tools.drawBoxes(image, []) is image
