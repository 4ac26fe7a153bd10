"""Oracle: ``Pipeline.recognize`` chained from the oracle stages (reference pipeline.py:28-75,
detection.py:745-785, recognition.py:491-537).  TEST INFRASTRUCTURE -- also the CPU baseline
(`bench.py --impl reference`, kind "port") because the reference itself cannot run without
TensorFlow (see oracle/__init__.py)."""
import time

import numpy as np
import torch

from . import craft, crnn, imageops


class OraclePipeline:
    def __init__(self, craft_weights, crnn_weights, scale=2, max_size=2048, color=False):
        self.color = color                                   # recognizer built with color=True: crops keep their RGB channels
        self.craft_weights = {k: torch.as_tensor(np.asarray(v), dtype=torch.float32) for k, v in craft_weights.items()}
        self.crnn_weights = crnn_weights
        self.scale = scale
        self.max_size = max_size
        self.timings = {}

    def detect_scores(self, images):
        """compute_input + CRAFT forward for a uint8 batch (N,H,W,3) -> (N,H/2,W/2,2) float32."""
        x = np.stack([imageops.compute_input(im) for im in images])
        xt = torch.from_numpy(x).permute(0, 3, 1, 2).contiguous()
        with torch.no_grad():
            return craft.craft_forward(self.craft_weights, xt).numpy()

    def detect(self, images, **thresholds):
        return imageops.get_boxes(self.detect_scores(images), **thresholds)

    def crops(self, images, box_groups):
        out = []
        for image, boxes in zip(images, box_groups):
            gray = image if self.color else imageops.rgb_to_gray(image)       # recognition.py:508-510
            for box in boxes:
                out.append(imageops.warp_box(gray, box))
        return out

    def recognize_from_boxes(self, images, box_groups):
        assert len(box_groups) == len(images), "You must provide the same number of box groups as images."
        crops = self.crops(images, box_groups)
        if not crops:
            return [[]] * len(images)
        with torch.no_grad():
            texts = crnn.recognize_crops(self.crnn_weights, np.array(crops))
        out, start = [], 0
        for boxes in box_groups:
            out.append(texts[start:start + len(boxes)])
            start += len(boxes)
        return out

    def prepare(self, images):
        resized = [imageops.resize_image(im, self.scale, self.max_size) for im in images]
        max_h, max_w = np.array([im.shape[:2] for im, _ in resized]).max(axis=0)
        scales = [s for _, s in resized]
        batch = np.array([imageops.pad(im, width=max_w, height=max_h) for im, _ in resized])
        return batch, scales

    def recognize(self, images, detection_kwargs=None):
        t0 = time.perf_counter()
        batch, scales = self.prepare(images)
        t1 = time.perf_counter()
        scores = self.detect_scores(batch)
        t2 = time.perf_counter()
        box_groups = imageops.get_boxes(scores, **(detection_kwargs or {}))
        t3 = time.perf_counter()
        predictions = self.recognize_from_boxes(batch, box_groups)
        t4 = time.perf_counter()
        box_groups = [np.array(b) * (1 / s) if s != 1 else b for b, s in zip(box_groups, scales)]
        self.timings = {"prepare": t1 - t0, "craft": t2 - t1, "get_boxes": t3 - t2, "warp_crnn": t4 - t3}
        return [list(zip(p, b)) for p, b in zip(predictions, box_groups)]
