"""Pipeline: drop-in for ``keras_ocr.pipeline.Pipeline`` (reference pipeline.py:7-75)."""
import numpy as np
import torch

from . import detection, recognition, tools


class Pipeline:
    """A wrapper for a combination of detector and recognizer.

    Args:
        detector: The detector to use (default: ``detection.Detector()``)
        recognizer: The recognizer to use (default: ``recognition.Recognizer()``)
        scale: The scale factor to apply to input images
        max_size: The maximum single-side dimension of images for inference.

    Any object with ``detect`` / ``recognize_from_boxes`` can be injected, as in the reference
    (pipeline.py:18-26, 62-65); with this package's own Detector and Recognizer the whole call
    stays on the GPU between the host->device copy of the images and the device->host copy of
    (counts, boxes, labels).
    """

    def __init__(self, detector=None, recognizer=None, scale=2, max_size=2048, inflight=1):
        if detector is None:
            detector = detection.Detector()
        if recognizer is None:
            recognizer = recognition.Recognizer()
        self.scale = scale
        self.detector = detector
        self.recognizer = recognizer
        self.max_size = max_size
        # inflight > 1: batches of >= inflight * min_chunk images are processed as `inflight` sub-batches,
        # software-pipelined on one stream: while the host waits for the box counts of one sub-batch (the path's
        # one data-dependent synchronisation) or decodes its labels, the GPU already runs the next one.  Results
        # are identical to the unsplit batch (images are independent; the padding is that of the whole batch).
        # Measured on the 32-page bench: ~2 % faster device-resident, but the half-size recurrent / box kernels
        # lose what the overlap gains and the end-to-end number does not move, hence the default of 1.
        self.inflight = inflight
        self.min_chunk = 4
        self.last_stats = {}
        self._h2d_stream = None

    def _native(self):
        return isinstance(self.detector, detection.Detector) and isinstance(self.recognizer, recognition.Recognizer)

    def _upload(self, array):
        """Host uint8 array -> CUDA tensor through pinned memory on a side stream, so the copy overlaps kernels
        already queued on the compute stream; the compute stream waits for the copy's event."""
        det = self.detector
        if self._h2d_stream is None:
            self._h2d_stream = torch.cuda.Stream(device=det.device)
        main = torch.cuda.current_stream(det.device)
        pinned = torch.from_numpy(np.ascontiguousarray(array))
        if not pinned.is_pinned():                       # callers that already hold pinned pages skip the staging copy
            pinned = pinned.pin_memory()
        if self.inflight <= 1:                           # nothing queued to overlap with: plain stream-ordered copy
            return pinned.to(det.device, non_blocking=True)
        with torch.cuda.stream(self._h2d_stream):
            t = pinned.to(det.device, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self._h2d_stream)
        t.record_stream(main)
        main.wait_event(done)
        return t

    def prepare_device(self, images, pad_to=None):
        """resize_image + pad (pipeline.py:44-57) on the GPU.  Returns ((N,H,W,3) u8 CUDA tensor, scales).
        ``pad_to``: (height, width) of the padded batch when ``images`` is part of a larger batch."""
        det = self.detector
        plans = [tools.resize_plan(image.shape, self.scale, self.max_size) for image in images]
        scales = [p[0] for p in plans]
        hp, wp = pad_to if pad_to is not None else (max(p[1] for p in plans), max(p[2] for p in plans))
        n = len(images)
        stream = torch.cuda.current_stream(det.device).cuda_stream
        batch = torch.empty((n, hp, wp, 3), dtype=torch.uint8, device=det.device)
        same = isinstance(images, np.ndarray) and images.ndim == 4
        h2d = 0
        if isinstance(images, torch.Tensor):
            # sources already resident in HBM (bench.py's device-resident leg): no copy
            assert images.is_cuda and images.dtype == torch.uint8 and images.dim() == 4
            src_all, same = images.contiguous(), True
        elif same:
            src_all = self._upload(images)
            h2d = src_all.numel()
        for i, image in enumerate(images):
            if same:
                src = src_all[i]
            else:
                assert image.ndim == 3 and image.shape[2] == 3 and image.dtype == np.uint8, "images must be HxWx3 uint8"
                src = self._upload(image)
                h2d += src.numel()
            _, hr, wr = plans[i]
            det.ctx.resize_pad(src.data_ptr(), image.shape[0], image.shape[1], hr, wr, batch.data_ptr(), i, hp, wp, stream)
        self.last_stats["h2d_bytes"] = self.last_stats.get("h2d_bytes", 0) + int(h2d)
        return batch, scales

    # ---------------------------------------------------------------- the three stages of one sub-batch
    def _stage_detect(self, images, pad_to, thresholds):
        batch, scales = self.prepare_device(images, pad_to)
        scores = self.detector.predict_device(batch)
        return {"batch": batch, "scales": scales, "boxes_state": self.detector.boxes_enqueue(scores, **thresholds)}

    def _stage_recognize(self, st):
        det, rec = self.detector, self.recognizer
        boxes, counts = det.boxes_finish(st.pop("boxes_state"))
        labels = rec.recognize_from_boxes_device(st["batch"], boxes, counts)
        st["counts"] = counts
        st["boxes_host"] = torch.empty(boxes.shape, dtype=boxes.dtype, pin_memory=True)
        st["boxes_host"].copy_(boxes, non_blocking=True)
        if labels is not None:
            st["labels_host"] = torch.empty(labels.shape, dtype=labels.dtype, pin_memory=True)
            st["labels_host"].copy_(labels, non_blocking=True)
        st["keep"] = (boxes, labels)                     # alive until the copies have run
        st["done"] = torch.cuda.Event()
        st["done"].record(torch.cuda.current_stream(det.device))

    def _stage_finish(self, st):
        st["done"].synchronize()
        boxes_host, counts = st["boxes_host"].numpy(), st["counts"]
        d2h = boxes_host.nbytes + counts.nbytes
        if "labels_host" in st:
            labels_host = st["labels_host"].numpy()
            d2h += labels_host.nbytes
            texts = recognition.labels_to_text(labels_host, self.recognizer.alphabet)
        else:
            texts = []
        self.last_stats["d2h_bytes"] += int(d2h)
        out, start = [], 0
        for i, (c, scale) in enumerate(zip(counts, st["scales"])):
            c = int(c)
            group = boxes_host[i, :c].copy()
            if scale != 1:
                group = tools.adjust_boxes(boxes=group, boxes_format="boxes", scale=1 / scale)
            out.append(list(zip(texts[start:start + c], group)))
            start += c
        return out

    def recognize(self, images, detection_kwargs=None, recognition_kwargs=None):
        """Run the pipeline on one or multiple images (reference pipeline.py:28-75).

        Returns a list (one entry per image) of lists of (text, box) tuples, boxes (4,2) float32 in
        the coordinates of the *input* image.
        """
        if not isinstance(images, (np.ndarray, torch.Tensor)):
            images = [tools.read(image) for image in images]
        if detection_kwargs is None:
            detection_kwargs = {}
        if recognition_kwargs is None:
            recognition_kwargs = {}
        if not self._native():
            return self._recognize_generic(images, detection_kwargs, recognition_kwargs)
        thresholds = {k: detection_kwargs[k] for k in ("detection_threshold", "text_threshold", "link_threshold",
                                                        "size_threshold") if k in detection_kwargs}
        n = len(images)
        self.last_stats = {"h2d_bytes": 0, "d2h_bytes": 0}
        if n == 0:
            return []
        plans = [tools.resize_plan(image.shape, self.scale, self.max_size) for image in images]
        pad_to = (max(p[1] for p in plans), max(p[2] for p in plans))      # of the WHOLE batch (pipeline.py:48-57)
        k = max(1, min(int(self.inflight), n // self.min_chunk))
        bounds = [n * i // k for i in range(k + 1)]
        states, out = [], []
        for step in range(k + 2):                        # detect(i) | recognize(i-1) | finish(i-2)
            if step < k:
                states.append(self._stage_detect(images[bounds[step]:bounds[step + 1]], pad_to, thresholds))
            if 1 <= step <= k:
                self._stage_recognize(states[step - 1])
            if step >= 2:
                out.extend(self._stage_finish(states[step - 2]))
                states[step - 2] = None
        return out

    def _recognize_generic(self, images, detection_kwargs, recognition_kwargs):
        """Reference flow for injected (duck-typed) detectors / recognizers: host arrays between stages."""
        import cv2

        resized = []
        for image in images:
            scale, hr, wr = tools.resize_plan(image.shape, self.scale, self.max_size)
            resized.append((cv2.resize(image, dsize=(wr, hr)), scale))
        max_height, max_width = np.array([im.shape[:2] for im, _ in resized]).max(axis=0)
        scales = [s for _, s in resized]
        padded = []
        for im, _ in resized:
            canvas = np.zeros((max_height, max_width, 3), dtype=im.dtype) + 255
            canvas[: im.shape[0], : im.shape[1]] = im
            padded.append(canvas)
        batch = np.array(padded)
        box_groups = self.detector.detect(images=batch, **detection_kwargs)
        prediction_groups = self.recognizer.recognize_from_boxes(images=batch, box_groups=box_groups, **recognition_kwargs)
        box_groups = [tools.adjust_boxes(boxes=boxes, boxes_format="boxes", scale=1 / scale) if scale != 1 else boxes
                      for boxes, scale in zip(box_groups, scales)]
        return [list(zip(predictions, boxes)) for predictions, boxes in zip(prediction_groups, box_groups)]
