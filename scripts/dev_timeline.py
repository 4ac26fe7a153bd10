"""Development probe: where the GPU idles inside back-to-back Pipeline.recognize() calls.

Every stage boundary gets a host timestamp and a CUDA event on the compute stream; an event recorded while the GPU
is idle completes at once, so `gpu - host` ~ 0 there, and a GPU time running ahead of the next host time is a bubble.

    python scripts/dev_timeline.py            # 32 pages 768x768, scale 2
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keras_ocr_b200 import weights as W
from keras_ocr_b200.detection import Detector
from keras_ocr_b200.pipeline import Pipeline
from keras_ocr_b200.recognition import Recognizer
from oracle import synth

marks = []


def mark(name):
    ev = torch.cuda.Event(enable_timing=True)
    ev.record()
    marks.append((name, time.perf_counter(), ev))


def wrap(obj, attr, name):
    fn = getattr(obj, attr)

    def inner(*a, **k):
        mark(name + " >")
        r = fn(*a, **k)
        mark(name + " <")
        return r
    setattr(obj, attr, inner)


def main():
    pages, _ = synth.text_images(seed=1000, n=int(os.environ.get("PAGES", 32)), h=768, w=768, n_words=32)
    det = Detector(weights=W.synthetic_craft_weights(3, textlike=True))
    rec = Recognizer(weights=W.synthetic_crnn_weights(2))
    pipe = Pipeline(detector=det, recognizer=rec, scale=2)
    dev = torch.from_numpy(pages).cuda()
    for _ in range(3):
        pipe.recognize(dev)
    wrap(pipe, "prepare_device", "prepare")
    wrap(det, "predict_device", "craft")
    wrap(det, "boxes_enqueue", "boxes_enqueue")
    wrap(det, "boxes_finish", "boxes_finish(sync)")
    wrap(rec, "recognize_from_boxes_device", "warp+crnn")
    wrap(pipe, "_stage_finish", "finish(sync+decode)")
    torch.cuda.synchronize()
    for rep in range(3):
        mark(f"step {rep} >")
        pipe.recognize(dev)
        mark(f"step {rep} <")
    torch.cuda.synchronize()
    t0, e0 = marks[0][1], marks[0][2]
    print(f"{'mark':28s} {'host ms':>9s} {'gpu ms':>9s}   (gpu = when the stream reached the mark)")
    for name, t, ev in marks:
        print(f"{name:28s} {1e3 * (t - t0):9.3f} {e0.elapsed_time(ev):9.3f}")
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for fn, label in ((lambda: pipe.recognize(dev), "recognize(device)"), (lambda: pipe.recognize(pages), "recognize(numpy)"),
                      (lambda: pipe.recognize_records(dev), "recognize_records(device)")):
        fn()
        torch.cuda.synchronize()
        a0.record()
        for _ in range(5):
            fn()
        a1.record()
        torch.cuda.synchronize()
        print(f"{label}: {a0.elapsed_time(a1) / 5:.2f} ms/step")


if __name__ == "__main__":
    main()
