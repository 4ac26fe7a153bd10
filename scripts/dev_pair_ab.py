"""Development A/B of a context-creation switch, both settings alternated in ONE process on ONE box: prints whether the
score maps / recognised texts are bit-identical and the step / conv-kernel times of both.

    python scripts/dev_pair_ab.py                        # B2O_TC_PAIR 0 (single-CTA tiles) vs 1 (CTA pairs, the default)
    python scripts/dev_pair_ab.py B2O_TC_BOX16 0 1       # any other switch: VARIABLE off-value on-value
    python scripts/dev_pair_ab.py B2O_TC_PAIR 1 2        # generic-tile pairs on top of the default
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keras_ocr_b200 import weights as W
from keras_ocr_b200.detection import Detector
from keras_ocr_b200.pipeline import Pipeline
from keras_ocr_b200.recognition import Recognizer
from oracle import synth


VAR, OFF, ON = (sys.argv[1:4] + ["B2O_TC_PAIR", "0", "1"][len(sys.argv) - 1:])[:3] if len(sys.argv) > 1 else ("B2O_TC_PAIR", "0", "1")


def make(on):
    os.environ[VAR] = ON if on else OFF
    det = Detector(weights=W.synthetic_craft_weights(3, textlike=True))
    rec = Recognizer(weights=W.synthetic_crnn_weights(2, decisive=True))
    os.environ.pop(VAR, None)
    return Pipeline(detector=det, recognizer=rec, scale=2)


def main():
    n = int(os.environ.get("PAGES", 32))
    pages, _ = synth.text_images(seed=1000, n=n, h=768, w=768, n_words=32)
    dev = torch.from_numpy(pages).cuda()
    pipes = {0: make(False), 1: make(True)}
    small = dev[:2].contiguous()
    batch0, _ = pipes[0].prepare_device(small)
    s0 = pipes[0].detector.predict_device(batch0).clone()
    s1 = pipes[1].detector.predict_device(batch0).clone()
    torch.cuda.synchronize()
    print("score maps identical:", bool(torch.equal(s0, s1)), "max|diff|", float((s0 - s1).abs().max()))
    out = {k: p.recognize(dev) for k, p in pipes.items()}
    same = [[t for t, _ in g] for g in out[0]] == [[t for t, _ in g] for g in out[1]]
    print("recognize() texts identical:", same, "words", sum(len(g) for g in out[0]), sum(len(g) for g in out[1]))
    res = {0: [], 1: []}
    for it in range(6):
        k = it & 1
        p = pipes[k]
        p.recognize(dev)
        p.detector.ctx.profile_enable(1); p.recognizer.ctx.profile_enable(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        p.recognize(dev)
        e1.record()
        torch.cuda.synchronize()
        ms_d, _, _ = p.detector.ctx.profile_read()
        ms_r, _, _ = p.recognizer.ctx.profile_read()
        p.detector.ctx.profile_enable(0); p.recognizer.ctx.profile_enable(0)
        res[k].append((e0.elapsed_time(e1), ms_d, ms_r))
    for k in (0, 1):
        print(f"{VAR}={ON if k else OFF}", " | ".join(f"step {a:.2f} conv craft {b:.2f} crnn {c:.2f}" for a, b, c in res[k]))


if __name__ == "__main__":
    main()
