"""Development A/B: conv_tc kernel time per recognize() step under two settings of an environment switch that
conv_tc_run reads at every launch, alternated in ONE process on ONE box (boxes differ by +-5 %).

    python scripts/dev_ab.py B2O_TC_NO_COALESCE
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


def main():
    var = sys.argv[1]
    pages, _ = synth.text_images(seed=1000, n=int(os.environ.get("PAGES", 32)), h=768, w=768, n_words=32)
    det = Detector(weights=W.synthetic_craft_weights(3, textlike=True))
    rec = Recognizer(weights=W.synthetic_crnn_weights(2))
    pipe = Pipeline(detector=det, recognizer=rec, scale=2)
    dev = torch.from_numpy(pages).cuda()
    for _ in range(2):
        pipe.recognize(dev)
    res = {0: [], 1: []}
    for it in range(6):
        on = it & 1
        if on:
            os.environ[var] = "1"
        else:
            os.environ.pop(var, None)
        det.ctx.profile_enable(1); rec.ctx.profile_enable(1)
        pipe.recognize(dev)
        torch.cuda.synchronize()
        ms_d, _, _ = det.ctx.profile_read()
        ms_r, _, _ = rec.ctx.profile_read()
        det.ctx.profile_enable(0); rec.ctx.profile_enable(0)
        res[on].append((ms_d, ms_r))
    for on in (0, 1):
        print(f"{var}={'1' if on else 'unset'}:", " ".join(f"craft {a:.2f} crnn {b:.2f} |" for a, b in res[on]))


if __name__ == "__main__":
    main()
