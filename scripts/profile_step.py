"""One profiled recognize() step for ncu (use with --profile-from-start off).

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python scripts/profile_step.py
    PAGES=2 ncu --profile-from-start off --set full --clock-control none --import-source on \
        -k regex:conv_tc -c 6 -o gpurun_out/conv_tc python scripts/profile_step.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keras_ocr_b200 import weights as W
from keras_ocr_b200.detection import Detector
from keras_ocr_b200.pipeline import Pipeline
from keras_ocr_b200.recognition import Recognizer
from oracle import synth


def main():
    pages_n = int(os.environ.get("PAGES", 32))
    pages, _ = synth.text_images(seed=1000, n=pages_n, h=768, w=768, n_words=32)
    pipe = Pipeline(detector=Detector(weights=W.synthetic_craft_weights(3, textlike=True)),
                    recognizer=Recognizer(weights=W.synthetic_crnn_weights(2, decisive=True)), scale=2)
    dev = torch.from_numpy(pages).cuda()
    for _ in range(2):
        out = pipe.recognize(dev)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    out = pipe.recognize(dev)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("words", sum(len(g) for g in out))


if __name__ == "__main__":
    main()
