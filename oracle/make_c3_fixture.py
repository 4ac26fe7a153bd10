"""tests/golden/c3_crops.npz -- the BASELINE.json configs[2] fixture (CRNN only, 256 crops 31x200 + greedy CTC).  TEST INFRASTRUCTURE.

The 256 crops are what the ORACLE pipeline cuts out of the first pages of the bench page set (oracle.synth.text_images,
seed 1000: resize/pad -> oracle CRAFT with the textlike weights -> getBoxes -> RGB2GRAY + warpBox); the expected side is
oracle/crnn.py with ``synthetic_crnn_weights(2, decisive=True)``: fp32 logits (stored as float16), greedy-CTC labels, and
the rendered word each crop shows ("" when a box does not sit on exactly one word).

    python -m oracle.make_c3_fixture            # ~3 min on 8 cores
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import crnn                                         # noqa: E402
from oracle.word_crops import labelled_crops               # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "c3_crops.npz")


def main():
    from keras_ocr_b200 import weights as W
    craft_w = W.synthetic_craft_weights(3, textlike=True)
    crnn_w = W.synthetic_crnn_weights(2, decisive=True)
    crops, labels = labelled_crops((1000, 9, 768, 768, 32), craft_w, crnn_w)      # 9 pages -> >= 256 word boxes
    assert len(crops) >= 256, len(crops)
    crops, labels = crops[:256], labels[:256]
    with torch.no_grad():
        probs, inter = crnn.crnn_logits(crnn_w, crops.astype(np.float32) / 255, return_intermediates=True)
    logits = inter["logits"].numpy()
    lab = crnn.ctc_greedy(probs)
    texts = crnn.labels_to_text(lab)
    words = np.array([w or "" for w in labels])
    top2 = torch.topk(torch.log_softmax(inter["logits"], -1), 2, -1).values
    margin = (top2[..., 0] - top2[..., 1]).numpy()
    ok = np.mean([t == w for t, w in zip(texts, words) if w])
    print(f"256 crops; oracle reads {ok:.3f} of the labelled words; steps with margin < 0.3: {(margin < 0.3).mean():.5f}, "
          f"< 1: {(margin < 1).mean():.5f}, min {margin.min():.3f}; |logits| max {np.abs(logits).max():.1f}")
    np.savez_compressed(OUT, crops=crops, logits_f16=logits.astype(np.float16), labels=lab.astype(np.int16), words=words)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
