"""Word crops of rendered pages, cut by the ORACLE pipeline and labelled with the rendered word -- TEST INFRASTRUCTURE.

Used by ``oracle/make_c3_fixture.py`` (the 256 crops of BASELINE configs[2]) and, packed into
``oracle/_train_data/real_crops.npz``, by ``oracle/train_crnn_full.py`` (the recognizer the parity tests and the bench use).

    python -m oracle.word_crops        # ~12 min on 8 cores (oracle CRAFT on ~110 pages): writes oracle/_train_data/real_crops.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import imageops, synth                             # noqa: E402
from oracle.pipeline import OraclePipeline                      # noqa: E402

OUT = os.path.join(ROOT, "oracle", "_train_data", "real_crops.npz")
CRAFT_SEED = 3                                                  # the textlike CRAFT weights tests/ and bench.py use
JITTER_COPIES = 4                                               # perturbed-box copies of every crop

# (seed, n, h, w, n_words): every page set the GPU tests / smoke / bench render, plus extra pages
PAGE_SETS = [
    (1000, 32, 768, 768, 32),        # bench.py C4 pages; tests use the first pages of the same stream
    (21, 2, 192, 384, 4), (77, 9, 256, 320, 6), (5, 3, 192, 384, 4), (4, 2, 96, 128, 3),
    (2000, 12, 768, 768, 32), (2001, 24, 384, 384, 10), (2002, 24, 256, 512, 8),
]
HOLDOUT = (3000, 4, 768, 768, 32)                               # pages no training crop comes from


JITTER_PX = 3.0          # detector-input pixels: the fp16 CUDA chain's boxes differ from the oracle's by up to ~3 such pixels at 1536^2


def detect_pages(page_set, craft_w, crnn_w):
    """Oracle chain up to the boxes: (padded batch, scales, box groups, rendered words, glyph rectangles)."""
    seed, n, h, w, n_words = page_set
    r = np.random.default_rng(seed)
    pages, words, rects = [], [], []
    for _ in range(n):
        img, ws, rc = synth.text_image(r, h, w, n_words, return_layout=True)
        pages.append(img); words.append(ws); rects.append(rc)
    pipe = OraclePipeline(craft_w, crnn_w, scale=2)
    batch, scales = pipe.prepare(np.stack(pages))
    groups = []
    for i in range(0, n, 4):                                    # bounded memory: 4 pages of fp32 CRAFT at a time
        groups += imageops.get_boxes(pipe.detect_scores(batch[i:i + 4]))
    return batch, scales, groups, words, rects


def labelled_crops(page_set, craft_w, crnn_w, jitter=0, rng=None, detected=None):
    """Oracle chain up to the crops, each crop paired with the rendered word whose glyph rectangle holds the box
    centre (None when a box does not sit on exactly one word, e.g. a split word).  ``jitter`` extra copies of every
    crop are cut from the box with its corners moved by up to JITTER_PX: the recognizer must give the same string for
    the slightly different boxes the fp16 chain finds."""
    batch, scales, groups, words, rects = detected if detected is not None else detect_pages(page_set, craft_w, crnn_w)
    crops, labels = [], []
    for img, boxes, ws, rc, s in zip(batch, groups, words, rects, scales):
        gray = imageops.rgb_to_gray(img)
        hits = {}
        for bi, box in enumerate(boxes):
            c = np.asarray(box).mean(0) / s
            inside = [k for k, (x0, y0, x1, y1) in enumerate(rc) if x0 <= c[0] <= x1 and y0 <= c[1] <= y1]
            hits[bi] = inside[0] if len(inside) == 1 else None
        counts = {}
        for k in hits.values():
            counts[k] = counts.get(k, 0) + 1
        for bi, box in enumerate(boxes):
            k = hits[bi]
            word = ws[k] if k is not None and counts[k] == 1 else None
            variants = [np.asarray(box, np.float32)]
            for _ in range(jitter):
                variants.append(variants[0] + rng.uniform(-JITTER_PX, JITTER_PX, (4, 2)).astype(np.float32))
            for v in variants:
                crops.append(imageops.warp_box(gray, v))
                labels.append(word)
    return np.array(crops), labels


def main():
    from keras_ocr_b200 import weights as W
    craft_w = W.synthetic_craft_weights(CRAFT_SEED, textlike=True)
    rng = np.random.default_rng(0)
    crops, labels = [], []
    for ps in PAGE_SETS:
        c, l = labelled_crops(ps, craft_w, None, JITTER_COPIES, rng)
        keep = [i for i, x in enumerate(l) if x is not None]
        crops.append(c[keep]); labels += [l[i] for i in keep]
        print(f"pages seed {ps[0]}: {len(c)} crops, {len(keep)} labelled", flush=True)
    hc, hl = labelled_crops(HOLDOUT, craft_w, None, 0, rng)
    hk = [i for i, x in enumerate(hl) if x is not None]
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, crops=np.concatenate(crops), labels=np.array(labels), hold_crops=hc[hk], hold_labels=np.array([hl[i] for i in hk]))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
