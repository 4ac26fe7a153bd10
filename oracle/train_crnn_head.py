"""Fit the recurrent head of the CRNN (fc_9, the four LSTMs, fc_12) on rendered words -- TEST INFRASTRUCTURE.

No pretrained ``crnn_kurapan.h5`` exists offline, and with purely random weights the greedy CTC argmax sits on near-ties
that fp16 may flip, which makes "identical decoded strings" (BASELINE.json north_star) impossible to assert.  This
script produces a small *decisive* checkpoint instead: the convolutional backbone + STN keep their seeded random weights
(``weights.synthetic_crnn_weights(seed)``), and everything after the spatial transformer is trained with CTC loss on
the crops that the ORACLE pipeline cuts out of ``oracle.synth.text_images`` pages (cv2 Hershey font).  The result reads
those pages, so the parity tests can assert ``texts == oracle texts`` for every crop and, beyond parity, ``texts ==
rendered words``.

    python -m oracle.train_crnn_head            # ~1.5 h on 8 cores; writes keras-ocr_b200/data/crnn_hershey_head.npz

The model trained here is the restatement in ``oracle/crnn.py`` expressed with ``torch.nn.LSTM`` (Keras gate order
i,f,c,o == torch's i,f,g,o; one bias; ``go_backwards`` outputs left in processing order, reference
recognition.py:292-319); the export is checked against ``oracle.crnn.crnn_logits`` before it is written.
"""
import os
import sys
import time

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import crnn, imageops, synth                      # noqa: E402
from oracle.pipeline import OraclePipeline                      # noqa: E402

OUT = os.path.join(ROOT, "keras-ocr_b200", "data", "crnn_hershey_head.npz")
CACHE = os.environ.get("B2O_TRAIN_CACHE", "/tmp/b2o_train_cache")
CRAFT_SEED, CRNN_SEED = 3, 2                                    # the seeds tests/ and bench.py use
JITTER_COPIES = int(os.environ.get("B2O_TRAIN_JITTER", 4))      # perturbed-box copies of every training crop

# (seed, n, h, w, n_words): every page set the GPU tests / smoke / bench render, plus extra pages for generalisation
PAGE_SETS = [
    (1000, 32, 768, 768, 32),        # bench.py C4 pages; tests use the first pages of the same stream
    (21, 2, 192, 384, 4), (77, 9, 256, 320, 6), (5, 3, 192, 384, 4), (4, 2, 96, 128, 3),
    (2000, 12, 768, 768, 32), (2001, 24, 384, 384, 10), (2002, 24, 256, 512, 8),
]
HOLDOUT = (3000, 4, 768, 768, 32)


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


def backbone_sequences(crnn_w, crops, batch=64):
    """Frozen part: conv stack + STN -> (N, 50, 3584) fp32, exactly what fc_9 consumes in oracle.crnn.crnn_logits."""
    w = crnn._t(crnn_w)
    out = []
    with torch.no_grad():
        for i in range(0, len(crops), batch):
            x = crops[i:i + batch].astype(np.float32) / 255
            feat = crnn.crnn_features(crnn_w, x)
            warped = crnn.stn_sample(feat.permute(0, 2, 3, 1).contiguous(), crnn.stn_theta(w, feat))
            out.append(warped.reshape(warped.shape[0], warped.shape[1], -1).clone())
    return torch.cat(out)


class Head(nn.Module):
    def __init__(self, classes=37):
        super().__init__()
        self.fc9 = nn.Linear(3584, 128)
        self.l10, self.l10b = nn.LSTM(128, 128, batch_first=True), nn.LSTM(128, 128, batch_first=True)
        self.l11, self.l11b = nn.LSTM(128, 128, batch_first=True), nn.LSTM(128, 128, batch_first=True)
        self.fc12 = nn.Linear(256, classes)

    def forward(self, seq):
        x = F.relu(self.fc9(seq))
        l1 = self.l10(x)[0] + self.l10b(torch.flip(x, [1]))[0]                    # Add, no re-reversal (recognition.py:305)
        l2 = torch.cat([self.l11(l1)[0], self.l11b(torch.flip(l1, [1]))[0]], -1)  # Concatenate (319)
        return self.fc12(l2)[:, crnn.STEPS_TO_DISCARD:]

    def export(self):
        out = {"fc_9.kernel": self.fc9.weight.T, "fc_9.bias": self.fc9.bias,
               "fc_12.kernel": self.fc12.weight.T, "fc_12.bias": self.fc12.bias}
        for name, m in (("lstm_10", self.l10), ("lstm_10_back", self.l10b), ("lstm_11", self.l11), ("lstm_11_back", self.l11b)):
            out[name + ".kernel"] = m.weight_ih_l0.T
            out[name + ".recurrent_kernel"] = m.weight_hh_l0.T
            out[name + ".bias"] = m.bias_ih_l0 + m.bias_hh_l0
        # stored as fp16 (what the device holds anyway); oracle and device both load these exact values
        return {k: v.detach().numpy().astype(np.float16) for k, v in out.items()}


def encode(words):
    flat = torch.tensor([crnn.ALPHABET.index(c) for wd in words for c in wd], dtype=torch.long)
    return flat, torch.tensor([len(wd) for wd in words], dtype=torch.long)


def greedy_texts(logits):
    return crnn.labels_to_text(crnn.ctc_greedy(torch.softmax(logits, -1)))


def margins(logits):
    top2 = torch.topk(torch.log_softmax(logits, -1), 2, -1).values
    return (top2[..., 0] - top2[..., 1])


def main():
    from keras_ocr_b200 import weights as W
    torch.manual_seed(0)
    craft_w = W.synthetic_craft_weights(CRAFT_SEED, textlike=True)
    crnn_w = W.synthetic_crnn_weights(CRNN_SEED)
    os.makedirs(CACHE, exist_ok=True)
    rng = np.random.default_rng(0)
    t0 = time.time()

    def cached(tag, page_set, jitter):
        path = os.path.join(CACHE, f"{tag}_j{jitter}.npz")
        if os.path.exists(path):
            d = np.load(path, allow_pickle=True)
            return torch.from_numpy(d["seq"]), list(d["labels"])
        crops, labels = labelled_crops(page_set, craft_w, crnn_w, jitter, rng)
        seq = backbone_sequences(crnn_w, crops)
        np.savez(path, seq=seq.numpy(), labels=np.array(labels, dtype=object), crops=crops)
        print(f"[{time.time() - t0:6.0f}s] {tag}: {len(crops)} crops, {sum(l is None for l in labels)} unlabelled", flush=True)
        return seq, labels

    seqs, labels = [], []
    for ps in PAGE_SETS:
        s, l = cached("set_%d" % ps[0], ps, jitter=JITTER_COPIES)
        seqs.append(s); labels += l
    seq = torch.cat(seqs)
    keep = [i for i, l in enumerate(labels) if l is not None]
    x_train, y_train = seq[keep], [labels[i] for i in keep]
    x_hold, y_hold = cached("hold_%d" % HOLDOUT[0], HOLDOUT, jitter=0)
    print(f"train {len(y_train)} crops, holdout {len(y_hold)}", flush=True)

    head = Head()
    if os.path.exists(OUT) and os.environ.get("B2O_TRAIN_WARM", "1") != "0":      # warm start from the previous fit
        prev = np.load(OUT)
        with torch.no_grad():
            head.fc9.weight.copy_(torch.from_numpy(prev["fc_9.kernel"].astype(np.float32).T)); head.fc9.bias.copy_(torch.from_numpy(prev["fc_9.bias"].astype(np.float32)))
            head.fc12.weight.copy_(torch.from_numpy(prev["fc_12.kernel"].astype(np.float32).T)); head.fc12.bias.copy_(torch.from_numpy(prev["fc_12.bias"].astype(np.float32)))
            for name, m in (("lstm_10", head.l10), ("lstm_10_back", head.l10b), ("lstm_11", head.l11), ("lstm_11_back", head.l11b)):
                m.weight_ih_l0.copy_(torch.from_numpy(prev[name + ".kernel"].astype(np.float32).T))
                m.weight_hh_l0.copy_(torch.from_numpy(prev[name + ".recurrent_kernel"].astype(np.float32).T))
                m.bias_ih_l0.copy_(torch.from_numpy(prev[name + ".bias"].astype(np.float32))); m.bias_hh_l0.zero_()
        print("warm start from", OUT, flush=True)
    epochs = int(os.environ.get("B2O_TRAIN_EPOCHS", 120))
    opt = torch.optim.Adam(head.parameters(), lr=2e-3, weight_decay=1e-5)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=float(os.environ.get("B2O_TRAIN_LR", 3e-3)), total_steps=epochs * ((len(y_train) + 63) // 64))
    scale = float(x_train.std())
    for ep in range(epochs):
        perm = torch.randperm(len(y_train))
        total = 0.0
        for i in range(0, len(perm), 64):
            idx = perm[i:i + 64]
            xb = x_train[idx]
            xb = xb + torch.randn_like(xb) * (0.02 * scale)          # robustness to the fp16 chain's small deviations
            tgt, tl = encode([y_train[j] for j in idx])
            lp = F.log_softmax(head(xb), -1).permute(1, 0, 2)
            loss = F.ctc_loss(lp, tgt, torch.full((len(idx),), lp.shape[0], dtype=torch.long), tl, blank=crnn.BLANK, zero_infinity=True)
            opt.zero_grad(); loss.backward()
            nn.utils.clip_grad_norm_(head.parameters(), 5.0)
            opt.step(); sched.step()
            total += float(loss) * len(idx)
        if ep % 5 == 4 or ep == epochs - 1:
            with torch.no_grad():
                acc = np.mean([a == b for a, b in zip(greedy_texts(head(x_train[:2000])), y_train[:2000])])
                hk = [i for i, l in enumerate(y_hold) if l is not None]
                hacc = np.mean([a == b for a, b in zip(greedy_texts(head(x_hold[hk])), [y_hold[i] for i in hk])])
                m = margins(head(x_train[:2000]))
            print(f"[{time.time() - t0:6.0f}s] epoch {ep + 1}: loss {total / len(perm):.4f} train acc {acc:.4f} holdout acc {hacc:.4f} "
                  f"steps with margin<1: {float((m < 1).float().mean()):.5f}", flush=True)

    exported = head.export()
    full = dict(crnn_w); full.update({k: v.astype(np.float32) for k, v in exported.items()})
    with torch.no_grad():                                        # the export really is the oracle's network
        d = np.load(os.path.join(CACHE, "set_%d_j%d.npz" % (PAGE_SETS[1][0], JITTER_COPIES)), allow_pickle=True)
        probs = crnn.crnn_logits(full, d["crops"].astype(np.float32) / 255)
        print("oracle texts on set 21:", crnn.labels_to_text(crnn.ctc_greedy(probs)), "labels:", list(d["labels"]))
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez(OUT, **exported)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
