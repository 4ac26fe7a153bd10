"""Development probe: wall-clock of each phase of Pipeline.recognize with explicit synchronisation."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keras_ocr_b200 import weights as W, recognition, tools
from keras_ocr_b200.detection import Detector
from keras_ocr_b200.pipeline import Pipeline
from keras_ocr_b200.recognition import Recognizer
from oracle import synth

pages, _ = synth.text_images(seed=1000, n=32, h=768, w=768, n_words=32)
pipe = Pipeline(detector=Detector(weights=W.synthetic_craft_weights(3, textlike=True)),
                recognizer=Recognizer(weights=W.synthetic_crnn_weights(2)), scale=2)
det, rec = pipe.detector, pipe.recognizer
dev = torch.from_numpy(pages).cuda()
for _ in range(3):
    pipe.recognize(dev)
torch.cuda.synchronize()
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for rep in range(3):
    t = [T()]
    batch, scales = pipe.prepare_device(dev); t.append(T())
    scores = det.predict_device(batch); t.append(T())
    boxes, counts = det.boxes_device(scores); t.append(T())
    labels = rec.recognize_from_boxes_device(batch, boxes, counts); t.append(T())
    boxes_host = boxes.cpu().numpy(); labels_host = labels.cpu().numpy(); t.append(T())
    texts = recognition.labels_to_text(labels_host, rec.alphabet); t.append(T())
    out, start = [], 0
    for i, (c, scale) in enumerate(zip(counts, scales)):
        c = int(c); group = boxes_host[i, :c]
        group = tools.adjust_boxes(boxes=group, boxes_format="boxes", scale=1 / scale)
        out.append(list(zip(texts[start:start + c], group))); start += c
    t.append(T())
    names = ["prepare", "craft", "get_boxes", "warp+crnn", "d2h", "strings", "assemble"]
    print(" ".join(f"{n}={1e3*(b-a):.2f}ms" for n, a, b in zip(names, t, t[1:])), f"total={1e3*(t[-1]-t[0]):.2f}ms")
for inflight in (1, 2, 1, 2):
    pipe.inflight = inflight
    pipe.recognize(dev)
    t0 = T()
    for _ in range(5):
        pipe.recognize(dev)
    t1 = T()
    pipe.recognize(pages)
    t2 = T()
    for _ in range(5):
        pipe.recognize(pages)
    t3 = T()
    print(f"inflight={inflight}: recognize(device) {(t1-t0)/5*1e3:.2f} ms/step, recognize(numpy) {(t3-t2)/5*1e3:.2f} ms/step")
