"""Development check: results must not depend on how images / crops are batched."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keras_ocr_b200 import weights as W
from keras_ocr_b200.detection import Detector
from keras_ocr_b200.recognition import Recognizer
from oracle import synth

det = Detector(weights=W.synthetic_craft_weights(3, textlike=True))
rec = Recognizer(weights=W.synthetic_crnn_weights(2))
rec.keep_workspace = True
pages, _ = synth.text_images(seed=77, n=9, h=512, w=640, n_words=6)
dev = torch.from_numpy(pages).cuda()
for rep in range(2):
    s_all = det.predict_device(dev).clone()
    s_a = det.predict_device(dev[:4].contiguous()).clone()
    s_b = det.predict_device(dev[4:].contiguous()).clone()
    print("scores equal:", torch.equal(s_all[:4], s_a), torch.equal(s_all[4:], s_b),
          float((s_all[:4] - s_a).abs().max()), float((s_all[4:] - s_b).abs().max()))
    b_all, c_all = det.boxes_device(s_all)
    b_a, c_a = det.boxes_device(s_a)
    b_b, c_b = det.boxes_device(s_b)
    print("counts", c_all, c_a, c_b)
    ok = all(torch.equal(b_all[i, :c_all[i]], (b_a if i < 4 else b_b)[i if i < 4 else i - 4, :c_all[i]]) for i in range(9))
    print("boxes equal:", ok)
rng = np.random.default_rng(3)
crops = torch.from_numpy(rng.integers(0, 256, (24, 31, 200), dtype=np.uint8)).cuda()
def run(c):
    b = c.shape[0]
    x = torch.empty((b, 200, 31), dtype=torch.float16, device="cuda")
    rec.ctx.crops_to_input(c.contiguous().data_ptr(), b, x.data_ptr(), torch.cuda.current_stream().cuda_stream)
    lab = rec.predict_device(x).clone()
    taps = {}
    for name, shape, dt in (("features", (b, 50, 7, 512), torch.float16), ("theta", (b, 6), torch.float32),
                            ("warped", (b, 50, 7, 512), torch.float16), ("fc_9", (b, 50, 128), torch.float16),
                            ("l1", (b, 50, 128), torch.float16), ("l2", (b, 50, 256), torch.float16),
                            ("logits", (b, 48, 37), torch.float32)):
        taps[name] = rec.tap(name, shape, dt).clone()
    return lab, taps
for rep in range(2):
    l24, t24 = run(crops)
    for lo, hi in ((0, 10), (10, 24), (3, 4), (0, 8), (5, 18)):
        l, t = run(crops[lo:hi])
        msg = " ".join(f"{k}:{'=' if torch.equal(t24[k][lo:hi], v) else 'X%.3g' % float((t24[k][lo:hi].float() - v.float()).abs().max())}" for k, v in t.items())
        print(f"crops[{lo}:{hi}] labels {'=' if torch.equal(l24[lo:hi], l) else 'X'} {msg}")
    l24b, t24b = run(crops)
    print("repeat 24:", " ".join(f"{k}:{'=' if torch.equal(t24[k], t24b[k]) else 'X'}" for k in t24))
