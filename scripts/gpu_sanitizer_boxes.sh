#!/bin/bash
# compute-sanitizer (memcheck, racecheck) over getBoxes alone: ordinary words, and the components that take the queue /
# large-plane / global-scratch paths of the quads kernels.
set -x
O=${1:-gpurun_out/sanitizer_boxes}
mkdir -p $O
cat > /tmp/san_boxes.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from keras_ocr_b200 import weights as W
from keras_ocr_b200.detection import Detector
from oracle import imageops, synth
det = Detector(weights=W.synthetic_craft_weights(3, textlike=True))
big = np.zeros((1, 1000, 1000, 2), np.float32)
big[0, 10:340, 50:950, 0] = 0.9
big[0, 420:750, 40:940, 0] = 0.85
big[0, 500:600, 300:500, 1] = 0.9
big[0, 830:990, 100:400, 0] = 0.9
big[0, 900:910, 500:560, 0] = 0.8
for scores in (synth.score_maps(101, 2, 384, 384, 16), big):
    boxes, counts = det.boxes_device(torch.from_numpy(np.ascontiguousarray(scores)).cuda())
    torch.cuda.synchronize()
    ref = imageops.get_boxes(scores)
    assert list(counts) == [len(r) for r in ref], (counts, [len(r) for r in ref])
    print("boxes", list(counts))
PY
for tool in memcheck racecheck; do
  timeout 100 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san_boxes.py > $O/$tool.log 2>&1
  echo "$tool exit $?" >> $O/summary.txt
  grep -E "boxes \[|ERROR SUMMARY|RACECHECK SUMMARY|hazard" $O/$tool.log | tail -n 6 >> $O/summary.txt
done
cat $O/summary.txt
