#!/bin/bash
# compute-sanitizer over one small end-to-end recognize() (memcheck + racecheck + synccheck): every kernel of the path, tiny sizes.
set -x
O=${1:-gpurun_out/sanitizer}
mkdir -p $O
cat > /tmp/san_step.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from keras_ocr_b200 import weights as W
from keras_ocr_b200.detection import Detector
from keras_ocr_b200.pipeline import Pipeline
from keras_ocr_b200.recognition import Recognizer
from oracle import synth
pages, _ = synth.text_images(seed=21, n=2, h=192, w=384, n_words=4)
pipe = Pipeline(detector=Detector(weights=W.synthetic_craft_weights(3, textlike=True)),
                recognizer=Recognizer(weights=W.synthetic_crnn_weights(2, decisive=True)), scale=2)
out = pipe.recognize(pages)
torch.cuda.synchronize()
print("words", [[t for t, _ in g] for g in out])
PY
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san_step.py > $O/$tool.log 2>&1
  echo "$tool exit $?" >> $O/summary.txt
  tail -n 4 $O/$tool.log >> $O/summary.txt
done
cat $O/summary.txt
