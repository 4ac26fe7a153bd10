"""Development probe for the CTA-pair (cta_group::2) convolution path: one small convolution per kernel mode through
b2o_conv2d_test (pairs are the default; B2O_TC_PAIR=0 disables them), compared with torch; on a mismatch prints WHERE it is wrong (which CTA of the
pair = tile-column parity, which half of the output channels = which CTA's half of B).

    B2O_TC_PAIR=1 python scripts/dev_pair_probe.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("B2O_TC_PAIR", "1")
from keras_ocr_b200 import _lib

CASES = [  # n, h, w, cin, cout   (3x3, dilation 1 -> halo tiles)
    (1, 16, 16, 64, 64),      # one pair, resident bank (MODE 3)
    (1, 16, 32, 64, 128),     # two pairs
    (2, 50, 7, 512, 512),     # one tile column: the peer CTA works on a dummy tile; streamed B, 2 n-tiles (MODE 1)
    (1, 48, 40, 128, 128),    # odd number of tile columns (5)
    (1, 96, 96, 64, 64),      # many tiles per CTA pair
    (1, 32, 24, 256, 512),    # two n-tiles of 256
]


def main():
    ctx = _lib.Context(0)
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(0)
    for n, h, w, cin, cout in CASES:
        x = torch.from_numpy(rng.standard_normal((n, h, w, cin)).astype(np.float32)).to(dev).half()
        wgt = (rng.standard_normal((cout, 3, 3, cin)) / np.sqrt(9 * cin)).astype(np.float32)
        s1, t1 = np.ones(cout, np.float32), np.zeros(cout, np.float32)
        out = torch.zeros((n, h, w, cout), dtype=torch.float16, device=dev)
        try:
            ctx.conv2d_test(x.data_ptr(), n, h, w, cin, wgt, cout, 3, 1, s1, t1, 0, None, None, out.data_ptr(), _lib.CONV_AUTO, st)
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print(f"case {n}x{h}x{w} {cin}->{cout}: FAILED {e}")
            break
        wt = torch.from_numpy(wgt).to(dev).half().float().permute(0, 3, 1, 2)
        ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), wt, padding=1).permute(0, 2, 3, 1)
        err = (out.float() - ref).abs()
        scale = float(ref.abs().max())
        line = f"case {n}x{h}x{w} {cin}->{cout}: max err {float(err.max()):.3e} (scale {scale:.2f})"
        if float(err.max()) > 5e-3 * scale:
            col = (torch.arange(w, device=dev) // 8) % 2               # tile-column parity = CTA rank
            for r in (0, 1):
                for half in (0, 1):
                    e = err[:, :, col == r][..., half * cout // 2:(half + 1) * cout // 2]
                    line += f"\n    rank {r} channels half {half}: max err {float(e.max()) if e.numel() else 0:.3e}"
        print(line)


if __name__ == "__main__":
    main()
