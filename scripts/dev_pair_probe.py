"""Development probe for the CTA-pair (cta_group::2) convolution path: one small convolution per kernel mode through
b2o_conv2d_test (pairs are the default; B2O_TC_PAIR=0 disables them), compared with torch; on a mismatch prints WHERE it is wrong (which CTA of the
pair = tile-column parity, which half of the output channels = which CTA's half of B).

    python scripts/dev_pair_probe.py                 # halo-tile pairs (default)
    B2O_TC_PAIR=2 python scripts/dev_pair_probe.py   # + generic-tile pairs (opt-in, to be validated)
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keras_ocr_b200 import _lib

CASES = [  # n, h, w, cin, cout, ksize, dilation   (3x3 dilation 1 -> halo tiles; anything else -> generic tiles)
    (1, 16, 16, 64, 64, 3, 1),      # one pair, resident bank (MODE 3)
    (1, 16, 32, 64, 128, 3, 1),     # two pairs
    (2, 50, 7, 512, 512, 3, 1),     # one tile column: the peer CTA works on a dummy tile; streamed B, 2 n-tiles (MODE 1)
    (1, 48, 40, 128, 128, 3, 1),    # odd number of tile columns (5)
    (1, 96, 96, 64, 64, 3, 1),      # many tiles per CTA pair
    (1, 32, 24, 256, 512, 3, 1),    # two n-tiles of 256
    # generic tiles: paired only with B2O_TC_PAIR=2 (not validated yet)
    (1, 8, 8, 1536, 512, 1, 1),     # upconv1.conv.0
    (1, 9, 13, 512, 1024, 3, 6),    # dilated slice5.1, odd spatial size
    (1, 1, 300, 3584, 128, 1, 1),   # fc_9 as a 1x1 conv over rows
    (2, 24, 24, 1024, 1024, 1, 1),  # slice5.2
]


def main():
    ctx = _lib.Context(0)
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(0)
    for n, h, w, cin, cout, k, dil in CASES:
        x = torch.from_numpy(rng.standard_normal((n, h, w, cin)).astype(np.float32)).to(dev).half()
        wgt = (rng.standard_normal((cout, k, k, cin)) / np.sqrt(k * k * cin)).astype(np.float32)
        s1, t1 = np.ones(cout, np.float32), np.zeros(cout, np.float32)
        out = torch.zeros((n, h, w, cout), dtype=torch.float16, device=dev)
        try:
            ctx.conv2d_test(x.data_ptr(), n, h, w, cin, wgt, cout, k, dil, s1, t1, 0, None, None, out.data_ptr(), _lib.CONV_AUTO, st)
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print(f"case {n}x{h}x{w} {cin}->{cout} k{k}d{dil}: FAILED {e}")
            break
        wt = torch.from_numpy(wgt).to(dev).half().float().permute(0, 3, 1, 2)
        ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), wt, padding=dil * (k // 2), dilation=dil).permute(0, 2, 3, 1)
        err = (out.float() - ref).abs()
        scale = float(ref.abs().max())
        line = f"case {n}x{h}x{w} {cin}->{cout} k{k}d{dil}: max err {float(err.max()):.3e} (scale {scale:.2f})"
        if float(err.max()) > 5e-3 * scale:
            col = (torch.arange(w, device=dev) // 8) % 2               # tile-column parity = CTA rank
            for r in (0, 1):
                for half in (0, 1):
                    e = err[:, :, col == r][..., half * cout // 2:(half + 1) * cout // 2]
                    line += f"\n    rank {r} channels half {half}: max err {float(e.max()) if e.numel() else 0:.3e}"
        print(line)


if __name__ == "__main__":
    main()
