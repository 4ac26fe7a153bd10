"""Development probe: MMA-warp wait breakdown of single conv layers (reads g_tc_debug)."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keras_ocr_b200 import _lib
ctx = _lib.Context(0)
lib = ctx.lib
lib.b2o_debug_read_tc.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
cases = [("crnn conv_6 512->512", 256, 50, 7, 512, 512, 3), ("crnn conv_4 256->256", 256, 100, 15, 256, 256, 3),
         ("slice5.2 1x1 1024", 32, 96, 96, 1024, 1024, 1),
         ("stem 16->64 3x3", 8, 1536, 1536, 16, 64, 3), ("conv2 64->64 3x3", 8, 1536, 1536, 64, 64, 3), ("cls 32->32 3x3", 8, 768, 768, 32, 32, 3),
         ("conv3 64->128", 8, 768, 768, 64, 128, 3), ("conv4 128->128", 8, 768, 768, 128, 128, 3),
         ("256->256", 8, 384, 384, 256, 256, 3), ("1x1 64->64", 8, 1536, 1536, 64, 64, 1)]
rng = np.random.default_rng(0)
for name, n, h, w, cin, cout, k in cases:
    x = torch.randn((n, h, w, cin), device="cuda").half()
    wgt = (rng.standard_normal((cout, k, k, cin)) * 0.05).astype(np.float32)
    s1 = np.ones(cout, np.float32); t1 = np.zeros(cout, np.float32)
    out = torch.empty((n, h, w, cout), dtype=torch.float16, device="cuda")
    for _ in range(2):
        ctx.conv2d_test(x.data_ptr(), n, h, w, cin, wgt, cout, k, 1, s1, t1, 1, None, None, out.data_ptr(), _lib.CONV_AUTO,
                        torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * (148 * 8))()
    assert lib.b2o_debug_read_tc(buf, 148 * 8) == 0
    a = np.array(buf, dtype=np.float64).reshape(148, 8)
    tiles = a[:, 4].mean()
    print(f"{name:20s} tiles/CTA {tiles:7.1f}  cycles/tile {a[:,0].mean()/tiles:8.1f}  wait tmem_empty {a[:,1].mean()/tiles:7.1f}  "
          f"wait a_full {a[:,2].mean()/tiles:7.1f}  wait b_full {a[:,3].mean()/tiles:7.1f}  issue+other {(a[:,0]-a[:,1]-a[:,2]-a[:,3]).mean()/tiles:7.1f}")
