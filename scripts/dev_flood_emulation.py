"""Lane-by-lane emulation of the row-sweep flood fill of csrc/boxes.cu (quad_of_component: one warp, lane = word of the row,
alternating downward / upward sweeps until a sweep changes nothing) against scipy's 8-connected labelling, on random masks
at several densities, multi-word and >32-word rows, and serpentine shapes that need many sweeps.  Development aid: the CUDA
code is checked on the GPU by the getBoxes tests; this checks the ALGORITHM (closure = the seed's 8-connected blob).

    python scripts/dev_flood_emulation.py
"""
import numpy as np
from scipy import ndimage
M32 = 0xffffffff
def pack(mask):
    rh, rw = mask.shape; stride = (rw + 31)//32
    P = np.zeros((rh, stride), dtype=np.uint64)
    for r in range(rh):
        for x in range(rw):
            if mask[r, x]: P[r, x//32] |= np.uint64(1 << (x % 32))
    return P, stride
def unpack(P, rw):
    rh, stride = P.shape
    out = np.zeros((rh, rw), bool)
    for r in range(rh):
        for x in range(rw):
            out[r, x] = (int(P[r, x//32]) >> (x % 32)) & 1
    return out
def flood(B, stride, rh, seed_rk, seed_bit):
    A = [[0]*stride for _ in range(rh)]
    A[seed_rk[0]][seed_rk[1]] = seed_bit
    down = True; sweeps = 0
    while True:
        changed = False; sweeps += 1
        for rr in range(rh):
            r = rr if down else rh-1-rr
            for k0 in range(0, stride, 32):
                writes = []
                for lane in range(32):
                    k = k0 + lane
                    if k >= stride: continue
                    d = int(B[r][k])
                    if not d: continue
                    nb = 0; cur = 0
                    for dr in (-1, 0, 1):
                        r2 = r + dr
                        if r2 < 0 or r2 >= rh: continue
                        row = A[r2]
                        mid = row[k]; prev = row[k-1] if k > 0 else 0; nxt = row[k+1] if k+1 < stride else 0
                        if dr == 0: cur = mid
                        nb |= mid | ((mid << 1) & M32) | (mid >> 1) | (prev >> 31) | ((nxt << 31) & M32)
                    grown = nb & d
                    while True:
                        pg = grown
                        grown |= (((grown << 1) & M32) | (grown >> 1)) & d
                        if grown == pg: break
                    grown &= ~cur & M32
                    if grown: writes.append((k, cur | grown))
                for k, v in writes:
                    A[r][k] = v; changed = True
        if not changed: break
        down = not down
    return np.array(A, dtype=np.uint64), sweeps
rng = np.random.default_rng(0)
worst = 0
for trial in range(300):
    rh = int(rng.integers(1, 40)); rw = int(rng.integers(1, 200 if trial % 10 else 1200))
    dens = rng.choice([0.3, 0.5, 0.62, 0.8])
    mask = rng.random((rh, rw)) < dens
    if trial % 7 == 0:   # snake / spiral-like structures
        mask = np.zeros((rh, rw), bool); mask[::2, :] = True
        for i, r in enumerate(range(1, rh, 2)): mask[r, (rw-1) if i % 2 == 0 else 0] = True
    if not mask.any(): continue
    P, stride = pack(mask)
    # seed = first raster pixel
    ys, xs = np.nonzero(mask); y0, x0 = ys[0], xs[0]
    A, sweeps = flood(P.tolist(), stride, rh, (y0, x0//32), 1 << (x0 % 32))
    lab, n = ndimage.label(mask, structure=np.ones((3, 3)))
    ref = lab == lab[y0, x0]
    got = unpack(A, rw)
    assert (got == ref).all(), (trial, rh, rw)
    worst = max(worst, sweeps)
print('ok, worst sweeps', worst)
