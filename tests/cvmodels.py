"""NumPy models of the integer / fixed-point arithmetic the CUDA kernels in
keras-ocr_b200/csrc/{image,boxes}.cu implement.  They restate what OpenCV 4.x does inside
cv2.resize, cv2.getPerspectiveTransform, cv2.warpPerspective, cv2.convexHull, cv2.minAreaRect and
cv2.boxPoints, and are pinned against cv2 itself in tests/test_cv_models.py (CPU only), so a
kernel that follows the model is bit-compatible with the reference's OpenCV calls
(reference tools.py:96-107,394-396; detection.py:267-273)."""
import numpy as np

f32 = np.float32

def cv_round_short(x):
    # saturate_cast<short>(float) = cvRound (round half to even)
    return np.clip(np.rint(x), -32768, 32767).astype(np.int32)
def resize_model(src, dw, dh):
    sh, sw = src.shape[:2]
    inv_x = dw / sw; scale_x = 1.0 / inv_x
    inv_y = dh / sh; scale_y = 1.0 / inv_y
    dx = np.arange(dw)
    fx = ((dx + 0.5) * scale_x - 0.5).astype(np.float32)
    sx = np.floor(fx).astype(np.int32)
    fx = (fx - sx).astype(np.float32)
    lo = sx < 0
    fx[lo] = 0; sx[lo] = 0
    hi = sx >= sw - 1
    fx[hi] = 0; sx[hi] = sw - 1
    a0 = cv_round_short((np.float32(1.0) - fx) * np.float32(2048))
    a1 = cv_round_short(fx * np.float32(2048))
    sx1 = np.minimum(sx + 1, sw - 1)
    dy = np.arange(dh)
    fy = ((dy + 0.5) * scale_y - 0.5).astype(np.float32)
    sy = np.floor(fy).astype(np.int32)
    fy = (fy - sy).astype(np.float32)
    b0 = cv_round_short((np.float32(1.0) - fy) * np.float32(2048))
    b1 = cv_round_short(fy * np.float32(2048))
    y0 = np.clip(sy, 0, sh - 1); y1 = np.clip(sy + 1, 0, sh - 1)
    s = src.astype(np.int32)
    # horizontal pass on needed rows
    H = s[:, sx] * a0[None, :, None] + s[:, sx1] * a1[None, :, None]   # (sh, dw, c)
    S0 = H[y0]; S1 = H[y1]
    out = (((b0[:, None, None] * (S0 >> 4)) >> 16) + ((b1[:, None, None] * (S1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)

def lu_solve8(A, b):
    A = A.copy(); b = b.copy(); m = 8
    for i in range(m):
        k = i
        for j in range(i+1, m):
            if abs(A[j,i]) > abs(A[k,i]): k = j
        if k != i:
            A[[i,k], i:] = A[[k,i], i:]
            b[[i,k]] = b[[k,i]]
        d = -1.0 / A[i,i]
        for j in range(i+1, m):
            alpha = A[j,i]*d
            for kk in range(i+1, m):
                A[j,kk] += alpha*A[i,kk]
            b[j] += alpha*b[i]
    for i in range(m-1, -1, -1):
        s = b[i]
        for kk in range(i+1, m):
            s -= A[i,kk]*b[kk]
        b[i] = s / A[i,i]
    return b

def get_persp(src, dst):
    A = np.zeros((8,8)); b = np.zeros(8)
    for i in range(4):
        sx, sy = float(src[i,0]), float(src[i,1]); dx, dy = float(dst[i,0]), float(dst[i,1])
        A[i,0]=A[i+4,3]=sx; A[i,1]=A[i+4,4]=sy; A[i,2]=A[i+4,5]=1
        A[i,6] = -sx*dx; A[i,7] = -sy*dx; A[i+4,6] = -sx*dy; A[i+4,7] = -sy*dy
        b[i]=dx; b[i+4]=dy
    x = lu_solve8(A,b)
    return np.append(x, 1.0).reshape(3,3)

def inv3(S):
    m = S
    det = m[0,0]*(m[1,1]*m[2,2]-m[1,2]*m[2,1]) - m[0,1]*(m[1,0]*m[2,2]-m[1,2]*m[2,0]) + m[0,2]*(m[1,0]*m[2,1]-m[1,1]*m[2,0])
    d = 1.0/det
    t = np.zeros(9)
    t[0]=(m[1,1]*m[2,2]-m[1,2]*m[2,1])*d; t[1]=(m[0,2]*m[2,1]-m[0,1]*m[2,2])*d; t[2]=(m[0,1]*m[1,2]-m[0,2]*m[1,1])*d
    t[3]=(m[1,2]*m[2,0]-m[1,0]*m[2,2])*d; t[4]=(m[0,0]*m[2,2]-m[0,2]*m[2,0])*d; t[5]=(m[0,2]*m[1,0]-m[0,0]*m[1,2])*d
    t[6]=(m[1,0]*m[2,1]-m[1,1]*m[2,0])*d; t[7]=(m[0,1]*m[2,0]-m[0,0]*m[2,1])*d; t[8]=(m[0,0]*m[1,1]-m[0,1]*m[1,0])*d
    return t.reshape(3,3)

def warp_model(gray, M, dw, dh, blocked=True):
    Mi = inv3(M).reshape(-1)
    H, W = gray.shape
    out = np.zeros((dh, dw), np.uint8)
    g = gray.astype(np.int64)
    bh0 = min(16, dh); bw0 = min(1024//bh0, dw); bh0 = min(1024//bw0, dh)
    for y in range(dh):
        for x in range(dw):
            if blocked:
                bx = (x // bw0) * bw0; x1 = x - bx
                X0 = Mi[0]*bx + Mi[1]*y + Mi[2]; Y0 = Mi[3]*bx + Mi[4]*y + Mi[5]; W0 = Mi[6]*bx + Mi[7]*y + Mi[8]
                Wv = W0 + Mi[6]*x1
                Wv = 32.0/Wv if Wv != 0 else 0.0
                fX = max(-2147483648.0, min(2147483647.0, (X0 + Mi[0]*x1)*Wv))
                fY = max(-2147483648.0, min(2147483647.0, (Y0 + Mi[3]*x1)*Wv))
            X = int(np.rint(fX)); Y = int(np.rint(fY))
            sx = X >> 5; sy = Y >> 5; ax = X & 31; ay = Y & 31
            sx = max(-32768, min(32767, sx)); sy = max(-32768, min(32767, sy))
            def px(yy, xx):
                return g[yy, xx] if (0 <= yy < H and 0 <= xx < W) else 0
            w00 = (32-ax)*(32-ay)*32; w01 = ax*(32-ay)*32; w10 = (32-ax)*ay*32; w11 = ax*ay*32
            v = px(sy,sx)*w00 + px(sy,sx+1)*w01 + px(sy+1,sx)*w10 + px(sy+1,sx+1)*w11
            out[y,x] = (v + 16384) >> 15
    return out

def cross(o, a, b):
    return (a[0]-o[0])*(b[1]-o[1]) - (a[1]-o[1])*(b[0]-o[0])
def hull_from_rows(ys, xmin, xmax):
    """rows sorted by y ascending. Returns hull in cv2.convexHull(clockwise=False) order."""
    # right chain: top -> bottom along the right side; left chain: bottom -> top along left side
    # Walk order wanted: start at max-x point (max y among ties), then increasing y along right side to bottom,
    # then along left side going up, then along the top back to right.
    # Build full cycle: right chain (y asc) then left chain (y desc); screen-clockwise => cross sign check.
    right = []
    for y, x in zip(ys, xmax):
        p = (x, y)
        while len(right) >= 2 and cross(right[-2], right[-1], p) <= 0: right.pop()
        right.append(p)
    left = []
    for y, x in zip(ys[::-1], xmin[::-1]):
        p = (x, y)
        while len(left) >= 2 and cross(left[-2], left[-1], p) <= 0: left.pop()
        left.append(p)
    cyc = right + left
    # remove duplicates at joins
    out = []
    for p in cyc:
        if not out or out[-1] != p: out.append(p)
    if len(out) > 1 and out[0] == out[-1]: out.pop()
    # joins can create collinear/concave points: run a cleanup pass
    changed = True
    while changed and len(out) > 2:
        changed = False
        n = len(out)
        for i in range(n):
            if cross(out[i-1], out[i], out[(i+1) % n]) <= 0:
                out.pop(i); changed = True; break
    # rotate to start at max x (max y among ties)
    first = (xmin[0], ys[0])          # first raster pixel of the blob = contour start
    k = out.index(first)
    out = out[k+1:] + out[:k+1]
    return np.array(out, np.int32)

def rotating_calipers(points):
    """points: (n,2) float32 convex hull. returns out[6] float32 like OpenCV's rotatingCalipers(MINAREARECT)."""
    n = len(points)
    pts = points.astype(np.float32)
    vect = np.zeros((n,2), np.float32); inv_len = np.zeros(n, np.float32)
    left = bottom = right = top = 0
    pt0 = pts[0]
    left_x = right_x = pt0[0]; top_y = bottom_y = pt0[1]
    for i in range(n):
        if pt0[0] < left_x: left_x = pt0[0]; left = i
        if pt0[0] > right_x: right_x = pt0[0]; right = i
        if pt0[1] > top_y: top_y = pt0[1]; top = i
        if pt0[1] < bottom_y: bottom_y = pt0[1]; bottom = i
        pt = pts[(i+1) % n]
        dx = float(pt[0]) - float(pt0[0]); dy = float(pt[1]) - float(pt0[1])
        vect[i] = (f32(dx), f32(dy))
        inv_len[i] = f32(1.0/np.sqrt(dx*dx+dy*dy))
        pt0 = pt
    orientation = f32(0)
    ax = float(vect[n-1][0]); ay = float(vect[n-1][1])
    for i in range(n):
        bx = float(vect[i][0]); by = float(vect[i][1])
        conv = ax*by - ay*bx
        if conv != 0:
            orientation = f32(1) if conv > 0 else f32(-1); break
        ax, ay = bx, by
    assert orientation != 0
    base_a = orientation; base_b = f32(0)
    seq = [bottom, right, top, left]
    minarea = f32(np.finfo(np.float32).max)
    best = None
    for k in range(n):
        dp = [ base_a*vect[seq[0]][0] + base_b*vect[seq[0]][1],
              -base_b*vect[seq[1]][0] + base_a*vect[seq[1]][1],
              -base_a*vect[seq[2]][0] - base_b*vect[seq[2]][1],
               base_b*vect[seq[3]][0] - base_a*vect[seq[3]][1]]
        maxcos = dp[0]*inv_len[seq[0]]; main = 0
        for i in range(1,4):
            c = dp[i]*inv_len[seq[i]]
            if c > maxcos: main = i; maxcos = c
        pidx = seq[main]
        lead_x = vect[pidx][0]*inv_len[pidx]; lead_y = vect[pidx][1]*inv_len[pidx]
        if main == 0: base_a, base_b = lead_x, lead_y
        elif main == 1: base_a, base_b = lead_y, -lead_x
        elif main == 2: base_a, base_b = -lead_x, -lead_y
        else: base_a, base_b = -lead_y, lead_x
        seq[main] = (seq[main]+1) % n
        dx = pts[seq[1]][0] - pts[seq[3]][0]; dy = pts[seq[1]][1] - pts[seq[3]][1]
        width = dx*base_a + dy*base_b
        dx = pts[seq[2]][0] - pts[seq[0]][0]; dy = pts[seq[2]][1] - pts[seq[0]][1]
        height = -dx*base_b + dy*base_a
        area = width*height
        if area <= minarea:
            minarea = area
            best = (seq[3], base_a, width, base_b, height, seq[0])
    l, A1, wdt, B1, hgt, b = best
    A2 = -B1; B2 = A1
    C1 = A1*pts[l][0] + pts[l][1]*B1
    C2 = A2*pts[b][0] + pts[b][1]*B2
    idet = f32(1)/(A1*B2 - A2*B1)
    px = (C1*B2 - C2*B1)*idet; py = (A1*C2 - A2*C1)*idet
    return np.array([px, py, A1*wdt, B1*wdt, A2*hgt, B2*hgt], np.float32)

def min_area_rect(hull):
    hull = hull.astype(np.float32)
    n = len(hull)
    if n > 2:
        out = rotating_calipers(hull)
        cx = out[0] + (out[2]+out[4])*f32(0.5); cy = out[1] + (out[3]+out[5])*f32(0.5)
        w = f32(np.sqrt(float(out[2])*float(out[2]) + float(out[3])*float(out[3])))
        h = f32(np.sqrt(float(out[4])*float(out[4]) + float(out[5])*float(out[5])))
        ang = f32(np.arctan2(float(out[3]), float(out[2])))
    elif n == 2:
        cx = (hull[0][0]+hull[1][0])*f32(0.5); cy = (hull[0][1]+hull[1][1])*f32(0.5)
        dx = float(hull[1][0])-float(hull[0][0]); dy = float(hull[1][1])-float(hull[0][1])
        w = f32(np.sqrt(dx*dx+dy*dy)); h = f32(0); ang = f32(np.arctan2(dy,dx))
    else:
        cx, cy = hull[0]; w = h = f32(0); ang = f32(0)
    ang = f32(float(ang)*180/np.pi)
    return (cx, cy), (w, h), ang

def box_points(rect):
    (cx, cy), (w, h), ang = rect
    a_ = float(ang)*np.pi/180.
    b = f32(np.cos(a_))*f32(0.5); a = f32(np.sin(a_))*f32(0.5)
    p0 = (cx - a*h - b*w, cy + b*h - a*w)
    p1 = (cx + a*h - b*w, cy - b*h - a*w)
    p2 = (f32(2)*cx - p0[0], f32(2)*cy - p0[1])
    p3 = (f32(2)*cx - p1[0], f32(2)*cy - p1[1])
    return np.array([p0,p1,p2,p3], np.float32)
