"""Seeded synthetic inputs for the parity tests and the benchmark (test infrastructure).

* ``score_maps``   -- CRAFT-like (text, link) heat-maps built from per-character Gaussians,
                      in the spirit of the reference's training-label synthesis
                      (detection.py:55-62,106-198), so ``getBoxes`` sees realistic components.
* ``text_images``  -- white pages with words rendered by ``cv2.putText``.
* ``random_quads`` -- rotated rectangles for ``warpBox`` parity.
"""
import cv2
import numpy as np

ALPHABET = "0123456789abcdefghijklmnopqrstuvwxyz"


def _gaussian_tile(size=64, ratio=1.8):
    v = np.abs(np.linspace(-size / 2, size / 2, num=size))
    gx, gy = np.meshgrid(v, v)
    g = np.sqrt(gx ** 2 + gy ** 2) * (ratio / (size / 2))
    return np.exp(-0.5 * g ** 2).astype(np.float32)


def _stamp(canvas, tile, quad):
    src = np.array([[0, 0], [tile.shape[1], 0], [tile.shape[1], tile.shape[0]], [0, tile.shape[0]]],
                   dtype=np.float32)
    M = cv2.getPerspectiveTransform(src, quad.astype(np.float32))
    canvas += cv2.warpPerspective(tile, M, dsize=(canvas.shape[1], canvas.shape[0]))


def _rot(points, centre, angle):
    c, s = np.cos(angle), np.sin(angle)
    R = np.array([[c, -s], [s, c]])
    return (points - centre) @ R.T + centre


def score_map(rng, h, w, n_words, max_angle=0.5, peak=(0.75, 1.0), weak_fraction=0.1):
    """One (h,w,2) float32 score map with ``n_words`` words placed on a jittered grid.

    ``weak_fraction`` of the words get a peak below the 0.7 detection threshold and some words
    are tiny, so that the size/max filters of getBoxes (detection.py:233-241) are exercised.
    """
    tile = _gaussian_tile()
    text = np.zeros((h, w), np.float32)
    link = np.zeros((h, w), np.float32)
    cols = int(np.ceil(np.sqrt(n_words * w / h)))
    rows = int(np.ceil(n_words / cols))
    cell_w, cell_h = w / cols, h / rows
    for k in range(n_words):
        gx, gy = k % cols, k // cols
        n_chars = int(rng.integers(2, 9))
        ch_h = float(rng.uniform(0.25, 0.5) * cell_h)
        ch_w = min(ch_h * float(rng.uniform(0.5, 0.9)), 0.8 * cell_w / n_chars)
        centre = np.array([(gx + 0.5) * cell_w + rng.uniform(-0.05, 0.05) * cell_w,
                           (gy + 0.5) * cell_h + rng.uniform(-0.05, 0.05) * cell_h])
        angle = float(rng.uniform(-max_angle, max_angle)) if rng.random() < 0.6 else 0.0
        amp = float(rng.uniform(*peak))
        if rng.random() < weak_fraction:
            amp = float(rng.uniform(0.45, 0.68))
        x0 = centre[0] - n_chars * ch_w / 2
        prev = None
        for i in range(n_chars):
            cx = x0 + (i + 0.5) * ch_w
            quad = np.array([[cx - ch_w / 2, centre[1] - ch_h / 2], [cx + ch_w / 2, centre[1] - ch_h / 2],
                             [cx + ch_w / 2, centre[1] + ch_h / 2], [cx - ch_w / 2, centre[1] + ch_h / 2]])
            quad = _rot(quad, centre, angle)
            _stamp(text, tile * amp, quad)
            cur = np.array([cx, centre[1]])
            if prev is not None:
                lq = np.array([[prev[0], centre[1] - ch_h / 4], [cur[0], centre[1] - ch_h / 4],
                               [cur[0], centre[1] + ch_h / 4], [prev[0], centre[1] + ch_h / 4]])
                _stamp(link, tile * amp, _rot(lq, centre, angle))
            prev = cur
    return np.stack([text.clip(0, 1), link.clip(0, 1)], -1)


def score_maps(seed, n, h, w, n_words):
    rng = np.random.default_rng(seed)
    return np.stack([score_map(rng, h, w, n_words) for _ in range(n)])


def random_word(rng, lo=3, hi=10):
    return "".join(ALPHABET[i] for i in rng.integers(0, len(ALPHABET), int(rng.integers(lo, hi + 1))))


def text_image(rng, h, w, n_words, return_layout=False):
    """White RGB page with ``n_words`` random words on a jittered grid.  Returns (image, words), or with
    ``return_layout`` (image, words, rects) where rects[k] = (x0, y0, x1, y1) bounds word k's glyphs in page pixels
    (same random stream either way, so the pages are identical)."""
    img = np.full((h, w, 3), 255, np.uint8)
    cols = max(1, int(np.floor(np.sqrt(n_words * w / h / 2.0))))
    rows = int(np.ceil(n_words / cols))
    cell_w, cell_h = w / cols, h / rows
    words, rects = [], []
    for k in range(n_words):
        gx, gy = k % cols, k // cols
        word = random_word(rng)
        scale = 0.9
        (tw, th), _ = cv2.getTextSize(word, cv2.FONT_HERSHEY_SIMPLEX, scale, 2)
        fit = min(0.8 * cell_w / tw, 0.45 * cell_h / th)
        scale *= fit
        (tw, th), base = cv2.getTextSize(word, cv2.FONT_HERSHEY_SIMPLEX, scale, 2)
        x = int(gx * cell_w + (cell_w - tw) / 2 + rng.uniform(-0.05, 0.05) * cell_w)
        y = int(gy * cell_h + (cell_h + th) / 2 + rng.uniform(-0.05, 0.05) * cell_h)
        colour = tuple(int(c) for c in rng.integers(0, 90, 3))
        cv2.putText(img, word, (x, y), cv2.FONT_HERSHEY_SIMPLEX, scale, colour, 2, cv2.LINE_AA)
        words.append(word)
        rects.append((x, y - th, x + tw, y + base))
    if return_layout:
        return img, words, rects
    return img, words


def text_images(seed, n, h, w, n_words):
    rng = np.random.default_rng(seed)
    pages = [text_image(rng, h, w, n_words) for _ in range(n)]
    return np.stack([p[0] for p in pages]), [p[1] for p in pages]


def random_quads(rng, n, h, w, min_side=12, max_side=300):
    """(n,4,2) float32 rotated rectangles inside an h x w image, arbitrary starting corner."""
    out = []
    for _ in range(n):
        bw = rng.uniform(min_side, min(max_side, w / 2))
        bh = rng.uniform(min_side, min(bw, 80))
        ang = rng.uniform(-np.pi, np.pi) if rng.random() < 0.7 else 0.0
        c = np.array([rng.uniform(bw / 2, w - bw / 2), rng.uniform(bw / 2, h - bw / 2)])
        quad = np.array([[-bw / 2, -bh / 2], [bw / 2, -bh / 2], [bw / 2, bh / 2], [-bw / 2, bh / 2]]) + c
        quad = _rot(quad, c, ang)
        out.append(np.roll(quad, int(rng.integers(0, 4)), 0))
    return np.array(out, dtype=np.float32)


def noise_gray(rng, h, w):
    """Blurred-noise gray image (uint8) so that bilinear sampling differences are visible."""
    g = rng.integers(0, 256, (h, w)).astype(np.float32)
    return cv2.GaussianBlur(g, (0, 0), 2.0).clip(0, 255).astype(np.uint8)
