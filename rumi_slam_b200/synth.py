"""Deterministic synthetic grayscale frames of the shapes BASELINE.json names.

Band-limited noise (three octaves) plus ~100 filled rectangles / triangles, clipped to u8.  Tuned so that
level 0 of a 640x480 frame yields a few thousand FAST candidates, some 35-px cells are empty at
iniThFAST=20 (exercising the minThFAST fallback) and the quad-tree's sorted final phase triggers.
numpy + scipy only; seeds are `np.random.default_rng(seed)`.
"""
import numpy as np
from scipy import ndimage


def _octave(rng, h, w, down, sigma):
    hh, ww = -(-h // down) + 2, -(-w // down) + 2
    n = ndimage.gaussian_filter(rng.standard_normal((hh, ww)).astype(np.float32), sigma, mode="reflect")
    n /= n.std() + 1e-6
    if down > 1:
        n = ndimage.zoom(n, down, order=1)
    return n[:h, :w]


def synthetic_frame(seed, w=640, h=480, texture=1.0, shapes=100):
    rng = np.random.default_rng(seed)
    img = 22.0 * texture * _octave(rng, h, w, 1, 2.5)
    img += 18.0 * _octave(rng, h, w, 2, 3.0)
    img += 30.0 * _octave(rng, h, w, 5, 3.0)
    # smooth low-texture mask so that some cells fall below iniThFAST
    m = _octave(rng, h, w, 8, 2.0)
    img *= np.clip(0.55 + 0.6 * m, 0.05, 1.4)
    img += 128.0
    yy, xx = np.mgrid[0:h, 0:w]
    for _ in range(shapes):
        val = float(rng.integers(20, 236))
        cx, cy = int(rng.integers(0, w)), int(rng.integers(0, h))
        sw, sh = int(rng.integers(6, 70)), int(rng.integers(6, 70))
        x0, x1 = max(cx - sw // 2, 0), min(cx + sw // 2 + 1, w)
        y0, y1 = max(cy - sh // 2, 0), min(cy + sh // 2 + 1, h)
        if x1 <= x0 or y1 <= y0:
            continue
        if rng.random() < 0.5:
            img[y0:y1, x0:x1] = val
        else:
            px = rng.integers(x0, x1, 3).astype(np.float32)
            py = rng.integers(y0, y1, 3).astype(np.float32)
            sx, sy = xx[y0:y1, x0:x1].astype(np.float32), yy[y0:y1, x0:x1].astype(np.float32)

            def edge(i, j):
                return (px[j] - px[i]) * (sy - py[i]) - (py[j] - py[i]) * (sx - px[i])

            e0, e1, e2 = edge(0, 1), edge(1, 2), edge(2, 0)
            inside = ((e0 >= 0) & (e1 >= 0) & (e2 >= 0)) | ((e0 <= 0) & (e1 <= 0) & (e2 <= 0))
            img[y0:y1, x0:x1][inside] = val
    img += 2.0 * rng.standard_normal((h, w)).astype(np.float32)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def synthetic_batch(n, w=640, h=480, seed0=0, unique=None):
    """n frames [n,h,w] u8.  `unique` < n generates that many distinct frames and fills the rest with
    circular shifts of them (distinct content, cheap to make) -- used by bench.py for the 1024-frame batch."""
    unique = n if unique is None else min(unique, n)
    if unique >= 8:                                  # scipy / numpy release the GIL in the filters: ~3x on 8 cores
        import os
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(16, len(os.sched_getaffinity(0)))) as ex:
            base = list(ex.map(lambda i: synthetic_frame(seed0 + i, w, h), range(unique)))
    else:
        base = [synthetic_frame(seed0 + i, w, h) for i in range(unique)]
    out = np.empty((n, h, w), np.uint8)
    for i in range(n):
        b = base[i % unique]
        k = i // unique
        out[i] = b if k == 0 else np.roll(b, (7 * k, 13 * k), axis=(0, 1))
    return out


def stereo_pair(seed, w=752, h=480, max_disp=40):
    """Left frame + right frame = left warped by a per-row-constant disparity plus independent noise."""
    left = synthetic_frame(2 * seed, w, h)
    rng = np.random.default_rng(2 * seed + 1)
    disp = ndimage.gaussian_filter1d(rng.uniform(2, max_disp, h).astype(np.float32), 25.0, mode="reflect")
    right = np.empty_like(left)
    for y in range(h):
        right[y] = np.roll(left[y], -int(round(float(disp[y]))))
    right = np.clip(right.astype(np.float32) + 1.5 * rng.standard_normal((h, w)), 0, 255)
    return left, np.rint(right).astype(np.uint8)


def perturbed_descriptors(desc, n, seed=7, flip_p=0.1):
    """Tile descriptor rows to n rows and flip each bit with probability flip_p (SURVEY.md 8d cfg 5b)."""
    rng = np.random.default_rng(seed)
    reps = -(-n // len(desc))
    d = np.tile(desc, (reps, 1))[:n].copy()
    flips = np.packbits(rng.random((n, 256)) < flip_p, axis=1, bitorder="little")
    return d ^ flips


def synthetic_vocabulary(k=10, L=6, seed=0, flip=0.08, stop_every=0):
    """A DBoW2-shaped vocabulary tree (ORBvoc.txt is a missing blob of the reference): k children per node, L levels,
    nodes in breadth-first id order (node 0 = root, parent id < child id, children of a node contiguous), child
    descriptors = parent with a fraction `flip` of the 256 bits flipped, leaf weights ~ idf-like positive values
    (every `stop_every`-th word gets weight 0 = a stopped word).  Returns (parent, is_leaf, desc, weight)."""
    rng = np.random.default_rng(seed)
    parent = [np.zeros(1, np.int32)]
    leaf = [np.zeros(1, np.uint8)]
    desc = [np.zeros((1, 32), np.uint8)]
    weight = [np.zeros(1, np.float64)]
    cur_desc = rng.integers(0, 256, (1, 32), dtype=np.uint8)     # virtual root centre
    cur_ids = np.zeros(1, np.int64)
    nxt = 1
    for lev in range(1, L + 1):
        n = len(cur_ids) * k
        pd = np.repeat(cur_desc, k, axis=0)
        mask = np.packbits(rng.random((n, 256)) < flip, axis=1)
        d = pd ^ mask
        parent.append(np.repeat(cur_ids, k).astype(np.int32))
        leaf.append(np.full(n, 1 if lev == L else 0, np.uint8))
        desc.append(d)
        w = np.zeros(n, np.float64)
        if lev == L:
            w = 0.5 + 4.0 * rng.random(n)
            if stop_every:
                w[::stop_every] = 0.0
        weight.append(w)
        cur_desc, cur_ids = d, np.arange(nxt, nxt + n, dtype=np.int64)
        nxt += n
    return np.concatenate(parent), np.concatenate(leaf), np.concatenate(desc), np.concatenate(weight)


def descriptors_near_vocabulary(desc_nodes, is_leaf, n, seed=0, flip=0.05):
    """n descriptors = random leaves of the vocabulary with a few bits flipped (realistic descent: clear winners and
    near ties)."""
    rng = np.random.default_rng(seed)
    leaves = np.flatnonzero(is_leaf)
    pick = rng.choice(leaves, n)
    mask = np.packbits(rng.random((n, 256)) < flip, axis=1)
    return desc_nodes[pick] ^ mask


def motion_sequence(n, w=640, h=480, seed=0, vx=1.3, vy=-0.7, noise=2):
    """n frames of a camera translating over one large synthetic scene with sub-pixel velocity (vx, vy) px/frame
    (bilinear crops + independent sensor noise): the input KFDSample::Step sees on untracked frames."""
    m = int(np.ceil(n * max(abs(vx), abs(vy)))) + 2
    scene = synthetic_frame(seed, w + 2 * m, h + 2 * m).astype(np.float32)
    rng = np.random.default_rng(seed + 1)
    out = np.empty((n, h, w), np.uint8)
    for k in range(n):
        ox, oy = m + k * vx, m + k * vy
        ix, iy = int(np.floor(ox)), int(np.floor(oy))
        fx, fy = np.float32(ox - ix), np.float32(oy - iy)
        c = lambda a, b: scene[iy + a:iy + a + h, ix + b:ix + b + w]
        f = (1 - fx) * (1 - fy) * c(0, 0) + fx * (1 - fy) * c(0, 1) + (1 - fx) * fy * c(1, 0) + fx * fy * c(1, 1)
        out[k] = np.clip(np.rint(f) + rng.integers(-noise, noise + 1, (h, w)), 0, 255).astype(np.uint8)
    return out
