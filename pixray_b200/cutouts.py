"""Host-side sampling of MakeCutouts' per-cutout augmentation parameters (numpy).

Mirrors pixray_b200/csrc/transforms.h (the engine's own Philox-driven sampler): the *distributions* kornia's
RandomPerspective / RandomResizedCrop / RandomAffine draw from at the reference's call sites (pixray.py:411-437,
SURVEY.md Appendix A), composed into one 3x3 "dst_pix <- src_pix" homography per cutout -- what
MakeCutouts.transforms caches (pixray.py:498).  Used to feed explicit parameters to both the engine and the oracle.
"""
import numpy as np


def _perspective_from_points(src, dst):
    A, b = np.zeros((8, 8)), np.zeros(8)
    for k in range(4):
        x, y = src[k]
        u, v = dst[k]
        A[2 * k] = [x, y, 1, 0, 0, 0, -u * x, -u * y]
        A[2 * k + 1] = [0, 0, 0, x, y, 1, -v * x, -v * y]
        b[2 * k], b[2 * k + 1] = u, v
    h = np.linalg.solve(A, b)
    return np.append(h, 1.0).reshape(3, 3)


def _sample_perspective(g, size, distortion, p):
    apply = g.uniform() <= p
    w1, f = size - 1.0, distortion * size / 2.0
    start = np.array([[0, 0], [w1, 0], [w1, w1], [0, w1]], dtype=np.float64)
    sgn = np.array([[1, 1], [-1, 1], [-1, -1], [1, -1]], dtype=np.float64)
    end = start + f * g.uniform(size=(4, 2)) * sgn
    return _perspective_from_points(start, end) if apply else np.eye(3)


def _sample_resized_crop(g, size):
    w = h = float(size)
    for _ in range(10):
        area = g.uniform(0.25, 0.95) * size * size
        ratio = np.exp(g.uniform(np.log(0.85), np.log(1.2)))
        cw, ch = round(np.sqrt(area * ratio)), round(np.sqrt(area / ratio))
        if 0 < cw <= size and 0 < ch <= size:
            w, h = float(cw), float(ch)
            break
    x0, y0 = np.floor(g.uniform() * (size - w + 1) * 0.999999), np.floor(g.uniform() * (size - h + 1) * 0.999999)
    src = np.array([[x0, y0], [x0 + w - 1, y0], [x0 + w - 1, y0 + h - 1], [x0, y0 + h - 1]])
    s1 = size - 1.0
    return _perspective_from_points(src, np.array([[0, 0], [s1, 0], [s1, s1], [0, s1]]))


def _sample_affine(g, size, n_s, n_t):
    c = size / 2.0 - 0.5
    tx, ty = g.uniform(-n_t * size, n_t * size, 2)
    return np.array([[n_s, 0, (1 - n_s) * c + tx], [0, n_s, (1 - n_s) * c + ty], [0, 0, 1.0]])


def sample_transforms(cutn, cut_size, seed):
    """[cutn, 3, 3] float32; zoom group first: global index < int(0.6 * cutn) (pixray.py:407, 493-494)."""
    g = np.random.default_rng(seed)
    out = np.zeros((cutn, 3, 3), dtype=np.float32)
    zoom = int(0.6 * cutn)
    for n in range(cutn):
        if n < zoom:
            H = _sample_resized_crop(g, cut_size) @ _sample_perspective(g, cut_size, 0.40, 0.7)
        else:
            n_s = 0.95
            H = _sample_perspective(g, cut_size, 0.20, 0.7) @ _sample_affine(g, cut_size, n_s, (1 - n_s) / 2)
        out[n] = H
    return out


def jitter_code(order):
    """Application order (a permutation of 0 brightness, 1 contrast, 2 saturation, 3 hue) -> the engine's code."""
    return 256 + order[0] + 4 * order[1] + 16 * order[2] + 64 * order[3]


def sample_color_jitter(cutn, seed, p=0.8, saturation=0.1, hue=0.1):
    """[cutn, 3] float32 rows {code, saturation_factor, hue_factor} of K.ColorJitter(hue=0.1, saturation=0.1, p=0.8),
    the last stage of both augmentation stacks (pixray.py:416, 436): Bernoulli(p) per cutout (code 0 = missed),
    saturation_factor ~ U(1-s, 1+s), hue_factor ~ U(-h, h), one randperm(4) order per stack and call."""
    g = np.random.default_rng(seed)
    zoom = int(0.6 * cutn)
    codes = [jitter_code(list(g.permutation(4))) for _ in range(2)]
    out = np.zeros((cutn, 3), dtype=np.float32)
    for n in range(cutn):
        apply = g.uniform() <= p
        out[n] = [codes[0 if n < zoom else 1] if apply else 0, g.uniform(1 - saturation, 1 + saturation), g.uniform(-hue, hue)]
    return out
