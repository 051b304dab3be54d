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


def source_size(cut_size, aspect=1.0):
    """(height, width) of the tensor the warps sample from: the pooled cut_size x cut_size image, stretched by
    kornia.geometry.transform.rescale when the canvas is not square (global_aspect_width = size[0] / size[1],
    pixray.py:468-472, 1931; rescale truncates: int(size * factor))."""
    if aspect == 1.0:
        return cut_size, cut_size
    if aspect > 1.0:
        return cut_size, int(cut_size * aspect)
    return int(cut_size * (1.0 / aspect)), cut_size


def _sample_perspective_hw(g, h, w, distortion, p):
    apply = g.uniform() <= p
    fx, fy = distortion * w / 2.0, distortion * h / 2.0
    start = np.array([[0, 0], [w - 1.0, 0], [w - 1.0, h - 1.0], [0, h - 1.0]], dtype=np.float64)
    sgn = np.array([[1, 1], [-1, 1], [-1, -1], [1, -1]], dtype=np.float64)
    end = start + np.array([fx, fy]) * g.uniform(size=(4, 2)) * sgn
    return _perspective_from_points(start, end) if apply else np.eye(3)


def _sample_resized_crop_hw(g, h, w, out):
    """RandomResizedCrop(size=(out, out), scale=(0.25, 0.95), ratio=(0.85, 1.2)) of an h x w source: ten tries for a box
    that fits, else the central-ratio fallback (whole height or width, ratio clamped), then a uniform position."""
    cw = ch = None
    for _ in range(10):
        area = g.uniform(0.25, 0.95) * h * w
        ratio = np.exp(g.uniform(np.log(0.85), np.log(1.2)))
        tw, th = round(np.sqrt(area * ratio)), round(np.sqrt(area / ratio))
        if 0 < tw <= w and 0 < th <= h:
            cw, ch = float(tw), float(th)
            break
    if cw is None:
        in_ratio = w / h
        if in_ratio < 0.85:
            cw, ch = float(w), float(round(w / 0.85))
        elif in_ratio > 1.2:
            ch, cw = float(h), float(round(h * 1.2))
        else:
            cw, ch = float(w), float(h)
    x0, y0 = np.floor(g.uniform() * (w - cw + 1) * 0.999999), np.floor(g.uniform() * (h - ch + 1) * 0.999999)
    src = np.array([[x0, y0], [x0 + cw - 1, y0], [x0 + cw - 1, y0 + ch - 1], [x0, y0 + ch - 1]])
    s1 = out - 1.0
    return _perspective_from_points(src, np.array([[0, 0], [s1, 0], [s1, s1], [0, s1]]))


def _sample_affine_hw(g, h, w, scale_lo, scale_hi, tfx, tfy):
    """MyRandomAffine(degrees=0, translate=(tfx, tfy), scale=(lo, hi)) about the centre of an h x w image."""
    s = g.uniform(scale_lo, scale_hi)
    cx, cy = w / 2.0 - 0.5, h / 2.0 - 0.5
    tx, ty = g.uniform(-tfx * w, tfx * w) if tfx > 0 else 0.0, g.uniform(-tfy * h, tfy * h) if tfy > 0 else 0.0
    return np.array([[s, 0, (1 - s) * cx + tx], [0, s, (1 - s) * cy + ty], [0, 0, 1.0]])


def _center_crop(h, w, out):
    """K.CenterCrop(size=out, cropping_mode='resample'): the central out x out box, a pure translation."""
    return np.array([[1, 0, -float((w - out) // 2)], [0, 1, -float((h - out) // 2)], [0, 0, 1.0]])


def sample_transforms(cutn, cut_size, seed, aspect=1.0):
    """[cutn, 3, 3] float32; zoom group first: global index < int(0.6 * cutn) (pixray.py:407, 493-494).  `aspect` =
    global_aspect_width (canvas width / height): the source of the warps is then source_size(cut_size, aspect) and the wide
    stack shrinks by 1 / aspect (or aspect) so the whole canvas fits the square cutout (pixray.py:420-432)."""
    g = np.random.default_rng(seed)
    out = np.zeros((cutn, 3, 3), dtype=np.float32)
    zoom = int(0.6 * cutn)
    h, w = source_size(cut_size, aspect)
    for n in range(cutn):
        if aspect == 1.0:
            if n < zoom:
                H = _sample_resized_crop(g, cut_size) @ _sample_perspective(g, cut_size, 0.40, 0.7)
            else:
                n_s = 0.95
                H = _sample_perspective(g, cut_size, 0.20, 0.7) @ _sample_affine(g, cut_size, n_s, (1 - n_s) / 2)
        elif n < zoom:
            H = _sample_resized_crop_hw(g, h, w, cut_size) @ _sample_perspective_hw(g, h, w, 0.40, 0.7)
        else:
            if aspect > 1.0:
                n_s = 1.0 / aspect
                A = _sample_affine_hw(g, h, w, 0.9 * n_s, n_s, 0.0, (1 - n_s) / 2)   # translate=(0, n_t), pixray.py:424-427
            else:
                n_s = aspect
                A = _sample_affine_hw(g, h, w, 0.9 * n_s, n_s, (1 - n_s) / 2, 0.0)   # translate=(n_t, 0), pixray.py:428-431
            H = _sample_perspective(g, cut_size, 0.20, 0.7) @ _center_crop(h, w, cut_size) @ A
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
