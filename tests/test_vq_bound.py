"""The rounding bound behind the tensor-core nearest-code search (pixray_b200/csrc/kernels_vq_tc.cu): vq_select keeps every
code whose APPROXIMATE distance (fp16 operands, fp32 accumulate) lies within 4 e of the minimum, where

    e = 2^-9 |x|_2 |c|_2 + 2^-22 (|x~|_1 + |c~|_1) / (scale_x scale_c)        (x~, c~: the power-of-two scaled operands)

bounds |x.c~ - x.c|.  If that bound holds, the exact fp32 arg-min is always among the candidates, so the index the engine
picks is the one the all-fp32 search picks.  Here the bound is checked numerically, on the CPU, with the same scaling and the
same fp16 rounding, on data chosen to stress it: unit-scale codes, taming's tiny uniform initialisation (values far below
fp16's normal range before scaling), heavy-tailed latents, latents far outside the codebook, sparse rows and near-duplicate
codes -- and the candidate rule is run end to end against the exact arg-min."""
import zlib

import numpy as np
import pytest


def pow2_scale(amax):
    """the power of two that brings amax into [0.5, 1): frexp, as vq_prep_kernel and build_vqgan do"""
    if not np.isfinite(amax) or amax <= 0:
        return 1.0
    _, e = np.frexp(np.float32(amax))
    return float(np.ldexp(1.0, -int(e)))


def approx_scores(z, cb):
    """z [P, C], cb [N, C] fp32 -> (scores of the scaled fp16 operands accumulated in fp32, per-position scale, cb scale)"""
    sx = np.array([pow2_scale(np.abs(r).max()) for r in z], dtype=np.float32)
    sc = np.float32(pow2_scale(np.abs(cb).max()))
    zh = (z * sx[:, None]).astype(np.float16).astype(np.float32)
    ch = (cb * sc).astype(np.float16).astype(np.float32)
    # fp32 accumulation of exact fp16 products (the tensor core's accumulator is at least that precise)
    return (zh.astype(np.float64) @ ch.astype(np.float64).T).astype(np.float32), sx, sc


def eps_dot(z, cb, sx, sc):
    xn2 = np.sqrt((z.astype(np.float64) ** 2).sum(1)) * 1.0000002
    xn1 = np.abs(z).astype(np.float64).sum(1) * 1.0001
    cmax2 = np.sqrt((cb.astype(np.float64) ** 2).sum(1)).max() * 1.000001
    cmax1 = np.abs(cb).astype(np.float64).sum(1).max() * 1.000001
    inv = 1.0 / (sx.astype(np.float64) * sc)
    return 2.0 ** -9 * xn2 * cmax2 + 2.0 ** -22 * inv * (xn1 * sx + cmax1 * sc), cmax2


CASES = {
    "unit-scale codes, latents on a code + noise": lambda g: (None, g.normal(size=(512, 256)).astype(np.float32) * 0.5, 0.05),
    "taming init: uniform(-1/n, 1/n)": lambda g: (None, g.uniform(-1 / 16384, 1 / 16384, size=(512, 256)).astype(np.float32), 0.05),
    "heavy-tailed latents": lambda g: (g.standard_t(2, size=(64, 256)).astype(np.float32), g.normal(size=(512, 256)).astype(np.float32), None),
    "latents 40x outside the codebook": lambda g: (g.normal(size=(64, 256)).astype(np.float32) * 40, g.normal(size=(512, 256)).astype(np.float32) * 0.5, None),
    "sparse rows with outliers": lambda g: (g.normal(size=(64, 256)).astype(np.float32) * (g.random((64, 256)) < 0.03) * 100,
                                            g.normal(size=(512, 256)).astype(np.float32) * (g.random((512, 256)) < 0.2), None),
    "wide dynamic range inside a row": lambda g: (g.normal(size=(64, 256)).astype(np.float32) * np.exp(g.normal(size=(64, 256)) * 4).astype(np.float32),
                                                  g.normal(size=(512, 256)).astype(np.float32) * np.exp(g.normal(size=(512, 256)) * 4).astype(np.float32), None),
}


@pytest.mark.parametrize("name", list(CASES))
def test_fp16_dot_error_stays_inside_the_bound(name):
    g = np.random.default_rng(zlib.crc32(name.encode()))
    z, cb, sigma = CASES[name](g)
    if z is None:
        pick = g.integers(0, cb.shape[0], 64)
        z = (cb[pick] + sigma * cb.std() * g.normal(size=(64, cb.shape[1]))).astype(np.float32)
    s, sx, sc = approx_scores(z, cb)
    exact = z.astype(np.float64) @ cb.astype(np.float64).T
    approx = s.astype(np.float64) / (sx.astype(np.float64)[:, None] * sc)
    e, _ = eps_dot(z, cb, sx, sc)
    ratio = (np.abs(approx - exact) / e[:, None]).max()
    print(f"[vq bound] {name}: max |error| / e = {ratio:.3f}")
    assert ratio <= 0.75, "the bound is meant to be about twice the worst case"


def test_candidate_rule_always_contains_the_exact_argmin():
    """End to end: d~ = |c|^2 - 2 x.c~, threshold min d~ + 4 e + 1e-5 (|x|^2 + cmax^2) as vq_select_kernel computes it;
    near-duplicate codes make the exact minimum a near-tie."""
    g = np.random.default_rng(5)
    cb = g.normal(size=(2048, 256)).astype(np.float32) * 0.5
    cb[1024:1536] = cb[:512] + g.normal(size=(512, 256)).astype(np.float32) * 1e-4      # near-duplicates of the first rows
    pick = g.integers(0, 512, 256)
    z = (cb[pick] + 0.02 * cb.std() * g.normal(size=(256, 256))).astype(np.float32)
    s, sx, sc = approx_scores(z, cb)
    e, cmax2 = eps_dot(z, cb, sx, sc)
    c2 = (cb.astype(np.float64) ** 2).sum(1)
    x2 = (z.astype(np.float64) ** 2).sum(1)
    d_approx = c2[None, :] - 2.0 * s.astype(np.float64) / (sx.astype(np.float64)[:, None] * sc)
    thr = d_approx.min(1) + 4 * e + 1e-5 * (x2 + cmax2 ** 2)
    d_exact = c2[None, :] - 2.0 * (z.astype(np.float64) @ cb.astype(np.float64).T)
    best = d_exact.argmin(1)
    inside = d_approx[np.arange(len(z)), best] <= thr
    ncand = (d_approx <= thr[:, None]).sum(1)
    print(f"[vq bound] candidates per position: mean {ncand.mean():.2f}, max {ncand.max()}")
    assert inside.all()
    assert ncand.max() <= 256, "the kernel's candidate list holds 256 entries before it falls back to the full search"
