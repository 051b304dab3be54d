"""Diagnostic (not collected by pytest): where does the engine-vs-oracle difference of d loss / d image through the
ColorJitter stage come from?  Runs tests/test_color_jitter.py::test_iteration_gradient_with_color_jitter's part (a) and
splits the difference into  (1) the cutout_bwd kernel vs a CPU emulation of its algorithm fed the ENGINE's own upstream
gradient (kernel check),  (2) the upstream gradient itself (fp16 CLIP vs fp32 oracle) pushed through the oracle's exact
Jacobian (conditioning of the stage),  per cutout.   python tests/diag_color_jitter.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import test_pipeline_gpu as P  # noqa: E402
from oracle import ref_path as R  # noqa: E402
from pixray_b200 import cutouts  # noqa: E402
from test_color_jitter import _host_jitter  # noqa: E402

cutn, cs, S = 8, 224, 4096.0
vq, clip, eng, prompts, z = P.build(cutn=cutn, seed=3)
T = P.random_transforms(cutn, cs, 13)
J = cutouts.sample_color_jitter(cutn, 22, p=1.0)
g = torch.Generator().manual_seed(19)
facs, noise = P.plant_extremes(torch.rand(cutn, generator=g) * 0.1, torch.randn(cutn, 3, cs, cs, generator=g))
img = R.vqgan_synth(vq, z).detach()
Tt, Jt = torch.from_numpy(T), torch.from_numpy(J)


def oracle(jitter):
    ir = img.clone().requires_grad_(True)
    b = R.make_cutouts(ir, Tt, cs, "border", 0.3, facs, noise, jitter=Jt if jitter else None)
    b.retain_grad()
    emb = R.encode_image(clip, b).float()
    sum(R.prompt_loss(emb, *p) for p in prompts).backward()
    return ir.grad.clone(), b.grad.clone(), b.detach()


def engine(jitter):
    eng.synth(z)
    eng.make_cutouts(img, transforms=T, zoom_padding=1, fill=0.3, noise_facs=facs.numpy(), noise=noise,
                     color_jitter=J if jitter else None)
    eng.encode_image(0)
    eng.prompt_loss(0)
    eng.backward()
    gb = eng.debug_read("g_batch", (cutn, 3, cs, cs)).cpu() / S          # direct term only
    gi = eng.debug_read("g_img", (1, 3, 32, 32)).cpu() / S
    rng = eng.debug_read("range", (4,)).cpu()
    sums = eng.debug_read("sums", (4,)).cpu() / S
    ir = eng.debug_read("irange", (4,), dtype=torch.int32).cpu()
    return gi, gb, rng, sums, ir


def emulate(g_full, jitter):
    """cutout_bwd's algorithm on the CPU: VJP of the pixel body at the recomputed pre-jitter colour, then the warp's
    adjoint (autograd through the oracle's warp-only make_cutouts)."""
    ir = img.clone().requires_grad_(True)
    pre = R.make_cutouts(ir, Tt, cs, "border", 0.3)
    g_pre = torch.zeros_like(pre)
    for n in range(cutn):
        if jitter:
            rgb = pre[n].detach().reshape(3, -1).t().contiguous().numpy()
            go = g_full[n].reshape(3, -1).t().contiguous().numpy()
            _, gi = _host_jitter(rgb, int(J[n, 0]), float(J[n, 1]), float(J[n, 2]), go)
            g_pre[n] = torch.from_numpy(gi).t().reshape(3, cs, cs)
        else:
            g_pre[n] = g_full[n]
    got, = torch.autograd.grad(pre, ir, g_pre)
    return got


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


for jitter in (False, True):
    print(f"================ ColorJitter {'ON' if jitter else 'OFF'}")
    gi_ref, gb_ref, b_ref = oracle(jitter)
    gi_eng, gb_eng, rng, sums, ir = engine(jitter)
    # the engine's full upstream gradient = direct term + the range terms on the two extreme elements
    dR = -float(sums[1]) if float(rng[3]) != 0 else 0.0
    dMin = -(float(sums[0]) + dR)
    gb_eng_full = gb_eng.clone().reshape(-1)
    gb_eng_full[int(ir[1])] += dR
    gb_eng_full[int(ir[0])] += dMin
    gb_eng_full = gb_eng_full.reshape(gb_eng.shape)
    flat = b_ref.reshape(-1)
    print("extreme elements engine", ir[:2].tolist(), "oracle", [int(flat.argmin()), int(flat.argmax())])
    print(f"upstream gradient d/d batch (incl. range terms): engine vs oracle rel-to-max {rel(gb_eng_full, gb_ref):.3e}")
    m = torch.ones(flat.numel(), dtype=torch.bool)
    m[int(ir[0])] = m[int(ir[1])] = False
    print(f"   direct term only (extremes masked): {float((gb_eng_full.reshape(-1)[m] - gb_ref.reshape(-1)[m]).abs().max() / gb_ref.reshape(-1)[m].abs().max()):.3e}"
          f"   range terms engine dR={dR:.4e} dMin={dMin:.4e}  oracle at those elements {float(gb_ref.reshape(-1)[int(ir[1])]):.4e} {float(gb_ref.reshape(-1)[int(ir[0])]):.4e}")
    print(f"(0) engine d/d image vs oracle:                                   {rel(gi_eng, gi_ref):.3e}")
    em_ref = emulate(gb_ref, jitter)
    print(f"(a) algorithm: CPU emulation fed the ORACLE's upstream vs oracle:  {rel(em_ref, gi_ref):.3e}")
    em_eng = emulate(gb_eng_full, jitter)
    print(f"(b) kernel:    engine d/d image vs CPU emulation fed the ENGINE's upstream: {rel(gi_eng, em_eng):.3e}")
    print(f"(c) upstream:  CPU emulation fed the engine's upstream vs oracle:  {rel(em_eng, gi_ref):.3e}")
    # per-cutout contribution of the upstream difference
    for n in range(cutn):
        d = torch.zeros_like(gb_ref)
        d[n] = gb_eng_full[n] - gb_ref[n]
        c = emulate(d, jitter)
        print(f"    cutout {n} (code {int(J[n, 0]) if jitter else 0}): upstream diff rel {float(d.abs().max() / gb_ref.abs().max()):.2e} -> d/d image {float(c.abs().max() / gi_ref.abs().max()):.3e}")
