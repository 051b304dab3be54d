"""GPU parity of the auxiliary losses (pixray_b200/csrc/kernels_losses.cu behind pxr_add_aux_loss; the reference's
Losses/*.py, pixray.py:1384-1393) against oracle/ref_path.py, which tests/test_oracle_golden.py pins to vectors
produced by the real reference classes.

Isolation: every loss ADDS its gradient into a buffer the rest of the chain also writes, so each case runs the same
inputs twice (without / with the loss) and compares the difference of the buffers with torch autograd of the oracle
loss evaluated on the ENGINE's own tensor (read back) -- that removes the fp16 noise of the networks from the check.
Tolerance: fp32 kernels vs fp32 torch, 1e-4 relative to the largest gradient entry (the difference of two fp32
buffers that also hold the much larger CLIP gradient carries its rounding); loss values 1e-5 relative.
"""
import types

import numpy as np
import pytest
import torch

from oracle import ref_path as R
from pixray_b200 import engine as E
from pixray_b200 import losses as L
from test_pipeline_gpu import build, plant_extremes, random_transforms

pytestmark = pytest.mark.gpu
S = 4096.0
CUTN, CS = 8, 224
PALETTE = [[0.9, 0.1, 0.1], [0.1, 0.8, 0.2], [0.2, 0.2, 0.9], [0.95, 0.95, 0.9], [0.05, 0.05, 0.05]]


def _head(D, seed=5):
    g = torch.Generator().manual_seed(seed)
    return dict(weight=torch.randn(1, D, generator=g) * 0.3, bias=torch.tensor([0.7]))


def _run(eng, z, T, facs, noise):
    eng.synth(z)
    eng.make_cutouts(None, transforms=T, zoom_padding=E.PAD_REFLECTION, fill=0.4, noise_facs=facs.numpy(), noise=noise)
    eng.encode_image(0)
    zg = eng.backward()
    return dict(zg=zg.cpu(), g_img=eng.debug_read("g_img", (1, 3, 32, 32)).cpu() / S,
                g_batch=eng.debug_read("g_batch", (CUTN, 3, CS, CS)).cpu() / S,
                de=eng.debug_read("clip0.de", (CUTN, 64)).cpu() / S, img=eng.debug_read("img", (1, 3, 32, 32)).cpu(),
                batch=eng.debug_read("batch", (CUTN, 3, CS, CS)).cpu(), e=eng.debug_read("clip0.e", (CUTN, 64)).cpu(),
                losses=eng.read_losses().copy())


CASES = {
    # name: (kind, params, weight, which tensor, oracle fn on that tensor)
    "symmetry": (E.LOSS_SYMMETRY, [0.7], 1.5, "img", lambda x: R.symmetry_loss(x, 0.7)),
    "edge": (E.LOSS_EDGE, [0.1, 0.05, 3, 6, 4, 0, 0.2, 0.6, 0.9], 2.0, "img",
             lambda x: R.edge_loss(x, [0.2, 0.6, 0.9], (3, 6, 4, 0), 0.1, 0.05)),
    "gaussian": (E.LOSS_GAUSSIAN, [0.6, 6.0, 9.0, 255, 128, 0], 1.0, "img",
                 lambda x: R.gaussian_loss(x, (6.0, 9.0), (255, 128, 0), 0.6)),
    "saturation": (E.LOSS_SATURATION, [1.3], 0.5, "batch", lambda x: R.saturation_loss(x, 1.3)),
    "palette": (E.LOSS_PALETTE, [0.8] + [c for row in PALETTE for c in row], 1.0, "batch",
                lambda x: R.palette_loss(x, PALETTE, 0.8)[0]),
    "smooth_default": (E.LOSS_SMOOTHNESS, [0.9, 0, 1], 1.0, "batch", lambda x: R.smoothness_loss(x, 0.9, "default")),
    "smooth_clipped": (E.LOSS_SMOOTHNESS, [0.9, 1, 1], 1.0, "batch", lambda x: R.smoothness_loss(x, 0.9, "clipped")),
    "smooth_log_sp2": (E.LOSS_SMOOTHNESS, [1.1, 2, 2], 0.7, "batch", lambda x: R.smoothness_loss(x, 1.1, "log", spacing=2)),
}


@pytest.fixture(scope="module")
def setup():
    vq, clip, eng, prompts, z = build(cutn=CUTN, seed=13)
    T = random_transforms(CUTN, CS, 6)
    g = torch.Generator().manual_seed(17)
    facs, noise = plant_extremes(torch.rand(CUTN, generator=g) * 0.1, torch.randn(CUTN, 3, CS, CS, generator=g))
    base = _run(eng, z, T, facs, noise)
    return dict(vq=vq, clip=clip, eng=eng, prompts=prompts, z=z, T=T, facs=facs, noise=noise, base=base)


@pytest.mark.parametrize("name", list(CASES))
def test_aux_loss_value_and_gradient(setup, name):
    kind, params, weight, which, fn = CASES[name]
    eng, base = setup["eng"], setup["base"]
    eng.clear_aux_losses()
    idx = eng.add_aux_loss(kind, weight, params)
    assert idx == 2 and eng.num_losses() == 3  # two prompts, then the auxiliary loss
    cur = _run(eng, setup["z"], setup["T"], setup["facs"], setup["noise"])
    eng.clear_aux_losses()
    x = base[which].clone().requires_grad_(True)
    ref = weight * fn(x)
    ref.backward()
    key = "g_img" if which == "img" else "g_batch"
    delta = cur[key] - base[key]
    err = (delta - x.grad).abs().max().item()
    mag = x.grad.abs().max().item()
    print(f"[aux] {name}: loss engine {cur['losses'][idx]:.7f} oracle {ref.item():.7f}; grad max_abs_err {err:.3e} (max {mag:.3e})")
    assert np.isfinite(cur["losses"]).all()
    assert abs(cur["losses"][idx] - ref.item()) <= 1e-5 * max(1.0, abs(ref.item()))
    assert np.allclose(cur["losses"][:2], base["losses"][:2], atol=0)  # prompt losses untouched
    # the two runs differ by more than the added term: the range-normalise sums use fp32 atomics (run-to-run last-bit
    # noise in the CLIP gradient the term is added to), on top of the fp32 rounding of the accumulation itself
    floor = 1e-5 * base[key].abs().max().item()
    assert mag > 0 and err <= 1e-4 * mag + floor
    if name == "palette":  # integer bookkeeping: nearest-palette index per pixel, bit-exact away from near-ties
        best = eng.debug_read("palette_best", (CUTN * CS * CS,), dtype=torch.int32).cpu().long()
        px = base["batch"].permute(0, 2, 3, 1).reshape(-1, 3)
        d = torch.cdist(torch.tensor(PALETTE), px)
        top2 = d.topk(2, dim=0, largest=False).values
        clear = (top2[1] - top2[0]) > 1e-5
        assert torch.equal(best[clear], d.argmin(0)[clear])
        assert clear.float().mean() > 0.999


def test_aesthetic_head(setup):
    eng, base = setup["eng"], setup["base"]
    head = _head(64)
    loss = L.AestheticLoss(device="cuda")
    args = types.SimpleNamespace(aesthetic_target=10.0, aesthetic_head=head)
    eng.clear_aux_losses()
    session = types.SimpleNamespace(engine=eng)
    loss.attach(session, loss.parse_settings(args), 0.8)
    cur = _run(eng, setup["z"], setup["T"], setup["facs"], setup["noise"])
    val = float(loss.get_loss(None, None, args))
    eng.clear_aux_losses()
    e = base["e"].clone().requires_grad_(True)
    unit = torch.nn.functional.normalize(e, dim=-1)  # CLIP_Base.encode_image's own normalisation (slip.py:66)
    ref = 0.8 * R.aesthetic_loss(unit, head["weight"], head["bias"], 10.0)
    ref.backward()
    delta = cur["de"] - base["de"]
    err, mag = (delta - e.grad).abs().max().item(), e.grad.abs().max().item()
    print(f"[aux] aesthetic: loss engine {val:.6f} oracle {ref.item():.6f}; d/de max_abs_err {err:.3e} (max {mag:.3e})")
    assert abs(val - ref.item()) <= 1e-4 * max(1.0, abs(ref.item()))
    assert err <= 1e-3 * mag  # de also holds the (large) prompt gradient: the difference carries its fp32 rounding


def test_all_aux_losses_through_the_whole_path(setup):
    """ascend_txt with custom losses (pixray.py:1384-1393): loss vector and z.grad against the oracle's autograd."""
    eng, vq, clip = setup["eng"], setup["vq"], setup["clip"]
    head = _head(64)
    args = types.SimpleNamespace(symmetry_weight=0.7, saturation_weight=1.3, palette=PALETTE, palette_weight=0.8,
                                 smoothness_weight=0.9, smoothness_type="clipped", smoothness_gaussian_kernel=0,
                                 smoothness_spacing=1, smoothness_edge_order=1, edge_thickness=10, edge_margins=None,
                                 edge_color="[0.2+0.6+0.9]", edge_color_weight=0.1, global_color_weight=0.05,
                                 edge_input_image="", edge_mask_image="", gaussian_weight=0.6, gaussian_std=(6.0, 9.0),
                                 gaussian_color=(255, 128, 0), aesthetic_target=10.0, aesthetic_head=head)
    session = types.SimpleNamespace(engine=eng)
    eng.clear_aux_losses()
    spec = [("symmetry", 1.5), ("saturation", 0.5), ("palette", 1.0), ("smoothness", 2.0), ("edge", 2.0),
            ("gaussian", 1.0), ("aesthetic", 0.8)]
    objs = []
    for name, w in spec:
        o = L.loss_class_table[name](device="cuda")
        args = o.parse_settings(args)
        objs.append(o.attach(session, args, w))
    cur = _run(eng, setup["z"], setup["T"], setup["facs"], setup["noise"])
    eng.clear_aux_losses()
    m = R.edge_margins_px(args.edge_margins, 32, 32)
    aux = [(1.5, lambda o, b, e: R.symmetry_loss(o, 0.7)), (0.5, lambda o, b, e: R.saturation_loss(b, 1.3)),
           (1.0, lambda o, b, e: R.palette_loss(b, PALETTE, 0.8)[0]),
           (2.0, lambda o, b, e: R.smoothness_loss(b, 0.9, "clipped")),
           (2.0, lambda o, b, e: R.edge_loss(o, [0.2, 0.6, 0.9], m, 0.1, 0.05)),
           (1.0, lambda o, b, e: R.gaussian_loss(o, (6.0, 9.0), (255, 128, 0), 0.6)),
           (0.8, lambda o, b, e: R.aesthetic_loss(e, head["weight"], head["bias"], 10.0))]
    ref = R.iterate(lambda zz: R.vqgan_synth(vq, zz), setup["z"], [clip], [setup["prompts"]], torch.from_numpy(setup["T"]), CS,
                    "reflection", 0.4, setup["facs"], setup["noise"], aux=aux)
    ref_l = np.array([float(l) for l in ref["losses"]], dtype=np.float32)
    print("[aux] losses engine", cur["losses"], "\n      oracle", ref_l)
    assert cur["losses"].shape == ref_l.shape == (9,)
    assert np.abs(cur["losses"] - ref_l).max() < 5e-3
    err, mag = (cur["zg"] - ref["z_grad"]).abs().max().item(), ref["z_grad"].abs().max().item()
    print(f"[aux] z.grad with all auxiliary losses: max_abs_err {err:.3e} (max {mag:.3e})")
    assert err <= 3e-2 * mag
    for o, (name, _) in zip(objs, spec):  # LossInterface.get_loss returns the engine's value for the iteration
        assert o._index is not None
