"""The anchor terms of ascend_txt (pixray.py:1344-1375): init_weight (spherical), init_weight_dist (mse / 2),
init_weight_pix (l1 / 2 on the image), init_weight_cos (cosine embedding), image_labels (spherical to an encoded label)
inside the fused iteration -- loss vector (anchors sit between the prompts and the custom losses) and z.grad vs the oracle."""
import numpy as np
import pytest
import torch

from oracle import ref_path as R
from pixray_b200 import cutouts
from pixray_b200 import engine as E

pytestmark = pytest.mark.gpu

KINDS = {"spherical": E.ANCHOR_SPHERICAL, "mse": E.ANCHOR_MSE, "cos": E.ANCHOR_COS, "pix": E.ANCHOR_PIX}


def _run(eng, z, T, facs, noise, n_loss):
    zc = z.clone().cuda()
    losses = np.zeros(n_loss, dtype=np.float32)
    eng.iterate(zc, 0.05, 0, params=dict(transforms=T, zoom_padding=E.PAD_REFLECTION, fill=0.4, noise_facs=facs.numpy(),
                                         noise=noise), losses_out=losses)
    return losses, eng.debug_read("z_grad", z.shape)


@pytest.mark.parametrize("which", ["spherical", "mse", "cos", "pix", "all", "all + symmetry"])
def test_anchor_terms_match_the_oracle(which):
    from test_pipeline_gpu import build, plant_extremes, report
    cutn, cs = 8, 224
    vq, clip, eng, prompts, z = build(cutn=cutn, seed=4)
    g = torch.Generator().manual_seed(41)
    z_orig = z + 0.3 * torch.randn(z.shape, generator=g)
    init_img = torch.rand(1, 3, 32, 32, generator=g)
    refs = {"spherical": z_orig, "mse": z_orig, "cos": z_orig, "pix": init_img}
    order = ["spherical", "mse", "pix", "cos"] if which.startswith("all") else [which]   # ascend_txt's order
    T = cutouts.sample_transforms(cutn, cs, 43)
    facs, noise = plant_extremes(torch.rand(cutn, generator=g) * 0.1, torch.randn(cutn, 3, cs, cs, generator=g))
    synth = lambda zz: R.vqgan_synth(vq, zz)  # noqa: E731
    aux = [(0.6, lambda out, batch, emb: R.symmetry_loss(out, 1.0))] if which.endswith("symmetry") else []
    plain = R.iterate(synth, z, [clip], [prompts], torch.from_numpy(T), cs, "reflection", 0.4, facs, noise, aux=aux)
    # weights: each term's own gradient (unit weight, autograd) scaled to half of the prompt gradient's maximum, so every
    # term is a visible share of z.grad
    anchors = []
    for k in order:
        zz = z.clone().requires_grad_(True)
        R.anchor_loss(k, 1.0, zz, synth(zz) if k == "pix" else None, refs[k]).backward()
        anchors.append((k, float(0.5 * plain["z_grad"].abs().max() / zz.grad.abs().max()), refs[k]))
    ref = R.iterate(synth, z, [clip], [prompts], torch.from_numpy(T), cs, "reflection", 0.4, facs, noise, anchors=anchors,
                    aux=aux)
    # the terms must matter for the gradient, or the comparison proves nothing
    assert (ref["z_grad"] - plain["z_grad"]).abs().max() > 0.2 * plain["z_grad"].abs().max()
    for (kind, weight, r) in anchors:
        eng.add_anchor(KINDS[kind], weight, r)
    if aux:
        eng.add_aux_loss(E.LOSS_SYMMETRY, 0.6, [1.0])
    n_loss = len(prompts) + len(anchors) + len(aux)
    assert eng.num_losses() == n_loss
    losses, zg = _run(eng, z, T, facs, noise, n_loss)
    ref_l = np.array([float(l) for l in ref["losses"]], dtype=np.float32)
    print(f"[parity] anchors {which}: losses engine {losses} oracle {ref_l}")
    a0, a1 = len(prompts), len(prompts) + len(anchors)
    rest = np.concatenate([losses[:a0] - ref_l[:a0], losses[a1:] - ref_l[a1:]])
    assert np.abs(rest).max() < 5e-3
    for k, got, want in zip(order, losses[a0:a1], ref_l[a0:a1]):
        # latent terms: fp32 kernel vs autograd of the same closed form; pix: l1 against the engine's fp16-decoder image
        assert abs(got - want) <= (2e-2 if k == "pix" else 2e-5) * abs(want), (k, got, want)
    e_g, m_g = report(f"z.grad with anchors {which}", zg, ref["z_grad"])
    assert e_g <= 3e-2 * m_g
    if "pix" not in order:
        # the latent terms never pass the drawer: engine(z.grad with) - engine(z.grad without) is the fp32 kernel against
        # autograd of the same closed forms
        eng.clear_anchors()
        _, zg0 = _run(eng, z, T, facs, noise, len(prompts) + len(aux))
        _, zg0b = _run(eng, z, T, facs, noise, len(prompts) + len(aux))
        # run-to-run difference of the engine's own prompt gradient (the cutout backward scatters with float atomics, and a
        # last-bit difference there can flip fp16 roundings downstream): the subtraction cannot be more exact than that
        jitter = (zg0.cpu() - zg0b.cpu()).abs().max().item()
        print(f"[parity] run-to-run max |z.grad difference| without anchors: {jitter:.3e}")
        e_a, m_a = report(f"anchor gradient alone ({which})", zg.cpu() - zg0.cpu(), ref["z_grad"] - plain["z_grad"])
        # two-state jitter: a run lands on one of two z.grad values 5e-6 apart (6.8e-4 of max|z.grad|), so the difference of
        # two runs is either exact to 1e-9 or off by that step
        assert e_a <= 2e-4 * m_a + max(3 * jitter, 1e-3 * m_g) + 1e-9
    else:
        eng.clear_anchors()
    assert eng.num_losses() == len(prompts) + len(aux)


def test_anchor_at_the_starting_point_is_zero_and_finite():
    """Iteration 0 of a run with init_weight: z == z_orig exactly, so d = 0 (torch's norm backward gives a zero gradient there),
    the mse and cosine terms vanish too."""
    from test_pipeline_gpu import build
    vq, clip, eng, prompts, z = build(cutn=8, seed=4)
    for k in (E.ANCHOR_SPHERICAL, E.ANCHOR_MSE, E.ANCHOR_COS):
        eng.add_anchor(k, 1.0, z)
    zc = z.clone().cuda()
    losses = np.zeros(len(prompts) + 3, dtype=np.float32)
    eng.iterate(zc, 0.05, 0, losses_out=losses)
    assert np.isfinite(losses).all() and np.abs(losses[len(prompts):]).max() < 1e-6
    assert torch.isfinite(eng.debug_read("z_grad", z.shape)).all()
    with pytest.raises(Exception):
        eng.add_anchor(E.ANCHOR_MSE, 1.0, torch.zeros(7))
