"""Spot prompts (args.spot_prompts / args.spot_prompts_off: pixray.py:917-931; the mask of fetch_spot_indexes 370-394 zeroes
part of the pooled image, 453-466; one more cutout batch per kind on the iteration's cached transforms, 1262-1293).  The
fused iteration against the oracle: loss vector in the reference's order and z.grad with gradient through all passes."""
import numpy as np
import pytest
import torch

from oracle import ref_path as R
from pixray_b200 import cutouts
from pixray_b200 import engine as E

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("with_off,hw", [(True, (32, 32)), (False, (32, 32)), (True, (32, 48))])
def test_spot_prompts_match_the_oracle(with_off, hw):
    from test_pipeline_gpu import SMALL_CLIP, SMALL_VQ, plant_extremes, report
    H, W = hw
    aspect = W / H
    cutn, cs, seed = 8, 224, 5
    torch.manual_seed(seed)
    vq = R.init_vqgan_weights(R.VQModel(n_embed=1024, embed_dim=128, ch=128, ch_mult=(1, 2), num_res_blocks=1,
                                        attn_resolutions=(16,), resolution=32, z_channels=128), seed)
    clip = R.init_clip_weights(R.ClipVisual(224, 32, 128, 2, 2, 64), seed + 1)
    eng = E.B200Engine(drawer=E.DRAWER_VQGAN, image_hw=(H, W), vqgan=SMALL_VQ, cutn=cutn, clip=[SMALL_CLIP], noise_fac=0.1,
                       seed=seed, cut_aspect=aspect)
    eng.load_module(E.MOD_VQGAN, vq.state_dict())
    eng.load_module(E.MOD_CLIP0, clip.state_dict())
    eng.finalize()
    g = torch.Generator().manual_seed(seed + 2)
    rnd = lambda: torch.randn(1, 64, generator=g)  # noqa: E731
    prompts = [(rnd(), 1.0, float("-inf")), (rnd(), -0.3, float("-inf"))]
    spot_on = [(rnd(), 0.8, float("-inf")), (rnd(), 0.4, -0.1)]
    spot_off = [(rnd(), 0.6, float("-inf"))] if with_off else []
    eng.set_prompts(0, torch.cat([p[0] for p in prompts]).numpy(), [p[1] for p in prompts], [p[2] for p in prompts])
    # an off-centre disc, like inputs/spot_square.png: True inside
    yy, xx = torch.meshgrid(torch.arange(cs), torch.arange(cs), indexing="ij")
    mask = (((yy - 100) ** 2 + (xx - 130) ** 2) < 60 ** 2)[None].expand(3, -1, -1).contiguous()
    eng.set_spot_mask(mask)
    eng.set_spot_prompts(0, 1, torch.cat([p[0] for p in spot_on]).numpy(), [p[1] for p in spot_on], [p[2] for p in spot_on])
    if with_off:
        eng.set_spot_prompts(0, 0, torch.cat([p[0] for p in spot_off]).numpy(), [p[1] for p in spot_off], [p[2] for p in spot_off])
    n_loss = len(spot_on) + len(spot_off) + len(prompts)
    assert eng.num_losses() == n_loss
    h, w = H // 2, W // 2
    idx = torch.randint(1024, (h * w,), generator=g)
    z = (vq.quantize.embedding.weight[idx].T.reshape(1, 128, h, w) + 0.05 * torch.randn(1, 128, h, w, generator=g)).contiguous()
    T = cutouts.sample_transforms(cutn, cs, 21, aspect=aspect)
    facs, noise = plant_extremes(torch.rand(cutn, generator=g) * 0.1, torch.randn(cutn, 3, cs, cs, generator=g))
    synth = lambda zz: R.vqgan_synth(vq, zz)  # noqa: E731
    ref = R.iterate(synth, z, [clip], [prompts], torch.from_numpy(T), cs, "border", 0.35, facs, noise, aspect=aspect,
                    spot_mask=mask, spot_prompts=[spot_on], spot_prompts_off=[spot_off])
    plain = R.iterate(synth, z, [clip], [prompts], torch.from_numpy(T), cs, "border", 0.35, facs, noise, aspect=aspect)
    assert (ref["z_grad"] - plain["z_grad"]).abs().max() > 0.05 * ref["z_grad"].abs().max()  # the spot passes matter
    zc = z.clone().cuda()
    losses = np.zeros(n_loss, dtype=np.float32)
    eng.iterate(zc, 0.05, 1, params=dict(transforms=T, zoom_padding=E.PAD_BORDER, fill=0.35, noise_facs=facs.numpy(),
                                         noise=noise), losses_out=losses)
    ref_l = np.array([float(l) for l in ref["losses"]], dtype=np.float32)
    print("[parity] spot losses engine", losses, "oracle", ref_l)
    assert len(ref_l) == n_loss and np.abs(losses - ref_l).max() < 5e-3
    e_g, m_g = report("z.grad with spot prompts", eng.debug_read("z_grad", z.shape), ref["z_grad"])
    assert e_g <= 3e-2 * m_g
    # clearing the spot prompts restores the plain iteration
    eng.set_spot_prompts(0, 1, [], [], [])
    eng.set_spot_prompts(0, 0, [], [], [])
    assert eng.num_losses() == len(prompts)
    zc2 = z.clone().cuda()
    l2 = np.zeros(len(prompts), dtype=np.float32)
    eng.reset_optimizer()
    eng.iterate(zc2, 0.05, 1, params=dict(transforms=T, zoom_padding=E.PAD_BORDER, fill=0.35, noise_facs=facs.numpy(),
                                          noise=noise), losses_out=l2)
    e_p, m_p = report("z.grad after clearing", eng.debug_read("z_grad", z.shape), plain["z_grad"])
    assert e_p <= 3e-2 * m_p
