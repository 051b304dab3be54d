"""Image prompts (pixray.py:1308-1336): per iteration the target image is cut with the cached transforms, encoded, and scored
as a multi-row Prompt(embed [cutn, D]).  Engine (pxr_set_image_prompts + the stage-wise C-ABI calls) vs the oracle."""
import numpy as np
import pytest
import torch

from oracle import ref_path as R
from pixray_b200 import cutouts
from pixray_b200 import engine as E

pytestmark = pytest.mark.gpu


def test_image_prompt_losses_and_gradient():
    from test_pipeline_gpu import build, plant_extremes, random_transforms, report
    cutn, cs = 8, 224
    vq, clip, eng, prompts, z = build(cutn=cutn, seed=7)
    T = random_transforms(cutn, cs, 17)
    J = cutouts.sample_color_jitter(cutn, 23)
    g = torch.Generator().manual_seed(29)
    facs, noise = plant_extremes(torch.rand(cutn, generator=g) * 0.1, torch.randn(cutn, 3, cs, cs, generator=g))
    # the second target has its OWN size (resize_image keeps the source aspect, pixray.py:514-518; MakeCutouts pools any
    # size to cut_size x cut_size, pixray.py:463), larger than the canvas in one direction: ragged pooling windows
    targets = [torch.rand(1, 3, 32, 32, generator=g), torch.rand(1, 3, 300, 100, generator=g) * 0.5 + 0.25]
    weights = [0.8, -0.4]
    tp = list(zip(targets, weights))
    synth = lambda zz: R.vqgan_synth(vq, zz)  # noqa: E731
    ref_text_only = R.iterate(synth, z, [clip], [prompts], torch.from_numpy(T), cs, "reflection", 0.45, facs, noise)
    eng.set_image_prompts(targets, weights)
    assert eng.num_losses() == len(prompts) + 2
    losses = None
    # with the main pass jittered (the targets' cutouts are not, pixray.py:480-486): loss vector; without: z.grad at the
    # bound of the smooth path (tests/test_color_jitter.py explains the looser whole-chain bound with the stage on)
    for jit in (J, None):
        ref = R.iterate(synth, z, [clip], [prompts], torch.from_numpy(T), cs, "reflection", 0.45, facs, noise,
                        jitter=None if jit is None else torch.from_numpy(jit), image_prompts=tp)
        eng.synth(z)
        eng.make_cutouts(None, transforms=T, zoom_padding=E.PAD_REFLECTION, fill=0.45, noise_facs=facs.numpy(),
                         noise=noise, color_jitter=jit)
        eng.encode_image(0)
        losses = eng.prompt_loss(0)
        ref_losses = torch.stack([l.reshape(()) for l in ref["losses"]])
        assert losses.numel() == ref_losses.numel() == len(prompts) + 2
        e_l, _ = report("losses (text + image prompts)" + (" jittered" if jit is not None else ""), losses, ref_losses)
        assert e_l < 2e-3
        zg = eng.backward()
        e_g, m_g = report("z.grad with image prompts" + (" jittered" if jit is not None else ""), zg, ref["z_grad"])
        if jit is None:
            assert (ref["z_grad"] - ref_text_only["z_grad"]).abs().max() > 0.05 * ref["z_grad"].abs().max()
            assert e_g <= 3e-2 * m_g
        else:
            assert e_g <= 4e-2 * m_g   # measured 2.85e-2; bound explained in tests/test_color_jitter.py (b)

    # the fused iteration takes the same path: loss vector of pxr_iterate == the stage-wise one
    out = np.zeros(eng.num_losses(), dtype=np.float32)
    zz = z.clone().cuda()
    eng.iterate(zz, 0.1, 0, params=dict(transforms=T, zoom_padding=E.PAD_REFLECTION, fill=0.45, noise_facs=facs.numpy(),
                                        noise=noise), losses_out=out)
    assert np.abs(out - losses.cpu().numpy()).max() < 1e-5
    # clearing restores the text-only loss vector
    eng.set_image_prompts(None)
    assert eng.num_losses() == len(prompts)
