"""BASELINE configs 3 and 5 at their full shapes on one GPU (C3 = vqgan 512x512, ViT-B/16 + ViT-B/32,
cutn 128; C5 = fft 512x512, ViT-L/14, cutn 256).  The CPU oracle cannot run 128 / 256 cutouts through a ViT in test
time, so parity is taken where it is size-independent:

  * drawer forward at full size against the oracle (the image);
  * all cutouts against the oracle (cheap gather);
  * embeddings of a SUBSET of cutouts -- an embedding depends on its own cutout and the batch-global range only
    (slip.py:21-36), which the oracle takes from the full batch;
  * d loss / d cutout for the same subset -- the Prompt loss is a mean over cutouts (pixray.py:275-280), so a row's
    direct gradient depends on that row alone;
  * drawer backward at full size: the oracle's autograd fed the ENGINE's d loss / d image gives z.grad.
"""
import numpy as np
import pytest
import torch
from torch.nn import functional as F

from oracle import ref_path as R
from pixray_b200 import engine as E
from test_pipeline_gpu import plant_extremes, random_transforms, report

pytestmark = [pytest.mark.gpu, pytest.mark.slow]

S = 4096.0


def subset_oracle(clip, batch, rows, prompts, cutn):
    """Embeddings (unit) and d loss / d batch[rows] for `rows` of `batch`, range taken from the whole batch."""
    mn = batch.min()
    rng = (batch - mn).max()
    b = batch[rows].clone().requires_grad_(True)
    x = (b - mn) / rng
    mean = torch.tensor(R.CLIP_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(R.CLIP_STD).view(1, 3, 1, 1)
    e = clip.encode_image((x - mean) / std).float()
    e = e / e.norm(dim=-1, keepdim=True)
    loss = 0.0
    for (embed, weight, stop) in prompts:  # Prompt.forward with the mean taken over all `cutn` rows
        w = torch.as_tensor(float(weight))
        d = F.normalize(e.unsqueeze(1), dim=2).sub(F.normalize(embed.unsqueeze(0), dim=2)).norm(dim=2).div(2).arcsin().pow(2).mul(2)
        d = d * w.sign()
        loss = loss + w.abs() * R.replace_grad(d, torch.maximum(d, torch.as_tensor(float(stop)))).sum() / cutn
    loss.backward()
    return e.detach(), b.grad.detach()


def check_subset(eng, clips, prompt_sets, batch_ref, rows, cutn, cs):
    for i, (clip, pms) in enumerate(zip(clips, prompt_sets)):
        e_eng = eng.encode_image(i).cpu()
        eng.prompt_loss(i)
    eng.backward()
    gb_eng = eng.debug_read("g_batch", (cutn, 3, cs, cs)).cpu() / S
    gb_ref = torch.zeros(len(rows), 3, cs, cs)
    for i, (clip, pms) in enumerate(zip(clips, prompt_sets)):
        e_ref, g = subset_oracle(clip, batch_ref, rows, pms, cutn)
        gb_ref += g
        e_eng = eng.debug_read(f"clip{i}.e", (cutn, e_ref.shape[1])).cpu()
        e_eng = e_eng / e_eng.norm(dim=-1, keepdim=True)
        err, _ = report(f"perceptor {i} embeddings of cutouts {rows}", e_eng[rows], e_ref)
        assert err < 5e-3
    err, mag = report(f"d loss / d cutout, rows {rows}", gb_eng[rows], gb_ref)
    assert err <= 3e-2 * mag


def test_config3_shape_vqgan512_two_perceptors_cutn128():
    cutn, cs, seed, H = 128, 224, 0, 512
    vq = R.init_vqgan_weights(R.VQModel(), seed)
    clip_a = R.init_clip_weights(R.ClipVisual(224, 16, 768, 12, 12, 512), seed + 1)
    clip_b = R.init_clip_weights(R.ClipVisual(224, 32, 768, 12, 12, 512), seed + 2)
    eng = E.B200Engine(drawer=E.DRAWER_VQGAN, image_hw=(H, H), cutn=cutn, clip=[E.CLIP_ARCH["ViT-B/16"], E.CLIP_ARCH["ViT-B/32"]],
                       noise_fac=0.1, seed=seed)
    eng.load_module(E.MOD_VQGAN, vq.state_dict())
    eng.load_module(E.MOD_CLIP0, clip_a.state_dict())
    eng.load_module(E.MOD_CLIP1, clip_b.state_dict())
    eng.finalize()
    g = torch.Generator().manual_seed(seed + 3)
    pa = [(torch.randn(1, 512, generator=g), 1.0, float("-inf")), (torch.randn(1, 512, generator=g), 0.1, float("-inf"))]
    pb = [(torch.randn(1, 512, generator=g), 1.0, float("-inf")), (torch.randn(1, 512, generator=g), -0.2, float("-inf"))]
    for i, pm in enumerate((pa, pb)):
        eng.set_prompts(i, torch.cat([p[0] for p in pm]).numpy(), [p[1] for p in pm], [p[2] for p in pm])
    idx = torch.randint(16384, (32 * 32,), generator=g)
    z = vq.quantize.embedding.weight[idx].T.reshape(1, 256, 32, 32).contiguous()
    z = (z + 0.05 * torch.randn(z.shape, generator=g)).contiguous()
    T = random_transforms(cutn, cs, 5)
    facs, noise = plant_extremes(torch.rand(cutn, generator=g) * 0.1, torch.randn(cutn, 3, cs, cs, generator=g))

    zr = z.clone().requires_grad_(True)
    img_ref = R.vqgan_synth(vq, zr)
    img = eng.synth(z)
    e_img, _ = report("c3 image 512x512", img, img_ref.detach())
    assert e_img < 5e-3
    idx_eng = eng.debug_read("vq_idx", (1024,), dtype=torch.int32).cpu().long()
    assert torch.equal(idx_eng, R.vector_quantize(z.movedim(1, 3), vq.quantize.embedding.weight)[1].reshape(-1))
    batch_ref = R.make_cutouts(img_ref.detach(), torch.from_numpy(T), cs, "reflection", 0.4, facs, noise)
    batch = eng.make_cutouts(img_ref.detach(), transforms=T, zoom_padding=E.PAD_REFLECTION, fill=0.4, noise_facs=facs.numpy(),
                             noise=noise)
    e_b, _ = report("c3 cutouts (oracle image in)", batch, batch_ref)
    assert e_b < 1e-4
    ir = eng.debug_read("irange", (4,), dtype=torch.int32).cpu()
    assert ir[0].item() == batch_ref.reshape(-1).argmin().item() and ir[1].item() == batch_ref.reshape(-1).argmax().item()
    check_subset(eng, [clip_a, clip_b], [pa, pb], batch_ref, [2, 40, 90, 127], cutn, cs)
    # drawer backward at 512x512: the oracle's autograd fed the engine's own d loss / d image
    g_img = eng.debug_read("g_img", (1, 3, H, H)).cpu() / S
    zg = eng.debug_read("z_grad", z.shape).cpu()
    img_ref.backward(g_img)
    e_g, m_g = report("c3 z.grad (decoder backward of the engine's image gradient)", zg, zr.grad)
    assert torch.isfinite(zg).all() and e_g <= 3e-2 * m_g
    # and the fused iteration runs at this shape
    zc = z.clone().cuda()
    losses = np.zeros(4, dtype=np.float32)
    for it in range(3):
        eng.iterate(zc, 0.1, it, losses_out=losses)
        assert np.isfinite(losses).all(), losses
    assert torch.isfinite(zc).all()
    print("[c3] losses after 3 iterations", losses)


def test_config5_shape_fft512_vit_l14_cutn256():
    cutn, cs, seed, H = 256, 224, 0, 512
    arch = E.CLIP_ARCH["ViT-L/14"]
    clip = R.init_clip_weights(R.ClipVisual(224, 14, 1024, 24, 16, 768), seed + 1)
    eng = E.B200Engine(drawer=E.DRAWER_FFT, image_hw=(H, H), cutn=cutn, clip=[arch], noise_fac=0.1, seed=seed)
    eng.load_module(E.MOD_CLIP0, clip.state_dict())
    eng.finalize()
    g = torch.Generator().manual_seed(seed + 2)
    pm = [(torch.randn(1, 768, generator=g), 1.0, float("-inf")), (torch.randn(1, 768, generator=g), 0.1, float("-inf"))]
    eng.set_prompts(0, torch.cat([p[0] for p in pm]).numpy(), [p[1] for p in pm], [p[2] for p in pm])
    z = (0.01 * torch.randn(1, 3, H, H // 2 + 1, 2, generator=g)).contiguous()
    T = random_transforms(cutn, cs, 6)
    facs, noise = plant_extremes(torch.rand(cutn, generator=g) * 0.1, torch.randn(cutn, 3, cs, cs, generator=g))
    zr = z.clone().requires_grad_(True)
    img_ref = R.fft_synth(zr)
    img = eng.synth(z)
    e_img, _ = report("c5 fft image 512x512", img, img_ref.detach())
    assert e_img < 2e-4
    batch_ref = R.make_cutouts(img_ref.detach(), torch.from_numpy(T), cs, "border", 0.6, facs, noise)
    batch = eng.make_cutouts(img_ref.detach(), transforms=T, zoom_padding=E.PAD_BORDER, fill=0.6, noise_facs=facs.numpy(),
                             noise=noise)
    e_b, _ = report("c5 cutouts", batch, batch_ref)
    assert e_b < 1e-4
    check_subset(eng, [clip], [pm], batch_ref, [3, 200], cutn, cs)
    g_img = eng.debug_read("g_img", (1, 3, H, H)).cpu() / S
    zg = eng.debug_read("z_grad", z.shape).cpu()
    img_ref.backward(g_img)
    e_g, m_g = report("c5 spectrum grad (drawer backward of the engine's image gradient)", zg, zr.grad)
    assert torch.isfinite(zg).all() and e_g <= 3e-2 * m_g
    zc = z.clone().cuda()
    losses = np.zeros(2, dtype=np.float32)
    for it in range(2):
        eng.iterate(zc, 0.3, it, losses_out=losses)
        assert np.isfinite(losses).all(), losses
    assert torch.isfinite(zc).all()
    print("[c5] losses after 2 iterations", losses)
