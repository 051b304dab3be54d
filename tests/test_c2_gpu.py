"""BASELINE config 2 (vqgan imagenet_f16_16384 256x256, ViT-B/16, cutn=64) on the GPU: the engine runs the full-size
path with seeded synthetic weights; checks finiteness / determinism-level properties and prints a first timing.
The full-size z.grad parity against the CPU oracle lives in test_c2_parity_gpu (slow: the oracle needs ~1 min)."""
import time

import numpy as np
import pytest
import torch

from oracle import ref_path as R
from pixray_b200 import engine as E

pytestmark = pytest.mark.gpu


def build_c2(cutn=64, seed=0):
    vq = R.init_vqgan_weights(R.VQModel(), seed)
    clip = R.init_clip_weights(R.ClipVisual(224, 16, 768, 12, 12, 512), seed + 1)
    eng = E.B200Engine(drawer=E.DRAWER_VQGAN, image_hw=(256, 256), cutn=cutn, clip=[E.CLIP_ARCH["ViT-B/16"]],
                       noise_fac=0.1, seed=seed)
    eng.load_module(E.MOD_VQGAN, vq.state_dict())
    eng.load_module(E.MOD_CLIP0, clip.state_dict())
    eng.finalize()
    g = torch.Generator().manual_seed(seed + 2)
    prompts = [(torch.randn(1, 512, generator=g), 1.0, float("-inf")), (torch.randn(1, 512, generator=g), 0.1, float("-inf"))]
    eng.set_prompts(0, torch.cat([p[0] for p in prompts]).numpy(), [p[1] for p in prompts], [p[2] for p in prompts])
    idx = torch.randint(16384, (256,), generator=g)
    z = vq.quantize.embedding.weight[idx].T.reshape(1, 256, 16, 16).contiguous()
    z = (z + 0.05 * torch.randn(z.shape, generator=g)).contiguous()
    return vq, clip, eng, prompts, z


def test_c2_iterate_runs_and_times():
    vq, clip, eng, prompts, z = build_c2()
    zc = z.clone().cuda()
    losses = np.zeros(2, dtype=np.float32)
    eng.iterate(zc, 0.1, 0, losses_out=losses)
    print("[c2] iter0 losses", losses)
    assert np.isfinite(losses).all()
    first = losses.copy()
    for it in range(1, 4):
        eng.iterate(zc, 0.1, it)
    eng.sync()
    n0 = eng.num_launches()
    t0 = time.time()
    K = 10
    for it in range(4, 4 + K):
        eng.iterate(zc, 0.1, it)
    eng.sync()
    dt = (time.time() - t0) / K
    per_iter = (eng.num_launches() - n0) / K
    print(f"[c2] {dt * 1e3:.2f} ms/iter  ({1 / dt:.1f} it/s)  {per_iter:.0f} kernel launches / iter")
    eng.iterate(zc, 0.1, 4 + K, losses_out=losses)
    print("[c2] later losses", losses)
    assert np.isfinite(losses).all() and torch.isfinite(zc).all()
    # optimisation must make progress on the weight-1 prompt
    assert losses[0] < first[0]


@pytest.mark.slow
def test_c2_parity_gpu():
    """Full-size z.grad parity (the second half of BASELINE.json's metric).  Tolerance: 3e-2 * max|grad|."""
    from test_pipeline_gpu import plant_extremes, random_transforms, report
    vq, clip, eng, prompts, z = build_c2(seed=3)
    T = random_transforms(64, 224, 5)
    g = torch.Generator().manual_seed(13)
    facs, noise = plant_extremes(torch.rand(64, generator=g) * 0.1, torch.randn(64, 3, 224, 224, generator=g))
    import bench
    print("[c2] oracle threads:", bench.pick_threads())
    t0 = time.time()
    ref = R.iterate(lambda zz: R.vqgan_synth(vq, zz), z, [clip], [prompts], torch.from_numpy(T), 224, "reflection", 0.4,
                    facs, noise)
    print(f"[c2] oracle iteration on CPU: {time.time() - t0:.1f} s")
    img = eng.synth(z)
    report("c2 image", img, ref["image"])
    eng.make_cutouts(img, transforms=T, zoom_padding=E.PAD_REFLECTION, fill=0.4, noise_facs=facs.numpy(), noise=noise)
    emb = eng.encode_image(0)
    report("c2 embeds", emb, ref["embeds"][0])
    zg = eng.backward()
    err, mag = report("c2 z.grad", zg, ref["z_grad"])
    assert torch.isfinite(zg).all()
    assert err <= 3e-2 * mag
