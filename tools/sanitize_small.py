"""Small-config pxr_iterate + per-op path for compute-sanitizer (memcheck / racecheck / initcheck):

    PXR_GN_COOP=0 compute-sanitizer --tool memcheck python tools/sanitize_small.py

PXR_GN_COOP=0: the single-kernel GroupNorm synchronises its blocks through a device counter and needs them co-resident,
which the sanitizer's serialised execution does not guarantee; the three-kernel variant computes the same thing."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from pixray_b200 import cutouts, synthetic  # noqa: E402
from pixray_b200 import engine as E  # noqa: E402

vq_cfg = dict(z_channels=128, n_embed=1024, ch=128, ch_mult=(1, 2), num_res_blocks=1, attn_resolution=16, resolution=32)
clip_cfg = dict(width=128, layers=2, heads=2, patch=16, image_res=224, out_dim=64)
cutn = 8
eng = E.B200Engine(drawer=E.DRAWER_VQGAN, image_hw=(32, 32), vqgan=vq_cfg, cutn=cutn, clip=[clip_cfg], seed=7)
vq_sd = synthetic.vqgan_state_dict(vq_cfg, 7)
eng.load_module(E.MOD_VQGAN, vq_sd)
eng.load_module(E.MOD_CLIP0, synthetic.clip_state_dict(clip_cfg, 8))
eng.finalize()
pr = synthetic.prompts(64, (1.0, -0.3), 9)
eng.set_prompts(0, torch.cat([p[0] for p in pr]).numpy(), [p[1] for p in pr], [p[2] for p in pr])
z = synthetic.z0_vqgan(vq_sd["quantize.embedding.weight"], (16, 16), 3)
T = cutouts.sample_transforms(cutn, 224, 4)
g = torch.Generator().manual_seed(5)
facs = (torch.rand(cutn, generator=g) * 0.1).numpy()
noise = torch.randn(cutn, 3, 224, 224, generator=g)
zc = z.clone().cuda()
torch.cuda.synchronize()
losses = np.zeros(2, dtype=np.float32)
eng.iterate(zc, 0.05, 0, params=dict(transforms=T, zoom_padding=0, fill=0.5, noise_facs=facs, noise=noise), losses_out=losses)
img_it = eng.debug_read("img", (1, 3, 32, 32)).cpu()
g_it = eng.debug_read("z_grad", z.shape).cpu()
img_op = eng.synth(z).cpu()
eng.make_cutouts(None, transforms=T, zoom_padding=0, fill=0.5, noise_facs=facs, noise=noise, it=0)
eng.encode_image(0)
g_op = eng.backward().cpu()
img_op2 = eng.synth(z).cpu()
print("per-op again vs per-op", (img_op2 - img_op).abs().max().item())
zc2 = z.clone().cuda()
torch.cuda.synchronize()
eng.iterate(zc2, 0.05, 1, params=dict(transforms=T, zoom_padding=1, fill=0.5, noise_facs=facs, noise=noise), losses_out=losses)
img_it2 = eng.debug_read("img", (1, 3, 32, 32)).cpu()
print("second iterate image vs first", (img_it2 - img_it).abs().max().item())
zb = eng.debug_read("z", z.shape).cpu()
print("engine z after iterate vs z (Adam moved it by ~lr):", (zb - z).abs().max().item())
print("losses", losses)
print("image iterate vs per-op max diff", (img_it - img_op).abs().max().item())
print("z.grad iterate vs per-op max diff", (g_it - g_op).abs().max().item(), "max", g_op.abs().max().item())
