"""Run a few full-size (BASELINE config 2) iterations of the engine and nothing else -- the target for ncu:

    ncu --set full --clock-control none --import-source on -s <skip> -c <n> -o gpurun_out/prof python tools/profile_c2.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixray_b200 import engine as E  # noqa: E402
from pixray_b200 import synthetic as S  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
vq_sd = S.vqgan_state_dict(E.VQGAN_F16_16384, 0)
clip_sd = S.clip_state_dict(E.CLIP_ARCH["ViT-B/16"], 1)
prompts = S.prompts(512, (1.0, 0.1), 2)
z = S.z0_vqgan(vq_sd["quantize.embedding.weight"], (16, 16), 3).cuda()
eng = E.B200Engine(drawer=E.DRAWER_VQGAN, image_hw=(256, 256), cutn=64, clip=[E.CLIP_ARCH["ViT-B/16"]], seed=0)
eng.load_module(E.MOD_VQGAN, vq_sd)
eng.load_module(E.MOD_CLIP0, clip_sd)
eng.finalize()
eng.set_prompts(0, torch.cat([p[0] for p in prompts]).numpy(), [p[1] for p in prompts], [p[2] for p in prompts])
for it in range(iters):
    eng.iterate(z, 0.2, it)
eng.sync()
print("done", eng.num_launches(), "launches")
