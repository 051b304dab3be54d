"""Per-rank compute time of the cutout-sharded config 2 on ONE GPU: for N in (1, 2, 4, 8) run the engine with cutn = 64 / N
(the drawer is replicated, so a rank of an N-GPU job does exactly this much work, minus the NCCL calls).  Prints ms per
iteration and the implied ceiling of the scaling curve -- what the kernels allow before any communication cost.

    python tools/time_rank_share.py [iters]
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixray_b200 import engine as E  # noqa: E402
from pixray_b200 import synthetic as S  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 30
vq_sd = S.vqgan_state_dict(E.VQGAN_F16_16384, 0)
clip_sd = S.clip_state_dict(E.CLIP_ARCH["ViT-B/16"], 1)
prompts = S.prompts(512, (1.0, 0.1), 2)
base = None
for n in (1, 2, 4, 8):
    z = S.z0_vqgan(vq_sd["quantize.embedding.weight"], (16, 16), 3).cuda()
    eng = E.B200Engine(drawer=E.DRAWER_VQGAN, image_hw=(256, 256), cutn=64 // n, clip=[E.CLIP_ARCH["ViT-B/16"]], seed=0)
    eng.load_module(E.MOD_VQGAN, vq_sd)
    eng.load_module(E.MOD_CLIP0, clip_sd)
    eng.finalize()
    eng.set_prompts(0, torch.cat([p[0] for p in prompts]).numpy(), [p[1] for p in prompts], [p[2] for p in prompts])
    for it in range(5):
        eng.iterate(z, 0.2, it)
    eng.sync()
    t0 = time.time()
    for it in range(5, 5 + iters):
        eng.iterate(z, 0.2, it)
    eng.sync()
    ms = (time.time() - t0) / iters * 1e3
    base = base or ms
    print(f"N={n}: cutn/rank {64 // n:2d}  {ms:6.3f} ms/iter  -> {1e3 / ms:6.1f} it/s ceiling, efficiency ceiling {base / ms / n:.2f}")
    del eng
