"""Per-op CUDA-event profile of one full-size (BASELINE config 2) iteration: sets PXR_PROFILE_DUMP so that
pxr_profile_iteration writes one CSV row per op, then prints the rows grouped by label (sum of 3 profiled iterations / 3).

    python tools/profile_ops.py [out.csv] [cutn]

cutn = 8 reproduces, on ONE GPU, the per-rank shapes of the 8-way cutout-sharded run (M = 8 x 197 tokens per GEMM) without
the collectives: where the small-M regime loses time.
"""
import collections
import csv
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ops.csv"
cutn = int(sys.argv[2]) if len(sys.argv) > 2 else 64
os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
os.environ["PXR_PROFILE_DUMP"] = out + ".tmp"
from pixray_b200 import engine as E  # noqa: E402
from pixray_b200 import synthetic as S  # noqa: E402

vq_sd = S.vqgan_state_dict(E.VQGAN_F16_16384, 0)
clip_sd = S.clip_state_dict(E.CLIP_ARCH["ViT-B/16"], 1)
prompts = S.prompts(512, (1.0, 0.1), 2)
z = S.z0_vqgan(vq_sd["quantize.embedding.weight"], (16, 16), 3).cuda()
eng = E.B200Engine(drawer=E.DRAWER_VQGAN, image_hw=(256, 256), cutn=cutn, clip=[E.CLIP_ARCH["ViT-B/16"]], seed=0)
eng.load_module(E.MOD_VQGAN, vq_sd)
eng.load_module(E.MOD_CLIP0, clip_sd)
eng.finalize()
eng.set_prompts(0, torch.cat([p[0] for p in prompts]).numpy(), [p[1] for p in prompts], [p[2] for p in prompts])
for it in range(5):
    eng.iterate(z, 0.2, it)
eng.sync()
agg = collections.OrderedDict()
R = 3
for r in range(R):
    prof = eng.profile_iteration(z, 0.2, 5 + r)
    for row in csv.DictReader(open(out + ".tmp")):
        a = agg.setdefault(row["op"], [0, 0.0, 0.0])
        a[0] += 1
        a[1] += float(row["ms"])
        a[2] += float(row["gflop"])
os.remove(out + ".tmp")
with open(out, "w") as f:
    f.write("op,count_per_iter,us_per_iter,us_each,tflops\n")
    for k, (n, ms, gf) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k},{n / R:.0f},{ms / R * 1e3:.1f},{ms / n * 1e3:.1f},{gf / ms if gf else 0:.0f}\n")
print(open(out).read())
print(prof)
