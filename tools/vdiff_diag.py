"""Diagnostics for the vdiff drawer backward: z.grad error vs the oracle for several gradient scales (fp16 range check)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref_path as R  # noqa: E402
from pixray_b200 import engine as E  # noqa: E402
from pixray_b200 import util as U  # noqa: E402
from test_pipeline_gpu import SMALL_CLIP, plant_extremes, random_transforms  # noqa: E402

hw = int(sys.argv[1]) if len(sys.argv) > 1 else 64
it = 6
torch.manual_seed(0)
model = R.VDiffCC12M1().eval().requires_grad_(False)
sd = model.ref_state_dict()
clip = R.init_clip_weights(R.ClipVisual(224, SMALL_CLIP["patch"], SMALL_CLIP["width"], SMALL_CLIP["layers"], SMALL_CLIP["heads"], SMALL_CLIP["out_dim"]), 4)
g = torch.Generator().manual_seed(5)
prompts = [(torch.randn(1, SMALL_CLIP["out_dim"], generator=g), w, float("-inf")) for w in (1.0, -0.3)]
steps, alphas, sigmas = (torch.from_numpy(a) for a in U.vdiff_schedule(20))
ce = torch.randn(1, 512, generator=g)
x = torch.randn(1, 3, hw, hw, generator=g) * float(sigmas[it]) + 0.3 * torch.rand(1, 3, hw, hw, generator=g)
t = steps[it:it + 1]
cutn, cs = 8, 224
T = random_transforms(cutn, cs, 5)
facs, noise = plant_extremes(torch.rand(cutn, generator=g) * 0.1, torch.randn(cutn, 3, cs, cs, generator=g))
ref = R.iterate(lambda zz: R.vdiff_synth(model, zz, t, ce, alphas[it], sigmas[it])[0], x, [clip], [prompts], torch.from_numpy(T), cs,
                "reflection", 0.4, facs, noise)
rz = ref["z_grad"]
print(f"oracle z.grad max {rz.abs().max():.3e} norm {rz.norm():.3e}", flush=True)
for S in (4096.0, 256.0, 32768.0):
    eng = E.B200Engine(drawer=E.DRAWER_VDIFF, image_hw=(hw, hw), cutn=cutn, clip=[SMALL_CLIP], noise_fac=0.1, seed=3, grad_scale=S)
    eng.load_module(E.MOD_VQGAN, sd)
    eng.load_module(E.MOD_CLIP0, clip.state_dict())
    eng.finalize()
    eng.set_prompts(0, torch.cat([p[0] for p in prompts]).numpy(), [p[1] for p in prompts], [p[2] for p in prompts])
    eng.vdiff_set_schedule(steps.numpy(), alphas.numpy(), sigmas.numpy())
    eng.vdiff_set_clip_embed(ce.numpy())
    eng.vdiff_set_iteration(it)
    eng.synth(x)
    eng.make_cutouts(None, transforms=T, zoom_padding=E.PAD_REFLECTION, fill=0.4, noise_facs=facs.numpy(), noise=noise)
    eng.encode_image(0)
    zg = eng.backward().cpu()
    d = zg - rz
    gi = eng.debug_read("g_img", (1, 3, hw, hw)).cpu() / S
    print(f"S={S:8.0f}: z.grad max_abs_err {d.abs().max():.3e} ({d.abs().max() / rz.abs().max():.3e} of max), rel-L2 {d.norm() / rz.norm():.3e}, "
          f"finite {torch.isfinite(zg).all().item()}, d/d image err {(gi - ref['image_grad']).abs().max():.3e}", flush=True)
    eng.close()
    del eng
    torch.cuda.empty_cache()
