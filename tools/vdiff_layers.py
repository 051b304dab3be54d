"""Layer-by-layer parity of the vdiff U-Net: activations and gradients of every element's output, engine (NHWC fp16
taps "vda.<key>" / "vdg.<key>") vs the oracle (autograd with retain_grad).  Shows where an error enters."""
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
it, S = 6, 4096.0
torch.manual_seed(0)
model = R.VDiffCC12M1().eval().requires_grad_(False)
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
model.taps = []
ref = R.iterate(lambda zz: R.vdiff_synth(model, zz, t, ce, alphas[it], sigmas[it])[0], x, [clip], [prompts], torch.from_numpy(T), cs,
                "reflection", 0.4, facs, noise)
taps = model.taps
eng = E.B200Engine(drawer=E.DRAWER_VDIFF, image_hw=(hw, hw), cutn=cutn, clip=[SMALL_CLIP], noise_fac=0.1, seed=3)
eng.load_module(E.MOD_VQGAN, model.ref_state_dict())
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
print(f"z.grad rel err {(zg - ref['z_grad']).abs().max() / ref['z_grad'].abs().max():.3e}")
print(f"{'element':44s} {'shape':>18s}  act_rel_err  grad_rel_err (max-abs / max-abs), grad rel-L2")
for key, tns in taps:
    n, c, h, w = tns.shape
    try:
        a = eng.debug_read("vda." + key, (h * w, c), dtype=torch.float16).float().cpu().T.reshape(1, c, h, w)
        gg = eng.debug_read("vdg." + key, (h * w, c), dtype=torch.float16).float().cpu().T.reshape(1, c, h, w) / S
    except Exception as e:  # the last block has no fp16 output tap
        print(f"{key:44s} {str(tuple(tns.shape)):>18s}  (no tap: {str(e)[:40]})")
        continue
    ra, rg = tns.detach(), tns.grad
    ea = (a - ra).abs().max() / ra.abs().max()
    eg = (gg - rg).abs().max() / rg.abs().max()
    l2 = (gg - rg).norm() / rg.norm()
    print(f"{key:44s} {str(tuple(tns.shape)):>18s}  {ea:.3e}    {eg:.3e}   {l2:.3e}")
