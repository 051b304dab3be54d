"""Diagnostic (run by hand on a GPU box: python tests/diag_aspect.py): stage-wise gradients of the non-square canvas path.
The oracle's image goes in, so the decoder's fp16 rounding is out of the picture up to d loss / d image; then the engine's
own chain from z.  Prints, per canvas, the error of d/d batch, d/d stretched source (via d/d pooled), d/d pooled, d/d image
and z.grad against autograd on the oracle."""
import sys

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
sys.path.insert(0, __file__.rsplit("/", 1)[0])

from oracle import ref_path as R  # noqa: E402
from pixray_b200 import cutouts  # noqa: E402
from pixray_b200 import engine as E  # noqa: E402
from test_pipeline_gpu import SMALL_CLIP, SMALL_VQ, plant_extremes  # noqa: E402


def rel(name, got, ref):
    got, ref = got.detach().cpu().double(), ref.detach().cpu().double()
    e = (got - ref).abs().max().item()
    m = ref.abs().max().item()
    l2 = ((got - ref).norm() / max(ref.norm().item(), 1e-30)).item()
    print(f"  {name:34s} max_abs_err {e:.3e}  max {m:.3e}  rel {e / max(m, 1e-30):.3e}  rel-L2 {l2:.3e}")


def run(hw, seed=3):
    H, W = hw
    aspect = W / H
    cutn, cs = 8, 224
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
    prompts = [(torch.randn(1, 64, generator=g), 1.0, float("-inf")), (torch.randn(1, 64, generator=g), -0.3, float("-inf"))]
    eng.set_prompts(0, torch.cat([p[0] for p in prompts]).numpy(), [1.0, -0.3], [float("-inf")] * 2)
    h, w = H // 2, W // 2
    idx = torch.randint(1024, (h * w,), generator=g)
    z = (vq.quantize.embedding.weight[idx].T.reshape(1, 128, h, w) + 0.05 * torch.randn(1, 128, h, w, generator=g)).contiguous()
    T = cutouts.sample_transforms(cutn, cs, 11, aspect=aspect)
    sh, sw = cutouts.source_size(cs, aspect)
    facs, noise = plant_extremes(torch.rand(cutn, generator=g) * 0.1, torch.randn(cutn, 3, cs, cs, generator=g))
    print(f"canvas {H}x{W}  aspect {aspect:.3f}  stretched source {sh}x{sw}")

    # oracle chain with every intermediate kept
    zz = z.clone().requires_grad_(True)
    img_r = R.vqgan_synth(vq, zz)
    img_r.retain_grad()
    pooled = R.pool_avg_max(img_r, cs)
    pooled.retain_grad()
    src = R.rescale_for_aspect(pooled, aspect)
    src.retain_grad()
    srcb = src.expand(cutn, -1, -1, -1)
    nz = int(0.6 * cutn)
    Tt = torch.from_numpy(T)
    parts = [R.warp_perspective(srcb[:nz], Tt[:nz], (cs, cs), padding_mode="reflection"),
             R.warp_perspective(srcb[nz:], Tt[nz:], (cs, cs), padding_mode="fill", fill_value=[0.4, 0.4, 0.4])]
    batch = torch.cat(parts) + facs.reshape(cutn, 1, 1, 1) * noise
    batch.retain_grad()
    emb = R.encode_image(clip, batch).float()
    loss = sum(R.prompt_loss(emb, *p) for p in prompts)
    loss.backward()

    S = 4096.0
    # engine on the ORACLE's image (per-op path)
    eng.synth(z)
    eng.make_cutouts(img_r.detach(), transforms=T, zoom_padding=E.PAD_REFLECTION, fill=0.4, noise_facs=facs.numpy(), noise=noise)
    eng.encode_image(0)
    zg = eng.backward()
    ir = eng.debug_read("irange", (4,), dtype=torch.int32).cpu()
    flat = batch.detach().reshape(-1)
    print("  argmin/argmax element engine", ir[:2].tolist(), "oracle", [flat.argmin().item(), flat.argmax().item()])
    gb = eng.debug_read("g_batch", (cutn, 3, cs, cs)) / S
    ref_gb = batch.grad.clone()
    # the d/dmin, d/dmax terms are added inside cutout_backward: compare away from the two extreme elements
    mask = torch.ones_like(ref_gb, dtype=torch.bool).reshape(-1)
    mask[flat.argmin()] = False
    mask[flat.argmax()] = False
    rel("d/d batch (direct term)", gb.cpu().reshape(-1)[mask], ref_gb.reshape(-1)[mask])
    rel("d/d pooled", eng.debug_read("g_pooled", (1, 3, cs, cs)) / S, pooled.grad)
    rel("d/d image (oracle image in)", eng.debug_read("g_img", (1, 3, H, W)) / S, img_r.grad)
    rel("z.grad (oracle image in cutouts)", zg, zz.grad)
    # engine's own chain
    zc = z.clone().cuda()
    losses = np.zeros(2, dtype=np.float32)
    eng.iterate(zc, 0.05, 0, params=dict(transforms=T, zoom_padding=E.PAD_REFLECTION, fill=0.4, noise_facs=facs.numpy(),
                                         noise=noise), losses_out=losses)
    rel("d/d image (engine chain)", eng.debug_read("g_img", (1, 3, H, W)) / S, img_r.grad)
    rel("z.grad (engine chain)", eng.debug_read("z_grad", z.shape), zz.grad)
    # decoder backward alone: feed the ORACLE's image gradient through the engine's decoder?  not exposed; instead the
    # oracle's decoder backward of the ENGINE's image gradient
    gi = (eng.debug_read("g_img", (1, 3, H, W)) / S).cpu()
    zz2 = z.clone().requires_grad_(True)
    R.vqgan_synth(vq, zz2).backward(gi)
    rel("engine z.grad vs oracle-bwd(engine g_img)", eng.debug_read("z_grad", z.shape), zz2.grad)
    # ClampWithGrad (vqgan.py:76-79) passes a gradient where the unclamped pixel is inside [0, 1] or the gradient pushes it
    # back: a pixel within the fp16 forward error of the clamp boundary can sit on the other side in the engine
    with torch.no_grad():
        zq, _ = R.vector_quantize(z.movedim(1, 3), vq.quantize.embedding.weight)
        pre_o = vq.decode(zq.movedim(3, 1)).add(1).div(2)
    pre_e = eng.debug_read("img_pre", (1, 3, H, W)).cpu()
    side = lambda t: (t > 1).int() - (t < 0).int()  # noqa: E731
    flip = side(pre_e) != side(pre_o)
    print(f"  pixels whose clamp side differs between engine and oracle: {int(flip.sum())} of {flip.numel()} "
          f"(max |pre-clamp difference| {float((pre_e - pre_o).abs().max()):.3e})")
    if int(flip.sum()):
        # oracle backward with the ENGINE's clamp decisions: the same gradient through the same piece of the function
        zz3 = z.clone().requires_grad_(True)
        zq3, _ = R.vector_quantize(zz3.movedim(1, 3), vq.quantize.embedding.weight)
        pre3 = vq.decode(zq3.movedim(3, 1)).add(1).div(2)
        gi_in = gi.clone()
        pe = pre_e
        passes = ((pe >= 0) & (pe <= 1)) | ((pe < 0) & (gi_in < 0)) | ((pe > 1) & (gi_in > 0))  # g * (g * (x - clamp(x)) >= 0)
        pre3.backward(gi_in * passes)
        rel("engine z.grad vs oracle-bwd, engine's clamp mask", eng.debug_read("z_grad", z.shape), zz3.grad)


if __name__ == "__main__":
    import os
    if len(sys.argv) > 1 and sys.argv[1] == "seeds":
        # is the 48x32 decoder-backward error a property of the shape or of the data?  (also: PXR_GN_GROUP=0 /
        # PXR_CONV_SPLITK=0 in the environment switch the candidate kernels off)
        print("env:", {k: v for k, v in os.environ.items() if k.startswith("PXR_")})
        for hw in [(48, 32), (32, 48)]:
            for seed in (3, 4, 5, 6):
                run(hw, seed)
    elif len(sys.argv) > 1 and sys.argv[1] == "one":
        print("env:", {k: v for k, v in os.environ.items() if k.startswith("PXR_")})
        run((48, 32), 3)
    else:
        for hw in [(32, 48), (48, 32), (18, 32), (64, 32)]:
            run(hw)
