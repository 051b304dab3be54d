"""Generate tests/golden/reference_vectors.npz by running the REAL reference code (container-only).

    python -m oracle.make_golden

Each entry pins one in-tree piece of the hot path (the reference's own Python executed unmodified through
oracle/shim.py); tests/test_oracle_golden.py then requires oracle/ref_path.py to reproduce every vector.
Un-vendored leaves (kornia warp, CLIP ViT, taming Decoder) are the restatements in both runs, so for them this pins
the reference's *use* of the leaf (argument order, group split, padding mode, noise, normalisation), not the leaf.
"""
import argparse
import os
import types

import numpy as np
import torch

from oracle import ref_path as R
from oracle import shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                   "reference_vectors.npz")


def main():
    px = shim.install()
    import fast_pixeldrawer
    import slip
    import vqgan as ref_vqgan

    G = {}
    g = torch.Generator().manual_seed(1234)

    def rnd(*s):
        return torch.randn(*s, generator=g)

    # 1. Prompt.forward (pixray.py:268-280) value + gradient, positive / negative weight, finite stop
    emb_in = rnd(6, 16)
    for tag, (w, stop) in {"pos": (1.0, float("-inf")), "neg": (-0.5, -0.3), "small": (0.1, float("-inf"))}.items():
        embed = rnd(2, 16)
        x = emb_in.clone().requires_grad_(True)
        p = px.Prompt(embed, w, stop)
        val = p(x)
        val.backward()
        G[f"prompt_{tag}_embed"], G[f"prompt_{tag}_w"], G[f"prompt_{tag}_stop"] = embed.numpy(), np.float32(w), np.float32(max(stop, -3e38))
        G[f"prompt_{tag}_val"], G[f"prompt_{tag}_grad"] = val.detach().numpy(), x.grad.numpy()
    G["prompt_input"] = emb_in.numpy()

    # 2. spherical_dist_loss (pixray.py:262-265)
    a, b = rnd(3, 8), rnd(3, 8)
    G["sph_x"], G["sph_y"], G["sph_out"] = a.numpy(), b.numpy(), px.spherical_dist_loss(a, b).numpy()

    # 3. vector_quantize (vqgan.py:60-64) value + straight-through gradient
    cb = rnd(32, 8)
    xq = (cb[torch.randint(32, (16,), generator=g)] + 0.05 * rnd(16, 8)).reshape(1, 4, 4, 8).requires_grad_(True)
    out = ref_vqgan.vector_quantize(xq, cb)
    wsum = rnd(1, 4, 4, 8)
    (out * wsum).sum().backward()
    G["vq_x"], G["vq_codebook"], G["vq_out"], G["vq_w"], G["vq_grad"] = xq.detach().numpy(), cb.numpy(), out.detach().numpy(), wsum.numpy(), xq.grad.numpy()

    # 4. ClampWithGrad (vqgan.py:66-79)
    xc = (rnd(64) * 0.8 + 0.5).requires_grad_(True)
    gc = rnd(64)
    yc = ref_vqgan.clamp_with_grad(xc, 0, 1)
    yc.backward(gc)
    G["clamp_x"], G["clamp_g"], G["clamp_y"], G["clamp_dx"] = xc.detach().numpy(), gc.numpy(), yc.detach().numpy(), xc.grad.numpy()

    # 5. MakeCutouts.forward, cached-transform path (pixray.py:445-511): pooling, zoom/wide split, padding, fill, noise
    cutn, cs = 5, 16
    mc = px.MakeCutouts(cs, cutn)
    img = torch.rand(1, 3, 24, 20, generator=g).requires_grad_(True)
    T = torch.eye(3).repeat(cutn, 1, 1)
    T[:, 0, 0] = torch.tensor([1.3, 1.1, 0.9, 0.85, 0.95])
    T[:, 1, 1] = torch.tensor([1.2, 1.4, 1.0, 0.9, 0.8])
    T[:, 0, 1] = torch.tensor([0.05, -0.1, 0.0, 0.02, 0.0])
    T[:, 0, 2] = torch.tensor([-3.0, 1.5, -6.0, 1.0, 2.0])
    T[:, 1, 2] = torch.tensor([2.0, -4.0, 0.5, 0.7, 1.5])
    T[:, 2, 0] = torch.tensor([1e-3, -2e-3, 0.0, 1e-3, 0.0])
    for mode in ("reflection", "border"):
        px.global_padding_mode = mode
        px.global_fill_color = torch.tensor([0.3, 0.3, 0.3])
        mc.transforms = T.clone()
        torch.manual_seed(77)
        batch = mc(img)
        wb = torch.randn(batch.shape, generator=torch.Generator().manual_seed(5))
        img.grad = None
        (batch * wb).sum().backward()
        G[f"cut_{mode}_batch"], G[f"cut_{mode}_dimg"] = batch.detach().numpy(), img.grad.numpy().copy()
    G["cut_img"], G["cut_T"], G["cut_w"] = img.detach().numpy(), T.numpy(), wb.numpy()
    G["cut_zoom_n"] = np.int32(mc.cutn_zoom)

    # 6. CLIP_Base.preprocess / encode_image (slip.py:21-66) on a small restated ViT
    vit = R.init_clip_weights(R.ClipVisual(32, 8, 64, 2, 1, 16), 3)
    base = slip.CLIP_Base(vit, None, "cpu")
    imgs = (torch.rand(3, 3, 32, 32, generator=g) * 1.3 - 0.2).requires_grad_(True)
    pre = base.preprocess(imgs)
    emb = base.encode_image(imgs)
    wv = rnd(3, 16)
    (emb * wv).sum().backward()
    G["clip_imgs"], G["clip_pre"], G["clip_emb"], G["clip_w"], G["clip_dimgs"] = imgs.detach().numpy(), pre.detach().numpy(), emb.detach().numpy(), wv.numpy(), imgs.grad.numpy()

    # 7. FastPixelDrawer.synth (fast_pixeldrawer.py:83-91)
    settings = types.SimpleNamespace(size=[24, 32], pixel_size=[6, 8], pixel_scale=None)
    d = fast_pixeldrawer.FastPixelDrawer(settings)
    d.z = (torch.rand(1, 3, 8, 6, generator=g) * 1.4 - 0.2).requires_grad_(True)
    o = d.synth(0)
    wo = rnd(*o.shape)
    (o * wo).sum().backward()
    G["pixel_z"], G["pixel_out"], G["pixel_w"], G["pixel_dz"] = d.z.detach().numpy(), o.detach().numpy(), wo.numpy(), d.z.grad.numpy()

    # 8. VqganDrawer.synth + clip_z (vqgan.py:190-204) over the restated VQModel
    vq = R.init_vqgan_weights(R.VQModel(n_embed=64, embed_dim=32, ch=32, ch_mult=(1, 2), num_res_blocks=1,
                                        attn_resolutions=(4,), resolution=8, z_channels=32), 4)
    drawer = ref_vqgan.VqganDrawer.__new__(ref_vqgan.VqganDrawer)
    drawer.model, drawer.gumbel = vq, False
    drawer.z_min = vq.quantize.embedding.weight.min(dim=0).values[None, :, None, None]
    drawer.z_max = vq.quantize.embedding.weight.max(dim=0).values[None, :, None, None]
    z = (vq.quantize.embedding.weight[torch.randint(64, (16,), generator=g)].T.reshape(1, 32, 4, 4) + 0.05 * rnd(1, 32, 4, 4))
    drawer.z = z.clone().requires_grad_(True)
    o = drawer.synth(0)
    wo = rnd(*o.shape)
    (o * wo).sum().backward()
    G["vqsynth_z"], G["vqsynth_out"], G["vqsynth_w"], G["vqsynth_dz"] = z.numpy(), o.detach().numpy(), wo.numpy(), drawer.z.grad.numpy()
    drawer.z = (z * 3.0).clone().requires_grad_(True)
    drawer.clip_z()
    G["clipz_in"], G["clipz_out"] = (z * 3.0).numpy(), drawer.z.detach().numpy()

    # 9. optim.Adam as rebuild_optimisers builds it (pixray.py:538-539), 3 steps
    za = rnd(40).requires_grad_(True)
    opt = torch.optim.Adam([za], lr=0.2)
    grads = [rnd(40) for _ in range(3)]
    G["adam_z0"] = za.detach().numpy().copy()
    for k, gr in enumerate(grads):
        opt.zero_grad()
        za.grad = gr.clone()
        opt.step()
        G[f"adam_g{k}"], G[f"adam_z{k + 1}"] = gr.numpy(), za.detach().numpy().copy()

    # 10. learning-rate drops / parse_unit (the reference's own unit tests, tests/test_pixray.py:54-64)
    G["lr_drops_75_300"] = np.array(px.get_learning_rate_drops([75], 300), dtype=np.int32)
    G["lr_drops_50_225_300"] = np.array(px.get_learning_rate_drops([50, 22.5], 300), dtype=np.int32)

    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **G)
    print("wrote", OUT, "with", len(G), "arrays,", os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    argparse.ArgumentParser().parse_args()
    main()
