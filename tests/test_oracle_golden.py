"""CPU: pin oracle/ref_path.py against vectors produced by the REAL reference code (oracle/make_golden.py ran the
reference's own pixray.py / vqgan.py / slip.py / fast_pixeldrawer.py under oracle/shim.py in the authoring container).
Also cross-checks the restated CLIP VisionTransformer against transformers' CLIPVisionModelWithProjection."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_path as R

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"))


def t(name):
    return torch.from_numpy(np.asarray(G[name]))


def close(a, b, tol=1e-6):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    err = (a - b).abs().max().item()
    assert err <= tol * max(1.0, b.abs().max().item()), err


@pytest.mark.parametrize("tag", ["pos", "neg", "small"])
def test_prompt_forward_backward(tag):
    x = t("prompt_input").clone().requires_grad_(True)
    val = R.prompt_loss(x, t(f"prompt_{tag}_embed"), float(G[f"prompt_{tag}_w"]), float(G[f"prompt_{tag}_stop"]))
    val.backward()
    close(val.detach(), t(f"prompt_{tag}_val"))
    close(x.grad, t(f"prompt_{tag}_grad"))


def test_spherical_dist_loss():
    close(R.spherical_dist_loss(t("sph_x"), t("sph_y")), t("sph_out"))


def test_vector_quantize():
    x = t("vq_x").clone().requires_grad_(True)
    out, idx = R.vector_quantize(x, t("vq_codebook"))
    (out * t("vq_w")).sum().backward()
    close(out.detach(), t("vq_out"))
    close(x.grad, t("vq_grad"))
    assert idx.dtype == torch.int64


def test_clamp_with_grad():
    x = t("clamp_x").clone().requires_grad_(True)
    y = R.clamp_with_grad(x, 0, 1)
    y.backward(t("clamp_g"))
    close(y.detach(), t("clamp_y"))
    close(x.grad, t("clamp_dx"))


@pytest.mark.parametrize("mode", ["reflection", "border"])
def test_make_cutouts_cached_path(mode):
    img = t("cut_img").clone().requires_grad_(True)
    T = t("cut_T")
    cutn, cs = T.shape[0], 16
    # replay the reference's draws (pixray.py:509-510): uniform_ for the factors, then randn_like
    torch.manual_seed(77)
    facs = torch.empty(cutn, 1, 1, 1).uniform_(0, 0.1)
    noise = torch.randn(cutn, 3, cs, cs)
    batch = R.make_cutouts(img, T, cs, mode, 0.3, facs, noise)
    assert int(G["cut_zoom_n"]) == int(0.6 * cutn)
    close(batch.detach(), t(f"cut_{mode}_batch"), 1e-6)
    (batch * t("cut_w")).sum().backward()
    close(img.grad, t(f"cut_{mode}_dimg"), 1e-5)


def test_clip_base_preprocess_and_encode():
    vit = R.init_clip_weights(R.ClipVisual(32, 8, 64, 2, 1, 16), 3)
    imgs = t("clip_imgs").clone().requires_grad_(True)
    close(R.clip_preprocess(imgs).detach(), t("clip_pre"))
    emb = R.encode_image(vit, imgs)
    close(emb.detach(), t("clip_emb"), 1e-5)
    (emb * t("clip_w")).sum().backward()
    close(imgs.grad, t("clip_dimgs"), 1e-4)


def test_fast_pixel_synth():
    z = t("pixel_z").clone().requires_grad_(True)
    out = R.pixel_synth(z, tuple(t("pixel_out").shape[-2:]))
    close(out.detach(), t("pixel_out"))
    (out * t("pixel_w")).sum().backward()
    close(z.grad, t("pixel_dz"))


def test_vqgan_synth_and_clip_z():
    vq = R.init_vqgan_weights(R.VQModel(n_embed=64, embed_dim=32, ch=32, ch_mult=(1, 2), num_res_blocks=1,
                                        attn_resolutions=(4,), resolution=8, z_channels=32), 4)
    z = t("vqsynth_z").clone().requires_grad_(True)
    out = R.vqgan_synth(vq, z)
    close(out.detach(), t("vqsynth_out"), 1e-5)
    (out * t("vqsynth_w")).sum().backward()
    close(z.grad, t("vqsynth_dz"), 1e-4)
    zmin, zmax = R.vqgan_z_bounds(vq)
    close(torch.maximum(torch.minimum(t("clipz_in"), zmax), zmin), t("clipz_out"))


def test_adam_matches_torch_optim():
    z = t("adam_z0").clone()
    st = R.AdamState(z)
    for k in range(3):
        z = st.step(z, t(f"adam_g{k}"), 0.2)
        close(z, t(f"adam_z{k + 1}"), 1e-6)


def test_pool_window_bounds_are_atens():
    """Integer bookkeeping: the window bounds the CUDA pool kernel uses must be ATen's, checked via the real op."""
    for H, cs in [(256, 224), (512, 224), (32, 224), (24, 16), (20, 16)]:
        starts, ends = R.adaptive_pool_bounds(H, cs)
        x = torch.arange(H, dtype=torch.float32).reshape(1, 1, H, 1)
        mx = torch.nn.functional.adaptive_max_pool2d(x, (cs, 1)).reshape(-1)
        av = torch.nn.functional.adaptive_avg_pool2d(x, (cs, 1)).reshape(-1)
        assert [int(v) + 1 for v in mx] == ends
        assert all(abs(float(av[i]) - (starts[i] + ends[i] - 1) / 2) < 1e-4 for i in range(cs))


def test_vit_restatement_matches_transformers_clip():
    """Un-vendored leaf: openai-CLIP VisionTransformer.  Cross-check against HF transformers (quick_gelu)."""
    tr = pytest.importorskip("transformers")
    cfg = tr.CLIPVisionConfig(hidden_size=64, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2,
                              image_size=32, patch_size=8, projection_dim=16, hidden_act="quick_gelu",
                              layer_norm_eps=1e-5, attn_implementation="eager")
    hf = tr.CLIPVisionModelWithProjection(cfg).eval()
    mine = R.init_clip_weights(R.ClipVisual(32, 8, 64, 2, 2, 16), 9)
    v = mine.visual
    sd = hf.state_dict()
    with torch.no_grad():
        sd["vision_model.embeddings.patch_embedding.weight"].copy_(v.conv1.weight)
        sd["vision_model.embeddings.class_embedding"].copy_(v.class_embedding)
        sd["vision_model.embeddings.position_embedding.weight"].copy_(v.positional_embedding)
        for nm, ln in (("pre_layrnorm", v.ln_pre), ("post_layernorm", v.ln_post)):
            sd[f"vision_model.{nm}.weight"].copy_(ln.weight)
            sd[f"vision_model.{nm}.bias"].copy_(ln.bias)
        sd["visual_projection.weight"].copy_(v.proj.T)
        for i, blk in enumerate(v.transformer.resblocks):
            p = f"vision_model.encoder.layers.{i}."
            W, b = blk.attn.in_proj_weight, blk.attn.in_proj_bias
            for j, nm in enumerate(("q_proj", "k_proj", "v_proj")):
                sd[p + f"self_attn.{nm}.weight"].copy_(W[j * 64:(j + 1) * 64])
                sd[p + f"self_attn.{nm}.bias"].copy_(b[j * 64:(j + 1) * 64])
            sd[p + "self_attn.out_proj.weight"].copy_(blk.attn.out_proj.weight)
            sd[p + "self_attn.out_proj.bias"].copy_(blk.attn.out_proj.bias)
            sd[p + "layer_norm1.weight"].copy_(blk.ln_1.weight)
            sd[p + "layer_norm1.bias"].copy_(blk.ln_1.bias)
            sd[p + "layer_norm2.weight"].copy_(blk.ln_2.weight)
            sd[p + "layer_norm2.bias"].copy_(blk.ln_2.bias)
            sd[p + "mlp.fc1.weight"].copy_(blk.mlp.c_fc.weight)
            sd[p + "mlp.fc1.bias"].copy_(blk.mlp.c_fc.bias)
            sd[p + "mlp.fc2.weight"].copy_(blk.mlp.c_proj.weight)
            sd[p + "mlp.fc2.bias"].copy_(blk.mlp.c_proj.bias)
    x = torch.randn(2, 3, 32, 32, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        a = mine.encode_image(x)
        b = hf(pixel_values=x).image_embeds
    close(a, b, 1e-4)


# ------------------------------------------------------------------ auxiliary losses (Losses/*.py), oracle/make_golden_aux.py
GA = np.load(os.path.join(os.path.dirname(__file__), "golden", "aux_loss_vectors.npz"))


def ta(name):
    return torch.from_numpy(np.asarray(GA[name]))


def _check_aux(tag, fn, on_cut):
    x = (ta("cut") if on_cut else ta("out")).clone().requires_grad_(True)
    val = fn(x)
    val.backward()
    close(val.detach(), ta(tag + "_val"), 2e-6)
    close(x.grad, ta(tag + "_grad"), 2e-6)


def test_aux_symmetry():
    _check_aux("symmetry", lambda o: R.symmetry_loss(o, 0.7), False)


def test_aux_saturation():
    _check_aux("saturation", lambda c: R.saturation_loss(c, 1.3), True)


def test_aux_palette():
    _check_aux("palette", lambda c: R.palette_loss(c, ta("palette"), 0.8)[0], True)


@pytest.mark.parametrize("kind", ["default", "clipped", "log"])
def test_aux_smoothness(kind):
    _check_aux("smooth_" + kind, lambda c: R.smoothness_loss(c, 0.9, kind), True)


def test_aux_smoothness_spacing():
    _check_aux("smooth_spacing2", lambda c: R.smoothness_loss(c, 1.0, "default", spacing=2), True)


def test_aux_edge():
    margins = R.edge_margins_px([10, 20, 15, 0], 20, 28)
    _check_aux("edge", lambda o: R.edge_loss(o, ta("edge_color").tolist(), margins, 0.1, 0.05), False)


def test_aux_gaussian():
    _check_aux("gaussian", lambda o: R.gaussian_loss(o, (6.0, 9.0), (255, 128, 0), 0.6), False)


def test_aux_aesthetic():
    e = ta("aes_emb").clone().requires_grad_(True)
    val = R.aesthetic_loss(e, ta("aes_w"), ta("aes_b"), 10.0)
    val.backward()
    close(val.detach(), ta("aes_val"), 2e-6)
    close(e.grad, ta("aes_grad"), 2e-6)


# ------------------------------------------------------------------ vdiff drawer (cc12m_1), oracle/make_golden_vdiff.py
@pytest.mark.slow
def test_vdiff_restatement_matches_the_reference_model():
    """torch.manual_seed(0) + the same constructor order reproduce the reference's seeded weights, hence its outputs."""
    GV = np.load(os.path.join(os.path.dirname(__file__), "golden", "vdiff_vectors.npz"))
    torch.manual_seed(0)
    m = R.VDiffCC12M1().eval().requires_grad_(False)
    sd = m.ref_state_dict()
    assert len(sd) == int(GV["n_tensors"]) and sum(p.numel() for p in sd.values()) == int(GV["n_params"])
    for k in GV.files:
        if k.startswith("w:"):
            assert np.array_equal(sd[k[2:]].reshape(-1)[:64].numpy(), GV[k]), k
    x = torch.from_numpy(GV["x"]).requires_grad_(True)
    v = m(x, torch.from_numpy(GV["t"]), torch.from_numpy(GV["clip_embed"]))
    (v * torch.from_numpy(GV["w"])).sum().backward()
    close(v.detach(), torch.from_numpy(GV["v"]), 1e-5)
    close(x.grad, torch.from_numpy(GV["dx"]), 1e-5)


def test_vdiff_schedule_and_renoise():
    GV = np.load(os.path.join(os.path.dirname(__file__), "golden", "vdiff_vectors.npz"))
    steps, alphas, sigmas = R.vdiff_schedule(20)
    close(steps, torch.from_numpy(GV["steps"]), 1e-6)
    close(alphas, torch.from_numpy(GV["alphas"]), 1e-6)
    close(sigmas, torch.from_numpy(GV["sigmas"]), 1e-6)
    x, pred, v = (torch.from_numpy(GV[k]) for k in ("rn_x", "rn_pred", "rn_v"))
    for i in (0, 7, 20):
        torch.manual_seed(123)
        noise = torch.randn_like(x)
        close(R.vdiff_renoise(x, pred, v, alphas, sigmas, i, noise), torch.from_numpy(GV[f"renoise_{i}"]), 1e-6)


def test_filters_reproduce_the_reference_classes():
    """oracle.filter_* against filters/{tiler,wallpaper,colorlookup}.py run for real (oracle/make_golden_filters.py): output
    image, loss and the gradient of sum(out * up) + loss, on the (rand_h, rand_w) the reference class itself drew."""
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "filter_vectors.npz"))
    img = torch.from_numpy(G["img"])

    def check(tag, fn):
        rh, rw = (int(v) for v in G[tag + "_rand"])
        x = img.clone().requires_grad_(True)
        out, loss = fn(x, rh, rw)
        ((out * torch.from_numpy(G[tag + "_up"])).sum() + loss).backward()
        assert tuple(out.shape) == tuple(G[tag + "_out"].shape), tag
        assert np.abs(out.detach().numpy() - G[tag + "_out"]).max() <= 1e-6, tag
        assert abs(float(loss) - float(G[tag + "_loss"])) <= 1e-6 * max(1.0, abs(float(G[tag + "_loss"]))), tag
        assert np.abs(x.grad.numpy() - G[tag + "_grad"]).max() <= 1e-6, tag

    check("tiler", lambda x, rh, rw: R.filter_tiler(x, rh, rw))
    for wt, em in (("shift", 0), ("horizontal", 0), ("horizontal", 6), ("vertical", 4), (None, 0), (None, 6)):
        check(f"wallpaper_{wt}_{em}", lambda x, rh, rw, wt=wt, em=em: R.filter_wallpaper(x, wt, em, rh, rw))
    pal = G["palette"].tolist()
    check("lookup", lambda x, rh, rw: R.filter_colorlookup(x, pal, 3.0))
    check("lookup_default", lambda x, rh, rw: R.filter_colorlookup(x, None, 10.0))
