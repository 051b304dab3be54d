"""Integer bookkeeping of the path, bit-exact on the GPU (pool window bounds, VQ argmin code indices,
max-pool arg-max elements), and the pixel drawer (fast_pixeldrawer.py:83-91) through the engine against the oracle."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import ref_path as R
from pixray_b200 import _lib
from pixray_b200 import engine as E

pytestmark = pytest.mark.gpu


def test_pool_window_bounds_equal_aten_on_the_device():
    """pool_fwd / pool_bwd take their windows from pool_start / pool_end (csrc/pool_bounds.cuh); the hook runs those
    same functions on the device.  Sizes: every BASELINE canvas side against cut_size 224, plus ragged ones."""
    lib = _lib.load()
    f = lib.pxr_test_pool_bounds
    f.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    for in_size, out_size in [(32, 224), (256, 224), (512, 224), (144, 224), (400, 224), (225, 224), (223, 224), (1, 224),
                              (7, 3), (1000, 224), (224, 224)]:
        s = torch.empty(out_size, dtype=torch.int32, device="cuda")
        e = torch.empty(out_size, dtype=torch.int32, device="cuda")
        assert f(in_size, out_size, s.data_ptr(), e.data_ptr()) == 0
        rs, re_ = R.adaptive_pool_bounds(in_size, out_size)
        assert s.cpu().tolist() == rs and e.cpu().tolist() == re_, (in_size, out_size)


def test_pool_values_and_argmax_elements_equal_torch():
    """(AdaptiveAvgPool2d + AdaptiveMaxPool2d) / 2 (pixray.py:463) on ragged canvases (windows of 1..3 pixels, both
    down- and up-sampling) and the arg-max element each max-pool window routes its gradient to (first maximum, like
    ATen), read back from the engine."""
    from test_pipeline_gpu import SMALL_CLIP
    from pixray_b200 import synthetic as S
    for (H, W) in [(300, 260), (32, 32), (512, 448)]:
        eng = E.B200Engine(drawer=E.DRAWER_PIXEL, image_hw=(H, W), grid=(10, 10), cutn=8, clip=[SMALL_CLIP], seed=1)
        eng.load_module(E.MOD_CLIP0, S.clip_state_dict(SMALL_CLIP, 1))
        eng.finalize()
        torch.manual_seed(4)
        img = torch.rand(1, 3, H, W)
        img[0, :, 3:9, 3:9] = 0.75                                     # exact ties inside windows
        img[0, 1, 20:, :] = (img[0, 1, 20:, :] * 4).round() / 4        # a channel full of ties
        eng.make_cutouts(img, transforms=np.tile(np.eye(3, dtype=np.float32), (8, 1, 1)), zoom_padding=E.PAD_BORDER, fill=0.0)
        pooled = eng.debug_read("pooled", (1, 3, 224, 224)).cpu()
        amax = eng.debug_read("pool_argmax", (3, 224, 224), dtype=torch.int32).cpu()
        ref = R.pool_avg_max(img, 224)
        _, ref_idx = torch.nn.functional.adaptive_max_pool2d(img, (224, 224), return_indices=True)
        assert (pooled - ref).abs().max().item() <= 1e-6, (H, W)
        assert torch.equal(amax.long(), ref_idx[0]), (H, W)


def test_vq_code_indices_bit_exact():
    """vector_quantize's argmin (vqgan.py:60-64) on the GPU equals the oracle's indices -- at the small test codebook and
    at the full imagenet_f16_16384 size (16384 x 256), with latents that sit between codes (0.05-sigma noise)."""
    from pixray_b200 import synthetic as S
    from test_pipeline_gpu import build
    vq, clip, eng, prompts, z = build(cutn=8, seed=21)
    eng.synth(z)
    idx = eng.debug_read("vq_idx", (z.shape[2] * z.shape[3],), dtype=torch.int32).cpu().long()
    _, ref_idx = R.vector_quantize(z.movedim(1, 3), vq.quantize.embedding.weight)
    assert torch.equal(idx, ref_idx.reshape(-1))
    # full-size codebook: only the quantiser matters, so run the 256x256 engine's synth and read the indices
    vq_sd = S.vqgan_state_dict(E.VQGAN_F16_16384, 0)
    eng2 = E.B200Engine(drawer=E.DRAWER_VQGAN, image_hw=(256, 256), cutn=8, clip=[dict(width=128, layers=1, heads=2, patch=32, image_res=224, out_dim=64)], seed=0)
    eng2.load_module(E.MOD_VQGAN, vq_sd)
    eng2.load_module(E.MOD_CLIP0, S.clip_state_dict(dict(width=128, layers=1, heads=2, patch=32, image_res=224, out_dim=64), 1))
    eng2.finalize()
    cb = vq_sd["quantize.embedding.weight"]
    g = torch.Generator().manual_seed(5)
    for trial in range(3):
        pick = torch.randint(16384, (256,), generator=g)
        z2 = (cb[pick].T.reshape(1, 256, 16, 16) + (0.02 + 0.2 * trial) * cb.std() * torch.randn(1, 256, 16, 16, generator=g)).contiguous()
        eng2.synth(z2)
        idx2 = eng2.debug_read("vq_idx", (256,), dtype=torch.int32).cpu().long()
        _, ref2 = R.vector_quantize(z2.movedim(1, 3).double(), cb.double())      # exact reference for the argmin
        _, ref2f = R.vector_quantize(z2.movedim(1, 3), cb)
        same = idx2 == ref2.reshape(-1)
        # fp32 near-ties: where the fp64 argmin and torch's own fp32 argmin disagree, either is a valid fp32 answer
        amb = ref2.reshape(-1) != ref2f.reshape(-1)
        assert bool((same | amb).all()), f"trial {trial}: {int((~(same | amb)).sum())} indices differ"
        print(f"[parity] VQ indices trial {trial}: {int(same.sum())}/256 equal the fp64 argmin, {int(amb.sum())} fp32-ambiguous")


def test_vq_tensor_core_search_equals_the_fp32_search():
    """The nearest-code search with the distance matrix on the tensor cores + exact fp32 recheck of the candidates
    (kernels_vq_tc.cu) picks the same index as the all-fp32 search (PXR_VQ_TC=0) -- full-size codebook, latents from
    'on a code' to 'far between codes', plus a degenerate codebook with duplicated rows (exact ties: the first index wins)."""
    import os
    from pixray_b200 import synthetic as S
    clip_cfg = dict(width=128, layers=1, heads=2, patch=32, image_res=224, out_dim=64)
    vq_sd = S.vqgan_state_dict(E.VQGAN_F16_16384, 0)
    cb = vq_sd["quantize.embedding.weight"]
    cb_dup = cb.clone()
    cb_dup[8000:8100] = cb[100:200]          # duplicated rows: ties between j and j + 7900
    engines = {}
    for name, sd_cb in (("plain", cb), ("dup", cb_dup)):
        sd = dict(vq_sd)
        sd["quantize.embedding.weight"] = sd_cb
        for tc in ("1", "0"):
            os.environ["PXR_VQ_TC"] = tc
            try:
                e = E.B200Engine(drawer=E.DRAWER_VQGAN, image_hw=(256, 256), cutn=8, clip=[clip_cfg], seed=0)
                e.load_module(E.MOD_VQGAN, sd)
                e.load_module(E.MOD_CLIP0, S.clip_state_dict(clip_cfg, 1))
                e.finalize()
            finally:
                os.environ.pop("PXR_VQ_TC", None)
            engines[(name, tc)] = e
    g = torch.Generator().manual_seed(9)
    for name, sd_cb in (("plain", cb), ("dup", cb_dup)):
        for trial, sigma in enumerate([0.0, 0.02, 0.3, 1.0, 3.0]):
            pick = torch.randint(16384, (256,), generator=g)
            if name == "dup":
                pick[:64] = torch.randint(100, 200, (64,), generator=g)      # land on duplicated rows
            z = (sd_cb[pick].T.reshape(1, 256, 16, 16) + sigma * sd_cb.std() * torch.randn(1, 256, 16, 16, generator=g)).contiguous()
            if trial == 4:
                z = z * 40.0        # far outside the codebook's range (the scaling path)
            out = {}
            for tc in ("1", "0"):
                engines[(name, tc)].synth(z)
                out[tc] = engines[(name, tc)].debug_read("vq_idx", (256,), dtype=torch.int32).cpu()
            assert torch.equal(out["1"], out["0"]), (name, trial, int((out["1"] != out["0"]).sum()))
            if name == "dup" and sigma == 0.0:
                assert int(out["1"][:64].max()) < 200   # first index of the tied pair
        st = engines[(name, "1")].debug_read("vq_stats", (2,), dtype=torch.int32).cpu()
        print(f"[parity] VQ tensor-core search ({name}): {int(st[0])} candidates rechecked over {5 * 256} positions, "
              f"{int(st[1])} positions fell back to the full search")


def test_pixel_drawer_matches_the_oracle():
    """FastPixelDrawer.synth (fast_pixeldrawer.py:83-91): nearest upsample of the colour grid + clamp_with_grad; forward
    image, z.grad and the Adam / clip_z([0,1]) update, through the engine's PXR_DRAWER_PIXEL path.  Includes a grid that
    does not divide the canvas and colours outside [0,1] (the clamp's backward mask)."""
    from test_pipeline_gpu import SMALL_CLIP, plant_extremes, random_transforms, report
    for (H, W, rows, cols, seed) in [(32, 32, 8, 8, 0), (48, 64, 9, 16, 1), (64, 64, 64, 64, 2)]:
        cutn, cs = 8, 224
        clip = R.init_clip_weights(R.ClipVisual(224, SMALL_CLIP["patch"], SMALL_CLIP["width"], SMALL_CLIP["layers"],
                                                SMALL_CLIP["heads"], SMALL_CLIP["out_dim"]), seed + 1)
        eng = E.B200Engine(drawer=E.DRAWER_PIXEL, image_hw=(H, W), grid=(rows, cols), cutn=cutn, clip=[SMALL_CLIP],
                           noise_fac=0.1, seed=seed)
        eng.load_module(E.MOD_CLIP0, clip.state_dict())
        eng.finalize()
        g = torch.Generator().manual_seed(seed + 2)
        prompts = [(torch.randn(1, 64, generator=g), 1.0, float("-inf")), (torch.randn(1, 64, generator=g), -0.3, float("-inf"))]
        eng.set_prompts(0, torch.cat([p[0] for p in prompts]).numpy(), [1.0, -0.3], [float("-inf")] * 2)
        z = (torch.rand(1, 3, rows, cols, generator=g) * 1.4 - 0.2).contiguous()   # some colours outside [0, 1]
        T = random_transforms(cutn, cs, 3 + seed)
        facs, noise = plant_extremes(torch.rand(cutn, generator=g) * 0.1, torch.randn(cutn, 3, cs, cs, generator=g))
        ref = R.iterate(lambda zz: R.pixel_synth(zz, (H, W)), z, [clip], [prompts], torch.from_numpy(T), cs, "border", 0.4,
                        facs, noise)
        img = eng.synth(z)
        e_img, _ = report(f"pixel image {H}x{W} grid {rows}x{cols}", img, ref["image"])
        assert e_img == 0.0                                          # a gather + clamp: exact
        zc = z.clone().cuda()
        losses = np.zeros(2, dtype=np.float32)
        eng.iterate(zc, 0.03, 1, params=dict(transforms=T, zoom_padding=E.PAD_BORDER, fill=0.4, noise_facs=facs.numpy(),
                                             noise=noise), losses_out=losses)
        zg = eng.debug_read("z_grad", z.shape).cpu()
        e_g, m_g = report("pixel z.grad", zg, ref["z_grad"])
        ref_l = np.array([float(l) for l in ref["losses"]], dtype=np.float32)
        assert np.abs(losses - ref_l).max() < 5e-3
        assert e_g <= 3e-2 * m_g
        # colours clamped from outside with the gradient pushing them further out get no gradient (ClampWithGrad)
        z_next = R.AdamState(z).step(z, zg, 0.03).clip(0, 1)         # FastPixelDrawer.clip_z, fast_pixeldrawer.py:101-103
        e_z, _ = report("pixel z after Adam + clip_z", zc, z_next)
        assert e_z < 2e-5
