"""Cross-checks of the oracle's restatements of UN-VENDORED leaves (kornia 0.6.2 is not in /root/reference
and not installable here, so these pieces cannot be pinned to kornia itself) against an independent implementation of
the same published operation.  They guard the conventions a self-written oracle and a self-written kernel could share
by mistake: direction of the homography, pixel-centre convention, border modes."""
import numpy as np
import pytest
import torch

from oracle import ref_path as R
from pixray_b200 import cutouts

cv2 = pytest.importorskip("cv2")


def _smooth_image(cs):
    yy, xx = np.mgrid[0:cs, 0:cs].astype(np.float32)
    return np.stack([0.5 + 0.4 * np.sin(xx / 9.0 + c) * np.cos(yy / 7.0 - c) for c in range(3)], 0).astype(np.float32)


@pytest.mark.parametrize("mode,cvmode,tol", [("reflection", cv2.BORDER_REFLECT_101, 2e-3), ("border", cv2.BORDER_REPLICATE, 2e-3),
                                             ("fill", cv2.BORDER_CONSTANT, 2.5e-2)])
def test_warp_perspective_conventions_match_opencv(mode, cvmode, tol):
    """kornia.geometry.transform.warp_perspective(src, M, dsize, align_corners=True, padding_mode=...) as MakeCutouts calls it
    (pixray.py:334, 351-352, 482-485) == cv2.warpPerspective(src, M, dsize, INTER_LINEAR, borderMode): M maps source pixels
    to destination pixels, pixel centres sit on integers.  OpenCV quantises the bilinear weights to 1/32, which bounds the
    agreement (2e-3 on a smooth image; 2.5e-2 at the image / fill-colour step of the constant border); a half-pixel
    convention error would be an order of magnitude above that, which the test demonstrates."""
    cs, cutn = 64, 8
    img = _smooth_image(cs)
    T = cutouts.sample_transforms(cutn, cs, 3)
    src = torch.from_numpy(img)[None].expand(cutn, -1, -1, -1)
    kw = {"fill_value": [0.3, 0.3, 0.3]} if mode == "fill" else {}
    got = R.warp_perspective(src, torch.from_numpy(T), (cs, cs), padding_mode=mode, **kw).numpy()
    worst = shifted = 0.0
    half = np.array([[1, 0, 0.5], [0, 1, 0.5], [0, 0, 1]], dtype=np.float64)
    for n in range(cutn):
        def cv(M):
            return cv2.warpPerspective(img.transpose(1, 2, 0), M, (cs, cs), flags=cv2.INTER_LINEAR, borderMode=cvmode,
                                       borderValue=(0.3, 0.3, 0.3)).transpose(2, 0, 1)
        worst = max(worst, float(np.abs(got[n] - cv(T[n].astype(np.float64))).max()))
        shifted = max(shifted, float(np.abs(got[n] - cv(half @ T[n].astype(np.float64))).max()))
    assert worst <= tol, worst
    if mode != "fill":
        assert shifted > 5 * worst      # the check has the power to see a half-pixel convention error


def test_adaptive_pool_bounds_match_torch():
    """AdaptiveAvg/MaxPool2d window bounds (pixray.py:442-443, 461) as the engine's pool kernel computes them."""
    for n_in, n_out in ((256, 224), (512, 224), (32, 224), (224, 224), (300, 224)):
        x = torch.arange(n_in, dtype=torch.float32).reshape(1, 1, 1, n_in)
        want_max = torch.nn.functional.adaptive_max_pool2d(x, (1, n_out)).reshape(-1)
        starts, ends = R.adaptive_pool_bounds(n_in, n_out)
        assert torch.equal(want_max, torch.as_tensor(ends, dtype=torch.float32) - 1)      # max of an increasing ramp = end - 1
        want_avg = torch.nn.functional.adaptive_avg_pool2d(x, (1, n_out)).reshape(-1)
        mine = torch.tensor([(s + e - 1) / 2.0 for s, e in zip(starts, ends)])
        assert torch.allclose(want_avg, mine, atol=1e-4)


def test_last_vit_layer_on_class_rows_only_is_exact():
    """The engine runs the LAST transformer block's out_proj / ln_2 / MLP (and their backward) on the class-token row of each
    image only (pixray_b200/csrc/engine.cu, build_clip: `cls_only`).  VisionTransformer.forward keeps x[:, 0, :] after the last
    block and everything behind that block's attention is row-wise, so this must change neither the embedding nor the gradient
    with respect to the input: checked here on the oracle's ViT in float64, with the attention output's gradient zeroed outside
    the class rows exactly as the engine does it."""
    import torch.nn.functional as F
    torch.manual_seed(0)
    clip = R.init_clip_weights(R.ClipVisual(64, 16, 64, 3, 1, 32), 5).double()
    v = clip.visual
    x_img = torch.randn(3, 3, 64, 64, dtype=torch.float64, requires_grad=True)
    full = v(x_img)
    (g_full,) = torch.autograd.grad(full.square().sum() + full.sum(), x_img)

    # the same network with the last block's tail restricted to the class rows
    x2 = x_img.detach().clone().requires_grad_(True)
    t = v.conv1(x2)
    t = t.reshape(t.shape[0], t.shape[1], -1).permute(0, 2, 1)
    cls = v.class_embedding + torch.zeros(t.shape[0], 1, t.shape[-1], dtype=t.dtype)
    t = v.ln_pre(torch.cat([cls, t], dim=1) + v.positional_embedding)
    t = t.permute(1, 0, 2)                                      # [T, B, W]
    blocks = list(v.transformer.resblocks)
    for blk in blocks[:-1]:
        t = blk(t)
    last = blocks[-1]
    h = last.ln_1(t)
    W = h.shape[-1]
    q, k, vv = F.linear(h, last.attn.in_proj_weight, last.attn.in_proj_bias).chunk(3, dim=-1)
    nh = last.attn.num_heads
    hd = W // nh
    Tn, Bn = h.shape[0], h.shape[1]
    split = lambda z: z.reshape(Tn, Bn * nh, hd).transpose(0, 1)      # noqa: E731  [B*heads, T, hd]
    att = torch.softmax(split(q) @ split(k).transpose(1, 2) / hd ** 0.5, dim=-1) @ split(vv)
    o = att.transpose(0, 1).reshape(Tn, Bn, W)                   # attention output of EVERY token (the engine computes all)
    o_cls = o[0]                                                 # ... but only the class rows go on
    x_mid = t[0] + F.linear(o_cls, last.attn.out_proj.weight, last.attn.out_proj.bias)
    x_out = x_mid + last.mlp(last.ln_2(x_mid))
    short = v.ln_post(x_out) @ v.proj
    (g_short,) = torch.autograd.grad(short.square().sum() + short.sum(), x2)
    assert (short - full).abs().max().item() < 1e-12
    assert (g_short - g_full).abs().max().item() < 1e-12 * max(1.0, g_full.abs().max().item())
