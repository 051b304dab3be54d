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
