"""GroupNorm(32, C, eps=1e-6) (+ swish) of the taming decoder (oracle/ref_path.py Normalize / nonlinearity) on the device:
the cluster-per-group kernels (kernels_gn_group.cu) against torch fp32 autograd and against the single-kernel grid-barrier
version, plain and as the fused epilogue of a split-K convolution (fp32 partial sums in, reduced tensor + normalised tensor
out; backward straight from the partials)."""
import ctypes as C

import pytest
import torch

from pixray_b200 import _lib

pytestmark = pytest.mark.gpu

SHAPES = [(256, 512), (1024, 256), (1024, 512), (4096, 256), (16384, 256), (4096, 512)]


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _run(variant, x, gamma, beta, swish, dy=None, dres=None, ws=None, bias=None, res=None, ws_dy=None, repeat=1):
    lib = _lib.load()
    px, c = (x.shape if x is not None else ws.shape[1:])
    y = torch.empty(px, c, dtype=torch.half, device="cuda")
    x_out = torch.empty(px, c, dtype=torch.half, device="cuda") if ws is not None else None
    stats = torch.zeros(64, dtype=torch.float32, device="cuda")
    dx = torch.empty(px, c, dtype=torch.half, device="cuda") if (dy is not None or ws_dy is not None) else None
    nsm = torch.cuda.get_device_properties(0).multi_processor_count
    scratch = torch.zeros(64 * (nsm + 2) + 2, dtype=torch.float32, device="cuda")
    rc = lib.pxr_test_groupnorm(variant, _ptr(x), _ptr(ws), 0 if ws is None else ws.shape[0], _ptr(bias), _ptr(res), _ptr(x_out),
                                _ptr(gamma), _ptr(beta), px, c, swish, _ptr(y), _ptr(stats), _ptr(dy), _ptr(ws_dy),
                                0 if ws_dy is None else ws_dy.shape[0], _ptr(dres), _ptr(dx), _ptr(scratch), repeat)
    assert rc == 0, rc
    return y, stats, dx, x_out


def _reference(x16, gamma, beta, swish, dy16=None, dres16=None):
    x = x16.float().requires_grad_(True)
    c = x.shape[1]
    h = torch.nn.functional.group_norm(x.t().reshape(1, c, -1), 32, gamma, beta, eps=1e-6).reshape(c, -1).t()
    y = h * torch.sigmoid(h) if swish else h
    dx = None
    if dy16 is not None:
        (dx,) = torch.autograd.grad(y, x, dy16.float())
        if dres16 is not None:
            dx = dx + dres16.float()
    return y.detach(), dx


@pytest.mark.parametrize("px,c", SHAPES)
@pytest.mark.parametrize("swish", [1, 0])
def test_group_kernel_matches_torch_and_the_grid_barrier_kernel(px, c, swish):
    g = torch.Generator(device="cuda").manual_seed(px + c + swish)
    x = (torch.randn(px, c, generator=g, device="cuda") * 1.5 + 0.3 * torch.randn(1, c, generator=g, device="cuda")).half()
    gamma = 1 + 0.2 * torch.randn(c, generator=g, device="cuda")
    beta = 0.2 * torch.randn(c, generator=g, device="cuda")
    dy = (torch.randn(px, c, generator=g, device="cuda") * 0.7).half()
    dres = (torch.randn(px, c, generator=g, device="cuda") * 0.7).half()
    y_ref, dx_ref = _reference(x, gamma, beta, swish, dy, dres)
    y1, st1, dx1, _ = _run(1, x, gamma, beta, swish, dy, dres)
    y0, st0, dx0, _ = _run(0, x, gamma, beta, swish, dy, dres)
    e_y, e_dx = (y1.float() - y_ref).abs().max().item(), (dx1.float() - dx_ref).abs().max().item()
    e_y0, e_dx0 = (y0.float() - y_ref).abs().max().item(), (dx0.float() - dx_ref).abs().max().item()
    d_st = (st1 - st0).abs().max().item()
    print(f"[gn] px={px} C={c} swish={swish}: group y {e_y:.2e} dx {e_dx:.2e} | grid-barrier y {e_y0:.2e} dx {e_dx0:.2e} | "
          f"stats diff {d_st:.2e}; differing fp16 outputs: y {(y1 != y0).sum().item()} dx {(dx1 != dx0).sum().item()}")
    # fp16 outputs: half an ulp at |y| <= 8 is 4e-3; dx has |.| <= ~6
    assert e_y <= 6e-3 and e_dx <= 8e-3
    assert d_st <= 2e-5 * max(1.0, st0.abs().max().item())
    # in-place residual gradient (dres aliases dx), as the engine calls it
    lib = _lib.load()
    buf = dres.clone()
    st = st1.clone()
    rc = lib.pxr_test_groupnorm(1, _ptr(x), None, 0, None, None, None, _ptr(gamma), _ptr(beta), px, c, swish,
                                _ptr(torch.empty_like(x)), _ptr(st), _ptr(dy), None, 0, _ptr(buf), _ptr(buf),
                                _ptr(torch.zeros(64 * 200, device="cuda")), 1)
    assert rc == 0 and torch.equal(buf, dx1)


@pytest.mark.parametrize("px,c,splits", [(256, 512, 9), (1024, 256, 4), (1024, 512, 2), (256, 256, 18), (4096, 256, 3)])
def test_group_kernel_as_the_split_k_epilogue(px, c, splits):
    g = torch.Generator(device="cuda").manual_seed(7 * px + c + splits)
    ws = torch.randn(splits, px, c, generator=g, device="cuda") * 0.6
    bias = 0.3 * torch.randn(c, generator=g, device="cuda")
    res = torch.randn(px, c, generator=g, device="cuda").half()
    gamma = 1 + 0.2 * torch.randn(c, generator=g, device="cuda")
    beta = 0.2 * torch.randn(c, generator=g, device="cuda")
    ws_dy = torch.randn(splits, px, c, generator=g, device="cuda") * 0.4
    dres = (torch.randn(px, c, generator=g, device="cuda") * 0.7).half()
    # what splitk_reduce writes: bias first, partials in split order, the residual last, one rounding
    acc = bias.expand(px, c).clone()
    for s in range(splits):
        acc = acc + ws[s]
    x = (acc + res.float()).half()
    accd = torch.zeros(px, c, device="cuda")
    for s in range(splits):
        accd = accd + ws_dy[s]
    dy = accd.half()
    y_f, st_f, dx_f, x_out = _run(1, None, gamma, beta, 1, dres=dres, ws=ws, bias=bias, res=res, ws_dy=ws_dy)
    assert torch.equal(x_out, x), "the reduced tensor must equal splitk_reduce's, bit for bit"
    y_u, st_u, dx_u, _ = _run(1, x, gamma, beta, 1, dy, dres)
    assert torch.equal(y_f, y_u) and torch.equal(dx_f, dx_u) and torch.equal(st_f, st_u), "fused == unfused, bit for bit"
    y_ref, dx_ref = _reference(x, gamma, beta, 1, dy, dres)
    e_y, e_dx = (y_f.float() - y_ref).abs().max().item(), (dx_f.float() - dx_ref).abs().max().item()
    print(f"[gn split-K] px={px} C={c} splits={splits}: y {e_y:.2e} dx {e_dx:.2e}")
    assert e_y <= 6e-3 and e_dx <= 8e-3


def test_groupnorm_timing_report():
    """Per-call time of the two variants at the small-spatial decoder shapes (CUDA events over `rep` back-to-back pairs)."""
    for px, c in [(256, 512), (1024, 256), (4096, 256), (16384, 256)]:
        g = torch.Generator(device="cuda").manual_seed(1)
        x = torch.randn(px, c, generator=g, device="cuda").half()
        dy = torch.randn(px, c, generator=g, device="cuda").half()
        gamma, beta = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
        out = []
        for variant in (0, 1):
            _run(variant, x, gamma, beta, 1, dy, None, repeat=3)
            rep = 50
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            _run(variant, x, gamma, beta, 1, dy, None, repeat=rep)
            t1.record()
            torch.cuda.synchronize()
            out.append(t0.elapsed_time(t1) * 1e3 / rep)
        print(f"[gn timing] px={px} C={c}: fwd+bwd pair {out[0]:.1f} us grid-barrier, {out[1]:.1f} us cluster-per-group")
