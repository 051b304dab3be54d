"""Generate tests/golden/aux_loss_vectors.npz by running the REAL reference Losses/*.py (container-only).

    python -m oracle.make_golden_aux

Each entry holds one LossInterface.get_loss call of the unmodified reference class (imported under oracle/shim.py)
with its inputs, value and autograd gradient; tests/test_oracle_golden.py requires oracle/ref_path.py to reproduce
them.  EdgeLoss.get_loss builds its accumulator with `torch.tensor(0.0).cuda()` (EdgeLoss.py:75); this container has
no GPU, so Tensor.cuda is patched to the identity for that one call.  AestheticLoss.__init__ downloads its weights
(AestheticLoss.py:18-20), so the instance is created with __new__ and given a seeded nn.Linear(512, 1).
"""
import os
import types

import numpy as np
import torch
from torch import nn

from oracle import shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                   "aux_loss_vectors.npz")


def main():
    shim.install()
    from Losses.AestheticLoss import AestheticLoss
    from Losses.EdgeLoss import EdgeLoss
    from Losses.GaussianLoss import GaussianLoss
    from Losses.PaletteLoss import PaletteLoss
    from Losses.SaturationLoss import SaturationLoss
    from Losses.SmoothnessLoss import SmoothnessLoss
    from Losses.SymmetryLoss import SymmetryLoss

    G = {}
    g = torch.Generator().manual_seed(4321)
    out = torch.rand(1, 3, 20, 28, generator=g)
    cut = torch.rand(5, 3, 12, 12, generator=g) * 1.2 - 0.1
    G["out"], G["cut"] = out.numpy(), cut.numpy()

    def run(tag, loss_obj, args, use_cut, globals_=None):
        o = out.clone().requires_grad_(True)
        c = cut.clone().requires_grad_(True)
        res = loss_obj.get_loss({12: c}, o, args, globals=globals_, lossGlobals={})
        val = sum(res) if isinstance(res, (list, tuple)) else res
        val.backward()
        G[tag + "_val"] = val.detach().numpy()
        G[tag + "_grad"] = (c.grad if use_cut else o.grad).numpy()

    run("symmetry", SymmetryLoss(device="cpu"), types.SimpleNamespace(symmetry_weight=0.7), False)
    run("saturation", SaturationLoss(device="cpu"), types.SimpleNamespace(saturation_weight=1.3), True)
    palette = [[0.9, 0.1, 0.1], [0.1, 0.8, 0.2], [0.2, 0.2, 0.9], [0.95, 0.95, 0.9], [0.05, 0.05, 0.05]]
    G["palette"] = np.array(palette, dtype=np.float32)
    run("palette", PaletteLoss(device="cpu"), types.SimpleNamespace(palette=palette, palette_weight=0.8), True)
    for kind in ("default", "clipped", "log"):
        run("smooth_" + kind, SmoothnessLoss(device="cpu"),
            types.SimpleNamespace(smoothness_weight=0.9, smoothness_type=kind, smoothness_gaussian_kernel=0,
                                  smoothness_gaussian_std=1, smoothness_spacing=1, smoothness_edge_order=1), True)
    run("smooth_spacing2", SmoothnessLoss(device="cpu"),
        types.SimpleNamespace(smoothness_weight=1.0, smoothness_type="default", smoothness_gaussian_kernel=0,
                              smoothness_gaussian_std=1, smoothness_spacing=2, smoothness_edge_order=1), True)
    # EdgeLoss: parse_settings maps the colour name and the percent margins; get_loss needs .cuda() (see docstring)
    el = EdgeLoss(device="cpu")
    eargs = types.SimpleNamespace(edge_thickness=5, edge_margins=[10, 20, 15, 0], edge_color="[0.2+0.6+0.9]",
                                  edge_color_weight=0.1, global_color_weight=0.05, edge_input_image="",
                                  edge_mask_image="")
    eargs = el.parse_settings(eargs)
    G["edge_color"], G["edge_margins"] = np.array(eargs.edge_color, dtype=np.float32), np.array(eargs.edge_margins, dtype=np.int32)
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        run("edge", el, eargs, False)
    finally:
        torch.Tensor.cuda = orig_cuda
    run("gaussian", GaussianLoss(device="cpu"),
        types.SimpleNamespace(gaussian_weight=0.6, gaussian_std=(6.0, 9.0), gaussian_color=(255, 128, 0)), False)
    # AestheticLoss on unit embeddings [cutn, 512]
    ae = AestheticLoss.__new__(AestheticLoss)
    ae.device = "cpu"
    torch.manual_seed(99)
    ae.ae_reg = nn.Linear(512, 1)
    ae.target_rating = torch.ones(5, 1) * 10.0
    emb = torch.nn.functional.normalize(torch.randn(5, 512, generator=g), dim=-1).requires_grad_(True)
    val = ae.get_loss(None, None, None, globals={"embeds": emb})
    val.backward()
    G["aes_emb"], G["aes_w"], G["aes_b"] = emb.detach().numpy(), ae.ae_reg.weight.detach().numpy(), ae.ae_reg.bias.detach().numpy()
    G["aes_val"], G["aes_grad"] = val.detach().numpy(), emb.grad.numpy()

    np.savez_compressed(OUT, **G)
    print("wrote", OUT, "with", len(G), "arrays,", os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
