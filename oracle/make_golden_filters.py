"""Generate tests/golden/filter_vectors.npz by running the REAL reference filters/*.py (container-only).

    python -m oracle.make_golden_filters

Each entry holds one FilterInterface.forward call of the unmodified reference class (imported under oracle/shim.py) with
its input, the (rand_h, rand_w) the class drew (recovered by replaying torch.randint under the same seed), the output image,
the loss and the autograd gradient of  sum(out * G) + loss;  tests/test_oracle_golden.py requires oracle/ref_path.py's
restatements to reproduce them."""
import os
import types

import numpy as np
import torch

from oracle import shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "filter_vectors.npz")


def main():
    shim.install()
    from filters.colorlookup import ColorLookup
    from filters.tiler import TilerFilter
    from filters.wallpaper import WallpaperFilter

    G = {}
    g = torch.Generator().manual_seed(777)
    img = torch.rand(1, 3, 20, 28, generator=g)
    G["img"] = img.numpy()

    def run(tag, filt, seed):
        torch.manual_seed(seed)
        B, C, H, W = img.shape
        rw, rh = int(torch.randint(0, W, (1,))), int(torch.randint(0, H, (1,)))  # the draws forward() is about to make
        torch.manual_seed(seed)
        x = img.clone().requires_grad_(True)
        out, loss = filt(x)
        up = torch.rand(out.shape, generator=torch.Generator().manual_seed(seed + 1))
        total = (out * up).sum() + loss
        total.backward()
        G[tag + "_rand"] = np.array([rh, rw], dtype=np.int64)
        G[tag + "_out"] = out.detach().numpy()
        G[tag + "_loss"] = np.asarray(float(loss), dtype=np.float32)
        G[tag + "_up"] = up.numpy()
        G[tag + "_grad"] = x.grad.numpy()

    run("tiler", TilerFilter(types.SimpleNamespace(), device="cpu"), 11)
    for wt, em in (("shift", 0), ("horizontal", 0), ("horizontal", 6), ("vertical", 4), (None, 0), (None, 6)):
        s = types.SimpleNamespace(wallpaper_type=wt, wallpaper_edge_match=em)
        run(f"wallpaper_{wt}_{em}", WallpaperFilter(s, device="cpu"), 23 + em)
    pal = [[0.9, 0.1, 0.1], [0.1, 0.8, 0.2], [0.2, 0.2, 0.9], [0.95, 0.95, 0.9], [0.05, 0.05, 0.05]]
    G["palette"] = np.array(pal, dtype=np.float32)
    run("lookup", ColorLookup(types.SimpleNamespace(lookup_beta=3.0, palette=pal), device="cpu"), 5)
    run("lookup_default", ColorLookup(types.SimpleNamespace(lookup_beta=10.0, palette=None), device="cpu"), 6)
    np.savez_compressed(OUT, **G)
    print("wrote", OUT, len(G), "arrays")


if __name__ == "__main__":
    main()
