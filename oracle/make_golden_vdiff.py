"""Generate tests/golden/vdiff_vectors.npz from the REAL reference vdiff code (container-only):

    python -m oracle.make_golden_vdiff

`/root/reference/v-diffusion-pytorch/diffusion` imports unmodified (no stubs needed).  The cc12m_1 U-Net has 603 M
parameters, so instead of its weights the fixture holds what a seeded instance computes: CC12M1Model() built under
torch.manual_seed(0) (the reference's own initialisers), its output v and the gradient of a random projection of v
w.r.t. x at 64 x 64, plus the sampler pieces (schedule, alpha/sigma, sample_step_noise).  oracle/ref_path.py must
reproduce all of them from the same seed (tests/test_oracle_golden.py).
"""
import os
import sys

import numpy as np
import torch

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "vdiff_vectors.npz")


def main():
    sys.path.insert(0, "/root/reference/v-diffusion-pytorch")
    from diffusion import get_model, sampling, utils

    G = {}
    torch.manual_seed(0)
    model = get_model("cc12m_1")().eval().requires_grad_(False)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 3, 64, 64, generator=g).requires_grad_(True)
    t = torch.tensor([0.7])
    ce = torch.randn(1, 512, generator=g)
    v = model(x, t, ce)
    w = torch.randn(v.shape, generator=g)
    (v * w).sum().backward()
    G["x"], G["t"], G["clip_embed"], G["w"] = x.detach().numpy(), t.numpy(), ce.numpy(), w.numpy()
    G["v"], G["dx"] = v.detach().numpy(), x.grad.numpy()
    sd = model.state_dict()
    G["n_tensors"] = np.int64(len(sd))
    G["n_params"] = np.int64(sum(p.numel() for p in sd.values()))
    # a few weight slices pin the seeded initialisation order
    for k in ("net.0.main.0.weight", "net.4.main.1.skip.weight", "net.4.main.5.main.5.main.5.main.2.qkv_proj.weight",
              "mapping.0.skip.weight", "net.8.main.4.weight"):
        G["w:" + k] = sd[k].reshape(-1)[:64].numpy().copy()
    # sampler: schedule as VdiffDrawer.init_from_tensor builds it (vdiff.py:113-126), 20 iterations
    iterations = 20
    ts = torch.linspace(1.0, 0, iterations + 2)[:-1]
    steps = utils.get_spliced_ddpm_cosine_schedule(ts)
    alphas, sigmas = utils.t_to_alpha_sigma(steps)
    G["steps"], G["alphas"], G["sigmas"] = steps.numpy(), alphas.numpy(), sigmas.numpy()
    xs, pred, vv = (torch.randn(1, 3, 8, 8, generator=g) for _ in range(3))
    for i in (0, 7, iterations):
        torch.manual_seed(123)
        out = sampling.sample_step_noise(None, xs, steps, 1, {}, None, alphas, sigmas, i, pred, vv)
        G[f"renoise_{i}"] = out.numpy()
    G["rn_x"], G["rn_pred"], G["rn_v"] = xs.numpy(), pred.numpy(), vv.numpy()
    np.savez_compressed(OUT, **G)
    print("wrote", OUT, len(G), "arrays", os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
