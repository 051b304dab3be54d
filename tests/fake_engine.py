"""A recording stand-in for pixray_b200.engine.B200Engine, for the CPU tests of the host-side control flow
(pixray_b200/api.py, pixray_b200/plugins.py).  Test infrastructure only: it computes nothing -- every method logs its call and
returns tensors of the right shape; `iterate` reports a scripted loss sequence so the learning-rate-drop / auto-stop
policy can be driven.  The product has no such path (the real engine refuses to start without a CUDA device)."""
from types import SimpleNamespace

import numpy as np
import torch

from pixray_b200 import engine as E


class FakeEngine:
    loss_script = staticmethod(lambda it, n: np.full(n, 1.0 / (1 + it), dtype=np.float32))

    def __init__(self, *, drawer=E.DRAWER_VQGAN, image_hw=(256, 256), vqgan=None, grid=None, cutn=64, cut_size=224,
                 clip=(), noise_fac=0.1, seed=0, device=0, rank=0, world=1, **extra):
        self.calls = []
        self.kw = dict(drawer=drawer, image_hw=tuple(image_hw), cutn=cutn, clip=list(clip), seed=seed, grid=grid, **extra)
        self.cut_aspect = extra.get('cut_aspect', 1.0)
        self.device = torch.device("cpu")
        self.cutn, self.cut_size, self.image_hw = cutn, cut_size, tuple(image_hw)
        self.n_local = cutn
        v = E.VQGAN_F16_16384 if vqgan is None else vqgan
        if drawer == E.DRAWER_VQGAN:
            f = 2 ** (len(v["ch_mult"]) - 1)
            self.z_shape = (1, v["z_channels"], image_hw[0] // f, image_hw[1] // f)
        elif drawer == E.DRAWER_VDIFF:
            self.z_shape = (1, 3, image_hw[0], image_hw[1])
        elif drawer == E.DRAWER_FFT:
            self.z_shape = (1, 3, image_hw[0], image_hw[1] // 2 + 1, 2)
        else:
            self.z_shape = (1, 3, grid[0], grid[1])
        self.cfg = SimpleNamespace(n_levels=len(v["ch_mult"]), clip=[SimpleNamespace(**c) for c in clip])
        self.clip_dims = [c["out_dim"] for c in clip]
        self.prompts = {}
        self.image_prompts = None
        self.aux = []
        self.modules = {}

    def _log(self, name, **kw):
        self.calls.append((name, kw))

    def names(self):
        return [c[0] for c in self.calls]

    # ---- setup
    def load_module(self, module_id, state_dict):
        self.modules[module_id] = state_dict
        self._log("load_module", module=module_id, n=len(state_dict))

    def finalize(self):
        self._log("finalize")

    def set_prompts(self, clip_idx, embeds, weights, stops):
        self.prompts[clip_idx] = (np.asarray(embeds), list(weights), list(stops))
        self._log("set_prompts", clip=clip_idx, n=len(weights))

    def set_image_prompts(self, imgs, weights=None):
        self.image_prompts = (imgs, weights)
        self._log("set_image_prompts", n=0 if imgs is None else len(imgs))

    def add_filter(self, kind, weight, params):
        self.filters = getattr(self, "filters", [])
        self.filters.append((kind, weight, list(params)))
        self._log("add_filter", kind=kind, weight=weight)
        return len(self.filters) - 1

    def add_anchor(self, kind, weight, ref):
        if not hasattr(self, "anchors"):
            self.anchors = []
        self.anchors.append((kind, weight, np.asarray(ref.cpu() if hasattr(ref, "cpu") else ref, dtype=np.float32).copy()))
        self._log("add_anchor", kind=kind, weight=weight)

    def add_aux_loss(self, kind, weight, params):
        self.aux.append((kind, weight, list(params)))
        self._log("add_aux_loss", kind=kind, weight=weight)
        return self.num_losses() - 1

    def num_losses(self):
        n_img = 0 if self.image_prompts is None or self.image_prompts[0] is None else len(self.image_prompts[0])
        return sum(len(p[1]) + n_img for p in self.prompts.values()) + len(self.aux) + len(getattr(self, "filters", [])) + len(getattr(self, "anchors", []))

    def z_bounds(self):
        c = self.z_shape[1]
        return -torch.ones(c), torch.ones(c)

    def reset_optimizer(self):
        self._log("reset_optimizer")

    # ---- per-op
    def synth(self, z):
        self._log("synth", shape=tuple(z.shape))
        return torch.full((1, 3, *self.image_hw), 0.5)

    def vqgan_encode(self, img):
        self._log("vqgan_encode", shape=tuple(img.shape), lo=float(img.min()), hi=float(img.max()))
        return torch.zeros(self.z_shape)

    def make_cutouts(self, img=None, **kw):
        self._log("make_cutouts", **{k: (None if v is None else type(v).__name__) for k, v in kw.items()})
        return torch.zeros(self.cutn, 3, self.cut_size, self.cut_size)

    def encode_image(self, clip_idx=0, batch=None):
        self._log("encode_image", clip=clip_idx)
        return torch.zeros(self.cutn, self.clip_dims[clip_idx])

    def prompt_loss(self, clip_idx=0, embeds=None):
        self._log("prompt_loss", clip=clip_idx, passed_embeds=embeds is not None)
        return torch.arange(len(self.prompts[clip_idx][1]), dtype=torch.float32)

    def backward(self):
        self._log("backward")
        return torch.ones(self.z_shape)

    def set_spot_mask(self, mask):
        self.spot_mask = np.asarray(mask)
        self._log("set_spot_mask", shape=tuple(self.spot_mask.shape), frac=float(self.spot_mask.mean()))

    def set_spot_prompts(self, clip_idx, which, embeds, weights, stops):
        self._log("set_spot_prompts", clip=clip_idx, which=which, n=len(weights))

    def set_z_grad(self, g):
        self._log("set_z_grad", shape=tuple(g.shape))

    def set_batches(self, batches):
        self._log("set_batches", batches=batches)

    def step(self, z, lr, it=0):
        self._log("step", lr=lr, it=it)
        return z

    def read_losses(self):
        return np.zeros(self.num_losses(), dtype=np.float32)

    # ---- fused
    def iterate(self, z, lr, it, *, params=None, losses_out=None):
        assert tuple(z.shape) == tuple(self.z_shape) and z.is_contiguous()
        self._log("iterate", lr=lr, it=it, params=params is not None)
        if losses_out is not None:
            losses_out[:] = type(self).loss_script(it, losses_out.size)

    # ---- vdiff
    def vdiff_set_schedule(self, steps, alphas, sigmas):
        self._log("vdiff_set_schedule", n=len(steps))

    def vdiff_set_clip_embed(self, e):
        self._log("vdiff_set_clip_embed", n=int(np.asarray(e).size))

    def vdiff_set_iteration(self, i):
        self._log("vdiff_set_iteration", i=i)

    def vdiff_renoise(self, x, i, noise):
        self._log("vdiff_renoise", i=i)
        return x
