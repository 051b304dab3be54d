"""Import the UNMODIFIED reference (/root/reference/pixray.py and friends) in the authoring container.

TEST INFRASTRUCTURE, container-only: /root/reference does not exist on the GPU box, so nothing that runs there
imports this module.  It is used by oracle/make_golden.py to execute the reference's own code for the in-tree parts
of the hot path and freeze the results as tests/golden/*.npz.

The reference's third-party imports that are absent here are registered as stub modules in
sys.modules; the stubs for kornia / clip / taming are backed by the restatements in oracle/ref_path.py so that the
reference's MakeCutouts (cached-transform path), CLIP_Base and VqganDrawer.synth run on top of them.
"""
import importlib
import os
import sys
import types

REFERENCE = "/root/reference"


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Anything:
    """Stands in for classes the hot path never instantiates (optimisers, samplers...)."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return self

    def __getattr__(self, k):
        return _Anything()


def install():
    if "pixray" in sys.modules:
        return sys.modules["pixray"]
    if not os.path.isdir(REFERENCE):
        raise RuntimeError("the reference checkout is only available in the authoring container")
    import torch
    import torch.nn as nn

    from oracle import ref_path as R

    _mod("braceexpand", braceexpand=lambda s: [s])
    _mod("omegaconf", OmegaConf=_Anything)
    _mod("torch_optimizer", DiffGrad=_Anything, AdamP=_Anything)
    _mod("perlin_numpy", generate_fractal_noise_2d=_Anything(), generate_fractal_noise_3d=_Anything())
    _mod("imageio")
    _mod("colorthief", ColorThief=_Anything)
    _mod("ftfy")
    _mod("resmem", ResMem=_Anything, transformer=_Anything(), path="")
    mpl = _mod("matplotlib")
    mpl.colors = _mod("matplotlib.colors", to_rgb=lambda c: (0, 0, 0))
    timm = _mod("timm", create_model=_Anything())
    timm.models = _mod("timm.models")
    timm.models.registry = _mod("timm.models.registry", register_model=lambda f: f)
    timm.models.vision_transformer = _mod("timm.models.vision_transformer", VisionTransformer=_Anything,
                                          _cfg=lambda **k: {})
    # --- clip: available_models / load backed by the restated VisionTransformer
    clip_pkg = _mod("clip")
    clip_clip = _mod("clip.clip", tokenize=_Anything(), available_models=lambda: ["ViT-B/32", "ViT-B/16"],
                     load=_Anything())
    clip_pkg.clip = clip_clip
    clip_pkg.tokenize, clip_pkg.available_models, clip_pkg.load = clip_clip.tokenize, clip_clip.available_models, clip_clip.load
    # --- kornia: only what MakeCutouts needs.  Augmentation classes are inert holders (the golden script drives the
    # cached-transform path, pixray.py:480-486, which calls kornia.geometry.transform.warp_perspective only).
    kornia = _mod("kornia")

    class _Aug(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()
            self.flags = {}

    aug = _mod("kornia.augmentation", RandomPerspective=_Aug, RandomAffine=_Aug, RandomResizedCrop=_Aug,
               ColorJitter=_Aug, CenterCrop=_Aug, RandomCrop=_Aug)
    kornia.augmentation = aug
    geo = _mod("kornia.geometry")
    tr = _mod("kornia.geometry.transform", warp_perspective=R.warp_perspective, warp_affine=_Anything(),
              rescale=_Anything())
    kornia.geometry = geo
    geo.transform = tr
    # --- taming
    taming = _mod("taming")
    taming.models = _mod("taming.models")
    taming.models.vqgan = _mod("taming.models.vqgan", VQModel=R.VQModel, GumbelVQ=_Anything)
    taming.models.cond_transformer = _mod("taming.models.cond_transformer", Net2NetTransformer=_Anything)

    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    cwd = os.getcwd()
    os.makedirs("/tmp/pixray_oracle_cwd", exist_ok=True)
    os.chdir("/tmp/pixray_oracle_cwd")  # the reference writes files into cwd (pixray.py:717-721)
    try:
        return importlib.import_module("pixray")
    finally:
        os.chdir(cwd)
