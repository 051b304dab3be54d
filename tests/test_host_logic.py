"""CPU: host-side logic of the drop-in -- cutout parameter sampling, synthetic weights, FLOP accounting."""
import numpy as np
import torch

from oracle import ref_path as R
from pixray_b200 import cutouts, synthetic
from pixray_b200.engine import CLIP_ARCH, VQGAN_F16_16384


def test_transform_sampler_groups_and_invertibility():
    T = cutouts.sample_transforms(64, 224, 0)
    assert T.shape == (64, 3, 3) and np.isfinite(T).all()
    assert all(abs(np.linalg.det(T[n].astype(np.float64))) > 1e-3 for n in range(64))
    zoom = int(0.6 * 64)  # pixray.py:407
    assert zoom == 38
    # zoom group magnifies (RandomResizedCrop area in [0.25, 0.95]), wide group shrinks by 0.95
    corners = np.array([[0, 0, 1], [223, 0, 1], [223, 223, 1], [0, 223, 1]], dtype=np.float64).T

    def area_scale(H):
        p = H.astype(np.float64) @ corners
        p = (p[:2] / p[2]).T
        x, y = p[:, 0], p[:, 1]
        return 0.5 * abs(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1))) / 223.0 ** 2

    z = np.array([area_scale(T[n]) for n in range(zoom)])
    w = np.array([area_scale(T[n]) for n in range(zoom, 64)])
    assert z.mean() > 1.0 and w.mean() < 1.0


def test_transform_sampler_statistics():
    """Distribution check: the crop-area fraction of the zoom group is ~U(0.25, 0.95)
    when no perspective is applied (30 % of draws)."""
    fr = []
    for seed in range(40):
        g = np.random.default_rng(seed)
        H = cutouts._sample_resized_crop(g, 224)
        # crop box = preimage of the output corners
        Hi = np.linalg.inv(H)
        p = Hi @ np.array([[0, 0, 1], [223, 223, 1]], dtype=np.float64).T
        p = p[:2] / p[2]
        fr.append(((p[0, 1] - p[0, 0] + 1) * (p[1, 1] - p[1, 0] + 1)) / 224.0 ** 2)
    fr = np.array(fr)
    assert 0.2 < fr.min() and fr.max() < 1.0 and 0.45 < fr.mean() < 0.75


def test_synthetic_state_dicts_load_into_the_oracle_models():
    vq = R.VQModel()
    assert not vq.load_state_dict(synthetic.vqgan_state_dict(VQGAN_F16_16384, 0), strict=True).missing_keys
    clip = R.ClipVisual(224, 32, 768, 12, 12, 512)
    assert not clip.load_state_dict(synthetic.clip_state_dict(CLIP_ARCH["ViT-B/32"], 0), strict=True).missing_keys


def test_flop_accounting_matches_survey():
    assert abs(synthetic.vit_fwd_flops(CLIP_ARCH["ViT-B/16"]) / 1e9 - 35.13) < 0.02
    assert abs(synthetic.vit_fwd_flops(CLIP_ARCH["ViT-B/32"]) / 1e9 - 8.82) < 0.02
    assert abs(synthetic.vqgan_decoder_fwd_flops(VQGAN_F16_16384, (256, 256)) / 1e9 - 252.7) < 0.2
    assert abs(synthetic.vqgan_decoder_fwd_flops(VQGAN_F16_16384, (512, 512)) / 1e9 - 1017.4) < 6.0


def test_prompts_are_unit_norm_and_seeded():
    a, b = synthetic.prompts(512, (1.0, 0.1), 5), synthetic.prompts(512, (1.0, 0.1), 5)
    assert torch.equal(a[0][0], b[0][0]) and abs(a[1][0].norm().item() - 1) < 1e-5


def test_loss_plugin_settings_match_the_reference():
    """pixray_b200.losses mirrors Losses/*.py's add_settings: same option names, defaults, types and nargs
    (tests/golden/loss_settings.json was dumped from the reference classes' own parsers under oracle/shim.py)."""
    import argparse
    import json
    import os

    from pixray_b200 import losses as L
    want = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "loss_settings.json")))
    for name, ref in want.items():
        p = getattr(L, name).add_settings(argparse.ArgumentParser())
        got = {a.dest: {"default": list(a.default) if isinstance(a.default, (tuple, list)) else a.default,
                        "type": getattr(a.type, "__name__", None), "nargs": a.nargs} for a in p._actions if a.dest != "help"}
        # the one engine-side extra: where the reference downloads the AVA head, this package takes it as a setting
        got.pop("aesthetic_head", None)
        assert got == ref, name
    assert set(L.loss_class_table) >= {"palette", "saturation", "symmetry", "smoothness", "edge", "aesthetic"}


def test_edge_loss_parse_settings_follows_the_reference():
    import types

    from pixray_b200 import losses as L
    a = types.SimpleNamespace(edge_thickness=7, edge_margins=None, edge_color="(255+128+0)", edge_input_image="", edge_mask_image="")
    a = L.EdgeLoss(device=None).parse_settings(a)
    assert a.edge_margins == (7, 7, 7, 7)  # EdgeLoss.py:33-35
    assert [round(v, 6) for v in a.edge_color] == [1.0, round(128 / 255, 6), 0.0]  # util.parse_triple_to_rgb


def test_vdiff_schedule_helper_matches_the_oracle():
    import numpy as np

    from oracle import ref_path as R
    from pixray_b200 import util as U
    for n, skip in ((20, 0.0), (300, 0.0), (50, 25.0)):
        s, a, g = U.vdiff_schedule(n, skip)
        rs, ra, rg = R.vdiff_schedule(n, skip)
        assert np.allclose(s, rs.numpy(), atol=2e-6) and np.allclose(a, ra.numpy(), atol=2e-6) and np.allclose(g, rg.numpy(), atol=2e-6)
