"""Freeze what the REAL reference's settings pipeline (pixray.reset_settings / add_settings / apply_settings ->
process_args, pixray.py:2005-2102, 1824-1997) resolves a few settings dicts to, as tests/golden/api_settings.json.
Run in the authoring container (needs /root/reference; oracle/shim.py stubs the un-vendored imports):

    python oracle/make_golden_api.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import shim  # noqa: E402

shim.install()
import pixray  # noqa: E402  (the reference module itself)

CASES = {
    "defaults": dict(prompts="a cat"),
    "draft_square": dict(prompts="a cat|a dog:0.5", quality="draft", aspect="square", iterations=300),
    "best": dict(prompts="x", quality="best", aspect="square", learning_rate_drops=[50, 22.5]),
    "sized": dict(prompts="x:2:0.1", size=[256, 256], num_cuts=64, clip_models="ViT-B/16", iterations=300,
                  vector_prompts="none", save_every="50%", display_every="10 iterations"),
    "ez": dict(prompts="x", ezsize="large", aspect="portrait", drawer="fast_pixel",
               vector_prompts="textoff|textoff2:0.5", learning_rate_drops=None),
    "vdiff": dict(prompts="x", drawer="vdiff", scale=2.5, init_noise="none", custom_loss="smoothness:0.5,symmetry"),
}
KEYS = ["prompts", "clip_models", "iterations", "num_cuts", "batches", "size", "scale", "save_every", "display_every",
        "overlay_every", "overlay_offset", "overlay_until", "learning_rate_drops", "vector_prompts", "init_noise",
        "learning_rate", "drawer", "custom_loss", "quality"]

out = {}
cwd = os.getcwd()
os.chdir("/tmp")  # the reference writes scratch files relative to the cwd
for name, kw in CASES.items():
    pixray.reset_settings()
    pixray.add_settings(skip_args=True, outdir="", **kw)
    a = pixray.apply_settings()
    out[name] = {"settings": kw, "want": {k: getattr(a, k) for k in KEYS}}
os.chdir(cwd)
with open(os.path.join(ROOT, "tests", "golden", "api_settings.json"), "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)
print("wrote", len(out), "cases")
