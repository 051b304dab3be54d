"""The fused attention kernels alone at the config-2 shape (64 cutouts x 12 heads, T = 197), the target for

    ncu --set full --import-source on --clock-control none -k regex:attn_ -c 4 -o gpurun_out/prof_attn python tools/ncu_attn.py
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixray_b200 import _lib  # noqa: E402

B, T, H = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (64, 197, 12)))
W = 64 * H
lib = _lib.load()
torch.manual_seed(0)
qkv = (torch.randn(B * T, 3 * W, device="cuda") * 0.5).half()
d_o = (torch.randn(B * T, W, device="cuda") * 0.1).half()
o = torch.zeros(B * T, W, dtype=torch.half, device="cuda")
lse = torch.zeros(B * H * T, dtype=torch.float32, device="cuda")
gqkv = torch.zeros(B * T, 3 * W, dtype=torch.half, device="cuda")
err = C.create_string_buffer(512)
for _ in range(2):
    rc = lib.pxr_test_attention(C.c_void_p(qkv.data_ptr()), C.c_void_p(o.data_ptr()), C.c_void_p(lse.data_ptr()),
                                C.c_void_p(d_o.data_ptr()), C.c_void_p(gqkv.data_ptr()), B, T, H, W, C.c_float(0.125), 1, err, 512)
    assert rc == 0, err.value.decode()
torch.cuda.synchronize()
print("done")
