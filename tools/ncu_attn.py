"""Config-2 attention shape (64 images x 12 heads x 197 tokens) through the pxr_test_attention hook: the target for
`ncu --set full -k regex:attn_` (two warm-up launches of each kernel, then the profiled ones)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixray_b200 import _lib  # noqa: E402

B, T, H = 64, 197, 12
W = 64 * H
lib = _lib.load()
qkv = torch.randn(B * T, 3 * W, device="cuda").half()
d_o = torch.randn(B * T, W, device="cuda").half()
o = torch.zeros(B * T, W, dtype=torch.half, device="cuda")
lse = torch.zeros(B * H * T, device="cuda")
gqkv = torch.zeros(B * T, 3 * W, dtype=torch.half, device="cuda")
err = C.create_string_buffer(512)
rc = lib.pxr_test_attention(C.c_void_p(qkv.data_ptr()), C.c_void_p(o.data_ptr()), C.c_void_p(lse.data_ptr()),
                            C.c_void_p(d_o.data_ptr()), C.c_void_p(gqkv.data_ptr()), B, T, H, W, C.c_float(0.125), 3, err, 512)
torch.cuda.synchronize()
print("rc", rc, err.value.decode())
