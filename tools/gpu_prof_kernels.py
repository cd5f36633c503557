"""Small driver for `ncu --set full`: a few launches of the dominant kernels at the baseline shapes (PROF_REPS launches each,
default 2)."""
import os
import sys
import torch
sys.path.insert(0, ".")
from musev_b200 import ops

dev = "cuda"
REPS = int(os.environ.get("PROF_REPS", "2"))
torch.manual_seed(0)
# level-0 spatial self attention: (272, 4096, 8192, 40)
NF, T, Nq, heads, d, dp = 34, 17, 4096, 8, 40, 48
M = NF * Nq
hd = heads * dp
qkv = torch.randn(M, 3 * hd, device=dev).half()
qkv[:, 2 * hd:].view(M, heads, dp)[:, :, d] = 1.0
segs = [dict(k=qkv[:, hd:2 * hd], v=qkv[:, 2 * hd:], nk=Nq, fdiv=1, fmul=Nq, fadd=0),
        dict(k=qkv[:, hd:2 * hd], v=qkv[:, 2 * hd:], nk=Nq, fdiv=T, fmul=T * Nq, fadd=0)]
out = torch.empty(M, heads * d, device=dev, dtype=torch.half)
for _ in range(REPS):
    ops.attention(qkv[:, :hd], segs, NF, Nq, heads, d, dp, d ** -0.5, out=out, v_ones_col=True)
# GEGLU up-projection at level 0: M=139264, K=320, N=2560 packed
a = torch.randn(1, 1, M, 320, device=dev).half()
w = (torch.randn(2560, 320, device=dev) / 320 ** 0.5).half()
b = torch.randn(2560, device=dev)
for _ in range(REPS):
    ops.conv_gemm(a, w, bias=b, geglu=True)
# 3x3 conv 320->320 at 64x64 with residual
x = torch.randn(NF, 64, 64, 320, device=dev).half()
wc = (torch.randn(320, 2880, device=dev) / 2880 ** 0.5).half()
res = torch.randn(M, 320, device=dev).half()
for _ in range(REPS):
    ops.conv_gemm(x, wc, taps=ops.TAPS_3X3, bias=b[:320].contiguous(), residual=res)
# QKV projection K=320 -> N=1152
wq = (torch.randn(1152, 320, device=dev) / 320 ** 0.5).half()
for _ in range(REPS):
    ops.conv_gemm(a, wq)
# out-projection K=320 -> N=320 with residual
wo = (torch.randn(320, 320, device=dev) / 320 ** 0.5).half()
for _ in range(REPS):
    ops.conv_gemm(a, wo, bias=b[:320].contiguous(), residual=res)
# GroupNorm (4-D and 5-D statistics) and LayerNorm at level 0
g = torch.randn(320, device=dev); be = torch.randn(320, device=dev)
for _ in range(REPS):
    ops.groupnorm(x.view(NF, 4096, 320), g, be, groups=32, frames_per_stat=1, silu=True)
    ops.groupnorm(x.view(NF, 4096, 320), g, be, groups=32, frames_per_stat=17, silu=True)
    ops.layernorm(res, g, be, 1e-5)
# temporal attention at level 0
qkv_t = torch.randn(2 * 17 * 4096, 3 * 8 * 48, device=dev).half()
for _ in range(REPS):
    ops.temporal_attention(qkv_t, 2, 17, 4096, 8, 40, 48, 40 ** -0.5)
torch.cuda.synchronize()
print("done")
