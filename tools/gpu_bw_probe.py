"""What does a streaming kernel get on this box? torch copy / add / layer_norm vs the engine's LayerNorm / GroupNorm at the
level-0 tensor size (139264 x 320 fp16 = 89 MB) and 4x that."""
import sys
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from musev_b200 import ops

dev = "cuda"


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for M, C in ((139264, 320), (139264, 1280)):
    x = torch.randn(M, C, device=dev).half()
    y = torch.empty_like(x)
    nbytes = x.numel() * 2
    g = torch.ones(C, device=dev); b = torch.zeros(C, device=dev)
    gh, bh = g.half(), b.half()
    rows = [("torch copy_ (1R+1W)", lambda: y.copy_(x), 2),
            ("torch add scalar (1R+1W)", lambda: torch.add(x, 1.0, out=y), 2),
            ("torch sum (1R)", lambda: x.sum(), 1),
            ("torch layer_norm fp16 (1R+1W)", lambda: F.layer_norm(x, (C,), gh, bh, 1e-5), 2),
            ("engine layernorm (1R+1W)", lambda: ops.layernorm(x, g, b, 1e-5), 2)]
    if C == 320:
        x3 = x.view(34, 4096, C)
        rows.append(("engine groupnorm+silu (2R+1W)", lambda: ops.groupnorm(x3, g, b, 32, 1, 1e-5, True), 3))
    for name, fn, mult in rows:
        ms = timeit(fn)
        print(f"[bw] {M}x{C} {name}: {ms * 1e3:.1f} us  {mult * nbytes / ms / 1e6:.0f} GB/s", flush=True)
