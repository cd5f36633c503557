import sys, torch
sys.path.insert(0, ".")
from musev_b200 import ops
dev = "cuda"
def bench(NF, H, W, C, N, taps, iters=20):
    x = torch.randn(NF, H, W, C, device=dev).half()
    w = (torch.randn(N, len(taps) * C, device=dev) / (len(taps) * C) ** 0.5).half()
    out = torch.empty(NF * H * W, N, device=dev, dtype=torch.half)
    for _ in range(3): ops.conv_gemm(x, w, taps=taps, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): ops.conv_gemm(x, w, taps=taps, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"[bench] NF={NF} {H}x{W} C={C} N={N} taps={len(taps)}: {ms:.3f} ms {2.0*NF*H*W*N*len(taps)*C/ms/1e9:.1f} TFLOP/s", flush=True)
bench(34, 64, 64, 320, 320, ops.TAPS_3X3)
bench(34, 32, 32, 640, 640, ops.TAPS_3X3)
bench(34, 16, 16, 1280, 1280, ops.TAPS_3X3)
bench(2, 17 * 4096 // 64, 64, 320, 320, ops.TAPS_T3)
bench(1, 1, 139264, 320, 1152, ops.TAPS_1)
bench(1, 1, 139264, 320, 320, ops.TAPS_1)
bench(1, 1, 8192, 8192, 8192, ops.TAPS_1)
