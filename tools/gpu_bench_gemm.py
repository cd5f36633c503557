#!/usr/bin/env python
"""Times conv_gemm at chosen linear shapes through the op-level C ABI (L2 flushed between iterations by rotating buffers).
    python tools/gpu_bench_gemm.py --shapes 139264,320,320,1 139264,1152,320,0 34816,640,640,1
shape = M,N,K,residual"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", nargs="*", default=["139264,320,320,1", "139264,1152,320,0", "34816,640,640,1", "139264,2560,320,2"])
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    from musev_b200 import ops
    dev = "cuda"
    for sh in a.shapes:
        M, N, K, mode = (int(x) for x in sh.split(","))
        nbuf = max(2, int(400e6 // (M * K * 2)) + 1)          # rotate inputs so that A / residual never come from L2
        A = [torch.randn(1, 1, M, K, device=dev).half() for _ in range(nbuf)]
        W = (torch.randn(N, K, device=dev) / K ** 0.5).half()
        b = torch.randn(N, device=dev)
        geglu = mode == 2
        nout = N // 2 if geglu else N
        R = [torch.randn(M, nout, device=dev).half() for _ in range(nbuf)] if mode == 1 else None
        out = torch.empty(M, nout, dtype=torch.float16, device=dev)

        def run(i):
            ops.conv_gemm(A[i % nbuf], W, bias=b, residual=R[i % nbuf] if R else None, geglu=geglu, out=out)
        for i in range(3):
            run(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(a.iters):
            run(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        byt = M * K * 2 + M * nout * 2 * (2 if mode == 1 else 1) + N * K * 2
        print("GEMM_BENCH " + json.dumps({"M": M, "N": N, "K": K, "mode": mode, "ms": ms, "tflops": 2.0 * M * N * K / ms / 1e9,
                                          "algorithmic_GBps": byt / ms / 1e6}), flush=True)


if __name__ == "__main__":
    main()
