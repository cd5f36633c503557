#!/usr/bin/env python
"""Times the spatial attention kernels at the UNet's shapes (B=2 CFG, T=16+1, 64x64 latents) through the op-level C ABI.
    python tools/gpu_bench_attention.py [--variants 0 1] [--levels 0 1 2]
MVB_POLY (share of exp2 on the FMA pipe, read once per process) is taken from the environment."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pad_heads(x, heads, d, dp, ones=False):
    o = torch.zeros(x.shape[0], heads, dp, device=x.device, dtype=x.dtype)
    o[:, :, :d] = x.view(x.shape[0], heads, d)
    if ones:
        o[:, :, d] = 1.0
    return o.view(x.shape[0], heads * dp)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", type=int, nargs="*", default=[0, 1])
    ap.add_argument("--levels", type=int, nargs="*", default=[0, 1, 2])
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    from musev_b200 import ops
    dev = "cuda"
    NF, T, heads = 34, 17, 8
    shapes = {0: (4096, 40), 1: (1024, 80), 2: (256, 160)}
    for lvl in a.levels:
        Nq, d = shapes[lvl]
        dp = (d + 15) // 16 * 16
        hd = heads * dp
        M = NF * Nq
        torch.manual_seed(0)
        q, k, v = (torch.randn(M, heads * d, device=dev).half() for _ in range(3))
        ones = dp > d
        qkv = torch.cat([pad_heads(q, heads, d, dp), pad_heads(k, heads, d, dp), pad_heads(v, heads, d, dp, ones)], 1).contiguous()
        segs = [dict(k=qkv[:, hd:2 * hd], v=qkv[:, 2 * hd:], nk=Nq, fdiv=1, fmul=Nq, fadd=0),
                dict(k=qkv[:, hd:2 * hd], v=qkv[:, 2 * hd:], nk=Nq, fdiv=T, fmul=T * Nq, fadd=0)]
        outs = {}
        for var in a.variants:
            out = torch.zeros(M, heads * d, dtype=torch.float16, device=dev)
            for _ in range(3):
                ops.attention(qkv[:, :hd], segs, NF, Nq, heads, d, dp, d ** -0.5, out=out, v_ones_col=ones, variant=var)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.iters):
                ops.attention(qkv[:, :hd], segs, NF, Nq, heads, d, dp, d ** -0.5, out=out, v_ones_col=ones, variant=var)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.iters
            flop = 4.0 * NF * heads * Nq * (2 * Nq) * d
            outs[var] = out
            print("ATTN_BENCH " + json.dumps({"level": lvl, "Nq": Nq, "d": d, "variant": var, "poly": os.environ.get("MVB_POLY", "default"),
                                              "ms": ms, "tflops": flop / ms / 1e9}), flush=True)
        if len(outs) > 1:
            ks = sorted(outs)
            print("ATTN_DIFF", lvl, (outs[ks[0]].float() - outs[ks[1]].float()).abs().max().item(), flush=True)


if __name__ == "__main__":
    main()
