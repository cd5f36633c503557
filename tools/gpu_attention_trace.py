#!/usr/bin/env python
"""Phase timeline of the ping-pong attention kernel at the level-0 shape: CTA (0,0,0) stamps the SM clock at every phase of its
first 32 key/value tiles (mvb_debug_attention_trace); prints the mean cycles per phase for both softmax warpgroups and the
MMA-issuing warp, and the steady-state period per KV tile.
    python tools/gpu_attention_trace.py [--out gpurun_out/attn_trace.json]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.gpu_bench_attention import pad_heads  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--variant", type=int, default=0)
    a = ap.parse_args()
    from musev_b200 import _capi, ops
    dev = "cuda"
    NF, T, heads, Nq, d = 34, 17, 8, 4096, 40
    dp = 48
    hd = heads * dp
    M = NF * Nq
    torch.manual_seed(0)
    q, k, v = (torch.randn(M, heads * d, device=dev).half() for _ in range(3))
    qkv = torch.cat([pad_heads(q, heads, d, dp), pad_heads(k, heads, d, dp), pad_heads(v, heads, d, dp, True)], 1).contiguous()
    segs = [dict(k=qkv[:, hd:2 * hd], v=qkv[:, 2 * hd:], nk=Nq, fdiv=1, fmul=Nq, fadd=0),
            dict(k=qkv[:, hd:2 * hd], v=qkv[:, 2 * hd:], nk=Nq, fdiv=T, fmul=T * Nq, fadd=0)]
    out = torch.zeros(M, heads * d, dtype=torch.float16, device=dev)

    def run():
        ops.attention(qkv[:, :hd], segs, NF, Nq, heads, d, dp, d ** -0.5, out=out, v_ones_col=True, variant=a.variant)

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    buf = torch.zeros(10 * 32 * 8, dtype=torch.int64, device=dev)
    _capi.check(_capi.lib().mvb_debug_attention_trace(buf.data_ptr()))
    run()
    torch.cuda.synchronize()
    _capi.check(_capi.lib().mvb_debug_attention_trace(None))
    tr = buf.cpu().view(10, 32, 8)
    t0 = int(tr[tr > 0].min())
    rel = torch.where(tr > 0, tr - t0, torch.zeros_like(tr))
    lo, hi = 4, 30                                    # steady state
    res = {}
    names = ["wait_S", "tmem_load", "row_max", "exp", "wait_prev_PV", "store_P"]
    for t in range(2):
        for q in range(4):                            # lane quarter q runs on SM sub-partition q (warp index % 4)
            r = rel[4 * t + q].double()
            e = {n: round(float((r[lo:hi, i + 1] - r[lo:hi, i]).mean())) for i, n in enumerate(names)}
            e["period"] = round(float((r[lo + 1:hi + 1, 1] - r[lo:hi, 1]).mean()))
            e["busy"] = round(float((r[lo:hi, 6] - r[lo:hi, 1]).mean()))
            res[f"softmax_t{t}_q{q}"] = e
    two = bool((rel[9] != 0).any())                   # stamps of a second MMA-issuing warp (an experiment of round 2; the shipped kernel has one)
    for t in range(2):
        m = rel[8 + t].double() if two else rel[8].double()
        b = 0 if two else 4 * t
        prev = m[lo - 1:hi - 1, 3] if two else (m[lo:hi, b - 1] if t == 1 else m[lo - 1:hi - 1, 7])
        e = {"wait_S_free": round(float((m[lo:hi, b] - prev).mean())), "issue_S": round(float((m[lo:hi, b + 1] - m[lo:hi, b]).mean())),
             "wait_V_and_P": round(float((m[lo:hi, b + 2] - m[lo:hi, b + 1]).mean())), "issue_PV": round(float((m[lo:hi, b + 3] - m[lo:hi, b + 2]).mean()))}
        pubs = torch.stack([rel[4 * t + q, lo:hi, 6] for q in range(4)])
        e["P_last_publish_to_seen"] = round(float((m[lo:hi, b + 2] - pubs.max(0).values.double()).mean()))
        e["P_publish_spread_over_quarters"] = round(float((pubs.max(0).values - pubs.min(0).values).double().mean()))
        last_free = torch.stack([rel[4 * t + q, lo:hi, 2] for q in range(4)]).max(0).values.double()
        e["S_last_freed_to_next_S_issued"] = round(float((m[lo:hi, b + 1] - last_free).mean()))
        res[f"mma_tile{t}"] = e
    print("ATTN_TRACE " + json.dumps(res), flush=True)
    for j in range(6, 9):
        print("ATTN_TRACE_RAW j=%d" % j, flush=True)
        for r in range(10):
            print("ATTN_TRACE_RAW   role %d %s" % (r, rel[r, j].tolist()), flush=True)
    if a.out:
        with open(a.out, "w") as fh:
            json.dump({"summary": res, "stamps": rel.tolist()}, fh)


if __name__ == "__main__":
    main()
