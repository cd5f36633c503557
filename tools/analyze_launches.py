"""Joins an ncu `gpu__time_duration.sum` launch list with the MVB_TRACE shape lines of the same run.

usage: python tools/analyze_launches.py launches.csv run.log [out.txt]
"""
import csv, io, re, sys, collections

csv_path, log_path = sys.argv[1], sys.argv[2]
out_path = sys.argv[3] if len(sys.argv) > 3 else None
lines = [l for l in open(csv_path) if not l.startswith("==")]
rows = []
for row in csv.DictReader(io.StringIO("".join(lines))):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e6 if u == "ns" else v / 1e3 if u == "us" else v
    rows.append((row["Kernel Name"].split("(")[0].replace("void ", "").replace("mvb::", ""), v))
trace = [l.strip().replace("MVB_TRACE ", "") for l in open(log_path) if l.startswith("MVB_TRACE")]
ti = 0
agg = {}
bycat = collections.defaultdict(float)
for kn, t in rows:
    key = kn
    if kn.startswith("conv_gemm") or kn.startswith("attention_"):
        want = "gemm" if kn.startswith("conv_gemm") else "attn"
        while ti < len(trace) and not trace[ti].startswith(want):
            ti += 1
        if ti < len(trace):
            key = re.sub(r" tiles=\d+", "", trace[ti])
            ti += 1
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += t
    bycat[kn.split("<")[0]] += t
tot = sum(v[1] for v in agg.values())
out = [f"total {tot:.2f} ms over {len(rows)} launches"]
for k, v in sorted(bycat.items(), key=lambda x: -x[1]):
    out.append(f"  {v:8.2f} ms {100 * v / tot:5.1f}%  {k}")
out.append("by shape:")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:60]:
    m = re.search(r"gemm M=(\d+) N=(\d+) K=(\d+)", k)
    tf = ""
    if m:
        M, N, K = map(int, m.groups())
        tf = f"{2 * M * N * K * n / t / 1e9:7.1f} TF/s"
    out.append(f"{t:8.3f} ms {100 * t / tot:5.1f}% n={n:3d} {tf}  {k}")
print("\n".join(out))
if out_path:
    open(out_path, "w").write("\n".join(out) + "\n")
