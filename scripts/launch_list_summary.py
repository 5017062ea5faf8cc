"""Per-kernel shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list.  usage: python scripts/launch_list_summary.py X.csv"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum": continue
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    tot[name][0] += 1; tot[name][1] += us
total = sum(v[1] for v in tot.values())
print(f"# {sum(v[0] for v in tot.values())} launches, {total/1e3:.2f} ms total device time (cold-cache, serialised: compare SHARES)")
print(f"{'kernel':90s} {'launches':>8s} {'total_us':>10s} {'share':>7s} {'avg_us':>8s}")
for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:90]:90s} {n:8d} {us:10.1f} {us/total*100:6.2f}% {us/n:8.2f}")
grp = collections.defaultdict(float)
for k, (n, us) in tot.items():
    g = ("cuBLAS / cutlass / ATen" if (k.startswith("nvjet") or "cublas" in k.lower() or "cutlass" in k.lower() or k.startswith("at::") or "gemv" in k.lower() or k.startswith("std::") or "internal::" in k)
         else ("own tcgen05 attention" if k.startswith("attn::") else ("own tcgen05 GEMM (ln_proj, mlp_dh, wgrad, gemm)" if k.startswith(("lnproj::", "mlp::", "proj::", "wgrad::", "gemm::")) else "own HBM-bound kernels")))
    grp[g] += us
print("# groups:")
for g, us in sorted(grp.items(), key=lambda kv: -kv[1]): print(f"#   {g:40s} {us/total*100:6.2f}%")
