"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list into the markdown summary kept under profiles/.
    python tools/launch_list_md.py gpurun_out/launches_r2.csv profiles/launches_r2.md "<the command that was profiled>"
"""
import collections
import csv
import re
import sys

OURS = ("ta::", "fused_cluster_kernel", "fused_p2p_kernel", "dwconv", "dim_fwd", "dim_bwd", "aten_abs_mean", "spectrum_gemm", "adaea_drf",
        "abs_mean_kernel", "update_l2_kernel", "init_l2_kernel", "philox", "upload_tab")


def main():
    src, dst, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    idx = {h: i for i, h in enumerate(rows[hi])}
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[hi + 1:]:
        if len(r) < len(idx):
            continue
        v = float(r[idx["Metric Value"]]); u = r[idx["Metric Unit"]]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]])
        name = re.sub(r"^void ", "", name)
        tot[name] += v; cnt[name] += 1
    T = sum(tot.values()); N = sum(cnt.values())
    ours = {n for n in tot if any(k in n for k in OURS)}
    t_ours = sum(tot[n] for n in ours); n_ours = sum(cnt[n] for n in ours)
    out = ["# ncu launch list (%s)" % src.split("/")[-1].replace(".csv", ""), "", "`%s`" % cmd if cmd else "",
           "(cold-cache, serialised launches: compare SHARES, not absolutes)", "",
           "%d launches, %.1f ms of kernel time in total. Kernels of libta_b200.so: %d launches, %.3f ms = **%.2f %%**." % (N, T / 1e3, n_ours, t_ours / 1e3, 100 * t_ours / T), "",
           "| kernel | launches | total µs | share | ours |", "|---|---|---|---|---|"]
    shown = 0
    for n, v in tot.most_common():
        if shown >= 30 and n not in ours:
            continue
        out.append("| `%s` | %d | %.1f | %.2f %% | %s |" % (n[:110], cnt[n], v, 100 * v / T, "yes" if n in ours else ""))
        shown += 1
    open(dst, "w").write("\n".join(out) + "\n")
    print(dst, N, "launches; ours %.2f %%" % (100 * t_ours / T))


if __name__ == "__main__":
    main()
