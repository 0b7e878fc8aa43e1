#!/usr/bin/env python
"""ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X.csv`) -> profiles/<name>.md: per kernel launches,
total time, share.  usage: python tools/launch_summary.py gpurun_out/r2c_launches.csv profiles/r2_launches.md "<title>" """
import collections
import csv
import sys


def main(src, dst, title):
    rows = list(csv.reader(open(src, errors="ignore")))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:120]
        v, unit = float(r[ix["Metric Value"]].replace(",", "")), r[ix["Metric Unit"]]
        agg[name][0] += 1
        agg[name][1] += v / 1e6 if unit.startswith("n") else v / 1e3 if unit.startswith("u") else v
    tot, n = sum(v[1] for v in agg.values()), sum(v[0] for v in agg.values())
    ours = lambda k: not any(t in k for t in ("at::", "native::", "cutlass", "cublas", "std::enable_if", "dot_kernel", "gemv", "epilogue::",
                                              "nccl", "internal::", "Kernel2", "elementwise", "reduce_kernel"))
    lib = sum(v[1] for k, v in agg.items() if ours(k))
    with open(dst, "w") as fh:
        fh.write(f"# {title}\n\n{n} kernel launches, {tot:.1f} ms of kernel time (serialised, cold-cache under ncu: use the shares), "
                 f"{100 * lib / tot:.1f} % of it in libb3d kernels, {sum(v[0] for k, v in agg.items() if not ours(k))} library (torch / cuBLAS) launches.\n\n"
                 "| kernel | launches | total ms | share |\n|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
            fh.write(f"| `{k}` | {v[0]} | {v[1]:.3f} | {100 * v[1] / tot:.1f} % |\n")
    print("wrote", dst)


if __name__ == "__main__":
    main(*sys.argv[1:4])
