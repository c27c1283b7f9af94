"""Summaries of the ncu CSVs that scripts/r2_ncu_evidence.sh writes (run here, on the files brought back in gpurun_out/):

    python scripts/ncu_summary.py metrics  gpurun_out/r2h_ncu_kernel_metrics.csv > profiles/r02_ncu_kernel_metrics.md
    python scripts/ncu_summary.py launches gpurun_out/r2h_launches_bench.csv     > profiles/r02_launches_bench_b256.md
"""
import collections
import csv
import re
import sys


def rows_of(path):
    lines = [l for l in open(path) if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ix = {k: hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Value", "Grid Size")}
    for r in rd:
        if len(r) > ix["Metric Value"]:
            yield r[ix["ID"]], r[ix["Kernel Name"]].replace("void ", "").replace("mode::", ""), r[ix["Metric Name"]], \
                float(r[ix["Metric Value"]].replace(",", "") or 0), r[ix["Grid Size"]]


def metrics(path):
    per = collections.OrderedDict()
    for kid, name, metric, val, _ in rows_of(path):
        per.setdefault((kid, name), {})[metric] = val
    agg = collections.OrderedDict()
    for (_, name), m in per.items():
        agg.setdefault(name, []).append(m)
    cols = [("us", "gpu__time_duration.sum", 1e-3), ("MB rd", "dram__bytes_read.sum", 1e-6), ("MB wr", "dram__bytes_write.sum", 1e-6),
            ("dram%", "dram__throughput.avg.pct_of_peak_sustained_elapsed", 1), ("L2hit%", "lts__t_sector_hit_rate.pct", 1),
            ("MB L2->SM", "l1tex__m_xbar2l1tex_read_bytes.sum", 1e-6),
            ("utchmma%", "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed", 1),
            ("GHz", "sm__cycles_elapsed.avg.per_second", 1e-9), ("regs", "launch__registers_per_thread", 1)]
    print("kernel | launches | " + " | ".join(c for c, _, _ in cols))
    for name, ms in agg.items():
        vals = [sum(m.get(k, 0.0) for m in ms) / len(ms) * s for _, k, s in cols]
        print(f"{name} | {len(ms)} | " + " | ".join(f"{v:.2f}" for v in vals))


def launches(path):
    t, n = collections.OrderedDict(), collections.Counter()
    for _, name, metric, val, _ in rows_of(path):
        if metric != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", name)
        t[name] = t.get(name, 0.0) + val
        n[name] += 1
    total = sum(t.values())
    print("kernel | launches | total us | share of the listed launches | us per launch")
    for name, v in sorted(t.items(), key=lambda kv: -kv[1]):
        print(f"{name} | {n[name]} | {v / 1e3:.1f} | {100 * v / total:.1f} % | {v / 1e3 / n[name]:.2f}")
    print(f"\ntotal {total / 1e6:.3f} ms over {sum(n.values())} launches")


if __name__ == "__main__":
    {"metrics": metrics, "launches": launches}[sys.argv[1]](sys.argv[2])
