"""Summarise ncu outputs (run here, no GPU): launch list csv -> per-kernel table; .ncu-rep -> key metrics."""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    H, data = rows[h], rows[h + 1:]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        v = float(r[vi].replace(",", ""))
        v = v / 1e6 if r[ui] == "ns" else v / 1e3 if r[ui] == "us" else v
        k = r[ki].split("(")[0]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {len(data)} launches, {tot:.3f} ms of device time (ncu-serialised, cold cache: compare SHARES)")
    print("| kernel | launches | ms | share |\n|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {v[0]} | {v[1]:.3f} | {100 * v[1] / tot:.1f}% |")


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H, U = rows[0], rows[1]
    ni = H.index("Kernel Name")
    print(f"\n# {path}")
    for r in rows[2:]:
        print(f"\n## {r[ni][:110]}")
        for i, h in enumerate(H):
            if any(h == k or h.startswith(k + ".") and h.count(".") == k.count(".") + 0 for k in KEYS) or h in KEYS:
                print(f"- {h} = {r[i]} {U[i]}")


if __name__ == "__main__":
    for p in sys.argv[1:]:
        (launches if p.endswith(".csv") else report)(p)
