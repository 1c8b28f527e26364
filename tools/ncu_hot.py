"""Hottest SASS instructions of one kernel in an .ncu-rep by warp-stall samples, with the CUDA source line each maps to
(run here, no GPU).  Usage: python tools/ncu_hot.py report.ncu-rep [top]"""
import csv
import subprocess
import sys


def main():
    path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    if not h:
        out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        h = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    H, data = rows[h[0]], rows[h[0] + 1:]
    si, ii = H.index("Warp Stall Sampling (All Samples)"), H.index("Instructions Executed")
    srci = H.index("Source")
    tot = sum(int(r[si]) for r in data if len(r) > si and r[si].isdigit())
    print(rows[0][1][:100] if len(rows[0]) > 1 else "", "total samples", tot)
    best = sorted([(int(r[si]), i) for i, r in enumerate(data) if len(r) > si and r[si].isdigit()], reverse=True)[:top]
    for s, i in best:
        ctx = " | ".join(data[j][srci].strip()[:60] for j in range(max(0, i - 2), min(len(data), i + 2)))
        print(f"{100 * s / tot:5.1f}%  #{i:5d} x{data[i][ii]:>8s}  {ctx}")


if __name__ == "__main__":
    main()
