#!/usr/bin/env python
"""Executed warp instructions per CUDA source line from an .ncu-rep captured with --import-source on (-lineinfo build):
    python profiles/hot_lines.py gpurun_out/x.ncu-rep [top_n] > profiles/rNN_hot_lines_x.txt"""
import csv
import io
import subprocess
import sys


def main(path, top=40):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True,
                         text=True).stdout
    cur, agg, kernel = None, [], None
    for r in csv.reader(io.StringIO(raw)):
        if r and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r and r[0] == "Function Name" and kernel is None:
            kernel = r[1]
        elif len(r) > 8 and r[0].isdigit() and r[2] == "-":
            try:
                agg.append((int(r[7]), int(r[4]), cur, int(r[0]), r[1].strip()[:100]))
            except ValueError:
                pass
    tot = sum(a[0] for a in agg)
    print(f"# {kernel}: executed warp instructions per CUDA line (inlined callees counted at their own line), total {tot}")
    for n, st, f, ln, src in sorted(agg, reverse=True)[:top]:
        print(f"{100.0 * n / tot:5.1f}%  stall samples {st:7d}  {f}:{ln}  {src}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
