#!/usr/bin/env python
"""Executed warp instructions and stall samples by opcode from the source page of an .ncu-rep (ncu --set full
--import-source on):  python profiles/opcode_mix.py gpurun_out/x.ncu-rep "title" > profiles/rNN_sass_opcode_mix_x.txt"""
import collections
import csv
import io
import re
import subprocess
import sys


def main(path, title):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = next(r for r in rows if r and r[0] == "Address")
    i_src, i_exec, i_stall = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    ex, st = collections.Counter(), collections.Counter()
    for r in rows:
        if len(r) <= i_exec or not r[0].startswith("0x"):
            continue
        m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[i_src])
        if not m:
            continue
        op = m.group(1)
        op = ".".join(op.split(".")[:2]) if op.startswith(("MUFU", "LDTM", "STTM", "SYNCS")) else op.split(".")[0]
        ex[op] += int(r[i_exec] or 0)
        st[op] += int(r[i_stall] or 0)
    tot = sum(ex.values())
    print(f"# {title}: executed warp instructions by opcode (ncu source page), total {tot}")
    for op, n in ex.most_common(28):
        print(f"{op:34s} {n:14d} {100.0 * n / tot:6.2f}%  stall samples {st[op]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
