#!/usr/bin/env python3
"""Per-instruction view of an `ncu --set full --import-source on` capture: tools/ncu_hot.py report.ncu-rep [mode]

mode "ops" (default): executed warp instructions per opcode, with the share of the kernel's total and the stall samples
that landed on them; mode "regions": the same per run of addresses with equal execution count (basic-block-like), so
that rarely taken paths show up with their real weight; mode "top": the 40 instructions with the most stall samples."""
import collections
import csv
import io
import re
import signal
import subprocess
import sys

signal.signal(signal.SIGPIPE, signal.SIG_DFL)      # `... | head` is the usual way to read this


def rows(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    lines = raw.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    rd = csv.DictReader(io.StringIO("\n".join(lines[start:])))
    out = []
    for r in rd:
        try:
            out.append((int(r["Address"], 16), r["Source"].strip(), int(r["Instructions Executed"]), int(r["# Samples"]), r))
        except (ValueError, KeyError):
            pass
    return out


def opcode(src):
    s = re.sub(r"^@!?U?P\d+\s+", "", src)
    return s.split()[0].rstrip(";")


def main():
    rep = sys.argv[1]
    mode = sys.argv[2] if len(sys.argv) > 2 else "ops"
    rs = rows(rep)
    total = sum(r[2] for r in rs)
    samples = sum(r[3] for r in rs)
    print("instructions executed %d, stall samples %d" % (total, samples))
    if mode == "ops":
        c = collections.Counter()
        s = collections.Counter()
        for _, src, n, smp, _ in rs:
            c[opcode(src)] += n
            s[opcode(src)] += smp
        for op, n in c.most_common(45):
            print("%-22s %12d %6.2f %%   samples %6.2f %%" % (op, n, 100.0 * n / total, 100.0 * s[op] / max(samples, 1)))
    elif mode == "regions":
        base = rs[0][0]
        i = 0
        while i < len(rs):
            j = i
            while j + 1 < len(rs) and abs(rs[j + 1][2] - rs[i][2]) <= 0.02 * max(rs[i][2], 1):
                j += 1
            n = sum(r[2] for r in rs[i:j + 1])
            smp = sum(r[3] for r in rs[i:j + 1])
            if n > 0.002 * total:
                print("%05x-%05x  %4d instr  x %10d  = %6.2f %% of executed, %6.2f %% of samples" % (rs[i][0] - base, rs[j][0] - base, j - i + 1, rs[i][2], 100.0 * n / total, 100.0 * smp / max(samples, 1)))
            i = j + 1
    elif mode == "top":
        base = rs[0][0]
        for a, src, n, smp, r in sorted(rs, key=lambda r: -r[3])[:40]:
            why = sorted(((int(v), k) for k, v in r.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit()), reverse=True)[:2]
            print("%05x %-60s x %9d  samples %5d  %s" % (a - base, src[:60], n, smp, why))


if __name__ == "__main__":
    main()
