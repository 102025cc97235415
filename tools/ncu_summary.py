#!/usr/bin/env python3
"""Transpose `ncu --page raw --csv` of a .ncu-rep into metric,unit,<one column per captured launch>.

Usage: tools/ncu_summary.py report.ncu-rep out.csv [substring ...]   (only metrics containing one of the substrings;
default: everything).  The summaries under profiles/ are made with this."""
import csv
import subprocess
import sys


def main():
    rep, out = sys.argv[1], sys.argv[2]
    subs = sys.argv[3:]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, launches = rows[0], rows[1], rows[2:]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        names = [r[hdr.index("Kernel Name")] for r in launches]
        w.writerow(["metric", "unit"] + names)
        for i, h in enumerate(hdr):
            if subs and not any(s in h for s in subs) and h not in ("Kernel Name", "Block Size", "Grid Size"):
                continue
            w.writerow([h, units[i]] + [r[i] for r in launches])


if __name__ == "__main__":
    main()
