#!/usr/bin/env python
"""Hottest SASS instructions (by warp-stall samples) of one kernel of an ncu report.

    python tools/ncu_hot.py gpurun_out/prof.ncu-rep conv1_kernel [N]
"""
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern],
                         capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(raw.splitlines())]
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h]
    idx = {k: i for i, k in enumerate(hdr)}
    data = [r for r in rows[h + 1:] if len(r) == len(hdr) and r[0].startswith("0x")]
    # keep only the first kernel instance (the page repeats per profiled launch)
    seen, first = set(), []
    for r in data:
        if r[0] in seen:
            break
        seen.add(r[0])
        first.append(r)
    data = first
    tot = sum(int(r[idx["# Samples"]]) for r in data)
    print(f"kernel {kern}: {len(data)} instructions, {tot} samples")
    stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    agg = {k: sum(int(r[idx[k]]) for r in data) for k in stalls}
    print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
    for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]]))[:n]:
        s = {k: int(r[idx[k]]) for k in stalls if int(r[idx[k]]) > 0}
        best = sorted(s.items(), key=lambda kv: -kv[1])[:2]
        print(f"{int(r[idx['# Samples']]):6d} x{r[idx['Instructions Executed']]:>8s}  {r[idx['Source']].strip()[:64]:64s} {best}")


if __name__ == "__main__":
    main()
