#!/usr/bin/env python
"""Per-kernel share of the step from an ncu launch list (`--metrics gpu__time_duration.sum --csv`), for comparison with
the CUDA-event split bench.py reports in roofline.kernel_share_of_step (ncu times are cold-cache and serialised: the
SHARES must agree, not the absolute values).

    python tools/launch_share.py profiles/r1z_launches.csv [kernels-per-step=11]"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    per_step = int(sys.argv[2]) if len(sys.argv) > 2 else 11
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    seq = []
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        unit = r[ui]
        us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
        name = r[ki].split("(")[0].split("::")[-1]
        seq.append((name, us))
    # the detect+demod step is the repeating pattern (corr, peak) x chunks + demod: take whole steps only
    step_names = [n for n, _ in seq[:per_step]]
    steps = 0
    tot = collections.OrderedDict()
    i = 0
    while i + per_step <= len(seq) and [n for n, _ in seq[i:i + per_step]] == step_names:
        for n, us in seq[i:i + per_step]:
            tot[n] = tot.get(n, 0.0) + us
        steps += 1
        i += per_step
    total = sum(tot.values())
    print(f"{steps} whole steps of {per_step} launches; {total / max(steps, 1) / 1e3:.3f} ms per step under ncu")
    for n, us in tot.items():
        print(f"  {n:24s} {us / max(steps, 1) / 1e3:8.3f} ms/step  share {us / total:6.3f}")
    rest = collections.Counter(n for n, _ in seq[i:])
    if rest:
        print("  launches after the steps (profiling pass, e2e pipeline):", dict(rest))


if __name__ == "__main__":
    main()
