#!/usr/bin/env python
"""Is a kernel bound by its arithmetic / latency chains or by HBM?  Times the three kernels of the normal-burst step on a
batch that streams from HBM (2^20 bursts, 5.2 GB) and on one that stays resident in the 126 MB L2 (the same bursts, a
whole number of waves of the launch geometry), per burst.  A kernel whose per-burst time does not drop when its input
comes from L2 is not waiting for HBM.   python tools/l2_probe.py [--workload nb]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="nb")
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    import bench
    import osmo_trx_b200
    trx = osmo_trx_b200.Trx(0)
    n_big = 1 << 20
    rx, typ, tsc, mt, bound = bench.make_workload(trx, args.workload, n_big, seed=3, device=trx.device)
    trx.detect_config(40 if args.workload == "rach" else 16, 2 if args.workload == "edge" else 1)
    sms = torch.cuda.get_device_properties(trx.device).multi_processor_count
    out = {}
    for name, n in (("hbm", n_big), ("l2", sms * 16 * 7 * 4 // 7 * 1), ("l2x2", sms * 16 * 8)):
        n = int(n)
        res = trx.alloc_results(n, 148)
        r = dict(rx=rx[:n], typ=typ[:n], tsc=tsc[:n], mt=mt[:n])
        for _ in range(5):
            trx.detect_demod(r["rx"], r["typ"], r["tsc"], r["mt"], bound, n_gmsk_soft=148, out=res)
        torch.cuda.synchronize()
        reps = 20 if n == n_big else 200
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            trx.detect_demod(r["rx"], r["typ"], r["tsc"], r["mt"], bound, n_gmsk_soft=148, out=res)
        b.record()
        torch.cuda.synchronize()
        step_ms = a.elapsed_time(b) / reps
        trx.profile_begin()
        for _ in range(reps):
            trx.detect_demod(r["rx"], r["typ"], r["tsc"], r["mt"], bound, n_gmsk_soft=148, out=res)
        pr = trx.profile_end()
        row = {"bursts": n, "MB": n * 5000 / 1e6, "step_ns_per_burst": 1e6 * step_ms / n}
        for k, v in pr.items():
            row[k + "_ns_per_burst"] = 1e6 * v[0] / reps / n
        out[name] = row
        print(name, json.dumps(row))
    if args.json:
        json.dump(out, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
