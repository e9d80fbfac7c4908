#!/usr/bin/env python
"""Run one case of the hot path a few times, for `ncu` captures (never a bench number):

    ncu --set full --import-source on --clock-control none -k regex:demod -c 1 -o gpurun_out/x \\
        python tools/prof_case.py --case edge [--n 65536] [--reps 2]

cases: nb | rach | edge (fused detect+demod), vitac, pull, convolve, delay, resamp, chan4, chan64"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="nb")
    ap.add_argument("--n", type=int, default=1 << 16)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--prof", action="store_true", help="print the per-kernel CUDA-event split of one call (ms, launches)")
    ap.add_argument("--time", action="store_true", help="print the CUDA-event time per call after warm-up")
    args = ap.parse_args()
    import bench
    from osmo_trx_b200 import Trx, Resampler, Channelizer, Synthesis
    trx = Trx(0)
    dev = trx.device
    n = args.n
    if args.case in ("nb", "rach", "edge"):
        soft = 444 if args.case == "edge" else 148
        cfg = {"nb": (16, 1), "rach": (40, 1), "edge": (16, 2)}[args.case]
        rx, typ, tsc, mt, bound = bench.make_workload(trx, args.case, n, seed=7, device=dev)
        trx.detect_config(*cfg)
        res = trx.alloc_results(n, soft)
        fn = lambda: trx.detect_demod(rx, typ, tsc, mt, bound, n_gmsk_soft=148, out=res)  # noqa: E731
    elif args.case == "pull":
        rx, typ, tsc, mt, bound = bench.make_workload(trx, "nb", n, seed=7, device=dev)
        trx.detect_config(16, 1)
        iq = (rx * bench.IQ_SCALE).round().clamp(-32768, 32767).to(torch.int16)
        f = (torch.arange(n, device=dev, dtype=torch.int32) // 8)
        t = (torch.arange(n, device=dev) % 8).to(torch.uint8)
        po = trx.alloc_pull_results(n, 160)
        fn = lambda: trx.pull(iq, typ, tsc, mt, f, t, bound, out=po)  # noqa: E731
    elif args.case == "vitac":
        rx, typ, tsc, mt, bound = bench.make_workload(trx, "nb", n, seed=9, device=dev)
        buf = torch.zeros((n, 40 + 625 + 63, 2), dtype=torch.float32, device=dev)
        buf[:, 40:665] = rx
        fn = lambda: trx.vitac(buf, 40, tsc)  # noqa: E731
    elif args.case in ("convolve", "delay"):
        rx, typ, tsc, mt, bound = bench.make_workload(trx, "nb", n, seed=7, device=dev)
        if args.case == "delay":
            dl = torch.rand(n, device=dev) * 8 - 4
            fn = lambda: trx.delay_vector(rx, dl)  # noqa: E731
        else:
            hc = torch.randn((16, 2), device=dev)
            fn = lambda: trx.convolve(rx, 20, 600, hc, 0, 600, True)  # noqa: E731
    elif args.case == "resamp":
        rs = Resampler(trx, 65, 48)
        x = torch.randn((4096, 16 + 192 * 8, 2), device=dev)
        fn = lambda: rs.rotate(x, 260 * 8)  # noqa: E731
    elif args.case in ("chan4", "chan64"):
        m = 4 if args.case == "chan4" else 64
        nbl = 8192 // m * 4
        ch = Channelizer(trx, m, 192)
        sy = Synthesis(trx, m, 192)
        xw = torch.randn((nbl * 192 * m, 2), device=dev)
        xs = torch.randn((m, nbl * 192, 2), device=dev)
        fn = lambda: (ch.rotate(xw), sy.rotate(xs))  # noqa: E731
    else:
        raise SystemExit("unknown case " + args.case)
    for _ in range(args.reps):
        fn()
    torch.cuda.synchronize()
    if args.prof:
        trx.profile_begin()
        fn()
        print({k: (round(v[0], 4), v[1]) for k, v in trx.profile_end().items()})
    if args.time:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        print(f"{args.case} n={n}: {ms:.4f} ms per call, {n / ms * 1e3:.4g} units/s")


if __name__ == "__main__":
    main()
