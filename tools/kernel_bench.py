#!/usr/bin/env python
"""Per-kernel roofline table for every row of SURVEY.md 8(a): each batched entry point of the C ABI is timed alone
with CUDA events (inputs resident in HBM and larger than L2, warm-up first) and its ALGORITHMIC bytes per unit
 are divided by the time and by the measured HBM peak.

    python tools/kernel_bench.py [--json out.json] [--n 262144]

This is a measurement tool, not the bench contract (bench.py is); its table goes to profiles/."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1 << 18)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    import osmo_trx_b200
    import bench
    from osmo_trx_b200 import Trx, Resampler, Channelizer, Synthesis
    peak, peak_src = bench.peaks()
    trx = Trx(0)
    dev = trx.device
    n = args.n
    rows = []

    # FP32 pipe: 128 lanes per SM per clock; a multiply, an add or an FMA each occupy one lane slot ("lane-op")
    props = torch.cuda.get_device_properties(dev)
    fp32_peak = props.multi_processor_count * 128 * 1.965e9

    def add(name, unit, units, bytes_per_unit, ms, note="", ops=None):
        """ops: FP32 lane-ops per unit in the reference's operation order (exact chains cannot fuse mul+add), for the
        kernels whose arithmetic rather than their bytes bounds them"""
        gbs = units * bytes_per_unit / (ms * 1e-3) / 1e9
        row = dict(kernel=name, unit=unit, units_per_launch=units, alg_bytes_per_unit=bytes_per_unit, ms=ms,
                   units_per_s=units / (ms * 1e-3), achieved_gbs=gbs, frac_of_hbm_peak=gbs / peak, note=note)
        extra = ""
        if ops:
            f32 = units * ops / (ms * 1e-3)
            row.update(fp32_lane_ops_per_unit=ops, achieved_fp32_lane_ops=f32, frac_of_fp32_peak=f32 / fp32_peak,
                       bound="fp32" if f32 / fp32_peak > gbs / peak else "hbm")
            extra = f"  {100 * f32 / fp32_peak:5.1f} % of FP32 pipe"
        rows.append(row)
        print(f"{name:34s} {units / (ms * 1e-3):12.4g} {unit}/s  {gbs:8.1f} GB/s  {100 * gbs / peak:5.1f} % of HBM peak{extra}  {note}")

    # ---- modulators (a37-a40) ----
    bits = torch.randint(0, 2, (n, 148), dtype=torch.uint8, device=dev)
    out = torch.empty((n, 625, 2), dtype=torch.float32, device=dev)
    add("modulate_gmsk_kernel", "burst", n, 148 + 5000, timed(lambda: trx.modulate_gmsk(bits, out=out)))
    ebits = torch.randint(0, 2, (n, 444), dtype=torch.uint8, device=dev)
    add("modulate_edge_kernel", "burst", n, 444 + 5000, timed(lambda: trx.modulate_edge(ebits, out=out)))
    del bits, ebits, out

    # ---- detect / demod per burst type (a15-a36) ----
    for kind, soft, det_cfg in (("nb", 148, (16, 1)), ("rach", 148, (40, 1)), ("edge", 444, (16, 2))):
        rx, typ, tsc, mt, bound = bench.make_workload(trx, kind, n, seed=7, device=dev)
        trx.detect_config(*det_cfg)
        res = trx.alloc_results(n, soft)
        hl = 40 if kind == "rach" else 16
        win = (4 * (hl + 16 + bound - 1) + 12) * 8
        t_det = timed(lambda: trx.detect(rx, typ, tsc, mt, bound, out=res))
        nd = hl + 16 + bound - 1
        det_ops = nd * 62 + (16 + bound) * hl * 8 + 9 * 2 * 21 * 4  # decimate + correlate + early/late interpolation
        # the clip scan (all 625 samples) only visits the bursts detection left at rc == 0
        undet = float((res["rc"] == 0).float().mean().item())
        add(f"detect[{kind}] (+clip scan)", "burst", n, win + 24 + undet * 5000, t_det,
            f"standalone detectAnyBurst: correlator window {win} B + results + the 625-sample clip scan of the {100 * undet:.0f} % of "
            "bursts left undetected", ops=det_ops)
        t_dem = timed(lambda: trx.demod(rx, res["rc"], res["amp"], res["toa"], res["ci"], soft=res["soft"], n_gmsk_soft=148))
        add(f"demod_kernel[{kind}]", "burst", n, 5000 + soft * 4 + 16, t_dem, ops=156 * 35 * 2 + 32 * 20 * 2 + (148 * 40 if kind == "edge" else 0))
        t_dd = timed(lambda: trx.detect_demod(rx, typ, tsc, mt, bound, n_gmsk_soft=148, out=res))
        trx.profile_begin()
        trx.detect_demod(rx, typ, tsc, mt, bound, n_gmsk_soft=148, out=res)
        pr = trx.profile_end()
        split = ", ".join(f"{k} {v[0]:.3f} ms/{v[1]}" for k, v in pr.items())
        add(f"detect_demod[{kind}] fused step", "burst", n, 5000 + soft * 4 + 24, t_dd, "the bench.py step; " + split)
        if kind == "nb":
            # ---- pull path on device (8(f) rows 1-3): int16 in, TRXD out ----
            iq = (rx * bench.IQ_SCALE).round().clamp(-32768, 32767).to(torch.int16)
            fn = (torch.arange(n, device=dev, dtype=torch.int32) // 8)
            tn = (torch.arange(n, device=dev) % 8).to(torch.uint8)
            po = trx.alloc_pull_results(n, 160)
            t_pull = timed(lambda: trx.pull(iq, typ, tsc, mt, fn, tn, bound, out=po))
            add("pull[nb] int16 -> TRXD v1", "burst", n, 2500 + 159 + 10, t_pull,
                "gate + detect and demod on the int16 slot (soft bits written as datagram bytes) + header; issue bound in demod_kernel<true>, not HBM bound")
            # ---- helpers ----
            e_t = timed(lambda: trx.energy_detect(rx, 80))
            add("energy_detect_kernel", "burst", n, 80 * 32 + 4, e_t,
                "80 samples at stride 4 = one 8-byte sample per 32-byte DRAM sector: bytes counted per sector (2,564 B/burst; 644 B of samples)")
            flat = res["soft"].reshape(-1)
            add("vector_slicer_kernel", "value", flat.numel(), 8, timed(lambda: trx.vector_slicer(flat)))
            xi = iq.reshape(-1)
            add("convert_short_float_kernel", "value", xi.numel(), 6, timed(lambda: trx.convert_short_float(xi)))
            xf = rx.reshape(-1)
            add("convert_float_short_kernel", "value", xf.numel(), 6, timed(lambda: trx.convert_float_short(xf, 0.5)))
            dl = torch.rand(n, device=dev) * 8 - 4
            add("delay_vector_kernel", "burst", n, 10000, timed(lambda: trx.delay_vector(rx, dl), reps=5), ops=625 * 78)
            h = torch.randn((16, 2), device=dev)
            h[:, 1] = 0
            add("convolve_kernel real16", "output", n * 600, 16,
                timed(lambda: trx.convolve(rx, 20, 600, h, 0, 600, False), reps=5), "x read once + y written", ops=62)
            hc = torch.randn((16, 2), device=dev)
            add("convolve_kernel complex16", "output", n * 600, 16,
                timed(lambda: trx.convolve(rx, 20, 600, hc, 0, 600, True), reps=5), ops=126)
            del iq, po, dl
        del rx, res
    trx.detect_config(40, 3)

    # ---- SCH search (8(f) row 4): 64-symbol sequence, 156 correlation outputs over the whole decimated burst ----
    rx, typ, tsc, mt, bound = bench.make_workload(trx, "nb", n, seed=8, device=dev)
    sch_ops = 219 * 62 + 156 * 64 * 8 + 9 * 2 * 21 * 4
    add("detect_sch[full] corr+peak", "burst", n, 5000 + 24, timed(lambda: trx.detect_sch(rx), reps=5),
        "detectSCHBurst(SCH_DETECT_FULL): 80 k complex taps per burst", ops=sch_ops)
    del rx

    # ---- vitac (a42-a47) ----
    nv = n
    rx, typ, tsc, mt, bound = bench.make_workload(trx, "nb", nv, seed=9, device=dev)
    buf = torch.zeros((nv, 40 + 625 + 63, 2), dtype=torch.float32, device=dev)
    buf[:, 40:665] = rx
    del rx
    add("vitac_kernel[nb]", "burst", nv, 5000 + 148 + 8, timed(lambda: trx.vitac(buf, 40, tsc), reps=5))
    del buf

    # ---- resampler / filterbanks (a48-a51) ----
    ns = 4096
    rs = Resampler(trx, 65, 48)
    x = torch.randn((ns, 16 + 192 * 8, 2), device=dev)
    add("resampler_kernel 65/48", "sample-out", ns * 260 * 8, 8 + 8 * 48 / 65, timed(lambda: rs.rotate(x, 260 * 8)))
    rs14 = Resampler(trx, 1, 4)
    x4 = torch.randn((ns * 4, 16 + 624, 2), device=dev)
    add("resampler_kernel 1/4", "sample-out", ns * 4 * 156, 8 + 32, timed(lambda: rs14.rotate(x4, 156)))
    del x, x4
    for m in (4, 64):
        nb_blocks = 8192 // m * 4
        ch = Channelizer(trx, m, 192)
        xw = torch.randn((nb_blocks * 192 * m, 2), device=dev)
        add(f"channelizer_kernel M={m}", "block", nb_blocks, 2 * 192 * m * 8, timed(lambda: ch.rotate(xw), reps=10),
            f"{nb_blocks} blocks of 192x{m} samples per call")
        sy = Synthesis(trx, m, 192)
        xs = torch.randn((m, nb_blocks * 192, 2), device=dev)
        add(f"synthesis_kernel M={m}", "block", nb_blocks, 2 * 192 * m * 8, timed(lambda: sy.rotate(xs), reps=10))
        del xw, xs, ch, sy

    if args.json:
        json.dump(dict(hbm_peak_gbs=peak, peak_source=peak_src, n=n, rows=rows), open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
