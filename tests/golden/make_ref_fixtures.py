"""Generate golden input/output fixtures by running the UNMODIFIED reference (oracle/_ref) here.

The reference cannot travel to the GPU box as source, and its own tests pin only convolve_* (SURVEY §4), so
everything else is pinned by these vectors: seeded inputs + the reference's outputs.  Run in the dev container
(needs oracle/_ref built from /root/reference):   python tests/golden/make_ref_fixtures.py
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import cpulibs  # noqa: E402
import synth  # noqa: E402
from cpulibs import TSC, EXT_RACH, RACH, EDGE, IDLE  # noqa: E402


def main():
    r = cpulibs.Ref()
    rng = np.random.default_rng(20261017)
    d = {}
    # tables
    names = [("sinc", 0), ("rot4", 0), ("rrot1", 0), ("pulse1_c0", 0), ("sch", 0), ("sch_meta", 0), ("dummy", 0), ("dummy_meta", 0)]
    names += [("delay", i) for i in (0, 1, 31, 63)] + [("midamble", i) for i in range(8)] + [("midamble_meta", i) for i in range(8)]
    names += [("edge_midamble", i) for i in range(8)] + [("rach", i) for i in range(3)] + [("rach_meta", i) for i in range(3)]
    for nm, i in names:
        d[f"table/{nm}/{i}"] = r.get_table(nm, i)
    # NB / RACH / EDGE detect+demod
    n = 96
    tsc = (np.arange(n) % 8).astype(np.uint8)
    bits = synth.nb_bits(n, tsc, rng)
    tx = r.modulate_gmsk_batch(bits)
    rx, _ = synth.impair(tx, rng, snr_db=np.choose(np.arange(n) % 3, [30.0, 10.0, 6.0]), noise_only_frac=0.08)
    rx[5] *= 60000.0
    d["nb/bits"], d["nb/tx"], d["nb/rx"], d["nb/tsc"] = bits, tx, rx, tsc
    for k, v in r.detect_demod(rx, TSC, tsc, 4).items():
        if k != "flags":
            d[f"nb/out/{k}"] = v[:, :156] if k == "soft" else v
    outs = []
    for seq in range(3):
        for delay in (0, 33):
            b = synth.ab_bits(8, delay, rng, seq)
            outs.append(np.stack([r.modulate_burst(b[i], 68 - delay, 4) for i in range(8)]))
    rxr, _ = synth.impair(np.concatenate(outs), rng, snr_db=12.0, shift_lo=-22, shift_hi=-14)
    d["rach/rx"] = rxr
    for typ in (RACH, EXT_RACH):
        for k, v in r.detect_demod(rxr, typ, 0, 63).items():
            if k != "flags":
                d[f"rach{typ}/out/{k}"] = v[:, :156] if k == "soft" else v
    ne = 48
    tsce = (np.arange(ne) % 8).astype(np.uint8)
    eb = synth.edge_bits(ne, tsce, rng)
    etx = r.modulate_edge_batch(eb)
    erx, _ = synth.impair(etx, rng, snr_db=25.0)
    erx[:8] = rx[:8]  # GMSK bursts -> TSC fallback
    d["edge/bits"], d["edge/tx"], d["edge/rx"], d["edge/tsc"] = eb, etx, erx, tsce
    for k, v in r.detect_demod(erx, EDGE, tsce, 4).items():
        if k != "flags":
            d[f"edge/out/{k}"] = v
    # vitac
    nv = 48
    tscv = (np.arange(nv) % 8).astype(np.uint8)
    vb = synth.nb_bits(nv, tscv, rng)
    vrx, _ = synth.impair(synth.multipath(r.modulate_gmsk_batch(vb), rng), rng, snr_db=34.0, amp_range=(0.5, 1.0), shift_lo=-4, shift_hi=4)
    buf = np.zeros((nv, 40 + 625 + 63, 2), np.float32)
    buf[:, 40:665] = vrx
    d["vitac/buf"], d["vitac/tsc"] = buf, tscv
    for k, v in r.vitac(buf, 40, tscv).items():
        d[f"vitac/out/{k}"] = v
    # resampler / channelizer / synthesis
    x = rng.standard_normal((16 + 48 * 2, 2)).astype(np.float32)
    d["rs/x"] = x
    d["rs/y"] = r.resampler_rotate(r.resampler(65, 48), x, 16, 130)[1]
    m, bl = 4, 192
    xc = rng.standard_normal((2, m * bl, 2)).astype(np.float32)
    hc, hs = r.channelizer(m, bl), r.synthesis(m, bl)
    d["chan/x"] = xc
    d["chan/y"] = np.stack([r.channelizer_rotate(hc, xc[k], m, bl)[1] for k in range(2)])
    xs = rng.standard_normal((2, m, bl, 2)).astype(np.float32)
    d["synth/x"] = xs
    d["synth/y"] = np.stack([r.synthesis_rotate(hs, xs[k], m, bl)[1] for k in range(2)])
    # helpers
    d["misc/x"] = rx[0]
    d["misc/energy80"] = np.float32(r.energy_detect(rx[0], 80))
    d["misc/delay_3.3"] = r.delay_vector(rx[0], 3.3)
    d["misc/delay_-7.71"] = r.delay_vector(rx[0], -7.71)
    out = os.path.join(HERE, "ref_fixtures.npz")
    np.savez_compressed(out, **d)
    print("wrote", out, os.path.getsize(out) // 1024, "KiB,", len(d), "arrays")


if __name__ == "__main__":
    main()
