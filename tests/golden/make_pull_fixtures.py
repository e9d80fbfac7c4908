"""Golden vectors for the receive chain around the hot path (int16 slot -> TRXD uplink datagram), produced by the
UNMODIFIED reference functions in oracle/_ref (convert_short_float, energyDetect, detectAnyBurst, demodAnyBurst,
vectorSlicer, trxd_send_burst_ind_v0/_v1 writing into a pipe; driver ref_pull_batch in oracle/ref_capi.cpp).
Run in the dev container (needs /root/reference):   python tests/golden/make_pull_fixtures.py
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
    rng = np.random.default_rng(20261018)
    d = {}
    n = 64
    tsc = (np.arange(n) % 8).astype(np.uint8)
    nb, _ = synth.impair(r.modulate_gmsk_batch(synth.nb_bits(n, tsc, rng)), rng,
                         snr_db=np.choose(np.arange(n) % 3, [30.0, 10.0, 6.0]), noise_only_frac=0.1)
    ed, _ = synth.impair(r.modulate_edge_batch(synth.edge_bits(n, tsc, rng)), rng, snr_db=28.0)
    ab, _ = synth.impair(r.modulate_gmsk_batch(synth.ab_bits(n, 20, rng, 0)), rng, snr_db=15.0)
    rx = np.concatenate([nb, ed, ab])
    iq = np.clip(np.rint(rx * 8000.0), -32768, 32767).astype(np.int16)
    iq[3] = 0
    iq[4] = rng.integers(-32768, 32767, (625, 2))
    typ = np.concatenate([np.full(n, TSC), np.full(n, EDGE), np.full(n, RACH)]).astype(np.uint8)
    typ[7] = IDLE
    typ[9] = 0
    typ[n + 5] = IDLE
    typ[2 * n + 1::7] = EXT_RACH
    tscs = np.concatenate([tsc, tsc, np.zeros(n, np.uint8)])
    mt = np.concatenate([np.full(n, 4), np.full(n, 4), np.full(n, 63)]).astype(np.uint16)
    fn = rng.integers(0, 2715648, 3 * n).astype(np.uint32)
    tn = rng.integers(0, 8, 3 * n).astype(np.uint8)
    d["iq"], d["type"], d["tsc"], d["max_toa"], d["fn"], d["tn"] = iq, typ, tscs, mt, fn, tn
    d["rssi_offset"] = np.float64(-3.5)
    for v in (0, 1):
        o = r.pull(iq, typ, tscs, mt, fn, tn, version=v, rssi_offset=-3.5)
        for k in ("rc", "energy", "pkt", "pkt_len", "amp", "toa", "ci", "tsc"):
            d[f"v{v}/{k}"] = o[k]
    out = os.path.join(HERE, "pull_fixtures.npz")
    np.savez_compressed(out, **d)
    print("wrote", out, os.path.getsize(out) // 1024, "KiB")


if __name__ == "__main__":
    main()
