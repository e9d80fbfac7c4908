"""Parity criteria of BASELINE.json's north_star, as assertions.

* rc (burst-type decision), TSC and hard bits: bit-exact, except bursts whose peak-to-average ratio lies
  within 1e-5 of the threshold (flag THRESH_EDGE) - counted and reported;
* TOA within 1e-3 symbol (bursts with a bisection near-tie, flag BISECT_TIE, are counted and reported: the
  reference's own SSE and scalar builds disagree by 1/256 symbol on those);
* soft bits and |amp| within 1e-4 relative.  "Relative" for soft bits is taken against the burst's soft-bit
  scale max(|ref|) (soft bits cross zero, a per-element ratio is meaningless there); hard bits may only
  differ where |ref soft| < 1e-5 * scale (counted and reported).
"""
import numpy as np

TOA_TOL = 1e-3
REL_TOL = 1e-4


def compare_detect(gpu, ref, flags=None, what=""):
    """gpu/ref: dicts of numpy arrays rc, amp[n,2], toa, tsc, ci. Returns a report dict."""
    n = len(ref["rc"])
    flags = np.zeros(n, np.uint8) if flags is None else flags
    edge = (flags & 1) != 0
    tie = (flags & 2) != 0
    rc_bad = (gpu["rc"] != ref["rc"]) & ~edge
    assert not rc_bad.any(), f"{what}: rc mismatch on {rc_bad.sum()} bursts, first {np.nonzero(rc_bad)[0][:5]}: " \
        f"gpu {gpu['rc'][rc_bad][:5]} ref {ref['rc'][rc_bad][:5]}"
    same = gpu["rc"] == ref["rc"]
    det = same & (ref["rc"] > 0)
    tsc_bad = same & (gpu["tsc"] != ref["tsc"])
    assert not tsc_bad.any(), f"{what}: tsc mismatch on {tsc_bad.sum()} bursts"
    dtoa = np.abs(gpu["toa"].astype(np.float64) - ref["toa"])
    toa_bad = same & (dtoa > TOA_TOL) & ~tie
    assert not toa_bad.any(), f"{what}: TOA off by up to {dtoa[toa_bad].max()} on {toa_bad.sum()} bursts"
    ok = det & (dtoa <= TOA_TOL)
    ga = np.hypot(gpu["amp"][:, 0].astype(np.float64), gpu["amp"][:, 1])
    ra = np.hypot(ref["amp"][:, 0].astype(np.float64), ref["amp"][:, 1])
    damp = np.abs(gpu["amp"].astype(np.float64) - ref["amp"]).max(axis=1)
    amp_rel = np.where(ra > 0, damp / np.maximum(ra, 1e-30), damp)
    assert (amp_rel[ok] <= REL_TOL).all(), f"{what}: amp rel err {amp_rel[ok].max()}"
    # undetected bursts report zeros
    und = same & (ref["rc"] <= 0)
    assert (gpu["toa"][und] == 0).all() and (gpu["amp"][und] == 0).all()
    return dict(n=n, detected=int(det.sum()), thresh_edge=int(edge.sum()), bisect_tie=int(tie.sum()),
                rc_diff_on_edge=int(((gpu["rc"] != ref["rc"]) & edge).sum()),
                toa_exact=int((same & (dtoa == 0)).sum()), toa_max=float(dtoa[same].max() if same.any() else 0),
                toa_quantum_flips=int((same & (dtoa > TOA_TOL)).sum()),
                amp_rel_max=float(amp_rel[ok].max() if ok.any() else 0), ok_mask=ok, _ga=ga)


def compare_ci(gpu_ci, ref_ci, mask, what="", tol=1e-3):
    """C/I in dB: not part of the north_star list; checked to 1e-3 dB absolute or 1e-4 relative."""
    g, r = gpu_ci[mask].astype(np.float64), ref_ci[mask].astype(np.float64)
    fin = np.isfinite(r)
    assert (np.isfinite(g) == fin).all(), f"{what}: ci finiteness differs"
    d = np.abs(g[fin] - r[fin])
    bad = d > np.maximum(tol, REL_TOL * np.abs(r[fin]))
    assert not bad.any(), f"{what}: ci differs by up to {d.max()} dB"
    return float(d.max() if d.size else 0.0)


def compare_soft(gpu_soft, ref_soft, mask, nsoft, what=""):
    """Soft/hard bit parity on rows in `mask`, first nsoft columns."""
    g = gpu_soft[mask][:, :nsoft].astype(np.float64)
    r = ref_soft[mask][:, :nsoft].astype(np.float64)
    if g.size == 0:
        return dict(soft_rel_max=0.0, hard_flips=0, hard_flips_near_zero=0)
    scale = np.abs(r).max(axis=1, keepdims=True)
    scale = np.maximum(scale, 1e-30)
    rel = np.abs(g - r) / scale
    assert rel.max() <= REL_TOL, f"{what}: soft-bit rel err {rel.max()}"
    flips = (g > 0) != (r > 0)
    near0 = np.abs(r) < 1e-5 * scale
    assert not (flips & ~near0).any(), f"{what}: {int((flips & ~near0).sum())} hard-bit flips away from zero"
    return dict(soft_rel_max=float(rel.max()), hard_flips=int(flips.sum()), hard_flips_near_zero=int((flips & near0).sum()))


def compare_pkts(gpu, ref, ref_soft01=None, what="", version=1):
    """TRXD uplink datagrams of the pull path (proto_trxd.c:27-117): gpu/ref dicts with rc, energy, pkt, pkt_len, flags.

    Exact: pkt_len, tn/fn/rssi bytes, the v1 idle/modulation/tsc byte, TOA (1/256 symbol units) unless the burst
    carries a bisection near-tie flag.  Soft bits are floats (1e-4 tolerance) quantised to 0..255: a byte may
    differ by one LSB only where the reference's value sits within 1e-4 * 255 of a rounding boundary (checked when
    ref_soft01, the reference's 0..1 soft values, is given); C/I in cB may differ by one unit for 8-PSK bursts
    (computeEdgeCI is part of the demodulator, 1e-4 tolerance).  Bursts flagged THRESH_EDGE are excluded and counted.
    """
    n = len(ref["rc"])
    flags = gpu.get("flags")
    flags = np.zeros(n, np.uint8) if flags is None else flags
    edge = (flags & 1) != 0
    tie = (flags & 2) != 0
    live = ~edge
    assert np.array_equal(gpu["rc"][live], ref["rc"][live]), f"{what}: rc mismatch"
    assert np.array_equal(gpu["energy"].view(np.uint32), ref["energy"].view(np.uint32)), f"{what}: energyDetect differs"
    assert np.array_equal(gpu["pkt_len"][live].astype(np.int64), ref["pkt_len"][live].astype(np.int64)), f"{what}: pkt_len mismatch"
    hdr = 11 if version == 1 else 8
    soft_diff = 0
    soft_total = 0
    ci_off = 0
    toa_off = 0
    for b in np.nonzero(live & (ref["pkt_len"] > 0))[0]:
        L = int(ref["pkt_len"][b])
        g, r = gpu["pkt"][b, :L].astype(np.int64), ref["pkt"][b, :L].astype(np.int64)
        assert np.array_equal(g[:6], r[:6]), f"{what}: burst {b} common/rssi header {g[:6]} vs {r[:6]}"
        if not np.array_equal(g[6:8], r[6:8]):
            assert tie[b], f"{what}: burst {b} TOA bytes differ without a near-tie flag"
            toa_off += 1
            continue  # a different TOA quantum moves every soft bit: counted, not compared
        if version == 1:
            assert g[8] == r[8], f"{what}: burst {b} v1 flags byte"
            gc = int(np.int16((g[9] << 8) | g[10]))
            rcb = int(np.int16((r[9] << 8) | r[10]))
            if gc != rcb:
                assert abs(gc - rcb) <= 1 and ref["rc"][b] == 5, f"{what}: burst {b} ci {gc} vs {rcb} cB"
                ci_off += 1
        nb = L - hdr - (2 if version == 0 else 0)
        if nb > 0:
            d = g[hdr:hdr + nb] - r[hdr:hdr + nb]
            soft_total += nb
            if d.any():
                assert np.abs(d).max() <= 1, f"{what}: burst {b} soft byte off by {np.abs(d).max()}"
                soft_diff += int((d != 0).sum())
                if ref_soft01 is not None and not tie[b]:
                    x = ref_soft01[b, :nb].astype(np.float64) * 255.0
                    dist = np.abs(x - np.floor(x) - 0.5)
                    assert (dist[d != 0] <= REL_TOL * 255.0).all(), \
                        f"{what}: burst {b} soft byte differs away from a rounding boundary ({dist[d != 0].max()})"
        if version == 0:
            assert g[L - 1] == 0  # trailing NUL (proto_trxd.c:86); the byte before it is uninitialised in the reference
    return dict(n=n, sent=int((ref["pkt_len"] > 0).sum()), detected=int((ref["rc"] > 0).sum()), thresh_edge=int(edge.sum()),
                bisect_tie=int(tie.sum()), toa_bytes_off=toa_off, ci_cb_off_by_one=ci_off, soft_bytes=soft_total,
                soft_bytes_off_by_one=soft_diff)
