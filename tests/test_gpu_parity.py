"""GPU parity tests proper: CUDA path (through the C ABI) vs the CPU checker on identical seeded inputs."""
import contextlib
import sys

import numpy as np
import pytest
import torch

import cpulibs
import parity
import synth
from cpulibs import TSC, EXT_RACH, RACH, EDGE, IDLE

pytestmark = pytest.mark.gpu


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def run_gpu_dd(trx, rx, typ, tsc, max_toa, bound, n_soft=156, soft_stride=444, thresh=4.0):
    n = rx.shape[0]
    t_type = dev(np.broadcast_to(np.asarray(typ, np.uint8), (n,)).copy())
    t_tsc = dev(np.broadcast_to(np.asarray(tsc, np.uint8), (n,)).copy())
    t_toa = dev(np.broadcast_to(np.asarray(max_toa, np.int16), (n,)).copy())
    r = trx.detect_demod(dev(rx), t_type, t_tsc, t_toa, bound, thresh=thresh, n_gmsk_soft=n_soft, soft_stride=soft_stride)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in r.items()}


# detect_config settings the detection tests are run under: the default (long-window correlator, three attempts),
# and the two 16-symbol hints that select the register-blocked / fused normal-burst kernels (the benchmarked path)
DETECT_CFGS = [(40, 3), (16, 1), (16, 2)]


@contextlib.contextmanager
def detect_cfg(trx, cfg):
    try:
        trx.detect_config(*cfg)
        yield
    finally:
        trx.detect_config(40, 3)


def unsupported_under(cfg, typ, ref_rc, n):
    """Bursts the configuration hint cannot serve and that must therefore come back as -SIGERR_BOUNDS (never a silent
    wrong answer): access bursts under the 16-symbol hint; 8-PSK bursts that the reference only finds in a later
    attempt (EDGE -> TSC fall-through, sigProcLib.cpp:1933-1941) when fewer attempts are scheduled."""
    typ = np.broadcast_to(np.asarray(typ, np.uint8), (n,))
    bad = np.zeros(n, bool)
    if cfg[0] < 40:
        bad |= (typ == RACH) | (typ == EXT_RACH)
    if cfg[1] < 2:
        bad |= (typ == EDGE) & (ref_rc != EDGE) & (ref_rc != -3)
    if cfg[1] < 3:
        bad |= (typ == EXT_RACH)  # sequences 1, 2 are attempts 2, 3; a hit on sequence 0 is still reported (checked below)
    return bad


def check_dd(trx, checker, rx, typ, tsc, max_toa, what, n_soft=156, cfg=(40, 3), thresh=4.0):
    bound = int(np.max(max_toa))
    n = rx.shape[0]
    with detect_cfg(trx, cfg):
        g = run_gpu_dd(trx, rx, typ, tsc, max_toa, bound, n_soft=n_soft, thresh=thresh)
    c = checker.detect_demod(rx, typ, tsc, max_toa, thresh=thresh)
    bad = unsupported_under(cfg, typ, c["rc"], n)
    if bad.any():
        # loud failure on what the hint excludes (an EXT_RACH hit on the first sequence is a legitimate answer)
        gb = g["rc"][bad]
        assert ((gb == -1) | (gb == c["rc"][bad])).all(), f"{what}: unsupported bursts not reported as -SIGERR_BOUNDS"
        keep = ~bad
        g = {k: v[keep] for k, v in g.items()}
        c = {k: v[keep] for k, v in c.items()}
    rep = parity.compare_detect(g, c, g["flags"], what)
    ok = rep["ok_mask"]
    gmsk = ok & (c["rc"] != EDGE)
    edge = ok & (c["rc"] == EDGE)
    rep["ci_max_dB"] = parity.compare_ci(g["ci"], c["ci"], ok, what)
    rep["gmsk"] = parity.compare_soft(g["soft"], c["soft"], gmsk, n_soft, what + " gmsk")
    rep["edge"] = parity.compare_soft(g["soft"], c["soft"], edge, 444, what + " edge")
    rep.pop("ok_mask"); rep.pop("_ga")
    print(what, rep)
    return g, c, rep


def test_tables_bit_exact(trx, checker):
    names = [("sinc", 0), ("rot4", 0), ("rrot4", 0), ("rot1", 0), ("rrot1", 0), ("pulse4_c0", 0), ("pulse4_c1", 0),
             ("pulse4_c0inv", 0), ("pulse1_c0", 0), ("sch", 0), ("sch_meta", 0), ("dummy", 0), ("dummy_meta", 0), ("psk8", 0)]
    names += [("delay", i) for i in range(64)]
    for nm in ("midamble", "edge_midamble"):
        names += [(nm, i) for i in range(8)] + [(nm + "_meta", i) for i in range(8)]
    names += [("rach", i) for i in range(3)] + [("rach_meta", i) for i in range(3)]
    for nm, i in names:
        a, b = trx.get_table(nm, i), checker.get_table(nm, i)
        assert a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32)), (nm, i)
    for t in range(9):
        assert np.array_equal(trx.get_table("vitac_norm", t).reshape(-1, 2), checker.vitac_table(0, t))
    assert np.array_equal(trx.get_table("vitac_access").reshape(-1, 2), checker.vitac_table(1))
    assert np.array_equal(trx.get_table("vitac_sch").reshape(-1, 2), checker.vitac_table(2))


def test_modulate_gmsk_exact(trx, checker):
    rng = np.random.default_rng(11)
    n = 512
    bits = synth.nb_bits(n, np.arange(n) % 8, rng)
    bits[::5] |= 0xA2  # BitVector elements carry garbage above bit 0 (BitVector.cpp:59-68): only &1 counts
    ref = checker.modulate_gmsk_batch(bits & 1)
    out = trx.modulate_gmsk(dev(bits)).cpu().numpy()
    assert np.array_equal(out, ref)  # value-exact (signed zeros compare equal)
    # ragged: access-burst lengths
    for nb in (88, 100, 147, 155, 3):  # 156 overruns the reference's own buffer (sigProcLib.cpp:628)
        b = rng.integers(0, 2, (7, nb)).astype(np.uint8)
        ref = np.stack([checker.modulate_burst(b[i], 0, 4) for i in range(7)])
        out = trx.modulate_gmsk(dev(b)).cpu().numpy()
        assert np.array_equal(out, ref), nb


def test_modulate_edge_exact(trx, checker):
    rng = np.random.default_rng(12)
    n = 256
    bits = synth.edge_bits(n, np.arange(n) % 8, rng)
    ref = checker.modulate_edge_batch(bits)
    out = trx.modulate_edge(dev(bits)).cpu().numpy()
    assert np.array_equal(out, ref)


def test_modulate_basic_forms(trx, checker):
    """modulateBurst outside the 4-sps Laurent case: modulateBurstBasic (sps 1), rotateBurst (emptyPulse, sps 1 and 4),
    rotateEdgeBurst (modulateEdgeBurst with emptyPulse) - the forms sigProcLibSetup builds its correlation references
    with (sigProcLib.cpp:558-580,672-689,938-979).  Value-exact."""
    rng = np.random.default_rng(13)
    for nb, guard in ((148, 8), (148, 9), (88, 68), (26, 0), (41, 3), (64, 0), (1, 0)):
        b = rng.integers(0, 2, (6, nb)).astype(np.uint8)
        b[0] |= 0xF0  # only bit 0 counts
        for sps, mode, empty in ((1, 0, False), (1, 1, True), (4, 1, True)):
            if sps * (nb + guard) > (157 if sps == 1 else 625):
                continue
            out = trx.modulate_basic(dev(b), guard, sps, mode).cpu().numpy()
            for i in range(len(b)):
                ref = checker.modulate_burst(b[i] & 1, guard, sps, empty)
                assert out[i].shape == ref.shape and np.array_equal(out[i], ref), (nb, guard, sps, mode, i)
    eb = synth.edge_bits(5, np.arange(5), rng)
    for sps in (1, 4):
        out = trx.modulate_basic(dev(eb), 0, sps, 2).cpu().numpy()
        for i in range(5):
            ref = checker.modulate_edge(eb[i], sps, True)
            assert out[i].shape == ref.shape and np.array_equal(out[i], ref), (sps, i)
    with pytest.raises(Exception):
        trx.modulate_basic(dev(rng.integers(0, 2, (2, 148)).astype(np.uint8)), 8, 4, 0)  # no single-pulse shaper at 4 sps


@pytest.mark.parametrize("cfg", DETECT_CFGS)
def test_nb_detect_demod_cfg1(trx, checker, cfg):
    """BASELINE configs[0]: 10k GMSK normal bursts, TSC 0-7, sps=4, AWGN + random TOA."""
    rng = np.random.default_rng(1)
    n = 10000
    tsc = np.arange(n) % 8
    bits = synth.nb_bits(n, tsc, rng)
    w = checker.modulate_gmsk_batch(bits, nthreads=8)
    snr = np.choose(np.arange(n) % 3, [30.0, 10.0, 6.0])
    rx, _ = synth.impair(w, rng, snr_db=snr, noise_only_frac=0.05)
    g, c, rep = check_dd(trx, checker, rx, TSC, tsc, 4, f"cfg1-NB{cfg}", cfg=cfg)
    assert rep["detected"] > 0.9 * 0.95 * n
    det = c["rc"] > 0
    hard = (g["soft"][det][:, :148] > 0).astype(np.uint8)
    ber = (hard != bits[det]).mean()
    print("cfg1 BER", ber)
    assert ber < 0.02
    # also with the 148-wide output the transceiver consumes
    with detect_cfg(trx, cfg):
        g2 = run_gpu_dd(trx, rx[:512], TSC, tsc[:512], 4, 4, n_soft=148, soft_stride=148)
    assert np.array_equal(g2["soft"], g["soft"][:512, :148])


def test_detect_demod_pipeline_same_bits(trx, checker, monkeypatch):
    """The optional two-stream detect -> demod pipeline (TRXB200_OVERLAP=1, off by default: measured slower on B200)
    must give the serial path's results bit for bit, chunk boundaries and undetected-burst clip reports included."""
    import osmo_trx_b200
    rng = np.random.default_rng(77)
    n = 9000
    tsc = np.arange(n) % 8
    w = checker.modulate_gmsk_batch(synth.nb_bits(n, tsc, rng), nthreads=8)
    rx, _ = synth.impair(w, rng, snr_db=12.0, full_scale=20000.0, noise_only_frac=0.2)
    rx[::9] = (rng.standard_normal((len(rx[::9]), 625, 2)) * 15000.0).astype(np.float32)  # loud noise: clip, no burst
    ref = run_gpu_dd(trx, rx, TSC, tsc, 4, 4, n_soft=148, soft_stride=148)
    monkeypatch.setenv("TRXB200_OVERLAP", "1")
    monkeypatch.setenv("TRXB200_CHUNK", "4096")
    t2 = osmo_trx_b200.Trx(0)
    try:
        got = run_gpu_dd(t2, rx, TSC, tsc, 4, 4, n_soft=148, soft_stride=148)
    finally:
        t2.close()
    det = ref["rc"] > 0
    assert det.sum() > 0.7 * n and (ref["rc"] < 0).any()
    for k in ("rc", "flags"):
        assert np.array_equal(got[k], ref[k]), k
    for k in ("toa", "amp", "ci", "tsc", "soft"):
        assert np.array_equal(got[k][det], ref[k][det]), k


@pytest.mark.parametrize("cfg", DETECT_CFGS)
def test_nb_full_scale_and_clipping(trx, checker, cfg):
    rng = np.random.default_rng(5)
    n = 1024
    tsc = np.arange(n) % 8
    w = checker.modulate_gmsk_batch(synth.nb_bits(n, tsc, rng))
    rx, _ = synth.impair(w, rng, snr_db=20.0, full_scale=20000.0, noise_only_frac=0.2)
    rx[::7] *= 3.0  # drive some bursts beyond CLIP_THRESH (still detected -> no clip report)
    rx[::5] = (rng.standard_normal((len(rx[::5]), 625, 2)) * 15000.0).astype(np.float32)  # loud noise: clip, no burst
    g, c, rep = check_dd(trx, checker, rx, TSC, tsc, 4, f"clip{cfg}", cfg=cfg)
    assert (c["rc"] == -2).any() and (c["rc"] == 1).any()
    # standalone detect reports clipping too - but only for the types detectAnyBurst handles: OFF / SCH / unknown
    # types return 0 without a clipping check (sigProcLib.cpp:1949-1956)
    n2 = 256
    typ2 = np.full(n2, TSC, np.uint8)
    typ2[::10] = 0     # OFF on loud-noise bursts (every fifth burst is loud noise)
    typ2[5::20] = 4    # SCH is not a detectAnyBurst type (loud noise too); the loud bursts at 15 mod 20 stay TSC
    typ2[3::20] = 9    # unknown type on ordinary bursts
    with detect_cfg(trx, cfg):
        r = trx.detect(dev(rx[:n2]), dev(typ2), dev(tsc[:n2].astype(np.uint8)), dev(np.full(n2, 4, np.int16)), 4)
    c2 = checker.detect_demod(rx[:n2], typ2, tsc[:n2], 4)
    assert np.array_equal(r["rc"].cpu().numpy(), c2["rc"])
    assert (c2["rc"][typ2 != TSC] == 0).all() and (c2["rc"] == -2).any()


def test_rach(trx, checker):
    rng = np.random.default_rng(2)
    outs = []
    for seq in range(3):
        for delay in (0, 7, 33, 59):
            b = synth.ab_bits(40, delay, rng, seq)
            outs.append(np.stack([checker.modulate_burst(b[i], 68 - delay, 4) for i in range(40)]))
    w = np.concatenate(outs)
    rx, _ = synth.impair(w, rng, snr_db=12.0, shift_lo=-22, shift_hi=-14, noise_only_frac=0.05)
    for typ in (RACH, EXT_RACH):
        g, c, rep = check_dd(trx, checker, rx, typ, 0, 63, f"rach-{typ}")
        assert rep["detected"] > 100


@pytest.mark.parametrize("cfg", DETECT_CFGS)
def test_edge_and_fallback(trx, checker, cfg):
    rng = np.random.default_rng(3)
    n = 2000
    tsc = np.arange(n) % 8
    w = checker.modulate_edge_batch(synth.edge_bits(n, tsc, rng), nthreads=8)
    rx, _ = synth.impair(w, rng, snr_db=25.0, noise_only_frac=0.05)
    wn = checker.modulate_gmsk_batch(synth.nb_bits(300, tsc[:300], rng))
    rx[:300], _ = synth.impair(wn, rng, snr_db=20.0)
    g, c, rep = check_dd(trx, checker, rx, EDGE, tsc, 4, f"edge{cfg}", cfg=cfg)
    assert (c["rc"] == EDGE).sum() > 1000 and (cfg[1] < 2 or (c["rc"] == TSC).sum() > 200)


def test_idle_and_mixed_types(trx, checker):
    rng = np.random.default_rng(4)
    n = 1500
    tsc = np.arange(n) % 8
    w = checker.modulate_gmsk_batch(synth.nb_bits(n, tsc, rng))
    rx, _ = synth.impair(w, rng, snr_db=15.0)
    typ = np.choose(np.arange(n) % 6, [TSC, IDLE, EDGE, RACH, 0, 4]).astype(np.uint8)  # incl. OFF and SCH (invalid here)
    tscv = tsc.copy()
    tscv[::50] = 9  # unsupported TSC -> -SIGERR_UNSUPPORTED
    mt = np.choose(np.arange(n) % 4, [0, 4, 17, 63]).astype(np.int16)
    g, c, rep = check_dd(trx, checker, rx, typ, tscv, mt, "mixed")
    assert (c["rc"] == -3).any()


@pytest.mark.parametrize("cfg", DETECT_CFGS)
def test_mixed_types_nb_geometry(trx, checker, cfg):
    """Every burst type, unsupported TSCs and per-burst max_toa values inside the normal-burst kernels' geometry
    (16-symbol sequences, max_toa <= 4): under the 16-symbol hints this is the register-blocked / fused path."""
    rng = np.random.default_rng(14)
    n = 4200
    tsc = np.arange(n) % 8
    w = checker.modulate_gmsk_batch(synth.nb_bits(n, tsc, rng), nthreads=8)
    didx = np.arange(1, n, 6)[::2]  # half of the IDLE-typed slots carry the dummy burst detectDummyBurst looks for
    w[didx] = checker.modulate_gmsk_batch(np.array([[int(ch) for ch in synth.DUMMY_STR]], np.uint8))[0]
    eidx = np.arange(2, n, 6)[::2]  # half of the EDGE-typed slots carry real 8-PSK bursts
    w[eidx] = checker.modulate_edge_batch(synth.edge_bits(len(eidx), tsc[eidx], rng), nthreads=8)
    rx, _ = synth.impair(w, rng, snr_db=np.choose(np.arange(n) % 5, [30.0, 22.0, 15.0, 9.0, 5.0]), noise_only_frac=0.1,
                         full_scale=12000.0)
    rx[7::97] = (rng.standard_normal((len(rx[7::97]), 625, 2)) * 15000.0).astype(np.float32)  # clipping noise
    typ = np.choose(np.arange(n) % 6, [TSC, IDLE, EDGE, RACH, 0, 4]).astype(np.uint8)
    tscv = tsc.copy()
    tscv[::50] = 9
    mt = np.choose(np.arange(n) % 5, [0, 1, 2, 3, 4]).astype(np.int16)
    # threshold 2.5: the dummy burst's periodic midamble has a peak-to-average ratio below the transceiver's 4.0 even on
    # a clean signal, so only a lower threshold reaches detectDummyBurst's hit branch (and noise slots then hit as well)
    g, c, rep = check_dd(trx, checker, rx, typ, tscv, mt, f"mixed-nb{cfg}", cfg=cfg, thresh=2.5)
    assert (c["rc"] == -3).any() and (c["rc"] == -2).any() and (c["rc"] == IDLE).any() and (c["rc"] == EDGE).any()


def test_empty_and_tiny_batches(trx, checker):
    rng = np.random.default_rng(6)
    w = checker.modulate_gmsk_batch(synth.nb_bits(3, [0, 1, 2], rng))
    rx, _ = synth.impair(w, rng, snr_db=30.0)
    for n in (1, 3):
        check_dd(trx, checker, rx[:n], TSC, np.arange(n), 4, f"tiny{n}")
    r = trx.detect_demod(torch.zeros((0, 625, 2), device="cuda"), torch.zeros(0, dtype=torch.uint8, device="cuda"),
                         torch.zeros(0, dtype=torch.uint8, device="cuda"), torch.zeros(0, dtype=torch.int16, device="cuda"), 4)
    assert r["rc"].numel() == 0
    # all-zero bursts: nothing detected, no NaNs
    z = np.zeros((64, 625, 2), np.float32)
    g = run_gpu_dd(trx, z, TSC, 0, 4, 4)
    c = checker.detect_demod(z, TSC, 0, 4)
    assert np.array_equal(g["rc"], c["rc"]) and (g["rc"] == 0).all()


def test_convolve_golden(trx):
    """The reference's own KAT: tests/Transceiver52M/convolve_test_golden.h (fixture in tests/golden)."""
    import os
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "convolve_golden.npz"))
    x, h = gold["x"], gold["h"]
    for cplx in (False, True):
        for hl in (4, 8, 12, 16, 20, 24):
            ref = gold[f"y_{'complex' if cplx else 'real'}_base_{hl}"].reshape(-1, 2)
            start, ln = hl - 1, 100 - (hl - 1)
            for base in (False, True):
                rc, y = trx.convolve(dev(x.reshape(1, 100, 2)), 0, 100, dev(h.reshape(-1, 2)[:hl]), start, ln, cplx, base)
                assert rc == 0
                y = y.cpu().numpy()[0]
                ok = (np.abs(y - ref) < 1e-5) | (np.abs(1 - y / ref) < 1e-5)  # compare_floats, convolve_test.c:76-96
                assert ok.all(), (cplx, hl, base)


def test_convolve_exact_vs_checker(trx, checker):
    rng = np.random.default_rng(7)
    n, xl, head = 5, 300, 70
    x = rng.standard_normal((n, head + xl, 2)).astype(np.float32)
    for hl in (1, 4, 5, 8, 12, 16, 20, 24, 40, 64, 7):
        h = rng.standard_normal((hl, 2)).astype(np.float32)
        for start, ln in ((hl - 1, xl - hl + 1), (0, xl), (3, 100)):
            for cplx in (False, True):
                for base in (False, True):
                    rc, y = trx.convolve(dev(x), head, xl, dev(h), start, ln, cplx, base)
                    assert rc == 0
                    y = y.cpu().numpy()
                    for b in range(n):
                        f = {(False, False): checker.convolve_real, (True, False): checker.convolve_complex,
                             (False, True): checker.base_convolve_real, (True, True): checker.base_convolve_complex}[(cplx, base)]
                        _, yr = f(x[b], h, start, ln, x_off=head)
                        assert np.array_equal(y[b], yr), (hl, start, ln, cplx, base)
    # bounds failure mirrors the reference's -1
    rc, _ = trx.convolve(dev(x), head, xl, dev(h), 10, xl, False, True)
    assert rc == -5


def test_helpers(trx, checker):
    rng = np.random.default_rng(8)
    x = rng.standard_normal((16, 625, 2)).astype(np.float32)
    e = trx.energy_detect(dev(x), 80).cpu().numpy()
    for b in range(16):
        assert e[b] == np.float32(checker.energy_detect(x[b], 80))
    s = (rng.standard_normal(5000) * 1.5).astype(np.float32)
    assert np.array_equal(trx.vector_slicer(dev(s)).cpu().numpy(), checker.vector_slicer(s))
    delays = np.array([0.0, 0.005, -0.005, 3.3, -7.71, 12.999, -0.5, 1.0, -40.25, 80.6, 0.011, -0.011, 624.5, -700.0, 2.0, -3.0], np.float32)
    y = trx.delay_vector(dev(x), dev(delays)).cpu().numpy()
    for b in range(16):
        assert np.array_equal(y[b], checker.delay_vector(x[b], float(delays[b]))), delays[b]
    v = (rng.standard_normal(4096) * 20000).astype(np.float32)
    v[:6] = [0.5, 1.5, 2.5, -0.5, -1.5, 1e12]
    assert np.array_equal(trx.convert_float_short(dev(v), 1.7).cpu().numpy(), checker.convert_float_short(v, 1.7))
    # the x86 dispatcher for lengths that are not multiples of eight (scalar truncating tail) and the scalar routine
    v[4090:4096] = [0.5, 1.5, -0.5, -1.5, 40000.0, -40000.0]
    for ln in (4096, 4093, 4088, 7, 1):
        for mode in (1, 2):
            got = trx.convert_float_short(dev(v[:ln].copy()), 1.0, mode=mode).cpu().numpy()
            assert np.array_equal(got, checker.convert_float_short_mode(v[:ln], 1.0, mode)), (ln, mode)
    i16 = rng.integers(-32768, 32767, 4096).astype(np.int16)
    assert np.array_equal(trx.convert_short_float(dev(i16)).cpu().numpy(), checker.convert_short_float(i16))


def test_host_pipeline_matches_device_path(trx, checker):
    rng = np.random.default_rng(9)
    n = 40000  # > 2 chunks of the host pipeline
    tsc = (np.arange(n) % 8).astype(np.uint8)
    w = checker.modulate_gmsk_batch(synth.nb_bits(n, tsc, rng), nthreads=8)
    rx, _ = synth.impair(w, rng, snr_db=12.0, noise_only_frac=0.1)
    g = run_gpu_dd(trx, rx, TSC, tsc, 4, 4, n_soft=148, soft_stride=148)
    out = trx.alloc_results(n, 148, device="cpu")
    out = {k: v.pin_memory() for k, v in out.items()}
    trx.detect_demod_host(torch.from_numpy(rx).pin_memory(), torch.full((n,), TSC, dtype=torch.uint8).pin_memory(),
                          torch.from_numpy(tsc).pin_memory(), torch.full((n,), 4, dtype=torch.int16).pin_memory(), 4, out)
    for k in ("rc", "amp", "toa", "tsc", "ci", "flags"):
        assert np.array_equal(out[k].numpy(), g[k], equal_nan=(k == "ci")), k
    det = g["rc"] > 0
    assert np.array_equal(out["soft"].numpy()[det], g["soft"][det])


def test_vitac_nb_cfg4(trx, checker):
    """BASELINE configs[3] recipe: GMSK bursts through a random 4-tap multipath channel, MLSE equalised."""
    rng = np.random.default_rng(21)
    n = 4000
    tsc = (np.arange(n) % 8).astype(np.uint8)
    bits = synth.nb_bits(n, tsc, rng)
    w = synth.multipath(checker.modulate_gmsk_batch(bits, nthreads=8), rng)
    rx, _ = synth.impair(w, rng, snr_db=34.0, amp_range=(0.5, 1.0), shift_lo=-4, shift_hi=4)
    buf = np.zeros((n, 40 + 625 + 63, 2), np.float32)
    buf[:, 40:665] = rx
    c = checker.vitac(buf, 40, tsc, nthreads=8)
    g = trx.vitac(dev(buf), 40, dev(tsc), want_cir=True)
    torch.cuda.synchronize()
    assert np.array_equal(g["start"].cpu().numpy(), c["start"])
    assert np.array_equal(g["bits"].cpu().numpy(), c["bits"])
    assert np.array_equal(g["cir"].cpu().numpy(), c["cir"])
    cm = g["corr_max"].cpu().numpy()
    assert np.allclose(cm, c["corr_max"], rtol=1e-4, atol=0)
    ber = (((c["bits"] < 0).astype(np.uint8)) != bits).mean()
    print("vitac NB BER", ber, "start range", c["start"].min(), c["start"].max())
    assert ber < 0.01
    # odd batch size (pair kernel tail) and dummy-burst TSC index 8
    g1 = trx.vitac(dev(buf[:7]), 40, dev(np.full(7, 8, np.uint8)))
    c1 = checker.vitac(buf[:7], 40, np.full(7, 8, np.uint8))
    assert np.array_equal(g1["bits"].cpu().numpy(), c1["bits"]) and np.array_equal(g1["start"].cpu().numpy(), c1["start"])


def test_vitac_ab(trx, checker):
    rng = np.random.default_rng(22)
    n = 600
    ab = synth.ab_bits(n, 0, rng, 0)
    w = np.stack([checker.modulate_burst(ab[i], 68, 4) for i in range(n)])
    rx, _ = synth.impair(w, rng, snr_db=25.0, amp_range=(0.5, 1.0), shift_lo=-2, shift_hi=30)
    buf = np.zeros((n, 40 + 625 + 400, 2), np.float32)
    buf[:, 40:665] = rx
    for md in (0, 10):
        c = checker.vitac(buf, 40, 0, is_ab=True, max_delay=md, clamp=(0, 100))
        g = trx.vitac(dev(buf), 40, dev(np.zeros(n, np.uint8)), is_ab=True, max_delay=md, clamp=(0, 100))
        assert np.array_equal(g["start"].cpu().numpy(), c["start"]), md
        assert np.array_equal(g["bits"].cpu().numpy(), c["bits"]), md
        assert np.allclose(g["corr_max"].cpu().numpy(), c["corr_max"], rtol=1e-4, atol=0)


def test_resampler(trx, checker):
    import osmo_trx_b200
    rng = np.random.default_rng(23)
    for (p, q, bw, L) in [(1, 4, 1.0, 16), (65, 48, 1.0, 16), (48, 65, 1.0, 16), (65, 96, 1.0, 16), (52, 75, 0.45, 16), (3, 2, 1.0, 8), (2, 3, 1.0, 12)]:
        rs = osmo_trx_b200.Resampler(trx, p, q, L, bw)
        hc = checker.resampler(p, q, L, bw)
        # 4 periods (one reference block) and a long row that spans several tiles of the staged kernel
        for nblk, ns in ((4, 3), (min(16384 // p, 2500), 2)):
            x = rng.standard_normal((ns, L + q * nblk, 2)).astype(np.float32)
            y = rs.rotate(dev(x), p * nblk).cpu().numpy()
            for s in range(ns):
                rc, yr = checker.resampler_rotate(hc, x[s], L, p * nblk)
                assert rc == p * nblk and np.array_equal(y[s], yr), (p, q, s, nblk)


@pytest.mark.parametrize("m", [4, 64, 5, 8, 16])
def test_channelizer_synthesis(trx, checker, m):
    import osmo_trx_b200
    rng = np.random.default_rng(24)
    bl = 192
    ch, sy = osmo_trx_b200.Channelizer(trx, m, bl), osmo_trx_b200.Synthesis(trx, m, bl)
    cc, sc = checker.channelizer(m, bl), checker.synthesis(m, bl)
    for it, nb in enumerate((1, 3, 2)):  # history carried across calls, several blocks per call
        x = rng.standard_normal((nb * m * bl, 2)).astype(np.float32)
        y = ch.rotate(dev(x)).cpu().numpy()
        yr = np.concatenate([checker.channelizer_rotate(cc, x[k * m * bl:(k + 1) * m * bl], m, bl)[1] for k in range(nb)], axis=1)
        assert np.abs(y - yr).max() <= 1e-4 * np.abs(yr).max(), (m, it)
        xin = rng.standard_normal((m, nb * bl, 2)).astype(np.float32)
        s = sy.rotate(dev(xin)).cpu().numpy()
        sr = np.concatenate([checker.synthesis_rotate(sc, np.ascontiguousarray(xin[:, k * bl:(k + 1) * bl]), m, bl)[1] for k in range(nb)])
        assert np.abs(s - sr).max() <= 1e-4 * np.abs(sr).max(), (m, it)
    if m in (4, 16, 64):
        # one long call: more tiles than resident CTAs, so the kernels' grid-stride / carried-halo paths run
        nb = 100
        x = rng.standard_normal((nb * m * bl, 2)).astype(np.float32)
        y = ch.rotate(dev(x)).cpu().numpy()
        yr = np.concatenate([checker.channelizer_rotate(cc, x[k * m * bl:(k + 1) * m * bl], m, bl)[1] for k in range(nb)], axis=1)
        assert np.abs(y - yr).max() <= 1e-4 * np.abs(yr).max()
        xin = rng.standard_normal((m, nb * bl, 2)).astype(np.float32)
        s = sy.rotate(dev(xin)).cpu().numpy()
        sr = np.concatenate([checker.synthesis_rotate(sc, np.ascontiguousarray(xin[:, k * bl:(k + 1) * bl]), m, bl)[1] for k in range(nb)])
        assert np.abs(s - sr).max() <= 1e-4 * np.abs(sr).max()
    # synthesis -> channelizer loopback after reset(), against the checker's own loopback
    ch.reset(); sy.reset()
    cc2, sc2 = checker.channelizer(m, bl), checker.synthesis(m, bl)
    xin = rng.standard_normal((m, 2 * bl, 2)).astype(np.float32)
    back = ch.rotate(sy.rotate(dev(xin))).cpu().numpy()
    ref_back = []
    for k in range(2):
        wide = checker.synthesis_rotate(sc2, np.ascontiguousarray(xin[:, k * bl:(k + 1) * bl]), m, bl)[1]
        ref_back.append(checker.channelizer_rotate(cc2, wide, m, bl)[1])
    ref_back = np.concatenate(ref_back, axis=1)
    assert np.abs(back - ref_back).max() <= 2e-4 * np.abs(ref_back).max()


def test_detect_config_hint(trx, checker):
    """With the 16-symbol sizing hint TSC/EDGE/IDLE results are unchanged and RACH bursts fail loudly."""
    rng = np.random.default_rng(31)
    n = 512
    tsc = np.arange(n) % 8
    w = checker.modulate_gmsk_batch(synth.nb_bits(n, tsc, rng))
    rx, _ = synth.impair(w, rng, snr_db=15.0)
    typ = np.choose(np.arange(n) % 4, [TSC, EDGE, IDLE, RACH]).astype(np.uint8)
    try:
        trx.detect_config(16)
        g = run_gpu_dd(trx, rx, typ, tsc, 4, 4)
    finally:
        trx.detect_config(40)
    c = checker.detect_demod(rx, typ, tsc, 4)
    rach = typ == RACH
    assert (g["rc"][rach] == -1).all()  # -SIGERR_BOUNDS, not a silent wrong answer
    for k in ("rc", "toa", "amp", "tsc"):
        assert np.array_equal(g[k][~rach], c[k][~rach]), k


# ---- the receive chain around the hot path: int16 slots -> TRXD uplink datagrams ----
def to_i16(rx, scale):
    return np.clip(np.rint(rx * scale), -32768, 32767).astype(np.int16)


def run_pull(trx, checker, iq, typ, tsc, max_toa, fn, tn, what, version=1, pkt_stride=160, rssi_offset=7.25):
    n = iq.shape[0]
    bound = int(np.max(max_toa))
    typ = np.broadcast_to(np.asarray(typ, np.uint8), (n,)).copy()
    mt = np.broadcast_to(np.asarray(max_toa, np.uint16), (n,)).copy()
    out = trx.alloc_pull_results(n, pkt_stride)
    trx.pull(dev(iq), dev(typ), dev(tsc.astype(np.uint8)), dev(mt.astype(np.int16)), dev(fn.astype(np.int32)), dev(tn), bound,
             out=out, version=version, rssi_offset=rssi_offset)
    torch.cuda.synchronize()
    g = {k: v.cpu().numpy() for k, v in out.items()}
    g["pkt_len"] = g["pkt_len"].view(np.uint16)
    c = checker.pull(iq, typ, tsc, mt, fn, tn, version=version, rssi_offset=rssi_offset, pkt_stride=pkt_stride, nthreads=8)
    # the reference's float soft values (0..1) for the rounding-boundary rule
    x = iq.astype(np.float32)
    typ_det = np.where((typ == 0) | (typ == IDLE), 0, typ).astype(np.uint8)
    dd = checker.detect_demod(x, typ_det, tsc, mt, nthreads=8)
    soft01 = checker.vector_slicer(dd["soft"]).reshape(dd["soft"].shape)
    rep = parity.compare_pkts(g, c, soft01, what, version)
    ok = (g["rc"] == c["rc"]) & (c["rc"] > 0)
    assert np.array_equal(g["tsc"][ok], c["tsc"][ok])
    assert np.abs(g["toa"][ok].astype(np.float64) - c["toa"][ok]).max() <= 1.0 / 256 + 1e-9
    print(what, rep)
    return g, c, rep


def pull_inputs(checker, n, seed, kind="nb"):
    rng = np.random.default_rng(seed)
    tsc = (np.arange(n) % 8).astype(np.uint8)
    if kind == "edge":
        w = checker.modulate_edge_batch(synth.edge_bits(n, tsc, rng), nthreads=8)
        rx, _ = synth.impair(w, rng, snr_db=np.choose(np.arange(n) % 3, [45.0, 25.0, 21.0]), noise_only_frac=0.05)
    else:
        w = checker.modulate_gmsk_batch(synth.nb_bits(n, tsc, rng), nthreads=8)
        rx, _ = synth.impair(w, rng, snr_db=np.choose(np.arange(n) % 3, [30.0, 10.0, 6.0]), noise_only_frac=0.05)
    fn = rng.integers(0, 2715648, n).astype(np.uint32)
    tn = rng.integers(0, 8, n).astype(np.uint8)
    return rng, tsc, rx, fn, tn


@pytest.mark.parametrize("cfg", DETECT_CFGS)
def test_pull_nb_trxd_v1(trx, checker, cfg):
    n = 6000
    rng, tsc, rx, fn, tn = pull_inputs(checker, n, 31)
    iq = to_i16(rx, 8000.0)
    iq[100:110] = to_i16(rx[100:110], 40000.0)  # saturating slots: clipping report / detection on clipped data
    iq[200] = 0                                   # all-zero slot: energy 0, RSSI conversion corner
    iq[300:305] = rng.integers(-32768, 32767, (5, 625, 2)).astype(np.int16)  # undetectable and clipping: -SIGERR_CLIP
    typ = np.full(n, TSC, np.uint8)
    typ[::37] = IDLE
    typ[5::41] = 0  # OFF
    with detect_cfg(trx, cfg):
        g, c, rep = run_pull(trx, checker, iq, typ, tsc, 4, fn, tn, f"pull nb v1 {cfg}")
    assert rep["sent"] > 0.9 * n and (c["rc"] == -2).any() and (c["pkt_len"] == 0).any() and (c["pkt_len"] == 11).any()


@pytest.mark.parametrize("cfg", [(40, 3), (40, 1)])
def test_pull_trxd_v0_and_rach(trx, checker, cfg):
    n = 3000
    rng = np.random.default_rng(32)
    b = synth.ab_bits(n, 20, rng, 0)
    w = checker.modulate_gmsk_batch(b, nthreads=8)
    rx, _ = synth.impair(w, rng, snr_db=15.0, noise_only_frac=0.1)
    iq = to_i16(rx, 6000.0)
    fn = rng.integers(0, 2715648, n).astype(np.uint32)
    tn = rng.integers(0, 8, n).astype(np.uint8)
    typ = np.choose(np.arange(n) % 3, [RACH, EXT_RACH if cfg[1] == 3 else RACH, RACH]).astype(np.uint8)
    tsc = np.zeros(n, np.uint8)
    with detect_cfg(trx, cfg):
        g, c, rep = run_pull(trx, checker, iq, typ, tsc, 63, fn, tn, f"pull rach v0 {cfg}", version=0, pkt_stride=158)
    assert set(np.unique(c["pkt_len"])) == {0, 158}


@pytest.mark.parametrize("cfg", [(40, 3), (16, 2)])
def test_pull_edge_and_truncated_rows(trx, checker, cfg):
    n = 2000
    rng, tsc, rx, fn, tn = pull_inputs(checker, n, 33, "edge")
    wn = checker.modulate_gmsk_batch(synth.nb_bits(200, tsc[:200], rng))
    rx[:200], _ = synth.impair(wn, rng, snr_db=20.0)  # GMSK bursts in 8-PSK slots: the EDGE -> TSC fall-through
    iq = to_i16(rx, 8000.0)
    with detect_cfg(trx, cfg):
        g, c, rep = run_pull(trx, checker, iq, EDGE, tsc, 4, fn, tn, f"pull edge v1 {cfg}", pkt_stride=456)
    assert (c["rc"] == TSC).sum() > 150
    assert (c["pkt_len"] == 455).any()
    # rows that cannot hold an 8-PSK burst: flagged, nothing emitted, nothing overrun
    out = trx.alloc_pull_results(n, 160)
    out["pkt"].fill_(0xAA)
    trx.pull(dev(iq), dev(np.full(n, EDGE, np.uint8)), dev(tsc), dev(np.full(n, 4, np.int16)), dev(fn.astype(np.int32)), dev(tn), 4,
             out=out)
    torch.cuda.synchronize()
    fl, pl, rc = out["flags"].cpu().numpy(), out["pkt_len"].cpu().numpy(), out["rc"].cpu().numpy()
    assert ((fl & 8) != 0)[rc == 5].all() and (pl[rc == 5] == 0).all() and (out["pkt"].cpu().numpy()[rc == 5] == 0xAA).all()


@pytest.mark.parametrize("cfg", [(40, 3), (16, 1)])
def test_pull_host_matches_device_path(trx, checker, cfg):
    with detect_cfg(trx, cfg):
        _pull_host_matches_device_path(trx, checker)


def _pull_host_matches_device_path(trx, checker):
    n = 40000  # > 2 chunks of the host pipeline
    rng, tsc, rx, fn, tn = pull_inputs(checker, n, 34)
    iq = to_i16(rx, 8000.0)
    typ = np.full(n, TSC, np.uint8)
    typ[::29] = IDLE
    mt = np.full(n, 4, np.int16)
    out = trx.alloc_pull_results(n, 160)
    trx.pull(dev(iq), dev(typ), dev(tsc), dev(mt), dev(fn.astype(np.int32)), dev(tn), 4, out=out)
    torch.cuda.synchronize()
    h = {k: v.pin_memory() for k, v in trx.alloc_pull_results(n, 160, device="cpu").items()}
    trx.pull_host(torch.from_numpy(iq).pin_memory(), torch.from_numpy(typ), torch.from_numpy(tsc), torch.from_numpy(mt),
                  torch.from_numpy(fn.astype(np.int32)), torch.from_numpy(tn), 4, h)
    for k in out:
        assert np.array_equal(out[k].cpu().numpy(), h[k].numpy(), equal_nan=True), k
    assert trx.pull(dev(iq[:0]), dev(typ[:0]), dev(tsc[:0]), dev(mt[:0]), dev(fn[:0].astype(np.int32)), dev(tn[:0]), 4) is not None


def test_pull_slots_off_the_16_byte_grid(trx, checker):
    """int16 slot arrays that start 4, 8 or 12 bytes off the 16-byte grid, and slot counts that are not a multiple of four:
    detect_lane_kernel<int16> describes the slots to the TMA four at a time from the grid point in front of the array (the last
    partial tile takes the 4-byte copies).  Detection outputs, energies and header bytes must equal the aligned run's bit for
    bit; the soft bytes may differ by one step where the demodulator's other summation order crosses a rounding boundary."""
    for n in (4096, 4096 + 37):
        rng, tsc, rx, fn, tn = pull_inputs(checker, n, 35)
        iq = to_i16(rx, 8000.0)
        typ = np.full(n, TSC, np.uint8)
        typ[::31] = IDLE
        args = [dev(typ), dev(tsc), dev(np.full(n, 4, np.int16)), dev(fn.astype(np.int32)), dev(tn), 4]
        with detect_cfg(trx, (16, 1)):
            a = {k: v.cpu().numpy() for k, v in trx.pull(dev(iq), *args).items()}
            flat = torch.zeros((n * 625 + 4, 2), dtype=torch.int16, device="cuda")
            for off in (1, 2, 3):
                view = flat[off:off + n * 625].view(n, 625, 2)
                view.copy_(dev(iq))
                assert (view.data_ptr() >> 2) & 3 == off
                b = {k: v.cpu().numpy() for k, v in trx.pull(view, *args).items()}
                for k in a:
                    if k == "pkt":
                        d = np.abs(a[k].astype(np.int32) - b[k].astype(np.int32))
                        assert d.max() <= 1 and (d[:, :11] == 0).all(), (n, off, k)
                    else:
                        assert np.array_equal(a[k], b[k], equal_nan=True), (n, off, k)
        assert (a["rc"] > 0).sum() > 0.8 * n


def test_scheduler_expected_corr_type(trx, checker):
    """Burst-type scheduler (Transceiver::expectedCorrType + the max_toa choice of pullRadioVector) on the device:
    several channels with different timeslot configurations in one batch, against the CPU checker per channel."""
    rng = np.random.default_rng(41)
    n_chan, n = 6, 40000
    ct = rng.integers(0, 16, (n_chan, 8)).astype(np.uint8)
    ct[0] = [4, 7, 1, 2, 3, 13, 5, 15]  # a realistic C0: IV, VII, TCH/F, TCH/H x2, PDCH, V, loopback
    ho = rng.integers(0, 256, 8).astype(np.uint8)
    fn = rng.integers(0, 2715648, n).astype(np.uint32)
    tn = rng.integers(0, 8, n).astype(np.uint8)
    chan = rng.integers(0, n_chan + 1, n).astype(np.int16)  # n_chan itself: out of range -> OFF
    for xr, eg in ((0, 0), (1, 0), (0, 1), (1, 1)):
        typ, mt = trx.expected_corr_type(dev(fn.astype(np.int32)), dev(tn), dev(ct), dev(ho), chan=dev(chan), ext_rach=xr, egprs=eg,
                                         max_toa_nb=4, max_toa_ab=63)
        torch.cuda.synchronize()
        typ, mt = typ.cpu().numpy(), mt.cpu().numpy()
        want = np.zeros(n, np.uint8)
        for c in range(n_chan):
            m = chan == c
            want[m] = checker.expected_corr_type(ct[c], ho, xr, eg, fn[m], tn[m])
        assert np.array_equal(typ, want), (xr, eg)
        assert np.array_equal(mt, np.where((want == RACH) | (want == EXT_RACH), 63, 4))
    # no channel array: everything is channel 0; empty batch
    typ, _ = trx.expected_corr_type(dev(fn.astype(np.int32)), dev(tn), dev(ct), dev(ho))
    assert np.array_equal(typ.cpu().numpy(), checker.expected_corr_type(ct[0], ho, 0, 0, fn, tn))
    typ, mt = trx.expected_corr_type(dev(fn[:0].astype(np.int32)), dev(tn[:0]), dev(ct), dev(ho))
    assert typ.numel() == 0 and mt.numel() == 0


def test_wide_window_with_short_sequences(trx, checker):
    """16-symbol sizing hint with search windows wider than corr_nb_kernel's (max_toa up to 17): the long-window
    correlator runs with a 16-tap sequence and a small decimated row."""
    rng = np.random.default_rng(43)
    n = 3000
    tsc = np.arange(n) % 8
    w = checker.modulate_gmsk_batch(synth.nb_bits(n, tsc, rng))
    rx, _ = synth.impair(w, rng, snr_db=14.0, shift_lo=-40, shift_hi=20, noise_only_frac=0.05)
    typ = np.choose(np.arange(n) % 3, [TSC, EDGE, IDLE]).astype(np.uint8)
    mt = np.choose(np.arange(n) % 4, [0, 4, 9, 17]).astype(np.int16)
    try:
        trx.detect_config(16, 2)
        g = run_gpu_dd(trx, rx, typ, tsc, mt, 17)
    finally:
        trx.detect_config(40, 3)
    c = checker.detect_demod(rx, typ, tsc, mt)
    rep = parity.compare_detect(g, c, g["flags"], "wide16")
    parity.compare_soft(g["soft"], c["soft"], rep["ok_mask"] & (c["rc"] != EDGE), 156, "wide16")
    assert rep["detected"] > 1500


def test_detect_rows_off_the_16_byte_grid(trx):
    """A burst array that starts on an odd sample (8 bytes off the 16-byte grid: slots addressed in place in a resampled stream)
    takes detect_lane_kernel's TMA path through a tensor that starts one sample in front of it; an odd row count leaves the last
    tile to the per-row copies.  Every detection output must equal the aligned run's, bit for bit."""
    rng = np.random.default_rng(77)
    ref = cpulibs.Ref() if cpulibs.Ref.available() else cpulibs.Oracle()
    for n in (2048, 2049 + 32):
        tsc = np.arange(n) % 8
        w = ref.modulate_gmsk_batch(synth.nb_bits(n, tsc, rng))
        rx, _ = synth.impair(w, rng, snr_db=12.0, noise_only_frac=0.1)
        typ = np.choose(np.arange(n) % 2, [TSC, EDGE]).astype(np.uint8)
        args = [dev(typ), dev(tsc.astype(np.uint8)), dev(np.full(n, 4, np.int16)), 4]
        try:
            trx.detect_config(16, 2)
            a = {k: v.cpu().numpy() for k, v in trx.detect_demod(dev(rx), *args).items()}
            flat = torch.zeros((n * 625 + 3, 2), dtype=torch.float32, device="cuda")
            for off in (1, 3):
                view = flat[off:off + n * 625].view(n, 625, 2)
                view.copy_(dev(rx))
                assert (view.data_ptr() >> 3) & 1 == 1
                b = {k: v.cpu().numpy() for k, v in trx.detect_demod(view, *args).items()}
                for k in a:
                    if k == "soft":
                        # the demodulator folds the row's phase on the 16-byte grid into its tap index (another summation
                        # order): soft values carry the 1e-4 tolerance, detection outputs are exact
                        det = a["rc"] > 0
                        assert np.abs(a[k][det] - b[k][det]).max() <= 1e-5 * np.abs(a[k][det]).max(), (n, off, k)
                    else:
                        assert np.array_equal(a[k], b[k], equal_nan=True), (n, off, k)
        finally:
            trx.detect_config(40, 3)
        assert (a["rc"] > 0).sum() > 0.8 * n


@pytest.mark.parametrize("blen", [157, 156])
def test_one_sample_per_symbol(trx, checker, blen):
    """The rx_sps = 1 path (detectAnyBurst / demodAnyBurst with sps = 1, sigProcLib.cpp:1659-1662, 2038-2042): normal,
    8-PSK (with its TSC fall-through), access and idle slots of 156 / 157 symbols; decisions exact, TOA / amp / soft bits
    per the usual criteria; computeEdgeCI over blen - 16 symbols."""
    rng = np.random.default_rng(80 + blen)
    n = 3000
    rx, tsc, is_edge = synth.sps1_bursts(checker, n, rng, blen=blen, edge_every=4, snr_db=np.choose(np.arange(n) % 3, [25.0, 12.0, 7.0]))
    typ = np.where(is_edge, EDGE, TSC).astype(np.uint8)
    typ[::7] = np.choose(np.arange(len(typ[::7])) % 3, [IDLE, RACH, EXT_RACH])
    typ[5::50] = 0
    k = len(rx[11::100])  # slots that clip and hold no burst: -SIGERR_CLIP
    rx[11::100] = (rng.standard_normal((k, blen, 2)) * 9000.0).astype(np.float32)
    rx[11::100, 60, 0] = 31000.0
    mt = np.choose(np.arange(n) % 3, [4, 0, 9]).astype(np.int16)
    c = checker.detect_demod(rx, typ, tsc, mt.astype(np.uint16), sps=1, blen=blen, nthreads=8)
    r = trx.detect_demod_sps1(dev(rx), dev(typ), dev(tsc), dev(mt), 9, n_gmsk_soft=blen, soft_stride=444)
    torch.cuda.synchronize()
    g = {k: v.cpu().numpy() for k, v in r.items()}
    rep = parity.compare_detect(g, c, g["flags"], f"sps1/{blen}")
    ok = rep["ok_mask"]
    parity.compare_ci(g["ci"], c["ci"], ok, f"sps1/{blen}")
    parity.compare_soft(g["soft"], c["soft"], ok & (c["rc"] != EDGE), blen, f"sps1/{blen} gmsk")
    parity.compare_soft(g["soft"], c["soft"], ok & (c["rc"] == EDGE), 444, f"sps1/{blen} edge")
    assert (c["rc"] == TSC).sum() > 1200 and (c["rc"] == EDGE).sum() > 100 and (c["rc"] == -2).sum() > 5 and (c["rc"] == 0).sum() > 50
    print(f"sps1/{blen}", {k: v for k, v in rep.items() if k not in ("ok_mask", "_ga")}, np.bincount(c["rc"] + 4))


def test_detect_sch_full(trx, checker):
    """detectSCHBurst in its single-burst state (SCH_DETECT_FULL): 64-symbol sequence, 156 correlation outputs whose
    window reaches before the burst (zeros).  Decisions exact, TOA / amp per the usual criteria."""
    rng = np.random.default_rng(45)
    n = 1200
    w = checker.modulate_gmsk_batch(synth.sch_bits(n, rng))
    rx, _ = synth.impair(w, rng, snr_db=np.choose(np.arange(n) % 3, [25.0, 10.0, 5.0]), noise_only_frac=0.1, shift_lo=-60, shift_hi=30)
    c = checker.detect_sch(rx)
    g = {k: v.cpu().numpy() for k, v in trx.detect_sch(dev(rx)).items()}
    assert (c["rc"] > 0).sum() > 0.8 * n and (c["rc"] == 0).sum() > 50
    c["tsc"] = np.zeros(n, np.uint8)
    g["tsc"] = np.zeros(n, np.uint8)
    rep = parity.compare_detect(g, c, g["flags"], "sch")
    parity.compare_ci(g["ci"], c["ci"], rep["ok_mask"], "sch")
    print("sch", {k: v for k, v in rep.items() if k not in ("ok_mask", "_ga")})
    # an ordinary detect call cannot select the internal SCH type through its type array
    r = run_gpu_dd(trx, rx[:64], 7, 0, 4, 4)
    assert (r["rc"] <= 0).all()


def test_sch_first_acquisition(trx, checker):
    """The MS side's first SCH acquisition over a 12-frame capture: detectSCHBurst(SCH_DETECT_BUFFER)
    - 15,000 decimated positions x 64 taps - and get_sch_buffer_chan_imp_resp (59,488 windows x 54 symbols, running window
    energy) + detect_burst_nb.  rc / TOA / amp / C/I per the usual criteria; start, taps, corr_max and all decisions exact.
    A short capture and the committed fixture as well."""
    import os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_sch_buffer_fixture as mk
    rng = np.random.default_rng(49)
    for n, length, head in ((40, 60000, 192), (7, 2500, 192)):
        buf, pos = mk.captures(checker, rng, n, length, head)
        cap = np.ascontiguousarray(buf[:, head:head + length])
        if length == 60000:
            c = checker.detect_sch_buffer(cap, length)  # the reference's BUFFER state always reads 12 frames
        else:
            if checker.pfx != "orc_":
                checker_d = cpulibs.Oracle()           # length-generic restatement (pinned to the reference at 60,000)
            else:
                checker_d = checker
            c = checker_d.detect_sch_buffer(cap, length)
        g = {k: v.cpu().numpy() for k, v in trx.detect_sch_buffer(dev(cap), length).items()}
        assert np.array_equal(g["rc"], c["rc"]), (g["rc"], c["rc"])
        hit = c["rc"] > 0
        assert hit.sum() >= n // 2
        assert np.abs(g["toa"][hit].astype(np.float64) - c["toa"][hit]).max() <= 1e-3
        assert np.abs(g["amp"][hit] - c["amp"][hit]).max() <= 1e-4 * np.abs(c["amp"][hit]).max()
        assert np.abs(g["ci"][hit] - c["ci"][hit]).max() <= 1e-2
        assert np.all(g["toa"][~hit] == 0) and np.all(g["amp"][~hit] == 0)
        exact = np.array_equal(g["toa"], c["toa"]) and np.array_equal(g["amp"], c["amp"])
        cv = checker.vitac_sch_buffer(buf, head, length)
        gv = trx.vitac_sch_buffer(dev(buf), head, length)
        torch.cuda.synchronize()
        for k in ("start", "corr_max", "cir", "bits"):
            assert np.array_equal(gv[k].cpu().numpy().view(np.uint8), cv[k].view(np.uint8)), (k, length)
        print("sch acquisition", length, "detected", int(hit.sum()), "of", n, "toa/amp bit-identical:", exact)
    fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sch_buffer_fixture.npz"))
    x, head, length = fx["buf"].astype(np.float32), int(fx["head"]), int(fx["length"])
    g = {k: v.cpu().numpy() for k, v in trx.detect_sch_buffer(dev(fx["cap"].astype(np.float32)), 60000).items()}
    assert np.array_equal(g["rc"], fx["d_rc"]) and np.abs(g["toa"] - fx["d_toa"]).max() <= 1e-3
    gv = trx.vitac_sch_buffer(dev(x), head, length)
    for k in ("start", "corr_max", "cir", "bits"):
        assert np.array_equal(gv[k].cpu().numpy().view(np.uint8), fx["v_" + k].view(np.uint8)), k


def test_vitac_sch(trx, checker):
    """SCH burst of a tracked cell through the MLSE: get_sch_chan_imp_resp (54-symbol search over 40 symbols) +
    detect_burst_nb.  Start, CIR and every decision exact."""
    rng = np.random.default_rng(47)
    n = 1500
    bits = synth.sch_bits(n, rng)
    w = synth.multipath(checker.modulate_gmsk_batch(bits, nthreads=8), rng)
    rx, _ = synth.impair(w, rng, snr_db=30.0, amp_range=(0.5, 1.0), shift_lo=-6, shift_hi=6)
    buf = np.zeros((n, 40 + 625 + 63, 2), np.float32)
    buf[:, 40:665] = rx
    c = checker.vitac(buf, 40, 0, is_ab=2, nthreads=8)
    g = trx.vitac(dev(buf), 40, None, is_ab=2, want_cir=True)
    torch.cuda.synchronize()
    assert np.array_equal(g["start"].cpu().numpy(), c["start"])
    assert np.array_equal(g["bits"].cpu().numpy(), c["bits"])
    assert np.array_equal(g["cir"].cpu().numpy(), c["cir"])
    ber = (((c["bits"] < 0).astype(np.uint8)) != bits).mean()
    print("vitac SCH BER", ber, "start range", c["start"].min(), c["start"].max())
    assert ber < 0.02


def test_vitac_detect_with_given_cir(trx, checker):
    """detect_burst_nb / detect_burst_ab on their own: channel estimate and start supplied by the caller (here: the
    estimate of one call, the start moved by a sample for half of the bursts), every decision exact."""
    rng = np.random.default_rng(48)
    n = 800
    tsc = (np.arange(n) % 8).astype(np.uint8)
    bits = synth.nb_bits(n, tsc, rng)
    w = synth.multipath(checker.modulate_gmsk_batch(bits, nthreads=8), rng)
    rx, _ = synth.impair(w, rng, snr_db=30.0, amp_range=(0.5, 1.0), shift_lo=-4, shift_hi=4)
    buf = np.zeros((n, 40 + 625 + 63, 2), np.float32)
    buf[:, 40:665] = rx
    est = checker.vitac(buf, 40, tsc, nthreads=8)
    start = est["start"] + (np.arange(n) % 2).astype(np.int32)
    want = checker.vitac_detect(buf, 40, est["cir"], start)
    got = trx.vitac_detect(dev(buf), 40, dev(est["cir"]), dev(start)).cpu().numpy()
    assert np.array_equal(got, want)
    assert np.array_equal(want[::2], est["bits"][::2])  # unchanged start: the one-call result
    # the five-argument forms: other start states of the Viterbi detector
    for ss in (0, 7, 15):
        w2 = checker.vitac_detect(buf[:200], 40, est["cir"][:200], start[:200], ss=ss)
        g2 = trx.vitac_detect(dev(buf[:200]), 40, dev(est["cir"][:200]), dev(start[:200]), ss=ss).cpu().numpy()
        assert np.array_equal(g2, w2), ss
    # access bursts
    ab = synth.ab_bits(300, 5, rng, 0)
    wa = checker.modulate_gmsk_batch(ab, nthreads=8)
    rxa, _ = synth.impair(wa, rng, snr_db=25.0, amp_range=(0.5, 1.0), shift_lo=-2, shift_hi=2)
    bufa = np.zeros((300, 40 + 625 + 63, 2), np.float32)
    bufa[:, 40:665] = rxa
    ea = checker.vitac(bufa, 40, 0, is_ab=True, max_delay=20, nthreads=8)
    wa2 = checker.vitac_detect(bufa, 40, ea["cir"], ea["start"], is_ab=True)
    ga2 = trx.vitac_detect(dev(bufa), 40, dev(ea["cir"]), dev(ea["start"]), is_ab=True).cpu().numpy()
    assert np.array_equal(ga2, wa2) and np.array_equal(wa2, ea["bits"])
