"""The C++ mirror of the reference's per-burst API (osmo_trx_b200/host): builds and exports the reference's
symbols (CPU), fails loudly without a GPU (CPU), and reproduces the checker's results when driven the way
Transceiver.cpp / burst-gen.cpp drive sigProcLib (GPU)."""
import os
import struct
import subprocess

import numpy as np
import pytest

import cpulibs
import parity
import synth
from cpulibs import TSC, RACH, EDGE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "osmo_trx_b200", "host")


@pytest.fixture(scope="module")
def host_build():
    import osmo_trx_b200.buildlib as b
    b.build()
    subprocess.run(["make", "-s", "-C", HOST], check=True)
    return os.path.join(HOST, "host_demo")


def test_mirror_exports_reference_api(host_build):
    out = subprocess.run(["nm", "-D", "--defined-only", "-C", os.path.join(HOST, "libsigproc_b200.so")], capture_output=True,
                         text=True, check=True).stdout
    for sym in ["sigProcLibSetup()", "sigProcLibDestroy()", "vectorSlicer(float*, float const*, unsigned long)",
                "modulateBurst(BitVector const&, int, int, bool)", "modulateEdgeBurst(BitVector const&, int, bool)",
                "generateEdgeBurst(int)", "generateEmptyBurst(int, int)", "genRandNormalBurst(int, int, int)",
                "genRandAccessBurst(int, int, int)", "generateDummyBurst(int, int)", "scaleVector(signalVector&, Complex<float>)",
                "delayVector(signalVector const*, signalVector*, float)", "energyDetect(signalVector const&, unsigned int)",
                "detectAnyBurst(signalVector const&, unsigned int, float, int, CorrType, unsigned int, estim_burst_params*)",
                "demodAnyBurst(signalVector const&, CorrType, int, estim_burst_params*)",
                "convolve_real", "convolve_complex", "base_convolve_real", "base_convolve_complex", "convolve_init", "convolve_h_alloc",
                "Resampler::rotate(float const*, unsigned long, float*, unsigned long)", "Resampler::init(float)",
                "Channelizer::rotate(float const*, unsigned long)", "Channelizer::outputBuffer(unsigned long) const",
                "Synthesis::rotate(float*, unsigned long)", "Synthesis::inputBuffer(unsigned long) const", "ChannelizerBase::init()",
                "initvita()", "get_norm_chan_imp_resp(std::complex<float> const*, std::complex<float>*, float*, int)",
                "detect_burst_nb(std::complex<float> const*, std::complex<float>*, int, signed char*)",
                "detect_burst_ab(std::complex<float> const*, std::complex<float>*, int, signed char*)"]:
        assert sym in out, f"missing symbol {sym}"


def test_mirror_fails_loudly_without_gpu(host_build, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    fi = tmp_path / "in.bin"
    fi.write_bytes(struct.pack("<i", 0))
    r = subprocess.run([host_build, "dd", str(fi), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU path" in r.stderr


def _run(demo, mode, payload, tmp_path):
    fi, fo = tmp_path / f"{mode}_in.bin", tmp_path / f"{mode}_out.bin"
    fi.write_bytes(payload)
    r = subprocess.run([demo, mode, str(fi), str(fo)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    return fo.read_bytes()


@pytest.mark.gpu
def test_mirror_modulate_detect_demod(host_build, checker, tmp_path):
    rng = np.random.default_rng(77)
    n = 96
    tsc = (np.arange(n) % 8).astype(np.uint8)
    nb = synth.nb_bits(n, tsc, rng)
    ed = synth.edge_bits(n, tsc, rng)
    # modulators, per burst, bit-exact
    pay = struct.pack("<i", 2 * n)
    for k in range(n):
        pay += struct.pack("<ii", 0, 148) + nb[k].tobytes() + struct.pack("<ii", 1, 444) + ed[k].tobytes()
    w = np.frombuffer(_run(host_build, "mod", pay, tmp_path), np.float32).reshape(2 * n, 625, 2)
    assert np.array_equal(w[0::2], checker.modulate_gmsk_batch(nb))
    assert np.array_equal(w[1::2], checker.modulate_edge_batch(ed))
    # detect + demod per burst: NB, EDGE and RACH typed
    rx_nb, _ = synth.impair(w[0::2].copy(), rng, snr_db=14.0, noise_only_frac=0.1)
    rx_ed, _ = synth.impair(w[1::2].copy(), rng, snr_db=30.0)
    ab = synth.ab_bits(n, 7, rng, 0)
    rx_ab, _ = synth.impair(checker.modulate_gmsk_batch(ab), rng, snr_db=15.0)
    rx = np.concatenate([rx_nb, rx_ed, rx_ab])
    typ = np.concatenate([np.full(n, TSC), np.full(n, EDGE), np.full(n, RACH)]).astype(np.uint8)
    tscs = np.concatenate([tsc, tsc, np.zeros(n, np.uint8)])
    mt = np.concatenate([np.full(n, 4), np.full(n, 4), np.full(n, 63)]).astype(np.uint16)
    pay = struct.pack("<i", len(rx))
    for k in range(len(rx)):
        pay += struct.pack("<iii", int(typ[k]), int(tscs[k]), int(mt[k])) + rx[k].tobytes()
    rec = np.frombuffer(_run(host_build, "dd", pay, tmp_path), np.float32).reshape(len(rx), 6 + 1 + 444)
    g = dict(rc=rec[:, 0].astype(np.int32), amp=rec[:, 1:3].copy(), toa=rec[:, 3].copy(), tsc=rec[:, 4].astype(np.uint8),
             ci=rec[:, 5].copy(), soft=rec[:, 7:].copy())
    c = checker.detect_demod(rx, typ, tscs, mt)
    rep = parity.compare_detect(g, c, None, "host mirror")
    ok = rep["ok_mask"]
    assert np.array_equal(g["toa"][ok], c["toa"][ok]) and np.array_equal(g["amp"][ok], c["amp"][ok])
    parity.compare_soft(g["soft"], c["soft"], ok & (c["rc"] != EDGE), 156, "host mirror gmsk")
    parity.compare_soft(g["soft"], c["soft"], ok & (c["rc"] == EDGE), 444, "host mirror edge")
    nsoft = rec[:, 6].view(np.int32)
    assert (nsoft[c["rc"] == EDGE] == 444).all() and (nsoft[(c["rc"] > 0) & (c["rc"] != EDGE)] == 156).all()
    assert rep["detected"] > 2 * n


@pytest.mark.gpu
def test_mirror_convolve_golden(host_build, tmp_path):
    """The reference's own KAT (tests/Transceiver52M/convolve_test_golden.h) through the C symbols of convolve.h."""
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "convolve_golden.npz"))
    x, h = gold["x"].astype(np.float32), gold["h"].astype(np.float32)
    cases = []
    pay = b""
    for cplx in (0, 1):
        for hl in (4, 8, 12, 16, 20, 24):
            for base in (0, 1):
                start, ln = hl - 1, 100 - (hl - 1)
                pay += struct.pack("<7i", 100, hl, start, ln, cplx, base, 0) + x.reshape(-1)[:200].tobytes() + \
                    h.reshape(-1, 2)[:hl].tobytes()
                cases.append((cplx, hl, ln))
    out = _run(host_build, "conv", struct.pack("<i", len(cases)) + pay, tmp_path)
    off = 0
    for cplx, hl, ln in cases:
        rc = struct.unpack_from("<i", out, off)[0]
        y = np.frombuffer(out, np.float32, 2 * ln, off + 4).reshape(-1, 2)
        off += 4 + 8 * ln
        ref = gold[f"y_{'complex' if cplx else 'real'}_base_{hl}"].reshape(-1, 2)
        assert rc == ln
        ok = (np.abs(y - ref) < 1e-5) | (np.abs(1 - y / ref) < 1e-5)  # compare_floats, convolve_test.c:76-96
        assert ok.all(), (cplx, hl)


@pytest.mark.gpu
def test_mirror_vitac(host_build, checker, tmp_path):
    rng = np.random.default_rng(78)
    n = 64
    tsc = (np.arange(n) % 8).astype(np.uint8)
    bits = synth.nb_bits(n, tsc, rng)
    w = synth.multipath(checker.modulate_gmsk_batch(bits), rng)
    rx, _ = synth.impair(w, rng, snr_db=34.0, amp_range=(0.5, 1.0), shift_lo=-4, shift_hi=4)
    pay = struct.pack("<i", n)
    for k in range(n):
        pay += struct.pack("<i", int(tsc[k])) + rx[k].tobytes()
    out = _run(host_build, "vit", pay, tmp_path)
    rec = np.frombuffer(out, np.uint8).reshape(n, 8 + 148)
    start = rec[:, :4].copy().view(np.int32)[:, 0]
    cmax = rec[:, 4:8].copy().view(np.float32)[:, 0]
    hard = rec[:, 8:].view(np.int8)
    buf = np.zeros((n, 40 + 625 + 63, 2), np.float32)
    buf[:, 40:665] = rx
    c = checker.vitac(buf, 40, tsc)
    assert np.array_equal(start, c["start"]) and np.array_equal(hard, c["bits"])
    assert np.allclose(cmax, c["corr_max"], rtol=1e-4, atol=0)


@pytest.mark.gpu
def test_mirror_detect_sch(host_build, checker, tmp_path):
    """detectSCHBurst(SCH_DETECT_FULL) through the C++ mirror, per burst, against the CPU checker."""
    rng = np.random.default_rng(79)
    n = 64
    w = checker.modulate_gmsk_batch(synth.sch_bits(n, rng))
    rx, _ = synth.impair(w, rng, snr_db=12.0, noise_only_frac=0.15, shift_lo=-60, shift_hi=30)
    pay = struct.pack("<i", n) + b"".join(rx[k].tobytes() for k in range(n))
    rec = np.frombuffer(_run(host_build, "sch", pay, tmp_path), np.float32).reshape(n, 5)
    c = checker.detect_sch(rx)
    assert np.array_equal(rec[:, 0].astype(np.int32), c["rc"]) and (c["rc"] > 0).sum() > n // 2
    assert np.array_equal(rec[:, 1:3], c["amp"]) and np.array_equal(rec[:, 3], c["toa"])
    det = c["rc"] > 0
    assert np.allclose(rec[det, 4], c["ci"][det], rtol=1e-4, atol=1e-3)
