"""BASELINE.json's full sizes on the GPU: the CPU checker cannot run millions of bursts in a test, so the large
batches are checked through size-independent properties plus exact parity on a random subset.

* subset parity: a few thousand bursts drawn at random from the big batch go through the CPU checker
  (the same criteria as tests/parity.py everywhere else);
* power-of-two scaling: multiplying every sample by 2 is exact in float32, and every stage of the chain is
  homogeneous, so rc / TSC / TOA / C/I / soft bits must come back bit-identical and amp exactly doubled - except
  where the detection threshold's additive 1e-5 (sigProcLib.cpp:1541-1571) decides a burst sitting on the threshold;
* batch independence: the two halves processed separately equal the whole batch, bit for bit;
* sanity of the decisions against the generator (detected fraction, TOA inside the search window).
"""
import numpy as np
import pytest
import torch

import bench
import parity
from cpulibs import TSC, RACH, EDGE

pytestmark = pytest.mark.gpu

N = 1 << 20
SUBSET = 20000  # bursts of each 2^20 batch that also go through the CPU checker


def beq(a, b):
    """bitwise equality (NaNs of undetected bursts compare equal to themselves)"""
    if a.dtype == torch.float32:
        a, b = a.contiguous().view(torch.int32), b.contiguous().view(torch.int32)
    return bool(torch.equal(a, b))


def _run(trx, rx, typ, tsc, mt, bound, soft):
    out = trx.alloc_results(rx.shape[0], soft)
    trx.detect_demod(rx, typ, tsc, mt, bound, n_gmsk_soft=148, out=out)
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("kind,batches", [("nb", 1), ("rach", 1), ("edge", 10)])
def test_fullsize_detect_demod(trx, checker, kind, batches):
    """cfg1 recipe at GPU scale (2^20), cfg2 (1M access bursts), cfg3 (10M EDGE bursts as 10 batches of 2^20)."""
    soft = 444 if kind == "edge" else 148
    cfg = {"nb": (16, 1), "rach": (40, 1), "edge": (16, 2)}[kind]
    rng = np.random.default_rng(77)
    try:
        trx.detect_config(*cfg)
        for bi in range(batches):
            rx, typ, tsc, mt, bound = bench.make_workload(trx, kind, N, seed=500 + bi, device=trx.device)
            full = _run(trx, rx, typ, tsc, mt, bound, soft)
            rc = full["rc"].cpu().numpy()
            toa = full["toa"].cpu().numpy()
            det = rc > 0
            # the generator kills 5 % of the bursts and draws SNRs 30/10/6 dB (+15 dB for EDGE)
            assert 0.90 < det.mean() <= 0.9501 + 0.002, (kind, det.mean())
            assert (rc[det] == {"nb": TSC, "rach": RACH, "edge": EDGE}[kind]).mean() > 0.999
            assert np.all(np.abs(toa[det]) <= bound + 12.0)  # search window: head 10 (8 for RACH), tail 6 + max_toa
            assert not np.isnan(full["soft"][:1024].cpu().numpy()).any()

            # ---- exact parity on a random subset, through the CPU checker ----
            idx = np.sort(rng.choice(N, SUBSET, replace=False))
            tidx = torch.from_numpy(idx).to(trx.device)
            sub = {k: v[tidx].cpu().numpy() for k, v in full.items()}
            c = checker.detect_demod(rx[tidx].cpu().numpy(), typ[tidx].cpu().numpy(), tsc[tidx].cpu().numpy(),
                                     mt[tidx].cpu().numpy().astype(np.int64), nthreads=8)
            rep = parity.compare_detect(sub, c, sub["flags"], f"{kind} full-size subset")
            ok = rep["ok_mask"]
            parity.compare_ci(sub["ci"], c["ci"], ok, kind)
            # 8-PSK bursts carry 444 soft values, GMSK ones (incl. the EDGE -> TSC fall-through) the 148 asked for
            sr = parity.compare_soft(sub["soft"], c["soft"], ok & (c["rc"] != EDGE), 148, kind + " gmsk")
            se = parity.compare_soft(sub["soft"], c["soft"], ok & (c["rc"] == EDGE), 444, kind + " 8psk")
            print(kind, bi, {k: v for k, v in rep.items() if k not in ("ok_mask", "_ga")}, sr, se)

            if bi == 0:
                # ---- batch independence ----
                h = N // 2 + 13  # odd split: different warp / tile boundaries
                a = _run(trx, rx[:h], typ[:h], tsc[:h], mt[:h], bound, soft)
                b = _run(trx, rx[h:], typ[h:], tsc[h:], mt[h:], bound, soft)
                for k in ("rc", "toa", "amp", "tsc", "ci", "soft"):
                    assert beq(full[k][:h], a[k]) and beq(full[k][h:], b[k]), (kind, "split", k)
                del a, b

                # ---- power-of-two scaling ----
                keep = {k: full[k].clone() for k in ("rc", "toa", "amp", "tsc", "ci", "soft")}
                rx.mul_(2.0)
                dbl = _run(trx, rx, typ, tsc, mt, bound, soft)
                same_rc = keep["rc"] == dbl["rc"]
                flips = int((~same_rc).sum())
                assert flips <= N // 10000, (kind, "threshold flips under x2 scaling", flips)
                m = same_rc
                assert beq(keep["toa"][m], dbl["toa"][m]), kind
                assert beq(keep["tsc"][m], dbl["tsc"][m]), kind
                assert beq(keep["amp"][m] * 2.0, dbl["amp"][m]), kind
                assert beq(keep["soft"][m], dbl["soft"][m]), kind
                assert beq(keep["ci"][m], dbl["ci"][m]), kind
                print(kind, "x2 scaling: threshold flips", flips, "of", N)
                del keep, dbl
            del rx, full
    finally:
        trx.detect_config(40, 3)


def test_fullsize_vitac(trx, checker):
    """cfg4: the MLSE equaliser on 2^20-burst batches (10 M = ten of them; three are run here): subset parity with the
    CPU checker (start and all 148 decisions exact) and batch independence."""
    rng = np.random.default_rng(78)
    for bi in range(3):
        rx, typ, tsc, mt, bound = bench.make_workload(trx, "nb", N, seed=600 + bi, device=trx.device)
        buf = torch.zeros((N, 40 + 625 + 63, 2), dtype=torch.float32, device=trx.device)
        buf[:, 40:665] = rx
        del rx
        g = trx.vitac(buf, 40, tsc)
        torch.cuda.synchronize()
        idx = np.sort(rng.choice(N, SUBSET, replace=False))
        tidx = torch.from_numpy(idx).to(trx.device)
        c = checker.vitac(buf[tidx].cpu().numpy(), 40, tsc[tidx].cpu().numpy(), nthreads=8)
        assert np.array_equal(g["start"][tidx].cpu().numpy(), c["start"])
        assert np.array_equal(g["bits"][tidx].cpu().numpy(), c["bits"])
        assert np.allclose(g["corr_max"][tidx].cpu().numpy(), c["corr_max"], rtol=1e-4, atol=0)
        if bi == 0:
            h = N // 2 + 7  # odd split: the pair kernel's tail
            a = trx.vitac(buf[:h], 40, tsc[:h])
            b = trx.vitac(buf[h:], 40, tsc[h:])
            torch.cuda.synchronize()
            for k in ("start", "bits", "corr_max"):
                assert beq(g[k][:h], a[k]) and beq(g[k][h:], b[k]), k
        del buf, g


def test_fullsize_pull(trx, checker):
    """The int16 -> TRXD chain on 2^20 slots: datagrams of a random subset equal the CPU checker's, batch independent."""
    rng = np.random.default_rng(79)
    try:
        trx.detect_config(16, 1)
        rx, typ, tsc, mt, bound = bench.make_workload(trx, "nb", N, seed=700, device=trx.device)
        iq = (rx * bench.IQ_SCALE).round().clamp(-32768, 32767).to(torch.int16)
        del rx
        fn = (torch.arange(N, device=trx.device, dtype=torch.int32) // 8)
        tn = (torch.arange(N, device=trx.device) % 8).to(torch.uint8)
        po = trx.alloc_pull_results(N, 160)
        trx.pull(iq, typ, tsc, mt, fn, tn, bound, out=po)
        torch.cuda.synchronize()
        h = N // 2 + 5
        pa = trx.alloc_pull_results(h, 160)
        trx.pull(iq[:h], typ[:h], tsc[:h], mt[:h], fn[:h], tn[:h], bound, out=pa)
        torch.cuda.synchronize()
        for k in ("rc", "energy", "pkt_len"):
            assert beq(po[k][:h], pa[k]), k
        live = pa["pkt_len"] > 0
        assert torch.equal(po["pkt"][:h][live], pa["pkt"][live])
        idx = np.sort(rng.choice(N, SUBSET, replace=False))
        tidx = torch.from_numpy(idx).to(trx.device)
        g = {k: po[k][tidx].cpu().numpy() for k in ("rc", "energy", "pkt", "pkt_len", "flags")}
        g["pkt_len"] = g["pkt_len"].view(np.uint16)
        c = checker.pull(iq[tidx].cpu().numpy(), typ[tidx].cpu().numpy(), tsc[tidx].cpu().numpy(),
                         mt[tidx].cpu().numpy().astype(np.uint16), fn[tidx].cpu().numpy().astype(np.uint32), tn[tidx].cpu().numpy(),
                         version=1, rssi_offset=0.0, pkt_stride=160, nthreads=8)
        rep = parity.compare_pkts(g, c, None, "pull full-size subset", version=1)
        print("pull full-size", rep)
    finally:
        trx.detect_config(40, 3)


def test_cfg5_wideband_chain(trx, checker):
    """BASELINE configs[4] at test scale: 64 ARFCNs, 52 normal bursts each.  TX mirror on the CPU checker
    (Resampler(48,65) per channel -> Synthesis(64,192), radioInterfaceMulti.cpp:323-344); RX on the GPU
    (Channelizer(64,192) -> Resampler(65,48) -> slot slicing -> detectAnyBurst + demodAnyBurst), against the same RX
    chain on the CPU checker.  The channelizer's DFT carries the 1e-4 bar (FFTW is unpinned in the reference), so
    the chain is compared at decision level: rc and TSC equal, TOA within one interpolation quantum (1/256 symbol,
    the reference's own SSE/scalar noise floor), soft bits within 1e-3 of the burst's scale."""
    import osmo_trx_b200
    import synth
    M, BL, K = 64, 192, 52            # 52 bursts x 625 samples = 125 blocks of 260 samples at 4 sps per channel
    rng = np.random.default_rng(81)
    nblk = K * 625 // 260
    tsc = (np.arange(M * K) % 8).astype(np.uint8)
    w = checker.modulate_gmsk_batch(synth.nb_bits(M * K, tsc, rng), nthreads=8)
    gain = rng.uniform(0.3, 1.0, (M * K, 1, 1)).astype(np.float32)
    streams = (w * gain).reshape(M, K * 625, 2)
    def cpu_resample(p_, q_, x, n_in, n_out):
        """block by block as radioInterfaceMulti does (Resampler::rotate precomputes its paths for short outputs)"""
        hr = checker.resampler(p_, q_)
        xx = np.concatenate([np.zeros((16, 2), np.float32), x])
        out = []
        for k in range(len(x) // n_in):
            rc, y = checker.resampler_rotate(hr, xx[k * n_in:k * n_in + n_in + 16], 16, n_out)
            assert rc == n_out
            out.append(y)
        return np.concatenate(out)

    # ---- TX mirror on the checker ----
    chan = np.zeros((M, nblk * BL, 2), np.float32)
    for c in range(M):
        chan[c] = cpu_resample(48, 65, streams[c], 260, BL)
    sy = checker.synthesis(M, BL)
    wide = np.concatenate([checker.synthesis_rotate(sy, np.ascontiguousarray(chan[:, k * BL:(k + 1) * BL]), M, BL)[1]
                           for k in range(nblk)])
    wide += rng.standard_normal(wide.shape).astype(np.float32) * 1e-3
    # ---- RX on the checker ----
    cc = checker.channelizer(M, BL)
    rxc = np.concatenate([checker.channelizer_rotate(cc, wide[k * M * BL:(k + 1) * M * BL], M, BL)[1] for k in range(nblk)], axis=1)
    back_c = np.zeros((M, K * 625, 2), np.float32)
    for c in range(M):
        back_c[c] = cpu_resample(65, 48, rxc[c], BL, 260)
    # ---- RX on the GPU ----
    ch = osmo_trx_b200.Channelizer(trx, M, BL)
    rs = osmo_trx_b200.Resampler(trx, 65, 48)
    rxg = ch.rotate(torch.from_numpy(wide).cuda())                      # [M][nblk*BL][2]
    assert np.abs(rxg.cpu().numpy() - rxc).max() <= 1e-4 * np.abs(rxc).max()
    # every (channel, block) is one stream of the batched rotate: 16 samples of history, then the 192 new ones
    pad = torch.zeros((M, 16 + nblk * BL, 2), dtype=torch.float32, device=trx.device)
    pad[:, 16:] = rxg
    xin = pad.unfold(1, 16 + BL, BL).permute(0, 1, 3, 2).reshape(M * nblk, 16 + BL, 2).contiguous()
    back_g = rs.rotate(xin, 260).reshape(M, nblk * 260, 2)              # [M][K*625][2]
    torch.cuda.synchronize()
    # ---- the filterbank chain delays the streams: find the whole-sample lag on channel 0 and slice both alike ----
    ref0 = streams[0, :, 0] + 1j * streams[0, :, 1]
    got0 = back_c[0, :, 0] + 1j * back_c[0, :, 1]
    lags = np.arange(0, 200)
    corr = [np.abs(np.vdot(ref0[:20000], got0[l:l + 20000])) for l in lags]
    lag = int(lags[int(np.argmax(corr))])
    assert max(corr) > 0.5 * np.vdot(ref0[:20000], ref0[:20000]).real, "chain output does not resemble its input"
    nb = K - 1                                                          # the last slot runs off the end by `lag`

    def slots(a):
        return np.ascontiguousarray(a[:, lag:lag + nb * 625].reshape(M * nb, 625, 2))
    sg, sc = slots(back_g.cpu().numpy()), slots(back_c)
    tsc_s = np.ascontiguousarray(tsc.reshape(M, K)[:, :nb].reshape(-1))
    try:
        trx.detect_config(16, 1)
        n = M * nb
        r = trx.detect_demod(torch.from_numpy(sg).cuda(), torch.full((n,), TSC, dtype=torch.uint8, device=trx.device),
                             torch.from_numpy(tsc_s).cuda(), torch.full((n,), 4, dtype=torch.int16, device=trx.device), 4,
                             n_gmsk_soft=148)
        torch.cuda.synchronize()
    finally:
        trx.detect_config(40, 3)
    g = {k: v.cpu().numpy() for k, v in r.items()}
    c = checker.detect_demod(sc, TSC, tsc_s, 4, nthreads=8)
    assert (c["rc"] == TSC).mean() > 0.99, "the CPU chain itself must detect its bursts"
    assert np.array_equal(g["rc"], c["rc"]) and np.array_equal(g["tsc"][c["rc"] > 0], c["tsc"][c["rc"] > 0])
    det = c["rc"] > 0
    dtoa = np.abs(g["toa"][det].astype(np.float64) - c["toa"][det])
    assert dtoa.max() <= 1.0 / 256 + 1e-9, dtoa.max()
    same = det.copy()
    same[det] = dtoa == 0
    scale = np.abs(c["soft"][same][:, :148]).max(axis=1, keepdims=True)
    rel = np.abs(g["soft"][same][:, :148].astype(np.float64) - c["soft"][same][:, :148]) / scale
    assert rel.max() <= 1e-3, rel.max()
    print("cfg5 chain: lag", lag, "slots", n, "detected", int(det.sum()), "toa quantum flips", int((dtoa > 0).sum()),
          "soft rel max", float(rel.max()))
