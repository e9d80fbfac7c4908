"""Seeded synthetic burst generators for the parity tests (SURVEY.md §8(d) recipes).

Bits are laid out as the reference's own generators do (sigProcLib.cpp:768-806 normal burst,
:811-841 access burst, :868-907 EDGE burst) but with a numpy PRNG instead of libc rand().
Waveforms come from a modulator passed in by the caller (oracle, reference or the CUDA path),
then get an integer+fractional time shift, a complex gain and AWGN.
"""
import numpy as np

TSC_STR = [
    "00100101110000100010010111", "00101101110111100010110111", "01000011101110100100001110",
    "01000111101101000100011110", "00011010111001000001101011", "01001110101100000100111010",
    "10100111110110001010011111", "11101111000100101110111100",
]
EDGE_TSC_STR = [
    "111111001111111001111001001001111111111111001111111111001111111001111001001001",
    "111111001111001001111001001001111001001001001111111111001111001001111001001001",
    "111001111111111111001001001111001001001111001111111001111111111111001001001111",
    "111001111111111001001001001111001001111001111111111001111111111001001001001111",
    "111111111001001111001111001001001111111001111111111111111001001111001111001001",
    "111001111111001001001111001111001001111111111111111001111111001001001111001111",
    "001111001111111001001001001001111001001111111111001111001111111001001001001001",
    "001001001111001001001001111111111001111111001111001001001111001001001001111111",
]
RACH_SYNC_STR = [
    "01001011011111111001100110101010001111000",
    "01010100111110001000011000101111001001101",
    "11101111001001110101011000001101101110111",
]
# GSM::gDummyBurst (GSMCommon.cpp:56-58): the 148-bit dummy burst of 3GPP TS 45.002 5.2.6
DUMMY_STR = ("0001111101101110110000010100100111000001001000100000001111100011100010111000101110001010111010010100011001100111001111010011111"
             "000100101111101010000")
RACH_HEAD = "00111010"  # first 8 bits of GSM::gRACHBurst (GSMCommon.cpp:68)


def _bits(s):
    return np.frombuffer(s.encode(), np.uint8) - ord("0")


def nb_bits(n, tsc, rng):
    """[n,148] normal-burst bits: 3 tail, 57 data, steal, 26 TSC, steal, 57 data, 3 tail."""
    tsc = np.broadcast_to(np.asarray(tsc), (n,))
    b = np.zeros((n, 148), np.uint8)
    b[:, 3:60] = rng.integers(0, 2, (n, 57))
    b[:, 88:145] = rng.integers(0, 2, (n, 57))
    for t in range(8):
        b[tsc == t, 61:87] = _bits(TSC_STR[t])
    return b


def ab_bits(n, delay, rng, seq=0):
    """[n,88+delay] access-burst bits: `delay` zeros, 8 head, 41 sync, 36 data, 3 tail."""
    b = np.zeros((n, 88 + delay), np.uint8)
    b[:, delay:delay + 8] = _bits(RACH_HEAD)
    b[:, delay + 8:delay + 49] = _bits(RACH_SYNC_STR[seq])
    b[:, delay + 49:delay + 85] = rng.integers(0, 2, (n, 36))
    return b


PSK8_BITS = np.array([[(i >> 0) & 1, (i >> 1) & 1, (i >> 2) & 1] for i in range(8)], np.uint8)


def edge_bits(n, tsc, rng):
    """[n,444] EDGE bits: 3 tail symbols (index 7), 58 data, 26 TSC symbols, 58 data, 3 tail."""
    tsc = np.broadcast_to(np.asarray(tsc), (n,))
    sym = np.full((n, 148), 7, np.int64)
    sym[:, 3:61] = rng.integers(0, 8, (n, 58))
    sym[:, 87:145] = rng.integers(0, 8, (n, 58))
    b = PSK8_BITS[sym].reshape(n, 444).copy()
    for t in range(8):
        b[tsc == t, 61 * 3:87 * 3] = _bits(EDGE_TSC_STR[t])
    return b


def frac_shift(wave, shift):
    """Band-limited shift of complex [n,L] by `shift` samples (positive = later), zero padded (FFT method)."""
    n, L = wave.shape
    P = 1024
    x = np.zeros((n, P), np.complex128)
    x[:, 128:128 + L] = wave
    f = np.fft.fftfreq(P)
    X = np.fft.fft(x, axis=1) * np.exp(-2j * np.pi * f[None, :] * np.asarray(shift)[:, None])
    return np.fft.ifft(X, axis=1)[:, 128:128 + L]


def impair(wave, rng, snr_db=30.0, amp_range=(0.1, 1.0), full_scale=1.0, shift_lo=-26.0, shift_hi=-9.0,
           noise_only_frac=0.0):
    """wave: float32 [n,L,2] clean TX bursts (L = 625 at 4 sps, 156 / 157 at 1 sps) -> float32 [n,L,2] received bursts.

    shift ~ U[shift_lo, shift_hi) samples (the modulator output peaks ~18.5 samples late),
    gain A e^{j phi}, A ~ U[amp_range]*full_scale, AWGN at `snr_db` relative to the burst's mean power.
    """
    n, L = wave.shape[0], wave.shape[1]
    w = wave[..., 0].astype(np.float64) + 1j * wave[..., 1].astype(np.float64)
    shift = rng.uniform(shift_lo, shift_hi, n)
    y = frac_shift(w, shift)
    A = rng.uniform(amp_range[0], amp_range[1], n) * full_scale
    phi = rng.uniform(0, 2 * np.pi, n)
    y = y * (A * np.exp(1j * phi))[:, None]
    snr = np.broadcast_to(np.asarray(snr_db, np.float64), (n,))
    sigma = A * 10.0 ** (-snr / 20.0) / np.sqrt(2.0)
    noise = (rng.standard_normal((n, L)) + 1j * rng.standard_normal((n, L))) * sigma[:, None]
    if noise_only_frac > 0:
        kill = rng.uniform(0, 1, n) < noise_only_frac
        y[kill] = 0
    y = y + noise
    out = np.empty((n, L, 2), np.float32)
    out[..., 0] = y.real
    out[..., 1] = y.imag
    return out, shift


def multipath(wave, rng, taps_sigma=(0.3, 0.2, 0.1)):
    """Cfg-4 channel: h0 = 1, h1..3 ~ CN(0, sigma^2) at 4,8,12 samples."""
    n = wave.shape[0]
    w = wave[..., 0].astype(np.float64) + 1j * wave[..., 1].astype(np.float64)
    y = w.copy()
    for k, s in enumerate(taps_sigma, start=1):
        h = (rng.standard_normal(n) + 1j * rng.standard_normal(n)) * s / np.sqrt(2.0)
        y[:, 4 * k:] += h[:, None] * w[:, :-4 * k]
    out = np.empty((n, 625, 2), np.float32)
    out[..., 0] = y.real
    out[..., 1] = y.imag
    return out


SCH_SYNC_STR = "1011100101100010000001000000111100101101010001010111011000011011"


def sch_bits(n, rng):
    """Synchronisation bursts (3GPP TS 45.002 5.2.5): 3 tail, 39 data, 64-bit extended training sequence, 39 data, 3 tail."""
    bits = np.zeros((n, 148), np.uint8)
    bits[:, 3:42] = rng.integers(0, 2, (n, 39))
    bits[:, 42:106] = [int(c) for c in SCH_SYNC_STR]
    bits[:, 106:145] = rng.integers(0, 2, (n, 39))
    return bits


def sps1_bursts(mod, n, rng, blen=157, edge_every=0, snr_db=20.0, noise_only_frac=0.05):
    """Received bursts at ONE sample per symbol (the reference's rx_sps = 1 path): `mod` is a checker whose
    modulate_burst(bits, guard, sps=1) is the reference's 1-sps modulator (modulateBurstBasic, sigProcLib.cpp:938-968).
    Normal bursts with TSC b % 8; every `edge_every`-th burst 8-PSK (rotateEdgeBurst at 1 sps, no pulse shaping: the
    reference itself notes that 8-PSK is nearly unrecoverable at 1 sps - the point is equal arithmetic, not sensitivity).
    Returns (rx [n, blen, 2], tsc [n], is_edge [n])."""
    tsc = (np.arange(n) % 8).astype(np.uint8)
    bits = nb_bits(n, tsc, rng)
    ebits = edge_bits(n, tsc, rng)
    tx = np.zeros((n, blen, 2), np.float32)
    is_edge = np.zeros(n, bool)
    for b in range(n):
        if edge_every and b % edge_every == edge_every - 1:
            w = mod.modulate_edge(ebits[b], sps=1, empty=True)
            is_edge[b] = True
        else:
            w = mod.modulate_burst(bits[b], guard=blen - 148, sps=1)
        tx[b, :min(blen, len(w))] = w[:blen]
    rx, _ = impair(tx, rng, snr_db=snr_db, shift_lo=-2.0, shift_hi=1.5, noise_only_frac=noise_only_frac)
    return rx, tsc, is_edge
