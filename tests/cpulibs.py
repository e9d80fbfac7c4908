"""ctypes loaders for the two CPU checkers (TEST INFRASTRUCTURE):

* ``Oracle``  - oracle/_build/liboracle_trx.so, the from-scratch C restatement (oracle/oracle_*.c)
* ``Ref``     - oracle/_ref/libref_osmotrx.so, the unmodified reference compiled from /root/reference
                (present when built in the dev container; travels to the GPU box as a prebuilt .so)

Both expose the same batched array layout as the product C ABI (include/trxb200.h).
"""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "liboracle_trx.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_osmotrx.so")

OFF, TSC, EXT_RACH, RACH, SCH, EDGE, IDLE = range(7)


def build_oracle():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True,
                   stdout=subprocess.DEVNULL)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class _Base:
    """Shared batch API; subclasses set self.L and self.pfx ('orc_' / 'ref_') and self.has_flags."""

    def get_table(self, name, idx=0):
        buf = np.zeros(4096, np.float32)
        n = self.L[self.pfx + "get_table"](name.encode(), C.c_int(idx), _p(buf), C.c_int(4096))
        if n < 0:
            raise KeyError(name)
        return buf[:n].copy()

    def modulate_burst(self, bits, guard=0, sps=4, empty=False):
        bits = np.ascontiguousarray(bits, np.uint8)
        out = np.zeros((700, 2), np.float32)
        n = self.L[self.pfx + "modulate_burst"](_p(bits), C.c_int(len(bits)), C.c_int(guard), C.c_int(sps),
                                                  C.c_int(int(empty)), _p(out), C.c_int(700))
        assert n >= 0
        return out[:n]

    def modulate_edge(self, bits, sps=4, empty=False):
        bits = np.ascontiguousarray(bits, np.uint8)
        out = np.zeros((700, 2), np.float32)
        n = self.L[self.pfx + "modulate_edge"](_p(bits), C.c_int(len(bits)), C.c_int(sps), C.c_int(int(empty)),
                                                 _p(out), C.c_int(700))
        assert n >= 0
        return out[:n]

    def modulate_gmsk_batch(self, bits, nthreads=1):
        bits = np.ascontiguousarray(bits, np.uint8)
        n, nb = bits.shape
        out = np.zeros((n, 625, 2), np.float32)
        self.L[self.pfx + "modulate_gmsk_batch"](_p(bits), C.c_int(nb), C.c_int(n), _p(out), C.c_int(nthreads))
        return out

    def modulate_edge_batch(self, bits, nthreads=1):
        bits = np.ascontiguousarray(bits, np.uint8)
        n, nb = bits.shape
        out = np.zeros((n, 625, 2), np.float32)
        self.L[self.pfx + "modulate_edge_batch"](_p(bits), C.c_int(nb), C.c_int(n), _p(out), C.c_int(nthreads))
        return out

    def _dd_args(self, bursts, type_, tsc, max_toa):
        bursts = _f32(bursts)
        n, stride = bursts.shape[0], bursts.shape[1]
        type_ = np.ascontiguousarray(np.broadcast_to(type_, (n,)), np.uint8)
        tsc = np.ascontiguousarray(np.broadcast_to(tsc, (n,)), np.uint8)
        max_toa = np.ascontiguousarray(np.broadcast_to(max_toa, (n,)), np.uint16)
        return bursts, n, stride, type_, tsc, max_toa

    def detect(self, bursts, type_, tsc, max_toa, thresh=4.0, sps=4, nthreads=1, blen=625):
        bursts, n, stride, type_, tsc, max_toa = self._dd_args(bursts, type_, tsc, max_toa)
        r = dict(rc=np.zeros(n, np.int32), amp=np.zeros((n, 2), np.float32), toa=np.zeros(n, np.float32),
                 tsc=np.zeros(n, np.uint8), ci=np.zeros(n, np.float32), flags=np.zeros(n, np.uint8))
        args = [_p(bursts), C.c_int(stride), C.c_int(blen), C.c_int(n), _p(type_), _p(tsc), _p(max_toa),
                C.c_float(thresh), C.c_int(sps), _p(r["rc"]), _p(r["amp"]), _p(r["toa"]), _p(r["tsc"]), _p(r["ci"])]
        if self.has_flags:
            args.append(_p(r["flags"]))
        args.append(C.c_int(nthreads))
        self.L[self.pfx + "detect_batch"](*args)
        return r

    def demod(self, bursts, rc, amp, toa, ci, sps=4, soft_stride=444, nthreads=1, blen=625):
        bursts = _f32(bursts)
        n, stride = bursts.shape[0], bursts.shape[1]
        rc = np.ascontiguousarray(rc, np.int32)
        amp = _f32(amp)
        toa = _f32(toa)
        ci = _f32(ci).copy()
        soft = np.zeros((n, soft_stride), np.float32)
        nsoft = np.zeros(n, np.int32)
        self.L[self.pfx + "demod_batch"](_p(bursts), C.c_int(stride), C.c_int(blen), C.c_int(n), _p(rc), _p(amp),
                                           _p(toa), _p(ci), C.c_int(sps), _p(soft), C.c_int(soft_stride), _p(nsoft),
                                           C.c_int(nthreads))
        return dict(soft=soft, nsoft=nsoft, ci=ci)

    def detect_demod(self, bursts, type_, tsc, max_toa, thresh=4.0, sps=4, soft_stride=444, nthreads=1, blen=625, out=None):
        bursts, n, stride, type_, tsc, max_toa = self._dd_args(bursts, type_, tsc, max_toa)
        r = out if out is not None else dict(
            rc=np.zeros(n, np.int32), amp=np.zeros((n, 2), np.float32), toa=np.zeros(n, np.float32),
            tsc=np.zeros(n, np.uint8), ci=np.zeros(n, np.float32), flags=np.zeros(n, np.uint8),
            soft=np.zeros((n, soft_stride), np.float32), nsoft=np.zeros(n, np.int32))
        args = [_p(bursts), C.c_int(stride), C.c_int(blen), C.c_int(n), _p(type_), _p(tsc), _p(max_toa),
                C.c_float(thresh), C.c_int(sps), _p(r["rc"]), _p(r["amp"]), _p(r["toa"]), _p(r["tsc"]), _p(r["ci"])]
        if self.has_flags:
            args.append(_p(r["flags"]))
        args += [_p(r["soft"]), C.c_int(soft_stride), _p(r["nsoft"]), C.c_int(nthreads)]
        self.L[self.pfx + "detect_demod_batch"](*args)
        return r

    def _conv(self, fn, x, h, start, length, x_off=0):
        """x: complex array [x_total,2] with the addressed vector beginning x_off samples in (head-room)."""
        x = _f32(x)
        h = _f32(h)
        y = np.zeros((length, 2), np.float32)
        xp = C.c_void_p(x.ctypes.data + 8 * x_off)
        rc = self.L[self.pfx + fn](xp, C.c_int(x.shape[0] - x_off), _p(h), C.c_int(h.shape[0]), _p(y),
                                     C.c_int(length), C.c_int(start), C.c_int(length))
        return rc, y

    def convolve_real(self, x, h, start, length, x_off=0):
        return self._conv("convolve_real", x, h, start, length, x_off)

    def convolve_complex(self, x, h, start, length, x_off=0):
        return self._conv("convolve_complex", x, h, start, length, x_off)

    def base_convolve_real(self, x, h, start, length, x_off=0):
        return self._conv("base_convolve_real", x, h, start, length, x_off)

    def base_convolve_complex(self, x, h, start, length, x_off=0):
        return self._conv("base_convolve_complex", x, h, start, length, x_off)

    def energy_detect(self, burst, window):
        burst = _f32(burst)
        f = self.L[self.pfx + "energy_detect"]
        f.restype = C.c_float
        return f(_p(burst), C.c_int(burst.shape[0]), C.c_uint(window))

    def delay_vector(self, x, delay):
        x = _f32(x)
        out = np.zeros_like(x)
        self.L[self.pfx + "delay_vector"](_p(x), C.c_int(x.shape[0]), C.c_float(delay), _p(out))
        return out

    def downsample_burst(self, x):
        x = _f32(x)
        out = np.zeros((156, 2), np.float32)
        self.L[self.pfx + "downsample_burst"](_p(x), C.c_int(x.shape[0]), _p(out))
        return out

    def vector_slicer(self, src):
        src = _f32(src)
        dst = np.zeros_like(src)
        self.L[self.pfx + "vector_slicer"](_p(dst), _p(src), C.c_size_t(src.size))
        return dst

    def convert_float_short(self, x, scale):
        x = _f32(x)
        out = np.zeros(x.size, np.int16)
        self.L[self.pfx + "convert_float_short"](_p(out), _p(x), C.c_float(scale), C.c_int(x.size))
        return out

    def convert_float_short_mode(self, x, scale, mode):
        """mode 1: the x86 dispatcher (SSE body, truncating tail of len % 8); 2: base_convert_float_short"""
        x = _f32(x)
        out = np.zeros(x.size, np.int16)
        name = {1: "convert_float_short_x86", 2: "base_convert_float_short"}[mode]
        self.L[self.pfx + name](_p(out), _p(x), C.c_float(scale), C.c_int(x.size))
        return out

    def convert_short_float(self, x):
        x = np.ascontiguousarray(x, np.int16)
        out = np.zeros(x.size, np.float32)
        self.L[self.pfx + "convert_short_float"](_p(out), _p(x), C.c_int(x.size))
        return out

    # --- receive chain around the hot path: int16 slots -> TRXD uplink datagrams ---
    def pull(self, iq, type_, tsc, max_toa, fn, tn, thresh=4.0, full_scale=32767.0, rssi_offset=0.0, version=1,
             pkt_stride=None, nthreads=1, capture=True, out=None):
        """iq: int16 [n][stride][2].  Returns rc, energy, pkt [n][pkt_stride] u8, pkt_len, amp, toa, ci, tsc (+flags)."""
        iq = np.ascontiguousarray(iq, np.int16)
        n, stride = iq.shape[0], iq.shape[1]
        type_ = np.ascontiguousarray(np.broadcast_to(type_, (n,)), np.uint8)
        tsc = np.ascontiguousarray(np.broadcast_to(tsc, (n,)), np.uint8)
        max_toa = np.ascontiguousarray(np.broadcast_to(max_toa, (n,)), np.uint16)
        fn = np.ascontiguousarray(np.broadcast_to(fn, (n,)), np.uint32)
        tn = np.ascontiguousarray(np.broadcast_to(tn, (n,)), np.uint8)
        if pkt_stride is None:
            pkt_stride = 11 + 444 + 2
        r = out if out is not None else dict(
            rc=np.zeros(n, np.int32), energy=np.zeros(n, np.float32), pkt=np.zeros((n, pkt_stride), np.uint8),
            pkt_len=np.zeros(n, np.uint16), flags=np.zeros(n, np.uint8), amp=np.zeros((n, 2), np.float32),
            toa=np.zeros(n, np.float32), ci=np.zeros(n, np.float32), tsc=np.zeros(n, np.uint8))
        args = [_p(iq), C.c_int(stride), C.c_int(n), _p(type_), _p(tsc), _p(max_toa), _p(fn), _p(tn), C.c_float(thresh),
                C.c_double(full_scale), C.c_double(rssi_offset), C.c_int(version), _p(r["rc"]), _p(r["energy"]),
                _p(r["pkt"]), C.c_int(pkt_stride), _p(r["pkt_len"])]
        if self.has_flags:
            args.append(_p(r["flags"]))
        args += [_p(r["amp"]), _p(r["toa"]), _p(r["ci"]), _p(r["tsc"])]
        if not self.has_flags:
            args.append(C.c_int(int(capture)))
        args.append(C.c_int(nthreads))
        self.L[self.pfx + "pull_batch"](*args)
        return r

    # --- resampler / filterbanks ---
    def resampler(self, p, q, filt_len=16, bw=1.0):
        f = self.L[self.pfx + "resampler_create"]
        f.restype = C.c_void_p
        h = f(C.c_int(p), C.c_int(q), C.c_int(filt_len), C.c_float(bw))
        assert h
        return C.c_void_p(h)

    def resampler_rotate(self, h, in_with_hist, hist, out_len):
        x = _f32(in_with_hist)
        out = np.zeros((out_len, 2), np.float32)
        rc = self.L[self.pfx + "resampler_rotate"](h, _p(x), C.c_int(hist), C.c_int(x.shape[0] - hist), _p(out),
                                                     C.c_int(out_len))
        return rc, out

    def channelizer(self, m, block_len, h_len=16):
        f = self.L[self.pfx + "channelizer_create"]
        f.restype = C.c_void_p
        return C.c_void_p(f(C.c_int(m), C.c_int(block_len), C.c_int(h_len)))

    def synthesis(self, m, block_len, h_len=16):
        f = self.L[self.pfx + "synthesis_create"]
        f.restype = C.c_void_p
        return C.c_void_p(f(C.c_int(m), C.c_int(block_len), C.c_int(h_len)))

    def wideband_rx(self, wide, nblk, m=64, block_len=192, p=65, q=48):
        """Channelizer(m, block_len) + Resampler(p, q) per channel over nblk blocks, block by block as radioInterfaceMulti
        does it (compiled reference only).  wide: [nblk * block_len * m, 2] -> [m, nblk * block_len / q * p, 2]"""
        wide = _f32(wide)
        out = np.zeros((m, nblk * (block_len // q * p), 2), np.float32)
        rc = self.L[self.pfx + "wideband_rx"](_p(wide), C.c_int(nblk), C.c_int(m), C.c_int(block_len), C.c_int(p), C.c_int(q), _p(out))
        assert rc == 0
        return out

    def channelizer_rotate(self, h, x, m, block_len):
        x = _f32(x)
        out = np.zeros((m, block_len, 2), np.float32)
        rc = self.L[self.pfx + "channelizer_rotate"](h, _p(x), C.c_int(m), C.c_int(block_len), _p(out))
        return rc, out

    def synthesis_rotate(self, h, x, m, block_len):
        x = _f32(x)
        out = np.zeros((m * block_len, 2), np.float32)
        rc = self.L[self.pfx + "synthesis_rotate"](h, _p(x), C.c_int(m), C.c_int(block_len), _p(out))
        return rc, out

    # --- detectSCHBurst, SCH_DETECT_FULL ---
    def detect_sch(self, bursts, thresh=4.0):
        x = _f32(bursts)
        n, stride = x.shape[0], x.shape[1]
        r = dict(rc=np.zeros(n, np.int32), amp=np.zeros((n, 2), np.float32), toa=np.zeros(n, np.float32),
                 ci=np.zeros(n, np.float32), flags=np.zeros(n, np.uint8))
        args = [_p(x), C.c_int(stride), C.c_int(625), C.c_int(n), C.c_float(thresh), _p(r["rc"]), _p(r["amp"]),
                _p(r["toa"]), _p(r["ci"])]
        if self.has_flags:
            args.append(_p(r["flags"]))
        self.L[self.pfx + "detect_sch_batch"](*args)
        return r

    def detect_sch_buffer(self, bufs, in_len=60000, thresh=4.0):
        """detectSCHBurst(SCH_DETECT_BUFFER): bufs [n, >= in_len, 2] captures (12 frames = 60000 samples in the reference)"""
        x = _f32(bufs)
        n, stride = x.shape[0], x.shape[1]
        r = dict(rc=np.zeros(n, np.int32), amp=np.zeros((n, 2), np.float32), toa=np.zeros(n, np.float32),
                 ci=np.zeros(n, np.float32), flags=np.zeros(n, np.uint8))
        args = [_p(x), C.c_int(stride), C.c_int(in_len), C.c_int(n), C.c_float(thresh), _p(r["rc"]), _p(r["amp"]),
                _p(r["toa"]), _p(r["ci"])]
        if self.has_flags:
            args.append(_p(r["flags"]))
        self.L[self.pfx + "detect_sch_buffer_batch"](*args)
        return r

    # --- burst-type scheduler (Transceiver::expectedCorrType) ---
    def expected_corr_type(self, chan_type8, handover8, ext_rach, egprs, fn, tn):
        """chan_type8: ChannelCombination per timeslot of one channel (u8[8]); handover8: sub-slot bit mask per timeslot."""
        ct = np.ascontiguousarray(chan_type8, np.uint8)
        ho = np.ascontiguousarray(handover8, np.uint8)
        fn = np.ascontiguousarray(fn, np.uint32)
        tn = np.ascontiguousarray(tn, np.uint8)
        out = np.zeros(len(fn), np.uint8)
        self.L[self.pfx + "expected_corr_type"](_p(ct), _p(ho), C.c_int(int(ext_rach)), C.c_int(int(egprs)), _p(fn), _p(tn),
                                                C.c_int(len(fn)), _p(out))
        return out

    # --- vitac ---
    def vitac_table(self, which, idx=0):
        buf = np.zeros((64, 2), np.float32)
        n = self.L[self.pfx + "get_vitac_table"](C.c_int(which), C.c_int(idx), _p(buf))
        return buf[:n].copy()

    def vitac(self, bufs, offset, tsc, is_ab=False, max_delay=0, clamp=(-39, 39), nthreads=1):
        bufs = _f32(bufs)
        n, stride = bufs.shape[0], bufs.shape[1]
        tsc = np.ascontiguousarray(np.broadcast_to(tsc, (n,)), np.uint8)
        nb = 88 if int(is_ab) == 1 else 148  # is_ab: 0 normal, 1 access, 2 SCH burst
        r = dict(bits=np.zeros((n, nb), np.int8), start=np.zeros(n, np.int32), corr_max=np.zeros(n, np.float32),
                 cir=np.zeros((n, 20, 2), np.float32))
        self.L[self.pfx + "vitac_batch"](_p(bufs), C.c_int(stride), C.c_int(offset), C.c_int(n), C.c_int(int(is_ab)),
                                           _p(tsc), C.c_int(max_delay), C.c_int(clamp[0]), C.c_int(clamp[1]),
                                           _p(r["bits"]), _p(r["start"]), _p(r["corr_max"]), _p(r["cir"]),
                                           C.c_int(nthreads))
        return r

    def vitac_sch_buffer(self, bufs, offset, length, want_bits=True):
        """get_sch_buffer_chan_imp_resp over `length` samples from `offset` of every row + detect_burst_nb at the position found"""
        bufs = _f32(bufs)
        n, stride = bufs.shape[0], bufs.shape[1]
        r = dict(bits=np.zeros((n, 148), np.int8), start=np.zeros(n, np.int32), corr_max=np.zeros(n, np.float32),
                 cir=np.zeros((n, 20, 2), np.float32))
        rc = self.L[self.pfx + "vitac_sch_buffer_batch"](_p(bufs), C.c_int(stride), C.c_int(offset), C.c_int(length), C.c_int(n),
                                                         _p(r["bits"]) if want_bits else None, _p(r["start"]), _p(r["corr_max"]), _p(r["cir"]))
        assert rc == n, rc
        return r

    def vitac_detect(self, bufs, offset, cir, start, is_ab=False, ss=3):
        """detect_burst_nb / detect_burst_ab with a given channel estimate (cir [n,20,2]) and start per burst
        (ss: the Viterbi detector's start state, 3 in the four-argument forms)."""
        bufs = _f32(bufs)
        cir = _f32(cir)
        n, stride = bufs.shape[0], bufs.shape[1]
        start = np.ascontiguousarray(start, np.int32)
        nb = 88 if is_ab else 148
        bits = np.zeros((n, nb), np.int8)
        self.L[self.pfx + "vitac_detect_ss_batch"](_p(bufs), C.c_int(stride), C.c_int(offset), C.c_int(n), C.c_int(int(is_ab)),
                                                     _p(cir), _p(start), C.c_int(ss), _p(bits))
        return bits

    def viterbi(self, x, rhh, start_state=3):
        x = _f32(x)
        rhh = _f32(rhh)
        out = np.zeros(x.shape[0], np.float32)
        self.L[self.pfx + "viterbi"](_p(x), C.c_int(x.shape[0]), _p(rhh), C.c_int(start_state), _p(out))
        return out


class Oracle(_Base):
    pfx = "orc_"
    has_flags = True

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        self.L = C.CDLL(ORACLE_SO)
        self.L.orc_setup.restype = C.c_void_p
        self.L.orc_setup()


class Ref(_Base):
    pfx = "ref_"
    has_flags = False

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def __init__(self):
        self.L = C.CDLL(REF_SO)
        assert self.L.ref_setup() == 0
