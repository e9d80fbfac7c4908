"""osmo_trx_b200 — B200-native batched burst DSP behind the sigProcLib / Resampler / Channelizer /
grgsm_vitac interface of osmo-trx's Transceiver52M.

This package is a thin ctypes binding over ``libtrxb200.so`` (hand-written sm_100a CUDA kernels behind
the C ABI in ``include/trxb200.h``).  torch is used only for device memory, streams and
``torch.distributed``.  There is no CPU fallback: constructing :class:`Trx` raises when the shared
library is missing or no Blackwell GPU is visible.
"""
import ctypes as C
import os

import torch

from . import buildlib as _build

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtrxb200.so")

OFF, TSC, EXT_RACH, RACH, SCH, EDGE, IDLE = range(7)
SIGERR_NONE, SIGERR_BOUNDS, SIGERR_CLIP, SIGERR_UNSUPPORTED, SIGERR_INTERNAL = range(5)
FLAG_THRESH_EDGE, FLAG_BISECT_TIE, FLAG_CLIP, FLAG_PKT_TRUNC = 1, 2, 4, 8
BURST_LEN = 625
BURST_THRESH = 4.0  # sigProcLib.h:54

_lib = None


class SchedCfg(C.Structure):
    """trxb200_sched_cfg (include/trxb200.h)"""
    _fields_ = [("n_chan", C.c_int), ("chan_type", C.c_void_p), ("handover", C.c_void_p), ("ext_rach", C.c_int),
                ("egprs", C.c_int), ("max_toa_nb", C.c_int), ("max_toa_ab", C.c_int)]


class PullArgs(C.Structure):
    """trxb200_pull_args (include/trxb200.h)"""
    _fields_ = [("iq", C.c_void_p), ("stride", C.c_int), ("n", C.c_int), ("type", C.c_void_p), ("tsc", C.c_void_p),
                ("max_toa", C.c_void_p), ("fn", C.c_void_p), ("tn", C.c_void_p), ("max_toa_bound", C.c_int),
                ("thresh", C.c_float), ("rx_full_scale", C.c_double), ("rssi_offset", C.c_double),
                ("trxd_version", C.c_int), ("rc", C.c_void_p), ("energy", C.c_void_p), ("pkt", C.c_void_p),
                ("pkt_stride", C.c_int), ("pkt_len", C.c_void_p), ("flags", C.c_void_p), ("amp", C.c_void_p),
                ("toa", C.c_void_p), ("ci", C.c_void_p), ("tsc_out", C.c_void_p)]


class TrxError(RuntimeError):
    pass


def load_library(path=LIB_PATH):
    """dlopen libtrxb200.so (building it first only if sources are newer and nvcc exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise TrxError(
            f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(path)
    lib.trxb200_last_error.restype = C.c_char_p
    lib.trxb200_get_stream.restype = C.c_void_p
    lib.trxb200_launch_count.restype = C.c_uint64
    _lib = lib
    return lib


def _ptr(t):
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def _chk_dev(*ts):
    for t in ts:
        if t is not None and (not t.is_cuda or not t.is_contiguous()):
            raise TrxError("tensor arguments must be contiguous CUDA tensors")


class Trx:
    """One context per process/GPU (sigProcLibSetup + initvita equivalent)."""

    def __init__(self, device=None):
        self.lib = load_library()
        if not torch.cuda.is_available():
            raise TrxError("no CUDA device visible; osmo_trx_b200 has no CPU path")
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device)
        h = C.c_void_p()
        rc = self.lib.trxb200_init(C.c_int(self.device.index), C.byref(h))
        if rc != 0:
            raise TrxError(f"trxb200_init failed with {rc}")
        self.h = h
        torch.cuda.set_device(self.device)

    def close(self):
        if getattr(self, "h", None):
            self.lib.trxb200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- plumbing --
    def _check(self, rc, what):
        if rc != 0:
            msg = self.lib.trxb200_last_error(self.h)
            raise TrxError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    def use_current_stream(self):
        s = torch.cuda.current_stream(self.device).cuda_stream
        self.lib.trxb200_set_stream(self.h, C.c_void_p(s))

    @property
    def launch_count(self):
        return int(self.lib.trxb200_launch_count(self.h))

    def get_table(self, name, idx=0, max_floats=16384):
        import numpy as np
        buf = np.zeros(max_floats, np.float32)
        n = self.lib.trxb200_get_table(self.h, name.encode(), C.c_int(idx), buf.ctypes.data_as(C.c_void_p),
                                       C.c_int(max_floats))
        if n < 0:
            raise KeyError(name)
        return buf[:n].copy()

    def profile_begin(self):
        self._check(self.lib.trxb200_profile_begin(self.h), "profile_begin")

    def profile_end(self):
        """-> {kernel: (total_ms, launches)} for the detect/demod kernels launched since profile_begin()."""
        buf = C.create_string_buffer(4096)
        r = self.lib.trxb200_profile_end(self.h, buf, C.c_int(4096))
        if r < 0:
            self._check(r, "profile_end")
        out = {}
        for item in buf.value.decode().split(";"):
            if item:
                name, ms, cnt = item.split(":")
                out[name] = (float(ms), int(cnt))
        return out

    def detect_config(self, max_seq_len=40, max_attempts=3):
        """max_seq_len 16: only TSC/EDGE/IDLE bursts will be submitted (smaller on-chip buffers); 40: any type.
        max_attempts: detection rounds per batch (1: TSC/RACH/IDLE only, 2: + EDGE, 3: + EXT_RACH)."""
        self._check(self.lib.trxb200_detect_config(self.h, C.c_int(max_seq_len), C.c_int(max_attempts)), "detect_config")

    # -- modulators --
    def modulate_gmsk(self, bits, out=None):
        """bits: uint8 [n, nbits] on device -> float32 [n, 625, 2]"""
        _chk_dev(bits, out)
        n, nbits = bits.shape
        if out is None:
            out = torch.empty((n, BURST_LEN, 2), dtype=torch.float32, device=bits.device)
        self.use_current_stream()
        self._check(self.lib.trxb200_modulate_gmsk_batch(self.h, _ptr(bits), C.c_int(nbits), C.c_int(bits.stride(0)),
                                                         C.c_int(n), _ptr(out), C.c_int(out.stride(0) // 2)),
                    "modulate_gmsk_batch")
        return out

    def modulate_edge(self, bits, out=None):
        _chk_dev(bits, out)
        n, nbits = bits.shape
        if out is None:
            out = torch.empty((n, BURST_LEN, 2), dtype=torch.float32, device=bits.device)
        self.use_current_stream()
        self._check(self.lib.trxb200_modulate_edge_batch(self.h, _ptr(bits), C.c_int(nbits), C.c_int(bits.stride(0)),
                                                         C.c_int(n), _ptr(out), C.c_int(out.stride(0) // 2)),
                    "modulate_edge_batch")
        return out

    # -- detection / demod --
    def modulate_basic(self, bits, guard=0, sps=1, mode=0):
        """modulateBurst outside the 4-sps Laurent case (sigProcLib.cpp:558-580,672-689,938-979).
        mode 0: modulateBurstBasic (sps 1); 1: rotateBurst (emptyPulse, sps 1 or 4); 2: rotateEdgeBurst.
        bits: uint8 [n, nbits] on device -> float32 [n, sps * (symbols + guard), 2]"""
        _chk_dev(bits)
        n, nbits = bits.shape
        nsym = nbits // 3 if mode == 2 else nbits
        out = torch.empty((n, sps * (nsym + guard), 2), dtype=torch.float32, device=bits.device)
        self.use_current_stream()
        self._check(self.lib.trxb200_modulate_basic_batch(self.h, _ptr(bits), C.c_int(nbits), C.c_int(bits.stride(0)), C.c_int(n),
                                                          C.c_int(guard), C.c_int(sps), C.c_int(mode), _ptr(out),
                                                          C.c_int(out.stride(0) // 2 if n else 1)), "modulate_basic_batch")
        return out

    def alloc_results(self, n, soft_stride=148, device=None):
        d = device or self.device
        return dict(rc=torch.zeros(n, dtype=torch.int32, device=d), amp=torch.zeros((n, 2), dtype=torch.float32, device=d),
                    toa=torch.zeros(n, dtype=torch.float32, device=d), tsc=torch.zeros(n, dtype=torch.uint8, device=d),
                    ci=torch.zeros(n, dtype=torch.float32, device=d), flags=torch.zeros(n, dtype=torch.uint8, device=d),
                    soft=torch.zeros((n, soft_stride), dtype=torch.float32, device=d))

    def detect(self, bursts, type_, tsc, max_toa, max_toa_bound, thresh=BURST_THRESH, out=None):
        """bursts float32 [n, stride>=625, 2]; type_/tsc uint8 [n]; max_toa int16/uint16 [n] (device)."""
        _chk_dev(bursts, type_, tsc, max_toa)
        n = bursts.shape[0]
        r = out or self.alloc_results(n, 1)
        self.use_current_stream()
        self._check(self.lib.trxb200_detect_batch(self.h, _ptr(bursts), C.c_int(bursts.stride(0) // 2), C.c_int(n),
                                                  _ptr(type_), _ptr(tsc), _ptr(max_toa), C.c_int(max_toa_bound),
                                                  C.c_float(thresh), _ptr(r["rc"]), _ptr(r["amp"]), _ptr(r["toa"]),
                                                  _ptr(r["tsc"]), _ptr(r["ci"]), _ptr(r["flags"])), "detect_batch")
        return r

    def demod(self, bursts, rc, amp, toa, ci, soft=None, n_gmsk_soft=148, soft_stride=None):
        _chk_dev(bursts, rc, amp, toa, ci, soft)
        n = bursts.shape[0]
        if soft is None:
            soft = torch.zeros((n, soft_stride or n_gmsk_soft), dtype=torch.float32, device=bursts.device)
        self.use_current_stream()
        self._check(self.lib.trxb200_demod_batch(self.h, _ptr(bursts), C.c_int(bursts.stride(0) // 2), C.c_int(n),
                                                 _ptr(rc), _ptr(amp), _ptr(toa), _ptr(ci), _ptr(soft),
                                                 C.c_int(soft.stride(0)), C.c_int(n_gmsk_soft)), "demod_batch")
        return soft

    def detect_demod(self, bursts, type_, tsc, max_toa, max_toa_bound, thresh=BURST_THRESH, n_gmsk_soft=148,
                     soft_stride=None, out=None):
        _chk_dev(bursts, type_, tsc, max_toa)
        n = bursts.shape[0]
        r = out or self.alloc_results(n, soft_stride or n_gmsk_soft)
        self.use_current_stream()
        self._check(self.lib.trxb200_detect_demod_batch(
            self.h, _ptr(bursts), C.c_int(bursts.stride(0) // 2), C.c_int(n), _ptr(type_), _ptr(tsc), _ptr(max_toa),
            C.c_int(max_toa_bound), C.c_float(thresh), _ptr(r["rc"]), _ptr(r["amp"]), _ptr(r["toa"]), _ptr(r["tsc"]),
            _ptr(r["ci"]), _ptr(r["flags"]), _ptr(r["soft"]), C.c_int(r["soft"].stride(0)), C.c_int(n_gmsk_soft)),
            "detect_demod_batch")
        return r

    def detect_demod_sps1(self, bursts, type_, tsc, max_toa, max_toa_bound, thresh=BURST_THRESH, n_gmsk_soft=148,
                          soft_stride=None, out=None):
        """detectAnyBurst + demodAnyBurst at ONE sample per symbol (rx_sps = 1): bursts float32 [n, blen, 2] with
        148 <= blen <= 160 (a slot is 156 or 157 symbols); results as detect_demod."""
        _chk_dev(bursts, type_, tsc, max_toa)
        n, blen = bursts.shape[0], bursts.shape[1]
        r = out or self.alloc_results(n, soft_stride or n_gmsk_soft)
        self.use_current_stream()
        self._check(self.lib.trxb200_detect_sps1_batch(
            self.h, _ptr(bursts), C.c_int(bursts.stride(0) // 2), C.c_int(blen), C.c_int(n), _ptr(type_), _ptr(tsc),
            _ptr(max_toa), C.c_int(max_toa_bound), C.c_float(thresh), _ptr(r["rc"]), _ptr(r["amp"]), _ptr(r["toa"]),
            _ptr(r["tsc"]), _ptr(r["ci"]), _ptr(r["flags"])), "detect_sps1_batch")
        self._check(self.lib.trxb200_demod_sps1_batch(
            self.h, _ptr(bursts), C.c_int(bursts.stride(0) // 2), C.c_int(blen), C.c_int(n), _ptr(r["rc"]), _ptr(r["amp"]),
            _ptr(r["toa"]), _ptr(r["ci"]), _ptr(r["soft"]), C.c_int(r["soft"].stride(0)), C.c_int(n_gmsk_soft)),
            "demod_sps1_batch")
        return r

    def detect_demod_host(self, bursts, type_, tsc, max_toa, max_toa_bound, out, thresh=BURST_THRESH, n_gmsk_soft=148):
        """Same with HOST tensors (ideally pinned); copies are inside the call.  `out` from alloc_results(device='cpu')."""
        n = bursts.shape[0]
        self._check(self.lib.trxb200_detect_demod_host(
            self.h, _ptr(bursts), C.c_int(bursts.stride(0) // 2), C.c_int(n), _ptr(type_), _ptr(tsc), _ptr(max_toa),
            C.c_int(max_toa_bound), C.c_float(thresh), _ptr(out["rc"]), _ptr(out["amp"]), _ptr(out["toa"]),
            _ptr(out["tsc"]), _ptr(out["ci"]), _ptr(out["flags"]), _ptr(out["soft"]), C.c_int(out["soft"].stride(0)),
            C.c_int(n_gmsk_soft)), "detect_demod_host")
        return out

    # -- receive chain around the hot path: int16 slots -> TRXD uplink datagrams --
    def alloc_pull_results(self, n, pkt_stride=160, device=None, extras=True):
        d = device or self.device
        r = dict(rc=torch.zeros(n, dtype=torch.int32, device=d), energy=torch.zeros(n, dtype=torch.float32, device=d),
                 pkt=torch.zeros((n, pkt_stride), dtype=torch.uint8, device=d),
                 pkt_len=torch.zeros(n, dtype=torch.int16, device=d), flags=torch.zeros(n, dtype=torch.uint8, device=d))
        if extras:
            r.update(amp=torch.zeros((n, 2), dtype=torch.float32, device=d), toa=torch.zeros(n, dtype=torch.float32, device=d),
                     ci=torch.zeros(n, dtype=torch.float32, device=d), tsc=torch.zeros(n, dtype=torch.uint8, device=d))
        return r

    def _pull_args(self, iq, type_, tsc, max_toa, fn, tn, max_toa_bound, out, thresh, full_scale, rssi_offset, version):
        a = PullArgs()
        a.iq = _ptr(iq); a.stride = iq.stride(0) // 2; a.n = iq.shape[0]
        a.type = _ptr(type_); a.tsc = _ptr(tsc); a.max_toa = _ptr(max_toa); a.fn = _ptr(fn); a.tn = _ptr(tn)
        a.max_toa_bound = max_toa_bound; a.thresh = thresh; a.rx_full_scale = full_scale; a.rssi_offset = rssi_offset
        a.trxd_version = version
        a.rc = _ptr(out["rc"]); a.energy = _ptr(out["energy"]); a.pkt = _ptr(out["pkt"]); a.pkt_stride = out["pkt"].stride(0)
        a.pkt_len = _ptr(out["pkt_len"]); a.flags = _ptr(out.get("flags"))
        a.amp = _ptr(out.get("amp")); a.toa = _ptr(out.get("toa")); a.ci = _ptr(out.get("ci")); a.tsc_out = _ptr(out.get("tsc"))
        return a

    def pull(self, iq, type_, tsc, max_toa, fn, tn, max_toa_bound, out=None, thresh=BURST_THRESH, full_scale=32767.0,
             rssi_offset=0.0, version=1, pkt_stride=160):
        """iq int16 [n, stride>=625, 2] (device); fn int32/uint32 [n]; tn uint8 [n].  Datagrams in out['pkt']."""
        _chk_dev(iq, type_, tsc, max_toa, fn, tn)
        r = out or self.alloc_pull_results(iq.shape[0], pkt_stride)
        self.use_current_stream()
        a = self._pull_args(iq, type_, tsc, max_toa, fn, tn, max_toa_bound, r, thresh, full_scale, rssi_offset, version)
        self._check(self.lib.trxb200_pull_batch(self.h, C.byref(a)), "pull_batch")
        return r

    def pull_host(self, iq, type_, tsc, max_toa, fn, tn, max_toa_bound, out, thresh=BURST_THRESH, full_scale=32767.0,
                  rssi_offset=0.0, version=1):
        """Same with HOST tensors (ideally pinned); copies are inside the call."""
        a = self._pull_args(iq, type_, tsc, max_toa, fn, tn, max_toa_bound, out, thresh, full_scale, rssi_offset, version)
        self._check(self.lib.trxb200_pull_host(self.h, C.byref(a)), "pull_host")
        return out

    # -- SCH search (MS side) --
    def detect_sch(self, bursts, thresh=BURST_THRESH):
        """detectSCHBurst(SCH_DETECT_FULL) per burst.  bursts float32 [n, stride>=625, 2] (device)."""
        _chk_dev(bursts)
        n = bursts.shape[0]
        d = bursts.device
        r = dict(rc=torch.zeros(n, dtype=torch.int32, device=d), amp=torch.zeros((n, 2), dtype=torch.float32, device=d),
                 toa=torch.zeros(n, dtype=torch.float32, device=d), ci=torch.zeros(n, dtype=torch.float32, device=d),
                 flags=torch.zeros(n, dtype=torch.uint8, device=d))
        self.use_current_stream()
        self._check(self.lib.trxb200_detect_sch_batch(self.h, _ptr(bursts), C.c_int(bursts.stride(0) // 2), C.c_int(n),
                                                      C.c_float(thresh), _ptr(r["rc"]), _ptr(r["amp"]), _ptr(r["toa"]),
                                                      _ptr(r["ci"]), _ptr(r["flags"])), "detect_sch_batch")
        return r

    # -- burst-type scheduler --
    def detect_sch_buffer(self, bufs, in_len=60000, thresh=BURST_THRESH):
        """detectSCHBurst(SCH_DETECT_BUFFER): bufs float32 [n, stride >= in_len, 2] captures -> rc, amp, toa, ci, flags"""
        _chk_dev(bufs)
        n, d = bufs.shape[0], bufs.device
        r = dict(rc=torch.zeros(n, dtype=torch.int32, device=d), amp=torch.zeros((n, 2), dtype=torch.float32, device=d),
                 toa=torch.zeros(n, dtype=torch.float32, device=d), ci=torch.zeros(n, dtype=torch.float32, device=d),
                 flags=torch.zeros(n, dtype=torch.uint8, device=d))
        self.use_current_stream()
        self._check(self.lib.trxb200_detect_sch_buffer_batch(self.h, _ptr(bufs), C.c_int(bufs.stride(0) // 2), C.c_int(in_len),
                                                             C.c_int(n), C.c_float(thresh), _ptr(r["rc"]), _ptr(r["amp"]),
                                                             _ptr(r["toa"]), _ptr(r["ci"]), _ptr(r["flags"])), "detect_sch_buffer_batch")
        return r

    def expected_corr_type(self, fn, tn, chan_type, handover, chan=None, ext_rach=False, egprs=False, max_toa_nb=4,
                           max_toa_ab=63):
        """Transceiver::expectedCorrType per slot.  fn int32/uint32 [n], tn uint8 [n], chan_type uint8 [n_chan, 8]
        (ChannelCombination per timeslot), handover uint8 [8] (sub-slot bit mask per timeslot), chan int16 [n] or None.
        -> (type uint8 [n], max_toa int16 [n]) ready for pull() / detect()."""
        _chk_dev(fn, tn, chan_type, handover)
        n = fn.shape[0]
        typ = torch.empty(n, dtype=torch.uint8, device=fn.device)
        mt = torch.empty(n, dtype=torch.int16, device=fn.device)
        cfg = SchedCfg()
        cfg.n_chan = chan_type.shape[0]; cfg.chan_type = _ptr(chan_type); cfg.handover = _ptr(handover)
        cfg.ext_rach = int(ext_rach); cfg.egprs = int(egprs); cfg.max_toa_nb = int(max_toa_nb); cfg.max_toa_ab = int(max_toa_ab)
        self.use_current_stream()
        self._check(self.lib.trxb200_expected_corr_type_batch(self.h, C.byref(cfg), _ptr(fn), _ptr(tn),
                                                              _ptr(chan) if chan is not None else None, C.c_int(n), _ptr(typ),
                                                              _ptr(mt)), "expected_corr_type_batch")
        return typ, mt

    # -- helpers --
    def energy_detect(self, bursts, window, blen=BURST_LEN):
        _chk_dev(bursts)
        n = bursts.shape[0]
        e = torch.empty(n, dtype=torch.float32, device=bursts.device)
        self.use_current_stream()
        self._check(self.lib.trxb200_energy_detect_batch(self.h, _ptr(bursts), C.c_int(bursts.stride(0) // 2),
                                                         C.c_int(blen), C.c_int(n), C.c_uint(window), _ptr(e)),
                    "energy_detect_batch")
        return e

    def vector_slicer(self, src):
        _chk_dev(src)
        dst = torch.empty_like(src)
        self.use_current_stream()
        self._check(self.lib.trxb200_vector_slicer(self.h, _ptr(dst), _ptr(src), C.c_size_t(src.numel())), "vector_slicer")
        return dst

    def delay_vector(self, x, delay):
        """x float32 [n, len, 2], delay float32 [n] -> [n, len, 2]"""
        _chk_dev(x, delay)
        out = torch.empty_like(x)
        self.use_current_stream()
        self._check(self.lib.trxb200_delay_vector_batch(self.h, _ptr(x), C.c_int(x.stride(0) // 2), C.c_int(x.shape[1]),
                                                        C.c_int(x.shape[0]), _ptr(delay), _ptr(out),
                                                        C.c_int(out.stride(0) // 2)), "delay_vector_batch")
        return out

    def convolve(self, x, x_off, x_len, h, start, length, complex_taps, base=False):
        """x float32 [n, row, 2] where the addressed vector starts x_off samples into each row (head-room)."""
        _chk_dev(x, h)
        n = x.shape[0]
        y = torch.zeros((n, length, 2), dtype=torch.float32, device=x.device)
        fn = self.lib.trxb200_convolve_complex_batch if complex_taps else self.lib.trxb200_convolve_real_batch
        self.use_current_stream()
        rc = fn(self.h, C.c_void_p(x.data_ptr() + 8 * x_off), C.c_int(x_len), C.c_int(x.stride(0) // 2), _ptr(h),
                C.c_int(h.shape[0]), _ptr(y), C.c_int(length), C.c_int(y.stride(0) // 2), C.c_int(start),
                C.c_int(length), C.c_int(n), C.c_int(int(base)))
        return rc, y

    def convert_float_short(self, x, scale, mode=0, out=None):
        """mode 0: SSE semantics (round to nearest even, saturate); 1: the x86 dispatcher for any length (truncating
        scalar tail of len % 8); 2: base_convert_float_short (truncation)"""
        _chk_dev(x, out)
        if out is None:
            out = torch.empty(x.numel(), dtype=torch.int16, device=x.device)
        self.use_current_stream()
        self._check(self.lib.trxb200_convert_float_short_mode(self.h, _ptr(out), _ptr(x), C.c_float(scale),
                                                              C.c_size_t(x.numel()), C.c_int(mode)), "convert_float_short")
        return out

    def convert_short_float(self, x):
        _chk_dev(x)
        out = torch.empty(x.numel(), dtype=torch.float32, device=x.device)
        self.use_current_stream()
        self._check(self.lib.trxb200_convert_short_float(self.h, _ptr(out), _ptr(x), C.c_size_t(x.numel())),
                    "convert_short_float")
        return out

    # -- vitac --
    def vitac(self, bufs, offset, tsc, is_ab=False, max_delay=0, clamp=(-39, 39), want_cir=False, out=None):
        _chk_dev(bufs) if tsc is None else _chk_dev(bufs, tsc)
        n = bufs.shape[0]
        nb = 88 if int(is_ab) == 1 else 148  # is_ab: 0 normal, 1 access, 2 SCH burst (tsc unused)
        d = bufs.device
        r = out if out is not None else dict(
            bits=torch.zeros((n, nb), dtype=torch.int8, device=d), start=torch.zeros(n, dtype=torch.int32, device=d),
            corr_max=torch.zeros(n, dtype=torch.float32, device=d),
            cir=torch.zeros((n, 20, 2), dtype=torch.float32, device=d) if want_cir else None)
        self.use_current_stream()
        self._check(self.lib.trxb200_vitac_batch(self.h, _ptr(bufs), C.c_int(bufs.stride(0) // 2), C.c_int(offset),
                                                 C.c_int(n), C.c_int(int(is_ab)), _ptr(tsc), C.c_int(max_delay),
                                                 C.c_int(clamp[0]), C.c_int(clamp[1]), _ptr(r["bits"]), _ptr(r["start"]),
                                                 _ptr(r["corr_max"]), _ptr(r["cir"])), "vitac_batch")
        return r


def _vitac_sch_buffer(self, bufs, offset, length, want_bits=True):
    """First SCH acquisition (get_sch_buffer_chan_imp_resp + detect_burst_nb): bufs float32 [n, stride, 2], the capture of
    `length` samples starts at sample `offset` of every row.  -> dict(bits int8 [n,148], start, corr_max, cir [n,20,2])"""
    _chk_dev(bufs)
    n, d = bufs.shape[0], bufs.device
    r = dict(bits=torch.zeros((n, 148), dtype=torch.int8, device=d) if want_bits else None,
             start=torch.zeros(n, dtype=torch.int32, device=d), corr_max=torch.zeros(n, dtype=torch.float32, device=d),
             cir=torch.zeros((n, 20, 2), dtype=torch.float32, device=d))
    self.use_current_stream()
    self._check(self.lib.trxb200_vitac_sch_buffer_batch(self.h, _ptr(bufs), C.c_int(bufs.stride(0) // 2), C.c_int(offset),
                                                        C.c_int(length), C.c_int(n), _ptr(r["bits"]), _ptr(r["start"]),
                                                        _ptr(r["corr_max"]), _ptr(r["cir"])), "vitac_sch_buffer_batch")
    return r


Trx.vitac_sch_buffer = _vitac_sch_buffer


def _vitac_detect(self, bufs, offset, cir, start, is_ab=False, clamp=(-39, 39), ss=3):
    """detect_burst_nb / detect_burst_ab with the caller's channel estimate: cir float32 [n, 20, 2], start int32 [n];
    ss: the Viterbi detector's start state (the five-argument forms of the reference)."""
    _chk_dev(bufs, cir, start)
    n = bufs.shape[0]
    bits = torch.zeros((n, 88 if is_ab else 148), dtype=torch.int8, device=bufs.device)
    self.use_current_stream()
    self._check(self.lib.trxb200_vitac_detect_ss_batch(self.h, _ptr(bufs), C.c_int(bufs.stride(0) // 2), C.c_int(offset), C.c_int(n),
                                                       C.c_int(int(is_ab)), _ptr(cir), _ptr(start), C.c_int(clamp[0]),
                                                       C.c_int(clamp[1]), C.c_int(ss), _ptr(bits)), "vitac_detect_ss_batch")
    return bits


Trx.vitac_detect = _vitac_detect


def build(force=False):
    """Compile libtrxb200.so in-tree (nvcc, sm_100a)."""
    return _build.build(force=force)


class Resampler:
    """Rational p/q polyphase resampler (Resampler.h:31-61) on device streams."""

    def __init__(self, trx, p, q, filt_len=16, bw=1.0):
        self.trx, self.p, self.q, self.filt_len = trx, p, q, filt_len
        h = C.c_void_p()
        trx._check(trx.lib.trxb200_resampler_create(trx.h, C.c_int(p), C.c_int(q), C.c_int(filt_len), C.c_float(bw),
                                                    C.byref(h)), "resampler_create")
        self.h = h

    def __del__(self):
        if getattr(self, "h", None):
            self.trx.lib.trxb200_resampler_destroy(self.h)
            self.h = None

    def taps(self, path):
        import numpy as np
        out = np.zeros(self.filt_len, np.float32)
        self.trx.lib.trxb200_resampler_taps(self.h, C.c_int(path), out.ctypes.data_as(C.c_void_p))
        return out

    def rotate_streams(self, x, first, in_len, in_stride, n_streams, out, out_len):
        """Lower-level form for streams that already lie in one buffer: x float32 [..., 2] (flat samples); stream s reads
        its in_len new samples at sample index first + s * in_stride, with filt_len samples of history before them in
        the buffer; out float32 [n_streams, out_len, 2].  Consecutive segments of one long stream (in_stride == in_len)
        are each other's history, so a long rotate equals the reference's block-by-block calls whenever
        q * out_len / p == in_len per segment (Resampler.cpp:131-150 restarts its path indices at every call)."""
        _chk_dev(x, out)
        self.trx.use_current_stream()
        fn = self.trx.lib.trxb200_resampler_rotate_stream if out_len > 4096 else self.trx.lib.trxb200_resampler_rotate
        rc = fn(self.h, C.c_void_p(x.data_ptr() + 8 * first), C.c_int(in_len), C.c_int(in_stride),
                _ptr(out), C.c_int(out_len), C.c_int(out.stride(0) // 2), C.c_int(n_streams))
        self.trx._check(rc, "resampler_rotate")
        return out

    def rotate(self, x, out_len, out=None):
        """x float32 [n_streams, filt_len + in_len, 2]: history then the new block. -> [n_streams, out_len, 2]"""
        _chk_dev(x)
        ns, tot = x.shape[0], x.shape[1]
        if out is None:
            out = torch.empty((ns, out_len, 2), dtype=torch.float32, device=x.device)
        self.trx.use_current_stream()
        rc = self.trx.lib.trxb200_resampler_rotate(self.h, C.c_void_p(x.data_ptr() + 8 * self.filt_len),
                                                   C.c_int(tot - self.filt_len), C.c_int(x.stride(0) // 2), _ptr(out),
                                                   C.c_int(out_len), C.c_int(out.stride(0) // 2), C.c_int(ns))
        self.trx._check(rc, "resampler_rotate")
        return out


class _Filterbank:
    def __init__(self, trx, m, block_len, h_len, synth):
        self.trx, self.m, self.block_len, self.h_len = trx, m, block_len, h_len
        h = C.c_void_p()
        fn = trx.lib.trxb200_synthesis_create if synth else trx.lib.trxb200_channelizer_create
        trx._check(fn(trx.h, C.c_int(m), C.c_int(block_len), C.c_int(h_len), C.byref(h)), "filterbank_create")
        self.h = h

    def __del__(self):
        if getattr(self, "h", None):
            self.trx.lib.trxb200_filterbank_destroy(self.h)
            self.h = None

    def reset(self):
        self.trx.use_current_stream()
        self.trx._check(self.trx.lib.trxb200_filterbank_reset(self.h), "filterbank_reset")

    def taps(self, branch):
        import numpy as np
        out = np.zeros(self.h_len, np.float32)
        self.trx.lib.trxb200_filterbank_taps(self.h, C.c_int(branch), out.ctypes.data_as(C.c_void_p))
        return out


class Channelizer(_Filterbank):
    """M-channel analysis filterbank (Channelizer.h:13-31). rotate(): [n_blocks*block_len*m, 2] -> [m, n_blocks*block_len, 2]"""

    def __init__(self, trx, m, block_len, h_len=16):
        super().__init__(trx, m, block_len, h_len, False)

    def rotate(self, x, out=None):
        _chk_dev(x)
        nb = x.shape[0] // (self.m * self.block_len)
        if out is None:
            out = torch.empty((self.m, nb * self.block_len, 2), dtype=torch.float32, device=x.device)
        self.trx.use_current_stream()
        self.trx._check(self.trx.lib.trxb200_channelizer_rotate(self.h, _ptr(x), _ptr(out), C.c_int(nb)), "channelizer_rotate")
        return out

    def rotate_into(self, x, out, col0=0):
        """x [total_t * m, 2] (any total_t >= h_len), out [m, pitch, 2]: channel c's samples go to out[c, col0:col0 + total_t]
        (room in front of each row for the resampler's history)."""
        _chk_dev(x, out)
        total_t = x.shape[0] // self.m
        assert out.shape[0] == self.m and out.stride(1) == 2 and col0 + total_t <= out.shape[1]
        self.trx.use_current_stream()
        rc = self.trx.lib.trxb200_channelizer_rotate_strided(self.h, _ptr(x), C.c_long(total_t), C.c_void_p(out.data_ptr() + 8 * col0),
                                                            C.c_long(out.stride(0) // 2))
        self.trx._check(rc, "channelizer_rotate_strided")
        return out

    def prime(self, prev_in):
        """Set the carried history from the last h_len rows of prev_in [n_prev_t * m, 2] (time-block sharding: the halo
        is re-read from the source instead of carried from the previous call)."""
        _chk_dev(prev_in)
        self.trx.use_current_stream()
        self.trx._check(self.trx.lib.trxb200_channelizer_prime(self.h, _ptr(prev_in), C.c_long(prev_in.shape[0] // self.m)),
                        "channelizer_prime")


class Synthesis(_Filterbank):
    """M-channel synthesis filterbank (Synthesis.h:13-32). rotate(): [m, n_blocks*block_len, 2] -> [n_blocks*block_len*m, 2]"""

    def __init__(self, trx, m, block_len, h_len=16):
        super().__init__(trx, m, block_len, h_len, True)

    def rotate(self, x):
        _chk_dev(x)
        nb = x.shape[1] // self.block_len
        out = torch.empty((nb * self.block_len * self.m, 2), dtype=torch.float32, device=x.device)
        self.trx.use_current_stream()
        self.trx._check(self.trx.lib.trxb200_synthesis_rotate(self.h, _ptr(x), _ptr(out), C.c_int(nb)), "synthesis_rotate")
        return out


from .wideband import WidebandRx  # noqa: E402,F401  (after Channelizer / Resampler are defined)
