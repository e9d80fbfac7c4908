#!/usr/bin/env python
"""Golden vectors for the first SCH acquisition, generated HERE from the unmodified reference (oracle/_ref):
detectSCHBurst(SCH_DETECT_BUFFER) and get_sch_buffer_chan_imp_resp + detect_burst_nb over short captures.

    python tests/golden/make_sch_buffer_fixture.py  ->  tests/golden/sch_buffer_fixture.npz

detectSCHBurst's BUFFER state has a fixed size (12 frames, 60,000 samples): three such captures.  get_sch_buffer_chan_imp_resp
takes its length as an argument: ten captures of 2 frames keep the file small.  Values are stored float16-exact: the inputs
are rounded to float16 BEFORE the reference sees them."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import cpulibs  # noqa: E402
import synth  # noqa: E402


def captures(lib, rng, n, length, head):
    """n captures of `length` samples behind `head` samples of head-room; one SCH burst somewhere inside most of them"""
    tx = lib.modulate_gmsk_batch(synth.sch_bits(n, rng))
    buf = (rng.standard_normal((n, head + length + 64, 2)) * 0.03).astype(np.float32)
    pos = rng.integers(0, length - 1300, n)
    pos[0], pos[1] = 3, length - 700           # at the very start / running off the search range
    for b in range(n):
        if b % 5 != 4:                          # every fifth capture holds noise only
            g = rng.uniform(0.2, 1.0) * np.exp(1j * rng.uniform(0, 2 * np.pi))
            w = (tx[b, :, 0] + 1j * tx[b, :, 1]) * g
            buf[b, head + pos[b]: head + pos[b] + 625, 0] += w.real.astype(np.float32)
            buf[b, head + pos[b]: head + pos[b] + 625, 1] += w.imag.astype(np.float32)
    return buf, pos


def main():
    ref = cpulibs.Ref()
    rng = np.random.default_rng(48)
    n, length, head = 10, 10000, 192
    buf, pos = captures(ref, rng, n, length, head)
    b16 = buf.astype(np.float16)
    v = ref.vitac_sch_buffer(b16.astype(np.float32), head, length)
    cap, cpos = captures(ref, rng, 3, 60000, 0)
    c16 = cap[:, :60000].astype(np.float16)
    d = ref.detect_sch_buffer(c16.astype(np.float32), 60000)
    np.savez_compressed(os.path.join(HERE, "sch_buffer_fixture.npz"), buf=b16, head=head, length=length, pos=pos, cap=c16, cpos=cpos,
                        d_rc=d["rc"], d_amp=d["amp"], d_toa=d["toa"], d_ci=d["ci"],
                        v_bits=v["bits"], v_start=v["start"], v_corr_max=v["corr_max"], v_cir=v["cir"])
    print("wrote sch_buffer_fixture.npz: detected", int((d["rc"] > 0).sum()), "of 3", d["toa"], cpos / 4.0, "starts", v["start"], pos)


if __name__ == "__main__":
    main()
