#!/usr/bin/env python
"""Golden vectors for detectSCHBurst(SCH_DETECT_FULL), generated HERE from the unmodified reference (oracle/_ref):

    python tests/golden/make_sch_fixture.py   ->  tests/golden/sch_fixture.npz  (48 bursts, inputs stored as float16-exact
    values so that the file stays small: the inputs are rounded to float16 BEFORE the reference sees them)"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import cpulibs  # noqa: E402
import synth  # noqa: E402


def main():
    ref = cpulibs.Ref()
    rng = np.random.default_rng(46)
    n = 48
    w = ref.modulate_gmsk_batch(synth.sch_bits(n, rng))
    rx, _ = synth.impair(w, rng, snr_db=np.choose(np.arange(n) % 3, [25.0, 10.0, 5.0]), noise_only_frac=0.15, shift_lo=-60, shift_hi=30)
    rx16 = rx.astype(np.float16)
    r = ref.detect_sch(rx16.astype(np.float32))
    np.savez_compressed(os.path.join(HERE, "sch_fixture.npz"), rx=rx16, rc=r["rc"], amp=r["amp"], toa=r["toa"], ci=r["ci"])
    print("wrote sch_fixture.npz: detected", int((r["rc"] > 0).sum()), "of", n)


if __name__ == "__main__":
    main()
