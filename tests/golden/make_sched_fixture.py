#!/usr/bin/env python
"""Golden vectors for the burst-type scheduler, generated HERE from the reference's own
Transceiver::expectedCorrType (oracle/_ref, cut out of Transceiver.cpp at build time by oracle/gen_ref_sched.py):

    python tests/golden/make_sched_fixture.py   ->  tests/golden/sched_fixture.npz

24 timeslot configurations (every ChannelCombination on every timeslot at least once, random handover masks, the
ext_rach / egprs flags in all four states), each over one full period of the multiframe structures
(lcm(26, 51, 52, 102) = 2652 frames) with the timeslot number rotating."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import cpulibs  # noqa: E402


def cases(seed=5):
    rng = np.random.default_rng(seed)
    out = []
    for k in range(24):
        ct = rng.integers(0, 16, 8).astype(np.uint8)
        ct[k % 8] = k % 16            # every combination on every timeslot index over the set
        ho = rng.integers(0, 256, 8).astype(np.uint8)
        if k % 5 == 0:
            ho[:] = 0
        fn = np.arange(2652, dtype=np.uint32) + np.uint32(rng.integers(0, 1000) * 2652)
        tn = ((np.arange(2652) + k) % 8).astype(np.uint8)
        out.append(dict(chan_type=ct, handover=ho, ext_rach=(k >> 0) & 1, egprs=(k >> 1) & 1, fn=fn, tn=tn))
    return out


def main():
    ref = cpulibs.Ref()
    cs = cases()
    d = {}
    for i, c in enumerate(cs):
        for key in ("chan_type", "handover", "fn", "tn"):
            d[f"{key}_{i}"] = c[key]
        d[f"flags_{i}"] = np.array([c["ext_rach"], c["egprs"]], np.uint8)
        d[f"type_{i}"] = ref.expected_corr_type(c["chan_type"], c["handover"], c["ext_rach"], c["egprs"], c["fn"], c["tn"])
    np.savez_compressed(os.path.join(HERE, "sched_fixture.npz"), n=np.array(len(cs)), **d)
    print("wrote sched_fixture.npz:", len(cs), "cases,", sum(len(c["fn"]) for c in cs), "slots; types seen",
          sorted(set(np.concatenate([d[f"type_{i}"] for i in range(len(cs))]).tolist())))


if __name__ == "__main__":
    main()
