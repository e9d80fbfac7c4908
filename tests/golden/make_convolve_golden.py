"""Extract the reference's own known-answer vectors for convolve_* into a small .npz fixture.

Source: /root/reference/tests/Transceiver52M/convolve_test_golden.h (12 arrays) and the LCG input generator
of convolve_test.c:23-45 (re-implemented here; the generated x/h are cross-checked against
convolve_test.ok).  Run in the dev container only: python tests/golden/make_convolve_golden.py
"""
import os
import re
import numpy as np

REF = "/root/reference/tests/Transceiver52M"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "convolve_golden.npz")


def parse_arrays(text):
    out = {}
    for m in re.finditer(r"static const float (\w+)\[\] = \{(.*?)\};", text, re.S):
        vals = [float(v.rstrip("f")) for v in re.findall(r"[-+]?\d\.\d+e[-+]\d+f", m.group(2))]
        out[m.group(1)] = np.array(vals, np.float32)
    return out


def lcg_floats(n, state):
    out = np.zeros(n, np.float32)
    for i in range(n):
        state = (1103515245 * state + 12345) & 0x7FFFFFFF
        u = state
        e = 112 + ((u ^ (u >> 8)) & 15)
        bits = (u & 0x007FFFFF) | ((u & 0x00800000) << 8) | ((e & 0xFF) << 23)
        out[i] = np.array([bits], np.uint32).view(np.float32)[0]
    return out, state


def main():
    gold = parse_arrays(open(os.path.join(REF, "convolve_test_golden.h")).read())
    ok = parse_arrays(open(os.path.join(REF, "convolve_test.ok")).read())
    x, st = lcg_floats(200, 0)
    h, _ = lcg_floats(50, st)
    # the .ok file prints with 8 significant digits: compare at that precision
    assert np.allclose(x, ok["x"], rtol=1e-7, atol=0) and np.allclose(h, ok["h"], rtol=1e-7, atol=0)
    d = {"x": x, "h": h}
    for k, v in gold.items():
        d[k.replace("y_ref_", "y_")] = v
    np.savez(OUT, **d)
    print("wrote", OUT, sorted(d))


if __name__ == "__main__":
    main()
