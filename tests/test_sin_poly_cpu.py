"""CPU check of the polynomial sine kept as a compile-time option of the sine epilogues (csrc/tc_core.cuh sin_poly2, B2R_SIN_POLY_PAIRS,
default 0): the same fp32 arithmetic restated in numpy -- magic-number rounding of x / 2 pi, degree-4 polynomial in r^2 with the
coefficients read from the CUDA source -- stays within 7e-6 of sin(x) over the argument range the FiLM / SIREN layers produce."""
import os
import re
import struct

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _hex_floats():
    src = open(os.path.join(ROOT, "msra_practice_project_b200", "csrc", "tc_core.cuh")).read()
    body = src[src.index("void sin_poly2("):]
    body = body[:body.index("}\n")]
    return [struct.unpack("<f", struct.pack("<I", int(h, 16)))[0] for h in re.findall(r"0f([0-9A-Fa-f]{8})", body)]


def test_sin_poly2_coefficients_and_error():
    c = _hex_floats()
    # 1 / 2 pi, 1.5 * 2^23, -1, then the polynomial from the highest coefficient down
    assert len(c) == 8
    inv2pi, magic, neg1 = (np.float32(v) for v in c[:3])
    assert abs(float(inv2pi) - 1.0 / (2.0 * np.pi)) < 1e-8 and float(magic) == 12582912.0 and float(neg1) == -1.0
    coef = [np.float32(v) for v in c[3:]]
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-300.0, 300.0, 400000), np.linspace(-7.0, 7.0, 100001)]).astype(np.float32)
    # fma.rn.f32 emulated in float64 and rounded once (the products of two fp32 values are exact in fp64; the one sum may round
    # twice in rare cases, far below the tolerance checked here)
    fma = lambda a, b, d: (a.astype(np.float64) * np.float64(b) + d.astype(np.float64)).astype(np.float32) if not np.isscalar(d) else \
        (a.astype(np.float64) * np.float64(b) + np.float64(d)).astype(np.float32)
    m = fma(x, inv2pi, magic)                                     # u + magic: rounds u to an integer k
    nk = fma(m, neg1, magic)                                      # -k
    r = fma(x, inv2pi, nk)                                        # u - k
    assert np.all(np.abs(r) <= 0.5 + 1e-6)
    s = (r.astype(np.float64) * r.astype(np.float64)).astype(np.float32)
    p = np.full_like(s, coef[0])
    for k in coef[1:]:
        p = (p.astype(np.float64) * s.astype(np.float64) + np.float64(k)).astype(np.float32)
    out = (p.astype(np.float64) * r.astype(np.float64)).astype(np.float32)
    err = np.abs(out.astype(np.float64) - np.sin(x.astype(np.float64)))
    # the argument reduction itself costs |x| * 2^-24 * 2 pi for large |x| (as it does for MUFU.SIN's range reduction)
    bound = 7e-6 + np.abs(x.astype(np.float64)) * 2.0 ** -24 * 2 * np.pi * 1.5
    assert np.all(err <= bound), float((err - bound).max())
