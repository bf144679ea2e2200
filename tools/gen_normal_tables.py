#!/usr/bin/env python
"""Generates gym_fishing_b200/csrc/normal_tables.cuh: the lookup tables and polynomial coefficients of the table-assisted
Box-Muller in csrc/philox.cuh (round 2).  Run once; the generated header is committed.  Needs mpmath (a sympy dependency).

  log table, 257 entries, i = round(256 * (m - 1)), m in [1, 2) the significand of u1:
      c_i   = 1 + i/256,  rcp_i = RN(1 / c_i),  adj_i = (i >= ITHR)          (ITHR = 106: c >= sqrt 2 uses c/2 and e + 1)
      entry = { -2 * rcp_i,  T_i = RN(2 ln(rcp_i) + 2 adj_i ln 2) }
      so that  -2 ln(u1) = (e + adj_i) * (-2 ln 2) + T_i + g(rr),   rr = fma(m, -2 rcp_i, 2) = -2 (m rcp_i - 1),
               g(rr) = -2 log1p(-rr / 2) = rr + rr^2 * Q(rr),       |rr| <= 2^-8 (1 + 2^-44)
  trig table, 257 entries, j = round(256 * frac) of the second uniform u2 = frac + 2^-53:
      entry = { cos(2 pi (j/256 + 2^-53)), sin(2 pi (j/256 + 2^-53)) }        (the half-step 2^-53 lives in the table)
      sin(2 pi f) = f * PS(f^2),  cos(2 pi f) - 1 = f^2 * PC(f^2),            |f| <= 2^-9
Coefficients are Chebyshev-node interpolants computed in 60-digit arithmetic, then rounded to binary64; the script prints
the worst absolute error of each rounded polynomial over its interval."""
from __future__ import annotations

import os

import mpmath as mp

mp.mp.dps = 60
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gym_fishing_b200", "csrc", "normal_tables.cuh")
NLOG, NTRIG, ITHR = 256, 256, 106


def rn(x):
    """nearest binary64 as a Python float"""
    return float(mp.mpf(x))


def hexf(x):
    return float(x).hex()


def cheb_fit(fn, lo, hi, deg):
    """polynomial of degree `deg` through fn at the Chebyshev nodes of [lo, hi]; coefficients of x^0..x^deg (mpf)"""
    n = deg + 1
    xs = [(lo + hi) / 2 + (hi - lo) / 2 * mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
    A = mp.matrix(n, n)
    b = mp.matrix(n, 1)
    for r, x in enumerate(xs):
        for c in range(n):
            A[r, c] = x ** c
        b[r] = fn(x)
    sol = mp.lu_solve(A, b)
    return [sol[k] for k in range(n)]


def max_err(fn_exact, coefs, lo, hi, weight, samples=4001):
    worst = mp.mpf(0)
    for k in range(samples):
        x = lo + (hi - lo) * k / (samples - 1)
        p = mp.mpf(0)
        for c in reversed(coefs):
            p = p * x + mp.mpf(c)
        worst = max(worst, abs((p - fn_exact(x)) * weight(x)))
    return worst


def main():
    ln2 = mp.log(2)
    # ---- log table
    log_rows = []
    for i in range(NLOG + 1):
        c = mp.mpf(1) + mp.mpf(i) / NLOG
        rcp = rn(1 / c)
        adj = 1 if i >= ITHR else 0
        T = rn(2 * mp.log(mp.mpf(rcp)) + 2 * adj * ln2)
        if i == 0 or i == NLOG:
            assert T == 0.0 and rcp in (1.0, 0.5)
        log_rows.append((-2.0 * rcp, T))
    # rr range: m within 2^-9 of c_i, rcp_i <= 1 and rounded: |rr| <= 2 * 2^-9 * (1 + small)
    RR = mp.mpf(2) ** -8 * (1 + mp.mpf(2) ** -40)

    def Q(rr):
        if rr == 0:
            return mp.mpf(1) / 4
        return (-2 * mp.log1p(-rr / 2) - rr) / rr ** 2

    q = [rn(c) for c in cheb_fit(Q, -RR, RR, 3)]
    eq = max_err(Q, q, -RR, RR, lambda x: x * x)
    # ---- trig table
    trig_rows = []
    half = mp.mpf(2) ** -53
    for j in range(NTRIG + 1):
        ang = 2 * mp.pi * (mp.mpf(j) / NTRIG + half)
        trig_rows.append((rn(mp.cos(ang)), rn(mp.sin(ang))))
    F2 = (mp.mpf(2) ** -9) ** 2

    def PS(f2):
        f = mp.sqrt(f2)
        return 2 * mp.pi if f2 == 0 else mp.sin(2 * mp.pi * f) / f

    def PC(f2):
        f = mp.sqrt(f2)
        return -2 * mp.pi ** 2 if f2 == 0 else (mp.cos(2 * mp.pi * f) - 1) / f2

    ps = [rn(c) for c in cheb_fit(PS, mp.mpf(0), F2, 2)]
    pc = [rn(c) for c in cheb_fit(PC, mp.mpf(0), F2, 2)]
    es = max_err(PS, ps, mp.mpf(0), F2, lambda x: mp.sqrt(x))
    ec = max_err(PC, pc, mp.mpf(0), F2, lambda x: x)
    print(f"max |rr^2 (Q_fit - Q)|      = {mp.nstr(eq, 5)}  (2^{mp.nstr(mp.log(eq, 2), 5)})")
    print(f"max |f (PS_fit - PS)|       = {mp.nstr(es, 5)}  (2^{mp.nstr(mp.log(es, 2), 5)})")
    print(f"max |f^2 (PC_fit - PC)|     = {mp.nstr(ec, 5)}  (2^{mp.nstr(mp.log(ec, 2), 5)})")

    # ---- exp(x) for the tipping-point growth (fishing-v2): x = k ln2/64 + r, exp(x) = 2^(k >> 6) * 2^((k & 63)/64) * exp(r)
    exp_rows = [rn(mp.mpf(2) ** (mp.mpf(j) / 64)) for j in range(64)]
    RE = mp.log(2) / 128 * (1 + mp.mpf(2) ** -30)

    def PE(r):       # (exp(r) - 1 - r) / r^2
        return mp.mpf(1) / 2 if r == 0 else (mp.exp(r) - 1 - r) / r ** 2

    pe = [rn(c) for c in cheb_fit(PE, -RE, RE, 3)]
    ee = max_err(PE, pe, -RE, RE, lambda x: x * x)
    print(f"max |r^2 (PE_fit - PE)|     = {mp.nstr(ee, 5)}  (2^{mp.nstr(mp.log(ee, 2), 5)})")
    ln2_64 = mp.log(2) / 64
    ln2_64_hi = rn(mp.mpf(float(ln2_64)) - mp.mpf(float(ln2_64)) % mp.mpf(2) ** -40)   # 64/ln2 * hi is exact for |k| < 2^17
    ln2_64_lo = rn(ln2_64 - mp.mpf(ln2_64_hi))

    with open(OUT, "w") as fh:
        w = fh.write
        w("// GENERATED by tools/gen_normal_tables.py -- do not edit.  Tables and coefficients of the table-assisted Box-Muller\n"
          "// (csrc/philox.cuh).  Hex float literals are exact binary64 values.\n#pragma once\n\n")
        w(f"#define FISHING_NLOG {NLOG}\n#define FISHING_NTRIG {NTRIG}\n")
        w(f"// carry constant for e + adj: adds into the exponent field when the significand's high word >= {ITHR} * 4096 - 2048\n")
        w(f"#define FISHING_LOG_THR 0x{ITHR * 4096 - 2048:x}\n")
        w(f"#define FISHING_M2LN2 {hexf(rn(-2 * ln2))}   /* -2 ln 2 */\n")
        for name, vals in (("FISHING_Q", q), ("FISHING_PS", ps), ("FISHING_PC", pc)):
            for k, v in enumerate(vals):
                w(f"#define {name}{k} {hexf(v)}\n")
        w("\n// { -2 / c_i (rounded), 2 ln(rcp_i) + 2 adj_i ln 2 }\n")
        w(f"#define FISHING_LOG_TABLE_INIT {{ \\\n")
        for a, b in log_rows:
            w(f"  {{{hexf(a)}, {hexf(b)}}}, \\\n")
        w("}\n\n// exp(x): 64 / ln 2, -ln2/64 split hi + lo (hi has 24 trailing zero bits), (exp(r) - 1 - r) / r^2 cubic\n")
        w(f"#define FISHING_EXP_64_LN2 {hexf(rn(64 / mp.log(2)))}\n")
        w(f"#define FISHING_EXP_NLN2_64_HI {hexf(-ln2_64_hi)}\n#define FISHING_EXP_NLN2_64_LO {hexf(-ln2_64_lo)}\n")
        for k, v in enumerate(pe):
            w(f"#define FISHING_PE{k} {hexf(v)}\n")
        w("// 2^(j/64), j = 0..63\n#define FISHING_EXP_TABLE_INIT { \\\n")
        for v in exp_rows:
            w(f"  {hexf(v)}, \\\n")
        w("}\n\n// { cos, sin } of 2 pi (j/256 + 2^-53)\n")
        w(f"#define FISHING_TRIG_TABLE_INIT {{ \\\n")
        for a, b in trig_rows:
            w(f"  {{{hexf(a)}, {hexf(b)}}}, \\\n")
        w("}\n")
    print("wrote", OUT)


if __name__ == "__main__":
    main()
