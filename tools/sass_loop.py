"""Histogram of the SASS of one kernel's innermost hot loop, grouped by issue pipe.

usage: python tools/sass_loop.py <lib.so> <mangled-or-substring kernel name> [--all]

Finds the backward branch that closes the longest loop body (or, with --all, takes the whole kernel), and counts the
instructions of the body by opcode and by the pipe they issue to (FP64 / FMA-heavy integer multiply / ALU / LSU / XU / other).
Used for DESIGN.md's per-pair-step instruction budgets; the GPU is not needed."""
from __future__ import annotations

import collections
import re
import subprocess
import sys

PIPE = [
    ("fp64", r"^(DFMA|DADD|DMUL|DSETP|DMNMX)"),
    ("imad.wide", r"^IMAD\.WIDE"),
    ("imad.hi", r"^IMAD\.HI"),
    ("imad", r"^IMAD(\.|$)"),
    ("fma32", r"^(FFMA|FMUL|FADD)"),
    ("iadd", r"^(VIADD|IADD3|IADD)"),        # overlaps with LOP3 in the issue probe: not on the ALU's port
    ("xu", r"^(MUFU|F2F|I2F|F2I|FRND|POPC|FLO|BREV)"),
    ("lsu", r"^(LDG|STG|LDS|STS|LD\.|ST\.|ATOM|RED|LDSM|LDL|STL)"),
    ("uniform", r"^(U[A-Z]|LDCU|R2UR|S2UR|VOTEU)"),
    ("const", r"^LDC"),
    ("branch", r"^(BRA|EXIT|RET|CALL|BSSY|BSYNC|WARPSYNC|NOP|BAR|YIELD)"),
    ("alu", r".*"),
]


def kernels(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    cur, body = None, {}
    for line in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            cur = m.group(1)
            body[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur:
            body[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return body


def opcode(text):
    t = text.split()
    if t[0].startswith("@"):
        t = t[1:]
    return t[0]


PORT_A = ("fp64", "alu")     # the two busiest pipes of the fused loops (tools/issue_probe.cu: they run side by side)


def loop_mix(lib, pattern, whole=False):
    """{kernel name: dict(n=instructions of the hot loop, pipes={class: count}, ops={opcode: count})} for every kernel of
    `lib` whose mangled name contains `pattern`.  The hot loop = the body closed by the longest backward branch."""
    out = {}
    for name, ins in kernels(lib).items():
        if pattern not in name:
            continue
        lo, hi = 0, len(ins)
        if not whole:
            best = None
            for idx, (addr, text) in enumerate(ins):
                m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?`?\(?0x([0-9a-f]+)", text)
                if m:
                    tgt = int(m.group(1), 16)
                    if tgt < addr:
                        span = addr - tgt
                        if best is None or span > best[0]:
                            best = (span, tgt, addr)
            if best:
                lo = next(i for i, (a, _) in enumerate(ins) if a == best[1])
                hi = next(i for i, (a, _) in enumerate(ins) if a == best[2]) + 1
        ops = collections.Counter()
        pipes = collections.Counter()
        for _, text in ins[lo:hi]:
            op = opcode(text)
            ops[op] += 1
            for pname, rx in PIPE:
                if re.match(rx, op):
                    pipes[pname] += 1
                    break
        out[name] = dict(n=hi - lo, total=len(ins), pipes=dict(pipes), ops=dict(ops))
    return out


def main():
    lib, pat = sys.argv[1], sys.argv[2]
    for name, m in loop_mix(lib, pat, "--all" in sys.argv).items():
        print(f"== {name}\n   body = {m['n']} instructions (of {m['total']})")
        print("   pipes:", ", ".join(f"{k} {v}" for k, v in sorted(m["pipes"].items(), key=lambda kv: -kv[1])))
        print("   ops  :", ", ".join(f"{k} {v}" for k, v in sorted(m["ops"].items(), key=lambda kv: -kv[1])))


if __name__ == "__main__":
    main()
