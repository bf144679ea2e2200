#!/usr/bin/env python
"""Builds timing variants of libfishing_b200.so into gym_fishing_b200/ab/ (git-ignored, shipped to the GPU box):
    python tools/ab_build.py [name ...]
Run one with FISHING_B200_LIB=gym_fishing_b200/ab/lib_<name>.so python tools/k2_bench.py.  These are measurement aids (what
does the Philox block cost, what does the round-1 Box-Muller cost); only the default build is the product."""
from __future__ import annotations

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gym_fishing_b200.build import build  # noqa: E402

VARIANTS = {
    "normal_v1": ["-DFISHING_NORMAL_V1"],                 # round-1 Box-Muller (58 FP64 instructions per pair of draws)
    "rounds7": ["-DFISHING_PHILOX_ROUNDS=7"],             # NOT the product: the cost of three Philox rounds
    "rounds0": ["-DFISHING_PHILOX_ROUNDS=0"],             # NOT the product: counter words straight into the Box-Muller
    "split_mul": ["-DFISHING_AB_SPLIT_MUL"],              # mul.hi + mul.lo kept apart instead of IMAD.WIDE
    "unroll1": ["-DFISHING_K2_UNROLL=1"],                 # the fused loop not unrolled (default: 2)
    "mb6": ["-DFISHING_K2_MINBLOCKS=6"],                  # fused kernel capped at 80 registers (6 blocks of 128 per SM)
    "mb8": ["-DFISHING_K2_MINBLOCKS=8"],                  # ... 64 registers
    "mb10": ["-DFISHING_K2_MINBLOCKS=10"],                # ... 48 registers
    "unroll4": ["-DFISHING_K2_UNROLL=4"],
    "libexp": ["-DFISHING_AB_LIBEXP"],
    "v4bs256": ["-DFISHING_K1_BS_V4=256"], "v4bs64": ["-DFISHING_K1_BS_V4=64"],   # fishing-v4's K1 block size (default 128: shorter reset tails)
    "v4ilp2": ["-DFISHING_AB_V4_ILP2"], "v4ilp2bs256": ["-DFISHING_AB_V4_ILP2", "-DFISHING_K1_BS_V4=256"],
    "v4rcp0": ["-DFISHING_V4_RCP=0"],                      # fishing-v4's two quotients per step as hardware divisions (until r02n)
    "v4res1536": ["-DFISHING_V4_RESIDENT=1536"],           # the lean fishing-v4 K1 compiled for 1536 resident threads (40 registers)
    "k2bs64": ["-DFISHING_K2_FAST_BS=64"], "k2bs64mb9": ["-DFISHING_K2_FAST_BS=64", "-DFISHING_K2_MINBLOCKS=9"],   # fused kernel in 64-thread blocks (8 blocks = today's 16 warps per SM; 9 = 18 warps at 112 registers)
    "k2tip5": ["-DFISHING_K2_TIPPING_WARPS5=1"],           # only the tipping-point instantiations at 64 threads x 9 blocks
    "k2bs32mb18": ["-DFISHING_K2_FAST_BS=32", "-DFISHING_K2_MINBLOCKS=18"], "k2bs32mb17": ["-DFISHING_K2_FAST_BS=32", "-DFISHING_K2_MINBLOCKS=17"],
    "liblog": ["-DFISHING_AB_LIBLOG"],                     # fishing-v5..v11 with libdevice's log / pow instead of csrc/normal.cuh:log_tab
    "imad_select": ["-DFISHING_AB_IMAD_SELECT"],
    "sweep_unroll1": ["-DFISHING_SWEEP_UNROLL=1"], "sweep_unroll2": ["-DFISHING_SWEEP_UNROLL=2"], "sweep_unroll8": ["-DFISHING_SWEEP_UNROLL=8"],
    "k1bs128": ["-DFISHING_K1_BS=128"], "k1bs512": ["-DFISHING_K1_BS=512"],   # block size of the lock-step kernel
    "one_from_constant": ["-DFISHING_AB_ONE_FROM_CONSTANT"],   # 1.0's high word from the constant bank (one LOP3 less per use)
    "v4nodraw": ["-DFISHING_AB_V4_NODRAW"],               # NOT the product: fishing-v4 resets without the Philox / Box-Muller work
    "v4nowrite": ["-DFISHING_AB_V4_NOWRITE"],             # NOT the product: ... without the scattered K / r rewrites                    # fishing-v2 with libdevice's exp instead of csrc/normal.cuh:exp_tab
    "k1smem": ["-DFISHING_K1_SMEM_TABLES"],               # K1 stages the Box-Muller tables in shared memory per block
}


def main():
    names = sys.argv[1:] or list(VARIANTS)
    outdir = os.path.join(ROOT, "gym_fishing_b200", "ab")
    os.makedirs(outdir, exist_ok=True)
    for n in names:
        out = os.path.join(outdir, f"lib_{n}.so")
        build(force=True, extra_flags=VARIANTS[n], out=out)
        print("built", out)


if __name__ == "__main__":
    main()
