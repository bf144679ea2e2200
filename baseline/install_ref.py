#!/usr/bin/env python
"""Puts the UNMODIFIED reference package where `bench.py --impl reference` times it: baseline/_ref/gym_fishing.

    python baseline/install_ref.py            # also called by __graft_entry__.build()

The reference (boettiger-lab/gym_fishing) is pure Python, so "installing" it is a copy of its package directory from
/root/reference (present in the authoring container only).  baseline/_ref/ is git-ignored -- reference sources never enter
this repository's history -- but NOT gpurun-ignored, so the copy travels to the GPU box with the working tree.  The
reference imports `gym` and `matplotlib`, which are not in the image; bench.py puts the small stand-ins of oracle/refstub/
on sys.path for the reference arm (the same ones tests/golden/make_golden.py uses).  Nothing is patched.

`pip install --no-index --target baseline/_ref /root/reference` was tried first (2026-10-18): setup.py's
install_requires (gym, numpy, pandas, matplotlib) cannot be resolved offline and with --no-deps the build wants to write
gym_fishing.egg-info into the read-only source tree; the result of a working pip install would be this same directory."""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/gym_fishing"
DST = os.path.join(HERE, "_ref", "gym_fishing")


def install(quiet=False):
    """Returns the path of the installed package, or None when the reference is not on this machine."""
    if not os.path.isdir(SRC):
        return DST if os.path.isdir(DST) else None
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(os.path.dirname(DST), exist_ok=True)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    if not quiet:
        print("installed the unmodified reference at", DST)
    return DST


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
