#!/usr/bin/env python3
"""TEST INFRASTRUCTURE -- compile oracle/sph_oracle.c once per switch set.

Output: oracle/_build/liboracle_<config>.so (git-ignored; travels to the GPU box).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
CONFIGS = ("shocktube", "sedov", "rings", "impact", "giant_hydro", "giant_solid", "nakamura")


def lib_path(config: str) -> str:
    return os.path.join(HERE, "_build", f"liboracle_{config}.so")


def build(configs=CONFIGS, force: bool = False) -> list:
    os.makedirs(os.path.join(HERE, "_build"), exist_ok=True)
    src = os.path.join(HERE, "sph_oracle.c")
    out = []
    for cfg in configs:
        lib = lib_path(cfg)
        deps = [src, os.path.join(REPO, "include", "b200sph.h"),
                os.path.join(REPO, "miluphcuda_b200", "configs", cfg, "parameter.h")]
        if not force and os.path.exists(lib) and all(os.path.getmtime(lib) >= os.path.getmtime(d) for d in deps):
            out.append(lib)
            continue
        cmd = ["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared", "-std=gnu11", "-Wall",
               "-Wno-unused-function", "-Wno-unused-variable", "-Wno-unused-but-set-variable",
               "-I", os.path.join(REPO, "miluphcuda_b200", "configs", cfg), "-I", os.path.join(REPO, "include"),
               src, "-o", lib, "-lm"]
        subprocess.check_call(cmd)
        out.append(lib)
    return out


if __name__ == "__main__":
    for p in build(sys.argv[1:] or CONFIGS, force=True):
        print(p)
