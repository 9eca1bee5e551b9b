"""Build libb200sph_<config>.so (sm_100a) in-tree, one library per compile-time switch set.

The reference is configured by parameter.h + recompile (README "Usage"); so is this
library: `nvcc -I configs/<config>` picks the switch set.  Output goes to
miluphcuda_b200/lib/ (git-ignored, travels to the GPU box with the snapshot).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# B200SPH_LIBDIR: measurement sessions build compile-time variants into their own directory and pick them per run
LIBDIR = os.environ.get("B200SPH_LIBDIR") or os.path.join(HERE, "lib")
CONFIGS = ("shocktube", "sedov", "rings", "impact", "giant_hydro", "giant_solid", "nakamura")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CU_SOURCES = ("libb200sph.cu",)
C_SOURCES = ("materials.c", "libconfig_lite.c")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def lib_path(config: str) -> str:
    return os.path.join(LIBDIR, f"libb200sph_{config}.so")


def _deps(config: str) -> list:
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "configs", config, "parameter.h"))
    deps.append(os.path.join(os.path.dirname(HERE), "include", "b200sph.h"))
    return deps


def build_one(config: str, force: bool = False, verbose: bool = False, config_dir: str | None = None) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    lib = lib_path(config)
    if not force and os.path.exists(lib) and all(os.path.getmtime(lib) >= os.path.getmtime(d) for d in _deps(config)):
        return lib
    cdir = config_dir or os.path.join(HERE, "configs", config)
    objdir = os.path.join(LIBDIR, "obj", config)
    os.makedirs(objdir, exist_ok=True)
    common = ["-I", cdir, "-I", CSRC, f'-DB200SPH_CONFIG_NAME="{config}"', *os.environ.get("B200SPH_EXTRA_FLAGS", "").split()]
    objs = []
    for src in CU_SOURCES:
        obj = os.path.join(objdir, src + ".o")
        cmd = [NVCC, *ARCH, "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-c", *common,
               os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.check_call(cmd)
        objs.append(obj)
    for src in C_SOURCES:
        obj = os.path.join(objdir, src + ".o")
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-std=gnu11", "-c", *common, os.path.join(CSRC, src), "-o", obj])
        objs.append(obj)
    subprocess.check_call([NVCC, *ARCH, "-shared", "-o", lib, *objs, "-lcudart"])
    return lib


def build_all(configs=CONFIGS, force: bool = False, verbose: bool = False) -> list:
    with ThreadPoolExecutor(max_workers=min(len(configs), os.cpu_count() or 2)) as ex:
        return list(ex.map(lambda c: build_one(c, force=force, verbose=verbose), configs))


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("-")]
    for path in build_all(args or CONFIGS, force="--force" in sys.argv, verbose="-v" in sys.argv):
        print(path)
