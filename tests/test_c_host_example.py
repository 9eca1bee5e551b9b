"""integration/c_host_example.c: a host in plain C on the C-ABI (no Python, no torch).

CPU: the example compiles against include/b200sph.h and links against libb200sph_sedov.so (every entry point it
uses resolves).  GPU: it runs -- material.cfg parsed by the library, two evaluations through b200sph_rhs_eval_host --
and its own checks (vanishing total force, sane interaction counts, kernels launched) pass.
"""
import os
import shutil
import subprocess

import pytest

from miluphcuda_b200 import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "integration", "c_host_example.c")


def _compile(out_dir):
    lib = build.lib_path("sedov")
    if not os.path.exists(lib):
        build.build_one("sedov")
    exe = os.path.join(str(out_dir), "c_host_example")
    libdir = os.path.dirname(lib)
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-L", libdir, "-lb200sph_sedov", "-lm",
           f"-Wl,-rpath,{libdir}", "-o", exe]
    cuda_lib = "/usr/local/cuda/lib64"
    if os.path.isdir(cuda_lib):
        cmd += ["-L", cuda_lib, f"-Wl,-rpath,{cuda_lib}"]
    done = subprocess.run(cmd, capture_output=True, text=True)
    assert done.returncode == 0, done.stderr
    return exe


@pytest.mark.skipif(shutil.which("gcc") is None, reason="no gcc")
def test_c_host_example_compiles_and_links(tmp_path):
    exe = _compile(tmp_path)
    assert os.path.getsize(exe) > 0


@pytest.mark.gpu
def test_c_host_example_runs(tmp_path):
    exe = _compile(tmp_path)
    done = subprocess.run([exe, "30", str(tmp_path / "material.cfg")], capture_output=True, text=True, timeout=300)
    assert done.returncode == 0, done.stdout + done.stderr
    assert "C_HOST n=27000" in done.stdout
