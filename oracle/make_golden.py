#!/usr/bin/env python3
"""TEST INFRASTRUCTURE -- produce golden vectors from the reference itself.

Runs on the GPU box (the reference has no CPU path): for each config it writes
a small synthetic input with `miluphcuda_b200.scenarios`, runs the UNMODIFIED
reference build `oracle/_ref/miluphcuda_<config>` (built by oracle/build_ref.sh
from /root/reference) with the dump hook `oracle/ref_hook.cu`, and stores

    <out>/<config>.npz   in_*   : state before the first rightHandSide()
                         out1_* : everything after the first call (p pinned to 0 before it)
                         out2_* : everything after a second call on the same buffers
                         nbr_ptr / nbr_idx : neighbour sets (CSR, sorted) of call 1

The files are committed under tests/golden/ and are the pin for oracle/ and
for the CUDA path.  Nothing here reads /root/reference at run time.

usage: python oracle/make_golden.py [--out gpurun_out/golden] [--configs a,b] [--n 2500]
       python oracle/make_golden.py --time sedov:1000000 [--calls 10]
"""
from __future__ import annotations

import argparse
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from miluphcuda_b200 import scenarios  # noqa: E402

GOLDEN_N = {"shocktube": None, "sedov": 2500, "rings": 2500, "impact": 2500, "giant_hydro": 2500, "giant_solid": 2500,
            "giant_aneos": 2500}   # giant_aneos: the giant_hydro build with tabulated-EOS materials


def read_dump(path: str) -> dict:
    out = {}
    with open(path, "rb") as fh:
        while True:
            tag = fh.read(32)
            if len(tag) < 32:
                break
            name = tag.split(b"\0", 1)[0].decode()
            dtype = int(np.frombuffer(fh.read(4), dtype=np.int32)[0])
            count = int(np.frombuffer(fh.read(8), dtype=np.int64)[0])
            dt = np.float64 if dtype == 0 else np.int32
            out[name] = np.frombuffer(fh.read(count * np.dtype(dt).itemsize), dtype=dt).copy()
    return out


def ref_binary(config: str) -> str:
    path = os.path.join(REPO, "oracle", "_ref", f"miluphcuda_{config}")
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} missing: run oracle/build_ref.sh where /root/reference exists")
    return path


def run_reference(sc: scenarios.Scenario, workdir: str, env_extra: dict, log_name: str = "ref.log") -> str:
    data, cfg = sc.write_inputs(workdir)
    cmd = [ref_binary(sc.config), "-I", "euler", "-f", os.path.basename(data), "-m", os.path.basename(cfg),
           "-n", "1", "-t", "1e-6"]
    if sc.selfgravity:
        cmd += ["-s", "-a", str(sc.theta)]
    env = dict(os.environ)
    env.update(env_extra)
    log = os.path.join(workdir, log_name)
    with open(log, "w") as fh:
        rc = subprocess.call(cmd, cwd=workdir, stdout=fh, stderr=subprocess.STDOUT, env=env)
    if rc != 0:
        tail = open(log).read()[-3000:]
        raise RuntimeError(f"reference run failed rc={rc}\n{tail}")
    return log


def make_golden(config: str, n, out_dir: str, stirred: bool = False) -> str:
    sc = scenarios.make(config, n, stirred=stirred)
    with tempfile.TemporaryDirectory() as wd:
        run_reference(sc, wd, {"REF_DUMP": os.path.join(wd, "dump"), "REF_DUMP_LISTS": "1"})
        d_in = read_dump(os.path.join(wd, "dump.in.bin"))
        d1 = read_dump(os.path.join(wd, "dump.out1.bin"))
        d2 = read_dump(os.path.join(wd, "dump.out2.bin"))
    npart = sc.n
    maxni = int(d1.pop("max_num_interactions")[0])
    inter = d1.pop("interactions").reshape(npart, maxni)
    noi = d1["noi"]
    ptr = np.zeros(npart + 1, dtype=np.int64)
    ptr[1:] = np.cumsum(noi)
    idx = np.empty(int(ptr[-1]), dtype=np.int32)
    for i in range(npart):
        idx[ptr[i]:ptr[i + 1]] = np.sort(inter[i, : noi[i]])
    payload = {"nbr_ptr": ptr, "nbr_idx": idx, "n": np.int64(npart), "max_num_interactions": np.int64(maxni),
               "selfgravity": np.int64(1 if sc.selfgravity else 0), "theta": np.float64(sc.theta)}
    for k, v in d_in.items():
        payload["in_" + k] = v
    for k, v in d1.items():
        payload["out1_" + k] = v
    for k, v in d2.items():
        if k in d1 and np.array_equal(v, d1[k]):
            continue  # unchanged by the second call
        payload["out2_" + k] = v
    for k in list(d1):
        if k in d_in and np.array_equal(d1[k], d_in[k]) and k != "noi":
            payload.pop("out1_" + k)  # input passed through untouched
    payload["material_cfg"] = np.array(sc.material_cfg)
    for name, text in sc.includes.items():
        payload["include_" + name] = np.array(text)
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, f"{config}_stirred.npz" if stirred else f"{config}.npz")
    np.savez_compressed(path, **payload)
    return path


def time_reference(config: str, n: int, calls: int, warmup: int, keep_log: str | None = None) -> dict:
    sc = scenarios.make(config, n)
    with tempfile.TemporaryDirectory() as wd:
        t0 = time.time()
        log = run_reference(sc, wd, {"REF_TIMED": str(calls), "REF_WARMUP": str(warmup)})
        text = open(log).read()
        wall = time.time() - t0
        if keep_log:
            with open(keep_log, "w") as fh:
                fh.write(text[-200000:])
    m = re.search(r"REF_TIMING n=(\d+) calls=(\d+) warmup=(\d+) ms_per_call=([\d.eE+-]+) best_ms=([\d.eE+-]+) updates_per_s=([\d.eE+-]+)", text)
    if not m:
        raise RuntimeError("no REF_TIMING line in reference output:\n" + text[-2000:])
    kern = {}
    for km in re.finditer(r"duration ([^:]+): ([\d.]+) ms", text):
        kern.setdefault(km.group(1).strip(), []).append(float(km.group(2)))
    return {"config": config, "n": int(m.group(1)), "calls": int(m.group(2)), "ms_per_call": float(m.group(4)),
            "best_ms": float(m.group(5)), "updates_per_s": float(m.group(6)), "wall_s": wall,
            "kernel_ms_last": {k: v[-1] for k, v in kern.items()}}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(REPO, "gpurun_out", "golden"))
    ap.add_argument("--configs", default=",".join(GOLDEN_N))
    ap.add_argument("--n", type=int, default=None)
    ap.add_argument("--n-stirred", type=int, default=1500)
    ap.add_argument("--time", default=None, help="config:n[,config:n...] -> time the reference RHS")
    ap.add_argument("--calls", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    if args.time:
        import json
        for item in args.time.split(","):
            cfg, n = item.split(":")
            os.makedirs(args.out, exist_ok=True)
            res = time_reference(cfg, int(n), args.calls, args.warmup, keep_log=os.path.join(args.out, f"ref_{cfg}_{n}.log"))
            print("REF_RESULT " + json.dumps(res), flush=True)
        return
    for cfg in args.configs.split(","):
        n = args.n if args.n is not None else GOLDEN_N[cfg]
        t0 = time.time()
        for stirred in (False, True):
            if stirred and n is not None:
                n_use = args.n_stirred
            else:
                n_use = n
            path = make_golden(cfg, n_use, args.out, stirred=stirred)
            print(f"golden {cfg}: {path} ({os.path.getsize(path) / 1e6:.2f} MB, {time.time() - t0:.1f}s)", flush=True)


if __name__ == "__main__":
    main()
