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
            "giant_aneos": 2500,   # giant_aneos: the giant_hydro build with tabulated-EOS materials
            "nakamura": 2500}
# variants that exist in one flavour only (already stirred): deactivated particles (both the eos.type = IGNORE material and
# materialId = -1 set by the hook), the other crush curves of the p-alpha model
GOLDEN_SINGLE = {"sedov_ignore": 2500, "impact_ignore": 2500, "giant_ignore": 2500,
                 "impact_aneos": 1500, "impact_crush1": 1500, "impact_crush2": 1500, "impact_crush3": 1500, "impact_crush4": 1500}
DEACTIVATE_STRIDE = 13

# Evolved states ("a state after >= 20 accepted steps"): the reference's own rk2_adaptive integrates the
# scenario over `steps` x dt_max with -M dt_max, so at least `steps` steps are taken whatever the natural step is.
# dt_max(n) = c * h: a fraction of the signal-crossing time of one smoothing length.
EVOLVE_STEPS = 24


def evolve_args(sc, steps: int = EVOLVE_STEPS) -> list:
    """Command-line arguments that make the reference's rk2_adaptive take >= `steps` steps: -n 1 -t steps*dt -M dt -A."""
    import re as _re
    if sc.h is not None:
        h = float(np.min(sc.h))
    else:
        h = float(_re.search(r"sml\s*=\s*([0-9.eE+-]+)", sc.material_cfg).group(1))
    # signal speeds of the scenarios (sound speed of the material / of the hot blast centre / impact speed)
    speed = {"shocktube": 2.0, "sedov": 0.0, "rings": 1.0, "impact": 9.0e3, "giant_hydro": 8.0e3, "giant_solid": 8.0e3,
             "nakamura": 7.0e3}[sc.config]
    if sc.config == "sedov":
        speed = float(np.sqrt(1.4 * 0.4 * np.max(sc.e)))
    # sedov: the reference runs out of tree nodes once the blast has piled particles up (t ~ 3 h / c_s), stay well before
    dt = (0.04 if sc.config == "sedov" else 0.25) * h / speed
    eps = {"shocktube": "1e-8", "rings": "1e-5"}.get(sc.config, "1e-4")
    return ["-n", "1", "-t", repr(steps * dt), "-M", repr(dt), "-Q", eps, "-A"]


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


class ReferenceTimeout(RuntimeError):
    """The reference binary did not finish within the caller's limit (it is killed)."""


def run_reference(sc: scenarios.Scenario, workdir: str, env_extra: dict, log_name: str = "ref.log", evolve: bool = False,
                  decouple: bool = False, input_file: str | None = None, suffix: str = "", timeout_s: float | None = None) -> str:
    """Run oracle/_ref/miluphcuda_<config><suffix> on the scenario (or on an existing input file in `workdir`) with
    `-I euler`, i.e. through oracle/ref_hook.cu; evolve=True lets the reference's own rk2_adaptive run first."""
    if input_file is None:
        data, cfg = sc.write_inputs(workdir)
    else:
        data, cfg = input_file, os.path.join(workdir, "material.cfg")
    time_args = evolve_args(sc) if evolve else ["-n", "1", "-t", "1e-6"]
    cmd = [ref_binary(sc.config) + suffix, "-I", "euler", "-f", os.path.basename(data), "-m", os.path.basename(cfg)] + time_args
    if sc.selfgravity:
        cmd += ["-s", "-a", str(sc.theta)]
    if decouple:
        cmd += ["-g"]
    if evolve:
        env_extra = dict({"REF_EVOLVE": "1"}, **env_extra)   # "pc" from the caller: the predictor-corrector integrator
    env = dict(os.environ)
    env.update(env_extra)
    log = os.path.join(workdir, log_name)
    with open(log, "w") as fh:
        try:
            rc = subprocess.call(cmd, cwd=workdir, stdout=fh, stderr=subprocess.STDOUT, env=env, timeout=timeout_s)
        except subprocess.TimeoutExpired:
            raise ReferenceTimeout(f"{os.path.basename(cmd[0])} did not finish within {timeout_s:.0f} s ({' '.join(time_args)})") from None
    if rc != 0:
        tail = open(log).read()[-3000:]
        raise RuntimeError(f"reference run failed rc={rc}\n{tail}")
    return log


def arrays_from_dump(config: str, d_in: dict, selfgravity: bool):
    """({field: numpy array} for every member of the switch set, meta) from a state dump of the hook."""
    from miluphcuda_b200 import api
    sw = scenarios.read_switches(config)
    n = int(d_in["x"].shape[0])
    dim, max_flaws = sw["DIM"], sw.get("MAX_NUM_FLAWS", 1)
    p_fields, rhs_fields = api.fields_for(sw, selfgravity)
    arrays = {}
    for name in p_fields + rhs_fields:
        dtype = np.int32 if name in api.INT_FIELDS else np.float64
        shape = api.field_shape(name, n, dim, max_flaws)
        if name in d_in and d_in[name].shape == shape:
            arrays[name] = np.ascontiguousarray(d_in[name].astype(dtype))
        else:
            arrays[name] = np.zeros(shape, dtype=dtype)
    return arrays, dict(n=n, max_num_flaws=max_flaws, selfgravity=selfgravity)


def evolved_state(sc, workdir: str, timeout_s: float | None = None):
    """State of the scenario after the reference's own rk2_adaptive took >= EVOLVE_STEPS steps: (dump, accepted steps)."""
    log = run_reference(sc, workdir, {"REF_DUMP": os.path.join(workdir, "evolved"), "REF_DUMP_STATE_ONLY": "1"}, evolve=True,
                        timeout_s=timeout_s)
    text = open(log).read()
    acc = re.findall(r"Had to integrate (\d+) timesteps \((\d+) accepted, (\d+) rejected\)", text)
    d_in = read_dump(os.path.join(workdir, "evolved.in.bin"))
    os.remove(os.path.join(workdir, "evolved.in.bin"))
    return d_in, (int(acc[-1][1]) if acc else 0)


def make_golden(config: str, n, out_dir: str, stirred: bool = False) -> str:
    sc = scenarios.make(config, n, stirred=stirred)
    env = {"REF_DUMP": "dump", "REF_DUMP_LISTS": "1"}
    if config.endswith("_ignore"):
        env["REF_DEACTIVATE"] = str(DEACTIVATE_STRIDE)
    with tempfile.TemporaryDirectory() as wd:
        env["REF_DUMP"] = os.path.join(wd, "dump")
        run_reference(sc, wd, env)
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


def time_reference(config: str, n: int, calls: int, warmup: int, keep_log: str | None = None, evolve: bool = False,
                   timeout_s: float | None = None) -> dict:
    sc = scenarios.make(config, n)
    with tempfile.TemporaryDirectory() as wd:
        t0 = time.time()
        log = run_reference(sc, wd, {"REF_TIMED": str(calls), "REF_WARMUP": str(warmup)}, evolve=evolve, timeout_s=timeout_s)
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
    acc = re.findall(r"Had to integrate (\d+) timesteps \((\d+) accepted, (\d+) rejected\)", text)
    return {"config": config, "n": int(m.group(1)), "calls": int(m.group(2)), "ms_per_call": float(m.group(4)),
            "evolved_steps": ({"integrated": int(acc[-1][0]), "accepted": int(acc[-1][1]), "rejected": int(acc[-1][2])} if acc else None),
            "best_ms": float(m.group(5)), "updates_per_s": float(m.group(6)), "wall_s": wall,
            "kernel_ms_last": {k: v[-1] for k, v in kern.items()}}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(REPO, "gpurun_out", "golden"))
    ap.add_argument("--configs", default=",".join(list(GOLDEN_N) + list(GOLDEN_SINGLE)))
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
        t0 = time.time()
        if cfg in GOLDEN_SINGLE:
            path = make_golden(cfg, args.n if args.n is not None else GOLDEN_SINGLE[cfg], args.out)
            print(f"golden {cfg}: {path} ({os.path.getsize(path) / 1e6:.2f} MB, {time.time() - t0:.1f}s)", flush=True)
            continue
        n = args.n if args.n is not None else GOLDEN_N[cfg]
        for stirred in (False, True):
            if stirred and n is not None:
                n_use = args.n_stirred
            else:
                n_use = n
            path = make_golden(cfg, n_use, args.out, stirred=stirred)
            print(f"golden {cfg}: {path} ({os.path.getsize(path) / 1e6:.2f} MB, {time.time() - t0:.1f}s)", flush=True)


if __name__ == "__main__":
    main()
