#!/usr/bin/env python3
"""Conservation over a full run: the UNMODIFIED reference (oracle/_ref/miluphcuda_<cfg>) next to the reference host
running on the new kernels (oracle/_ref/miluphcuda_<cfg>_b200 = reference main/IO/rk2_adaptive + integration/rhs_b200.cu
+ libb200sph_<cfg>.so), same input, same command line.  north_star: "energy and momentum drift over a full run no
worse than the reference's".  Reads the reference's own conserved_quantities.log (src/io.cu:1980-2017).

    python tools/drift_check.py [--configs shocktube,sedov,impact,giant_hydro] [--out gpurun_out/drift.json]
Needs a GPU (run under gpurun); TEST/MEASUREMENT infrastructure, not part of the product.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from miluphcuda_b200 import scenarios  # noqa: E402

# particles, outputs, output interval, extra flags: a few hundred integrator steps each
RUNS = {
    "shocktube": dict(n=3400, nout=10, tout=0.0228, extra=["-Q", "1e-8"]),
    "sedov": dict(n=240000, nout=10, tout=3e-5, extra=[]),   # the reference aborts (tree out of nodes) at t = 3.6e-4 on this input
    "rings": dict(n=8000, nout=10, tout=4.0, extra=["-Q", "1e-5"]),
    "impact": dict(n=20000, nout=10, tout=2e-4, extra=["-Q", "1e-4"]),
    "giant_hydro": dict(n=20000, nout=5, tout=20.0, extra=["-Q", "1e-4"]),
}


def run_binary(binary: str, sc, wd: str, nout: int, tout: float, extra: list, timeout: int):
    data, cfg = sc.write_inputs(wd)
    cmd = [binary, "-I", "rk2_adaptive", "-f", os.path.basename(data), "-m", os.path.basename(cfg), "-n", str(nout), "-t", repr(tout)] + extra
    if sc.selfgravity:
        cmd += ["-s", "-a", str(sc.theta)]
    t0 = time.time()
    with open(os.path.join(wd, "run.log"), "w") as fh:
        rc = subprocess.call(cmd, cwd=wd, stdout=fh, stderr=subprocess.STDOUT, timeout=timeout)
    wall = time.time() - t0
    log = open(os.path.join(wd, "run.log")).read()
    if rc != 0:
        raise RuntimeError(f"{binary} failed rc={rc}\n{log[-2000:]}")
    path = os.path.join(wd, "conserved_quantities.log")
    rows = [l.split() for l in open(path) if l.strip() and not l.lstrip().startswith("#")]
    table = np.array([[float(v) for v in r] for r in rows])
    steps = log.count("time step accepted") or log.count("accepted")
    last = sorted(f for f in os.listdir(wd) if f.startswith(os.path.basename(data).rsplit(".", 1)[0] + ".") and f[-4:].isdigit())
    final = None
    if last:   # positions and velocities only: later columns are ragged (flaw lists)
        ncol = 2 * sc.dim
        final = np.array([[float(v) for v in line.split()[:ncol]] for line in open(os.path.join(wd, last[-1])) if line.strip()])
    return table, wall, steps, final


def summarise(table: np.ndarray, dim: int, selfgravity: bool) -> dict:
    """columns: time N Nignored Npointmass mass Ekin Einner [Egrav] |p| px [py [pz]] ..."""
    c = 4
    mass, ekin, eint = table[:, c], table[:, c + 1], table[:, c + 2]
    c += 3
    egrav = np.zeros_like(ekin)
    if selfgravity:
        egrav = table[:, c]
        c += 1
    c += 1                      # |p|
    mom = table[:, c: c + dim]
    etot = ekin + eint + egrav
    scale_e = max(np.abs(etot[0]), np.abs(ekin).max(), np.abs(eint).max(), 1e-300)
    # momentum scale: sum m |v| is not logged; use sqrt(2 M Ekin_max) (>= |p|)
    scale_p = max(np.sqrt(2.0 * mass[0] * np.abs(ekin).max()), 1e-300)
    return dict(t_end=float(table[-1, 0]), e_total_0=float(etot[0]), e_total_end=float(etot[-1]),
                energy_drift=float(np.abs(etot - etot[0]).max() / scale_e),
                momentum_drift=float(np.abs(mom - mom[0]).max() / scale_p),
                mass_drift=float(np.abs(mass - mass[0]).max() / mass[0]))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="shocktube,sedov,impact,giant_hydro")
    ap.add_argument("--out", default=os.path.join(REPO, "gpurun_out", "drift.json"))
    ap.add_argument("--timeout", type=int, default=600)
    ap.add_argument("--n", type=int, default=None, help="override the particle count")
    ap.add_argument("--tout", type=float, default=None, help="override the output interval")
    args = ap.parse_args()
    results = {}
    for config in args.configs.split(","):
        spec = dict(RUNS[config])
        if args.n:
            spec["n"] = args.n
        if args.tout:
            spec["tout"] = args.tout
        sc = scenarios.make(config, spec["n"])
        entry = {"particles": sc.n, "args": f"-I rk2_adaptive -n {spec['nout']} -t {spec['tout']} {' '.join(spec['extra'])}"}
        finals = {}
        for label, suffix in (("reference", ""), ("b200", "_b200")):
            binary = os.path.join(REPO, "oracle", "_ref", f"miluphcuda_{config}{suffix}")
            if not os.path.exists(binary):
                entry[label] = {"error": f"{binary} missing (oracle/build_ref.sh {config}{'+b200' if suffix else ''})"}
                continue
            with tempfile.TemporaryDirectory() as wd:
                try:
                    table, wall, steps, final = run_binary(binary, sc, wd, spec["nout"], spec["tout"], spec["extra"], args.timeout)
                    entry[label] = dict(summarise(table, sc.dim, sc.selfgravity), wall_s=round(wall, 2), outputs=len(table))
                    finals[label] = final
                except Exception as exc:  # noqa: BLE001
                    entry[label] = {"error": str(exc)[-1500:]}
        if len(finals) == 2 and finals["reference"] is not None and finals["b200"] is not None and finals["reference"].shape == finals["b200"].shape:
            a, b = finals["reference"], finals["b200"]
            dim = sc.dim
            span = float((a[:, :dim].max(axis=0) - a[:, :dim].min(axis=0)).max())
            entry["final_state"] = {"max_position_difference_over_extent": float(np.abs(a[:, :dim] - b[:, :dim]).max() / span),
                                    "rms_velocity_difference_over_rms_velocity":
                                        float(np.sqrt(np.mean((a[:, dim:2 * dim] - b[:, dim:2 * dim]) ** 2)) /
                                              max(np.sqrt(np.mean(a[:, dim:2 * dim] ** 2)), 1e-300))}
        results[config] = entry
        print(config, json.dumps(entry, indent=1), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump(results, fh, indent=1)


if __name__ == "__main__":
    main()
