#!/usr/bin/env python3
"""Wall clock of the time integration itself at full size: the reference's own
rk2Adaptive() (oracle/_ref/miluphcuda_<config>, timed around the call by oracle/ref_hook.cu) against b200sph_rk2_advance
(fused integrator kernels + this library's right-hand side) for the SAME run: same input, same end time, same step-size
control, so the same accepted / rejected steps.  MEASUREMENT infrastructure (GPU only).

usage: python tools/integrator_speed.py [workload:particles ...]  -> one JSON line per workload, also gpurun_out/integrator/*.json
"""
import json, os, re, sys, tempfile, time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "oracle"))
sys.path.insert(0, os.path.join(REPO, "tests"))
import numpy as np
import torch
import make_golden
import common
from miluphcuda_b200 import api, scenarios

INTEGRATED = ("x", "y", "z", "vx", "vy", "vz", "rho", "e", "h", "S", "d", "alpha_jutzi")


def measure(config: str, n: int) -> dict:
    sc = scenarios.make(config, n)
    with tempfile.TemporaryDirectory() as wd0, tempfile.TemporaryDirectory() as wd1:
        make_golden.run_reference(sc, wd0, {"REF_DUMP": os.path.join(wd0, "s"), "REF_DUMP_STATE_ONLY": "1"})
        start = make_golden.read_dump(os.path.join(wd0, "s.in.bin"))
        t0 = time.time()
        log = make_golden.run_reference(sc, wd1, {"REF_DUMP": os.path.join(wd1, "s"), "REF_DUMP_STATE_ONLY": "1"}, evolve=True,
                                        timeout_s=900)
        ref_process_s = time.time() - t0
        ref = make_golden.read_dump(os.path.join(wd1, "s.in.bin"))
        text = open(log).read()
        acc = re.findall(r"Had to integrate (\d+) timesteps \((\d+) accepted, (\d+) rejected\)", text)[-1]
        ref_ms = float(re.search(r"REF_EVOLVE_WALL_MS=([\d.]+)", text).group(1))
        # the drop-in binary: the reference's host and integrator (its ~120 device-to-device copies per step included) on
        # this library's right-hand side (integration/rhs_b200.cu)
        dropin_ms, dropin_acc, dropin_err = None, None, None
        if os.path.exists(make_golden.ref_binary(config) + "_b200"):
            with tempfile.TemporaryDirectory() as wd2:
                log2 = make_golden.run_reference(sc, wd2, {"REF_DUMP": os.path.join(wd2, "s"), "REF_DUMP_STATE_ONLY": "1"}, evolve=True,
                                                 suffix="_b200", timeout_s=900)
                text2 = open(log2).read()
                dropin_ms = float(re.search(r"REF_EVOLVE_WALL_MS=([\d.]+)", text2).group(1))
                a2 = re.findall(r"Had to integrate (\d+) timesteps \((\d+) accepted, (\d+) rejected\)", text2)[-1]
                dropin_acc = {"accepted": int(a2[1]), "rejected": int(a2[2])}
                d2 = make_golden.read_dump(os.path.join(wd2, "s.in.bin"))
                dropin_err = max(float(common.field_error(d2[k], ref[k])) for k in INTEGRATED if k in d2 and k in ref and d2[k].shape == ref[k].shape)
        args = make_golden.evolve_args(sc)
        t_end, dt_max, eps = float(args[args.index("-t") + 1]), float(args[args.index("-M") + 1]), float(args[args.index("-Q") + 1])
        arrays, meta = make_golden.arrays_from_dump(config, start, bool(sc.selfgravity))
        n = meta["n"]
        eng = api.RhsEngine(config, n_max=n, material_cfg=os.path.join(wd0, "material.cfg"))
    dev = {k: torch.from_numpy(v).cuda() for k, v in arrays.items()}
    rk_fields = [{k: torch.zeros_like(v) for k, v in dev.items() if k not in ("materialId", "flaws", "h0")} for _ in range(3)]
    view = api.make_view(dev, None, n, max_num_flaws=meta["max_num_flaws"], selfgravity=meta["selfgravity"], theta=sc.theta,
                         grav_const=eng.materials.grav_const)
    rk = eng.rk2_buffers(rk_fields)
    prm = eng.rk2_default_params()
    prm.rk_epsrel, prm.dt_max = eps, dt_max
    st = api.Rk2State()
    torch.cuda.synchronize()
    t0 = time.time()
    eng.rk2_init(view, rk)
    eng.rk2_advance(view, rk, prm, t_end, st)
    torch.cuda.synchronize()
    ours_ms = (time.time() - t0) * 1e3
    eng.pressure(view)
    torch.cuda.synchronize()
    worst = {}
    for name in INTEGRATED:
        if name in dev and name in ref and ref[name].shape == tuple(dev[name].shape):
            worst[name] = float(common.field_error(dev[name].cpu().numpy(), ref[name]))
    out = {"workload": config, "particles": n, "t_end": t_end, "steps_reference": {"accepted": int(acc[1]), "rejected": int(acc[2])},
           "steps_b200": {"accepted": st.accepted, "rejected": st.rejected, "rhs_calls": st.rhs_calls},
           "reference_rk2Adaptive_wall_ms": ref_ms, "reference_process_wall_s": ref_process_s,
           "b200_rk2_advance_wall_ms": ours_ms, "ratio": ref_ms / ours_ms,
           "dropin_binary_rk2Adaptive_wall_ms": dropin_ms, "dropin_ratio": (ref_ms / dropin_ms) if dropin_ms else None,
           "dropin_steps": dropin_acc, "dropin_max_field_error": dropin_err,
           "ms_per_accepted_step": {"reference": ref_ms / max(int(acc[1]), 1), "b200": ours_ms / max(st.accepted, 1)},
           "max_field_error_after_run": max(worst.values()) if worst else None, "field_errors": worst,
           "what": ("reference: rk2Adaptive() of the unmodified CUDA build (device-resident, no particle output), host clock around the call; "
                    "b200: b200sph_rk2_init + b200sph_rk2_advance on device-resident buffers in the input file's particle order, host clock")}
    eng.close()
    return out


def main() -> None:
    items = sys.argv[1:] or ["impact:1000000"]
    os.makedirs(os.path.join(REPO, "gpurun_out", "integrator"), exist_ok=True)
    for item in items:
        cfg, n = item.split(":")
        res = measure(cfg, int(n))
        with open(os.path.join(REPO, "gpurun_out", "integrator", f"{cfg}_{n}.json"), "w") as fh:
            json.dump(res, fh, indent=1)
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
