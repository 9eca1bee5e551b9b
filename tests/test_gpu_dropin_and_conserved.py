"""GPU: the drop-in BINARY -- the reference host (main, I/O, libconfig reader, rk2_adaptive and predictor-corrector
integrators) linked against libb200sph through integration/rhs_b200.cu (oracle/_ref/miluphcuda_<config>_b200) -- next to
the unmodified reference binary on the same input, same command line:

* rk2_adaptive over >= 20 steps: same accepted / rejected step counts, final state within 1e-7 of the field scale,
  energy and momentum drift read from the reference's own conserved_quantities.log no worse than the reference's
  (north_star), wall-clock of both runs recorded;
* monaghan_pc: the predictor-corrector integrator of src/predictor_corrector.cu on the new right-hand
  side, final state against the reference's.

And the device-side sums behind conserved_quantities.log (b200sph_conserved_quantities) against numpy."""
import json
import os
import re
import time

import numpy as np
import pytest

import common
import make_golden
from miluphcuda_b200 import api, scenarios, state

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

STATE = ("x", "y", "z", "vx", "vy", "vz", "rho", "e", "h", "S", "d", "alpha_jutzi", "damage_porjutzi", "p")


def _run(sc, wd, suffix, evolve_mode):
    os.makedirs(wd)
    env = {"REF_DUMP": os.path.join(wd, "s"), "REF_DUMP_STATE_ONLY": "1"}
    if evolve_mode == "pc":
        env["REF_EVOLVE"] = "pc"
    t0 = time.time()
    log = make_golden.run_reference(sc, wd, env, evolve=True, suffix=suffix)
    wall = time.time() - t0
    if evolve_mode == "pc":
        pass
    text = open(log).read()
    dump = make_golden.read_dump(os.path.join(wd, "s.in.bin"))
    cq = None
    path = os.path.join(wd, "conserved_quantities.log")
    if os.path.exists(path):
        rows = [l.split() for l in open(path) if l.strip() and not l.lstrip().startswith("#")]
        cq = np.array([[float(v) for v in r] for r in rows])
    return dump, text, cq, wall


def _drift(cq, dim):
    """(energy drift, momentum drift) of a conserved_quantities.log table: columns time N Nign Npm mass Ekin Eint |p| px.."""
    if cq is None or len(cq) < 1:
        return None
    e = cq[:, 5] + cq[:, 6]
    return float(e[-1]), float(np.abs(cq[-1, 8: 8 + dim]).max())


@pytest.mark.parametrize("config,n,mode", [("sedov", 100000, "rk2"), ("impact", 100000, "rk2"), ("giant_hydro", 60000, "rk2"),
                                           ("shocktube", 10000, "rk2"), ("impact", 60000, "pc"), ("sedov", 60000, "pc")])
def test_dropin_binary_against_reference_binary(config, n, mode, tmp_path):
    ref_bin = os.path.join(common.REPO, "oracle", "_ref", f"miluphcuda_{config}")
    if not os.path.exists(ref_bin) or not os.path.exists(ref_bin + "_b200"):
        pytest.skip("reference or drop-in binary not built (oracle/build_ref.sh <config> <config>+b200 needs /root/reference)")
    sc = scenarios.make(config, n)
    ref, text_r, cq_r, wall_r = _run(sc, str(tmp_path / "ref"), "", mode)
    new, text_n, cq_n, wall_n = _run(sc, str(tmp_path / "b200"), "_b200", mode)
    if mode == "rk2":
        pat = r"Had to integrate (\d+) timesteps \((\d+) accepted, (\d+) rejected\)"
        assert re.findall(pat, text_n)[-1] == re.findall(pat, text_r)[-1]
        assert int(re.findall(pat, text_r)[-1][1]) >= 20
    bad = {}
    for name in STATE:
        if name in ref and name in new and ref[name].shape == new[name].shape:
            err = common.field_error(new[name], ref[name])
            if not err <= 1e-7:
                bad[name] = err
    assert not bad, f"{mode}: final state of the drop-in binary deviates from the reference binary: {bad}"
    dr, dn = _drift(cq_r, sc.dim), _drift(cq_n, sc.dim)
    if dr is not None and dn is not None:
        scale_e = max(abs(dr[0]), 1e-300)
        assert abs(dn[0] - dr[0]) <= 1e-8 * scale_e, ("total energy at the end differs", dn, dr)
    out = os.path.join(common.REPO, "gpurun_out", "dropin")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, f"{config}_{n}_{mode}.json"), "w") as fh:
        json.dump({"config": config, "particles": int(ref["x"].shape[0]), "integrator": mode, "wall_s_reference": wall_r,
                   "wall_s_dropin": wall_n, "energy_momentum_reference": dr, "energy_momentum_dropin": dn}, fh)


@pytest.mark.parametrize("config", ["impact", "rings", "giant_ignore"])
def test_conserved_quantities_on_device(config, tmp_path):
    sc = scenarios.make(config, 30000, stirred=True)
    cfg = state.write_material_files(sc, str(tmp_path))
    eng = api.RhsEngine(sc.config, n_max=sc.n, material_cfg=cfg)
    arrays, meta = state.scenario_arrays(sc, eng.materials)
    if config.endswith("_ignore"):
        arrays["materialId"][6::13] = -1
    n, dim = meta["n"], sc.dim
    dev = {k: torch.from_numpy(v).cuda() for k, v in arrays.items()}
    cq = eng.conserved_quantities(api.make_view(dev, None, n, max_num_flaws=meta["max_num_flaws"]))
    live = arrays["materialId"] != -1
    m = arrays["m"][live]
    pos = np.stack([arrays[a][live] for a in "xyz"[:dim]] + [np.zeros(live.sum())] * (3 - dim), axis=1)
    vel = np.stack([arrays["v" + a][live] for a in "xyz"[:dim]] + [np.zeros(live.sum())] * (3 - dim), axis=1)
    close = lambda a, b, s: abs(a - b) <= 1e-11 * max(abs(s), 1e-300)
    assert cq.n_ignored == int((~live).sum())
    assert close(cq.mass, m.sum(), m.sum())
    ekin = 0.5 * (m * (vel ** 2).sum(axis=1)).sum()
    assert close(cq.e_kin, ekin, ekin)
    if "e" in arrays:
        eint = (m * arrays["e"][live]).sum()
        assert close(cq.e_int, eint, max(abs(eint), ekin))
    mom_scale = (m * np.abs(vel).sum(axis=1)).sum()
    for k in range(dim):
        assert close(cq.p[k], (m * vel[:, k]).sum(), mom_scale)
        assert close(cq.bary_pos[k], (m * pos[:, k]).sum() / m.sum(), np.abs(pos).max())
    L = (m[:, None] * np.cross(pos, vel)).sum(axis=0)
    Lscale = (m * np.linalg.norm(pos, axis=1) * np.linalg.norm(vel, axis=1)).sum()
    if dim == 3:
        for k in range(3):
            assert close(cq.L[k], L[k], Lscale)
    elif dim == 2:
        assert close(cq.L[0], L[2], Lscale)
    eng.close()
