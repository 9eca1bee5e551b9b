"""GPU: edge cases of the C-ABI call (SURVEY 8b "Error convention"; the reference asserts / exits where this
library returns a code): a neighbour list that overflows MAX_NUM_INTERACTIONS, a non-finite coordinate, bad
arguments, and the smallest particle sets (one particle without partner, two partners), the latter checked
against the CPU oracle.

Written at the end of round 1 after the GPU budget was spent: the file name makes it run last."""
import numpy as np
import pytest

import common
from miluphcuda_b200 import api, scenarios, state

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _scenario(config, n, tmp_path):
    sc = scenarios.make(config, n, stirred=True)
    cfg = state.write_material_files(sc, str(tmp_path))
    mats = api.MaterialTables(config, cfg)
    arrays, meta = state.scenario_arrays(sc, mats)
    return sc, cfg, mats, arrays, meta


def _head(arrays, n_all, n):
    """the first n particles of a field set"""
    return {k: np.ascontiguousarray(v.reshape(n_all, -1)[:n].reshape(-1)) for k, v in arrays.items()}


def _run(config, cfg, arrays, meta, n, n_max=None):
    eng = api.RhsEngine(config, n_max=n_max or n, material_cfg=cfg)
    dev = {k: torch.from_numpy(v.copy()).cuda() for k, v in arrays.items()}
    view = api.make_view(dev, None, n, max_num_flaws=meta["max_num_flaws"], selfgravity=meta["selfgravity"],
                         theta=meta["theta"], grav_const=eng.materials.grav_const)
    try:
        eng.rhs_eval(view)
        torch.cuda.synchronize()
    finally:
        eng.close()
    return {k: v.cpu().numpy() for k, v in dev.items()}


def test_list_overflow_reports_first_offender(tmp_path):
    """Every particle has more partners than MAX_NUM_INTERACTIONS (128 for the shocktube switch set): the call returns
    TOO_MANY_INTERACTIONS and the lowest offending caller index, like the oracle (the reference asserts, src/tree.cu:917)."""
    sc, cfg, mats, arrays, meta = _scenario("shocktube", 3000, tmp_path)
    n = 400
    sub = _head(arrays, sc.n, n)
    sub["x"] = np.linspace(0.0, 0.5, n)
    sub["h"][:] = 1.0
    sub["h0"][:] = 1.0
    meta_n = dict(meta, n=n)
    ref = {k: v.copy() for k, v in sub.items()}
    rc, off, _ = common.oracle_rhs("shocktube", ref, mats, meta_n)
    assert rc == 1 and off == 0
    with pytest.raises(api.B200SphError) as err:
        _run("shocktube", cfg, sub, meta_n, n)
    assert err.value.code == 1 and err.value.offender == 0


def test_nonfinite_coordinate_is_reported(tmp_path):
    sc, cfg, mats, arrays, meta = _scenario("sedov", 3000, tmp_path)
    arrays["y"][7] = np.nan
    with pytest.raises(api.B200SphError) as err:
        _run("sedov", cfg, arrays, meta, sc.n)
    assert err.value.code == 5


def test_bad_arguments(tmp_path):
    sc, cfg, mats, arrays, meta = _scenario("sedov", 3000, tmp_path)
    with pytest.raises(api.B200SphError) as err:          # more particles than the handle was created for
        _run("sedov", cfg, arrays, meta, sc.n, n_max=sc.n - 1)
    assert err.value.code == 2
    eng = api.RhsEngine("sedov", n_max=sc.n, material_cfg=cfg)
    dev = {k: torch.from_numpy(v.copy()).cuda() for k, v in arrays.items()}
    view = api.make_view(dev, None, sc.n, max_num_flaws=meta["max_num_flaws"], selfgravity=meta["selfgravity"],
                         theta=meta["theta"], grav_const=eng.materials.grav_const)
    view.p.cs = None                                       # a mandatory member is missing
    with pytest.raises(api.B200SphError) as err:
        eng.rhs_eval(view)
    assert err.value.code == 2
    eng.close()


@pytest.mark.parametrize("n", [1, 2])
def test_smallest_particle_sets_match_oracle(n, tmp_path):
    sc, cfg, mats, arrays, meta = _scenario("sedov", 3000, tmp_path)
    sub = _head(arrays, sc.n, n)
    if n == 2:                                             # make the two particles partners
        for a in ("x", "y", "z"):
            sub[a][1] = sub[a][0]
        sub["x"][1] += 0.3 * sub["h"][0]
    meta_n = dict(meta, n=n)
    ref = {k: v.copy() for k, v in sub.items()}
    rc, off, _ = common.oracle_rhs("sedov", ref, mats, meta_n)
    assert rc == 0
    out = _run("sedov", cfg, sub, meta_n, n)
    assert np.array_equal(out["noi"], ref["noi"]) and int(ref["noi"][0]) == n - 1
    for name in ("ax", "ay", "az", "drhodt", "dedt", "rho", "p", "cs", "dxdt", "dydt", "dzdt"):
        err = common.field_error(out[name], ref[name])
        assert err <= common.RTOL, (name, err, out[name], ref[name])
