"""GPU: b200sph_rhs_eval_host (HOST buffers, overlapped copies) against b200sph_rhs_eval (device buffers).

Both run the same kernels, so every member the host call reads back must be BIT-identical to the device
call's -- with the default options, and with cached immutables + skipped p_rhs scratch (where the scratch
members must stay untouched on the host and everything else must still be identical, call after call)."""
import numpy as np
import pytest

import common
from miluphcuda_b200 import api, scenarios, state

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

SCRATCH = ("sigma", "R", "plastic_f", "tensorialCorrectionMatrix")
IMMUTABLE = ("m", "h0", "materialId", "numFlaws", "flaws")


def device_calls(config, arrays, cfg, meta, calls):
    eng = api.RhsEngine(config, n_max=meta["n"], material_cfg=cfg)
    dev = {k: torch.from_numpy(v.copy()).cuda() for k, v in arrays.items()}
    view = api.make_view(dev, None, meta["n"], max_num_flaws=meta["max_num_flaws"], selfgravity=meta["selfgravity"],
                         theta=meta["theta"], grav_const=eng.materials.grav_const)
    outs = []
    for _ in range(calls):
        eng.rhs_eval(view)
        torch.cuda.synchronize()
        outs.append({k: v.cpu().numpy() for k, v in dev.items()})
    eng.close()
    return outs


def host_calls(config, arrays, cfg, meta, calls, options):
    eng = api.RhsEngine(config, n_max=meta["n"], material_cfg=cfg)
    eng.host_options(options)
    pinned = {k: torch.from_numpy(v.copy()).pin_memory() for k, v in arrays.items()}
    view = api.make_view(pinned, None, meta["n"], max_num_flaws=meta["max_num_flaws"], selfgravity=meta["selfgravity"],
                         theta=meta["theta"], grav_const=eng.materials.grav_const)
    outs, traffic = [], []
    for _ in range(calls):
        traffic.append(eng.rhs_eval_host(view))
        outs.append({k: v.numpy().copy() for k, v in pinned.items()})
    eng.close()
    return outs, traffic


@pytest.mark.parametrize("config", ["sedov", "impact", "rings", "giant_hydro"])
@pytest.mark.parametrize("options", [0, 3])
def test_host_call_is_bit_identical_to_device_call(config, options, tmp_path):
    sc = scenarios.make(config, 20000, stirred=True)
    cfg = state.write_material_files(sc, str(tmp_path))
    mats = api.MaterialTables(config, cfg)
    arrays, meta = state.scenario_arrays(sc, mats)
    ref = device_calls(config, arrays, cfg, meta, calls=3)
    got, traffic = host_calls(config, arrays, cfg, meta, calls=3, options=options)
    for call in range(3):
        for name, want in ref[call].items():
            if options & 2 and name in SCRATCH:
                assert np.array_equal(got[call][name], arrays[name]), f"{name} must not be written with SKIP_SCRATCH"
                continue
            assert np.array_equal(got[call][name], want, equal_nan=True), f"call {call + 1}: {name} differs from the device call"
    h2d = [t[0] for t in traffic]
    d2h = [t[1] for t in traffic]
    assert min(h2d) > 0 and min(d2h) > 0
    if options & 1:
        imm_bytes = sum(arrays[k].nbytes for k in IMMUTABLE if k in arrays)
        assert h2d[1] == h2d[2] == h2d[0] - imm_bytes, (h2d, imm_bytes)
    else:
        assert h2d[0] == h2d[1] == h2d[2]


def test_cached_immutables_are_dropped_on_request(tmp_path):
    """After b200sph_host_options() the next call uploads the immutables again (a changed mass must be seen)."""
    config = "sedov"
    sc = scenarios.make(config, 8000, stirred=True)
    cfg = state.write_material_files(sc, str(tmp_path))
    mats = api.MaterialTables(config, cfg)
    arrays, meta = state.scenario_arrays(sc, mats)
    eng = api.RhsEngine(config, n_max=meta["n"], material_cfg=cfg)
    eng.host_options(eng.HOST_CACHE_IMMUTABLES)
    pinned = {k: torch.from_numpy(v.copy()).pin_memory() for k, v in arrays.items()}
    view = api.make_view(pinned, None, meta["n"], max_num_flaws=meta["max_num_flaws"], selfgravity=meta["selfgravity"],
                         theta=meta["theta"], grav_const=eng.materials.grav_const)
    eng.rhs_eval_host(view)
    rho1 = pinned["rho"].numpy().copy()
    pinned["m"].mul_(2.0)
    eng.rhs_eval_host(view)                       # cached mass: kernel-sum density unchanged
    assert np.array_equal(pinned["rho"].numpy(), rho1)
    eng.host_options(eng.HOST_CACHE_IMMUTABLES)   # drop the cache
    eng.rhs_eval_host(view)
    assert np.allclose(pinned["rho"].numpy(), 2.0 * rho1, rtol=1e-12)
    eng.close()
