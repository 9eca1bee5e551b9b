"""GPU: the CUDA path at BASELINE.json's full size (10^6 particles), where the CPU oracle would take minutes,
checked through properties that do not depend on the size:

  * neighbour sets are symmetric (the rule d < h_i^2 and d < h_j^2 is), rows are -1 padded behind noi entries,
    and the statistics the library reports are those of the lists;
  * hydro without tensorial correction: the pair forces cancel, sum_i m_i a_i = 0 to rounding
    (reference src/internal_forces.cu:699-761: the pair term is antisymmetric for a fixed smoothing length);
  * relabelling invariance: the library never reorders the caller's buffers, so a permuted particle set must give
    the permuted answer -- identical neighbour counts, rates equal within the 1e-9 gate (summation order differs).
"""
import numpy as np
import pytest

import common
from miluphcuda_b200 import api, scenarios, state

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

FULL_SIZE = 1000000
COMPARE = ("ax", "ay", "az", "drhodt", "dedt", "dhdt", "dSdt", "dddt", "dalphadt", "rho", "p", "cs")


def _evaluate(eng, arrays, meta):
    dev = {k: torch.from_numpy(v).cuda() for k, v in arrays.items()}
    view = api.make_view(dev, None, meta["n"], max_num_flaws=meta["max_num_flaws"], selfgravity=meta["selfgravity"],
                         theta=meta["theta"], grav_const=eng.materials.grav_const)
    eng.rhs_eval(view)
    torch.cuda.synchronize()
    return dev


@pytest.mark.parametrize("config", ["sedov", "impact"])
def test_full_size_properties(config, tmp_path):
    sc = scenarios.make(config, FULL_SIZE, stirred=True)
    n = sc.n
    cfg = state.write_material_files(sc, str(tmp_path))
    eng = api.RhsEngine(config, n_max=n, material_cfg=cfg)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    arrays, meta = state.scenario_arrays(sc, eng.materials)
    dev = _evaluate(eng, arrays, meta)
    stats = eng.stats()
    noi = dev["noi"]
    assert stats["kernel_launches"] > 0
    assert int(noi.sum().item()) == stats["total_noi"]
    assert int(noi.max().item()) == stats["max_noi"]

    # ---- neighbour lists: padding and symmetry
    maxni = eng.lib.b200sph_switch_value(b"MAX_NUM_INTERACTIONS")
    lists = torch.empty((n, maxni), dtype=torch.int32, device="cuda")
    eng.export_interactions(lists, maxni)
    torch.cuda.synchronize()
    valid = lists >= 0
    assert torch.equal(valid.sum(dim=1).to(torch.int32), noi), "rows must hold exactly noi entries, -1 behind them"
    cols = torch.arange(maxni, device="cuda").unsqueeze(0)
    assert torch.equal(valid, cols < noi.unsqueeze(1)), "entries must be packed at the front of a row"
    i_idx = torch.arange(n, device="cuda").unsqueeze(1).expand(n, maxni)[valid]
    j_idx = lists[valid].to(torch.int64)
    del lists, valid
    assert bool((j_idx != i_idx).all()), "a particle is not its own interaction partner"
    forward = torch.sort(i_idx * n + j_idx).values
    backward = torch.sort(j_idx * n + i_idx).values
    assert torch.equal(forward, backward), "neighbour sets are not symmetric"
    del forward, backward, i_idx, j_idx

    # ---- momentum balance of the antisymmetric pair force (hydro, fixed h, no tensorial correction)
    sw = sc.switches()
    if not sw.get("SOLID", 0) and not sw.get("VARIABLE_SML", 0):
        m = dev["m"]
        total = torch.stack([(m * dev[a]).sum() for a in ("ax", "ay", "az")[: sc.dim]])
        scale = sum((m * dev[a]).abs().sum() for a in ("ax", "ay", "az")[: sc.dim])
        assert float(total.abs().max() / scale) < 1e-10, (total, scale)

    first = {k: dev[k].cpu().numpy() for k in COMPARE + ("noi",) if k in dev}
    del dev

    # ---- relabelling invariance
    perm = np.random.default_rng(20240229).permutation(n)
    shuffled = {k: np.ascontiguousarray(v.reshape(n, -1)[perm].reshape(-1)) for k, v in arrays.items()}
    dev2 = _evaluate(eng, shuffled, meta)
    assert np.array_equal(dev2["noi"].cpu().numpy(), first["noi"][perm])
    bad = {}
    for name in COMPARE:
        if name not in dev2:
            continue
        want = first[name].reshape(n, -1)[perm].reshape(-1)
        err = common.field_error(dev2[name].cpu().numpy(), want)
        if not err <= common.RTOL:
            bad[name] = err
    assert not bad, f"relabelled evaluation differs beyond {common.RTOL}: {bad}"
    eng.close()
