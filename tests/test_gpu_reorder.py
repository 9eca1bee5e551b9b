"""GPU: b200sph_reorder (persistent cell order of the caller's buffers).  The permuted buffers hold
the same particles, every array moved by the same permutation, and an evaluation on them gives the relabelled answer:
identical neighbour counts and numActiveFlaws, rates within 1e-9, neighbour sets equal after mapping the indices."""
import numpy as np
import pytest

import common
from miluphcuda_b200 import api, scenarios, state

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("config", ["sedov", "impact", "giant_hydro", "rings", "shocktube"])
def test_reorder_is_a_relabelling(config, tmp_path):
    sc = scenarios.make(config, 30000, stirred=True)
    cfg = state.write_material_files(sc, str(tmp_path))
    eng = api.RhsEngine(config, n_max=sc.n, material_cfg=cfg)
    arrays, meta = state.scenario_arrays(sc, eng.materials)
    n = meta["n"]
    rng = np.random.default_rng(3)
    shuffle = rng.permutation(n)   # start from an order unrelated to space, like an input file's
    arrays = {k: np.ascontiguousarray(v.reshape(n, -1)[shuffle].reshape(v.shape)) for k, v in arrays.items()}
    kw = dict(max_num_flaws=meta["max_num_flaws"], selfgravity=meta["selfgravity"], theta=meta["theta"], grav_const=eng.materials.grav_const)

    dev_a = {k: torch.from_numpy(v).cuda() for k, v in arrays.items()}
    eng.rhs_eval(api.make_view(dev_a, None, n, **kw))
    torch.cuda.synchronize()
    maxni = int(dev_a["noi"].max().item()) + 1
    nbr_a = torch.empty((n, maxni), dtype=torch.int32, device="cuda")
    eng.export_interactions(nbr_a, maxni)

    dev_b = {k: torch.from_numpy(v).cuda() for k, v in arrays.items()}
    extra_fields = [{"x": dev_b["x"].clone(), "noi": torch.arange(n, dtype=torch.int32, device="cuda")}]
    extra = eng.rk2_buffers(extra_fields + [{}, {}])
    perm = torch.empty(n, dtype=torch.int32, device="cuda")
    view_b = api.make_view(dev_b, None, n, **kw)
    eng.reorder(view_b, extra, 1, perm)
    torch.cuda.synchronize()
    pm = perm.cpu().numpy().astype(np.int64)
    assert np.array_equal(np.sort(pm), np.arange(n)), "not a permutation"
    for name, arr in arrays.items():   # every member moved by the same permutation
        per = arr.size // n
        assert np.array_equal(dev_b[name].cpu().numpy().reshape(n, per), arr.reshape(n, per)[pm]), name
    assert np.array_equal(extra_fields[0]["x"].cpu().numpy(), arrays["x"][pm])
    assert np.array_equal(extra_fields[0]["noi"].cpu().numpy(), pm.astype(np.int32))

    eng.rhs_eval(view_b)
    torch.cuda.synchronize()
    st = eng.stats()
    out_a = {k: v.cpu().numpy() for k, v in dev_a.items()}
    out_b = {k: v.cpu().numpy() for k, v in dev_b.items()}
    for name in common.INT_COMPARE:
        if name in out_a:
            assert np.array_equal(out_b[name], out_a[name][pm]), name
    bad = {}
    for name in common.RATE_FIELDS + common.STATE_FIELDS:
        if name in out_a:
            per = out_a[name].size // n
            err = common.field_error(out_b[name].reshape(n, per), out_a[name].reshape(n, per)[pm])
            if not err <= common.RTOL:
                bad[name] = err
    assert not bad, bad
    nbr_b = torch.empty((n, maxni), dtype=torch.int32, device="cuda")
    eng.export_interactions(nbr_b, maxni)
    a, b = nbr_a.cpu().numpy(), nbr_b.cpu().numpy()
    inv = np.empty(n, dtype=np.int64)
    inv[pm] = np.arange(n)
    a_mapped = np.where(a >= 0, inv[np.maximum(a, 0)], -1)[pm]   # old ids -> new ids, rows in the new order
    assert np.array_equal(np.sort(a_mapped, axis=1), np.sort(b, axis=1)), "neighbour sets differ after relabelling"
    # the point of it: after the reorder the library's own sort finds the buffers (nearly) in order already
    assert st["kernel_launches"] > 0
    eng.close()
