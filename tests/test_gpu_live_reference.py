"""GPU: the CUDA path against the reference's own build RUN LIVE on the same box (oracle/_ref/miluphcuda_<config>,
compiled unmodified from /root/reference by oracle/build_ref.sh), at BASELINE.json's sizes and on states the small
golden files cannot hold:

* evolved states -- the reference's own rk2_adaptive integrates the synthetic scenario over >= 20 accepted steps
  (oracle/ref_hook.cu, REF_EVOLVE), dumps what its buffers hold, evaluates rightHandSide() once and dumps again; the
  CUDA path is run on the first dump and compared with the second: neighbour sets bit-exact as sets, every state and
  rate field within 1e-9 (per-field scale).  10^5 particles per switch set, 10^6 for sedov and the impact;
* the reference's SHIPPED inputs: examples/impact/impact.0000.gz (58 402 particles, SEAGen arrangement,
  real flaw distribution), examples/giant_collisions/{hydro,solid}/impact.0000.gz -- staged next to the binaries by
  oracle/build_ref.sh (oracle/_ref/fixtures/), at step 0 and evolved.

Nothing here reads /root/reference: binaries and fixtures travel with the snapshot.  Skipped when they are absent."""
import gzip
import os
import re
import shutil
import types

import numpy as np
import pytest

import common
import make_golden
from miluphcuda_b200 import api, scenarios, state

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

FIXTURES = os.path.join(common.REPO, "oracle", "_ref", "fixtures")


def have_binary(config):
    return os.path.exists(os.path.join(common.REPO, "oracle", "_ref", f"miluphcuda_{config}"))


def pair_keys_dense(nbrs, n):
    rows = np.repeat(np.arange(n, dtype=np.int64), nbrs.shape[1]).reshape(nbrs.shape)
    mask = nbrs >= 0
    keys = rows[mask] * n + nbrs[mask].astype(np.int64)
    keys.sort()
    return keys


def pair_keys_csr(noi, idx, n):
    rows = np.repeat(np.arange(n, dtype=np.int64), noi)
    keys = rows * n + idx.astype(np.int64)
    keys.sort()
    return keys


def run_and_compare(config, sc, wd, evolve, input_file=None, min_steps=20, cfg_path=None):
    """Reference live (dump in/out1 with compacted lists) -> CUDA path on the `in` state -> compare."""
    env = {"REF_DUMP": os.path.join(wd, "dump"), "REF_DUMP_LISTS": "2"}
    log = make_golden.run_reference(sc, wd, env, evolve=evolve, input_file=input_file)
    text = open(log).read()
    if evolve:
        acc = re.findall(r"Had to integrate (\d+) timesteps \((\d+) accepted, (\d+) rejected\)", text)
        assert acc and int(acc[-1][1]) >= min_steps, f"reference took {acc} steps, wanted >= {min_steps} accepted"
    d_in = make_golden.read_dump(os.path.join(wd, "dump.in.bin"))
    d1 = make_golden.read_dump(os.path.join(wd, "dump.out1.bin"))
    for f in ("dump.in.bin", "dump.out1.bin", "dump.out2.bin"):
        os.remove(os.path.join(wd, f))
    arrays, meta = make_golden.arrays_from_dump(config, d_in, bool(sc.selfgravity))
    n = meta["n"]
    eng = api.RhsEngine(config, n_max=n, material_cfg=cfg_path or os.path.join(wd, "material.cfg"))
    dev = {k: torch.from_numpy(v).cuda() for k, v in arrays.items()}
    view = api.make_view(dev, None, n, max_num_flaws=meta["max_num_flaws"], selfgravity=meta["selfgravity"],
                         theta=sc.theta, grav_const=eng.materials.grav_const)
    eng.rhs_eval(view)
    torch.cuda.synchronize()
    out = {k: v.cpu().numpy() for k, v in dev.items()}
    # neighbour sets, bit-exact as sets
    assert np.array_equal(out["noi"], d1["noi"]), "neighbour counts differ from the live reference"
    maxni = eng.lib.b200sph_switch_value(b"MAX_NUM_INTERACTIONS")
    width = int(out["noi"].max()) + 1
    buf = torch.empty((n, width), dtype=torch.int32, device="cuda")
    eng.export_interactions(buf, width)
    ours = pair_keys_dense(buf.cpu().numpy(), n)
    del buf
    theirs = pair_keys_csr(d1["noi"], d1["nbr_idx"], n)
    assert ours.shape == theirs.shape and np.array_equal(ours, theirs), "neighbour sets differ from the live reference"
    eng.close()
    bad = {}
    for name in common.RATE_FIELDS + common.STATE_FIELDS:
        if name in out and name in d1 and d1[name].shape == out[name].shape:
            err = common.field_error(out[name], d1[name])
            if not err <= common.RTOL:
                bad[name] = err
    for name in common.INT_COMPARE:
        if name in out and name in d1:
            assert np.array_equal(out[name], d1[name]), name
    assert not bad, f"relative errors above {common.RTOL} against the live reference: {bad}"
    return n, maxni


LIVE_CASES = [("shocktube", 10000), ("sedov", 100000), ("rings", 100000), ("impact", 100000), ("giant_hydro", 100000),
              ("giant_solid", 60000), ("nakamura", 100000), ("sedov", 1000000), ("impact", 1000000)]


@pytest.mark.parametrize("config,n", LIVE_CASES)
def test_evolved_state_against_live_reference(config, n, tmp_path):
    if not have_binary(config):
        pytest.skip(f"oracle/_ref/miluphcuda_{config} not built (oracle/build_ref.sh needs /root/reference)")
    sc = scenarios.make(config, n)
    run_and_compare(config, sc, str(tmp_path), evolve=True)


SHIPPED = {
    # fixture -> (config, selfgravity, smallest smoothing length for the evolve step estimate)
    "impact": ("impact", False, 0.5),
    "giant_hydro": ("giant_hydro", True, 171776.0),
    "giant_solid": ("giant_solid", True, 171776.0),
}


@pytest.mark.parametrize("evolve", [False, True], ids=["step0", "evolved"])
@pytest.mark.parametrize("fixture", sorted(SHIPPED))
def test_shipped_input_against_live_reference(fixture, evolve, tmp_path):
    config, selfgravity, hmin = SHIPPED[fixture]
    src = os.path.join(FIXTURES, fixture)
    if not have_binary(config) or not os.path.exists(os.path.join(src, "impact.0000.gz")):
        pytest.skip("shipped fixture or reference binary not staged (oracle/build_ref.sh needs /root/reference)")
    wd = str(tmp_path)
    for f in os.listdir(src):
        if f.endswith(".cfg"):
            shutil.copy(os.path.join(src, f), os.path.join(wd, f))
    data = os.path.join(wd, "impact.0000")
    with gzip.open(os.path.join(src, "impact.0000.gz"), "rb") as fi, open(data, "wb") as fo:
        shutil.copyfileobj(fi, fo)
    sc = types.SimpleNamespace(config=config, selfgravity=selfgravity, theta=0.5, h=np.array([hmin]), material_cfg="",
                               e=np.zeros(1))
    n, _ = run_and_compare(config, sc, wd, evolve=evolve, input_file=data)
    assert n > 50000
