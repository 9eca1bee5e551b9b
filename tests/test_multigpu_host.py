"""N>1 host logic on CPU: two gloo ranks, Morton decomposition + halo exchange, and the CPU oracle on
owned + halo particles must reproduce the single-domain result for every owned particle
(identical neighbour counts; rates to 1e-11: the pieces only reorder the neighbour sums)."""
import numpy as np
import pytest

import mg_worker
from miluphcuda_b200 import multigpu


# rings: two separate bodies, so three ranks are needed for every rank to have foreign neighbours
@pytest.mark.parametrize("config,n,world", [("sedov", 12000, 2), ("impact", 8000, 2), ("giant_hydro", 8000, 2),
                                            ("rings", 8000, 3), ("shocktube", 3000, 2)])
def test_ranks_match_single_domain(config, n, world):
    lines = mg_worker.run(config, n, world, "oracle", "gloo")
    assert all(line.startswith("OK") for line in lines), lines


@pytest.mark.parametrize("config,n,world", [("sedov", 12000, 2), ("impact", 8000, 2), ("rings", 8000, 3), ("giant_hydro", 8000, 2)])
def test_plan_decision_survives_motion_within_its_tolerance(config, n, world):
    """The rule a reusable send plan is built from -- reach (h_k + h_max(box)) (1 + growth) + 2 D -- decided at the
    initial positions must still be complete after every particle moved by D TOWARDS the nearest foreign domain and h
    grew by up to `growth`: the oracle on owned + halo of the MOVED state reproduces the single-domain result of the
    moved state.  (The same motion breaks a decision taken without the head-room: neighbour counts differ.)"""
    lines = mg_worker.run(config, n, world, "oracle_plan", "gloo")
    assert all(line.startswith("OK") for line in lines), lines


def test_morton_partition_is_a_partition():
    rng = np.random.default_rng(5)
    x = rng.random((50000, 3))
    dec, parts = multigpu.morton_partition(x, 8)
    allidx = np.concatenate(parts)
    assert np.array_equal(np.sort(allidx), np.arange(50000))
    assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 0.02 * 50000 / 8
    # every particle lies inside one of its owner's boxes, and the boxes of all ranks tile the cube
    vol = 0.0
    for r, p in enumerate(parts):
        boxes = dec.boxes(r)
        inside = np.zeros(len(p), dtype=bool)
        for b in boxes:
            inside |= np.all((x[p] >= b[:3]) & (x[p] <= b[3:]), axis=1)
            vol += np.prod(b[3:] - b[:3])
        assert inside.all()
        assert len(boxes) <= 2 * 7 * dec.level
    assert abs(vol - dec.span ** 3) < 1e-9 * dec.span ** 3


def test_cell_ids_numpy_and_torch_agree():
    import torch
    rng = np.random.default_rng(6)
    for dim in (1, 2, 3):
        x = rng.random((1000, dim)) * 3.0 - 1.0
        dec = multigpu.MortonDecomposition(dim, x.min(axis=0), x.max(axis=0), 4)
        a = dec.cell_ids(x)
        b = dec.cell_ids(torch.from_numpy(x)).numpy()
        assert np.array_equal(a, b)
        assert a.min() >= 0 and a.max() < dec.n_cells


def test_halo_levels():
    assert multigpu.halo_levels({"INTEGRATE_DENSITY": 0}) == 2
    assert multigpu.halo_levels({"INTEGRATE_DENSITY": 1, "TENSORIAL_CORRECTION": 1}) == 2
    assert multigpu.halo_levels({"INTEGRATE_DENSITY": 1}) == 1
