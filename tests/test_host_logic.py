"""Host-side rules that need neither a GPU nor the CUDA library: which halo a switch set needs, and the time limit on the
reference binary that bench.py's evolved-state preparation relies on."""
import os

import pytest




def test_halo_levels_rule():
    """Which halo the switch set needs; the combination a two-level state halo cannot serve is refused, not guessed."""
    from miluphcuda_b200 import multigpu
    assert multigpu.halo_levels({"INTEGRATE_DENSITY": 1, "TENSORIAL_CORRECTION": 0}) == 1
    assert multigpu.halo_levels({"INTEGRATE_DENSITY": 0, "TENSORIAL_CORRECTION": 0}) == 2
    assert multigpu.halo_levels({"INTEGRATE_DENSITY": 1, "TENSORIAL_CORRECTION": 1}) == 2
    assert multigpu.halo_levels({"INTEGRATE_DENSITY": 0, "TENSORIAL_CORRECTION": 1}) == 3
    assert multigpu.halo_levels({"INTEGRATE_DENSITY": 1, "TENSORIAL_CORRECTION": 1}, kernel_sum_materials=True) == 3


def test_reference_run_time_limit(tmp_path, monkeypatch):
    """oracle/make_golden.run_reference kills a reference binary that exceeds the caller's limit and says so
    (bench.py turns that into the step-0 fallback of BOTH arms); a binary that finishes in time is unaffected."""
    import stat
    import sys
    import time
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import make_golden
    from miluphcuda_b200 import scenarios
    fake = tmp_path / "miluphcuda_fake"
    fake.write_text("#!/bin/sh\nsleep ${FAKE_SLEEP:-0}\necho REF_DONE\n")
    fake.chmod(fake.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setattr(make_golden, "ref_binary", lambda config: str(fake))
    sc = scenarios.make("shocktube", 200)
    wd = tmp_path / "run"
    wd.mkdir()
    log = make_golden.run_reference(sc, str(wd), {"FAKE_SLEEP": "0"}, timeout_s=30)
    assert "REF_DONE" in open(log).read()
    t0 = time.time()
    with pytest.raises(make_golden.ReferenceTimeout) as exc:
        make_golden.run_reference(sc, str(wd), {"FAKE_SLEEP": "30"}, evolve=True, timeout_s=1)
    assert time.time() - t0 < 15
    assert "did not finish within 1 s" in str(exc.value)


def _f32_up(v):
    import numpy as np
    f = np.float32(v)
    return np.where(f.astype(np.float64) < v, np.nextafter(f, np.float32(np.inf)), f).astype(np.float32)


def test_fp32_prefilter_never_rejects_an_exact_neighbour():
    """The FP32 pre-filter of the neighbour search (csrc/rhs_kernels.cu: search_threshold, k_gather, search_hit) must be
    conservative: a pair the exact FP64 test of the reference accepts (d < h_i^2 && d < h_j^2, src/tree.cu:851-865) always
    survives it.  Restated in numpy with the same operations and roundings; pairs are placed ON the edge of the kernel
    support (relative distance 1 -+ 1e-7 .. 1e-3), in grids of up to 1000 cells per axis, far from the origin."""
    import numpy as np
    rng = np.random.default_rng(7)
    checked = 0
    for ncm, cell, offset in ((8, 0.25, 0.0), (300, 0.01, 5.0), (1000, 3.0e4, -2.0e7), (1000, 1e-3, 1e3)):
        n = 200000
        lo = np.array([offset, offset - 1.0, offset + 2.0])
        cell_inv = 1.0 / cell
        xi = lo + rng.random((n, 3)) * (ncm * cell)
        hi = cell * (1.0 + rng.random(n) * 3.0)          # h between 1 and 4 cells (variable resolution)
        hj = hi * np.where(rng.random(n) < 0.5, 1.0, 1.0 + rng.random(n))
        direction = rng.normal(size=(n, 3))
        direction /= np.linalg.norm(direction, axis=1)[:, None]
        rel = 1.0 + rng.choice([-1.0, 1.0], n) * 10.0 ** rng.uniform(-7, -3, n)
        band = 2e-3 * max(1.0, ncm / 300.0)               # beyond this the filter has to reject (it IS a filter)
        far = rng.random(n) < 0.2
        rel = np.where(far, 1.0 + band + rng.random(n) * 0.05, rel)
        xj = xi + direction * (np.minimum(hi, hj) * rel)[:, None]
        # exact test, accumulated like the compiled reference: mul, then one fma per further axis (fma emulated in longdouble)
        d = xi - xj
        r2 = (np.longdouble(d[:, 0]) * d[:, 0]).astype(np.float64)
        for a in (1, 2):
            r2 = (np.longdouble(d[:, a]) * d[:, a] + r2).astype(np.float64)
        exact = (r2 < hi * hi) & (r2 < hj * hj)

        def thr(h):
            hc = h * cell_inv
            delta = 2.384185791015625e-07 * (ncm + hc + 1.0)
            margin = 3.4641016151377544 * hc * delta + 3.0 * delta * delta + 9.5367431640625e-07 * hc * hc
            return _f32_up((hc * hc + 2.0 * margin) * 1.000001)

        ui = ((xi - lo) * cell_inv).astype(np.float32)
        uj = ((xj - lo) * cell_inv).astype(np.float32)
        dd = None
        for a in range(3):
            da = (ui[:, a] - uj[:, a]).astype(np.float32)
            sq = da.astype(np.float64) * da.astype(np.float64)          # exact product of two floats
            dd = sq.astype(np.float32) if dd is None else (sq + dd.astype(np.float64)).astype(np.float32)   # fmaf
        hit = dd < np.minimum(thr(hi), thr(hj))
        missed = exact & ~hit
        assert not missed.any(), f"grid {ncm} cells of {cell}: {int(missed.sum())} exact neighbours rejected by the FP32 filter"
        checked += int(exact.sum())
        assert far.sum() > 1000 and not (hit & far).any(), "pairs clearly outside the support survive the FP32 filter"
    assert checked > 100000


def test_row_clipping_keeps_every_exact_neighbour():
    """k_neighbours scans rows of x-adjacent cells clipped to the sphere (csrc/rhs_kernels.cu: stencil_of, row_gap, the
    FP32 square root rounded up).  Restated in numpy: for random targets and random exact neighbours (d < h_i^2), the
    neighbour's cell lies inside the stencil, its row is not skipped, and its cell index lies inside the clipped x range --
    including particles on cell faces, in the clamped outermost cells and beyond the grid."""
    import numpy as np
    rng = np.random.default_rng(11)

    def cell_coord(x, lo, cell_inv, nc):
        c = ((x - lo) * cell_inv).astype(np.int64)        # truncation like the (int) cast (arguments are >= lo - reach)
        c = np.where((x - lo) * cell_inv < 0, np.ceil((x - lo) * cell_inv).astype(np.int64), c)
        return np.clip(c, 0, nc - 1)

    def row_gap(p, lo, cell, c, nc, slack):
        c_lo = c * cell + lo
        below = np.where(c > 0, c_lo - p, -1.0)
        above = np.where(c < nc - 1, p - (c_lo + cell), -1.0)
        return np.maximum(np.maximum(below, above) - slack, 0.0)

    total = 0
    for cell, nc, lo0 in ((0.5, 40, -3.0), (1.0e-2, 700, 100.0), (2.5e4, 64, -8.0e5)):
        n = 300000
        lo = np.array([lo0, lo0 + 0.3 * cell, lo0 - 0.7 * cell])
        ncs = np.array([nc, max(nc // 2, 3), max(nc // 3, 3)])
        cell_inv = 1.0 / cell
        # targets: inside the grid, a tenth exactly on cell faces, a tenth outside the grid (clamped outermost cells)
        p = lo + rng.random((n, 3)) * ncs * cell
        on_face = rng.random(n) < 0.1
        p[on_face] = lo + np.floor((p[on_face] - lo) * cell_inv) * cell
        outside = rng.random(n) < 0.1
        p[outside] += (rng.random((int(outside.sum()), 3)) - 0.5) * 6.0 * cell
        h = cell * (0.3 + rng.random(n) * 3.5)             # from a third of a cell to 3.8 cells
        direction = rng.normal(size=(n, 3))
        direction /= np.linalg.norm(direction, axis=1)[:, None]
        q = p + direction * (h * (1.0 - 10.0 ** rng.uniform(-12, 0, n)))[:, None]   # neighbours from r ~ 0 up to the edge
        d = p - q
        r2 = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]
        exact = r2 < h * h
        reach = (h * cell_inv + 1e-9).astype(np.int64) + 1
        cp = [cell_coord(p[:, a], lo[a], cell_inv, ncs[a]) for a in range(3)]
        cq = [cell_coord(q[:, a], lo[a], cell_inv, ncs[a]) for a in range(3)]
        for a in range(3):
            inside = (cq[a] >= np.maximum(cp[a] - reach, 0)) & (cq[a] <= np.minimum(cp[a] + reach, ncs[a] - 1))
            assert inside[exact].all(), f"axis {a}: an exact neighbour lies outside the stencil"
        reach2 = (h * h) * (1.0 + 1e-9)
        slack = 1e-9 * cell
        gz = row_gap(p[:, 2], lo[2], cell, cq[2], ncs[2], slack)
        rem_z = reach2 - gz * gz
        gy = row_gap(p[:, 1], lo[1], cell, cq[1], ncs[1], slack)
        rem = rem_z - gy * gy
        assert (rem_z[exact] > 0.0).all() and (rem[exact] > 0.0).all(), "the row of an exact neighbour is skipped"
        rem_f = _f32_up(np.maximum(rem, 0.0))
        root = np.sqrt(rem_f.astype(np.float64))
        root_f = _f32_up(root)                              # __fsqrt_ru
        w = root_f.astype(np.float64) * (1.0 + 1e-6) + slack
        xa = np.maximum(np.maximum(cp[0] - reach, 0), cell_coord(p[:, 0] - w, lo[0], cell_inv, ncs[0]))
        xb = np.minimum(np.minimum(cp[0] + reach, ncs[0] - 1), cell_coord(p[:, 0] + w, lo[0], cell_inv, ncs[0]))
        ok = (cq[0] >= xa) & (cq[0] <= xb)
        assert ok[exact].all(), f"{int((~ok & exact).sum())} exact neighbours fall outside the clipped x range of their row"
        total += int(exact.sum())
    assert total > 500000
