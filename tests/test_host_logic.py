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
