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
