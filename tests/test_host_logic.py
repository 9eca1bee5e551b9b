

def test_halo_levels_rule():
    """Which halo the switch set needs; the combination a two-level state halo cannot serve is refused, not guessed."""
    from miluphcuda_b200 import multigpu
    assert multigpu.halo_levels({"INTEGRATE_DENSITY": 1, "TENSORIAL_CORRECTION": 0}) == 1
    assert multigpu.halo_levels({"INTEGRATE_DENSITY": 0, "TENSORIAL_CORRECTION": 0}) == 2
    assert multigpu.halo_levels({"INTEGRATE_DENSITY": 1, "TENSORIAL_CORRECTION": 1}) == 2
    assert multigpu.halo_levels({"INTEGRATE_DENSITY": 0, "TENSORIAL_CORRECTION": 1}) == 3
    assert multigpu.halo_levels({"INTEGRATE_DENSITY": 1, "TENSORIAL_CORRECTION": 1}, kernel_sum_materials=True) == 3
