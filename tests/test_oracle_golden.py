"""CPU: the oracle (oracle/sph_oracle.c) against golden vectors produced by the reference's own
CUDA build on a B200 (oracle/make_golden.py).  This is the pin that lets the GPU parity tests
trust the oracle at sizes for which no golden file exists."""
import numpy as np
import pytest

import common
from miluphcuda_b200 import api


@pytest.mark.parametrize("case", common.GOLDEN_CASES)
def test_oracle_matches_reference(case):
    config = common.config_of(case)
    g = common.load_golden(case)
    arrays, meta = common.state_from_golden(g, config)
    td, cfg = common.tmp_material(g)
    try:
        mats = api.MaterialTables(config, cfg)
        # first call: p pinned to 0 by the dump hook
        rc, off, inter = common.oracle_rhs(config, arrays, mats, meta)
        assert rc == 0, f"oracle rc={rc} offender={off}"
        ref_nbrs = common.golden_neighbours(g)
        noi = arrays["noi"]
        dead = common.deactivated_rows(g)   # their own lists are undefined in the reference (common.deactivated_rows)
        live = np.ones(meta["n"], dtype=bool) if dead is None else ~dead
        assert np.array_equal(noi[live], g["out1_noi"][live]), "neighbour counts differ"
        for i in np.nonzero(live)[0]:
            assert np.array_equal(inter[i, : noi[i]], ref_nbrs[i]), f"neighbour set of particle {i} differs"
        rep1 = common.compare_fields(arrays, g, "out1", common.RATE_FIELDS + common.STATE_FIELDS)
        for name in common.INT_COMPARE:
            ref = common.golden_expected(g, "out1", name)
            if name in arrays and ref is not None:
                assert np.array_equal(arrays[name][live], ref[live]), name
        bad = {k: v for k, v in rep1.items() if not v <= common.RTOL}
        assert not bad, f"call 1 mismatches (rel. error): {bad}"
        # second call on the same buffers: c_s now sees the self-consistent pressure
        rc, off, inter = common.oracle_rhs(config, arrays, mats, meta)
        assert rc == 0
        rep2 = common.compare_fields(arrays, g, "out2", common.RATE_FIELDS + common.STATE_FIELDS)
        bad = {k: v for k, v in rep2.items() if not v <= common.RTOL}
        assert not bad, f"call 2 mismatches (rel. error): {bad}"
    finally:
        td.cleanup()
