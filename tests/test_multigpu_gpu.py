"""N>1 product path on GPUs: each rank runs libb200sph on its Morton piece + halo (b200sph_set_owned,
b200sph_set_gravity_sources) and must match the single-domain oracle for its owned particles.
With two or more GPUs the exchange runs over NCCL; on a one-GPU box the two ranks share cuda:0 and
exchange over gloo (host staging), which still exercises the library's owned/halo and gravity-source paths."""
import pytest

import mg_worker

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("config,n", [("sedov", 30000), ("impact", 20000), ("giant_hydro", 20000), ("giant_solid", 20000)])
def test_two_ranks_cuda(config, n):
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    lines = mg_worker.run(config, n, 2, "cuda", backend)
    assert all(line.startswith("OK") for line in lines), lines
