"""GPUs (>= 2): the multi-GPU host in C++ over NCCL (csrc/mg.cu) -- what a C host calls -- against the single-domain oracle.
Every rank starts with an arbitrary slice of the particle set (generator order); b200sph_mg_decompose cuts the Morton
curve, b200sph_mg_migrate moves full particle records to their owners, b200sph_mg_rhs_eval runs halo exchange + staged
evaluation.  Global ids ride along in the `depth` member (untouched without self-gravity)."""
import os
import socket
import tempfile
import traceback

import numpy as np
import pytest

import common
from miluphcuda_b200 import api, scenarios, state

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

FIELDS = ("ax", "ay", "az", "drhodt", "dedt", "dhdt", "dSdt", "dddt", "dalphadt", "rho", "p", "cs")


def _worker(rank, world, port, config, n, by_work, result_dir):
    import torch.distributed as dist
    try:
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        dev = torch.device("cuda", rank % torch.cuda.device_count())
        torch.cuda.set_device(dev)
        dist.init_process_group("gloo", rank=rank, world_size=world)   # only to hand the NCCL id around, as MPI_Bcast would
        sc = scenarios.make(config, n, stirred=True)
        with tempfile.TemporaryDirectory() as td:
            cfg = state.write_material_files(sc, td)
            mats = api.MaterialTables(config, cfg)
            full, meta = state.scenario_arrays(sc, mats)
            meta = dict(meta, selfgravity=False)
            N = sc.n
            full["depth"][:] = np.arange(N, dtype=np.int32)
            ref = {k: v.copy() for k, v in full.items()}
            rc, off, _ = common.oracle_rhs(config, ref, mats, dict(meta, n=N))
            assert rc == 0, (rc, off)
            if by_work:   # second pass of a real run: the interaction counts of the last evaluation weight the cut
                full["noi"][:] = ref["noi"]
            lo, hi = rank * N // world, (rank + 1) * N // world       # an arbitrary initial distribution
            n_held, capacity = hi - lo, int(2.5 * N / world) + 4096
            fields = {}
            for name, arr in full.items():
                per = arr.size // N
                buf = np.zeros(capacity * per, dtype=arr.dtype)
                buf.reshape(capacity, per)[:n_held] = arr.reshape(N, per)[lo:hi]
                fields[name] = torch.from_numpy(buf).to(dev)
            eng = api.RhsEngine(config, n_max=capacity, device=dev.index, material_cfg=cfg)
            ids = [api.NativeMultiGpu.unique_id(config) if rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            mg = api.NativeMultiGpu(eng, rank, world, ids[0])
            view = api.make_view(fields, None, n_held, max_num_flaws=meta["max_num_flaws"], grav_const=eng.materials.grav_const)
            mg.decompose(view, n_held, by_work=by_work)
            n_owned = mg.migrate(view, n_held, capacity)
            total = torch.tensor([n_owned], dtype=torch.int64)
            dist.all_reduce(total)
            assert int(total.item()) == N, "particles were lost or duplicated by the migration"
            bad = {}
            for call in range(3):   # plan build, plan reuse, and the fixed point of c_s(p)
                n_total = mg.rhs_eval(view, n_owned, capacity)
                assert n_total > n_owned
            torch.cuda.synchronize()
            gid = fields["depth"][:n_owned].cpu().numpy().astype(np.int64)
            assert len(np.unique(gid)) == n_owned
            for _ in range(2):
                rc, off, _ = common.oracle_rhs(config, ref, mats, dict(meta, n=N))
            got_noi = fields["noi"][:n_owned].cpu().numpy()
            if not np.array_equal(got_noi, ref["noi"][gid]):
                bad["noi"] = int(np.abs(got_noi - ref["noi"][gid]).max())
            for name in FIELDS:
                if name not in fields or name not in ref:
                    continue
                per = ref[name].size // N
                got = fields[name].cpu().numpy().reshape(capacity, per)[:n_owned]
                want = ref[name].reshape(N, per)[gid]
                scale = float(np.sqrt(np.mean(ref[name].astype(np.float64) ** 2)))
                err = common.field_error(got, want, scale)
                if not err <= common.RTOL:
                    bad[name] = err
            st = mg.stats()
            with open(os.path.join(result_dir, f"rank{rank}.txt"), "w") as fh:
                fh.write("OK\n" if not bad else f"MISMATCH {bad}\n")
                fh.write(f"n_owned={n_owned} stats={st}\n")
            assert st["plan_builds"] == 1 and st["stale_plans"] == 0, st
            mg.close()
            eng.close()
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        with open(os.path.join(result_dir, f"rank{rank}.txt"), "w") as fh:
            fh.write("EXCEPTION\n" + traceback.format_exc())
        raise


def _rk_worker(rank, world, port, config, n, result_dir):
    """rk2_adaptive over two GPUs (b200sph_mg_rk2_advance: evaluations through the halo exchange, step-size reductions
    all-reduced) against the same integration of the whole set on one GPU: same steps, same final state."""
    import torch.distributed as dist
    import make_golden
    try:
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        dev = torch.device("cuda", rank % torch.cuda.device_count())
        torch.cuda.set_device(dev)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        sc = scenarios.make(config, n)
        args = make_golden.evolve_args(sc, steps=12)
        t_end, dt_max, eps = float(args[args.index("-t") + 1]), float(args[args.index("-M") + 1]), float(args[args.index("-Q") + 1])
        with tempfile.TemporaryDirectory() as td:
            cfg = state.write_material_files(sc, td)
            mats = api.MaterialTables(config, cfg)
            full, meta = state.scenario_arrays(sc, mats)
            N = sc.n
            full["depth"][:] = np.arange(N, dtype=np.int32)
            kw = dict(max_num_flaws=meta["max_num_flaws"], grav_const=mats.grav_const)

            def integrate(engine, fields, n_rows, advance):
                rk_fields = [{k: torch.zeros_like(v) for k, v in fields.items() if k not in ("materialId", "flaws", "h0")} for _ in range(3)]
                view = api.make_view(fields, None, n_rows, **kw)
                rk = engine.rk2_buffers(rk_fields)
                engine.init_soundspeed(view)
                engine.rk2_init(view, rk)
                prm = engine.rk2_default_params()
                prm.rk_epsrel, prm.dt_max = eps, dt_max
                st = api.Rk2State()
                advance(view, rk, prm, st)
                torch.cuda.synchronize()
                return st

            # the whole set on this rank's GPU
            eng1 = api.RhsEngine(config, n_max=N, device=dev.index, material_cfg=cfg)
            one = {k: torch.from_numpy(v.copy()).to(dev) for k, v in full.items()}
            st1 = integrate(eng1, one, N, lambda view, rk, prm, st: eng1.rk2_advance(view, rk, prm, t_end, st))
            eng1.close()
            # the same set over two GPUs
            lo, hi = rank * N // world, (rank + 1) * N // world
            n_held, capacity = hi - lo, int(2.5 * N / world) + 4096
            fields = {}
            for name, arr in full.items():
                per = arr.size // N
                buf = np.zeros(capacity * per, dtype=arr.dtype)
                buf.reshape(capacity, per)[:n_held] = arr.reshape(N, per)[lo:hi]
                fields[name] = torch.from_numpy(buf).to(dev)
            eng = api.RhsEngine(config, n_max=capacity, device=dev.index, material_cfg=cfg)
            ids = [api.NativeMultiGpu.unique_id(config) if rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            mg = api.NativeMultiGpu(eng, rank, world, ids[0])
            view0 = api.make_view(fields, None, n_held, **kw)
            mg.decompose(view0, n_held)
            n_owned = mg.migrate(view0, n_held, capacity)
            st2 = integrate(eng, fields, n_owned,
                            lambda view, rk, prm, st: mg.rk2_advance(view, rk, prm, t_end, st, n_owned, capacity))
            bad = {}
            if (st1.accepted, st1.rejected) != (st2.accepted, st2.rejected) or st1.accepted < 12:
                bad["steps"] = ((st1.accepted, st1.rejected), (st2.accepted, st2.rejected))
            gid = fields["depth"][:n_owned].cpu().numpy().astype(np.int64)
            for name in ("x", "y", "z", "vx", "vy", "vz", "rho", "e", "h", "S", "d", "alpha_jutzi", "p"):
                if name not in fields:
                    continue
                per = fields[name].numel() // capacity
                got = fields[name].cpu().numpy().reshape(capacity, per)[:n_owned]
                want = one[name].cpu().numpy().reshape(N, per)[gid]
                scale = float(np.sqrt(np.mean(one[name].cpu().numpy().astype(np.float64) ** 2)))
                err = common.field_error(got, want, scale)
                if not err <= 1e-9:
                    bad[name] = err
            with open(os.path.join(result_dir, f"rank{rank}.txt"), "w") as fh:
                fh.write("OK\n" if not bad else f"MISMATCH {bad}\n")
                fh.write(f"accepted={st2.accepted} rejected={st2.rejected} t={st2.t} stats={mg.stats()}\n")
            mg.close()
            eng.close()
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        with open(os.path.join(result_dir, f"rank{rank}.txt"), "w") as fh:
            fh.write("EXCEPTION\n" + traceback.format_exc())
        raise


@pytest.mark.parametrize("config,n", [("sedov", 40000), ("impact", 40000)])
def test_native_distributed_integrator(config, n):
    if torch.cuda.device_count() < 2:
        pytest.skip("the NCCL host needs two GPUs")
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with tempfile.TemporaryDirectory() as rd:
        try:
            mp.spawn(_rk_worker, args=(2, port, config, n, rd), nprocs=2, join=True)
        finally:
            lines = [open(os.path.join(rd, f"rank{r}.txt")).read() if os.path.exists(os.path.join(rd, f"rank{r}.txt")) else "NO RESULT"
                     for r in range(2)]
    assert all(line.startswith("OK") for line in lines), lines


@pytest.mark.parametrize("config,n,by_work", [("sedov", 60000, False), ("impact", 40000, False), ("impact", 40000, True),
                                              ("rings", 40000, False)])
def test_native_host_two_ranks(config, n, by_work):
    if torch.cuda.device_count() < 2:
        pytest.skip("the NCCL host needs two GPUs")
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with tempfile.TemporaryDirectory() as rd:
        try:
            mp.spawn(_worker, args=(2, port, config, n, by_work, rd), nprocs=2, join=True)
        finally:
            lines = [open(os.path.join(rd, f"rank{r}.txt")).read() if os.path.exists(os.path.join(rd, f"rank{r}.txt")) else "NO RESULT"
                     for r in range(2)]
    assert all(line.startswith("OK") for line in lines), lines
