"""Worker of the multi-rank tests (spawned once per rank by torch.multiprocessing).

mode "oracle": halo exchange on CPU tensors over gloo, then the CPU oracle on owned + halo particles --
               checks that the decomposition hands every rank everything its particles interact with.
mode "cuda"  : the same exchange, then libb200sph on a GPU (rank -> cuda:(rank % device_count)) with
               b200sph_set_owned / b200sph_set_gravity_sources -- the N>1 product path.
Each rank compares ITS owned particles with the single-domain oracle result computed from the same inputs.
"""
from __future__ import annotations

import os
import sys
import tempfile
import traceback

import numpy as np

import common
from miluphcuda_b200 import api, multigpu, scenarios, state

FIELDS = ("ax", "ay", "az", "drhodt", "dedt", "dhdt", "dSdt", "dddt", "dalphadt", "rho", "p", "cs", "g_ax", "g_ay", "g_az")


def _worker(rank: int, world: int, port: int, config: str, n: int, mode: str, backend: str, result_dir: str):
    import torch
    import torch.distributed as dist
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.set_num_threads(2)
        use_cuda = mode == "cuda"
        dev = torch.device("cpu")
        if use_cuda:
            dev = torch.device("cuda", rank % torch.cuda.device_count())
            torch.cuda.set_device(dev)
        kw = {"device_id": dev} if (use_cuda and backend == "nccl") else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)

        sc = scenarios.make(config, n, stirred=True)
        sw = sc.switches()
        with tempfile.TemporaryDirectory() as td:
            cfg = state.write_material_files(sc, td)
            mats = api.MaterialTables(config, cfg)
            full, meta = state.scenario_arrays(sc, mats)
            gravity = bool(meta["selfgravity"]) and use_cuda      # the oracle has no notion of foreign gravity sources
            meta_run = dict(meta, selfgravity=gravity)

            # mode "oracle_plan": the halo decision is taken at the scenario's positions with the head-room of a
            # reusable plan (reach * (1 + growth) + 2 D), THEN every particle moves by up to D and h grows by up to
            # `growth`; the old decision must still hand every rank all it needs for the moved state.
            moved = None
            if mode == "oracle_plan":
                rng = np.random.default_rng(424242)
                h_evolves = bool(sw.get("VARIABLE_SML", 0) or sw.get("INTEGRATE_SML", 0))
                growth = multigpu.HaloExchange.H_GROWTH if h_evolves else 0.0
                max_move = multigpu.HaloExchange.SKIN * float(full["h"].min())
                # adversarial motion: every particle heads for the nearest foreign domain (pairs across a cut approach
                # each other by almost 2 D); particles that touch a foreign box move in a random direction
                pos0 = np.stack([full[a] for a in ["x", "y", "z"][: sc.dim]], axis=1)
                dec0, _ = multigpu.morton_partition(pos0, world)
                boxes0, box_rank0 = dec0.all_boxes()
                owner0 = dec0.owner_of(dec0.cell_ids(pos0))
                direction = rng.normal(size=(sc.n, sc.dim))
                best = np.full(sc.n, np.inf)
                for b in range(len(box_rank0)):
                    vec = np.maximum(boxes0[b, : sc.dim] - pos0, 0.0) - np.maximum(pos0 - boxes0[b, 3: 3 + sc.dim], 0.0)
                    d2 = (vec * vec).sum(axis=1)
                    better = (owner0 != box_rank0[b]) & (d2 < best) & (d2 > 0.0)
                    direction[better] = vec[better]
                    best[better] = d2[better]
                direction /= np.linalg.norm(direction, axis=1, keepdims=True)
                step = direction * (0.999 * max_move)
                moved = {k: v.copy() for k, v in full.items()}
                for a, name in enumerate(["x", "y", "z"][: sc.dim]):
                    moved[name] += step[:, a]
                moved["h"] *= 1.0 + 0.999 * growth * rng.random(sc.n)

            # single-domain answer
            ref = {k: v.copy() for k, v in (moved if moved is not None else full).items()}
            rc, off, _ = common.oracle_rhs(config, ref, mats, dict(meta_run, n=sc.n))
            assert rc == 0, (rc, off)

            local, n_owned, capacity, mine, dec = multigpu.scatter_scenario(full, sc.n, sc.dim, meta["max_num_flaws"], rank, world)
            comm_dev = dev if backend == "nccl" else torch.device("cpu")
            fields = {k: torch.from_numpy(v).to(comm_dev) for k, v in local.items()}
            eng = api.RhsEngine(config, n_max=capacity, device=dev.index, material_cfg=cfg) if use_cuda else None
            if eng is not None:
                eng.set_stream(torch.cuda.current_stream().cuda_stream)   # library kernels ordered with torch / NCCL work
            halo = multigpu.HaloExchange(fields, capacity, dec, levels=multigpu.halo_levels(sw), engine=eng)
            if moved is None:
                n_total = halo.run(n_owned)
            else:
                send_idx = halo.select_cpu(n_owned, 1.0 + growth, 2.0 * max_move)      # decided before anything moved
                for name in ["x", "y", "z"][: sc.dim] + ["h"]:
                    fields[name][:n_owned] = torch.from_numpy(moved[name][mine])
                n_total = halo.move_cpu(n_owned, send_idx)                             # current state of the old selection
            assert n_total > n_owned, "a rank without halo particles means the decomposition is not being exercised"

            if use_cuda:
                dfields = {k: v.to(dev) for k, v in fields.items()}
                if gravity:
                    gs = multigpu.GravitySources(sc.dim)
                    gsrc = {k: v for k, v in (fields if backend == "nccl" else fields).items()}
                    x, y, z, m, n_src, own_begin = gs.gather(gsrc, n_owned)
                    to = lambda t: None if t is None else t.to(dev)
                    x, y, z, m = to(x), to(y), to(z), to(m)
                    eng.set_gravity_sources(x, y, z, m, n_src, own_begin)
                view = api.make_view(dfields, None, n_total, max_num_flaws=meta["max_num_flaws"], selfgravity=gravity,
                                     theta=meta["theta"], grav_const=eng.materials.grav_const)
                eng.set_owned(n_owned)
                eng.rhs_eval(view)
                torch.cuda.synchronize()
                stats = eng.stats()
                assert stats["kernel_launches"] > 0
                out = {k: v.cpu().numpy() for k, v in dfields.items()}
                eng.close()
            else:
                out = {k: v.numpy() for k, v in fields.items()}
                sub = {k: v.reshape(capacity, -1)[:n_total].reshape(-1).copy() for k, v in out.items()}
                rc, off, _ = common.oracle_rhs(config, sub, mats, dict(meta_run, n=n_total))
                assert rc == 0, (rc, off)
                out = sub
                capacity_rows = n_total
            rows = capacity if use_cuda else n_total

            bad = {}
            got_noi = out["noi"].reshape(rows, -1)[:n_owned, 0]
            if not np.array_equal(got_noi, ref["noi"][mine]):
                bad["noi"] = int(np.abs(got_noi - ref["noi"][mine]).max())
            tol = common.RTOL if use_cuda else 1e-11
            for name in FIELDS:
                if name not in out or name not in ref:
                    continue
                per = ref[name].size // sc.n
                got = out[name].reshape(rows, per)[:n_owned]
                want = ref[name].reshape(sc.n, per)[mine]
                # scale of the whole field (not of this rank's piece) so a quiet piece is not judged against noise
                scale = float(np.sqrt(np.mean(ref[name].astype(np.float64) ** 2)))
                err = common.field_error(got, want, scale)
                if not err <= tol:
                    bad[name] = err
            with open(os.path.join(result_dir, f"rank{rank}.txt"), "w") as fh:
                fh.write("OK\n" if not bad else f"MISMATCH {bad}\n")
                fh.write(f"n_owned={n_owned} n_halo={n_total - n_owned} sent={halo.last}\n")
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        with open(os.path.join(result_dir, f"rank{rank}.txt"), "w") as fh:
            fh.write("EXCEPTION\n" + traceback.format_exc())
        raise


def _plan_worker(rank: int, world: int, port: int, config: str, n: int, result_dir: str):
    """NCCL only: DistributedRhs with the reusable send plan.  Evaluation 1 builds the plan; before evaluation 2
    particles deep inside every rank are moved by three times the plan's tolerance, so its check must fail and the
    plan must be rebuilt before the evaluation runs; evaluation 3 (nothing moved) must reuse the second plan.
    The owned particles of every evaluation are compared with the single-domain oracle."""
    import torch
    import torch.distributed as dist
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.set_num_threads(2)
        dev = torch.device("cuda", rank % torch.cuda.device_count())
        torch.cuda.set_device(dev)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        sc = scenarios.make(config, n, stirred=True)
        sw = sc.switches()
        notes, bad = [], {}
        with tempfile.TemporaryDirectory() as td:
            cfg = state.write_material_files(sc, td)
            mats = api.MaterialTables(config, cfg)
            full, meta = state.scenario_arrays(sc, mats)
            meta_run = dict(meta, selfgravity=False)
            local, n_owned, capacity, mine, dec = multigpu.scatter_scenario(full, sc.n, sc.dim, meta["max_num_flaws"], rank, world)
            fields = {k: torch.from_numpy(v).to(dev) for k, v in local.items()}
            eng = api.RhsEngine(config, n_max=capacity, device=dev.index, material_cfg=cfg)
            eng.set_stream(torch.cuda.current_stream().cuda_stream)
            drhs = multigpu.DistributedRhs(eng, fields, capacity, n_owned, dec, meta_run, sw)

            def check(label, ref_arrays):
                ref = {k: v.copy() for k, v in ref_arrays.items()}
                rc, off, _ = common.oracle_rhs(config, ref, mats, dict(meta_run, n=sc.n))
                assert rc == 0, (rc, off)
                out = {k: v.cpu().numpy() for k, v in fields.items()}
                got_noi = out["noi"].reshape(capacity, -1)[:n_owned, 0]
                if not np.array_equal(got_noi, ref["noi"][mine]):
                    bad[label + ":noi"] = int(np.abs(got_noi - ref["noi"][mine]).max())
                for name in FIELDS:
                    if name not in out or name not in ref or name.startswith("g_a"):
                        continue
                    per = ref[name].size // sc.n
                    got = out[name].reshape(capacity, per)[:n_owned]
                    want = ref[name].reshape(sc.n, per)[mine]
                    scale = float(np.sqrt(np.mean(ref[name].astype(np.float64) ** 2)))
                    err = common.field_error(got, want, scale)
                    if not err <= common.RTOL:
                        bad[f"{label}:{name}"] = err

            drhs.eval()
            check("eval1", full)
            assert drhs.halo.plan_builds == 1 and drhs.halo.stale_plans == 0, (drhs.halo.plan_builds, drhs.halo.stale_plans)

            # move particles that are far from every foreign domain (so halo membership cannot change) beyond the tolerance
            max_move = drhs.halo._plan["max_move"]
            boxes, box_rank = dec.all_boxes()
            ax = ["x", "y", "z"][: sc.dim]
            pos = np.stack([full[a] for a in ax], axis=1)
            owner = dec.owner_of(dec.cell_ids(pos))
            hmax = float(full["h"].max())
            deep = np.ones(sc.n, dtype=bool)
            for b in range(len(box_rank)):
                foreign = owner != box_rank[b]
                gap = np.maximum(np.maximum(boxes[b, : sc.dim] - pos, pos - boxes[b, 3: 3 + sc.dim]), 0.0)
                deep &= ~(foreign & ((gap * gap).sum(axis=1) < (3.0 * hmax) ** 2))
            assert deep.sum() > 0, "no particle is deep inside its rank: enlarge the test case"
            moved = {k: v.copy() for k, v in full.items()}
            moved["x"][deep] += 3.0 * max_move

            def load_state(src):
                # the integrator hands every evaluation the state it integrated: (re)load all owned rows
                for name, arr in src.items():
                    per = arr.size // sc.n
                    fields[name].view(capacity, per)[:n_owned] = torch.from_numpy(arr.reshape(sc.n, per)[mine]).to(dev)

            load_state(moved)
            drhs.eval()
            check("eval2", moved)
            assert drhs.halo.plan_builds == 2 and drhs.halo.stale_plans == 1, (drhs.halo.plan_builds, drhs.halo.stale_plans)
            load_state(moved)
            drhs.eval()
            check("eval3", moved)
            assert drhs.halo.plan_builds == 2 and drhs.halo.stale_plans == 1, (drhs.halo.plan_builds, drhs.halo.stale_plans)
            notes.append(f"n_owned={n_owned} n_halo={drhs.n_total - n_owned} moved={int(deep[mine].sum())} max_move={max_move:.3e}")
            eng.close()
        with open(os.path.join(result_dir, f"rank{rank}.txt"), "w") as fh:
            fh.write("OK\n" if not bad else f"MISMATCH {bad}\n")
            fh.write("\n".join(notes) + "\n")
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        with open(os.path.join(result_dir, f"rank{rank}.txt"), "w") as fh:
            fh.write("EXCEPTION\n" + traceback.format_exc())
        raise


def run_plan(config: str, n: int, world: int) -> list:
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with tempfile.TemporaryDirectory() as rd:
        try:
            mp.spawn(_plan_worker, args=(world, port, config, n, rd), nprocs=world, join=True)
        finally:
            lines = []
            for r in range(world):
                path = os.path.join(rd, f"rank{r}.txt")
                lines.append(open(path).read() if os.path.exists(path) else "NO RESULT")
    return lines


def run(config: str, n: int, world: int, mode: str, backend: str) -> list:
    """Spawn `world` ranks; returns the per-rank result lines."""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with tempfile.TemporaryDirectory() as rd:
        try:
            mp.spawn(_worker, args=(world, port, config, n, mode, backend, rd), nprocs=world, join=True)
        finally:
            lines = []
            for r in range(world):
                path = os.path.join(rd, f"rank{r}.txt")
                lines.append(open(path).read() if os.path.exists(path) else "NO RESULT")
    return lines


if __name__ == "__main__":
    cfg, n, world, mode, backend = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5]
    for r, text in enumerate(run(cfg, n, world, mode, backend)):
        print(f"rank {r}: {text.strip()}")
