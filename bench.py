#!/usr/bin/env python3
"""bench.py -- particle-updates/s of the SPH right-hand side on N B200s.

One "step" = one evaluation of the hot path (what miluphcuda's rightHandSide() does,
reference src/rhs.cu:143-861) over every particle of a synthetic scenario; the metric
is RHS evaluations x particles / second (BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload impact] [--particles 1000000] [--state evolved|step0]

Default workload: the 10^6-particle solid impact per GPU (BASELINE.json: the >= 5x target is defined on it; 16M on 8
GPUs = 2M per GPU with --particles 2000000).  Default state on one GPU: EVOLVED -- the scenario after >= 20 accepted
rk2_adaptive steps (SURVEY 8d's second measurement point), not the pristine step-0 lattice (zero velocities, zero
stress, no active flaws).  The evolved input is prepared, before anything is timed, by the reference's own integrator
(oracle/_ref/miluphcuda_<workload> with REF_EVOLVE: the same state the reference arm times); the step-0 state is
measured as well and reported under "step0".  Without that binary, and on several GPUs (the reference is single-GPU
and cannot evolve a 16M-particle set), the step-0 state is the timed one and `config.state` says so.
  python bench.py --impl reference ...      # the reference's own implementation, same metric/config

Arms
  default   : libb200sph_<workload>.so through the C-ABI.  `value` = inputs resident in HBM,
              per-step CUDA events on the launching stream, L2 flushed between steps (untimed);
              `e2e` = the same call with HOST (pinned) buffers, copies inside the timed region.
  reference : miluphcuda has no CPU path (north_star), so the reference arm runs the reference's
              own CUDA build (oracle/_ref/miluphcuda_<workload>, compiled unmodified from
              /root/reference for sm_100a) on the same GPU, one rightHandSide() per step.
              If that binary is missing, the CPU oracle port is timed instead (kind "port").
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

METRIC = "sph_particle_updates_per_s"
UNIT = "particle-updates/s"
DEFAULT_PARTICLES = {"shocktube": 10000, "sedov": 1000000, "rings": 1000000, "impact": 1000000,
                     "giant_hydro": 1000000, "giant_solid": 1000000, "nakamura": 1000000}


# ----------------------------------------------------------------------------- roofline constants
def load_constants() -> dict:
    with open(os.path.join(REPO, "miluphcuda_b200", "roofline_constants.json")) as fh:
        return json.load(fh)


def measured_peaks() -> dict:
    peaks = {"hbm_gbs": 6650.0, "hbm_source": "fallback (B200_PROFILING.md)", "fp64_tflops": 34.1,
             "fp64_source": "measured, tools/fp64_peak.cu on this pool (profiles/r01_fp64_peak.json)"}
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            mp = json.load(fh)
        if "hbm_gbs" in mp:
            peaks["hbm_gbs"] = float(mp["hbm_gbs"])
            peaks["hbm_source"] = "measured (MEASURED_PEAKS.json)"
    fp = os.path.join(REPO, "profiles", "r01_fp64_peak.json")
    if os.path.exists(fp):
        with open(fp) as fh:
            peaks["fp64_tflops"] = float(json.load(fh)["fp64_tflops_sustained"])
    return peaks


# ----------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[5:9]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# the reference integrator prepares the evolved state before anything is timed; a workload it cannot advance in this
# time (self-gravitating sets at 10^6 particles) is timed in its step-0 state by BOTH arms, and the line says so
EVOLVE_TIMEOUT_S = float(os.environ.get("B200SPH_EVOLVE_TIMEOUT_S", "240"))


def workload_label(workload: str, n_global: int, n_per_gpu: int, state: str, accepted) -> str:
    """Identical for both arms when they ran the same particle set in the same state."""
    what = (f"evolved: after {accepted} accepted rk2_adaptive steps of the reference integrator" if state == "evolved"
            else "step-0 state")
    return f"{workload} (synthetic, {n_global} particles, {what})"


# ----------------------------------------------------------------------------- reference arm
def run_reference_arm(args, rank: int, world: int) -> None:
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    workload, n = args.workload, args.particles
    binary = os.path.join(REPO, "oracle", "_ref", f"miluphcuda_{workload}")
    base = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{workload} (synthetic, {n} particles requested)", "particles": n}}
    if os.path.exists(binary):
        import make_golden
        # the same state rule as the library arm: evolved on one GPU, step 0 when the library arm is distributed
        evolve = (args.state or ("evolved" if world == 1 else "step0")) == "evolved"
        try:
            res = make_golden.time_reference(workload, n, calls=args.steps, warmup=args.warmup, evolve=evolve,
                                             timeout_s=EVOLVE_TIMEOUT_S if evolve else None)
        except make_golden.ReferenceTimeout as exc:
            # same rule as the library arm: both arms then time the step-0 state
            evolve = False
            base["config"]["state_note"] = f"{exc}: step-0 state instead"
            res = make_golden.time_reference(workload, n, calls=args.steps, warmup=args.warmup, evolve=False)
        value = res["updates_per_s"]
        base["config"]["particles"] = res["n"]
        base["config"]["workload"] = workload_label(workload, res["n"], res["n"], "evolved" if evolve else "step0",
                                                    (res.get("evolved_steps") or {}).get("accepted"))
        base["config"]["state"] = "evolved" if evolve else "step0"
        base.update({"value": value, "ms_per_step": res["ms_per_call"],
                     "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "reference",
                                      "sample": ("miluphcuda has no CPU path: the UNMODIFIED reference CUDA build (sm_100a, quiet debug flags) "
                                                 f"ran {args.steps} rightHandSide() calls on ONE B200 after {args.warmup} warm-up calls; 1 host thread; "
                                                 "a single GPU regardless of --gpus (the reference is single-GPU)")},
                     "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                     "reference_kernel_ms": res["kernel_ms_last"]})
    else:
        value, cores, sample, ms = time_oracle_port(workload, min(n, 200000), budget_s=20.0)
        base.update({"value": value, "ms_per_step": ms,
                     "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
                     "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(base), flush=True)


def time_oracle_port(workload: str, n: int, budget_s: float = 15.0):
    """CPU restatement (oracle/sph_oracle.c, OpenMP) on a bounded sample of the workload."""
    sys.path.insert(0, os.path.join(REPO, "tests"))
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import common  # tests/common.py: oracle binding
    from miluphcuda_b200 import api, scenarios, state
    sc = scenarios.make(workload, n)
    with tempfile.TemporaryDirectory() as td:
        cfg = state.write_material_files(sc, td)
        mats = api.MaterialTables(workload, cfg)
        arrays, meta = state.scenario_arrays(sc, mats)
        cores = os.cpu_count() or 1
        common.oracle_rhs(workload, arrays, mats, meta)  # warm-up (page faults, omp pool)
        calls, t0 = 0, time.time()
        while True:
            rc, off, _ = common.oracle_rhs(workload, arrays, mats, meta)
            if rc != 0:
                raise RuntimeError(f"oracle rc={rc}")
            calls += 1
            if time.time() - t0 > budget_s or calls >= 20:
                break
        dt = time.time() - t0
    value = sc.n * calls / dt
    sample = f"oracle/sph_oracle.c (OpenMP, {cores} threads) on {workload} with {sc.n} particles, {calls} RHS calls in {dt:.1f} s"
    return value, cores, sample, dt / calls * 1e3


class NativeDistributedRhs:
    """bench.py's view of the C++/NCCL host (csrc/mg.cu): decompose + migrate once, then one b200sph_mg_rhs_eval per step."""

    def __init__(self, api, torch, dist, workload, eng, dev, capacity, n, meta, mine, rank, world):
        import types
        self.api, self.eng, self.dev, self.capacity, self.meta = api, eng, dev, capacity, meta
        ids = [api.NativeMultiGpu.unique_id(workload) if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        self.mg = api.NativeMultiGpu(eng, rank, world, ids[0])
        self.ids = torch.zeros(capacity, dtype=torch.int32, device="cuda")   # global ids ride along as an extra member
        self.ids[:n] = torch.as_tensor(mine, dtype=torch.int32, device="cuda")
        extra = eng.rk2_buffers([{"depth": self.ids}, {}, {}])
        view = self._view(n)
        self.mg.decompose(view, n)
        self.n_owned = self.mg.migrate(view, n, capacity, extra, 1)
        self.n_total = self.n_owned
        self.external_sums, self.sum_exchanges = True, 0
        self.halo = types.SimpleNamespace(levels=1, SKIN=0.15, plan_builds=0, stale_plans=0, device_verdict=True, last={})

    def _view(self, n):
        m = self.meta
        return self.api.make_view(self.dev, None, n, max_num_flaws=m["max_num_flaws"], selfgravity=m["selfgravity"], theta=m["theta"],
                                  grav_const=self.eng.materials.grav_const)

    def global_ids(self):
        return self.ids[: self.n_owned].cpu().numpy().astype("int64")

    def exchange(self):
        pass   # exchange and evaluation are one C call

    def compute(self):
        self.n_total = self.mg.rhs_eval(self._view(self.n_owned), self.n_owned, self.capacity)
        st = self.mg.stats()
        self.sum_exchanges = st["sum_exchanges"]
        self.halo.plan_builds, self.halo.stale_plans = st["plan_builds"], st["stale_plans"]
        self.halo.last = {"bytes_sent": st["halo_bytes_sent"]}

    def eval(self):
        self.compute()


def _perturb(torch, dev, n, ids, amp_v, amp_e, integrate_density):
    """Deterministic perturbation keyed by the GLOBAL particle id (the same on whichever rank and row a particle lives):
    velocities, energy and density move so that every pair term of the rates is exercised on the pristine lattices."""
    ph = ids.to(torch.float64)
    for k, name in enumerate(("vx", "vy", "vz")):
        if name in dev:
            dev[name][:n] += amp_v * torch.sin(0.37 * (k + 1) * ph + k)
    if "e" in dev:
        dev["e"][:n] += amp_e * (1.0 + torch.sin(0.53 * ph))
    if integrate_density:
        dev["rho"][:n] *= 1.0 + 0.02 * torch.sin(0.29 * ph)


def single_domain_check(api, torch, dist, workload, cfg, full, meta, sc, M, rank, world, local_rank) -> dict:
    """Correctness of the distributed evaluation at the size it was timed at.  Every particle gets a deterministic
    perturbation keyed by its global id (on the pristine step-0 lattices most rates vanish: a comparison of zeros proves
    little); the distributed buffers are evaluated twice more; rank 0 evaluates the WHOLE particle set as one domain on
    its own GPU through the same sequence and sends every rank the rows it owns; each rank compares."""
    import numpy as np
    names_all = ("ax", "ay", "az", "drhodt", "dedt", "dhdt", "dSdt", "dddt", "dalphadt", "rho", "p", "cs", "g_ax", "g_ay", "g_az", "noi")
    n, cap, dev, drhs = M["n"], M["capacity"], M["dev"], M["drhs"]
    names = [f for f in names_all if f in dev]
    n_all = sc.n
    integrate_density = bool(sc.switches().get("INTEGRATE_DENSITY", 0))
    ids = torch.as_tensor(np.asarray(M["mine"]), dtype=torch.int64, device="cuda")
    # every rank's global ids, in its row order, to rank 0
    counts = torch.zeros(world, dtype=torch.int64, device="cuda")
    counts[rank] = n
    dist.all_reduce(counts)
    counts = [int(c) for c in counts.tolist()]
    amps = torch.zeros(2, dtype=torch.float64, device="cuda")
    if rank == 0:
        cs_mean = float(dev["cs"][:n].mean().item())
        amps[0], amps[1] = 0.05 * cs_mean, 0.01 * cs_mean * cs_mean
    dist.broadcast(amps, src=0)
    amp_v, amp_e = float(amps[0].item()), float(amps[1].item())
    _perturb(torch, dev, n, ids, amp_v, amp_e, integrate_density)
    for _ in range(2):
        drhs.eval()
    torch.cuda.synchronize()
    want = {}
    if rank == 0:
        all_ids = [ids] + [torch.empty(counts[r], dtype=torch.int64, device="cuda") for r in range(1, world)]
        for r in range(1, world):
            dist.recv(all_ids[r], src=r)
        eng1 = api.RhsEngine(workload, n_max=n_all, device=local_rank, material_cfg=cfg)
        eng1.set_stream(torch.cuda.current_stream().cuda_stream)
        dev1 = {k: torch.from_numpy(v).cuda() for k, v in full.items()}
        view1 = api.make_view(dev1, None, n_all, max_num_flaws=meta["max_num_flaws"], selfgravity=meta["selfgravity"], theta=meta["theta"],
                              grav_const=eng1.materials.grav_const)
        for _ in range(3):   # like the timed buffers: evaluated repeatedly, so that c_s has seen its own pressure
            eng1.rhs_eval(view1)
        _perturb(torch, dev1, n_all, torch.arange(n_all, dtype=torch.int64, device="cuda"), amp_v, amp_e, integrate_density)
        for _ in range(2):
            eng1.rhs_eval(view1)
        torch.cuda.synchronize()
        sc_t = torch.tensor([float(torch.sqrt(torch.mean(dev1[f].double() ** 2)).item()) for f in names], dtype=torch.float64, device="cuda")
        for r in range(world):
            for f in names:
                per = dev1[f].numel() // n_all
                rows = dev1[f].view(n_all, per)[all_ids[r]].double().contiguous()
                if r == 0:
                    want[f] = rows
                else:
                    dist.send(rows, dst=r)
        eng1.close()
        del dev1
    else:
        dist.send(ids, dst=0)
        sc_t = torch.empty(len(names), dtype=torch.float64, device="cuda")
        for f in names:
            per = dev[f].numel() // cap
            buf = torch.empty((n, per), dtype=torch.float64, device="cuda")
            dist.recv(buf, src=0)
            want[f] = buf
    dist.broadcast(sc_t, src=0)
    worst, worst_name, nonzero = 0.0, None, 0
    for k, f in enumerate(names):
        if f == "noi":
            continue
        per = dev[f].numel() // cap
        got = dev[f].view(cap, per)[:n].double()
        denom = torch.clamp(want[f].abs(), min=float(sc_t[k].item()))
        denom = torch.where(denom > 0, denom, torch.ones_like(denom))
        err = float(((got - want[f]).abs() / denom).max().item())
        nonzero += int(float(sc_t[k].item()) > 0.0)
        if err > worst:
            worst, worst_name = err, f
    noi_bad = int((dev["noi"][:n].double() != want["noi"].view(-1)).sum().item())
    torch.cuda.empty_cache()
    return {"what": ("owned particles of every rank vs a single-domain evaluation of all %d particles (rank 0's GPU), on a state "
                     "perturbed per global particle id so that every rate is non-zero" % n_all),
            "fields": [f for f in names if f != "noi"], "fields_nonzero": nonzero, "max_rel_err": worst, "worst_field": worst_name,
            "noi_mismatches": noi_bad, "tolerance": 1e-9}


# ----------------------------------------------------------------------------- our arm
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="impact", choices=list(DEFAULT_PARTICLES))
    ap.add_argument("--state", default=None, choices=["evolved", "step0"], help="default: evolved on one GPU, step0 on several")
    ap.add_argument("--particles", type=int, default=None, help="particles per GPU (weak scaling)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--mg-host", default="native", choices=["python", "native"],
                    help="several GPUs: host layer of the exchange -- miluphcuda_b200/multigpu.py over torch.distributed, or the C++/NCCL "
                         "host behind the C-ABI (csrc/mg.cu, b200sph_mg_*)")
    ap.add_argument("--halo-headroom", type=float, default=1.6, help="capacity of a rank's buffers over its owned particles")
    ap.add_argument("--no-parity-check", action="store_true", help="several GPUs: skip the single-domain comparison after the timed region")
    ap.add_argument("--no-reorder", action="store_true", help="keep the generator's particle order (no b200sph_reorder)")
    args = ap.parse_args()
    if args.particles is None:
        args.particles = DEFAULT_PARTICLES[args.workload]
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from miluphcuda_b200 import api, scenarios, state

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the SPH right-hand side has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from miluphcuda_b200 import multigpu

    # Weak scaling: `--particles` per GPU.  ONE particle set of world x particles is generated (identically on
    # every rank, the generators are deterministic) and cut along a Morton curve; every rank keeps its piece.
    workload = args.workload
    sc = scenarios.make(workload, args.particles * world)
    tmp = tempfile.TemporaryDirectory()
    cfg = state.write_material_files(sc, tmp.name)
    mats = api.MaterialTables(workload, cfg)
    full_step0, meta = state.scenario_arrays(sc, mats)
    n_global = sc.n

    # ---- input preparation (untimed): the evolved state, made by the reference's own integrator (see module docstring)
    state_kind = args.state or ("evolved" if world == 1 else "step0")
    state_note, evolved_steps, full_evolved = None, None, None
    if state_kind == "evolved":
        binary = os.path.join(REPO, "oracle", "_ref", f"miluphcuda_{workload}")
        if world > 1:
            state_kind, state_note = "step0", "the reference integrator that prepares the evolved state is single-GPU"
        elif not os.path.exists(binary):
            state_kind, state_note = "step0", f"{os.path.relpath(binary, REPO)} is not built: no integrator to evolve the state with"
        else:
            sys.path.insert(0, os.path.join(REPO, "oracle"))
            import make_golden
            try:
                d_in, evolved_steps = make_golden.evolved_state(sc, tmp.name, timeout_s=EVOLVE_TIMEOUT_S)
                full_evolved, _ = make_golden.arrays_from_dump(workload, d_in, bool(sc.selfgravity))
                del d_in
            except make_golden.ReferenceTimeout as exc:
                state_kind, state_note = "step0", f"{exc}: step-0 state instead"

    flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")  # 512 MiB > 126 MB L2
    stream = torch.cuda.current_stream()
    engines = {}

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def measure(full, steps, with_clocks):
        """Warm-up + exactly `steps` timed evaluations of one particle set; device times by CUDA events, max over ranks."""
        arrays, n, capacity, mine, dec = multigpu.scatter_scenario(full, sc.n, sc.dim, meta["max_num_flaws"], rank, world,
                                                                    headroom=args.halo_headroom)
        if capacity not in engines:
            engines[capacity] = api.RhsEngine(workload, n_max=capacity, device=local_rank, material_cfg=cfg)
            engines[capacity].set_stream(stream.cuda_stream)
        eng = engines[capacity]
        dev = {k: torch.from_numpy(v).cuda() for k, v in arrays.items()}
        native = world > 1 and args.mg_host == "native"
        drhs = None if native else multigpu.DistributedRhs(eng, dev, capacity, n, dec, meta, sc.switches())
        if not args.no_reorder:
            # persistent cell order: the product's own b200sph_reorder(), once, before anything is timed
            # -- what an integrator does every few hundred steps (on several GPUs: every rank for the particles it owns);
            # the host copies follow so that e2e uploads that order
            view0 = api.make_view(dev, None, n, max_num_flaws=meta["max_num_flaws"], selfgravity=meta["selfgravity"],
                                  theta=meta["theta"], grav_const=eng.materials.grav_const)
            perm = torch.empty(n, dtype=torch.int32, device="cuda")
            eng.set_owned(0)
            eng.reorder(view0, perm_out=perm)
            mine = np.asarray(mine)[perm.cpu().numpy()]
            arrays = {k: v.cpu().numpy() for k, v in dev.items()}
        if native:
            drhs = NativeDistributedRhs(api, torch, dist, workload, eng, dev, capacity, n, meta, mine, rank, world)
            n, mine = drhs.n_owned, drhs.global_ids()
            arrays = {k: v.cpu().numpy() for k, v in dev.items()}   # the rows a rank holds changed with the migration
        for _ in range(args.warmup):
            drhs.eval()
        barrier()
        sampler = ClockSampler(local_rank) if with_clocks else None
        if sampler:
            sampler.start()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
              for _ in range(steps)]
        stage_ms, launches = {}, 0
        t_wall0 = time.time()
        for k in range(steps):
            flush.fill_(float(k))           # evict L2 between timed steps (not timed)
            ev[k][0].record(stream)
            drhs.exchange()                 # halo exchange (+ gravity sources) over NCCL; a no-op on one GPU
            ev[k][2].record(stream)
            drhs.compute()                  # b200sph_rhs_eval on owned + halo particles
            ev[k][1].record(stream)
            st = eng.stats()
            launches += st["kernel_launches"]
            for key, val in st.items():
                if key.startswith("ms_"):
                    stage_ms[key] = stage_ms.get(key, 0.0) + val
        barrier()
        t_wall = time.time() - t_wall0
        clocks = sampler.stop() if sampler else None
        dev_ms = sum(a.elapsed_time(b) for a, b, _ in ev)
        exch_ms = sum(a.elapsed_time(c) for a, _, c in ev)
        # max over ranks of the device time for exactly K steps
        t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
        cnt = torch.tensor([float(n)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        return dict(arrays=arrays, n=n, capacity=capacity, dec=dec, eng=eng, dev=dev, drhs=drhs, stage_ms=stage_ms, launches=launches, mine=mine,
                    t_wall=t_wall, clocks=clocks, exch_ms=exch_ms, total_ms=float(t.item()), total_particles=float(cnt.item()),
                    stats=eng.stats())

    step0 = None
    if full_evolved is not None:
        # second measurement point of SURVEY 8d: the pristine step-0 state, reported next to the headline
        m0 = measure(full_step0, max(3, min(args.steps, 10)), with_clocks=False)
        steps0 = max(3, min(args.steps, 10))
        step0 = {"value": m0["total_particles"] * steps0 / (m0["total_ms"] * 1e-3), "unit": UNIT, "ms_per_step": m0["total_ms"] / steps0,
                 "steps": steps0, "mean_interactions": float(m0["dev"]["noi"][: m0["n"]].sum().item()) / m0["n"],
                 "stage_ms_per_step": {k: v / steps0 for k, v in m0["stage_ms"].items()}}
        del m0
        torch.cuda.empty_cache()
    full_timed = full_evolved if full_evolved is not None else full_step0
    M = measure(full_timed, args.steps, with_clocks=True)
    if rank != 0:
        full_timed = full_step0 = full_evolved = None   # only rank 0 keeps the whole set (for the parity check)

    # ---- several GPUs: the distributed answer against a SINGLE-DOMAIN evaluation of the whole particle set on this
    # rank's own GPU (after the timed region): neighbour counts of the owned particles equal, every rate within 1e-9
    parity = None
    if world > 1 and not args.no_parity_check:
        parity = single_domain_check(api, torch, dist, workload, cfg, full_timed if rank == 0 else None, meta, sc, M,
                                     rank, world, local_rank)
        worst = torch.tensor([parity["max_rel_err"], float(parity["noi_mismatches"])], dtype=torch.float64, device="cuda")
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        parity["max_rel_err"], parity["noi_mismatches"] = float(worst[0].item()), int(worst[1].item())
        parity["ok"] = bool(parity["max_rel_err"] <= 1e-9 and parity["noi_mismatches"] == 0)
    del full_step0, full_evolved, full_timed
    arrays, n, capacity, eng, dev, drhs = M["arrays"], M["n"], M["capacity"], M["eng"], M["dev"], M["drhs"]
    stage_ms, launches, t_wall, clocks, exch_ms = M["stage_ms"], M["launches"], M["t_wall"], M["clocks"], M["exch_ms"]
    total_ms, total_particles, stats = M["total_ms"], M["total_particles"], M["stats"]
    # per-rank picture (owned, halo, ms in b200sph_rhs_eval, ms in the exchange incl. waiting for peers): shows imbalance
    mine = torch.tensor([float(n), float(drhs.n_total - n), stage_ms.get("ms_total", 0.0) / args.steps, exch_ms / args.steps],
                        dtype=torch.float64, device="cuda")
    per_rank = [mine.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, mine)
    per_rank = [[round(float(x), 4) for x in row.tolist()] for row in per_rank]
    value = total_particles * args.steps / (total_ms * 1e-3)

    # ---- end to end: host (pinned) buffers, copies inside the timed region.
    # One GPU: the C-ABI's own host entry point (b200sph_rhs_eval_host).  Several GPUs: each rank uploads the
    # state of its owned particles, runs exchange + evaluation, and reads every rate/state output back.
    e2e = None
    if not args.no_e2e:
        pinned = {k: torch.from_numpy(v).pin_memory() for k, v in arrays.items()}
        if world == 1:
            eng.set_owned(0)
            # flaws, h0, m, materialId, numFlaws go up once (the reference uploads them once per run too,
            # src/memory_handling.cu:111-122,373-392); sigma/R/C/plastic_f scratch is not read back
            eng.host_options(eng.HOST_CACHE_IMMUTABLES | eng.HOST_SKIP_SCRATCH)
            hview = api.make_view(pinned, None, n, max_num_flaws=meta["max_num_flaws"], selfgravity=meta["selfgravity"],
                                  theta=meta["theta"], grav_const=eng.materials.grav_const)

            def e2e_step():
                return eng.rhs_eval_host(hview)
        else:
            p_fields, rhs_fields = api.fields_for(sc.switches(), meta["selfgravity"])
            # same contract as the one-GPU host call with CACHE_IMMUTABLES|SKIP_SCRATCH: the integrated state goes up
            # every step, immutables (m, h0, materialId, numFlaws, flaws) are resident, p_rhs scratch is not read back
            immutable = ("m", "h0", "materialId", "flaws", "numFlaws")
            scratch = ("sigma", "R", "plastic_f", "tensorialCorrectionMatrix")
            outputs = [f for f in p_fields + rhs_fields if f in dev and f not in ("x", "y", "z", "vx", "vy", "vz") + immutable + scratch]
            inputs = [f for f in dev if (f in multigpu.HALO_STATE_FIELDS or f in ("numActiveFlaws", "pold")) and f not in immutable]

            def e2e_step():
                nb_in = nb_out = 0
                for f in inputs:
                    per = dev[f].numel() // capacity
                    dev[f][: n * per].copy_(pinned[f][: n * per], non_blocking=True)
                    nb_in += n * per * dev[f].element_size()
                drhs.eval()
                for f in outputs:
                    per = dev[f].numel() // capacity
                    pinned[f][: n * per].copy_(dev[f][: n * per], non_blocking=True)
                    nb_out += n * per * dev[f].element_size()
                torch.cuda.synchronize()
                return nb_in, nb_out
        h2d = d2h = 0
        for _ in range(2):
            h2d, d2h = e2e_step()
        barrier()
        e_steps = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(e_steps):
            h2d, d2h = e2e_step()
        barrier()
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": total_particles * e_steps / float(te.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e_steps,
               "timing": "host wall clock around the host-buffer call (sync on both sides), max over ranks",
               "api": ("b200sph_rhs_eval_host, options CACHE_IMMUTABLES|SKIP_SCRATCH: per-step inputs are the integrated state, "
                       "immutables (m, h0, materialId, flaws) uploaded once; copies overlapped with the kernels on a second stream")
               if world == 1 else "pinned->device copies of the integrated state + DistributedRhs.eval + device->pinned copies of every rate/state output (immutables resident, p_rhs scratch not read back)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (the pair-force loop), timed live by CUDA events on its stream
    const = load_constants()
    peaks = measured_peaks()
    kind = const["workloads"][workload]
    total_noi = float(dev["noi"][:n].sum().item())
    pairs = total_noi
    ms_forces = stage_ms.get("ms_forces", 0.0) / args.steps
    flops = kind["force_flop_per_pair"] * pairs + kind["force_flop_per_particle"] * n
    bytes_alg = kind["force_bytes_per_particle"] * n
    tf = flops / (ms_forces * 1e-3) / 1e12 if ms_forces > 0 else 0.0
    gbs = bytes_alg / (ms_forces * 1e-3) / 1e9 if ms_forces > 0 else 0.0
    # DRAM bytes of one k_forces launch from the committed `ncu --set full` capture of this workload (None if none exists)
    traffic = None
    traffic_source = None
    for tname in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):  # newest capture that holds this workload
        tpath = os.path.join(REPO, "profiles", tname)
        if traffic is None and os.path.exists(tpath):
            with open(tpath) as fh:
                traffic = (json.load(fh).get(workload) or {}).get("k_forces_dram_bytes_per_launch")
            traffic_source = "profiles/" + tname if traffic is not None else None
    # busiest unit of the kernel in the committed full capture (the gather's L1TEX data pipe), reported beside the FP64 fraction
    l1_pipe = None
    lpath = os.path.join(REPO, "profiles", "r02_ncu_l1_pipe.json")
    if os.path.exists(lpath):
        with open(lpath) as fh:
            lp = json.load(fh)
        if workload in lp:
            l1_pipe = {"frac_of_peak_wavefront_rate": lp[workload].get("k_forces"), "fp64_pipe_busy": lp[workload].get("k_forces_fp64_pipe"),
                       "issue_active": lp[workload].get("k_forces_issue_active"), "source": "profiles/r02_ncu_l1_pipe.json (ncu --set full capture, not this run)"}
    roof = {"bound": "fp64", "kernel": "k_forces", "achieved": tf, "peak": peaks["fp64_tflops"], "unit": "TFLOP/s",
            "frac": tf / peaks["fp64_tflops"], "traffic": traffic, "traffic_source": traffic_source, "peak_source": peaks["fp64_source"],
            "ms_per_launch": ms_forces, "pairs_per_launch": pairs, "flop_per_pair": kind["force_flop_per_pair"],
            "hbm": {"achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                    "peak_source": peaks["hbm_source"], "bytes_per_particle": kind["force_bytes_per_particle"]},
            "l1tex_data_pipe": l1_pipe,
            "share_of_step": ms_forces * args.steps / total_ms if total_ms > 0 else None,
            "stage_ms_per_step": {k: v / args.steps for k, v in stage_ms.items()}}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        v_cpu, cores, sample, _ = time_oracle_port(workload, min(n, 250000), budget_s=12.0)
        cpu = {"value": v_cpu, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_label(workload, n_global, n_global // world, state_kind, evolved_steps),
                   "state": state_kind, "state_note": state_note,
                   "particle_order": ("generator order (as the reference arm)" if args.no_reorder else
                                      "search-cell order: b200sph_reorder() applied once before the timed region (SURVEY 8f row 2); "
                                      "the reference keeps the input file's order"), "particles_per_gpu": n_global // world, "particles": int(total_particles),
                   "mean_interactions": total_noi / n, "l2": "512 MiB buffer written between timed steps (untimed)",
                   "timing": "per-step CUDA events on the launching stream, summed over K steps, max over ranks",
                   "multi_gpu_host": (("C++ over NCCL behind the C-ABI (csrc/mg.cu: b200sph_mg_decompose / migrate / rhs_eval)"
                                       if args.mg_host == "native" else "Python over torch.distributed (miluphcuda_b200/multigpu.py)")
                                      if world > 1 else None),
                   "multi_gpu": ("Morton-curve domain decomposition, %d-level state halo per evaluation (NCCL all_to_all of the packed "
                                 "state)%s; the send plan is reused while no particle moved > %.2f h_min, its verdict %s: "
                                 "%d plan builds, %d stale plans in this run%s"
                                 % (drhs.halo.levels,
                                    " + neighbour-sum exchange (owners deliver density / correction matrix of the copies between the "
                                    "stages of the evaluation: %d per evaluation)" % (drhs.sum_exchanges // max(1, args.steps + args.warmup))
                                    if drhs.external_sums else "",
                                    drhs.halo.SKIN,
                                    "stays on the device as the evaluation's abort flag" if drhs.halo.device_verdict else "is awaited by the host",
                                    drhs.halo.plan_builds, drhs.halo.stale_plans,
                                    ", replicated gravity tree (NCCL all_gather of x,y,z,m)" if meta["selfgravity"] else ""))
                   if world > 1 else "single",
                   "rank0": {"owned": n, "halo": drhs.n_total - n, "halo_bytes_sent": drhs.halo.last.get("bytes_sent", 0),
                             "exchange_ms_per_step": exch_ms / args.steps},
                   "ranks": {"columns": ["owned", "halo", "rhs_ms", "exchange_ms"], "rows": per_rank}},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu, "step0": step0,
        "parity": parity,
        "wall_ms_per_step_incl_flush": t_wall / args.steps * 1e3,
        "search_grid": {"cells": stats["n_cells"], "cell_size": stats["cell_size"], "max_interactions": stats["max_noi"]},
    }
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
