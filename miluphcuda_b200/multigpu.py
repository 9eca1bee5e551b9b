"""Multi-GPU host layer of the SPH right-hand side: domain decomposition and halo exchange.

The reference is single-GPU ("Parallelism strategies"); this layer is new
(SURVEY section 8e).  One process per GPU; `torch.distributed` carries the collectives (NCCL on
GPUs, gloo in the CPU tests), torch tensors carry the buffers.  Nothing here computes physics:
per evaluation it decides which owned particles other ranks need, moves their state, and hands
`n_owned + n_halo` particles to `b200sph_rhs_eval`, which produces rates for the owned ones.

Decomposition
    The global bounding cube is cut into octree cells of a fixed level, numbered along the Morton
    (Z-order) curve; every rank owns a contiguous range of cells chosen on the global per-cell
    histogram so that the particle counts are equal (`MortonDecomposition`).  A rank's domain is then
    exactly a small union of aligned octree boxes.

Halo
    Rank r needs every foreign particle that can be a neighbour of one of its particles
    (|x_i - x_j| < min(h_i, h_j), reference src/tree.cu:851-865), and -- because density (kernel
    sum) and the tensorial correction matrix of those neighbours are themselves neighbour sums
    (src/density.cu:41-209, src/kernel.cu:585-713) -- the neighbours of those neighbours.  A
    particle k is therefore sent to rank r when its distance to one of r's boxes is below
    h_k + h_max(r) ("two levels"; h_max(r) = largest smoothing length on rank r), or below h_k when the
    switch set has neither neighbour sum ("one level").  Pointwise quantities (pressure, sound speed, stress, plasticity) are recomputed
    on the copies, so ONE exchange per evaluation suffices.

Gravity
    Self-gravity uses a replicated tree: x, y, z, m of all particles are all-gathered
    (32 bytes per particle) and every rank builds the same reference cells from them
    (`b200sph_set_gravity_sources`), walking them for its own particles only.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import api

# state a neighbour contributes through (inputs of the pointwise chain and of the pair loops);
# everything else on a halo copy is an output nobody reads
HALO_STATE_FIELDS = (
    "x", "y", "z", "vx", "vy", "vz", "m", "h", "h0", "rho", "e", "p", "cs", "materialId",
    "S", "d", "damage_porjutzi", "alpha_jutzi",
    # the pointwise chain recomputes the damage limit (numActiveFlaws / numFlaws)^(1/DIM) on the copies, exactly as the
    # owner does (src/damage.cu:33-82); the flaw thresholds themselves stay home (only owners run the flaw scan)
    "numFlaws", "numActiveFlaws",
)
# neighbour sums of the copies that their owners deliver between the stages of an evaluation
SUM_FIELDS = {1: ("rho",), 2: ("tensorialCorrectionMatrix",)}


class MortonDecomposition:
    """Morton-key domain decomposition with cuts aligned to the cells of octree level `level`.

    The global bounding cube is divided into 2^level cells per axis; cells are numbered along the Z-order
    curve and rank r owns the contiguous cell range [cuts[r], cuts[r+1]).  The cuts are chosen on the
    histogram of particles per cell so that the ranks hold (nearly) equal counts -- the imbalance is at
    most one cell's population.  Because cuts fall on cell boundaries, a rank's domain is exactly a union
    of at most 2*(2^dim - 1)*level aligned octree boxes (`boxes()`), which is what the halo test needs.
    """

    DEFAULT_LEVEL = {1: 15, 2: 9, 3: 6}

    def __init__(self, dim: int, lo, hi, world: int, level: int | None = None):
        self.dim = dim
        self.world = world
        self.level = level if level is not None else self.DEFAULT_LEVEL[dim]
        lo = np.asarray(lo, dtype=np.float64)[:dim]
        hi = np.asarray(hi, dtype=np.float64)[:dim]
        span = float((hi - lo).max())
        span = span * (1.0 + 1e-12) if span > 0 else 1.0
        centre = 0.5 * (lo + hi)
        self.lo = centre - 0.5 * span          # the bounding CUBE, so cells are cubes
        self.span = span
        self.n_cells = 1 << (dim * self.level)
        self.cuts = None                       # world + 1 cell ids

    # -- cell ids ---------------------------------------------------------------------------------
    def cell_ids(self, pos):
        """Z-order id of the level-`level` cell of every position; pos[n, dim] is a numpy array or a torch tensor."""
        g = 1 << self.level
        if isinstance(pos, np.ndarray):
            q = np.clip(((pos - self.lo) / self.span * g).astype(np.int64), 0, g - 1)
            ids = np.zeros(len(pos), dtype=np.int64)
        else:
            lo = torch.as_tensor(self.lo, dtype=pos.dtype, device=pos.device)
            q = ((pos - lo) / self.span * g).to(torch.int64).clamp_(0, g - 1)
            ids = torch.zeros(pos.shape[0], dtype=torch.int64, device=pos.device)
        for b in range(self.level):
            for a in range(self.dim):
                ids |= ((q[:, a] >> b) & 1) << (self.dim * b + a)
        return ids

    # -- cuts -------------------------------------------------------------------------------------
    def set_cuts_from_histogram(self, hist) -> None:
        """hist[c] = global particle count of cell c.  Rank r gets cells [cuts[r], cuts[r+1])."""
        hist = np.asarray(hist, dtype=np.int64)
        csum = np.concatenate([[0], np.cumsum(hist)])
        total = int(csum[-1])
        cuts = [0]
        for r in range(1, self.world):
            target = total * r / self.world
            c = int(np.searchsorted(csum, target, side="left"))
            # csum[c-1] < target <= csum[c]: pick the nearer boundary
            if c > 0 and abs(csum[c - 1] - target) <= abs(csum[c] - target):
                c -= 1
            cuts.append(min(max(c, cuts[-1]), self.n_cells))
        cuts.append(self.n_cells)
        self.cuts = np.asarray(cuts, dtype=np.int64)

    def owner_of(self, ids):
        """Rank owning each cell id (numpy or torch)."""
        if isinstance(ids, np.ndarray):
            return np.searchsorted(self.cuts, ids, side="right") - 1
        cuts = torch.as_tensor(self.cuts, device=ids.device)
        return torch.searchsorted(cuts, ids, right=True) - 1

    # -- geometry ---------------------------------------------------------------------------------
    def _cell_box(self, cell: int, level: int):
        """(lo[3], hi[3]) of octree cell `cell` (Z-order id at `level`)."""
        q = [0] * self.dim
        for b in range(level):
            for a in range(self.dim):
                q[a] |= ((cell >> (self.dim * b + a)) & 1) << b
        size = self.span / (1 << level)
        lo = [self.lo[a] + q[a] * size for a in range(self.dim)] + [0.0] * (3 - self.dim)
        hi = [self.lo[a] + (q[a] + 1) * size for a in range(self.dim)] + [0.0] * (3 - self.dim)
        return lo, hi

    def boxes(self, rank: int) -> np.ndarray:
        """Aligned octree boxes [nb, 6] whose union is exactly rank `rank`'s cell range."""
        c0, c1 = int(self.cuts[rank]), int(self.cuts[rank + 1])
        out = []
        fan = 1 << self.dim
        c = c0
        while c < c1:
            # largest aligned block starting at c that fits into [c, c1)
            up = 0
            while up < self.level and c % (fan ** (up + 1)) == 0 and c + fan ** (up + 1) <= c1:
                up += 1
            lo, hi = self._cell_box(c // (fan ** up), self.level - up)
            out.append(lo + hi)
            c += fan ** up
        return np.asarray(out, dtype=np.float64).reshape(-1, 6)

    def all_boxes(self):
        """(boxes[nb, 6], box_rank[nb]) of every rank."""
        boxes, ranks = [], []
        for r in range(self.world):
            b = self.boxes(r)
            boxes.append(b)
            ranks.append(np.full(len(b), r, dtype=np.int32))
        return np.concatenate(boxes), np.concatenate(ranks)


def morton_partition(x: np.ndarray, world: int, level: int | None = None):
    """Decompose a full (host) particle set: returns (decomposition, [index array per rank])."""
    n, dim = x.shape
    dec = MortonDecomposition(dim, x.min(axis=0), x.max(axis=0), world, level)
    ids = dec.cell_ids(x)
    dec.set_cuts_from_histogram(np.bincount(ids, minlength=dec.n_cells))
    owner = dec.owner_of(ids)
    return dec, [np.nonzero(owner == r)[0] for r in range(world)]


def halo_levels(switches: dict, kernel_sum_materials: bool = False) -> int:
    """2 when neighbours' own neighbour sums are needed (kernel-sum density or tensorial correction), else 1.

    Kernel-sum density TOGETHER with the tensorial correction (an INTEGRATE_DENSITY build whose material.cfg asks for
    density_via_kernel_sum, or a build without INTEGRATE_DENSITY) would need a third level in the state-only halo: the
    correction matrix of a first-level copy sums m/rho over second-level copies whose kernel-sum rho is incomplete.
    That combination is only served by the neighbour-sum exchange (DistributedRhs(external_sums=True) on GPUs, one level);
    the state-only halo refuses it instead of returning wrong matrices."""
    kernel_sum = (not switches.get("INTEGRATE_DENSITY", 0)) or kernel_sum_materials
    if kernel_sum and switches.get("TENSORIAL_CORRECTION", 0):
        return 3
    if kernel_sum or switches.get("TENSORIAL_CORRECTION", 0):
        return 2
    return 1


class HaloExchange:
    """Per-evaluation halo exchange over fixed-capacity particle buffers.

    `fields` maps member names of the reference's `struct Particle` to flat tensors with room for
    `capacity` particles (tensors: capacity*DIM*DIM, flaws: capacity*max_flaws); rows [0, n_owned)
    are this rank's particles, rows behind them receive the halo copies.  `engine` (an api.RhsEngine)
    supplies the selection kernel when the buffers live on a GPU; on CPU tensors (the gloo tests) the
    same test is evaluated with torch ops.
    """

    H_GROWTH = 0.02   # growth of a smoothing length a plan tolerates (only when h is integrated: VARIABLE_SML / INTEGRATE_SML)
    SKIN = 0.15       # a plan tolerates every particle moving this fraction of the smallest smoothing length

    def __init__(self, fields: dict, capacity: int, dec: MortonDecomposition, levels: int = 2, group=None, engine=None,
                 h_evolves: bool = True, reuse_plan: bool = True):
        self.fields = fields
        self.capacity = capacity
        self.dec = dec
        self.dim = dec.dim
        self.levels = levels
        self.group = group
        self.engine = engine
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.axes = ["x", "y", "z"][: self.dim]
        self.exchange = [f for f in HALO_STATE_FIELDS if f in fields]
        self.per = {f: fields[f].numel() // capacity for f in self.exchange}
        self.width = sum(self.per.values())
        self.boxes, self.box_rank = dec.all_boxes()
        self.my_boxes = dec.boxes(self.rank)
        self.box_counts = [int((self.box_rank == r).sum()) for r in range(self.world)]
        self._desc = None
        self._send = None
        self._recv = None
        self.last = {}
        self.h_evolves = h_evolves
        self.reuse_plan = reuse_plan
        self.device_verdict = False   # set by DistributedRhs when the engine carries the plan verdict as its abort flag
        self._plan = None
        self.plan_builds = 0
        self.stale_plans = 0          # evaluations whose plan had gone stale (re-decided before the evaluation ran)
        self._flag_host = None

    def _rows(self, name: str) -> torch.Tensor:
        return self.fields[name].view(self.capacity, -1)

    # ------------------------------------------------------------------ GPU path: reusable send plan, library kernels
    #
    # Deciding WHO needs WHICH particle costs three collectives and a host wait (per-box h_max, send counts, and the
    # split sizes NCCL wants on the host) -- 0.3 ms at 2 ranks, 1.3 ms at 8 for 10^6 particles per rank
    # (profiles/r01_bench_sedov_{2,8}gpu.json), as much as the evaluation itself.  But the answer changes slowly:
    # between the three evaluations of an RK step, and between consecutive steps, particles move a small fraction
    # of h.  So the decision is made once with head-room and REUSED:
    #
    #   build  reach = (h_k + h_max(box)) * (1 + growth) + 2 D,  D = SKIN * (smallest h anywhere); snapshot x, h
    #   reuse  valid while no particle on any rank has moved further than D from its snapshot and no h has grown by
    #          more than `growth` -- one kernel and one 4-byte all-reduce; the host waits for that verdict only
    #          (the rows already move behind it, speculatively) and re-decides when it is negative.
    #
    # A reused exchange is check -> pack -> all_to_all -> unpack with host-known sizes.
    def _setup_cuda(self, n_owned: int) -> None:
        f = self.fields
        dev = f["x"].device
        eng = self.engine
        # the library's pack / unpack / plan kernels and torch's collectives must share one stream:
        # do it here, not only in the callers
        eng.set_stream(torch.cuda.current_stream(dev).cuda_stream)
        eng.halo_set_domains(self.boxes, self.box_rank, self.world, self.rank)
        self._desc = eng.halo_fields(f, self.exchange, self.capacity)
        self._sum_desc = {}
        eng.set_abort_flag(None)
        self.width = eng.halo_row_width(self._desc)
        self._nb_max = max(self.box_counts)
        self._idx = torch.empty(max(self.capacity, 2 * n_owned), dtype=torch.int32, device=dev)
        self._counts = torch.zeros(self.world + 1, dtype=torch.int32, device=dev)
        self._flag = torch.zeros(1, dtype=torch.int32, device=dev)
        self._flag_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._flag_event = torch.cuda.Event()

    def _build_plan(self, n_owned: int) -> None:
        f, eng, nb, w = self.fields, self.engine, self._nb_max, self.world
        dev = f["x"].device
        growth = self.H_GROWTH if self.h_evolves else 0.0
        # D: the same number on every rank
        hmin = f["h"][:n_owned].min().reshape(1).clone() if n_owned > 0 else torch.full((1,), float("inf"), dtype=torch.float64, device=dev)
        dist.all_reduce(hmin, op=dist.ReduceOp.MIN, group=self.group)
        extra = None
        if self.levels == 2:
            mine = torch.zeros(nb, dtype=torch.float64, device=dev)
            eng.halo_box_hmax(f["x"], f.get("y"), f.get("z"), f["h"], n_owned, mine)
            extra = torch.zeros(w * nb, dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(extra, mine, group=self.group)
        max_move = self.SKIN * float(hmin.item())
        eng.halo_select_plan(f["x"], f.get("y"), f.get("z"), f["h"], n_owned, extra, nb, 1.0 + growth, 2.0 * max_move,
                             self._idx, self._counts)
        all_counts = torch.zeros(w * (w + 1), dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(all_counts, self._counts, group=self.group)
        table = all_counts.cpu().numpy().reshape(w, w + 1)
        if table[:, w].any():
            raise RuntimeError(f"halo send list of a rank does not fit its index buffer ({self._idx.numel()} entries here)")
        send_counts = [int(c) for c in table[self.rank, :w]]
        recv_counts = [int(c) for c in table[:w, self.rank]]
        n_send, n_recv = sum(send_counts), sum(recv_counts)
        if n_owned + n_recv > self.capacity:
            raise RuntimeError(f"halo of {n_recv} particles does not fit: capacity {self.capacity}, owned {n_owned}")
        if self._send is None or self._send.numel() < n_send * self.width:
            self._send = torch.empty(int(n_send * self.width * 1.2) + 64, dtype=torch.float64, device=dev)
        if self._recv is None or self._recv.numel() < n_recv * self.width:
            self._recv = torch.empty(int(n_recv * self.width * 1.2) + 64, dtype=torch.float64, device=dev)
        snap = {a: f[a][:n_owned].clone() for a in self.axes + ["h"]}
        i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=dev)
        self._plan = dict(n_owned=n_owned, send_counts=send_counts, recv_counts=recv_counts, n_send=n_send, n_recv=n_recv,
                          send_counts_dev=i32(send_counts), recv_counts_dev=i32(recv_counts),
                          max_move=max_move, growth=growth, snap=snap)
        self._flag_host.zero_()
        self._flag.zero_()
        if self.levels == 2:   # copies that need their own lists are told the plan's margins (reach of a copy to an owned particle)
            eng.halo_set_list_margin(1.0 + growth, 2.0 * max_move)
        self.plan_builds += 1

    def replan(self, n_owned: int) -> None:
        """The plan went stale: decide again and move the rows again."""
        self.stale_plans += 1
        self._build_plan(n_owned)
        self._move_rows(n_owned)

    def exchange_sums(self, which: int, n_owned: int) -> None:
        """Deliver the owners' neighbour sums (1: kernel-sum density, 2: tensorial correction matrix) of the particles of
        the send plan into the receivers' halo rows: the same plan, a row of 1 or DIM*DIM doubles; stream-ordered."""
        pl, eng = self._plan, self.engine
        if which not in self._sum_desc:
            names = [n for n in SUM_FIELDS[which] if n in self.fields]
            desc = eng.halo_fields(self.fields, names, self.capacity)
            width = eng.halo_row_width(desc)
            self._sum_desc[which] = (desc, width)
        desc, width = self._sum_desc[which]
        n_send, n_recv, w = pl["n_send"], pl["n_recv"], self.world
        if self._send.numel() < n_send * width or self._recv.numel() < n_recv * width:
            raise RuntimeError("neighbour-sum rows are wider than the state rows")
        send, recv = self._send[: n_send * width], self._recv[: n_recv * width]
        eng.halo_pack_by_rank(desc, self._idx, pl["send_counts_dev"], w, n_send, send)
        dist.all_to_all_single(recv, send, output_split_sizes=[c * width for c in pl["recv_counts"]],
                               input_split_sizes=[c * width for c in pl["send_counts"]], group=self.group)
        eng.halo_unpack_by_rank(desc, recv, pl["recv_counts_dev"], w, n_recv, n_owned)

    def _check_plan(self, n_owned: int, read_back: bool = True) -> None:
        """Queue the validity check of the current plan (kernel + 4-byte all-reduce + async read-back); no host wait."""
        f, pl = self.fields, self._plan
        sn = pl["snap"]
        self.engine.halo_plan_check(f["x"], f.get("y"), f.get("z"), f["h"], sn["x"], sn.get("y"), sn.get("z"), sn["h"], n_owned,
                                    pl["max_move"], pl["growth"], self._flag)
        dist.all_reduce(self._flag, op=dist.ReduceOp.MAX, group=self.group)
        if read_back:
            self._flag_host.copy_(self._flag, non_blocking=True)

    def invalidate(self) -> None:
        """Drop the plan (the owned set changed, or a check failed): the next run() decides again."""
        self._plan = None

    def _move_rows(self, n_owned: int) -> None:
        """pack -> all_to_all -> unpack with the plan's (host-known) sizes; nothing here waits for the device."""
        pl, eng = self._plan, self.engine
        n_send, n_recv = pl["n_send"], pl["n_recv"]
        w, width = self.world, self.width
        send = self._send[: n_send * width]
        recv = self._recv[: n_recv * width]
        # every rank's block is stored column by column (b200sph_halo_pack_by_rank); split sizes are in elements
        eng.halo_pack_by_rank(self._desc, self._idx, pl["send_counts_dev"], w, n_send, send)
        dist.all_to_all_single(recv, send, output_split_sizes=[c * width for c in pl["recv_counts"]],
                               input_split_sizes=[c * width for c in pl["send_counts"]], group=self.group)
        eng.halo_unpack_by_rank(self._desc, recv, pl["recv_counts_dev"], w, n_recv, n_owned)

    def _run_cuda(self, n_owned: int) -> int:
        if self.engine is None:
            raise RuntimeError("halo exchange on GPU buffers needs the b200sph engine (no torch fallback on the product path)")
        if self._desc is None:
            self._setup_cuda(n_owned)
        if self._plan is None or self._plan["n_owned"] != n_owned or not self.reuse_plan:
            self._build_plan(n_owned)
            self._move_rows(n_owned)
        elif self.device_verdict:
            # The verdict on the plan stays on the device: the all-reduced flag is the evaluation's abort flag
            # (b200sph_set_abort_flag), every state-modifying kernel reads it first, and a stale plan surfaces as
            # B200SPH_ERR_ABORTED at the end-of-call synchronisation (DistributedRhs.compute re-decides and repeats).
            # Nothing here waits for the device.
            self._check_plan(n_owned, read_back=False)
            self._move_rows(n_owned)
        else:
            # The verdict on the plan is needed before the evaluation starts (an evaluation mutates state -- p, c_s,
            # S, damage -- so it cannot simply be repeated).  Its 4-byte read-back is queued first and the rows move
            # speculatively behind it: the host waits for the verdict only, not for the exchange.
            self._check_plan(n_owned)
            self._flag_event.record()
            self._move_rows(n_owned)
            self._flag_event.synchronize()
            if int(self._flag_host[0]) != 0:
                self.replan(n_owned)
        pl = self._plan
        self.last = dict(n_halo=pl["n_recv"], sent=pl["n_send"], bytes_sent=pl["n_send"] * self.width * 8, plan_builds=self.plan_builds)
        return n_owned + pl["n_recv"]

    # ------------------------------------------------------------------ CPU tensors (gloo tests): the same rule in torch ops
    def _box_hmax(self, n_owned: int) -> np.ndarray:
        """Largest smoothing length per box of every rank (host array aligned with self.boxes)."""
        f = self.fields
        nb_max = max(self.box_counts)
        mine = torch.zeros(nb_max, dtype=torch.float64)
        pos = torch.stack([f[a][:n_owned] for a in self.axes], dim=1)
        h = f["h"][:n_owned]
        for b in range(len(self.my_boxes)):
            lo = torch.as_tensor(self.my_boxes[b, : self.dim])
            hi = torch.as_tensor(self.my_boxes[b, 3: 3 + self.dim])
            inside = ((pos >= lo) & (pos <= hi)).all(dim=1)
            if bool(inside.any()):
                mine[b] = h[inside].max()
        flat = torch.empty(self.world * nb_max, dtype=torch.float64)
        dist.all_gather_into_tensor(flat, mine, group=self.group)
        table = flat.view(self.world, nb_max).numpy()
        return np.concatenate([table[r, : self.box_counts[r]] for r in range(self.world)])

    def _needed_by(self, n_owned: int, extra: np.ndarray, reach_scale: float = 1.0, skin: float = 0.0) -> torch.Tensor:
        """int64 mask per owned particle: bit r set <=> rank r needs a copy (same rule as h_mask in csrc/halo.cu)."""
        f = self.fields
        pos = torch.stack([f[a][:n_owned] for a in self.axes], dim=1)
        h = f["h"][:n_owned]
        mask = torch.zeros(n_owned, dtype=torch.int64)
        for b in range(len(self.box_rank)):
            r = int(self.box_rank[b])
            if r == self.rank:
                continue
            lo = torch.as_tensor(self.boxes[b, : self.dim])
            hi = torch.as_tensor(self.boxes[b, 3: 3 + self.dim])
            gap = torch.clamp(torch.maximum(lo - pos, pos - hi), min=0.0)
            reach = (h + float(extra[b])) * reach_scale * (1.0 + 1e-9) + skin
            mask |= ((gap * gap).sum(dim=1) < reach * reach).to(torch.int64) << r
        return mask

    def run(self, n_owned: int) -> int:
        """Fill rows [n_owned, n_owned + n_halo) with the copies this rank needs; returns n_owned + n_halo."""
        f = self.fields
        dev = f["x"].device
        if self.world == 1:
            self.last = dict(n_halo=0, sent=0, bytes_sent=0)
            return n_owned
        if dev.type == "cuda":
            return self._run_cuda(n_owned)
        return self.move_cpu(n_owned, self.select_cpu(n_owned))

    def select_cpu(self, n_owned: int, reach_scale: float = 1.0, skin: float = 0.0) -> list:
        """CPU tensors: per destination rank, the ascending indices of the owned particles it needs (empty for this rank).
        With reach_scale = 1 + growth and skin = 2 D this is the decision a reusable plan is built from."""
        # second halo level: a copy must also be complete around its own neighbours, which reach up to the
        # largest smoothing length found in the owner's box it borders (all-gathered per box)
        extra = self._box_hmax(n_owned) if self.levels == 2 else np.zeros(len(self.box_rank))
        mask = self._needed_by(n_owned, extra, reach_scale, skin)
        empty = torch.empty(0, dtype=torch.int64)
        return [empty if r == self.rank else torch.nonzero((mask >> r) & 1, as_tuple=False).flatten() for r in range(self.world)]

    def move_cpu(self, n_owned: int, send_idx: list) -> int:
        """CPU tensors: send the CURRENT state of the listed particles and append what arrives behind the owned rows."""
        f = self.fields
        dev = f["x"].device
        send_counts = [int(idx.numel()) for idx in send_idx]
        sc = torch.tensor(send_counts, dtype=torch.int64, device=dev)
        rc = torch.empty_like(sc)
        dist.all_to_all_single(rc, sc, group=self.group)
        recv_counts = [int(c) for c in rc.tolist()]
        n_send, n_recv = sum(send_counts), sum(recv_counts)
        if n_owned + n_recv > self.capacity:
            raise RuntimeError(f"halo of {n_recv} particles does not fit: capacity {self.capacity}, owned {n_owned}")

        idx_all = torch.cat(send_idx) if send_idx else torch.empty(0, dtype=torch.int64, device=dev)
        send = torch.empty(n_send, self.width, dtype=torch.float64, device=dev)
        col = 0
        for name in self.exchange:
            w = self.per[name]
            send[:, col: col + w] = self._rows(name)[idx_all].to(torch.float64)
            col += w
        recv = torch.empty(n_recv, self.width, dtype=torch.float64, device=dev)
        dist.all_to_all_single(recv, send, output_split_sizes=recv_counts, input_split_sizes=send_counts, group=self.group)
        col = 0
        for name in self.exchange:
            w = self.per[name]
            rows = self._rows(name)
            rows[n_owned: n_owned + n_recv] = recv[:, col: col + w].to(rows.dtype)
            col += w
        self.last = dict(n_halo=n_recv, sent=n_send, bytes_sent=n_send * self.width * 8)
        return n_owned + n_recv


class GravitySources:
    """All-gather of x, y, z, m of every rank's owned particles (replicated-tree gravity)."""

    def __init__(self, dim: int, group=None):
        self.dim = dim
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.buffers = None

    def gather(self, fields: dict, n_owned: int):
        """Returns (x, y, z, m, n_total, own_begin): concatenation over ranks in rank order."""
        dev = fields["x"].device
        counts = torch.empty(self.world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(counts, torch.tensor([n_owned], dtype=torch.int64, device=dev), group=self.group)
        counts = [int(c) for c in counts.tolist()]
        n_total, own_begin = sum(counts), sum(counts[: self.rank])
        names = ["x", "y", "z"][: self.dim] + ["m"]
        # one padded all-gather of [x, y, z, m] (equal-sized pieces work on every backend)
        c_max = max(counts)
        mine = torch.zeros(len(names), c_max, dtype=torch.float64, device=dev)
        for k, name in enumerate(names):
            mine[k, :n_owned] = fields[name][:n_owned]
        flat = torch.empty(self.world * mine.numel(), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(flat, mine.view(-1), group=self.group)
        gathered = flat.view(self.world, len(names), c_max)
        out = {name: torch.cat([gathered[r, k, : counts[r]] for r in range(self.world)]) for k, name in enumerate(names)}
        self.buffers = out   # keep alive while the library reads them
        return out.get("x"), out.get("y"), out.get("z"), out["m"], n_total, own_begin


class DistributedRhs:
    """`rightHandSide()` for a particle set spread over the GPUs of one box.

    eval() = halo exchange -> (gravity sources) -> b200sph_rhs_eval on owned + halo particles.
    """

    def __init__(self, engine: "api.RhsEngine", fields: dict, capacity: int, n_owned: int, dec: MortonDecomposition, meta: dict,
                 switches: dict, group=None, external_sums: bool = True, device_verdict: bool = True):
        self.engine = engine
        self.fields = fields
        self.capacity = capacity
        self.n_owned = n_owned
        self.meta = meta
        h_evolves = bool(switches.get("VARIABLE_SML", 0) or switches.get("INTEGRATE_SML", 0))
        on_gpu = fields["x"].device.type == "cuda"
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        # GPU buffers: ONE halo level; density / correction matrix of the copies are delivered by their owners between
        # the stages of the evaluation (neighbour-sum exchange).  CPU tensors (gloo tests, oracle as the evaluator): the
        # two-level halo, where the evaluator completes those sums itself.
        need = halo_levels(switches, bool(meta.get("kernel_sum_materials")))
        self.external_sums = bool(on_gpu and world > 1 and external_sums and need >= 2)
        levels = 1 if self.external_sums else need
        if levels > 2:
            raise NotImplementedError("kernel-sum density together with the tensorial correction needs the neighbour-sum exchange "
                                      "(GPU buffers, external_sums=True); the state-only halo has two levels")
        self.halo = HaloExchange(fields, capacity, dec, levels=levels, group=group, engine=engine, h_evolves=h_evolves)
        self.halo.device_verdict = bool(on_gpu and world > 1 and device_verdict)
        self.sum_exchanges = 0
        self.gravity = GravitySources(dec.dim, group) if meta.get("selfgravity") else None
        self.n_total = n_owned

    def exchange(self) -> int:
        self.n_total = self.halo.run(self.n_owned)
        if self.gravity is not None and self.halo.world > 1:
            x, y, z, m, n_src, own_begin = self.gravity.gather(self.fields, self.n_owned)
            self.engine.set_gravity_sources(x, y, z, m, n_src, own_begin)
        return self.n_total

    def compute(self) -> None:
        view = api.make_view(self.fields, None, self.n_total, n_real=self.n_total, max_num_flaws=self.meta["max_num_flaws"],
                             selfgravity=self.meta["selfgravity"], theta=self.meta["theta"],
                             grav_const=self.engine.materials.grav_const)
        eng = self.engine
        eng.set_owned(self.n_owned)
        if self.halo.world == 1 or not (self.external_sums or self.halo.device_verdict):
            eng.set_halo_sums(False)
            eng.rhs_eval(view)
            return
        eng.set_halo_sums(self.external_sums)
        if self.halo.device_verdict:
            eng.set_abort_flag(self.halo._flag)
        for attempt in range(2):
            try:
                for stage in (0, 1, 2):
                    pending = eng.rhs_eval_stage(view, stage)
                    if pending:
                        self.halo.exchange_sums(pending, self.n_owned)
                        self.sum_exchanges += 1
                return
            except api.B200SphError as exc:
                if exc.code != eng.ERR_ABORTED or attempt == 1:
                    raise
                # stale send plan: nothing was modified; decide again (every rank sees the same all-reduced flag)
                self.halo.replan(self.n_owned)
                self.n_total = self.n_owned + self.halo._plan["n_recv"]
                view = api.make_view(self.fields, None, self.n_total, n_real=self.n_total, max_num_flaws=self.meta["max_num_flaws"],
                                     selfgravity=self.meta["selfgravity"], theta=self.meta["theta"],
                                     grav_const=eng.materials.grav_const)

    def eval(self) -> None:
        self.exchange()
        self.compute()


def scatter_scenario(arrays: dict, n: int, dim: int, max_flaws: int, rank: int, world: int, headroom: float = 1.6,
                     min_extra: int = 4096):
    """Cut a full (host, numpy) particle set into this rank's Morton piece with halo headroom.

    Returns (local numpy arrays sized for `capacity` particles, n_owned, capacity, global indices of the owned
    particles, the decomposition)."""
    x = np.stack([arrays[a][:n] for a in ["x", "y", "z"][:dim]], axis=1)
    dec, parts = morton_partition(x, world)
    mine = parts[rank]
    n_owned = len(mine)
    capacity = n_owned if world == 1 else int(n_owned * headroom) + min_extra
    local = {}
    for name, arr in arrays.items():
        per = arr.size // n
        out = np.zeros(capacity * per, dtype=arr.dtype)
        out.reshape(capacity, per)[:n_owned] = arr.reshape(n, per)[mine]
        local[name] = out
    return local, n_owned, capacity, mine, dec
