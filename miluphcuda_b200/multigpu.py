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
)
# integer members that must be defined (zero) on halo copies
HALO_ZERO_FIELDS = ("numFlaws", "numActiveFlaws")


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


def halo_levels(switches: dict) -> int:
    """2 when neighbours' own neighbour sums are needed (kernel-sum density or tensorial correction), else 1."""
    if not switches.get("INTEGRATE_DENSITY", 0) or switches.get("TENSORIAL_CORRECTION", 0):
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

    H_GROWTH = 0.02   # head-room on last evaluation's per-box h_max when h is integrated (VARIABLE_SML / INTEGRATE_SML)

    def __init__(self, fields: dict, capacity: int, dec: MortonDecomposition, levels: int = 2, group=None, engine=None,
                 h_evolves: bool = True):
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
        self.last_retry = False
        self.h_evolves = h_evolves

    def _rows(self, name: str) -> torch.Tensor:
        return self.fields[name].view(self.capacity, -1)

    # ------------------------------------------------------------------ GPU path: library kernels, two collectives, one host wait
    def _setup_cuda(self, n_owned: int) -> None:
        f = self.fields
        dev = f["x"].device
        eng = self.engine
        eng.halo_set_domains(self.boxes, self.box_rank, self.world, self.rank)
        self._desc = eng.halo_fields(f, self.exchange, self.capacity, HALO_ZERO_FIELDS)
        self.width = eng.halo_row_width(self._desc)
        nb = self._nb_max = max(self.box_counts)
        w = self.world
        self._idx = torch.empty(max(self.capacity, 2 * n_owned), dtype=torch.int32, device=dev)
        self._counts = torch.zeros(w + 1, dtype=torch.int32, device=dev)
        # one row per rank: [largest h in each of its boxes (nb), rows it sends to every rank (w), send list overflow (1)]
        self._meta_mine = torch.zeros(nb + w + 1, dtype=torch.float64, device=dev)
        self._meta_all = torch.zeros(w * (nb + w + 1), dtype=torch.float64, device=dev)
        self._meta_host = torch.zeros(w * (nb + w + 1), dtype=torch.float64).pin_memory()
        self._hmax_used = None        # device [w * nb]: the per-box h_max the selection works with
        self._hmax_used_host = None   # the same numbers on the host, for the check after the all-gather

    def _select_and_share(self, n_owned: int):
        """Selection with the current `_hmax_used`, then ONE all-gather of (fresh box h_max, send counts); returns the host table."""
        f, eng, nb, w = self.fields, self.engine, self._nb_max, self.world
        mine = self._meta_mine
        if self.levels == 2:
            eng.halo_box_hmax(f["x"], f.get("y"), f.get("z"), f["h"], n_owned, mine[:nb])
        eng.halo_select(f["x"], f.get("y"), f.get("z"), f["h"], n_owned, self._hmax_used if self.levels == 2 else None, nb,
                        self._idx, self._counts)
        mine[nb:].copy_(self._counts)
        dist.all_gather_into_tensor(self._meta_all, mine, group=self.group)
        self._meta_host.copy_(self._meta_all, non_blocking=True)
        torch.cuda.current_stream().synchronize()          # the only host wait: NCCL needs the split sizes
        return self._meta_host.view(w, nb + w + 1).numpy()

    def _run_cuda(self, n_owned: int) -> int:
        """Stream-ordered exchange.  The second halo level needs every rank's largest h per box, which used to cost
        its own all-gather before the selection could start.  Now the selection runs with the values of the
        previous evaluation (inflated by `H_GROWTH` when h evolves), the fresh values travel in the same all-gather
        as the send counts, and the exchange is repeated with them only if a box outgrew what was assumed --
        over-selection is always safe, under-selection never happens."""
        f = self.fields
        dev = f["x"].device
        eng = self.engine
        if eng is None:
            raise RuntimeError("halo exchange on GPU buffers needs the b200sph engine (no torch fallback on the product path)")
        if self._desc is None:
            self._setup_cuda(n_owned)
        nb, w = self._nb_max, self.world
        if self.levels == 2 and self._hmax_used is None:
            # first evaluation: fetch the table once, the old way
            eng.halo_box_hmax(f["x"], f.get("y"), f.get("z"), f["h"], n_owned, self._meta_mine[:nb])
            first = torch.zeros(w * nb, dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(first, self._meta_mine[:nb].contiguous(), group=self.group)
            self._hmax_used = first
            self._hmax_used_host = first.cpu().numpy().reshape(w, nb).copy()
        table = self._select_and_share(n_owned)
        self.last_retry = False
        if self.levels == 2:
            fresh = table[:, :nb]
            if (fresh > self._hmax_used_host).any():
                # a box holds a larger h than the selection assumed: redo it with the fresh table (exact this time)
                self._hmax_used = self._meta_all.view(w, -1)[:, :nb].contiguous().view(-1)
                self._hmax_used_host = fresh.copy()
                table = self._select_and_share(n_owned)
                fresh = table[:, :nb]
                self.last_retry = True
            # assumption for the next evaluation
            growth = 1.0 + self.H_GROWTH if self.h_evolves else 1.0
            self._hmax_used = (self._meta_all.view(w, -1)[:, :nb] * growth).contiguous().view(-1)
            self._hmax_used_host = fresh * growth
        counts = table[:, nb: nb + w]
        if table[:, nb + w].any():
            raise RuntimeError(f"halo send list of a rank does not fit its index buffer ({self._idx.numel()} entries here)")
        send_counts = [int(c) for c in counts[self.rank]]
        recv_counts = [int(c) for c in counts[:, self.rank]]
        n_send, n_recv = sum(send_counts), sum(recv_counts)
        if n_owned + n_recv > self.capacity:
            raise RuntimeError(f"halo of {n_recv} particles does not fit: capacity {self.capacity}, owned {n_owned}")
        if self._send is None or self._send.numel() < n_send * self.width:
            self._send = torch.empty(int(n_send * self.width * 1.2) + 64, dtype=torch.float64, device=dev)
        if self._recv is None or self._recv.numel() < n_recv * self.width:
            self._recv = torch.empty(int(n_recv * self.width * 1.2) + 64, dtype=torch.float64, device=dev)
        send = self._send[: n_send * self.width].view(n_send, self.width)
        recv = self._recv[: n_recv * self.width].view(n_recv, self.width)
        eng.halo_pack(self._desc, self._idx, n_send, send)
        dist.all_to_all_single(recv, send, output_split_sizes=recv_counts, input_split_sizes=send_counts, group=self.group)
        eng.halo_unpack(self._desc, recv, n_recv, n_owned)
        self.last = dict(n_halo=n_recv, sent=n_send, bytes_sent=n_send * self.width * 8)
        return n_owned + n_recv

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

    def _needed_by(self, n_owned: int, extra: np.ndarray) -> torch.Tensor:
        """int64 mask per owned particle: bit r set <=> rank r needs a copy."""
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
            reach = (h + float(extra[b])) * (1.0 + 1e-9)
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
        # second halo level: a copy must also be complete around its own neighbours, which reach up to the
        # largest smoothing length found in the owner's box it borders (all-gathered per box)
        extra = self._box_hmax(n_owned) if self.levels == 2 else np.zeros(len(self.box_rank))
        mask = self._needed_by(n_owned, extra)

        send_idx, send_counts = [], []
        for r in range(self.world):
            if r == self.rank:
                send_counts.append(0)
                continue
            idx = torch.nonzero((mask >> r) & 1, as_tuple=False).flatten()
            send_idx.append(idx)
            send_counts.append(int(idx.numel()))
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
        for name in HALO_ZERO_FIELDS:
            if name in f:
                self._rows(name)[n_owned: n_owned + n_recv] = 0
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
                 switches: dict, group=None):
        self.engine = engine
        self.fields = fields
        self.capacity = capacity
        self.n_owned = n_owned
        self.meta = meta
        h_evolves = bool(switches.get("VARIABLE_SML", 0) or switches.get("INTEGRATE_SML", 0))
        self.halo = HaloExchange(fields, capacity, dec, levels=halo_levels(switches), group=group, engine=engine,
                                 h_evolves=h_evolves)
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
        self.engine.set_owned(self.n_owned)
        self.engine.rhs_eval(view)

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
