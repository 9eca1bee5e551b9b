#!/usr/bin/env python3
"""Device time of the halo kernels of ONE rank of a W-rank decomposition, on one GPU and without NCCL
(diagnosis: which of box_hmax / select / pack / unpack costs what as the rank count grows).
usage: python tools/halo_kernel_timing.py [workload] [particles_per_rank] [world]"""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from miluphcuda_b200 import api, multigpu, scenarios, state

workload = sys.argv[1] if len(sys.argv) > 1 else "sedov"
npart = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
world = int(sys.argv[3]) if len(sys.argv) > 3 else 8
rank = world // 2
torch.cuda.set_device(0)
sc = scenarios.make(workload, npart * world)
td = tempfile.TemporaryDirectory()
cfg = state.write_material_files(sc, td.name)
mats = api.MaterialTables(workload, cfg)
full, meta = state.scenario_arrays(sc, mats)
arrays, n, cap, _, dec = multigpu.scatter_scenario(full, sc.n, sc.dim, meta["max_num_flaws"], rank, world)
eng = api.RhsEngine(workload, n_max=cap, device=0, material_cfg=cfg)
eng.set_stream(torch.cuda.current_stream().cuda_stream)
dev = {k: torch.from_numpy(v).cuda() for k, v in arrays.items()}
boxes, box_rank = dec.all_boxes()
eng.halo_set_domains(boxes, box_rank, world, rank)
exchange = [f for f in multigpu.HALO_STATE_FIELDS if f in dev]
desc = eng.halo_fields(dev, exchange, cap)
width = eng.halo_row_width(desc)
nb = max(int((box_rank == r).sum()) for r in range(world))
hmax_mine = torch.zeros(nb, dtype=torch.float64, device="cuda")
hmax_all = torch.full((world * nb,), float(dev["h"][:n].max()), dtype=torch.float64, device="cuda")
idx = torch.empty(max(cap, 2 * n), dtype=torch.int32, device="cuda")
counts = torch.zeros(world + 1, dtype=torch.int32, device="cuda")
f = dev


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


out = {}
out["box_hmax"] = timed(lambda: eng.halo_box_hmax(f["x"], f.get("y"), f.get("z"), f["h"], n, hmax_mine))
out["select"] = timed(lambda: eng.halo_select(f["x"], f.get("y"), f.get("z"), f["h"], n, hmax_all, nb, idx, counts))
c = counts.cpu().numpy()
ns = int(c[:world].sum())
send = torch.empty(ns, width, dtype=torch.float64, device="cuda")
out["pack"] = timed(lambda: eng.halo_pack(desc, idx, ns, send))
nr = min(ns, cap - n)
out["unpack"] = timed(lambda: eng.halo_unpack(desc, send, nr, n))
print({k: round(v, 4) for k, v in out.items()}, "ms; n", n, "n_send", ns, "width", width, "boxes", len(box_rank), "world", world, flush=True)
