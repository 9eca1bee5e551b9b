#!/usr/bin/env python3
"""Stage-by-stage wall clock of the GPU halo exchange (diagnosis; run under torchrun with >= 2 GPUs).
Each stage is bracketed by torch.cuda.synchronize(), so the numbers add up to more than the pipelined cost."""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from miluphcuda_b200 import api, multigpu, scenarios, state

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
workload = sys.argv[1] if len(sys.argv) > 1 else "sedov"
npart = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
sc = scenarios.make(workload, npart * world)
td = tempfile.TemporaryDirectory()
cfg = state.write_material_files(sc, td.name)
mats = api.MaterialTables(workload, cfg)
full, meta = state.scenario_arrays(sc, mats)
arrays, n, cap, _, dec = multigpu.scatter_scenario(full, sc.n, sc.dim, meta["max_num_flaws"], rank, world)
eng = api.RhsEngine(workload, n_max=cap, device=lr, material_cfg=cfg)
eng.set_stream(torch.cuda.current_stream().cuda_stream)
dev = {k: torch.from_numpy(v).cuda() for k, v in arrays.items()}
hx = multigpu.HaloExchange(dev, cap, dec, levels=multigpu.halo_levels(sc.switches()), engine=eng)
for _ in range(3):
    hx.run(n)
torch.cuda.synchronize(); dist.barrier()

def timed(label, fn, acc):
    torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize()
    acc[label] = acc.get(label, 0.0) + (time.perf_counter() - t0) * 1e3
    return out

acc = {}
reps = 20
f = dev
nb, w = hx._nb_max, world
for _ in range(reps):
    dist.barrier()
    mine = hx._meta_mine
    timed("box_hmax", lambda: eng.halo_box_hmax(f["x"], f.get("y"), f.get("z"), f["h"], n, mine[:nb]), acc)
    timed("select", lambda: eng.halo_select(f["x"], f.get("y"), f.get("z"), f["h"], n, hx._hmax_used if hx.levels == 2 else None, nb,
                                            hx._idx, hx._counts), acc)
    def share():
        mine[nb:].copy_(hx._counts)
        dist.all_gather_into_tensor(hx._meta_all, mine)
        hx._meta_host.copy_(hx._meta_all, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return hx._meta_host.view(w, nb + w + 1).numpy()
    table = timed("all_gather_meta+read", share, acc)
    counts = table[:, nb: nb + w]
    sc_, rc_ = [int(c) for c in counts[rank]], [int(c) for c in counts[:, rank]]
    ns, nr = sum(sc_), sum(rc_)
    send = hx._send[: ns * hx.width].view(ns, hx.width); recv = hx._recv[: nr * hx.width].view(nr, hx.width)
    timed("pack", lambda: eng.halo_pack(hx._desc, hx._idx, ns, send), acc)
    timed("a2a_data", lambda: dist.all_to_all_single(recv, send, output_split_sizes=rc_, input_split_sizes=sc_), acc)
    timed("unpack", lambda: eng.halo_unpack(hx._desc, recv, nr, n), acc)
    dist.barrier()
    t0 = time.perf_counter(); hx.run(n); torch.cuda.synchronize(); acc["whole_run"] = acc.get("whole_run", 0.0) + (time.perf_counter() - t0) * 1e3
if rank == 0:
    print({k: round(v / reps, 4) for k, v in acc.items()}, "n_send", ns, "width", hx.width, "boxes", len(hx.box_rank), "retry", hx.last_retry, flush=True)
dist.destroy_process_group()
