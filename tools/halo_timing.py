#!/usr/bin/env python3
"""Wall clock of the GPU halo exchange, deciding every time vs reusing the send plan, and of its stages
(diagnosis; run under torchrun with >= 2 GPUs).  Back-to-back calls, one synchronisation per measurement."""
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

def wall(fn, reps):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3 / reps

hx.run(n)
acc = {}
hx.reuse_plan = False
acc["decide_every_time"] = wall(lambda: hx.run(n), 20)
hx.reuse_plan = True
hx.run(n)
acc["reuse_plan"] = wall(lambda: hx.run(n), 20)
pl = hx._plan
w = world
send = hx._send[: pl["n_send"] * hx.width]; recv = hx._recv[: pl["n_recv"] * hx.width]
acc["check"] = wall(lambda: hx._check_plan(n), 20)
acc["pack"] = wall(lambda: eng.halo_pack_by_rank(hx._desc, hx._idx, pl["send_counts_dev"], w, pl["n_send"], send), 20)
acc["a2a_data"] = wall(lambda: dist.all_to_all_single(recv, send, output_split_sizes=[c * hx.width for c in pl["recv_counts"]],
                                                      input_split_sizes=[c * hx.width for c in pl["send_counts"]]), 20)
acc["unpack"] = wall(lambda: eng.halo_unpack_by_rank(hx._desc, recv, pl["recv_counts_dev"], w, pl["n_recv"], n), 20)
if rank == 0:
    print({k: round(v, 4) for k, v in acc.items()}, "ms; n_send", pl["n_send"], "width", hx.width, "boxes", len(hx.box_rank),
          "plan_builds", hx.plan_builds, flush=True)
dist.destroy_process_group()
