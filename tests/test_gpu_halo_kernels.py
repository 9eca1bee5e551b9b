"""GPU (one device is enough): the device side of the multi-GPU halo exchange (csrc/halo.cu) against the same
rules written in numpy.  One rank of a 4-rank Morton decomposition is played on cuda:0 without any collective:

  b200sph_halo_box_hmax        largest smoothing length inside each of the rank's boxes
  b200sph_halo_select_plan     who needs which particle: distance to a foreign box < (h_k + extra) * scale + skin
  b200sph_halo_pack_by_rank /  rank blocks stored column by column; a round trip must reproduce the selected rows
  b200sph_halo_unpack_by_rank
  b200sph_halo_plan_check      a plan is stale once a particle moved beyond max_move or h outgrew the head-room

The multi-rank tests (tests/test_multigpu_*.py) cover the same code end to end where two GPUs exist."""
import numpy as np
import pytest

from miluphcuda_b200 import api, multigpu, scenarios, state

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

WORLD = 4


def _setup(config, n, rank, tmp_path):
    sc = scenarios.make(config, n, stirred=True)
    cfg = state.write_material_files(sc, str(tmp_path))
    mats = api.MaterialTables(config, cfg)
    full, meta = state.scenario_arrays(sc, mats)
    local, n_owned, capacity, mine, dec = multigpu.scatter_scenario(full, sc.n, sc.dim, meta["max_num_flaws"], rank, WORLD)
    eng = api.RhsEngine(config, n_max=capacity, material_cfg=cfg)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)   # library kernels ordered with the torch ops of this test
    boxes, box_rank = dec.all_boxes()
    eng.halo_set_domains(boxes, box_rank, WORLD, rank)
    dev = {k: torch.from_numpy(v).cuda() for k, v in local.items()}
    return sc, full, local, dev, eng, dec, boxes, box_rank, n_owned, capacity, mine


def _box_hmax_table(full, sc, dec, boxes, box_rank):
    """numpy: largest h inside every box (of its owner's particles), as the [WORLD, nb_max] table the kernels index."""
    ax = ["x", "y", "z"][: sc.dim]
    pos = np.stack([full[a] for a in ax], axis=1)
    owner = dec.owner_of(dec.cell_ids(pos))
    nb_max = max(int((box_rank == r).sum()) for r in range(WORLD))
    table = np.zeros((WORLD, nb_max))
    local_index = np.zeros(len(box_rank), dtype=int)
    seen = {}
    for b, r in enumerate(box_rank):
        local_index[b] = seen.get(int(r), 0)
        seen[int(r)] = local_index[b] + 1
        inside = (owner == r) & np.all((pos >= boxes[b, : sc.dim]) & (pos <= boxes[b, 3: 3 + sc.dim]), axis=1)
        if inside.any():
            table[r, local_index[b]] = full["h"][inside].max()
    return table, local_index, nb_max


@pytest.mark.parametrize("config,n,rank", [("sedov", 40000, 1), ("impact", 30000, 2), ("rings", 30000, 0)])
def test_selection_pack_unpack_match_numpy(config, n, rank, tmp_path):
    sc, full, local, dev, eng, dec, boxes, box_rank, n_owned, capacity, mine = _setup(config, n, rank, tmp_path)
    dim = sc.dim
    table, local_index, nb_max = _box_hmax_table(full, sc, dec, boxes, box_rank)

    # --- largest h per own box
    hmax_mine = torch.zeros(nb_max, dtype=torch.float64, device="cuda")
    eng.halo_box_hmax(dev["x"], dev.get("y"), dev.get("z"), dev["h"], n_owned, hmax_mine)
    torch.cuda.synchronize()
    assert np.array_equal(hmax_mine.cpu().numpy(), table[rank])

    # --- selection (two-level rule with head-room and skin)
    scale, skin = 1.02, 0.013 * float(full["h"].min())
    extra = torch.from_numpy(table.reshape(-1).copy()).cuda()
    idx = torch.full((2 * n_owned + 64,), -1, dtype=torch.int32, device="cuda")
    counts = torch.zeros(WORLD + 1, dtype=torch.int32, device="cuda")
    eng.halo_select_plan(dev["x"], dev.get("y"), dev.get("z"), dev["h"], n_owned, extra, nb_max, scale, skin, idx, counts)
    torch.cuda.synchronize()
    got_counts = counts.cpu().numpy()
    got_idx = idx.cpu().numpy()
    assert got_counts[WORLD] == 0
    ax = ["x", "y", "z"][:dim]
    pos = np.stack([local[a][:n_owned] for a in ax], axis=1)
    h = local["h"][:n_owned]
    want = []
    for r in range(WORLD):
        need = np.zeros(n_owned, dtype=bool)
        if r != rank:
            for b in np.nonzero(box_rank == r)[0]:
                gap = np.maximum(np.maximum(boxes[b, :dim] - pos, pos - boxes[b, 3: 3 + dim]), 0.0)
                reach = (h + table[r, local_index[b]]) * scale * (1.0 + 1e-9) + skin
                need |= (gap * gap).sum(axis=1) < reach * reach
        want.append(np.nonzero(need)[0])
    off = 0
    for r in range(WORLD):
        assert got_counts[r] == len(want[r]), f"rank {r}: {got_counts[r]} selected, numpy says {len(want[r])}"
        assert np.array_equal(got_idx[off: off + got_counts[r]], want[r]), f"send list for rank {r} differs (must be ascending)"
        off += got_counts[r]
    n_send = off
    assert n_send > 0

    # --- pack (column-major per rank) and the inverse
    names = [f for f in multigpu.HALO_STATE_FIELDS if f in dev]
    desc = eng.halo_fields(dev, names, capacity)
    width = eng.halo_row_width(desc)
    send = torch.zeros(n_send * width, dtype=torch.float64, device="cuda")
    eng.halo_pack_by_rank(desc, idx, counts, WORLD, n_send, send)
    torch.cuda.synchronize()
    buf = send.cpu().numpy()
    rows = np.concatenate([np.concatenate([local[f].reshape(capacity, -1)[got_idx[:n_send]].astype(np.float64) for f in names], axis=1)])
    assert rows.shape == (n_send, width)
    off = 0
    for r in range(WORLD):
        c = int(got_counts[r])
        block = buf[off * width: (off + c) * width].reshape(width, c)      # [column][row]
        assert np.array_equal(block.T, rows[off: off + c]), f"block for rank {r} is not the column-major image of its rows"
        off += c
    # unpack behind the owned rows of a fresh field set that has room for every row
    cap2 = n_owned + n_send
    back = {}
    for f in list(names):
        per = dev[f].numel() // capacity
        t = torch.full((cap2 * per,), 3, dtype=dev[f].dtype, device="cuda")
        t[: n_owned * per] = dev[f][: n_owned * per]
        back[f] = t
    desc2 = eng.halo_fields(back, names, cap2)
    eng.halo_unpack_by_rank(desc2, send, counts, WORLD, n_send, n_owned)
    torch.cuda.synchronize()
    for f in names:
        got = back[f].cpu().numpy().reshape(cap2, -1)
        assert np.array_equal(got[n_owned:], local[f].reshape(capacity, -1)[got_idx[:n_send]]), f
        assert np.array_equal(got[:n_owned], local[f].reshape(capacity, -1)[:n_owned]), f"{f}: owned rows were touched"
    eng.close()


def test_plan_check_flags_motion_and_growth(tmp_path):
    sc, full, local, dev, eng, dec, boxes, box_rank, n_owned, capacity, mine = _setup("sedov", 20000, 1, tmp_path)
    snap = {a: dev[a][:n_owned].clone() for a in ("x", "y", "z", "h")}
    flag = torch.full((1,), 7, dtype=torch.int32, device="cuda")
    hmin = float(full["h"].min())
    max_move, growth = 0.15 * hmin, 0.02

    def verdict():
        eng.halo_plan_check(dev["x"], dev["y"], dev["z"], dev["h"], snap["x"], snap["y"], snap["z"], snap["h"], n_owned, max_move, growth, flag)
        torch.cuda.synchronize()
        return int(flag.item())

    assert verdict() == 0                                   # nothing moved
    dev["y"][n_owned // 2] += 0.99 * max_move
    dev["h"][5] *= 1.0 + 0.99 * growth
    assert verdict() == 0                                   # inside the tolerance
    dev["y"][n_owned // 2] += 0.02 * max_move
    assert verdict() == 1                                   # one particle beyond the skin
    dev["y"][n_owned // 2] = snap["y"][n_owned // 2]
    assert verdict() == 0
    dev["h"][5] = snap["h"][5] * (1.0 + 1.5 * growth)
    assert verdict() == 1                                   # a smoothing length outgrew the head-room
    dev["h"][5] = snap["h"][5]
    dev["z"][n_owned - 1] = float("nan")
    assert verdict() == 1                                   # NaN counts as a violation
    eng.close()
