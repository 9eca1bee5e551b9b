"""Shared helpers for the parity tests: golden loading, oracle binding, field comparison."""
from __future__ import annotations

import ctypes as C
import os
import sys
import tempfile

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "oracle"))

from miluphcuda_b200 import api, scenarios  # noqa: E402

GOLDEN_DIR = os.path.join(REPO, "tests", "golden")
CONFIGS = scenarios.CONFIG_NAMES
# scenario variants that run on another config's switch set (library): the giant collision with tabulated-EOS
# (ANEOS-format) materials uses the giant_hydro build
VARIANT_CONFIG = scenarios.VARIANT_CONFIG
# every golden file present (written by oracle/make_golden.py on the GPU box from the reference's own build)
GOLDEN_CASES = sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz"))

# fields compared against the reference's output.  `depth` is excluded: the reference's value
# depends on the insertion order of its racy tree build (src/tree.cu:259).
RATE_FIELDS = ("ax", "ay", "az", "dxdt", "dydt", "dzdt", "drhodt", "dedt", "dhdt", "dSdt", "dddt", "dalphadt",
               "ddamage_porjutzidt", "edotp", "g_ax", "g_ay", "g_az")
STATE_FIELDS = ("rho", "p", "cs", "h", "e", "S", "d", "damage_total", "alpha_jutzi", "alpha_jutzi_old", "dalphadp",
                "dalphadrho", "delpdelrho", "delpdele", "f", "sigma", "tensorialCorrectionMatrix", "R", "plastic_f",
                "local_strain", "muijmax", "vx", "vy", "vz")
INT_COMPARE = ("noi", "numActiveFlaws")
RTOL = 1e-9   # north_star: "within 1e-9 relative in fp64"


def config_of(case: str) -> str:
    base = case[:-8] if case.endswith("_stirred") else case
    return VARIANT_CONFIG.get(base, base)


def load_golden(case: str):
    return np.load(os.path.join(GOLDEN_DIR, f"{case}.npz"))


def golden_expected(g, stage: str, field: str):
    """Value of `field` after call `stage` ("out1"/"out2"); pruned entries fall back to the earlier stage."""
    order = {"out1": ("out1_", "in_"), "out2": ("out2_", "out1_", "in_")}[stage]
    for prefix in order:
        if prefix + field in g.files:
            return g[prefix + field]
    return None


def golden_neighbours(g):
    ptr, idx = g["nbr_ptr"], g["nbr_idx"]
    return [idx[ptr[i]:ptr[i + 1]] for i in range(int(g["n"]))]


def write_material_cfg(g, directory: str) -> str:
    path = os.path.join(directory, "material.cfg")
    with open(path, "w") as fh:
        fh.write(str(g["material_cfg"]))
    for key in g.files:
        if key.startswith("include_"):
            with open(os.path.join(directory, key[len("include_"):]), "w") as fh:
                fh.write(str(g[key]))
    return path


def state_from_golden(g, config: str):
    """numpy arrays for every field of the switch set, initialised from the golden `in_` dump
    (what the reference's buffers held right before its first rightHandSide())."""
    sw = scenarios.read_switches(config)
    n = int(g["n"])
    dim = sw["DIM"]
    selfgrav = bool(int(g["selfgravity"]))
    max_flaws = sw.get("MAX_NUM_FLAWS", 1)
    p_fields, rhs_fields = api.fields_for(sw, selfgrav)
    arrays = {}
    for name in p_fields + rhs_fields:
        dtype = np.int32 if name in api.INT_FIELDS else np.float64
        shape = api.field_shape(name, n, dim, max_flaws)
        src = "in_" + name
        if src in g.files and g[src].shape == shape:
            arrays[name] = np.ascontiguousarray(g[src].astype(dtype))
        else:
            arrays[name] = np.zeros(shape, dtype=dtype)
    return arrays, dict(n=n, max_num_flaws=max_flaws, selfgravity=selfgrav, theta=float(g["theta"]))


# ----------------------------------------------------------------------------- oracle binding
_ORACLES = {}


def oracle_lib(config: str):
    if config not in _ORACLES:
        import build_oracle
        path = build_oracle.build((config,))[0]
        lib = C.CDLL(path)
        lib.oracle_rhs.restype = C.c_int
        lib.oracle_rhs.argtypes = [C.POINTER(api.View), C.POINTER(api.Materials), C.c_void_p, C.POINTER(C.c_int)]
        lib.oracle_max_num_interactions.restype = C.c_int
        lib.oracle_dim.restype = C.c_int
        for name in ("oracle_pressure", "oracle_init_soundspeed", "oracle_damage_limit"):
            fn = getattr(lib, name)
            fn.restype = C.c_int
            fn.argtypes = [C.POINTER(api.View), C.POINTER(api.Materials)]
        _ORACLES[config] = lib
    return _ORACLES[config]


def oracle_rhs(config: str, arrays: dict, materials: api.MaterialTables, meta: dict):
    """Run the CPU restatement in place on `arrays`; returns (rc, offender, interactions[n, MAX])."""
    lib = oracle_lib(config)
    n = meta["n"]
    maxni = lib.oracle_max_num_interactions()
    inter = np.full((n, maxni), -1, dtype=np.int32)
    view = api.make_view(arrays, None, n, max_num_flaws=meta["max_num_flaws"], selfgravity=meta["selfgravity"],
                         theta=meta["theta"], grav_const=materials.grav_const)
    off = C.c_int(-1)
    rc = lib.oracle_rhs(C.byref(view), materials.pointer(), inter.ctypes.data, C.byref(off))
    return rc, off.value, inter


# ----------------------------------------------------------------------------- comparison
def field_error(got: np.ndarray, ref: np.ndarray, min_scale: float = 0.0) -> float:
    """max |got-ref| / max(|ref|, rms(ref))  -- the per-field scale of SURVEY H4.

    `min_scale` lets a caller supply the field's natural magnitude when the reference values
    themselves are rounding noise (e.g. the plastic strain rate (1 - plastic_f) * edot on a second
    call, where plastic_f is 1 up to one ulp)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    if ref.size == 0:
        return 0.0
    scale = max(np.sqrt(np.mean(ref * ref)), min_scale)
    denom = np.maximum(np.abs(ref), scale)
    denom = np.where(denom > 0.0, denom, 1.0)
    bad = ~np.isfinite(got) | ~np.isfinite(ref)
    if bad.any():
        same = np.array_equal(np.isnan(got), np.isnan(ref))
        if not same:
            return float("inf")
    with np.errstate(invalid="ignore"):
        err = np.abs(got - ref) / denom
    err = err[np.isfinite(err)]
    return float(err.max()) if err.size else 0.0


def deactivated_rows(g):
    """Particles with materialId == EOS_TYPE_IGNORE (-1) in the golden input, or None.

    For those the reference indexes its material tables with -1 (matEOS[-1], matSml[-1], mat_f_sml_max[-1], ...:
    src/soundspeed.cu:48, src/tree.cu:940, src/plasticity.cu:150 read out of bounds), so what it leaves in their STATE
    members (c_s = 0, h = 0, S = 0, and with h = 0 an empty neighbour list) is whatever lies in front of the tables.
    They are compared on what is defined: rates, frozen velocities, g_a, and their absence from every neighbour set."""
    if "in_materialId" not in g.files:
        return None
    rows = np.asarray(g["in_materialId"]) == -1
    return rows if rows.any() else None


def _drop_rows(a, rows):
    a = np.asarray(a)
    per = a.size // rows.size
    return a.reshape(rows.size, per)[~rows].reshape(-1)


def compare_fields(arrays: dict, g, stage: str, fields, rtol: float = RTOL, skip=()) -> dict:
    """{field: error} for every field present both in `arrays` and in the golden stage."""
    report = {}
    dead = deactivated_rows(g)
    for name in fields:
        if name in skip or name not in arrays:
            continue
        ref = golden_expected(g, stage, name)
        if ref is None or ref.shape != np.asarray(arrays[name]).shape:
            continue
        min_scale = 0.0
        if stage == "out2":
            first = golden_expected(g, "out1", name)
            if first is not None and first.shape == ref.shape:
                min_scale = float(np.sqrt(np.mean(first.astype(np.float64) ** 2)))
        got = arrays[name]
        if dead is not None and name in STATE_FIELDS and name not in ("vx", "vy", "vz"):
            got, ref = _drop_rows(got, dead), _drop_rows(ref, dead)
        report[name] = field_error(got, ref, min_scale)
    return report


def tmp_material(g):
    """Context: temp dir holding the golden's material.cfg; returns (tmpdir object, path)."""
    td = tempfile.TemporaryDirectory()
    return td, write_material_cfg(g, td.name)
