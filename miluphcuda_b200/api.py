"""ctypes binding of include/b200sph.h and the host-side mirror of the reference's
`rightHandSide()` interface.

The product is the C-ABI library `libb200sph_<config>.so` (hand-written sm_100a CUDA,
built by `miluphcuda_b200.build`).  This module is what tests and bench.py use to call
it; PyTorch only provides device memory.  There is no CPU fallback: if the library
for a config is missing, loading raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

# ----------------------------------------------------------------------------- ABI mirror
PARTICLE_FIELDS = (
    "x", "y", "z", "vx", "vy", "vz", "dxdt", "dydt", "dzdt", "ax", "ay", "az", "g_ax", "g_ay", "g_az",
    "g_local_cellsize", "g_x", "g_y", "g_z", "m", "h", "h0", "dhdt", "rho", "drhodt", "p", "e", "dedt",
    "S", "dSdt", "local_strain", "ep", "edotp", "plastic_f", "sigma", "R", "d", "damage_total", "dddt",
    "numFlaws", "numActiveFlaws", "flaws", "damage_porjutzi", "ddamage_porjutzidt", "muijmax",
    "pold", "alpha_jutzi", "alpha_jutzi_old", "dalphadt", "dalphadp", "dalphadrho", "f", "delpdelrho", "delpdele",
    "tensorialCorrectionMatrix", "cs", "noi", "materialId", "depth",
)
INT_FIELDS = frozenset({"numFlaws", "numActiveFlaws", "noi", "materialId", "depth"})
TENSOR_FIELDS = frozenset({"S", "dSdt", "sigma", "R", "tensorialCorrectionMatrix"})

MATERIAL_INT_TABLES = frozenset({
    "matEOS", "matdensity_via_kernel_sum", "matcrushcurve_style", "aneos_n_rho", "aneos_n_e", "aneos_rho_id",
    "aneos_e_id", "aneos_matrix_id",
})
MATERIAL_TABLES = (
    "matEOS", "matSml", "mat_f_sml_min", "mat_f_sml_max", "matAlpha", "matBeta", "matPolytropicK", "matPolytropicGamma",
    "matIsothermalSoundSpeed", "matBulkmodulus", "matShearmodulus", "matYoungModulus", "matYieldStress", "matRho0", "matN",
    "matRhoLimit", "matcsLimit", "matTillRho0", "matTillA", "matTillB", "matTillE0", "matTillEiv", "matTillEcv", "matTilla",
    "matTillb", "matTillAlpha", "matTillBeta", "matCohesion", "matCohesionDamaged", "matInternalFriction",
    "matInternalFrictionDamaged", "matMeltEnergy", "matDensityFloor", "matEnergyFloor", "matdensity_via_kernel_sum",
    "matexponent_tensor", "matepsilon_stress", "matmean_particle_distance", "matporjutzi_p_elastic",
    "matporjutzi_p_transition", "matporjutzi_p_compacted", "matporjutzi_alpha_0", "matporjutzi_alpha_e",
    "matporjutzi_alpha_t", "matporjutzi_n1", "matporjutzi_n2", "matcs_porous", "matcs_solid", "matcrushcurve_style",
    "aneos_n_rho", "aneos_n_e", "aneos_rho_id", "aneos_e_id", "aneos_matrix_id", "aneos_rho", "aneos_e", "aneos_p",
    "aneos_cs", "aneos_bulk_cs", "aneos_gamma",
)


class ParticleArrays(C.Structure):
    _fields_ = [(name, C.c_void_p) for name in PARTICLE_FIELDS]


class View(C.Structure):
    _fields_ = [
        ("n", C.c_int), ("n_real", C.c_int), ("max_num_flaws", C.c_int), ("selfgravity", C.c_int),
        ("decouplegravity", C.c_int), ("is_relaxation_run", C.c_int), ("theta", C.c_double), ("grav_const", C.c_double),
        ("p", ParticleArrays), ("p_rhs", ParticleArrays),
    ]


class Materials(C.Structure):
    _fields_ = (
        [("n_materials", C.c_int)]
        + [(name, C.c_void_p) for name in MATERIAL_TABLES]
        + [("aneos_rho_len", C.c_int64), ("aneos_e_len", C.c_int64), ("aneos_matrix_len", C.c_int64)]
    )

    def table(self, name: str) -> np.ndarray:
        """Host copy of one per-material table (only valid for host-resident tables)."""
        ptr = getattr(self, name)
        if not ptr:
            return np.zeros(self.n_materials, dtype=np.int32 if name in MATERIAL_INT_TABLES else np.float64)
        if name in ("aneos_rho", "aneos_e", "aneos_p", "aneos_cs"):
            length = {"aneos_rho": self.aneos_rho_len, "aneos_e": self.aneos_e_len}.get(name, self.aneos_matrix_len)
        else:
            length = self.n_materials
        ctype = C.c_int if name in MATERIAL_INT_TABLES else C.c_double
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(int(length),)).copy()


class HaloField(C.Structure):
    _fields_ = [("data", C.c_void_p), ("per", C.c_int), ("kind", C.c_int)]


class Stats(C.Structure):
    _fields_ = [
        ("kernel_launches", C.c_int), ("n_cells", C.c_int), ("max_noi", C.c_int), ("total_noi", C.c_int64),
        ("cell_size", C.c_double), ("ms_total", C.c_float), ("ms_sort", C.c_float), ("ms_neighbours", C.c_float),
        ("ms_density", C.c_float), ("ms_pointwise", C.c_float), ("ms_correction", C.c_float), ("ms_forces", C.c_float),
        ("ms_gravity", C.c_float), ("ms_scatter", C.c_float), ("gravity_recomputed", C.c_int),
    ]

    def as_dict(self) -> dict:
        return {k: getattr(self, k) for k, _ in self._fields_}


class Rk2Params(C.Structure):
    """b200sph_rk2_params (include/b200sph.h): -Q / -M / -F and the RK2_* switches of the reference's rk2adaptive.h."""
    _fields_ = [
        ("rk_epsrel", C.c_double), ("dt_max", C.c_double), ("first_dt", C.c_double),
        ("use_courant_limit", C.c_int), ("use_forces_limit", C.c_int), ("use_damage_limit", C.c_int),
        ("use_velocity_error", C.c_int), ("use_density_error", C.c_int), ("use_energy_error", C.c_int),
        ("limit_pressure_change", C.c_int), ("limit_alpha_change", C.c_int),
        ("courant_fact", C.c_double), ("forces_fact", C.c_double),
        ("location_safety", C.c_double), ("min_vel_change", C.c_double), ("tiny_density", C.c_double), ("tiny_energy", C.c_double),
        ("timestep_safety", C.c_double), ("smallest_dt_allowed", C.c_double),
        ("max_damage_change", C.c_double), ("max_alpha_change", C.c_double), ("max_pressure_change", C.c_double),
    ]


class MgStats(C.Structure):
    _fields_ = [("n_boxes", C.c_int), ("n_halo", C.c_int), ("plan_builds", C.c_int), ("stale_plans", C.c_int), ("sum_exchanges", C.c_int),
                ("migrated_out", C.c_longlong), ("migrated_in", C.c_longlong), ("halo_bytes_sent", C.c_longlong)]


class Conserved(C.Structure):
    _fields_ = [("mass", C.c_double), ("e_kin", C.c_double), ("e_int", C.c_double), ("p_abs", C.c_double), ("p", C.c_double * 3),
                ("L_abs", C.c_double), ("L", C.c_double * 3), ("bary_pos", C.c_double * 3), ("bary_vel", C.c_double * 3),
                ("n_ignored", C.c_int)]


class Rk2State(C.Structure):
    _fields_ = [
        ("t", C.c_double), ("dt", C.c_double), ("dt_suggested", C.c_double), ("dt_done", C.c_double),
        ("accepted", C.c_int), ("rejected", C.c_int), ("rhs_calls", C.c_int), ("intervals", C.c_int),
        ("approaching_output_time", C.c_int), ("err", C.c_double * 6),
    ]


ERRORS = {1: "TOO_MANY_INTERACTIONS", 2: "BAD_ARGUMENT", 3: "SWITCH_MISMATCH", 4: "UNSUPPORTED", 5: "NONFINITE", 6: "ABORTED", -1: "CUDA"}


class B200SphError(RuntimeError):
    def __init__(self, code: int, message: str, offender: int = -1):
        super().__init__(f"b200sph error {code} ({ERRORS.get(code, '?')}): {message}")
        self.code = code
        self.offender = offender


# ----------------------------------------------------------------------------- library loading
_EXPORTS = {
    "b200sph_abi_version": (C.c_int, []),
    "b200sph_config_name": (C.c_char_p, []),
    "b200sph_switch_hash": (C.c_uint64, []),
    "b200sph_switch_value": (C.c_int, [C.c_char_p]),
    "b200sph_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_uint64]),
    "b200sph_destroy": (C.c_int, [C.c_void_p]),
    "b200sph_last_error": (C.c_char_p, [C.c_void_p]),
    "b200sph_materials_load": (C.c_int, [C.c_char_p, C.POINTER(C.POINTER(Materials)), C.POINTER(C.c_double), C.c_char_p, C.c_size_t]),
    "b200sph_materials_free": (None, [C.POINTER(Materials)]),
    "b200sph_set_materials": (C.c_int, [C.c_void_p, C.POINTER(Materials)]),
    "b200sph_rhs_eval": (C.c_int, [C.c_void_p, C.POINTER(View), C.POINTER(C.c_int)]),
    "b200sph_rhs_eval_host": (C.c_int, [C.c_void_p, C.POINTER(View), C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "b200sph_host_options": (C.c_int, [C.c_void_p, C.c_int]),
    "b200sph_pressure": (C.c_int, [C.c_void_p, C.POINTER(View)]),
    "b200sph_damage_limit": (C.c_int, [C.c_void_p, C.POINTER(View)]),
    "b200sph_init_soundspeed": (C.c_int, [C.c_void_p, C.POINTER(View)]),
    "b200sph_export_interactions": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "b200sph_mg_unique_id": (C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t]),
    "b200sph_mg_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "b200sph_mg_destroy": (C.c_int, [C.c_void_p]),
    "b200sph_mg_last_error": (C.c_char_p, [C.c_void_p]),
    "b200sph_mg_decompose": (C.c_int, [C.c_void_p, C.POINTER(View), C.c_int, C.c_int]),
    "b200sph_mg_migrate": (C.c_int, [C.c_void_p, C.POINTER(View), C.POINTER(ParticleArrays), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "b200sph_mg_rhs_eval": (C.c_int, [C.c_void_p, C.POINTER(View), C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "b200sph_mg_rk2_advance": (C.c_int, [C.c_void_p, C.POINTER(View), C.POINTER(ParticleArrays), C.POINTER(Rk2Params), C.c_double,
                                         C.POINTER(Rk2State), C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "b200sph_mg_get_stats": (C.c_int, [C.c_void_p, C.POINTER(MgStats)]),
    "b200sph_conserved_quantities": (C.c_int, [C.c_void_p, C.POINTER(View), C.POINTER(Conserved)]),
    "b200sph_reorder": (C.c_int, [C.c_void_p, C.POINTER(View), C.POINTER(ParticleArrays), C.c_int, C.c_void_p]),
    "b200sph_rk2_default_params": (C.c_int, [C.POINTER(Rk2Params)]),
    "b200sph_rk2_init": (C.c_int, [C.c_void_p, C.POINTER(View), C.POINTER(ParticleArrays)]),
    "b200sph_rk2_step": (C.c_int, [C.c_void_p, C.POINTER(View), C.POINTER(ParticleArrays), C.POINTER(Rk2Params), C.c_double,
                                   C.POINTER(Rk2State), C.POINTER(C.c_int)]),
    "b200sph_rk2_advance": (C.c_int, [C.c_void_p, C.POINTER(View), C.POINTER(ParticleArrays), C.POINTER(Rk2Params), C.c_double,
                                      C.POINTER(Rk2State), C.POINTER(C.c_int)]),
    "b200sph_get_stats": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "b200sph_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "b200sph_set_owned": (C.c_int, [C.c_void_p, C.c_int]),
    "b200sph_set_global_domain": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "b200sph_halo_set_domains": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "b200sph_halo_box_hmax": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "b200sph_halo_select": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                      C.c_void_p, C.c_int, C.c_void_p]),
    "b200sph_halo_select_plan": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                           C.c_double, C.c_double, C.c_void_p, C.c_int, C.c_void_p]),
    "b200sph_halo_plan_check": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_void_p]),
    "b200sph_halo_row_width": (C.c_int, [C.c_void_p, C.c_int]),
    "b200sph_halo_pack": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "b200sph_halo_unpack": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]),
    "b200sph_halo_pack_by_rank": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "b200sph_halo_unpack_by_rank": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "b200sph_set_halo_sums": (C.c_int, [C.c_void_p, C.c_int]),
    "b200sph_rhs_eval_stage": (C.c_int, [C.c_void_p, C.POINTER(View), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "b200sph_set_abort_flag": (C.c_int, [C.c_void_p, C.c_void_p]),
    "b200sph_halo_set_list_margin": (C.c_int, [C.c_void_p, C.c_double, C.c_double]),
    "b200sph_set_gravity_sources": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
}
_LIBS: dict = {}


def load_library(config: str) -> C.CDLL:
    """dlopen libb200sph_<config>.so; raises if it has not been built (no fallback)."""
    if config in _LIBS:
        return _LIBS[config]
    path = _build.lib_path(config)
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"{path} is missing: build the CUDA extension first (python -m miluphcuda_b200.build {config}); "
            "there is no CPU fallback for the SPH right-hand side")
    lib = C.CDLL(path)
    for name, (restype, argtypes) in _EXPORTS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    _LIBS[config] = lib
    return lib


def exported_symbols() -> tuple:
    return tuple(_EXPORTS)


# ----------------------------------------------------------------------------- materials
class MaterialTables:
    """material.cfg parsed by the library's own libconfig-format reader (host memory)."""

    def __init__(self, config: str, cfg_path: str):
        self.lib = load_library(config)
        self._ptr = C.POINTER(Materials)()
        g = C.c_double(0.0)
        err = C.create_string_buffer(1024)
        rc = self.lib.b200sph_materials_load(cfg_path.encode(), C.byref(self._ptr), C.byref(g), err, len(err))
        if rc != 0:
            raise B200SphError(rc, err.value.decode(errors="replace"))
        self.grav_const = g.value

    @property
    def struct(self) -> Materials:
        return self._ptr.contents

    def pointer(self):
        return self._ptr

    def table(self, name: str) -> np.ndarray:
        return self.struct.table(name)

    def __del__(self):
        try:
            if self._ptr:
                self.lib.b200sph_materials_free(self._ptr)
                self._ptr = C.POINTER(Materials)()
        except Exception:
            pass


# ----------------------------------------------------------------------------- particle state
def fields_for(switches: dict, selfgravity: bool = False) -> tuple:
    """(p fields, p_rhs fields) present for a switch set -- the members of `struct Particle`
    (reference include/miluph.h:67-275) that the hot path touches."""
    g = lambda k: switches.get(k, 0)
    dim = g("DIM")
    ax = ["x", "y", "z"][:dim]
    p = list(ax) + ["v" + a for a in ax] + ["d" + a + "dt" for a in ax] + ["a" + a for a in ax]
    p += ["m", "h", "rho", "drhodt", "p", "e", "cs", "noi", "depth"]
    rhs = ["materialId", "h0"]
    if selfgravity:
        p += ["g_a" + a for a in ax]
        rhs += ["g_" + a for a in ax] + ["g_local_cellsize"]
    if g("INTEGRATE_ENERGY"):
        p.append("dedt")
    if g("INTEGRATE_SML"):
        p.append("dhdt")
    if g("ARTIFICIAL_VISCOSITY"):
        p.append("muijmax")
    if g("SOLID"):
        p += ["S", "dSdt", "local_strain", "ep", "edotp"]
        rhs += ["plastic_f", "sigma"]
    if g("TENSORIAL_CORRECTION"):
        rhs.append("tensorialCorrectionMatrix")
    if g("ARTIFICIAL_STRESS"):
        rhs.append("R")
    if g("FRAGMENTATION"):
        p += ["d", "damage_total", "dddt", "numFlaws", "numActiveFlaws"]
        rhs.append("flaws")
        if g("PALPHA_POROSITY"):
            p += ["damage_porjutzi", "ddamage_porjutzidt"]
    if g("PALPHA_POROSITY"):
        p += ["pold", "alpha_jutzi", "alpha_jutzi_old", "dalphadt", "dalphadp", "dalphadrho", "f", "delpdelrho", "delpdele"]
    return tuple(p), tuple(rhs)


def field_shape(name: str, n: int, dim: int, max_flaws: int) -> tuple:
    if name in TENSOR_FIELDS:
        return (n * dim * dim,)
    if name == "flaws":
        return (n * max(max_flaws, 1),)
    return (n,)


def _ptr_of(arr) -> int:
    if arr is None:
        return 0
    if isinstance(arr, np.ndarray):
        if not arr.flags["C_CONTIGUOUS"]:
            raise ValueError("arrays passed through the C-ABI must be contiguous")
        return arr.ctypes.data
    return int(arr.data_ptr())  # torch tensor


def make_view(arrays: dict, rhs_arrays: dict | None, n: int, *, n_real: int | None = None, max_num_flaws: int = 1,
              selfgravity: bool = False, decouplegravity: bool = False, theta: float = 0.5,
              grav_const: float = 6.67408e-11, is_relaxation_run: bool = False) -> View:
    """Build a b200sph_view from {field: numpy array | torch tensor}.  The caller keeps the arrays alive."""
    v = View()
    v.n = n
    v.n_real = n if n_real is None else n_real
    v.max_num_flaws = max_num_flaws
    v.selfgravity = int(selfgravity)
    v.decouplegravity = int(decouplegravity)
    v.is_relaxation_run = int(is_relaxation_run)
    v.theta = theta
    v.grav_const = grav_const
    rhs_arrays = arrays if rhs_arrays is None else rhs_arrays
    for name in PARTICLE_FIELDS:
        setattr(v.p, name, _ptr_of(arrays.get(name)) or None)
        setattr(v.p_rhs, name, _ptr_of(rhs_arrays.get(name)) or None)
    return v


# ----------------------------------------------------------------------------- engine
class RhsEngine:
    """One b200sph handle: owns the device scratch; `rhs_eval` is the drop-in for rightHandSide()."""

    def __init__(self, config: str, n_max: int, device: int = 0, material_cfg: str | None = None):
        self.config = config
        self.lib = load_library(config)
        self.handle = C.c_void_p()
        rc = self.lib.b200sph_create(C.byref(self.handle), n_max, device, self.lib.b200sph_switch_hash())
        if rc != 0:
            raise B200SphError(rc, self.lib.b200sph_last_error(None).decode(errors="replace"))
        self.materials = None
        if material_cfg is not None:
            self.set_material_cfg(material_cfg)

    def _check(self, rc: int, offender: int = -1) -> None:
        if rc != 0:
            raise B200SphError(rc, self.lib.b200sph_last_error(self.handle).decode(errors="replace"), offender)

    def set_material_cfg(self, cfg_path: str) -> MaterialTables:
        self.materials = MaterialTables(self.config, cfg_path)
        self._check(self.lib.b200sph_set_materials(self.handle, self.materials.pointer()))
        return self.materials

    def rhs_eval(self, view: View) -> None:
        off = C.c_int(-1)
        rc = self.lib.b200sph_rhs_eval(self.handle, C.byref(view), C.byref(off))
        self._check(rc, off.value)

    def rhs_eval_host(self, view: View) -> tuple:
        off = C.c_int(-1)
        h2d, d2h = C.c_int64(0), C.c_int64(0)
        rc = self.lib.b200sph_rhs_eval_host(self.handle, C.byref(view), C.byref(off), C.byref(h2d), C.byref(d2h))
        self._check(rc, off.value)
        return h2d.value, d2h.value

    SUM_DENSITY, SUM_CORRECTION, ERR_ABORTED = 1, 2, 6

    def set_halo_sums(self, external: bool) -> None:
        self._check(self.lib.b200sph_set_halo_sums(self.handle, int(bool(external))))

    def rhs_eval_stage(self, view: View, stage: int) -> int:
        """One stage (0, 1, 2) of the evaluation; returns the neighbour sum (SUM_*) the host must deliver before the next."""
        pending, off = C.c_int(0), C.c_int(-1)
        rc = self.lib.b200sph_rhs_eval_stage(self.handle, C.byref(view), stage, C.byref(pending), C.byref(off))
        self._check(rc, off.value)
        return pending.value

    def set_abort_flag(self, device_flag) -> None:
        self._check(self.lib.b200sph_set_abort_flag(self.handle, _ptr_of(device_flag) or None))

    def halo_set_list_margin(self, reach_scale: float, skin: float) -> None:
        self._check(self.lib.b200sph_halo_set_list_margin(self.handle, float(reach_scale), float(skin)))

    HOST_CACHE_IMMUTABLES = 1
    HOST_SKIP_SCRATCH = 2

    def host_options(self, options: int) -> None:
        """B200SPH_HOST_* bits of the host-buffer call (include/b200sph.h); also drops cached immutables."""
        self._check(self.lib.b200sph_host_options(self.handle, int(options)))

    def pressure(self, view: View) -> None:
        self._check(self.lib.b200sph_pressure(self.handle, C.byref(view)))

    def damage_limit(self, view: View) -> None:
        self._check(self.lib.b200sph_damage_limit(self.handle, C.byref(view)))

    def init_soundspeed(self, view: View) -> None:
        self._check(self.lib.b200sph_init_soundspeed(self.handle, C.byref(view)))

    def conserved_quantities(self, view: View) -> Conserved:
        out = Conserved()
        self._check(self.lib.b200sph_conserved_quantities(self.handle, C.byref(view), C.byref(out)))
        return out

    def reorder(self, view: View, extra=None, n_extra: int = 0, perm_out=None) -> None:
        """Put the caller's buffers (and `extra`, a C array from rk2_buffers) into search-cell order: new[k] = old[perm[k]]."""
        self._check(self.lib.b200sph_reorder(self.handle, C.byref(view), extra, n_extra, _ptr_of(perm_out) or None))

    # ---- rk2_adaptive on the device (csrc/integrate.cu) ----
    def rk2_default_params(self) -> Rk2Params:
        prm = Rk2Params()
        self._check(self.lib.b200sph_rk2_default_params(C.byref(prm)))
        return prm

    @staticmethod
    def rk2_buffers(buffers) -> "C.Array":
        """C array of three b200sph_particle_arrays (RKSTART, RKFIRST, RKSECOND) from three {field: tensor} dicts."""
        arr = (ParticleArrays * 3)()
        for k, fields in enumerate(buffers):
            for name in PARTICLE_FIELDS:
                setattr(arr[k], name, _ptr_of(fields.get(name)) or None)
        return arr

    def rk2_init(self, view: View, rk) -> None:
        self._check(self.lib.b200sph_rk2_init(self.handle, C.byref(view), rk))

    def rk2_step(self, view: View, rk, prm: Rk2Params, t_end: float, state: Rk2State) -> None:
        off = C.c_int(-1)
        self._check(self.lib.b200sph_rk2_step(self.handle, C.byref(view), rk, C.byref(prm), t_end, C.byref(state), C.byref(off)), off.value)

    def rk2_advance(self, view: View, rk, prm: Rk2Params, t_end: float, state: Rk2State) -> None:
        off = C.c_int(-1)
        self._check(self.lib.b200sph_rk2_advance(self.handle, C.byref(view), rk, C.byref(prm), t_end, C.byref(state), C.byref(off)), off.value)

    def export_interactions(self, device_int_buffer, max_per_row: int) -> None:
        self._check(self.lib.b200sph_export_interactions(self.handle, _ptr_of(device_int_buffer), max_per_row))

    def stats(self) -> dict:
        st = Stats()
        self._check(self.lib.b200sph_get_stats(self.handle, C.byref(st)))
        return st.as_dict()

    def set_stream(self, cuda_stream: int | None) -> None:
        self._check(self.lib.b200sph_set_stream(self.handle, cuda_stream))

    def set_owned(self, n_owned: int) -> None:
        self._check(self.lib.b200sph_set_owned(self.handle, n_owned))

    def set_global_domain(self, lo, hi) -> None:
        if lo is None:
            self._check(self.lib.b200sph_set_global_domain(self.handle, None, None))
            return
        a = (C.c_double * 3)(*[float(x) for x in lo])
        b = (C.c_double * 3)(*[float(x) for x in hi])
        self._check(self.lib.b200sph_set_global_domain(self.handle, a, b))

    # ---- halo exchange, device side (csrc/halo.cu) ----
    def halo_set_domains(self, boxes: np.ndarray, box_rank: np.ndarray, n_ranks: int, my_rank: int) -> None:
        boxes = np.ascontiguousarray(boxes, dtype=np.float64)
        box_rank = np.ascontiguousarray(box_rank, dtype=np.int32)
        self._check(self.lib.b200sph_halo_set_domains(self.handle, boxes.ctypes.data, box_rank.ctypes.data, len(box_rank), n_ranks, my_rank))

    def halo_box_hmax(self, x, y, z, h, n: int, hmax_out) -> None:
        self._check(self.lib.b200sph_halo_box_hmax(self.handle, _ptr_of(x), _ptr_of(y) or None, _ptr_of(z) or None, _ptr_of(h), n,
                                                   _ptr_of(hmax_out), int(hmax_out.numel())))

    def halo_select(self, x, y, z, h, n: int, extra, extra_stride: int, idx_out, counts_out) -> None:
        self._check(self.lib.b200sph_halo_select(self.handle, _ptr_of(x), _ptr_of(y) or None, _ptr_of(z) or None, _ptr_of(h), n,
                                                 _ptr_of(extra) or None, extra_stride, _ptr_of(idx_out), int(idx_out.numel()),
                                                 _ptr_of(counts_out)))

    def halo_select_plan(self, x, y, z, h, n: int, extra, extra_stride: int, reach_scale: float, skin: float, idx_out, counts_out) -> None:
        self._check(self.lib.b200sph_halo_select_plan(self.handle, _ptr_of(x), _ptr_of(y) or None, _ptr_of(z) or None, _ptr_of(h), n,
                                                      _ptr_of(extra) or None, extra_stride, float(reach_scale), float(skin),
                                                      _ptr_of(idx_out), int(idx_out.numel()), _ptr_of(counts_out)))

    def halo_plan_check(self, x, y, z, h, x0, y0, z0, h0, n: int, max_move: float, growth: float, flag_out) -> None:
        self._check(self.lib.b200sph_halo_plan_check(self.handle, _ptr_of(x), _ptr_of(y) or None, _ptr_of(z) or None, _ptr_of(h),
                                                     _ptr_of(x0), _ptr_of(y0) or None, _ptr_of(z0) or None, _ptr_of(h0), n,
                                                     float(max_move), float(growth), _ptr_of(flag_out)))

    @staticmethod
    def halo_fields(fields: dict, names, capacity: int, zero_names=()):
        """ctypes array of b200sph_halo_field for the named members of `fields` (flat tensors sized for `capacity`)."""
        present = [n for n in names if n in fields] + [n for n in zero_names if n in fields]
        arr = (HaloField * len(present))()
        for k, name in enumerate(present):
            t = fields[name]
            arr[k].data = _ptr_of(t)
            arr[k].per = t.numel() // capacity
            arr[k].kind = 2 if name in zero_names else (1 if name in INT_FIELDS else 0)
        return arr

    def halo_row_width(self, desc) -> int:
        return int(self.lib.b200sph_halo_row_width(desc, len(desc)))

    def halo_pack(self, desc, idx, n_rows: int, out) -> None:
        self._check(self.lib.b200sph_halo_pack(self.handle, desc, len(desc), _ptr_of(idx), n_rows, _ptr_of(out)))

    def halo_unpack(self, desc, buf, n_rows: int, first_row: int) -> None:
        self._check(self.lib.b200sph_halo_unpack(self.handle, desc, len(desc), _ptr_of(buf), n_rows, first_row))

    def halo_pack_by_rank(self, desc, idx, counts, n_ranks: int, n_rows: int, out) -> None:
        self._check(self.lib.b200sph_halo_pack_by_rank(self.handle, desc, len(desc), _ptr_of(idx), _ptr_of(counts), n_ranks, n_rows,
                                                       _ptr_of(out)))

    def halo_unpack_by_rank(self, desc, buf, counts, n_ranks: int, n_rows: int, first_row: int) -> None:
        self._check(self.lib.b200sph_halo_unpack_by_rank(self.handle, desc, len(desc), _ptr_of(buf), _ptr_of(counts), n_ranks, n_rows,
                                                         first_row))

    def set_gravity_sources(self, x, y, z, m, n_sources: int, own_begin: int) -> None:
        """Multi-GPU gravity: device arrays of the global particle set (see include/b200sph.h)."""
        self._check(self.lib.b200sph_set_gravity_sources(self.handle, _ptr_of(x) or None, _ptr_of(y) or None, _ptr_of(z) or None,
                                                         _ptr_of(m) or None, n_sources, own_begin))

    def close(self) -> None:
        if self.handle:
            self.lib.b200sph_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class NativeMultiGpu:
    """The multi-GPU host of csrc/mg.cu (C++ over NCCL) seen from Python: what a C host calls, one object per rank."""

    def __init__(self, engine: RhsEngine, rank: int, world: int, unique_id: bytes):
        self.engine, self.lib = engine, engine.lib
        self.handle = C.c_void_p()
        buf = C.create_string_buffer(unique_id, 128)
        rc = self.lib.b200sph_mg_create(C.byref(self.handle), engine.handle, rank, world, buf)
        if rc != 0:
            raise B200SphError(rc, self._error())

    @staticmethod
    def unique_id(config: str) -> bytes:
        lib = load_library(config)
        buf, err = C.create_string_buffer(128), C.create_string_buffer(512)
        rc = lib.b200sph_mg_unique_id(buf, err, len(err))
        if rc != 0:
            raise B200SphError(rc, err.value.decode(errors="replace"))
        return buf.raw

    def _error(self) -> str:
        msg = self.lib.b200sph_mg_last_error(self.handle) if self.handle else b""
        return (msg or b"").decode(errors="replace") or self.lib.b200sph_last_error(self.engine.handle).decode(errors="replace")

    def _check(self, rc: int, offender: int = -1) -> None:
        if rc != 0:
            raise B200SphError(rc, self._error(), offender)

    def decompose(self, view: View, n_held: int, by_work: bool = False) -> None:
        self._check(self.lib.b200sph_mg_decompose(self.handle, C.byref(view), n_held, int(by_work)))

    def migrate(self, view: View, n_held: int, capacity: int, extra=None, n_extra: int = 0) -> int:
        out = C.c_int(0)
        self._check(self.lib.b200sph_mg_migrate(self.handle, C.byref(view), extra, n_extra, n_held, capacity, C.byref(out)))
        return out.value

    def rhs_eval(self, view: View, n_owned: int, capacity: int) -> int:
        n_total, off = C.c_int(0), C.c_int(-1)
        self._check(self.lib.b200sph_mg_rhs_eval(self.handle, C.byref(view), n_owned, capacity, C.byref(n_total), C.byref(off)), off.value)
        return n_total.value

    def rk2_advance(self, view: View, rk, prm: "Rk2Params", t_end: float, state: "Rk2State", n_owned: int, capacity: int) -> None:
        off = C.c_int(-1)
        self._check(self.lib.b200sph_mg_rk2_advance(self.handle, C.byref(view), rk, C.byref(prm), t_end, C.byref(state), n_owned, capacity,
                                                    C.byref(off)), off.value)

    def stats(self) -> dict:
        st = MgStats()
        self._check(self.lib.b200sph_mg_get_stats(self.handle, C.byref(st)))
        return {k: getattr(st, k) for k, _ in st._fields_}

    def close(self) -> None:
        if self.handle:
            self.lib.b200sph_mg_destroy(self.handle)
            self.handle = C.c_void_p()
