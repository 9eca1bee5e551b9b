"""Deterministic synthetic initial conditions for the five scored scenarios.

Each generator returns a `Scenario`: particle arrays in the caller's order plus
the material.cfg text, with the same columns the reference's ASCII reader
expects for that switch set (reference: src/io.cu:1022-1314).  Shapes follow
the reference's own generators / shipped inputs:

* shocktube  -- test_cases/shocktube/shocktube1D.py
* sedov      -- test_cases/sedov/sedov.cpp:66-182
* rings      -- test_cases/colliding_rings/generate_initial_rings.py
* impact     -- examples/impact (.internal/impact_ini.params), synthetic equivalent
* giant_*    -- examples/giant_collisions (.internal/spheres_ini.log), synthetic equivalent

No RNG except the impact thinning / flaw assignment, which is seeded.
"""
from __future__ import annotations

import os
import re
from dataclasses import dataclass, field

import numpy as np

CONFIG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs")
CONFIG_NAMES = ("shocktube", "sedov", "rings", "impact", "giant_hydro", "giant_solid", "nakamura")
# scenario variants that run on another config's switch set (library)
VARIANT_CONFIG = {"giant_aneos": "giant_hydro", "sedov_ignore": "sedov", "impact_ignore": "impact", "giant_ignore": "giant_hydro",
                  "impact_aneos": "impact",
                  "impact_crush1": "impact", "impact_crush2": "impact", "impact_crush3": "impact", "impact_crush4": "impact"}


def read_switches(config: str) -> dict:
    """Parse configs/<config>/parameter.h into {switch: int}; absent switches are 0."""
    path = os.path.join(CONFIG_DIR, config, "parameter.h")
    sw: dict = {}
    with open(path) as fh:
        for line in fh:
            m = re.match(r"\s*#\s*define\s+(\w+)\s+(-?\d+)\s*$", line)
            if m:
                sw[m.group(1)] = int(m.group(2))
    return sw


@dataclass
class Scenario:
    config: str
    dim: int
    x: np.ndarray            # (N, dim)
    v: np.ndarray            # (N, dim)
    m: np.ndarray
    material_cfg: str
    rho: np.ndarray | None = None
    e: np.ndarray | None = None
    h: np.ndarray | None = None      # per-particle sml (READ_INITIAL_SML_FROM_PARTICLE_FILE)
    mat: np.ndarray | None = None
    S: np.ndarray | None = None      # (N, dim*dim)
    d: np.ndarray | None = None      # DIM-root of tensile damage
    num_flaws: np.ndarray | None = None
    flaws: np.ndarray | None = None  # (N, MAX_NUM_FLAWS), ascending, padded with 0
    alpha: np.ndarray | None = None
    pold: np.ndarray | None = None
    includes: dict = field(default_factory=dict)   # extra files referenced by @include
    selfgravity: bool = False
    theta: float = 0.5

    @property
    def n(self) -> int:
        return int(self.x.shape[0])

    def switches(self) -> dict:
        return read_switches(self.config)

    # ------------------------------------------------------------------ I/O
    def columns(self) -> list:
        """Column list in the reference's input order for this switch set."""
        sw = self.switches()
        g = lambda k: sw.get(k, 0)
        cols = [self.x[:, k] for k in range(self.dim)] + [self.v[:, k] for k in range(self.dim)] + [self.m]
        if g("INTEGRATE_DENSITY"):
            cols.append(self.rho)
        if g("INTEGRATE_ENERGY"):
            cols.append(self.e)
        if g("READ_INITIAL_SML_FROM_PARTICLE_FILE"):
            cols.append(self.h)
        cols.append(self.mat)
        if g("FRAGMENTATION"):
            cols.append(self.num_flaws)
            cols.append(self.d)
        if g("SOLID"):
            cols += [self.S[:, k] for k in range(self.dim * self.dim)]
        if g("PALPHA_POROSITY"):
            cols += [self.alpha, self.pold]
        return cols

    def write_ascii(self, path: str) -> None:
        """Write the reference's ASCII input format (one particle per line)."""
        sw = self.switches()
        cols = self.columns()
        n = self.n
        try:
            # C++ writer: shortest round-trip decimal per double, seconds instead of minutes at 10^6 particles.
            # Unused flaw slots are nulls -> empty fields, i.e. runs of blanks the reference's fscanf reader skips.
            import pyarrow as pa
            import pyarrow.csv as pacsv
            arrs = [pa.array(np.ascontiguousarray(c)) for c in cols]
            if sw.get("FRAGMENTATION", 0):
                nf = np.asarray(self.num_flaws)
                for f in range(int(nf.max()) if n else 0):
                    arrs.append(pa.array(np.ascontiguousarray(self.flaws[:, f]), mask=(nf <= f)))
            table = pa.table(arrs, names=[f"c{k}" for k in range(len(arrs))])
            pacsv.write_csv(table, path, write_options=pacsv.WriteOptions(include_header=False, delimiter=" ", quoting_style="none"))
            if sw.get("FRAGMENTATION", 0):
                # the reader wants the newline right after the last value it expects (src/io.cu:1290-1313)
                with open(path, "rb") as fh:
                    data = fh.read()
                with open(path, "wb") as fh:
                    fh.write(re.sub(rb" +\n", b"\n", data))
            return
        except ImportError:
            pass
        fixed = []
        for c in cols:
            c = np.asarray(c)
            if c.dtype.kind in "iu":
                fixed.append(np.char.mod("%d", c))
            else:
                fixed.append(np.char.mod("%.17e", c))
        if sw.get("FRAGMENTATION", 0):
            nf = np.asarray(self.num_flaws)
            fl = np.char.mod("%.17e", self.flaws)
            with open(path, "w") as fh:
                for i in range(n):
                    parts = [f[i] for f in fixed] + list(fl[i, : nf[i]])
                    fh.write(" ".join(parts) + "\n")
        else:
            table = np.stack(fixed, axis=1)
            with open(path, "w") as fh:
                fh.write("\n".join(" ".join(row) for row in table))
                fh.write("\n")

    def write_inputs(self, directory: str, basename: str = "input.0000") -> tuple:
        os.makedirs(directory, exist_ok=True)
        cfg = os.path.join(directory, "material.cfg")
        with open(cfg, "w") as fh:
            fh.write(self.material_cfg)
        for name, text in self.includes.items():
            with open(os.path.join(directory, name), "w") as fh:
                fh.write(text)
        data = os.path.join(directory, basename)
        self.write_ascii(data)
        return data, cfg


# ---------------------------------------------------------------- helpers
def _cubic_spline_w3(r, h):
    """3-D cubic B-spline with support h (reference: src/kernel.cu:112-153)."""
    q = r / h
    f = 8.0 / np.pi / h**3
    return np.where(q > 1.0, 0.0, np.where(q > 0.5, 2.0 * f * (1.0 - q) ** 3, f * (6.0 * q**3 - 6.0 * q**2 + 1.0)))


def _lattice(lo, hi, delta, dim):
    ax = [np.arange(lo[k], hi[k], delta) for k in range(dim)]
    g = np.meshgrid(*ax, indexing="ij")
    return np.stack([a.ravel() for a in g], axis=1)


_AV = "artificial_viscosity = { alpha = 1.0; beta = 2.0; };"


# ---------------------------------------------------------------- shocktube
def shocktube(dx: float = 5e-4, sml_over_dx: float = 20.0) -> Scenario:
    """1-D Sod tube, x in [-1, 2], spacing dx left of 0.5 and 8 dx right of it."""
    xs, es = [], []
    x = -1.0
    while x < 2.0:
        if x > 0.5:
            x += 8.0 * dx
            e = 2.0
        else:
            x += dx
            e = 2.5
        xs.append(x)
        es.append(e)
    n = len(xs)
    sml = sml_over_dx * dx
    cfg = (
        "materials = (\n  {\n    ID = 0;\n    name = \"ideal gas\";\n"
        f"    sml = {sml:.17e}\n    {_AV}\n"
        "    eos = { type = 9; polytropic_gamma = 1.4; };\n  }\n);\n"
    )
    return Scenario(
        "shocktube", 1, np.asarray(xs)[:, None], np.zeros((n, 1)), np.full(n, dx), cfg,
        e=np.asarray(es), mat=np.zeros(n, dtype=np.int32),
    )


# ---------------------------------------------------------------- sedov
def sedov(delta: float = 0.013, sml_over_delta: float = 0.029 / 0.013) -> Scenario:
    """3-D Sedov-Taylor blast: cubic lattice clipped to a sphere of radius 0.5."""
    R = 0.5
    scale = delta / 0.013
    lo = -1.01 * R
    pts = _lattice([lo] * 3, [1.01 * R] * 3, delta, 3)
    r = np.sqrt((pts**2).sum(axis=1))
    keep = r < R
    pts, r = pts[keep], r[keep]
    n = pts.shape[0]
    m = (4.0 / 3.0) * np.pi * R**3 / n
    e = np.full(n, 1e-8)
    blast = r < 0.06 * R * scale
    e[blast] += _cubic_spline_w3(r[blast], 0.029 * scale)
    sml = sml_over_delta * delta
    cfg = (
        "global = {\n  c_gravity = 6.67408e-11\n}\nmaterials = (\n  {\n    ID = 0\n    name = \"Ideal_gas\"\n"
        f"    sml = {sml:.17e}\n    interactions = 30\n    {_AV}\n"
        "    eos = {\n      type = 9;\n      polytropic_gamma = 1.4\n    };\n  }\n);\n"
    )
    return Scenario(
        "sedov", 3, pts, np.zeros((n, 3)), np.full(n, m), cfg, e=e, mat=np.zeros(n, dtype=np.int32)
    )


def sedov_delta_for(n_target: int) -> float:
    """Lattice spacing giving about n_target particles inside the R=0.5 sphere."""
    return ((4.0 / 3.0) * np.pi * 0.5**3 / n_target) ** (1.0 / 3.0)


# ---------------------------------------------------------------- colliding rings
def rings(dx: float = 0.075, sml_over_dx: float = 0.25 / 0.075) -> Scenario:
    """2-D colliding rubber rings (Monaghan 2000): r in [3,4], centres at x = -/+5."""
    speed, rmin, rmax, off = 0.059, 3.0, 4.0, 5.0
    pts = _lattice([-off, -off], [off, off], dx, 2)
    r = np.sqrt((pts**2).sum(axis=1))
    pts = pts[(r >= rmin) & (r <= rmax)]
    k = pts.shape[0]
    x = np.empty((2 * k, 2))
    v = np.zeros((2 * k, 2))
    x[0::2] = pts + [-off, 0.0]
    x[1::2] = pts + [off, 0.0]
    v[0::2, 0] = speed
    v[1::2, 0] = -speed
    n = 2 * k
    sml = sml_over_dx * dx
    cfg = (
        "materials = (\n  {\n    ID = 0;\n    name = \"TestRubber (Murnaghan)\";\n"
        f"    sml = {sml:.17e}\n    {_AV}\n"
        "    physical_viscosity = { eta = 0.0; zeta = 0.0; };\n"
        "    artificial_stress = {\n      exponent_tensor = 4.;\n      epsilon_stress = 0.3;\n"
        f"      mean_particle_distance = {dx:.17e};\n    }};\n"
        "    eos = {\n      type = 1;\n      n = 1.0;\n      rho_0 = 1.0;\n      rho_limit = 0.0;\n"
        "      shear_modulus = 0.22;\n      bulk_modulus = 1.0;\n      yield_stress = 0.003;\n    };\n  }\n);\n"
    )
    return Scenario(
        "rings", 2, x, v, np.full(n, dx * dx), cfg, rho=np.ones(n), mat=np.zeros(n, dtype=np.int32),
        S=np.zeros((n, 4)),
    )


# ---------------------------------------------------------------- impact
_IMPACT_CFG = """materials = (
  {
    ID = 0
    name = "Basalt Nakamura porous (Tillotson)"
    interactions = 30
    factor_sml_min = 0.1
    factor_sml_max = 10.0
    artificial_viscosity = { alpha = 1.0; beta = 2.0; };
    eos = {
      type = 5
      shear_modulus = 22.7e9
      bulk_modulus = 26.7e9
      till_rho_0 = 2.7e3
      till_A = 26.7e9
      till_B = 26.7e9
      till_E_0 = 487.0e6
      till_E_iv = 4.72e6
      till_E_cv = 18.2e6
      till_a = 0.5
      till_b = 1.5
      till_alpha = 5.0
      till_beta = 5.0
      rho_limit = 0.0
      cs_limit = 3e1
      crushcurve_style = 0
      porjutzi_p_elastic = 2e8
      porjutzi_p_compacted = 2e9
      porjutzi_alpha_0 = 1.25
      porjutzi_alpha_e = 1.25
      cs_porous = 1.5e3
      yield_stress = 1.5e9
      cohesion = 1e5
      friction_angle = 0.98
      cohesion_damaged = 0.0
      friction_angle_damaged = 0.675
    };
  }
);
"""


def _weibull_flaws(volumes, max_flaws, rng, weibull_m=16.0, weibull_k=1e61):
    """Benz-Asphaug flaw assignment: N ln N flaws, i-th threshold (i/(kV))^(1/m), random owner."""
    n = volumes.shape[0]
    vtot = float(volumes.sum())
    total = max(int(np.ceil(n * np.log(max(n, 2)))), n)
    owners = rng.integers(0, n, size=total)
    have = np.zeros(n, dtype=bool)
    have[owners] = True
    missing = np.flatnonzero(~have)
    owners = np.concatenate([owners, missing])
    idx = np.arange(1, owners.shape[0] + 1, dtype=np.float64)
    eps = (idx / (weibull_k * vtot)) ** (1.0 / weibull_m)
    order = np.argsort(owners, kind="stable")
    owners_s, eps_s = owners[order], eps[order]
    counts = np.bincount(owners_s, minlength=n)
    starts = np.concatenate([[0], np.cumsum(counts)[:-1]])
    rank = np.arange(owners_s.shape[0]) - starts[owners_s]
    keep = rank < max_flaws
    flaws = np.zeros((n, max_flaws))
    flaws[owners_s[keep], rank[keep]] = eps_s[keep]
    return np.minimum(counts, max_flaws).astype(np.int32), flaws


def impact(n_target: int = 58402, seed: int = 20240229) -> Scenario:
    """Basalt half-sphere (R = 15 m) hit by a 0.5 m projectile at 6 km/s, 30 deg to the vertical.

    Variable resolution as in the shipped input: full particle density inside
    R_inner = 4 m falling linearly to 1 % at R_outer = 8 m, realised by seeded
    rejection thinning of a cubic lattice; h_i = 2.1 V_i^(1/3).
    """
    rng = np.random.default_rng(seed)
    R, r_in, r_out, f_out = 15.0, 4.0, 8.0, 0.01
    rho0, alpha0 = 2700.0 / 1.25, 1.25

    def keep_fraction(r):
        t = np.clip((r - r_in) / (r_out - r_in), 0.0, 1.0)
        return 1.0 - t * (1.0 - f_out)

    # expected kept volume fraction -> lattice spacing for n_target particles
    rr = np.linspace(0.0, R, 4001)
    shell = 2.0 * np.pi * rr**2  # half sphere
    eff_vol = np.trapezoid(shell * keep_fraction(rr), rr)
    delta = (eff_vol / n_target) ** (1.0 / 3.0)
    # the lattice of the bounding box is walked in slabs of x-planes (x is its slowest index, so the point order -- and with
    # it every random draw -- is that of the whole lattice at once, which needs > 60 GB at 8M particles)
    xs = np.arange(-R, R + delta, delta)
    ys = np.arange(-R, R + delta, delta)
    zs = np.arange(-R, 0.5 * delta, delta)
    shift = 0.5 * delta * np.array([0.37, 0.41, -1.0])  # lattice not aligned with cell planes
    slab = max(1, int(4.0e6 / (len(ys) * len(zs))))
    pts_list, frac_list = [], []
    for i0 in range(0, len(xs), slab):
        g = np.meshgrid(xs[i0: i0 + slab], ys, zs, indexing="ij")
        chunk = np.stack([a.ravel() for a in g], axis=1) + shift
        r = np.sqrt((chunk**2).sum(axis=1))
        sel = (r < R) & (chunk[:, 2] <= 0.0)
        chunk, r = chunk[sel], r[sel]
        f = keep_fraction(r)
        kept = rng.random(chunk.shape[0]) < f
        pts_list.append(chunk[kept])
        frac_list.append(f[kept])
    pts, frac = np.concatenate(pts_list), np.concatenate(frac_list)
    del pts_list, frac_list
    vol = delta**3 / frac
    # projectile: sphere radius 0.5 at (1.5, 0, 2.598) moving with (-3000, 0, -5196.15)
    pc = np.array([1.5, 0.0, 2.598])
    pp = _lattice(pc - 0.5, pc + 0.5 + 0.5 * delta, delta, 3)
    pp = pp[((pp - pc) ** 2).sum(axis=1) < 0.25]
    if pp.shape[0] == 0:
        pp = pc[None, :]
    n_t, n_p = pts.shape[0], pp.shape[0]
    x = np.concatenate([pp, pts])
    v = np.zeros_like(x)
    v[:n_p] = [-3000.0, 0.0, -5196.15]
    volume = np.concatenate([np.full(n_p, delta**3), vol])
    n = n_t + n_p
    h = 2.1 * volume ** (1.0 / 3.0)
    sw = read_switches("impact")
    num_flaws, flaws = _weibull_flaws(volume, sw["MAX_NUM_FLAWS"], rng)
    return Scenario(
        "impact", 3, x, v, rho0 * volume, _IMPACT_CFG,
        rho=np.full(n, rho0), e=np.zeros(n), h=h, mat=np.zeros(n, dtype=np.int32),
        S=np.zeros((n, 9)), d=np.zeros(n), num_flaws=num_flaws, flaws=flaws,
        alpha=np.full(n, alpha0), pold=np.zeros(n),
    )


# ---------------------------------------------------------------- nakamura (von Mises + Grady-Kipp on S)
_NAKAMURA_CFG = """materials = (
  {{
    ID = 0;
    name = "Basalt Nakamura (Tillotson)";
    sml = {sml:.17e};
    artificial_viscosity = {{ alpha = 1.0; beta = 2.0; }};
    eos = {{
      type = 2
      shear_modulus = 22.7e9
      bulk_modulus = 26.7e9
      yield_stress = 3.5e9
      till_rho_0 = 2.7e3
      till_A = 26.7e9
      till_B = 26.7e9
      till_E_0 = 487e6
      till_E_iv = 4.72e6
      till_E_cv = 18.2e6
      till_a = 0.5
      till_b = 1.5
      till_alpha = 5.0
      till_beta = 5.0
      rho_limit = 0.0
    }};
  }},
  {{
    ID = 1;
    name = "Lucite";
    sml = {sml:.17e};
    artificial_viscosity = {{ alpha = 1.0; beta = 2.0; }};
    eos = {{
      type = 2
      shear_modulus = 7.3e7
      bulk_modulus = 10.1e9
      yield_stress = 1e7
      till_rho_0 = 1.18e3
      till_A = 26.7e9
      till_B = 26.7e9
      till_E_0 = 487e6
      till_E_iv = 4.72e6
      till_E_cv = 18.2e6
      till_a = 0.5
      till_b = 1.5
      till_alpha = 5.0
      till_beta = 5.0
      rho_limit = 0.0
    }};
  }}
);
"""


def nakamura(n_target: int = 905000, seed: int = 1991) -> Scenario:
    """Nakamura & Fujiwara (1991): 0.2 g lucite bullet hitting a 3 cm basalt sphere at 3.2 km/s, 30 degrees
    (reference: test_cases/nakamura/input/target.c, projectile.c, create_input.sh).  Cubic lattices, fixed
    sml = 2.8 lattice spacings, Weibull flaws (k = 5e34, m = 8.5) on the target only, at most 28 per particle."""
    R = 3e-2
    delta = ((4.0 / 3.0) * np.pi * R**3 / n_target) ** (1.0 / 3.0)
    sml = 2.8 * delta
    rho_t, rho_p = 2.7e3, 1.18e3
    span = R + 2.0 * delta
    pts = _lattice([-span] * 3, [span] * 3, delta, 3) + delta * np.array([0.11, 0.23, 0.31])
    pts = pts[(pts**2).sum(axis=1) <= R * R]
    pts[:, 0] += np.sin(np.pi / 6.0) * R
    rp = (0.75 / np.pi * 0.2e-3 / rho_p) ** (1.0 / 3.0)
    pp = _lattice([-rp - delta] * 3, [rp + delta] * 3, delta, 3)
    pp = pp[(pp**2).sum(axis=1) <= rp * rp]
    if pp.shape[0] == 0:
        pp = np.zeros((1, 3))
    pp[:, 2] += delta * 3.04 + R + 2.0 * rp
    n_p, n_t = pp.shape[0], pts.shape[0]
    n = n_p + n_t
    x = np.concatenate([pp, pts])
    v = np.zeros_like(x)
    v[:n_p, 2] = -3.2e3
    rho = np.concatenate([np.full(n_p, rho_p), np.full(n_t, rho_t)])
    mat = np.concatenate([np.ones(n_p, dtype=np.int32), np.zeros(n_t, dtype=np.int32)])
    rng = np.random.default_rng(seed)
    sw = read_switches("nakamura")
    nf_t, fl_t = _weibull_flaws(np.full(n_t, delta**3), 28, rng, weibull_m=8.5, weibull_k=5e34)
    num_flaws = np.concatenate([np.zeros(n_p, dtype=np.int32), nf_t])
    flaws = np.zeros((n, sw["MAX_NUM_FLAWS"]))
    flaws[n_p:, :28] = fl_t
    return Scenario(
        "nakamura", 3, x, v, rho * delta**3, _NAKAMURA_CFG.format(sml=sml),
        rho=rho, e=np.zeros(n), mat=mat, S=np.zeros((n, 9)), d=np.zeros(n), num_flaws=num_flaws, flaws=flaws,
    )


# ---------------------------------------------------------------- giant collisions
_IRON_TILL = """till_rho_0 = 7.8e3
till_A = 128.0e9
till_B = 105.0e9
till_E_0 = 9.5e6
till_E_iv = 2.4e6
till_E_cv = 8.67e6
till_a = 0.5
till_b = 1.5
till_alpha = 5.0
till_beta = 5.0
rho_limit = 0.9
cs_limit = 40.0
"""
_GRANITE_TILL = """till_rho_0 = 2.68e3
till_A = 1.8e10
till_B = 1.8e10
till_E_0 = 1.6e7
till_E_iv = 3.5e6
till_E_cv = 1.8e7
till_a = 0.5
till_b = 1.3
till_alpha = 5.0
till_beta = 5.0
rho_limit = 0.9
cs_limit = 30.0
"""


def _giant_cfg(sml: float) -> str:
    def mat(i, name, floor, shear, bulk, ys, coh, inc):
        return (
            f"  {{\n    ID = {i}\n    name = \"{name}\";\n    sml = {sml:.17e}\n    interactions = 30\n    {_AV}\n"
            f"    density_floor = {floor}\n    eos = {{\n      type = 2\n      shear_modulus = {shear}\n"
            f"      bulk_modulus = {bulk}\n      yield_stress = {ys}\n      cohesion = {coh}\n"
            "      friction_angle = 1.11\n      cohesion_damaged = 0.0\n      friction_angle_damaged = 0.675\n"
            f"      @include \"{inc}\"\n    }};\n  }}"
        )

    return (
        "materials = (\n"
        + mat(0, "Iron", "100.", "105e9", "113.5e9", "10.5e9", "90.0e7", "iron.till.cfg")
        + ",\n"
        + mat(1, "Granite", "10.", "2.7e10", "5.0e10", "1.5e9", "90.0e6", "granite.till.cfg")
        + "\n);\n"
    )


# ANEOS variant of the giant collision (north_star: "the EOS including tabulated ANEOS").
# The real M-ANEOS tables are not shipped with the reference, so a table in the reference's format
# (src/aneos.cu:119-181: three header lines, then n_rho x n_e rows "rho e p T cs entropy phase", rho in the
# outer loop) is synthesised from a Tillotson-like closed form.  The grid is deliberately narrower than the
# particle states so that every branch of the lookup runs: rho below / above the table (edge-cell
# extrapolation), e below the table (clamped) and e above it (ideal-gas fallback, src/pressure.cu / aneos.cu).
_ANEOS_PARAMS_BASALT = dict(rho0=2700.0, A=2.67e10, B=2.67e10, E0=4.87e8, a=0.5, b=1.5, bulk_cs=3144.0)
_ANEOS_PARAMS = {
    "basalt": _ANEOS_PARAMS_BASALT,
    "iron": dict(rho0=7800.0, A=1.28e11, B=1.05e11, E0=9.5e6, a=0.5, b=1.5, bulk_cs=4050.0),
    "granite": dict(rho0=2680.0, A=1.8e10, B=1.8e10, E0=1.6e7, a=0.5, b=1.3, bulk_cs=2590.0),
}
ANEOS_N_RHO, ANEOS_N_E = 40, 36


def aneos_table_text(name: str) -> str:
    q = _ANEOS_PARAMS[name]
    rho = q["rho0"] * np.geomspace(0.62, 1.10, ANEOS_N_RHO)
    e = np.geomspace(2.0e4, 1.0e7, ANEOS_N_E)
    lines = [f"# synthetic ANEOS-format table for {name} (miluphcuda_b200/scenarios.py)",
             f"# n_rho = {ANEOS_N_RHO}, n_e = {ANEOS_N_E}", "# rho e p T cs entropy phase"]
    for r in rho:
        eta = r / q["rho0"]
        mu = eta - 1.0
        for en in e:
            om = en / (q["E0"] * eta * eta) + 1.0
            pres = (q["a"] + q["b"] / om) * r * en + q["A"] * mu + q["B"] * mu * mu
            cs2 = q["a"] * en + q["b"] * en / (om * om) * (3.0 * om - 2.0) + (q["A"] + 2.0 * q["B"] * mu) / r
            cs = np.sqrt(max(cs2, (0.05 * q["bulk_cs"]) ** 2))
            temp = 300.0 + en / 800.0
            lines.append(f"{r:.16e} {en:.16e} {pres:.16e} {temp:.6e} {cs:.16e} {1000.0 + 0.1 * temp:.6e} 1")
    return "\n".join(lines) + "\n"


def _giant_aneos_cfg(sml: float) -> str:
    def mat(i, name, key, floor):
        q = _ANEOS_PARAMS[key]
        return (
            f"  {{\n    ID = {i}\n    name = \"{name}\";\n    sml = {sml:.17e}\n    interactions = 30\n    {_AV}\n"
            f"    density_floor = {floor}\n    eos = {{\n      type = 7\n      table_path = \"{key}.aneos.table\"\n"
            f"      n_rho = {ANEOS_N_RHO}\n      n_e = {ANEOS_N_E}\n      aneos_rho_0 = {q['rho0']}\n"
            f"      aneos_bulk_cs = {q['bulk_cs']}\n      aneos_gamma = 1.4\n      rho_limit = 0.9\n    }};\n  }}"
        )

    return "materials = (\n" + mat(0, "Iron", "iron", "100.") + ",\n" + mat(1, "Granite", "granite", "10.") + "\n);\n"


def giant(n_target: int = 59899, solid: bool = False, seed: int = 7, aneos: bool = False) -> Scenario:
    """Two differentiated bodies (iron core id 0, granite mantle id 1) about to collide.

    Geometry and kinematics from the shipped run (target R = 1.70e6 m, core
    8.63e5 m; projectile R = 8.03e5 m, core 4.1e5 m; v = 2 v_esc at 45 deg,
    three touching distances apart).  Particles sit on cubic lattices clipped to
    each sphere; rho, e follow a smooth self-compression profile (monotone in
    r), m_i = rho(r_i) delta^3, uniform h ~ N^(-1/3) from 171776 m at 59899.
    """
    Rt, Rtc, Rp, Rpc = 1.70e6, 8.63e5, 8.03e5, 4.10e5
    vol = (4.0 / 3.0) * np.pi * (Rt**3 + Rp**3)
    delta = (vol / n_target) ** (1.0 / 3.0)
    sml = 171776.0 * (59899.0 / n_target) ** (1.0 / 3.0)

    def body(R, Rc, centre, vel):
        pts = _lattice([-R] * 3, [R + delta] * 3, delta, 3) + delta * np.array([0.13, 0.29, 0.47])
        r = np.sqrt((pts**2).sum(axis=1))
        sel = r < R
        pts, r = pts[sel], r[sel]
        core = r < Rc
        s = 1.0 - (r / R) ** 2
        rho = np.where(core, 7800.0 * (1.0 + 0.08 * s), 2680.0 * (1.0 + 0.05 * s))
        e = np.where(core, 4.0e5 * s + 1.0e4, 3.0e5 * s + 1.0e4)
        mat = np.where(core, 0, 1).astype(np.int32)
        v = np.broadcast_to(np.asarray(vel, dtype=np.float64), pts.shape).copy()
        return pts + np.asarray(centre), v, rho * delta**3, rho, e, mat

    G = 6.67408e-11
    mt = (4.0 / 3.0) * np.pi * (7800.0 * Rtc**3 + 2680.0 * (Rt**3 - Rtc**3))
    mp = (4.0 / 3.0) * np.pi * (7800.0 * Rpc**3 + 2680.0 * (Rp**3 - Rpc**3))
    vesc = np.sqrt(2.0 * G * (mt + mp) / (Rt + Rp))
    dist = 3.0 * (Rt + Rp)
    ang = np.pi / 4.0
    pos_p = np.array([dist * np.sin(ang), dist * np.cos(ang), 0.0])
    vel_p = np.array([0.0, -2.0 * vesc * 0.5, 0.0])  # slowed to the shipped approach speed scale
    parts = [body(Rt, Rtc, [0.0, 0.0, 0.0], [0.0, 0.0, 0.0]), body(Rp, Rpc, pos_p, vel_p)]
    x, v, m, rho, e, mat = [np.concatenate([p[k] for p in parts]) for k in range(6)]
    n = x.shape[0]
    cfg = _giant_cfg(sml)
    inc = {"iron.till.cfg": _IRON_TILL, "granite.till.cfg": _GRANITE_TILL}
    if aneos:
        cfg = _giant_aneos_cfg(sml)
        inc = {"iron.aneos.table": aneos_table_text("iron"), "granite.aneos.table": aneos_table_text("granite")}
    if not solid:
        return Scenario("giant_hydro", 3, x, v, m, cfg, rho=rho, e=e, mat=mat, includes=inc, selfgravity=True)
    rng = np.random.default_rng(seed)
    sw = read_switches("giant_solid")
    num_flaws, flaws = _weibull_flaws(np.full(n, delta**3), sw["MAX_NUM_FLAWS"], rng, weibull_m=16.0, weibull_k=1e61)
    return Scenario(
        "giant_solid", 3, x, v, m, cfg, rho=rho, e=e, mat=mat, S=np.zeros((n, 9)), d=np.zeros(n),
        num_flaws=num_flaws, flaws=flaws, includes=inc, selfgravity=True,
    )


def stir(sc: Scenario, seed: int = 1234) -> Scenario:
    """Perturb a pristine initial condition so that every term of the RHS is exercised.

    Step-0 states of the shipped scenarios are at rest / stress-free, which makes most
    rates vanish.  This applies seeded, bounded perturbations to positions (a fraction of
    the local spacing), velocities, density, energy, deviatoric stress (non-symmetric on
    purpose), damage and distension, staying inside each EOS's valid range.
    """
    rng = np.random.default_rng(seed)
    n, dim = sc.n, sc.dim
    u = lambda *shape: rng.uniform(-1.0, 1.0, size=shape)
    sw = sc.switches()
    if sc.h is not None:
        spacing = sc.h / 2.1
    else:
        m = re.search(r"sml\s*=\s*([0-9.eE+-]+)", sc.material_cfg)
        sml = float(m.group(1))
        ratio = {"shocktube": 20.0, "sedov": 0.029 / 0.013, "rings": 0.25 / 0.075}.get(sc.config, 2.1)
        spacing = np.full(n, sml / ratio)
    if sc.config == "shocktube":
        sc.x = np.sort(sc.x + 0.2 * spacing[:, None] * u(n, dim), axis=0)
    else:
        sc.x = sc.x + 0.15 * spacing[:, None] * u(n, dim)
    cfgname = sc.config
    if cfgname in ("shocktube", "sedov"):
        sc.v = sc.v + 0.3 * u(n, dim)
        sc.e = sc.e * (1.0 + 0.2 * u(n)) + (0.05 if cfgname == "sedov" else 0.0)
        sc.m = sc.m * (1.0 + 0.05 * u(n))
    elif cfgname == "rings":
        sc.v = sc.v + 0.02 * u(n, dim)
        sc.rho = sc.rho * (1.0 + 0.05 * u(n))
        sc.S = 0.05 * u(n, dim * dim)
    elif cfgname == "impact":
        sc.v = sc.v + 60.0 * u(n, dim)
        sc.rho = sc.rho * (1.0 + 0.04 * u(n))
        sc.e = 2.0e4 * (1.0 + u(n)) + 1.0e3
        hot = rng.random(n) < 0.03          # a few particles in the intermediate / vapour regimes
        sc.e[hot] = rng.uniform(4.0e6, 3.0e7, size=int(hot.sum()))
        sc.S = 2.0e7 * u(n, dim * dim)
        sc.d = np.clip(0.45 * rng.random(n) - 0.05, 0.0, 1.0)
        sc.alpha = 1.0 + 0.25 * rng.random(n)
        sc.alpha[rng.random(n) < 0.05] = 1.0
        sc.alpha[rng.random(n) < 0.01] = 0.995
        sc.h = sc.h * (1.0 + 0.1 * u(n))
    elif cfgname == "nakamura":
        sc.v = sc.v + 40.0 * u(n, dim)
        sc.rho = sc.rho * (1.0 + 0.03 * u(n))
        sc.e = 2.0e4 * (1.0 + u(n)) + 1.0e3
        hot = rng.random(n) < 0.03
        sc.e[hot] = rng.uniform(4.0e6, 3.0e7, size=int(hot.sum()))
        sc.S = 3.0e9 * u(n, dim * dim)          # straddles the von Mises yield surface (Y = 3.5e9 / 1e7)
        sc.d = np.clip(0.6 * rng.random(n) - 0.05, 0.0, 1.0)
    else:  # giant_hydro / giant_solid
        sc.v = sc.v + 150.0 * u(n, dim)
        sc.rho = sc.rho * (1.0 + 0.03 * u(n))
        sc.e = sc.e * (1.0 + 0.3 * u(n))
        hot = rng.random(n) < 0.03
        sc.e[hot] = rng.uniform(2.5e6, 3.0e7, size=int(hot.sum()))
        sc.rho[hot] *= rng.uniform(0.5, 1.0, size=int(hot.sum()))
        if sw.get("SOLID", 0):
            sc.S = 1.0e8 * u(n, dim * dim)
            sc.d = np.clip(0.45 * rng.random(n) - 0.05, 0.0, 1.0)
    return sc


def with_ignored_material(sc: Scenario) -> Scenario:
    """Adds a material whose eos.type is EOS_TYPE_IGNORE (-1) and hands it to every 11th particle: the
    `matEOS[materialId] == EOS_TYPE_IGNORE` half of the reference's deactivation tests (src/boundary.cu:98-145,
    src/internal_forces.cu:145,271, src/density.cu:66-107).  The `materialId == -1` half is produced by the dump
    hook (REF_DEACTIVATE), which mimics what BoundaryConditionsAfterIntegratorStep does in a run."""
    n_mat = len(re.findall(r"\bID\s*=", sc.material_cfg))
    m = re.search(r"sml\s*=\s*([0-9.eE+-]+)", sc.material_cfg)
    sml_line = f"    sml = {m.group(1)}\n" if m else ""
    extra = (f"  {{\n    ID = {n_mat}\n    name = \"ignored\"\n{sml_line}    {_AV}\n"
             "    eos = {\n      type = -1\n    };\n  }\n);")
    head = sc.material_cfg.rstrip()
    assert head.endswith(");")
    sc.material_cfg = head[:-2].rstrip() + ",\n" + extra + "\n"
    sc.mat = sc.mat.copy()
    sc.mat[3::11] = n_mat
    return sc


def with_tabulated_matrix(sc: Scenario) -> Scenario:
    """The impact scenario with the p-alpha model on a TABULATED matrix EOS (eos.type = 13, EOS_TYPE_JUTZI_ANEOS;
    reference: src/pressure.cu:313-363, src/soundspeed.cu:167-206): the Tillotson keys give way to a synthetic
    ANEOS-format table, narrower than the particle states so that the edge-cell extrapolation, the cold-curve clamp and
    the ideal-gas fallback of the lookup all run."""
    q = _ANEOS_PARAMS["basalt"]
    cfg = sc.material_cfg.replace("type = 5", "type = 13")
    cfg = cfg.replace("till_rho_0 = 2.7e3", f"table_path = \"basalt.aneos.table\"\n      n_rho = {ANEOS_N_RHO}\n      n_e = {ANEOS_N_E}\n"
                      f"      aneos_rho_0 = {q['rho0']}\n      aneos_bulk_cs = {q['bulk_cs']}\n      aneos_gamma = 1.4\n      till_rho_0 = 2.7e3")
    sc.material_cfg = cfg
    sc.includes = dict(sc.includes, **{"basalt.aneos.table": aneos_table_text("basalt")})
    return sc


def with_crush_curve(sc: Scenario, style: int) -> Scenario:
    """The impact scenario on another crush curve of the p-alpha model (reference: src/pressure.cu:365-440)."""
    extra = f"crushcurve_style = {style}"
    if style == 1:
        extra += "\n      porjutzi_p_transition = 6e8\n      porjutzi_alpha_t = 1.1\n      porjutzi_n1 = 12.0\n      porjutzi_n2 = 3.0"
        sc.material_cfg = sc.material_cfg.replace("porjutzi_alpha_e = 1.25", "porjutzi_alpha_e = 1.2")
    sc.material_cfg = sc.material_cfg.replace("crushcurve_style = 0", extra)
    return sc


def make(config: str, n: int | None = None, stirred: bool = False) -> Scenario:
    """Scenario by config name at roughly n particles (None = the shipped resolution)."""
    if stirred and not (config.endswith("_ignore") or config.startswith("impact_crush") or config == "impact_aneos"):   # those variants are stirred already
        return stir(make(config, n))
    if config == "shocktube":
        return shocktube() if n is None else shocktube(dx=5e-4 * 3376.0 / n)
    if config == "sedov":
        return sedov() if n is None else sedov(delta=sedov_delta_for(n))
    if config == "rings":
        return rings() if n is None else rings(dx=0.075 * np.sqrt(7800.0 / n))
    if config == "impact":
        return impact() if n is None else impact(n_target=n)
    if config == "giant_hydro":
        return giant() if n is None else giant(n_target=n)
    if config == "giant_aneos":   # the giant_hydro switch set (same library) with tabulated-EOS materials
        return giant(aneos=True) if n is None else giant(n_target=n, aneos=True)
    if config == "giant_solid":
        return giant(solid=True) if n is None else giant(n_target=n, solid=True)
    if config == "nakamura":
        return nakamura() if n is None else nakamura(n_target=n)
    if config.endswith("_ignore"):
        return with_ignored_material(stir(make({"giant_ignore": "giant_hydro"}.get(config, config[:-7]), n)))
    if config == "impact_aneos":
        return with_tabulated_matrix(stir(make("impact", n)))
    if config.startswith("impact_crush"):
        return with_crush_curve(stir(make("impact", n)), int(config[-1]))
    raise ValueError(f"unknown config {config!r}")
