"""Particle buffers in the caller's layout for a generated scenario.

Mirrors what the reference's reader + init_values() leave in `p_host`/`p_device`
before the first rightHandSide() (reference: src/io.cu:162-207, 1013-1314;
src/memory_handling.cu:1108): inputs from the file, h from material.cfg's `sml`
unless read per particle, h0 = h, numActiveFlaws from the initial damage,
everything else zero.
"""
from __future__ import annotations

import numpy as np

from . import api


def scenario_arrays(sc, materials: api.MaterialTables | None = None):
    """(arrays, meta): {field: numpy array} for every member of the switch set, plus view scalars."""
    sw = sc.switches()
    n, dim = sc.n, sc.dim
    max_flaws = sw.get("MAX_NUM_FLAWS", 1)
    p_fields, rhs_fields = api.fields_for(sw, sc.selfgravity)
    arrays = {}
    for name in p_fields + rhs_fields:
        dtype = np.int32 if name in api.INT_FIELDS else np.float64
        arrays[name] = np.zeros(api.field_shape(name, n, dim, max_flaws), dtype=dtype)
    for k, ax in enumerate("xyz"[:dim]):
        arrays[ax][:] = sc.x[:, k]
        arrays["v" + ax][:] = sc.v[:, k]
    arrays["m"][:] = sc.m
    arrays["materialId"][:] = sc.mat
    if sc.rho is not None:
        arrays["rho"][:] = sc.rho
    if sc.e is not None:
        arrays["e"][:] = sc.e
    if sc.S is not None:
        arrays["S"][:] = sc.S.reshape(-1)
    if sc.d is not None:
        arrays["d"][:] = sc.d
        arrays["numFlaws"][:] = sc.num_flaws
        arrays["flaws"][:] = sc.flaws.reshape(-1)
        arrays["numActiveFlaws"][:] = np.minimum(np.ceil(sc.num_flaws * sc.d ** dim), sc.num_flaws).astype(np.int32)
    if sc.alpha is not None:
        arrays["alpha_jutzi"][:] = sc.alpha
        arrays["pold"][:] = sc.pold
    if sc.h is not None:
        arrays["h"][:] = sc.h
    elif materials is not None:
        arrays["h"][:] = materials.table("matSml")[arrays["materialId"]]
    arrays["h0"][:] = arrays["h"]
    meta = dict(n=n, max_num_flaws=max_flaws, selfgravity=sc.selfgravity, theta=sc.theta)
    return arrays, meta


def write_material_files(sc, directory: str) -> str:
    """material.cfg (+ @include files) of a scenario; returns the cfg path."""
    import os
    os.makedirs(directory, exist_ok=True)
    path = os.path.join(directory, "material.cfg")
    with open(path, "w") as fh:
        fh.write(sc.material_cfg)
    for name, text in sc.includes.items():
        with open(os.path.join(directory, name), "w") as fh:
            fh.write(text)
    return path
