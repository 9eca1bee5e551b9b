"""CPU: the C-ABI library loads, exports every symbol include/b200sph.h declares, the ctypes
mirror matches the header, and the host-side material reader reproduces the reference's
defaults and libconfig typing rules.  No compute call is made (no GPU needed)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import common
from miluphcuda_b200 import api, build, scenarios

HEADER = os.path.join(common.REPO, "include", "b200sph.h")


def header_text():
    with open(HEADER) as fh:
        return fh.read()


def test_header_declares_what_binding_uses():
    text = header_text()
    declared = set(re.findall(r"\b(b200sph_[a-z_0-9]+)\s*\(", text))
    assert set(api.exported_symbols()) == declared


@pytest.mark.parametrize("config", common.CONFIGS)
def test_library_loads_and_exports_symbols(config):
    lib = api.load_library(config)
    for name in api.exported_symbols():
        assert hasattr(lib, name), name
    assert lib.b200sph_abi_version() == 1
    assert lib.b200sph_config_name().decode() == config
    sw = scenarios.read_switches(config)
    for key, val in sw.items():
        if key.startswith("_"):
            continue
        assert lib.b200sph_switch_value(key.encode()) == val, key
    assert lib.b200sph_switch_value(b"NO_SUCH_SWITCH") == -999
    assert lib.b200sph_switch_hash() != 0


def test_switch_hash_distinguishes_configs():
    hashes = {api.load_library(c).b200sph_switch_hash() for c in common.CONFIGS}
    assert len(hashes) == len(common.CONFIGS)


def _struct_members(text, struct_name):
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct_name, struct_name), text, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        decl = re.sub(r"^(const\s+)?(double|float|int|int64_t|b200sph_particle_arrays)\s+", "", decl)
        for part in decl.split(","):
            names.append(part.strip().lstrip("*").strip())
    return names


def test_ctypes_mirror_matches_header_layout():
    text = header_text()
    assert tuple(_struct_members(text, "b200sph_particle_arrays")) == api.PARTICLE_FIELDS
    assert _struct_members(text, "b200sph_view") == [n for n, _ in api.View._fields_]
    assert _struct_members(text, "b200sph_materials") == [n for n, _ in api.Materials._fields_]
    assert _struct_members(text, "b200sph_stats") == [n for n, _ in api.Stats._fields_]


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(build, "lib_path", lambda c: "/nonexistent/libb200sph_%s.so" % c)
    api._LIBS.pop("sedov", None)
    with pytest.raises(FileNotFoundError):
        api.load_library("sedov")
    monkeypatch.undo()
    api.load_library("sedov")


def test_material_defaults_match_reference_rules(tmp_path):
    sc = scenarios.make("impact", 200)
    cfg = tmp_path / "material.cfg"
    cfg.write_text(sc.material_cfg)
    m = api.MaterialTables("impact", str(cfg))
    t = m.table
    assert t("matEOS")[0] == 5
    assert t("matAlpha")[0] == 1.0 and t("matBeta")[0] == 2.0
    assert t("matN")[0] == 1.0                      # default n (src/config_parameter.cu:763-765)
    assert t("matEnergyFloor")[0] == -1e30          # src/config_parameter.cu:843-844
    assert t("matDensityFloor")[0] == pytest.approx(27.0)   # 0.01 * till_rho_0
    assert t("matInternalFriction")[0] == pytest.approx(np.tan(0.98))
    assert t("matInternalFrictionDamaged")[0] == pytest.approx(np.tan(0.675))
    k, mu = 26.7e9, 22.7e9
    assert t("matYoungModulus")[0] == pytest.approx(9 * k * mu / (3 * k + mu))
    assert t("matcs_porous")[0] == 1.5e3 and t("matcsLimit")[0] == 30.0
    assert t("mat_f_sml_min")[0] == 0.1 and t("mat_f_sml_max")[0] == 10.0
    assert m.grav_const == 6.67408e-11


def test_material_cs_porous_default_uses_unread_till_A(tmp_path):
    # reference quirk: the cs_porous default is evaluated before till_A is read -> 0
    text = scenarios.make("impact", 200).material_cfg.replace("cs_porous = 1.5e3", "")
    cfg = tmp_path / "material.cfg"
    cfg.write_text(text)
    m = api.MaterialTables("impact", str(cfg))
    assert m.table("matcs_porous")[0] == 0.0


def test_libconfig_typing_rules(tmp_path):
    # an integer literal is not a float for config_setting_lookup_float (real libconfig behaviour)
    cfg = tmp_path / "material.cfg"
    cfg.write_text('materials = ( { ID = 0; sml = 1; eos = { type = 9; polytropic_gamma = 1.4 } } );\n')
    m = api.MaterialTables("sedov", str(cfg))
    assert m.table("matSml")[0] == 0.0
    assert m.table("matPolytropicGamma")[0] == 1.4
    cfg.write_text('// c\n# c\n/* c */ materials : ( { ID : 0, sml : 2.5e-1, eos : { type : 9 } } )')
    m = api.MaterialTables("sedov", str(cfg))
    assert m.table("matSml")[0] == 0.25


def test_material_errors(tmp_path):
    cfg = tmp_path / "material.cfg"
    cfg.write_text("materials = ( { ID = 1; eos = { type = 9 } } );")
    with pytest.raises(api.B200SphError):
        api.MaterialTables("sedov", str(cfg))
    cfg.write_text("materials = ( { ID = 0 } );")
    with pytest.raises(api.B200SphError):
        api.MaterialTables("sedov", str(cfg))
    with pytest.raises(api.B200SphError):
        api.MaterialTables("sedov", str(tmp_path / "missing.cfg"))


def test_include_directive_and_two_materials(tmp_path):
    sc = scenarios.make("giant_hydro", 300)
    (tmp_path / "material.cfg").write_text(sc.material_cfg)
    for name, text in sc.includes.items():
        (tmp_path / name).write_text(text)
    m = api.MaterialTables("giant_hydro", str(tmp_path / "material.cfg"))
    assert list(m.table("matEOS")) == [2, 2]
    assert list(m.table("matTillRho0")) == [7.8e3, 2.68e3]
    assert list(m.table("matDensityFloor")) == [100.0, 10.0]
    assert list(m.table("matRhoLimit")) == [0.9, 0.9]


def test_aneos_table_reader(tmp_path):
    n_rho, n_e = 4, 3
    rho = np.array([1.0, 2.0, 4.0, 8.0])
    e = np.array([10.0, 20.0, 40.0])
    lines = ["# header 1", "# header 2", "# header 3"]
    for r in rho:
        for en in e:
            lines.append(f"{r:e} {en:e} {r * en:e} 300.0 {np.sqrt(r + en):e} 0.0 1")
    (tmp_path / "table.dat").write_text("\n".join(lines) + "\n")
    (tmp_path / "material.cfg").write_text(
        'materials = ( { ID = 0; sml = 1.0; eos = { type = 7; table_path = "table.dat"; n_rho = 4; n_e = 3; '
        'aneos_rho_0 = 2.0; aneos_bulk_cs = 5.0; aneos_gamma = 1.5; } } );')
    m = api.MaterialTables("giant_hydro", str(tmp_path / "material.cfg"))
    assert np.allclose(m.table("aneos_rho"), rho) and np.allclose(m.table("aneos_e"), e)
    assert np.allclose(m.table("aneos_p").reshape(n_rho, n_e), np.outer(rho, e))
    assert m.table("matcsLimit")[0] == pytest.approx(0.05)
    assert m.table("matDensityFloor")[0] == pytest.approx(0.02)
