/*
 * materials.c -- host side of the material tables: material.cfg -> b200sph_materials.
 *
 * Mirrors what the reference's transferMaterialsToGPU() reads and derives
 * (reference: src/config_parameter.cu:346-878): same keys, same defaults
 * (alpha = 1, beta = 2, rho_limit = 0.9, n = 1, cs_limit, density_floor,
 * energy_floor = -1e30, ...), same derived quantities (young_modulus,
 * internal_friction = tan(friction_angle)), in the same order -- the order
 * matters once: the default for cs_porous is evaluated BEFORE till_A is read
 * (config_parameter.cu:710-720 vs :744), so it is 0 unless cs_porous is given.
 * ANEOS-format tables are read like initialize_aneos_eos_basic()
 * (src/aneos.cu:119-181): three comment lines, then n_rho*n_e rows
 * "rho e p T cs ...", rho in the outer loop.
 */
#include "switches.h"
#include "libconfig_lite.h"
#include "../../include/b200sph.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    b200sph_materials view;   /* must be first: the public pointer is a pointer to this */
    void *blocks[96];
    int nblocks;
} material_store;

static void *store_alloc(material_store *s, size_t count, size_t elem)
{
    void *ptr = calloc(count ? count : 1, elem);
    if (ptr && s->nblocks < (int)(sizeof(s->blocks) / sizeof(s->blocks[0]))) s->blocks[s->nblocks++] = ptr;
    return ptr;
}

void b200sph_materials_free(b200sph_materials *m)
{
    material_store *s = (material_store *)m;
    int i;
    if (!s) return;
    for (i = 0; i < s->nblocks; i++) free(s->blocks[i]);
    free(s);
}

static int fail(char *err, size_t errlen, const char *fmt, const char *a, int b)
{
    if (err && errlen) snprintf(err, errlen, fmt, a, b);
    return B200SPH_ERR_BAD_ARGUMENT;
}

static int read_aneos_table(const char *path, int n_rho, int n_e, double *rho, double *e, double *ptab, double *cstab,
                            char *err, size_t errlen)
{
    FILE *f = fopen(path, "r");
    char line[4096];
    int i, j;
    if (!f) return fail(err, errlen, "cannot open ANEOS table '%s' (material %d)", path, 0);
    for (i = 0; i < 3; i++)
        if (!fgets(line, sizeof(line), f)) { fclose(f); return fail(err, errlen, "short ANEOS table '%s' (line %d)", path, i); }
    for (i = 0; i < n_rho; i++)
        for (j = 0; j < n_e; j++) {
            double r, en, pp, tt, cc;
            if (!fgets(line, sizeof(line), f) || sscanf(line, "%le %le %le %le %le", &r, &en, &pp, &tt, &cc) != 5) {
                fclose(f);
                return fail(err, errlen, "bad row in ANEOS table '%s' (data line %d)", path, i * n_e + j + 1);
            }
            rho[i] = r;
            e[j] = en;
            ptab[(size_t)i * n_e + j] = pp;
            cstab[(size_t)i * n_e + j] = cc;
        }
    fclose(f);
    return 0;
}

int b200sph_materials_load(const char *cfg_path, b200sph_materials **out, double *grav_const, char *err, size_t errlen)
{
    config_t cfg;
    config_setting_t *materials, *global;
    material_store *s;
    b200sph_materials *m;
    int n, i, max_id = 0, rc = 0;
    double g = 6.67408e-11;
    char dir[1024];

    if (!cfg_path || !out) return B200SPH_ERR_BAD_ARGUMENT;
    *out = NULL;
    config_init(&cfg);
    if (!config_read_file(&cfg, cfg_path)) {
        if (err && errlen)
            snprintf(err, errlen, "Error reading config file %s: %s (%s:%d)", cfg_path, config_error_text(&cfg),
                     config_error_file(&cfg), config_error_line(&cfg));
        config_destroy(&cfg);
        return B200SPH_ERR_BAD_ARGUMENT;
    }
    {
        const char *slash = strrchr(cfg_path, '/');
        size_t dl = slash ? (size_t)(slash - cfg_path) + 1 : 0;
        if (dl >= sizeof(dir)) dl = 0;
        memcpy(dir, cfg_path, dl);
        dir[dl] = '\0';
    }
    global = config_lookup(&cfg, "global");
    if (global) config_setting_lookup_float(global, "c_gravity", &g);
    if (grav_const) *grav_const = g;

    materials = config_lookup(&cfg, "materials");
    if (!materials) {
        config_destroy(&cfg);
        return fail(err, errlen, "no 'materials' list in %s%.0d", cfg_path, 0);
    }
    n = config_setting_length(materials);
    for (i = 0; i < n; i++) {
        int id;
        if (!config_setting_lookup_int(config_setting_get_elem(materials, i), "ID", &id)) {
            config_destroy(&cfg);
            return fail(err, errlen, "Found material without ID in config file %s%.0d", cfg_path, 0);
        }
        if (id > max_id) max_id = id;
    }
    if (max_id != n - 1) {
        config_destroy(&cfg);
        return fail(err, errlen, "Material-IDs in config file %s have to be 0, 1, 2,...%.0d", cfg_path, 0);
    }

    s = (material_store *)calloc(1, sizeof(*s));
    m = &s->view;
    m->n_materials = n;
#define DTAB(name) double *name = (double *)store_alloc(s, n, sizeof(double)); m->name = name
#define ITAB(name) int *name = (int *)store_alloc(s, n, sizeof(int)); m->name = name
    ITAB(matEOS); DTAB(matSml); DTAB(mat_f_sml_min); DTAB(mat_f_sml_max); DTAB(matAlpha); DTAB(matBeta);
    DTAB(matPolytropicK); DTAB(matPolytropicGamma); DTAB(matIsothermalSoundSpeed);
    DTAB(matBulkmodulus); DTAB(matShearmodulus); DTAB(matYoungModulus); DTAB(matYieldStress);
    DTAB(matRho0); DTAB(matN); DTAB(matRhoLimit); DTAB(matcsLimit);
    DTAB(matTillRho0); DTAB(matTillA); DTAB(matTillB); DTAB(matTillE0); DTAB(matTillEiv); DTAB(matTillEcv);
    DTAB(matTilla); DTAB(matTillb); DTAB(matTillAlpha); DTAB(matTillBeta);
    DTAB(matCohesion); DTAB(matCohesionDamaged); DTAB(matInternalFriction); DTAB(matInternalFrictionDamaged);
    DTAB(matMeltEnergy); DTAB(matDensityFloor); DTAB(matEnergyFloor); ITAB(matdensity_via_kernel_sum);
    DTAB(matexponent_tensor); DTAB(matepsilon_stress); DTAB(matmean_particle_distance);
    DTAB(matporjutzi_p_elastic); DTAB(matporjutzi_p_transition); DTAB(matporjutzi_p_compacted);
    DTAB(matporjutzi_alpha_0); DTAB(matporjutzi_alpha_e); DTAB(matporjutzi_alpha_t);
    DTAB(matporjutzi_n1); DTAB(matporjutzi_n2); DTAB(matcs_porous); DTAB(matcs_solid); ITAB(matcrushcurve_style);
    ITAB(aneos_n_rho); ITAB(aneos_n_e); ITAB(aneos_rho_id); ITAB(aneos_e_id); ITAB(aneos_matrix_id);
    DTAB(aneos_bulk_cs); DTAB(aneos_gamma);
#undef DTAB
#undef ITAB
    {
        double *aneos_rho_0 = (double *)store_alloc(s, n, sizeof(double));
        const char **tab_file = (const char **)store_alloc(s, n, sizeof(char *));
        int run_rho = 0, run_e = 0, run_mat = 0;

        for (i = 0; i < n; i++) {
            mat_f_sml_min[i] = 1.0;
            mat_f_sml_max[i] = 1.0;
            aneos_rho_id[i] = aneos_e_id[i] = aneos_matrix_id[i] = -1;
        }
        /* pass 1: everything except the table payload (sizes of the concatenated tables are needed first) */
        for (i = 0; i < n && !rc; i++) {
            config_setting_t *mat = config_setting_get_elem(materials, i), *sub;
            double friction_angle = 0.0, friction_angle_damaged = 0.0;
            int id;
            config_setting_lookup_int(mat, "ID", &id);
            config_setting_lookup_float(mat, "sml", &matSml[id]);
#if VARIABLE_SML
            config_setting_lookup_float(mat, "factor_sml_min", &mat_f_sml_min[id]);
            config_setting_lookup_float(mat, "factor_sml_max", &mat_f_sml_max[id]);
#endif
#if ARTIFICIAL_VISCOSITY
            matAlpha[id] = 1.0;
            matBeta[id] = 2.0;
            if ((sub = config_setting_get_member(mat, "artificial_viscosity"))) {
                config_setting_lookup_float(sub, "alpha", &matAlpha[id]);
                config_setting_lookup_float(sub, "beta", &matBeta[id]);
            }
#endif
#if ARTIFICIAL_STRESS
            if (!(sub = config_setting_get_member(mat, "artificial_stress"))) {
                rc = fail(err, errlen, "Error reading material config file %s. Subgroup 'artificial_stress' is missing for material with ID %d.", cfg_path, id);
                break;
            }
            config_setting_lookup_float(sub, "exponent_tensor", &matexponent_tensor[id]);
            config_setting_lookup_float(sub, "epsilon_stress", &matepsilon_stress[id]);
            config_setting_lookup_float(sub, "mean_particle_distance", &matmean_particle_distance[id]);
#endif
            if (!(sub = config_setting_get_member(mat, "eos"))) {
                rc = fail(err, errlen, "Error reading material config file %s. Subgroup 'eos' is missing for material with ID %d.", cfg_path, id);
                break;
            }
            if (!config_setting_lookup_int(sub, "type", &matEOS[id])) {
                rc = fail(err, errlen, "Each material needs an eos.type in the material config file %s (material %d).", cfg_path, id);
                break;
            }
            config_setting_lookup_float(sub, "polytropic_K", &matPolytropicK[id]);
            config_setting_lookup_float(sub, "polytropic_gamma", &matPolytropicGamma[id]);
            config_setting_lookup_float(sub, "isothermal_soundspeed", &matIsothermalSoundSpeed[id]);
            config_setting_lookup_float(sub, "bulk_modulus", &matBulkmodulus[id]);
            config_setting_lookup_float(sub, "shear_modulus", &matShearmodulus[id]);
            config_setting_lookup_float(sub, "yield_stress", &matYieldStress[id]);
            config_setting_lookup_float(sub, "rho_0", &matRho0[id]);
            config_setting_lookup_float(sub, "till_rho_0", &matTillRho0[id]);
            config_setting_lookup_float(sub, "till_E_0", &matTillE0[id]);
            config_setting_lookup_float(sub, "till_E_iv", &matTillEiv[id]);
            config_setting_lookup_float(sub, "till_E_cv", &matTillEcv[id]);
            config_setting_lookup_float(sub, "till_a", &matTilla[id]);
            config_setting_lookup_float(sub, "till_b", &matTillb[id]);
            config_setting_lookup_string(sub, "table_path", &tab_file[id]);
            config_setting_lookup_int(sub, "n_rho", &aneos_n_rho[id]);
            config_setting_lookup_int(sub, "n_e", &aneos_n_e[id]);
            config_setting_lookup_int(sub, "density_via_kernel_sum", &matdensity_via_kernel_sum[id]);
            config_setting_lookup_float(sub, "aneos_rho_0", &aneos_rho_0[id]);
            config_setting_lookup_float(sub, "aneos_bulk_cs", &aneos_bulk_cs[id]);
            config_setting_lookup_float(sub, "aneos_gamma", &aneos_gamma[id]);
            if (matEOS[id] == EOS_TYPE_ANEOS || matEOS[id] == EOS_TYPE_JUTZI_ANEOS) {
                if (aneos_n_rho[id] < 2 || aneos_n_e[id] < 2 || !tab_file[id]) {
                    rc = fail(err, errlen, "ANEOS material in %s needs table_path, n_rho, n_e (material %d)", cfg_path, id);
                    break;
                }
                aneos_rho_id[id] = run_rho; run_rho += aneos_n_rho[id];
                aneos_e_id[id] = run_e; run_e += aneos_n_e[id];
                aneos_matrix_id[id] = run_mat; run_mat += aneos_n_rho[id] * aneos_n_e[id];
            }
#if PALPHA_POROSITY
            config_setting_lookup_float(sub, "porjutzi_p_elastic", &matporjutzi_p_elastic[id]);
            config_setting_lookup_float(sub, "porjutzi_p_transition", &matporjutzi_p_transition[id]);
            config_setting_lookup_float(sub, "porjutzi_p_compacted", &matporjutzi_p_compacted[id]);
            if (!config_setting_lookup_float(sub, "porjutzi_alpha_0", &matporjutzi_alpha_0[id])) matporjutzi_alpha_0[id] = 1.0;
            if (!config_setting_lookup_float(sub, "porjutzi_alpha_e", &matporjutzi_alpha_e[id])) matporjutzi_alpha_e[id] = 1.0;
            if (!config_setting_lookup_float(sub, "porjutzi_alpha_t", &matporjutzi_alpha_t[id])) matporjutzi_alpha_t[id] = 1.0;
            config_setting_lookup_float(sub, "porjutzi_n1", &matporjutzi_n1[id]);
            config_setting_lookup_float(sub, "porjutzi_n2", &matporjutzi_n2[id]);
            if (!config_setting_lookup_float(sub, "cs_porous", &matcs_porous[id])) {
                /* till_A has not been read yet at this point in the reference either */
                if (matEOS[id] == EOS_TYPE_JUTZI) matcs_porous[id] = 0.5 * sqrt(matTillA[id] / matTillRho0[id]);
                else if (matEOS[id] == EOS_TYPE_JUTZI_ANEOS) matcs_porous[id] = 0.5 * aneos_bulk_cs[id];
                else if (matEOS[id] == EOS_TYPE_JUTZI_MURNAGHAN) matcs_porous[id] = 0.5 * sqrt(matBulkmodulus[id] / matRho0[id]);
            }
            if (matEOS[id] == EOS_TYPE_JUTZI_MURNAGHAN)
                matcs_solid[id] = sqrt(matBulkmodulus[id] / matRho0[id] / matporjutzi_alpha_0[id]);
            config_setting_lookup_int(sub, "crushcurve_style", &matcrushcurve_style[id]);
#endif
            config_setting_lookup_float(sub, "till_A", &matTillA[id]);
            config_setting_lookup_float(sub, "till_B", &matTillB[id]);
            config_setting_lookup_float(sub, "till_alpha", &matTillAlpha[id]);
            config_setting_lookup_float(sub, "till_beta", &matTillBeta[id]);
            if (!config_setting_lookup_float(sub, "cs_limit", &matcsLimit[id])) {
                if (matEOS[id] == EOS_TYPE_TILLOTSON || matEOS[id] == EOS_TYPE_JUTZI)
                    matcsLimit[id] = 0.01 * sqrt(matTillA[id] / matTillRho0[id]);
                else if (matEOS[id] == EOS_TYPE_ANEOS || matEOS[id] == EOS_TYPE_JUTZI_ANEOS)
                    matcsLimit[id] = 0.01 * aneos_bulk_cs[id];
            }
            if (!config_setting_lookup_float(sub, "rho_limit", &matRhoLimit[id])) {
                if (matEOS[id] == EOS_TYPE_TILLOTSON || matEOS[id] == EOS_TYPE_JUTZI || matEOS[id] == EOS_TYPE_MURNAGHAN ||
                    matEOS[id] == EOS_TYPE_JUTZI_MURNAGHAN)
                    matRhoLimit[id] = 0.9;
            }
            if (!config_setting_lookup_float(sub, "n", &matN[id])) matN[id] = 1.0;
            config_setting_lookup_float(sub, "cohesion", &matCohesion[id]);
            config_setting_lookup_float(sub, "cohesion_damaged", &matCohesionDamaged[id]);
            config_setting_lookup_float(sub, "friction_angle", &friction_angle);
            config_setting_lookup_float(sub, "friction_angle_damaged", &friction_angle_damaged);
            config_setting_lookup_float(sub, "melt_energy", &matMeltEnergy[id]);
            matInternalFriction[id] = tan(friction_angle);
            matInternalFrictionDamaged[id] = tan(friction_angle_damaged);
#if SOLID
            matYoungModulus[id] = 9.0 * matBulkmodulus[id] * matShearmodulus[id] / (3.0 * matBulkmodulus[id] + matShearmodulus[id]);
#endif
            if (!config_setting_lookup_float(mat, "density_floor", &matDensityFloor[id])) {
                switch (matEOS[id]) {
                    case EOS_TYPE_MURNAGHAN: case EOS_TYPE_JUTZI_MURNAGHAN: case EOS_TYPE_VISCOUS_REGOLITH:
                        matDensityFloor[id] = matRho0[id] * 0.01; break;
                    case EOS_TYPE_TILLOTSON: case EOS_TYPE_JUTZI: case EOS_TYPE_EPSILON:
                        matDensityFloor[id] = matTillRho0[id] * 0.01; break;
                    case EOS_TYPE_ANEOS: case EOS_TYPE_JUTZI_ANEOS:
                        matDensityFloor[id] = aneos_rho_0[id] * 0.01; break;
                    default: matDensityFloor[id] = 0.0; break;
                }
            }
            if (!config_setting_lookup_float(mat, "energy_floor", &matEnergyFloor[id])) matEnergyFloor[id] = -1e30;
        }
        /* pass 2: table payload */
        if (!rc && run_mat > 0) {
            double *t_rho = (double *)store_alloc(s, run_rho, sizeof(double));
            double *t_e = (double *)store_alloc(s, run_e, sizeof(double));
            double *t_p = (double *)store_alloc(s, run_mat, sizeof(double));
            double *t_cs = (double *)store_alloc(s, run_mat, sizeof(double));
            m->aneos_rho = t_rho; m->aneos_e = t_e; m->aneos_p = t_p; m->aneos_cs = t_cs;
            m->aneos_rho_len = run_rho; m->aneos_e_len = run_e; m->aneos_matrix_len = run_mat;
            for (i = 0; i < n && !rc; i++) {
                char path[2048];
                if (aneos_matrix_id[i] < 0) continue;
                if (tab_file[i][0] == '/') snprintf(path, sizeof(path), "%s", tab_file[i]);
                else snprintf(path, sizeof(path), "%s%s", dir, tab_file[i]);
                rc = read_aneos_table(path, aneos_n_rho[i], aneos_n_e[i], t_rho + aneos_rho_id[i], t_e + aneos_e_id[i],
                                      t_p + aneos_matrix_id[i], t_cs + aneos_matrix_id[i], err, errlen);
            }
        }
    }
    config_destroy(&cfg);
    if (rc) {
        b200sph_materials_free(m);
        return rc;
    }
    *out = m;
    return 0;
}
