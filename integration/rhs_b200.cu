/*
 * rhs_b200.cu -- the reference-side binding: miluphcuda's `void rightHandSide(void)`
 * implemented on top of libb200sph_<config>.so.
 *
 * A maintainer drops this file into the reference's src/ in place of src/rhs.cu
 * (reference: include/rhs.h:30 declares the symbol; the eleven call sites in
 * src/rk2adaptive.cu:223,293,314, src/predictor_corrector.cu:783,816,
 * src/predictor_corrector_euler.cu:734,822, src/euler.cu:164 and
 * src/coupled_heun_rk4_sph_nbody.cu:735,824 stay untouched), adds
 * `-I<b200sph>/include -I<b200sph>/miluphcuda_b200/csrc` to NVFLAGS and
 * `-lb200sph_<config>` to LDFLAGS.  Everything else of the reference -- parameter.h,
 * material.cfg via libconfig, the input/HDF5 formats, the command line, the
 * integrators -- is unchanged.  oracle/build_ref.sh builds exactly this variant as
 * oracle/_ref/miluphcuda_<config>_b200 for the drop-in tests.
 *
 * What the reference passes through globals is gathered here:
 *   - the bound buffer: the integrator copies a `struct Particle` into the __constant__
 *     symbol `p` right before each call (e.g. src/rk2adaptive.cu:219); it is read back;
 *   - `p_rhs` is always `p_device` (src/timeintegration.cu:200-201);
 *   - material tables: the device pointers stored in the __constant__ symbols mat*
 *     (include/config_parameter.h:172-290) are read once;
 *   - scalars: numberOfParticles, numberOfRealParticles, maxNumFlaws_host, treeTheta,
 *     param.selfgravity / decouplegravity, gravConst, isRelaxationRun.
 * Errors follow the reference's convention (include/cuda_utils.h:29-48): print and exit(1).
 */
#include "miluph.h"
#include "timeintegration.h"
#include "config_parameter.h"
#include "rhs.h"
#include "parameter.h"
#include "pressure.h"
#include "little_helpers.h"

#define B200SPH_NO_EOS_ENUM
#include "b200sph.h"
#include "switches.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

extern volatile int terminate_flag;
extern __device__ double gravConst;
extern __constant__ int isRelaxationRun;

static b200sph_handle *g_b200 = NULL;
static double g_grav_const = 0.0;

static uint64_t b200_expected_hash(void)
{
    /* same FNV-1a over "NAME=value;" as b200sph_switch_hash(), evaluated with THIS build's parameter.h */
    struct { const char *name; int value; } sw[] = {
#define X(name) {#name, name},
        B200SPH_SWITCH_LIST(X)
#undef X
    };
    uint64_t h = 1469598103934665603ull;
    char buf[96];
    for (size_t i = 0; i < sizeof(sw) / sizeof(sw[0]); i++) {
        snprintf(buf, sizeof(buf), "%s=%d;", sw[i].name, sw[i].value);
        for (const char *c = buf; *c; c++) {
            h ^= (unsigned char)*c;
            h *= 1099511628211ull;
        }
    }
    return h;
}

static void b200_bind(b200sph_particle_arrays *a, const struct Particle *q)
{
    memset(a, 0, sizeof(*a));
    a->x = q->x; a->vx = q->vx; a->dxdt = q->dxdt; a->ax = q->ax; a->g_ax = q->g_ax; a->g_x = q->g_x;
#if DIM > 1
    a->y = q->y; a->vy = q->vy; a->dydt = q->dydt; a->ay = q->ay; a->g_ay = q->g_ay; a->g_y = q->g_y;
#endif
#if DIM > 2
    a->z = q->z; a->vz = q->vz; a->dzdt = q->dzdt; a->az = q->az; a->g_az = q->g_az; a->g_z = q->g_z;
#endif
    a->g_local_cellsize = q->g_local_cellsize;
    a->m = q->m; a->h = q->h; a->h0 = q->h0; a->rho = q->rho; a->drhodt = q->drhodt; a->p = q->p; a->e = q->e;
    a->cs = q->cs; a->noi = q->noi; a->materialId = q->materialId; a->depth = q->depth;
#if INTEGRATE_SML
    a->dhdt = q->dhdt;
#endif
#if INTEGRATE_ENERGY
    a->dedt = q->dedt;
#endif
#if SOLID
    a->S = q->S; a->dSdt = q->dSdt; a->local_strain = q->local_strain; a->ep = q->ep; a->edotp = q->edotp;
    a->plastic_f = q->plastic_f; a->sigma = q->sigma;
#endif
#if ARTIFICIAL_STRESS
    a->R = q->R;
#endif
#if FRAGMENTATION
    a->d = q->d; a->damage_total = q->damage_total; a->dddt = q->dddt; a->numFlaws = q->numFlaws;
    a->numActiveFlaws = q->numActiveFlaws; a->flaws = q->flaws;
# if PALPHA_POROSITY
    a->damage_porjutzi = q->damage_porjutzi; a->ddamage_porjutzidt = q->ddamage_porjutzidt;
# endif
#endif
#if ARTIFICIAL_VISCOSITY
    a->muijmax = q->muijmax;
#endif
#if PALPHA_POROSITY
    a->pold = q->pold; a->alpha_jutzi = q->alpha_jutzi; a->alpha_jutzi_old = q->alpha_jutzi_old; a->dalphadt = q->dalphadt;
    a->dalphadp = q->dalphadp; a->dalphadrho = q->dalphadrho; a->f = q->f; a->delpdelrho = q->delpdelrho; a->delpdele = q->delpdele;
#endif
#if TENSORIAL_CORRECTION
    a->tensorialCorrectionMatrix = q->tensorialCorrectionMatrix;
#endif
}

#define B200_DIE(what, rc)                                                                          \
    do {                                                                                            \
        fprintf(stderr, "b200sph: %s failed (%d): %s\n", what, rc, b200sph_last_error(g_b200));     \
        exit(1);                                                                                    \
    } while (0)

#define B200_TABLE(field, symbol) cudaVerify(cudaMemcpyFromSymbol(&m.field, symbol, sizeof(void *)))

static void b200_init(void)
{
    int device = 0, rc;
    b200sph_materials m;
    cudaVerify(cudaGetDevice(&device));
    rc = b200sph_create(&g_b200, maxNumberOfParticles, device, b200_expected_hash());
    if (rc) B200_DIE("b200sph_create", rc);
    rc = b200sph_set_stream(g_b200, NULL);   /* the reference runs everything on the legacy default stream */
    if (rc) B200_DIE("b200sph_set_stream", rc);

    memset(&m, 0, sizeof(m));
    m.n_materials = numberOfMaterials;
    B200_TABLE(matEOS, matEOS); B200_TABLE(matSml, matSml);
    B200_TABLE(mat_f_sml_min, mat_f_sml_min); B200_TABLE(mat_f_sml_max, mat_f_sml_max);
    B200_TABLE(matAlpha, matAlpha); B200_TABLE(matBeta, matBeta);
    B200_TABLE(matPolytropicK, matPolytropicK); B200_TABLE(matPolytropicGamma, matPolytropicGamma);
    B200_TABLE(matIsothermalSoundSpeed, matIsothermalSoundSpeed);
    B200_TABLE(matBulkmodulus, matBulkmodulus); B200_TABLE(matShearmodulus, matShearmodulus);
    B200_TABLE(matYieldStress, matYieldStress);
    B200_TABLE(matRho0, matRho0); B200_TABLE(matN, matN); B200_TABLE(matRhoLimit, matRhoLimit); B200_TABLE(matcsLimit, matcsLimit);
    B200_TABLE(matTillRho0, matTillRho0); B200_TABLE(matTillA, matTillA); B200_TABLE(matTillB, matTillB);
    B200_TABLE(matTillE0, matTillE0); B200_TABLE(matTillEiv, matTillEiv); B200_TABLE(matTillEcv, matTillEcv);
    B200_TABLE(matTilla, matTilla); B200_TABLE(matTillb, matTillb); B200_TABLE(matTillAlpha, matTillAlpha);
    B200_TABLE(matTillBeta, matTillBeta);
    B200_TABLE(matCohesion, matCohesion); B200_TABLE(matCohesionDamaged, matCohesionDamaged);
    B200_TABLE(matInternalFriction, matInternalFriction); B200_TABLE(matInternalFrictionDamaged, matInternalFrictionDamaged);
    B200_TABLE(matMeltEnergy, matMeltEnergy);
    B200_TABLE(matDensityFloor, matDensityFloor); B200_TABLE(matEnergyFloor, matEnergyFloor);
    B200_TABLE(matdensity_via_kernel_sum, matdensity_via_kernel_sum);
#if SOLID
    B200_TABLE(matYoungModulus, matYoungModulus);
#endif
#if ARTIFICIAL_STRESS
    B200_TABLE(matexponent_tensor, matexponent_tensor); B200_TABLE(matepsilon_stress, matepsilon_stress);
    B200_TABLE(matmean_particle_distance, matmean_particle_distance);
#endif
#if PALPHA_POROSITY
    B200_TABLE(matporjutzi_p_elastic, matporjutzi_p_elastic); B200_TABLE(matporjutzi_p_transition, matporjutzi_p_transition);
    B200_TABLE(matporjutzi_p_compacted, matporjutzi_p_compacted); B200_TABLE(matporjutzi_alpha_0, matporjutzi_alpha_0);
    B200_TABLE(matporjutzi_alpha_e, matporjutzi_alpha_e); B200_TABLE(matporjutzi_alpha_t, matporjutzi_alpha_t);
    B200_TABLE(matporjutzi_n1, matporjutzi_n1); B200_TABLE(matporjutzi_n2, matporjutzi_n2);
    B200_TABLE(matcs_porous, matcs_porous); B200_TABLE(matcs_solid, matcs_solid);
    B200_TABLE(matcrushcurve_style, matcrushcurve_style);
#endif
    /* tabulated EOS: the concatenated tables already live on the device (src/config_parameter.cu) */
    m.aneos_n_rho = aneos_n_rho_d; m.aneos_n_e = aneos_n_e_d; m.aneos_rho_id = aneos_rho_id_d; m.aneos_e_id = aneos_e_id_d;
    m.aneos_matrix_id = aneos_matrix_id_d; m.aneos_bulk_cs = aneos_bulk_cs_d; m.aneos_gamma = aneos_gamma_d;
    m.aneos_rho = aneos_rho_d; m.aneos_e = aneos_e_d; m.aneos_p = aneos_p_d; m.aneos_cs = aneos_cs_d;
    if (aneos_p_d && aneos_n_rho_d) {
        int *nr = (int *)malloc(sizeof(int) * numberOfMaterials), *ne = (int *)malloc(sizeof(int) * numberOfMaterials);
        int *id = (int *)malloc(sizeof(int) * numberOfMaterials), k;
        cudaVerify(cudaMemcpy(nr, aneos_n_rho_d, sizeof(int) * numberOfMaterials, cudaMemcpyDeviceToHost));
        cudaVerify(cudaMemcpy(ne, aneos_n_e_d, sizeof(int) * numberOfMaterials, cudaMemcpyDeviceToHost));
        cudaVerify(cudaMemcpy(id, aneos_matrix_id_d, sizeof(int) * numberOfMaterials, cudaMemcpyDeviceToHost));
        for (k = 0; k < numberOfMaterials; k++)
            if (id[k] >= 0) {
                m.aneos_rho_len += nr[k];
                m.aneos_e_len += ne[k];
                m.aneos_matrix_len += (int64_t)nr[k] * ne[k];
            }
        free(nr); free(ne); free(id);
    }
    rc = b200sph_set_materials(g_b200, &m);
    if (rc) B200_DIE("b200sph_set_materials", rc);
    cudaVerify(cudaMemcpyFromSymbol(&g_grav_const, gravConst, sizeof(double)));
}

void rightHandSide()
{
    struct Particle bound;
    b200sph_view view;
    int offender = -1, rc, relax = 0;

#if USE_SIGNAL_HANDLER
    /* SIGINT / SIGTERM: write what there is and leave (src/rhs.cu:170-174, src/little_helpers.cu:33-41) */
    if (terminate_flag) {
        copyToHostAndWriteToFile(-2, -2);
    }
#endif
    if (!g_b200) b200_init();
    cudaVerify(cudaMemcpyFromSymbol(&bound, p, sizeof(struct Particle)));
    cudaVerify(cudaMemcpyFromSymbol(&relax, isRelaxationRun, sizeof(int)));

    memset(&view, 0, sizeof(view));
    view.n = numberOfParticles;
    view.n_real = numberOfRealParticles;
#if FRAGMENTATION
    view.max_num_flaws = maxNumFlaws_host;
#else
    view.max_num_flaws = 1;
#endif
    view.selfgravity = param.selfgravity;
    view.decouplegravity = param.decouplegravity;
    view.is_relaxation_run = relax;
    view.theta = treeTheta;
    view.grav_const = g_grav_const;
    b200_bind(&view.p, &bound);
    b200_bind(&view.p_rhs, &p_device);

    rc = b200sph_rhs_eval(g_b200, &view, &offender);
    if (rc == B200SPH_ERR_TOO_MANY_INTERACTIONS) {
        fprintf(stderr, "ERROR: Maximum number of interactions exceeded for particle %d (MAX_NUM_INTERACTIONS = %d)\n",
                offender, MAX_NUM_INTERACTIONS);
        exit(1);
    }
    if (rc) B200_DIE("b200sph_rhs_eval", rc);
}
