/*
 * capi.cu -- lifetime, material upload and the host-buffer entry point of the C-ABI
 * declared in include/b200sph.h.  The hot path itself is in rhs_kernels.cu.
 */
#include "rhs_internal.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int upload_materials(b200sph_handle *h, const MatParams *host, int n, const AneosTables &tables);
int sort_temp_bytes(int n_max, int bits, size_t *bytes);
int gravity_tree_create(b200sph_handle *h);
void gravity_tree_destroy(b200sph_handle *h);
void halo_state_destroy(b200sph_handle *h);

/* ------------------------------------------------------------------ switch set identity */
struct SwitchEntry {
    const char *name;
    int value;
};
#define X(name) {#name, name},
static const SwitchEntry k_switches[] = {B200SPH_SWITCH_LIST(X)};
#undef X

extern "C" int b200sph_abi_version(void) { return B200SPH_ABI_VERSION; }
extern "C" const char *b200sph_config_name(void) { return B200SPH_CONFIG_NAME; }

extern "C" uint64_t b200sph_switch_hash(void)
{
    uint64_t hsh = 1469598103934665603ull;
    char buf[96];
    for (size_t i = 0; i < sizeof(k_switches) / sizeof(k_switches[0]); i++) {
        snprintf(buf, sizeof(buf), "%s=%d;", k_switches[i].name, k_switches[i].value);
        for (const char *c = buf; *c; c++) {
            hsh ^= (unsigned char)*c;
            hsh *= 1099511628211ull;
        }
    }
    return hsh;
}

extern "C" int b200sph_switch_value(const char *name)
{
    if (!name) return -999;
    for (size_t i = 0; i < sizeof(k_switches) / sizeof(k_switches[0]); i++)
        if (strcmp(k_switches[i].name, name) == 0) return k_switches[i].value;
    return -999;
}

/* ------------------------------------------------------------------ lifetime */
static char g_create_error[512] = "";

extern "C" const char *b200sph_last_error(const b200sph_handle *h) { return h ? h->err : g_create_error; }

template <typename T>
static cudaError_t dev_alloc(T **ptr, size_t count)
{
    return cudaMalloc((void **)ptr, (count ? count : 1) * sizeof(T));
}

extern "C" int b200sph_create(b200sph_handle **out, int n_max, int device, uint64_t expected_switch_hash)
{
    if (!out || n_max <= 0) return B200SPH_ERR_BAD_ARGUMENT;
    *out = nullptr;
    if (expected_switch_hash != 0 && expected_switch_hash != b200sph_switch_hash()) {
        snprintf(g_create_error, sizeof(g_create_error),
                 "switch-set mismatch: caller was built with a different parameter.h than libb200sph_%s", B200SPH_CONFIG_NAME);
        return B200SPH_ERR_SWITCH_MISMATCH;
    }
    b200sph_handle *h = (b200sph_handle *)calloc(1, sizeof(b200sph_handle));
    if (!h) return B200SPH_ERR_BAD_ARGUMENT;
    h->n_max = n_max;
    h->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        snprintf(g_create_error, sizeof(g_create_error), "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
        free(h);
        return B200SPH_ERR_CUDA;
    }
    {
        cudaDeviceProp prop;
        h->n_sm = (cudaGetDeviceProperties(&prop, device) == cudaSuccess && prop.multiProcessorCount > 0) ? prop.multiProcessorCount : 148;
    }
    const size_t n = (size_t)n_max;
    const size_t tiles = (n + NBR_TILE - 1) / NBR_TILE;
    h->max_cells = (int)(8 * n + 4096 < (size_t)0x3fffffff ? 8 * n + 4096 : (size_t)0x3fffffff);
    h->sort_bits = 1;
    while ((1ll << h->sort_bits) < (long long)h->max_cells + 1) h->sort_bits++;
    if (sort_temp_bytes(n_max, h->sort_bits, &h->cub_tmp_bytes) != 0) h->cub_tmp_bytes = 0;

#define ALLOC(ptr, count)                                                                              \
    do {                                                                                               \
        e = dev_alloc(&(ptr), (count));                                                                \
        if (e != cudaSuccess) {                                                                        \
            snprintf(g_create_error, sizeof(g_create_error), "cudaMalloc(%s, %zu elements): %s", #ptr, (size_t)(count), cudaGetErrorString(e)); \
            b200sph_destroy(h);                                                                        \
            return B200SPH_ERR_CUDA;                                                                   \
        }                                                                                              \
    } while (0)
    Sorted &s = h->s;
    ALLOC(s.perm, n); ALLOC(s.keys, n); ALLOC(s.cell_start, (size_t)h->max_cells + 2);
    ALLOC(s.pos4, n); ALLOC(s.vel4, n); ALLOC(s.gas4, n); ALLOC(s.mat, n); ALLOC(s.srch, n);
#if SOLID
    ALLOC(s.ten, n * TEN_RECS);
#endif
    ALLOC(s.nbr, tiles * NBR_TILE * (size_t)MAX_NUM_INTERACTIONS);
    ALLOC(s.noi, n);
    ALLOC(h->keys_in, n); ALLOC(h->idx_in, n);
    ALLOC(h->rho_sorted, n);
    ALLOC(h->block_partials, (size_t)h->n_sm * 4 * 16 + 16);
    ALLOC(h->block_counter, 4);
    ALLOC(h->d_flags, 8);
    ALLOC(h->d_domain, 1);
    {
        char *tmp = nullptr;
        ALLOC(tmp, h->cub_tmp_bytes + 16);
        h->cub_tmp = tmp;
    }
#undef ALLOC
    e = cudaMemset(h->block_counter, 0, 4 * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    h->own_stream = (e == cudaSuccess);
    for (int k = 0; k < 12 && e == cudaSuccess; k++) e = cudaEventCreate(&h->ev[k]);
    if (e != cudaSuccess) {
        snprintf(g_create_error, sizeof(g_create_error), "stream/event setup: %s", cudaGetErrorString(e));
        b200sph_destroy(h);
        return B200SPH_ERR_CUDA;
    }
    if (gravity_tree_create(h) != 0) {
        snprintf(g_create_error, sizeof(g_create_error), "gravity tree allocation failed: %.400s", h->err);
        b200sph_destroy(h);
        return B200SPH_ERR_CUDA;
    }
    h->flag_force_gravity_calc = 0;
    h->gravity_index = 0;
    h->forces_threads = SOLID ? 64 : 128;
    {
        /* tuning knob for profiling sessions; the default above is what ships */
        const char *env = getenv("B200SPH_FORCES_THREADS");
        const int t = env ? atoi(env) : 0;
        if (t == 32 || t == 64 || t == 96 || t == 128) h->forces_threads = t;
        const char *ov = getenv("B200SPH_OVERLAP_POINTWISE");   /* measurement switch; the default is what ships */
        h->overlap_pointwise = ov ? atoi(ov) : 1;
        const char *pad = getenv("B200SPH_PAD_SMEM");
        h->pad_smem = pad ? atoi(pad) : 0;
        if (h->pad_smem < 0 || h->pad_smem > 48 * 1024) h->pad_smem = 0;
    }
    *out = h;
    return B200SPH_OK;
}

extern "C" int b200sph_destroy(b200sph_handle *h)
{
    if (!h) return B200SPH_OK;
    cudaSetDevice(h->device);
    gravity_tree_destroy(h);
    Sorted &s = h->s;
    cudaFree(s.perm); cudaFree(s.keys); cudaFree(s.cell_start); cudaFree(s.pos4); cudaFree(s.vel4); cudaFree(s.gas4);
    cudaFree(s.mat); cudaFree(s.srch); cudaFree(s.ten); cudaFree(s.nbr); cudaFree(s.noi);
    cudaFree(h->rk_scalars); cudaFree(h->rk_partials); cudaFree(h->rk_counter);
    cudaFree(h->keys_in); cudaFree(h->idx_in); cudaFree(h->rho_sorted); cudaFree(h->block_partials);
    cudaFree(h->block_counter); cudaFree(h->d_flags); cudaFree(h->d_domain); cudaFree(h->cub_tmp);
    cudaFree(h->stage);
    if (h->aux_stream) {
        cudaStreamDestroy(h->aux_stream);
        cudaEventDestroy(h->ev_fork);
        cudaEventDestroy(h->ev_join);
    }
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    for (int k = 0; k < 4; k++)
        if (h->ev_copy[k]) cudaEventDestroy(h->ev_copy[k]);
    halo_state_destroy(h);
    cudaFree(h->aneos_buf);
    for (int k = 0; k < 12; k++)
        if (h->ev[k]) cudaEventDestroy(h->ev[k]);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    free(h);
    return B200SPH_OK;
}

/* ------------------------------------------------------------------ materials */
template <typename T>
static int fetch(const T *src, int n, T *dst, b200sph_handle *h)
{
    if (!src) {
        memset(dst, 0, sizeof(T) * n);
        return 0;
    }
    CU(cudaMemcpy(dst, src, sizeof(T) * n, cudaMemcpyDefault));
    return 0;
}

extern "C" int b200sph_set_materials(b200sph_handle *h, const b200sph_materials *mat)
{
    if (!h || !mat || mat->n_materials <= 0 || !mat->matEOS) return B200SPH_ERR_BAD_ARGUMENT;
    if (mat->n_materials > B200_MAX_MATERIALS) {
        snprintf(h->err, sizeof(h->err), "%d materials exceed the built-in limit of %d", mat->n_materials, B200_MAX_MATERIALS);
        return B200SPH_ERR_UNSUPPORTED;
    }
    CU(cudaSetDevice(h->device));
    const int n = mat->n_materials;
    MatParams host[B200_MAX_MATERIALS];
    memset(host, 0, sizeof(host));
    int itmp[B200_MAX_MATERIALS];
    double dtmp[B200_MAX_MATERIALS];
#define GETI(src, field)                                   \
    do {                                                   \
        if (fetch(src, n, itmp, h)) return B200SPH_ERR_CUDA; \
        for (int k = 0; k < n; k++) host[k].field = itmp[k]; \
    } while (0)
#define GETD(src, field)                                   \
    do {                                                   \
        if (fetch(src, n, dtmp, h)) return B200SPH_ERR_CUDA; \
        for (int k = 0; k < n; k++) host[k].field = dtmp[k]; \
    } while (0)
    GETI(mat->matEOS, eos);
    GETI(mat->matdensity_via_kernel_sum, density_via_kernel_sum);
    GETI(mat->matcrushcurve_style, crushcurve_style);
    GETI(mat->aneos_n_rho, aneos_n_rho); GETI(mat->aneos_n_e, aneos_n_e);
    GETI(mat->aneos_rho_id, aneos_rho_id); GETI(mat->aneos_e_id, aneos_e_id); GETI(mat->aneos_matrix_id, aneos_matrix_id);
    GETD(mat->matSml, sml); GETD(mat->mat_f_sml_min, f_sml_min); GETD(mat->mat_f_sml_max, f_sml_max);
    GETD(mat->matAlpha, av_alpha); GETD(mat->matBeta, av_beta);
    GETD(mat->matPolytropicK, poly_K); GETD(mat->matPolytropicGamma, poly_gamma); GETD(mat->matIsothermalSoundSpeed, iso_cs);
    GETD(mat->matBulkmodulus, bulk); GETD(mat->matShearmodulus, shear); GETD(mat->matYoungModulus, young);
    GETD(mat->matYieldStress, yield_stress);
    GETD(mat->matRho0, rho0); GETD(mat->matN, n); GETD(mat->matRhoLimit, rho_limit); GETD(mat->matcsLimit, cs_limit);
    GETD(mat->matTillRho0, till_rho0); GETD(mat->matTillA, till_A); GETD(mat->matTillB, till_B); GETD(mat->matTillE0, till_E0);
    GETD(mat->matTillEiv, till_Eiv); GETD(mat->matTillEcv, till_Ecv); GETD(mat->matTilla, till_a); GETD(mat->matTillb, till_b);
    GETD(mat->matTillAlpha, till_alpha); GETD(mat->matTillBeta, till_beta);
    GETD(mat->matCohesion, cohesion); GETD(mat->matCohesionDamaged, cohesion_damaged);
    GETD(mat->matInternalFriction, friction); GETD(mat->matInternalFrictionDamaged, friction_damaged);
    GETD(mat->matMeltEnergy, melt_energy);
    GETD(mat->matDensityFloor, density_floor); GETD(mat->matEnergyFloor, energy_floor);
    GETD(mat->matexponent_tensor, exponent_tensor); GETD(mat->matepsilon_stress, epsilon_stress);
    GETD(mat->matmean_particle_distance, mean_particle_distance);
    GETD(mat->matporjutzi_p_elastic, pj_p_elastic); GETD(mat->matporjutzi_p_transition, pj_p_transition);
    GETD(mat->matporjutzi_p_compacted, pj_p_compacted); GETD(mat->matporjutzi_alpha_0, pj_alpha_0);
    GETD(mat->matporjutzi_alpha_e, pj_alpha_e); GETD(mat->matporjutzi_alpha_t, pj_alpha_t);
    GETD(mat->matporjutzi_n1, pj_n1); GETD(mat->matporjutzi_n2, pj_n2);
    GETD(mat->matcs_porous, cs_porous); GETD(mat->matcs_solid, cs_solid);
    GETD(mat->aneos_bulk_cs, aneos_bulk_cs); GETD(mat->aneos_gamma, aneos_gamma);
#undef GETI
#undef GETD
    int kernel_sum = !INTEGRATE_DENSITY;
    h->s.any_eos_ignore = 0;
    for (int k = 0; k < n; k++) {
        const int eos = host[k].eos;
        if (eos == EOS_TYPE_IGNORE) h->s.any_eos_ignore = 1;
        if (host[k].density_via_kernel_sum > 0) kernel_sum = 1;
        if (!mat->mat_f_sml_min) host[k].f_sml_min = 1.0;
        if (!mat->mat_f_sml_max) host[k].f_sml_max = 1.0;
        const bool known = eos == EOS_TYPE_IGNORE || eos == EOS_TYPE_POLYTROPIC_GAS || eos == EOS_TYPE_MURNAGHAN ||
                           eos == EOS_TYPE_TILLOTSON || eos == EOS_TYPE_ISOTHERMAL_GAS || eos == EOS_TYPE_ANEOS ||
                           eos == EOS_TYPE_IDEAL_GAS
#if PALPHA_POROSITY
                           || eos == EOS_TYPE_JUTZI || eos == EOS_TYPE_JUTZI_MURNAGHAN || eos == EOS_TYPE_JUTZI_ANEOS
#endif
            ;
        if (!known) {
            snprintf(h->err, sizeof(h->err), "material %d: eos.type = %d is outside the hot-path scope of libb200sph_%s", k, eos,
                     B200SPH_CONFIG_NAME);
            return B200SPH_ERR_UNSUPPORTED;
        }
#if PALPHA_POROSITY
        if ((eos == EOS_TYPE_JUTZI || eos == EOS_TYPE_JUTZI_MURNAGHAN || eos == EOS_TYPE_JUTZI_ANEOS) && (host[k].crushcurve_style < 0 || host[k].crushcurve_style > 4)) {
            snprintf(h->err, sizeof(h->err), "material %d: crushcurve_style = %d is not one of the reference's crush curves 0..4 (src/pressure.cu:365-440)",
                     k, host[k].crushcurve_style);
            return B200SPH_ERR_UNSUPPORTED;
        }
#endif
        if ((eos == EOS_TYPE_ANEOS || eos == EOS_TYPE_JUTZI_ANEOS) && (!mat->aneos_rho || !mat->aneos_e || !mat->aneos_p || !mat->aneos_cs || host[k].aneos_matrix_id < 0)) {
            snprintf(h->err, sizeof(h->err), "material %d: ANEOS tables missing", k);
            return B200SPH_ERR_BAD_ARGUMENT;
        }
    }
    /* tabulated EOS payload lives in global memory owned by the handle */
    static_assert(sizeof(AneosTables) == 4 * sizeof(void *), "AneosTables layout");
    AneosTables tables = {nullptr, nullptr, nullptr, nullptr};
    if (mat->aneos_matrix_len > 0 && mat->aneos_rho && mat->aneos_p) {
        const size_t total = (size_t)mat->aneos_rho_len + mat->aneos_e_len + 2 * (size_t)mat->aneos_matrix_len;
        double *buf = nullptr;
        if (h->aneos_buf) cudaFree(h->aneos_buf);
        h->aneos_buf = nullptr;
        CU(cudaMalloc((void **)&buf, total * sizeof(double)));
        h->aneos_buf = buf;
        double *d_rho = buf, *d_e = d_rho + mat->aneos_rho_len, *d_p = d_e + mat->aneos_e_len, *d_cs = d_p + mat->aneos_matrix_len;
        CU(cudaMemcpy(d_rho, mat->aneos_rho, sizeof(double) * mat->aneos_rho_len, cudaMemcpyDefault));
        CU(cudaMemcpy(d_e, mat->aneos_e, sizeof(double) * mat->aneos_e_len, cudaMemcpyDefault));
        CU(cudaMemcpy(d_p, mat->aneos_p, sizeof(double) * mat->aneos_matrix_len, cudaMemcpyDefault));
        CU(cudaMemcpy(d_cs, mat->aneos_cs, sizeof(double) * mat->aneos_matrix_len, cudaMemcpyDefault));
        tables.rho = d_rho; tables.e = d_e; tables.p = d_p; tables.cs = d_cs;
    }
    if (upload_materials(h, host, n, tables)) return B200SPH_ERR_CUDA;
    h->kernel_sum_density = kernel_sum;
    h->materials_set = 1;
    return B200SPH_OK;
}

/* ------------------------------------------------------------------ misc */
extern "C" int b200sph_get_stats(const b200sph_handle *h, b200sph_stats *out)
{
    if (!h || !out) return B200SPH_ERR_BAD_ARGUMENT;
    *out = h->stats;
    return B200SPH_OK;
}

extern "C" int b200sph_set_stream(b200sph_handle *h, void *cuda_stream)
{
    if (!h) return B200SPH_ERR_BAD_ARGUMENT;
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    h->own_stream = 0;
    h->stream = (cudaStream_t)cuda_stream;
    return B200SPH_OK;
}

extern "C" int b200sph_set_owned(b200sph_handle *h, int n_owned)
{
    if (!h || n_owned < 0) return B200SPH_ERR_BAD_ARGUMENT;
    h->n_owned = n_owned;
    return B200SPH_OK;
}

extern "C" int b200sph_set_gravity_sources(b200sph_handle *h, const double *x, const double *y, const double *z, const double *m,
                                           int n_sources, int own_begin)
{
    if (!h) return B200SPH_ERR_BAD_ARGUMENT;
    if (n_sources <= 0 || !x || !m) {
        h->grav_src_n = 0;
        return B200SPH_OK;
    }
    h->grav_src[0] = x; h->grav_src[1] = y; h->grav_src[2] = z; h->grav_src[3] = m;
    h->grav_src_n = n_sources;
    h->grav_own_begin = own_begin;
    return B200SPH_OK;
}

extern "C" int b200sph_set_global_domain(b200sph_handle *h, const double lo[3], const double hi[3])
{
    if (!h) return B200SPH_ERR_BAD_ARGUMENT;
    if (!lo || !hi) {
        h->have_global_domain = 0;
        return B200SPH_OK;
    }
    for (int a = 0; a < 3; a++) {
        h->global_lo[a] = lo[a];
        h->global_hi[a] = hi[a];
    }
    h->have_global_domain = 1;
    return B200SPH_OK;
}

/* ------------------------------------------------------------------ host-buffer entry point
 * Field table: which members are inputs (copied host->device before the call) and which are
 * read back.  `per` = elements per particle: 1, DD (tensors) or -1 (flaws: max_num_flaws). */
struct FieldDesc {
    size_t offset;   /* offset of the pointer inside b200sph_particle_arrays */
    int is_int;
    int per;
    int in_p, out_p;       /* role inside view.p */
    int in_rhs, out_rhs;   /* role inside view.p_rhs */
};
#define OFF(f) offsetof(b200sph_particle_arrays, f)
static const FieldDesc k_fields[] = {
    {OFF(x), 0, 1, 1, 0, 0, 0}, {OFF(y), 0, 1, 1, 0, 0, 0}, {OFF(z), 0, 1, 1, 0, 0, 0},
    {OFF(vx), 0, 1, 1, 1, 0, 0}, {OFF(vy), 0, 1, 1, 1, 0, 0}, {OFF(vz), 0, 1, 1, 1, 0, 0},
    {OFF(dxdt), 0, 1, 0, 1, 0, 0}, {OFF(dydt), 0, 1, 0, 1, 0, 0}, {OFF(dzdt), 0, 1, 0, 1, 0, 0},
    {OFF(ax), 0, 1, 0, 1, 0, 0}, {OFF(ay), 0, 1, 0, 1, 0, 0}, {OFF(az), 0, 1, 0, 1, 0, 0},
    {OFF(g_ax), 0, 1, 1, 1, 0, 0}, {OFF(g_ay), 0, 1, 1, 1, 0, 0}, {OFF(g_az), 0, 1, 1, 1, 0, 0},
    {OFF(g_local_cellsize), 0, 1, 0, 0, 1, 1}, {OFF(g_x), 0, 1, 0, 0, 1, 1}, {OFF(g_y), 0, 1, 0, 0, 1, 1}, {OFF(g_z), 0, 1, 0, 0, 1, 1},
    {OFF(m), 0, 1, 1, 0, 0, 0}, {OFF(h), 0, 1, 1, 1, 0, 0}, {OFF(h0), 0, 1, 0, 0, 1, 0}, {OFF(dhdt), 0, 1, 0, 1, 0, 0},
    {OFF(rho), 0, 1, 1, 1, 0, 0}, {OFF(drhodt), 0, 1, 0, 1, 0, 0}, {OFF(p), 0, 1, 1, 1, 0, 0}, {OFF(e), 0, 1, 1, 1, 0, 0},
    {OFF(dedt), 0, 1, 0, 1, 0, 0},
    {OFF(S), 0, DD, 1, 1, 0, 0}, {OFF(dSdt), 0, DD, 0, 1, 0, 0}, {OFF(local_strain), 0, 1, 0, 1, 0, 0},
    {OFF(ep), 0, 1, 0, 0, 0, 0}, {OFF(edotp), 0, 1, 0, 1, 0, 0}, {OFF(plastic_f), 0, 1, 0, 0, 0, 1}, {OFF(sigma), 0, DD, 0, 0, 0, 1},
    {OFF(R), 0, DD, 0, 0, 0, 1},
    {OFF(d), 0, 1, 1, 1, 0, 0}, {OFF(damage_total), 0, 1, 1, 1, 0, 0}, {OFF(dddt), 0, 1, 0, 1, 0, 0},
    {OFF(numFlaws), 1, 1, 1, 0, 0, 0}, {OFF(numActiveFlaws), 1, 1, 1, 1, 0, 0}, {OFF(flaws), 0, -1, 0, 0, 1, 0},
    {OFF(damage_porjutzi), 0, 1, 1, 1, 0, 0}, {OFF(ddamage_porjutzidt), 0, 1, 0, 1, 0, 0},
    {OFF(muijmax), 0, 1, 0, 1, 0, 0},
    {OFF(pold), 0, 1, 0, 0, 0, 0}, {OFF(alpha_jutzi), 0, 1, 1, 1, 0, 0}, {OFF(alpha_jutzi_old), 0, 1, 0, 1, 0, 0},
    {OFF(dalphadt), 0, 1, 0, 1, 0, 0}, {OFF(dalphadp), 0, 1, 0, 1, 0, 0}, {OFF(dalphadrho), 0, 1, 0, 1, 0, 0},
    {OFF(f), 0, 1, 0, 1, 0, 0}, {OFF(delpdelrho), 0, 1, 0, 1, 0, 0}, {OFF(delpdele), 0, 1, 0, 1, 0, 0},
    {OFF(tensorialCorrectionMatrix), 0, DD, 0, 0, 0, 1},
    {OFF(cs), 0, 1, 1, 1, 0, 0},
    {OFF(noi), 1, 1, 0, 1, 0, 0}, {OFF(materialId), 1, 1, 0, 0, 1, 0}, {OFF(depth), 1, 1, 0, 1, 0, 0},
};
#undef OFF

static inline void **field_ptr(b200sph_particle_arrays *a, size_t off) { return (void **)((char *)a + off); }
static inline void *const *field_ptr(const b200sph_particle_arrays *a, size_t off) { return (void *const *)((const char *)a + off); }

/* roles of a member in the overlapped host-buffer call */
static bool off_in(size_t off, const size_t *list, int n)
{
    for (int k = 0; k < n; k++)
        if (list[k] == off) return true;
    return false;
}
#define OFF(f) offsetof(b200sph_particle_arrays, f)
/* never change during a run: the reference allocates/copies them once (allocate_immutables,
 * src/memory_handling.cu:111-122; copy_particles_immutables_device_to_device, :373-392) */
static const size_t k_immutable[] = {OFF(m), OFF(h0), OFF(materialId), OFF(numFlaws), OFF(flaws)};
/* read by k_prepare / k_cell_keys / k_gather / k_neighbours / k_density: must land before the first kernel */
static const size_t k_first_inputs[] = {OFF(x), OFF(y), OFF(z), OFF(vx), OFF(vy), OFF(vz), OFF(m), OFF(h), OFF(h0),
                                        OFF(rho), OFF(e), OFF(materialId)};
/* final once k_pointwise has run (k_correction / k_forces / gravity do not write them) */
static const size_t k_early_outputs[] = {OFF(p), OFF(cs), OFF(S), OFF(alpha_jutzi_old), OFF(dalphadp), OFF(dalphadrho),
                                         OFF(delpdele), OFF(delpdelrho), OFF(f), OFF(damage_total), OFF(damage_porjutzi),
                                         OFF(h), OFF(e), OFF(sigma), OFF(R), OFF(plastic_f)};
/* p_rhs scratch no caller reads (SURVEY 8b "outputs that callers read" does not list them) */
static const size_t k_scratch_outputs[] = {OFF(sigma), OFF(R), OFF(plastic_f), OFF(tensorialCorrectionMatrix)};
static const size_t k_velocity[] = {OFF(vx), OFF(vy), OFF(vz)};
#undef OFF
#define IN_LIST(off, list) off_in(off, list, (int)(sizeof(list) / sizeof(list[0])))

struct HostCall {
    const b200sph_view *hv;
    b200sph_view dv;
    size_t n, nf;
    int64_t out_bytes;
    cudaError_t err;
};

static size_t host_field_bytes(const FieldDesc &f, size_t n, size_t nf)
{
    const size_t per = f.per == -1 ? nf : (size_t)f.per;
    return n * per * (f.is_int ? sizeof(int) : sizeof(double));
}

/* device->host copies of one output class on the copy stream: which = 0 early, 1 late */
static void host_enqueue_outputs(b200sph_handle *h, HostCall *c, int which)
{
    const int nfields = (int)(sizeof(k_fields) / sizeof(k_fields[0]));
    const bool skip_scratch = (h->host_options & B200SPH_HOST_SKIP_SCRATCH) != 0;
    for (int k = 0; k < nfields && c->err == cudaSuccess; k++) {
        const FieldDesc &f = k_fields[k];
        const bool early = IN_LIST(f.offset, k_early_outputs);
        if ((which == 0) != early) continue;
        if (skip_scratch && IN_LIST(f.offset, k_scratch_outputs)) continue;
        /* velocities come back unchanged unless k_prepare froze a particle (BoundaryConditionsBeforeRHS) */
        if (IN_LIST(f.offset, k_velocity) && h->h_domain.n_frozen == 0) continue;
        const size_t raw = host_field_bytes(f, c->n, c->nf);
        void *hp = *field_ptr(&c->hv->p, f.offset);
        void *hr = *field_ptr(&c->hv->p_rhs, f.offset);
        if (hp && (f.out_p || (hr == hp && f.out_rhs))) {
            c->err = cudaMemcpyAsync(hp, *field_ptr(&c->dv.p, f.offset), raw, cudaMemcpyDeviceToHost, h->copy_stream);
            c->out_bytes += (int64_t)raw;
        }
        if (c->err == cudaSuccess && hr && hr != hp && f.out_rhs) {
            c->err = cudaMemcpyAsync(hr, *field_ptr(&c->dv.p_rhs, f.offset), raw, cudaMemcpyDeviceToHost, h->copy_stream);
            c->out_bytes += (int64_t)raw;
        }
    }
}

static void host_after_pointwise(b200sph_handle *h, void *ctx)
{
    HostCall *c = (HostCall *)ctx;
    c->err = cudaEventRecord(h->ev_copy[2], h->stream);
    if (c->err == cudaSuccess) c->err = cudaStreamWaitEvent(h->copy_stream, h->ev_copy[2], 0);
    if (c->err == cudaSuccess) host_enqueue_outputs(h, c, 0);
}

extern "C" int b200sph_host_options(b200sph_handle *h, int options)
{
    if (!h) return B200SPH_ERR_BAD_ARGUMENT;
    h->host_options = options;
    h->host_imm_valid = 0;
    return B200SPH_OK;
}

/* Three overlapped stages instead of copy-in / compute / copy-out:
 *   compute stream : [first inputs H2D] prepare, sort, search, density | wait | pointwise | correction, forces, gravity
 *   copy stream    :                    [late inputs H2D ............]          | [early outputs D2H ......] [late outputs D2H]
 * "first inputs" are what the search reads (positions, h, velocities, rho, e); "early outputs" are the state
 * k_pointwise finalises (p, c_s, S, porosity partials).  PCIe is full duplex but one call only has one
 * direction busy at a time, so the gain is the kernels hidden behind the copies. */
extern "C" int b200sph_rhs_eval_host(b200sph_handle *h, const b200sph_view *hv, int *offender, int64_t *h2d_bytes, int64_t *d2h_bytes)
{
    if (!h || !hv || hv->n <= 0 || hv->n > h->n_max) return B200SPH_ERR_BAD_ARGUMENT;
    CU(cudaSetDevice(h->device));
    const size_t n = (size_t)hv->n;
    const size_t nf = (size_t)(hv->max_num_flaws > 0 ? hv->max_num_flaws : 1);
    const int nfields = (int)(sizeof(k_fields) / sizeof(k_fields[0]));
    if (!h->copy_stream) {
        CU(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        for (int k = 0; k < 4; k++) CU(cudaEventCreateWithFlags(&h->ev_copy[k], cudaEventDisableTiming));
    }
    auto bytes_of = [&](const FieldDesc &f) -> size_t { return (host_field_bytes(f, n, nf) + 255) & ~(size_t)255; };
    /* device mirror: one slab, carved per present field (p first, then p_rhs when it has its own buffer) */
    size_t need = 0;
    for (int k = 0; k < nfields; k++) {
        if (*field_ptr(&hv->p, k_fields[k].offset)) need += bytes_of(k_fields[k]);
        void *rp = *field_ptr(&hv->p_rhs, k_fields[k].offset);
        if (rp && rp != *field_ptr(&hv->p, k_fields[k].offset)) need += bytes_of(k_fields[k]);
    }
    if (need > h->stage_bytes) {
        if (h->stage) cudaFree(h->stage);
        h->stage = nullptr;
        h->stage_bytes = 0;
        h->host_imm_valid = 0;
        CU(cudaMalloc(&h->stage, need));
        h->stage_bytes = need;
    }
    /* cached immutables stay valid only for the same host arrays and particle count */
    const void *imm_key[6] = {hv->p.m, hv->p_rhs.h0, hv->p_rhs.materialId, hv->p.numFlaws ? (const void *)hv->p.numFlaws : (const void *)hv->p_rhs.numFlaws,
                              hv->p_rhs.flaws, hv->p.x};
    bool imm_cached = (h->host_options & B200SPH_HOST_CACHE_IMMUTABLES) && h->host_imm_valid && h->host_imm_n == hv->n;
    for (int k = 0; k < 6 && imm_cached; k++) imm_cached = (imm_key[k] == h->host_imm_key[k]);

    HostCall call;
    call.hv = hv;
    call.dv = *hv;
    call.n = n;
    call.nf = nf;
    call.out_bytes = 0;
    call.err = cudaSuccess;
    b200sph_view &dv = call.dv;
    memset(&dv.p, 0, sizeof(dv.p));
    memset(&dv.p_rhs, 0, sizeof(dv.p_rhs));
    int64_t in_bytes = 0;
    cudaStream_t st = h->stream, cs = h->copy_stream;
    /* pass 0 queues the first-stage inputs on the compute stream, pass 1 the late ones on the copy stream
     * (behind the first stage, so the two do not share the host->device engine) */
    for (int pass = 0; pass < 2; pass++) {
        char *cursor = (char *)h->stage;
        if (pass == 1) {
            CU(cudaEventRecord(h->ev_copy[0], st));
            CU(cudaStreamWaitEvent(cs, h->ev_copy[0], 0));
        }
        for (int k = 0; k < nfields; k++) {
            const FieldDesc &f = k_fields[k];
            const size_t raw = host_field_bytes(f, n, nf);
            const bool first = IN_LIST(f.offset, k_first_inputs);
            const bool skip = imm_cached && IN_LIST(f.offset, k_immutable);
            cudaStream_t q = first ? st : cs;
            const bool mine = (first == (pass == 0));
            void *hp = *field_ptr(&hv->p, f.offset);
            void *hr = *field_ptr(&hv->p_rhs, f.offset);
            void *dp = nullptr;
            if (hp) {
                dp = cursor;
                cursor += bytes_of(f);
                *field_ptr(&dv.p, f.offset) = dp;
                if (mine && !skip && (f.in_p || (hr == hp && f.in_rhs))) {
                    CU(cudaMemcpyAsync(dp, hp, raw, cudaMemcpyHostToDevice, q));
                    in_bytes += (int64_t)raw;
                }
            }
            if (hr) {
                if (hr == hp) {
                    *field_ptr(&dv.p_rhs, f.offset) = dp;
                } else {
                    void *dr = cursor;
                    cursor += bytes_of(f);
                    *field_ptr(&dv.p_rhs, f.offset) = dr;
                    if (mine && !skip && f.in_rhs) {
                        CU(cudaMemcpyAsync(dr, hr, raw, cudaMemcpyHostToDevice, q));
                        in_bytes += (int64_t)raw;
                    }
                }
            }
        }
    }
    CU(cudaEventRecord(h->ev_copy[1], cs));
    h->hook_wait_before_pointwise = h->ev_copy[1];
    h->hook_after_pointwise = host_after_pointwise;
    h->hook_ctx = &call;
    const int rc = b200sph_rhs_eval(h, &dv, offender);   /* returns with the compute stream drained */
    h->hook_wait_before_pointwise = nullptr;
    h->hook_after_pointwise = nullptr;
    h->hook_ctx = nullptr;
    if (rc == B200SPH_OK && call.err == cudaSuccess) host_enqueue_outputs(h, &call, 1);
    const cudaError_t e_sync = cudaStreamSynchronize(cs);
    if (rc != B200SPH_OK) {
        h->host_imm_valid = 0;
        return rc;
    }
    CU(call.err);
    CU(e_sync);
    for (int k = 0; k < 6; k++) h->host_imm_key[k] = imm_key[k];
    h->host_imm_n = hv->n;
    h->host_imm_valid = 1;
    if (h2d_bytes) *h2d_bytes = in_bytes;
    if (d2h_bytes) *d2h_bytes = call.out_bytes;
    return B200SPH_OK;
}
