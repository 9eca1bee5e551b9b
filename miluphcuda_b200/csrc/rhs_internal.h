/*
 * rhs_internal.h -- state owned by a b200sph handle and the device-side views
 * the kernels work on.
 *
 * Layout in HBM (all arrays sized for n_max particles, allocated once in
 * b200sph_create; nothing is allocated per call):
 *
 *   caller order (the integrator's buffers, never reordered)  --perm-->  cell-sorted scratch
 *
 *   keys[n], perm[n]            cell index of each particle / sorted -> caller index
 *   cell_start[n_cells + 1]     first sorted slot of every cell (x fastest), so the
 *                               particles of a row of x-adjacent cells are one range
 *   pos4[n]   = {x, y, z, h}    32-byte records, one LDG.256 per candidate
 *   vel4[n]   = {vx, vy, vz, m}
 *   gas4[n]   = {p/rho^2 (hydro) or 1/rho^2 (solid), c_s, rho, m/rho}
 *   srch[n]   = {u_x, u_y, u_z, thr} FP32, 16 bytes: conservative pre-filter of the neighbour search
 *   ten[n*TEN_RECS]             sigma/rho^2, correction matrix, R/rho^2 packed (solid)
 *   nbr[(tile*MAX_NUM_INTERACTIONS + k)*32 + lane], noi[n]
 *                               neighbour lists, interleaved per 32-particle tile so the
 *                               k-th entries of a warp's particles share one 128-byte line
 */
#ifndef B200SPH_RHS_INTERNAL_H
#define B200SPH_RHS_INTERNAL_H

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200sph.h"
#include "sph_math.cuh"

#define NBR_TILE 32

/* ------------------------------------------------------------------ 32-byte records
 * Every per-candidate quantity a pair loop gathers lives in 32-byte records that are fetched
 * with ONE 256-bit load (sm_100a LDG.E.256): the pair loops are bound by L1TEX wavefronts
 * (one per distinct 128-byte line per instruction, profiles/r01_ncu_full_sedov_v0_summary.csv),
 * so halving the number of load instructions per pair halves their cost. */
struct __align__(32) Rec4 {
    double x, y, z, w;
};

#ifdef __CUDACC__
__device__ __forceinline__ Rec4 ld_rec(const Rec4 *ptr)
{
    /* read-only for the lifetime of the reading kernel (written by an earlier launch) */
    Rec4 r;
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(ptr));
    return r;
}
__device__ __forceinline__ void st_rec(Rec4 *ptr, const Rec4 &v)
{
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(ptr), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
}
#endif

/* Solid tensors of a particle, packed into TEN_RECS records: sigma/rho^2, then the tensorial
 * correction matrix, then R/rho^2 (artificial stress).  sigma is stored as its upper triangle
 * when the switch set symmetrises S (symmetrizeStress runs under FRAGMENTATION or a plasticity
 * model, reference src/rhs.cu:462-470) -- it is then bitwise symmetric -- and in full otherwise.
 * C = pinv(symmetric) and R = V^T diag V are symmetric up to one rounding; the upper triangle is
 * used for both halves (difference ~1e-16 relative, far inside the 1e-9 gate). */
#if SOLID
#define TEN_SIG_SYM (FRAGMENTATION || B200_PLASTICITY)
#define TEN_NSYM (DIM * (DIM + 1) / 2)
#define TEN_NSIG (TEN_SIG_SYM ? TEN_NSYM : DIM * DIM)
#define TEN_SIG_OFF 0
#define TEN_C_OFF (TEN_SIG_OFF + TEN_NSIG)
#define TEN_R_OFF (TEN_C_OFF + (TENSORIAL_CORRECTION ? TEN_NSYM : 0))
#define TEN_DOUBLES (TEN_R_OFF + (ARTIFICIAL_STRESS ? TEN_NSYM : 0))
#define TEN_RECS ((TEN_DOUBLES + 3) / 4)
#ifdef __CUDACC__
__host__ __device__ constexpr int ten_sym(int a, int b)
{
    return (a <= b) ? (a * DIM - a * (a - 1) / 2 + (b - a)) : (b * DIM - b * (b - 1) / 2 + (a - b));
}
__host__ __device__ constexpr int ten_sig(int a, int b) { return TEN_SIG_OFF + (TEN_SIG_SYM ? ten_sym(a, b) : a * DIM + b); }
__host__ __device__ constexpr int ten_c(int a, int b) { return TEN_C_OFF + ten_sym(a, b); }
__host__ __device__ constexpr int ten_r(int a, int b) { return TEN_R_OFF + ten_sym(a, b); }
#endif
#else
#define TEN_RECS 0
#endif

/* search grid and root cube, computed on the device each call */
struct Domain {
    double lo[3], hi[3];        /* bounding box of all particles */
    double root_centre[3];      /* reference root node: 0.5*(max+min), src/tree.cu:1080-1084 */
    double root_radius;         /* half of the largest extent, src/tree.cu:1071-1078 */
    double cell;                /* edge of a search cell */
    double cell_inv;
    double h_max, h_mean;
    int nc[3];                  /* cells per axis */
    int n_cells;
    int nonfinite;              /* bounding box or h statistics are NaN/Inf: the evaluation is void */
    int n_frozen;               /* particles whose velocity k_prepare zeroed (deactivated / boundary material) */
};

struct Sorted {
    int n;
    int n_owned;                /* caller indices >= n_owned are halo copies: no rates are produced for them */
    int any_eos_ignore;         /* a material with eos.type IGNORE exists: pair loops must look at mat[j] */
    int halo_sums_external;     /* multi-GPU: density / correction matrix of halo copies are delivered by their owners */
    const int *abort;           /* device flag (may be NULL): non-zero = this evaluation must not touch the caller's state */
    int *perm;                  /* sorted slot -> caller index */
    int *keys;
    int *cell_start;
    Rec4 *pos4, *vel4, *gas4;
    float4 *srch;               /* FP32 search record {u_x, u_y, u_z, thr}: position in cell units, acceptance threshold */
    int *mat;
    Rec4 *ten;                  /* TEN_RECS records per particle (solid) */
    int *nbr, *noi;
};

/* multi-GPU: every rank's domain as a union of axis-aligned octree boxes (halo.cu) */
#define HALO_MAX_BOXES 1024
#define HALO_MAX_RANKS 64
struct HaloDomains {
    double lo[HALO_MAX_BOXES][3], hi[HALO_MAX_BOXES][3];
    double list_reach_scale, list_skin;   /* margins of the reusable send plan (halo_copy_needs_list) */
    int rank[HALO_MAX_BOXES];       /* owner of box b */
    int local[HALO_MAX_BOXES];      /* index of box b among its owner's boxes */
    int n_boxes, n_ranks, my_rank, my_first, my_count;
};

struct HaloState {
    HaloDomains host;
    HaloDomains *dev;
    unsigned long long *mask;       /* per particle: bit r = rank r needs it */
    int *blk_counts;                /* [n_ranks][n_blocks] -> exclusive offsets after h_scan */
    int mask_capacity, blk_capacity;
};

struct b200sph_handle {
    int n_max, device;
    int n_sm;                   /* multiProcessorCount: grids of the reduction kernels are n_sm * 4 blocks */
    cudaStream_t stream;
    int own_stream;
    cudaEvent_t ev[12];
    Sorted s;
    Domain *d_domain;           /* device */
    Domain h_domain;            /* host copy of the last call (stats) */
    int max_cells;
    int sort_bits;
    void *cub_tmp;
    size_t cub_tmp_bytes;
    int *keys_in, *idx_in;
    double *block_partials;     /* bbox / h reductions */
    unsigned int *block_counter;
    int *d_flags;               /* [0] offender slot (min caller index with overflow), [1] max noi, [2..3] total noi (64 bit) */
    int materials_set;
    int kernel_sum_density;     /* 1 if k_density must run (no INTEGRATE_DENSITY, or a material with density_via_kernel_sum) */
    double *rho_sorted;
    double *aneos_buf;          /* device copy of the tabulated-EOS payload */
    int n_owned;
    int halo_sums_external;     /* b200sph_set_halo_sums */
    const int *abort_flag;      /* b200sph_set_abort_flag */
    int lists_validated, stage_launches;   /* carried between the stages of one evaluation */
    int overlap_pointwise, pointwise_early;   /* k_pointwise beside k_neighbours on aux_stream */
    cudaStream_t aux_stream;
    cudaEvent_t ev_fork, ev_join;
    const double *grav_src[4];  /* x, y, z, m of the global particle set (multi-GPU gravity), device pointers */
    int grav_src_n, grav_own_begin;
    int pad_smem;               /* profiling knob (B200SPH_PAD_SMEM): dynamic shared memory per pair-loop block, throttles occupancy */
    int forces_threads;         /* block size of k_forces: small blocks keep more warps resident at high register counts */
    int have_global_domain;
    double global_lo[3], global_hi[3];
    void *halo;                 /* HaloState (halo.cu): domain boxes and selection scratch */
    /* host-view staging (b200sph_rhs_eval_host) */
    void *stage;
    size_t stage_bytes;
    cudaStream_t copy_stream;   /* second DMA queue: late inputs / early outputs overlap the kernels */
    cudaEvent_t ev_copy[4];     /* [0] first-stage inputs queued, [1] late inputs landed, [2] k_pointwise done */
    int host_options;           /* B200SPH_HOST_* bits */
    int host_imm_valid;         /* immutables of host_imm_key are resident in `stage` */
    const void *host_imm_key[6];/* host pointers (m, h0, materialId, numFlaws, flaws, x) the cached copy came from */
    int host_imm_n;
    /* hooks of b200sph_rhs_eval used by the host-buffer entry point (NULL otherwise) */
    cudaEvent_t hook_wait_before_pointwise;
    void (*hook_after_pointwise)(struct b200sph_handle *, void *);
    void *hook_ctx;
    /* integrate.cu on several GPUs (set by mg.cu): the evaluation goes through the halo exchange, and the step-size
     * reductions (device doubles) are all-reduced over the ranks, min or max */
    int (*rk_rhs_hook)(void *ctx, const b200sph_view *bound, int *offender);
    int (*rk_allreduce)(void *ctx, double *dev_values, int n, int is_min);
    void *rk_hook_ctx;
    void *rk_scalars;           /* integrate.cu: device step state, reduction partials, ticket */
    double *rk_partials;
    unsigned int *rk_counter;
    b200sph_stats stats;
    /* gravity */
    struct GravityTree *tree;
    int gravity_index, flag_force_gravity_calc;
    char err[512];
};

#endif
