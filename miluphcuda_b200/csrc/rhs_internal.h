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
 *   pos4[n]   = {x, y, z, h}    32-byte records, one LDG.128 pair per candidate
 *   vel4[n]   = {vx, vy, vz, m}
 *   gas4[n]   = {p/rho^2 (hydro) or 1/rho^2 (solid), c_s, rho, m/rho}
 *   sig[n*DD], cmat[n*DD], rart[n*DD]   sigma/rho^2, correction matrix, R/rho^2 (solid)
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
};

struct Sorted {
    int n;
    int *perm;                  /* sorted slot -> caller index */
    int *keys;
    int *cell_start;
    double4 *pos4, *vel4, *gas4;
    int *mat;
    double *sig, *cmat, *rart;
    int *nbr, *noi;
};

struct b200sph_handle {
    int n_max, device;
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
    int have_global_domain;
    double global_lo[3], global_hi[3];
    /* host-view staging (b200sph_rhs_eval_host) */
    void *stage;
    size_t stage_bytes;
    b200sph_stats stats;
    /* gravity */
    struct GravityTree *tree;
    int gravity_index, flag_force_gravity_calc;
    char err[512];
};

#endif
