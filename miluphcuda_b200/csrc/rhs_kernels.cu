/*
 * rhs_kernels.cu -- the SPH right-hand side on one B200.
 *
 * Replaces the ~20 kernels miluphcuda's rightHandSide() launches
 * (reference: src/rhs.cu:181-840) by eight:
 *
 *   k_prepare      floors / h clamp (boundary hooks, check_sml_boundary) + bbox + h statistics
 *   k_cell_keys    uniform-grid cell index per particle          } replaces the lock-based octree
 *   (cub radix sort of (cell, index))                            } build and the per-thread DFS
 *   k_cell_start   first sorted slot of every cell               } (src/tree.cu:71-267, 786-926)
 *   k_gather       caller order -> 32-byte cell-sorted records
 *   k_neighbours   exact neighbour lists, tile-interleaved
 *   k_density      kernel-sum density                            (src/density.cu:41-209)
 *   k_pointwise    c_s, p, p-alpha, symmetrise S, damage limit, yield, sigma, artificial stress
 *                  (src/soundspeed.cu, pressure.cu, timeintegration.cu:116, damage.cu, plasticity.cu,
 *                   stress.cu, artificial_stress.cu -- seven reference launches fused)
 *   k_correction   tensorial correction matrix                   (src/kernel.cu:585-713)
 *   k_forces       pair loop + per-particle epilogue             (src/internal_forces.cu:41-1258,
 *                  boundary.cu:215-329, velocity.cu:31-57)
 *
 * Particles are never reordered in the caller's buffers: every kernel after the
 * sort works on cell-sorted scratch and writes results back through `perm`.
 */
#include "rhs_internal.h"

#include <cub/cub.cuh>
#include <math.h>
#include <stdio.h>

#define FULL_MASK 0xffffffffu

/* ------------------------------------------------------------------ helpers */
__device__ __forceinline__ double warp_min(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}

__device__ __forceinline__ double coord_of(const b200sph_particle_arrays &p, int i, int axis)
{
#if DIM > 2
    if (axis == 2) return p.z[i];
#endif
#if DIM > 1
    if (axis == 1) return p.y[i];
#endif
    return p.x[i];
}

/* ------------------------------------------------------------------ k_prepare
 * values per block: min[3], max[3], sum h, max h, min h, number of non-finite coordinates,
 * number of frozen particles (their velocities were zeroed) */
/* unroll factor of the pair loops.  Measured on a B200 (gpurun_out v6, round 1): two pairs per iteration take the
 * hydro force loop from 0.354 to 0.324 ms (two independent FP64 chains and twice the loads in flight per warp, at
 * 124 instead of 106 registers); the solid loops already sit at 168 registers and keep one pair per iteration. */
#ifndef B200_PAIR_UNROLL
#if SOLID
#define B200_PAIR_UNROLL 1
#else
#define B200_PAIR_UNROLL 2
#endif
#endif
constexpr int kPairUnroll = B200_PAIR_UNROLL;
#if B200_PAIR_UNROLL > 1
#define PAIR_UNROLL _Pragma("unroll kPairUnroll")
#else
#define PAIR_UNROLL   /* no pragma at all: "#pragma unroll 1" would change the code the compiler emits today */
#endif
constexpr int kSearchUnroll = 4;   /* candidates fetched per trip of the search loop */
#define PREP_VALUES 11
#define PREP_THREADS 256

__global__ void __launch_bounds__(PREP_THREADS)
k_prepare(b200sph_view v, double *partials, unsigned int *counter, Domain *dom, int max_cells,
          int use_global, double3 glo, double3 ghi, int apply_hooks)
{
    const b200sph_particle_arrays &p = v.p;
    const b200sph_particle_arrays &pr = v.p_rhs;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300}, hsum = 0.0, hmax = 0.0, hmin = 1e300, bad = 0.0, frozen = 0.0;

    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < v.n; i += gridDim.x * blockDim.x) {
        const int matId = pr.materialId[i];
        if (!apply_hooks) {
            /* b200sph_reorder: only the bounding box and the h statistics are wanted, the state is left alone */
        } else if (matId == EOS_TYPE_IGNORE || matId == BOUNDARY_PARTICLE_ID) {
            /* BoundaryConditionsBeforeRHS, src/boundary.cu:98-145: deactivated particles are frozen */
            p.vx[i] = 0.0;
#if DIM > 1
            p.vy[i] = 0.0;
#endif
#if DIM > 2
            p.vz[i] = 0.0;
#endif
            frozen += 1.0;
        }
        if (apply_hooks && matId >= 0 && matId != BOUNDARY_PARTICLE_ID) {
            const MatParams &M = c_mat[matId];
            if (p.rho[i] < M.density_floor) p.rho[i] = M.density_floor;
            if (p.e && p.e[i] < M.energy_floor) p.e[i] = M.energy_floor;
#if VARIABLE_SML
            /* check_sml_boundary, src/tree.cu:930-952 */
            const double smin = pr.h0[i] * M.f_sml_min, smax = pr.h0[i] * M.f_sml_max;
            if (p.h[i] < smin) p.h[i] = smin;
            else if (p.h[i] > smax) p.h[i] = smax;
#endif
        }
        const double h = p.h[i];
        hsum += h;
        hmax = fmax(hmax, h);
        hmin = fmin(hmin, h);
#pragma unroll
        for (int a = 0; a < DIM; a++) {
            const double c = coord_of(p, i, a);
            lo[a] = fmin(lo[a], c);
            hi[a] = fmax(hi[a], c);
            if (!isfinite(c)) bad += 1.0;   /* fmin/fmax drop NaNs silently */
        }
    }

    __shared__ double sh[PREP_THREADS / 32][PREP_VALUES];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double vals[PREP_VALUES];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        vals[a] = warp_min(lo[a]);
        vals[3 + a] = warp_max(hi[a]);
    }
    vals[6] = warp_sum(hsum);
    vals[7] = warp_max(hmax);
    vals[8] = warp_min(hmin);
    vals[9] = warp_sum(bad);
    vals[10] = warp_sum(frozen);
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < PREP_VALUES; k++) sh[warp][k] = vals[k];
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < PREP_THREADS / 32; w++) {
            for (int a = 0; a < 3; a++) {
                sh[0][a] = fmin(sh[0][a], sh[w][a]);
                sh[0][3 + a] = fmax(sh[0][3 + a], sh[w][3 + a]);
            }
            sh[0][6] += sh[w][6];
            sh[0][7] = fmax(sh[0][7], sh[w][7]);
            sh[0][8] = fmin(sh[0][8], sh[w][8]);
            sh[0][9] += sh[w][9];
            sh[0][10] += sh[w][10];
        }
        for (int k = 0; k < PREP_VALUES; k++) partials[blockIdx.x * PREP_VALUES + k] = sh[0][k];
        __threadfence();
        const unsigned int ticket = atomicAdd(counter, 1u);
        is_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;

    /* last block: all of its threads combine the per-block partials (one thread doing this alone cost
     * 170 us for 592 blocks, profiles/r01_launches_sedov_v2.csv), then thread 0 sets up the search grid */
    __threadfence();
    double r[PREP_VALUES];
#pragma unroll
    for (int a = 0; a < 3; a++) { r[a] = 1e300; r[3 + a] = -1e300; }
    r[6] = 0.0; r[7] = 0.0; r[8] = 1e300; r[9] = 0.0; r[10] = 0.0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
        const double *q = partials + b * PREP_VALUES;
#pragma unroll
        for (int a = 0; a < 3; a++) {
            r[a] = fmin(r[a], __ldcg(q + a));
            r[3 + a] = fmax(r[3 + a], __ldcg(q + 3 + a));
        }
        r[6] += __ldcg(q + 6);
        r[7] = fmax(r[7], __ldcg(q + 7));
        r[8] = fmin(r[8], __ldcg(q + 8));
        r[9] += __ldcg(q + 9);
        r[10] += __ldcg(q + 10);
    }
#pragma unroll
    for (int a = 0; a < 3; a++) {
        r[a] = warp_min(r[a]);
        r[3 + a] = warp_max(r[3 + a]);
    }
    r[6] = warp_sum(r[6]);
    r[7] = warp_max(r[7]);
    r[8] = warp_min(r[8]);
    r[9] = warp_sum(r[9]);
    r[10] = warp_sum(r[10]);
    __syncthreads();   /* sh[] is being reused */
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < PREP_VALUES; k++) sh[warp][k] = r[k];
    __syncthreads();
    if (threadIdx.x != 0) return;
    for (int k = 0; k < PREP_VALUES; k++) r[k] = sh[0][k];
    for (int w = 1; w < PREP_THREADS / 32; w++) {
        for (int a = 0; a < 3; a++) {
            r[a] = fmin(r[a], sh[w][a]);
            r[3 + a] = fmax(r[3 + a], sh[w][3 + a]);
        }
        r[6] += sh[w][6];
        r[7] = fmax(r[7], sh[w][7]);
        r[8] = fmin(r[8], sh[w][8]);
        r[9] += sh[w][9];
        r[10] += sh[w][10];
    }
    *counter = 0;
    Domain d;
    for (int a = 0; a < 3; a++) {
        d.lo[a] = (a < DIM) ? r[a] : 0.0;
        d.hi[a] = (a < DIM) ? r[3 + a] : 0.0;
    }
    if (use_global) {
        d.lo[0] = glo.x; d.lo[1] = glo.y; d.lo[2] = glo.z;
        d.hi[0] = ghi.x; d.hi[1] = ghi.y; d.hi[2] = ghi.z;
    }
    /* root cube of the reference's octree (src/tree.cu:1071-1086) */
    double radius = d.hi[0] - d.lo[0];
#if DIM > 1
    radius = fmax(d.hi[0] - d.lo[0], d.hi[1] - d.lo[1]);
#endif
#if DIM > 2
    radius = fmax(radius, d.hi[2] - d.lo[2]);
#endif
    d.root_radius = 0.5 * radius;
    for (int a = 0; a < 3; a++) d.root_centre[a] = 0.5 * (d.hi[a] + d.lo[a]);
    d.h_max = r[7];
    d.h_mean = r[6] / (double)v.n;
    d.n_frozen = (int)r[10];
    /* NaN/Inf coordinates or smoothing lengths: no grid can be built; report instead of indexing with garbage */
    bool finite = isfinite(r[6]) && isfinite(r[7]) && r[8] > 0.0 && r[9] == 0.0;
    for (int a = 0; a < DIM; a++) finite = finite && isfinite(d.lo[a]) && isfinite(d.hi[a]);
    if (!finite) {
        for (int a = 0; a < 3; a++) { d.lo[a] = 0.0; d.hi[a] = 0.0; d.root_centre[a] = 0.0; d.nc[a] = 1; }
        d.root_radius = 0.0; d.cell = 1.0; d.cell_inv = 1.0; d.n_cells = 1; d.h_max = 0.0; d.h_mean = 0.0;
        d.nonfinite = 1;
        *dom = d;
        return;
    }
    d.nonfinite = 0;
#if VARIABLE_SML
    /* Cells as small as the smallest smoothing lengths: the bulk of a variable-resolution set sits at
     * the finest resolution, and a cell edge of 1.3 h_mean made those particles test ~1200 candidates
     * for ~35 neighbours (profiles/r01_ncu_full_impact_v1_summary.csv).  Coarse particles reach further
     * (stencil_of) but live where cells are nearly empty.  Floored at h_mean/2 against outliers. */
    double cell = fmin(d.h_max, fmax(r[8], 0.5 * d.h_mean)) * 1.0001;
#else
    /* fixed h: cells of half the smoothing length, 5x5(x5) stencil clipped to the sphere (k_neighbours) */
    double cell = d.h_max * (1.0001 / 2.0);
#endif
    if (!(cell > 0.0)) cell = 1.0;
    for (;;) {
        /* counted in double: a far-flung particle makes the first guesses overflow an int */
        double cells = 1.0, ncd[3];
        for (int a = 0; a < 3; a++) {
            ncd[a] = (a < DIM) ? floor((d.hi[a] - d.lo[a]) / cell) + 1.0 : 1.0;
            cells *= ncd[a];
        }
        if (cells <= (double)max_cells) {
            for (int a = 0; a < 3; a++) d.nc[a] = (int)ncd[a];
            d.n_cells = d.nc[0] * d.nc[1] * d.nc[2];
            break;
        }
        cell *= 1.26;
    }
    d.cell = cell;
    d.cell_inv = 1.0 / cell;
    *dom = d;
}

__device__ __forceinline__ int cell_coord(double x, double lo, double cell_inv, int nc)
{
    int c = (int)((x - lo) * cell_inv);
    c = c < 0 ? 0 : c;
    return c >= nc ? nc - 1 : c;
}

__global__ void k_cell_keys(b200sph_view v, const Domain *dom, int *keys, int *idx)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.n) return;
    const Domain &d = *dom;
    int key = cell_coord(v.p.x[i], d.lo[0], d.cell_inv, d.nc[0]);
#if DIM > 1
    key += d.nc[0] * cell_coord(v.p.y[i], d.lo[1], d.cell_inv, d.nc[1]);
#endif
#if DIM > 2
    key += d.nc[0] * d.nc[1] * cell_coord(v.p.z[i], d.lo[2], d.cell_inv, d.nc[2]);
#endif
    keys[i] = key;
    idx[i] = i;
}

/* cell_start[c] = first sorted slot whose key >= c, for c in [0, n_cells].  One thread per sorted slot
 * fills the cells between its predecessor's key and its own; runs of empty cells longer than 32 (sparse
 * ejecta, variable resolution) are filled by the whole warp instead of one lane. */
__global__ void k_cell_start(const int *keys, int n, const Domain *dom, int *cell_start)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int n_cells = dom->n_cells;
    int prev = 0, cur = -1;
    if (s <= n) {
        prev = (s == 0) ? -1 : keys[s - 1];
        cur = (s == n) ? n_cells : keys[s];
    }
    const bool long_run = cur - prev > 32;
    if (!long_run)
        for (int c = prev + 1; c <= cur; c++) cell_start[c] = s;
    unsigned int todo = __ballot_sync(FULL_MASK, long_run);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int p0 = __shfl_sync(FULL_MASK, prev, src), c0 = __shfl_sync(FULL_MASK, cur, src);
        const int s0 = __shfl_sync(FULL_MASK, s, src);
        for (int c = p0 + 1 + lane; c <= c0; c += 32) cell_start[c] = s0;
    }
}

/* FP32 pre-filter of the neighbour search.  Positions are stored in cell units, u = (x - lo)/cell,
 * rounded to FP32; a candidate is kept when its FP32 distance is below BOTH particles' thresholds
 * thr = (h/cell)^2 + margin.  The margin bounds every rounding on the FP32 path, so the filter never
 * rejects a pair the exact FP64 test accepts; the few extra survivors (a shell of relative width
 * ~1e-4) are removed by the exact test in the first pair loop that walks the list (LIST_VALIDATE).
 *   |u_hat - u| <= 2^-24 nc          (conversion; nc = cells along the longest axis)
 *   |D_hat - D| <= delta = 2^-22 nc + 2^-22 (hc + 1)     per axis (two conversions + the subtraction)
 *   |d_hat - d| <= 2 sqrt(3) hc delta + 3 delta^2 + 2^-20 hc^2   for d <= hc^2 (hc = h/cell) */
__device__ __forceinline__ float search_threshold(double h, const Domain &d)
{
    const double ncm = (double)max(d.nc[0], max(d.nc[1], d.nc[2]));
    const double hc = h * d.cell_inv;
    const double delta = 2.384185791015625e-07 * (ncm + hc + 1.0);
    const double margin = 3.4641016151377544 * hc * delta + 3.0 * delta * delta + 9.5367431640625e-07 * hc * hc;
    return __double2float_ru((hc * hc + 2.0 * margin) * 1.000001);
}

__global__ void k_gather(b200sph_view v, Sorted s, const Domain *dom)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= s.n) return;
    const int i = s.perm[k];
    const Domain &d = *dom;
    const b200sph_particle_arrays &p = v.p;
    Rec4 a, b;
    a.x = p.x[i]; b.x = p.vx[i];
#if DIM > 1
    a.y = p.y[i]; b.y = p.vy[i];
#else
    a.y = 0.0; b.y = 0.0;
#endif
#if DIM > 2
    a.z = p.z[i]; b.z = p.vz[i];
#else
    a.z = 0.0; b.z = 0.0;
#endif
    a.w = p.h[i];
    b.w = p.m[i];
    st_rec(&s.pos4[k], a);
    st_rec(&s.vel4[k], b);
    const int matId = v.p_rhs.materialId[i];
    s.mat[k] = matId;
    float4 f;
    f.x = (float)((a.x - d.lo[0]) * d.cell_inv);
    f.y = (DIM > 1) ? (float)((a.y - d.lo[1]) * d.cell_inv) : 0.0f;
    f.z = (DIM > 2) ? (float)((a.z - d.lo[2]) * d.cell_inv) : 0.0f;
    /* deactivated particles are never anybody's neighbour (src/tree.cu:845-847) */
    f.w = (matId == EOS_TYPE_IGNORE) ? -1.0f : search_threshold(a.w, d);
    s.srch[k] = f;
}

/* ------------------------------------------------------------------ k_neighbours
 * Membership: d < h_i^2 && d < h_j^2, j != i, materialId[j] != IGNORE, with d accumulated as the
 * reference's compiled code does (src/tree.cu:851-865: mul, then one fma per further axis). */
__device__ __forceinline__ double pair_d2(const Rec4 &a, const Rec4 &b, double &dx, double &dy, double &dz)
{
    dx = a.x - b.x;
    double d = __dmul_rn(dx, dx);
#if DIM > 1
    dy = a.y - b.y;
    d = __fma_rn(dy, dy, d);
#else
    dy = 0.0;
#endif
#if DIM > 2
    dz = a.z - b.z;
    d = __fma_rn(dz, dz, d);
#else
    dz = 0.0;
#endif
    return d;
}

#define NBR_SLOT(s, k) ((((size_t)((s) / NBR_TILE)) * MAX_NUM_INTERACTIONS + (k)) * NBR_TILE + ((s) % NBR_TILE))

struct Stencil {
    int x0, x1, y0, y1, z0, z1;
};

__device__ __forceinline__ Stencil stencil_of(const Rec4 &pi, const Domain &d)
{
    Stencil st;
    const int reach = (int)(pi.w * d.cell_inv + 1e-9) + 1;
    const int cx = cell_coord(pi.x, d.lo[0], d.cell_inv, d.nc[0]);
    st.x0 = max(cx - reach, 0); st.x1 = min(cx + reach, d.nc[0] - 1);
#if DIM > 1
    const int cy = cell_coord(pi.y, d.lo[1], d.cell_inv, d.nc[1]);
    st.y0 = max(cy - reach, 0); st.y1 = min(cy + reach, d.nc[1] - 1);
#else
    st.y0 = 0; st.y1 = 0;
#endif
#if DIM > 2
    const int cz = cell_coord(pi.z, d.lo[2], d.cell_inv, d.nc[2]);
    st.z0 = max(cz - reach, 0); st.z1 = min(cz + reach, d.nc[2] - 1);
#else
    st.z0 = 0; st.z1 = 0;
#endif
    return st;
}

/* exact FP64 scan; only used for a particle whose pre-filtered candidates do not fit the list */
__device__ __noinline__ int neighbours_exact(const Sorted &s, const Domain &d, int t, int k, const Rec4 &pi, const Stencil &st)
{
    const double h2 = __dmul_rn(pi.w, pi.w);
    int cnt = 0;
    for (int z = st.z0; z <= st.z1; z++)
        for (int y = st.y0; y <= st.y1; y++) {
            const int row = d.nc[0] * (y + d.nc[1] * z);
            const int jb = s.cell_start[row + st.x0], je = s.cell_start[row + st.x1 + 1];
            for (int j = jb; j < je; j++) {
                const Rec4 pj = ld_rec(&s.pos4[j]);
                double dx, dy, dz;
                const double dd = pair_d2(pi, pj, dx, dy, dz);
                if (dd < h2 && dd < __dmul_rn(pj.w, pj.w) && j != k && s.mat[j] != EOS_TYPE_IGNORE) {
                    if (cnt < MAX_NUM_INTERACTIONS) s.nbr[NBR_SLOT(t, cnt)] = j;
                    cnt++;
                }
            }
        }
    return cnt;
}

/* A halo copy needs its own neighbour list only when one of this rank's particles reads a neighbour SUM of
 * it (kernel-sum density, tensorial correction matrix) -- that is, when it can be a neighbour of an owned
 * particle at all: distance to this rank's boxes below its own h (criterion d < h_i^2 && d < h_j^2).  The
 * outer halo level (copies that only complete those sums) and every copy of a one-level halo skip the search. */
__device__ __forceinline__ bool halo_copy_needs_list(const Rec4 &pi, const HaloDomains *hd)
{
    /* the send plan lets owned particles drift up to its skin out of their boxes and h grow by its growth factor
     * before it is rebuilt: a copy can then be a neighbour of an owned particle from that much further away */
    const double reach = pi.w * hd->list_reach_scale * (1.0 + 1e-9) + hd->list_skin;
    const double reach2 = reach * reach;
    const double pos[3] = {pi.x, pi.y, pi.z};
    const int first = hd->my_first, count = hd->my_count;
    for (int b = 0; b < count; b++) {
        double g2 = 0.0;
#pragma unroll
        for (int a = 0; a < DIM; a++) {
            const double g = fmax(fmax(hd->lo[first + b][a] - pos[a], pos[a] - hd->hi[first + b][a]), 0.0);
            g2 = fma(g, g, g2);
        }
        if (g2 < reach2) return true;
    }
    return false;
}

/* distance (>= 0) along one axis between coordinate p and the cells with index c; the outermost cells
 * also hold the particles beyond the grid (cell_coord clamps), so they extend to infinity */
__device__ __forceinline__ double row_gap(double p, double lo, double cell, int c, int nc, double slack)
{
    const double c_lo = fma((double)c, cell, lo);
    const double below = (c > 0) ? c_lo - p : -1.0;
    const double above = (c < nc - 1) ? p - (c_lo + cell) : -1.0;
    return fmax(fmax(below, above) - slack, 0.0);
}

__device__ __forceinline__ bool search_hit(const float4 &si, float thr_i, const float4 c)
{
    const float dx = si.x - c.x;
    float dd = dx * dx;
#if DIM > 1
    const float dy = si.y - c.y;
    dd = fmaf(dy, dy, dd);
#endif
#if DIM > 2
    const float dz = si.z - c.z;
    dd = fmaf(dz, dz, dd);
#endif
    return dd < fminf(thr_i, c.w);
}

__global__ void __launch_bounds__(128)
k_neighbours(Sorted s, const Domain *dom, int n_targets, int *flags, const HaloDomains *hd, int halo_sums)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_targets) return;
    const Domain &d = *dom;
    if (d.nonfinite) {   /* void evaluation (reported by the host): do not scan the degenerate one-cell grid */
        s.noi[t] = 0;
        return;
    }
    const int k = t;
    const Rec4 pi = ld_rec(&s.pos4[k]);
    if (s.perm[k] >= s.n_owned && (s.halo_sums_external || (hd != nullptr && (!halo_sums || !halo_copy_needs_list(pi, hd))))) {
        s.noi[t] = 0;
        return;
    }
    const float4 si = s.srch[k];
    const float thr_i = search_threshold(pi.w, d);   /* srch.w is -1 for a deactivated target, which still collects neighbours */
    const Stencil st = stencil_of(pi, d);
    /* The particle itself passes the filter (distance 0) and is stored like any survivor; the
     * validating pair loop drops it.  That keeps the j != k comparison out of the candidate loop.
     *
     * Rows of x-adjacent cells are clipped to the sphere: a neighbour in row (y, z) is at least
     * (g_y, g_z) away in y and z -- the gaps between the particle and the row's cells -- so it lies within
     * w = sqrt(h_i^2 - g_y^2 - g_z^2) in x; rows with w^2 <= 0 are skipped.  With cells of half the largest h
     * (k_prepare) this leaves ~120 candidates for ~47 hits instead of the 300 of a 3x3x3 block of h-sized cells. */
    int cnt = 0;
    int *const base = s.nbr + NBR_SLOT(t, 0);
    const double reach2 = __dmul_rn(pi.w, pi.w) * (1.0 + 1e-9);
    const double slack = 1e-9 * d.cell;
    for (int z = st.z0; z <= st.z1; z++) {
#if DIM > 2
        const double gz = row_gap(pi.z, d.lo[2], d.cell, z, d.nc[2], slack);
        const double rem_z = reach2 - gz * gz;
        if (rem_z <= 0.0) continue;
#else
        const double rem_z = reach2;
#endif
        for (int y = st.y0; y <= st.y1; y++) {
#if DIM > 1
            const double gy = row_gap(pi.y, d.lo[1], d.cell, y, d.nc[1], slack);
            const double rem = rem_z - gy * gy;
            if (rem <= 0.0) continue;
#else
            const double rem = rem_z;
#endif
            /* FP32 square root rounded up, then widened: never narrower than the exact half-width */
            const double w = (double)__fsqrt_ru(__double2float_ru(rem)) * (1.0 + 1e-6) + slack;
            const int xa = max(st.x0, cell_coord(pi.x - w, d.lo[0], d.cell_inv, d.nc[0]));
            const int xb = min(st.x1, cell_coord(pi.x + w, d.lo[0], d.cell_inv, d.nc[0]));
            const int row = d.nc[0] * (y + d.nc[1] * z);
            const int jb = s.cell_start[row + xa], je = s.cell_start[row + xb + 1];
            if (cnt + (je - jb) <= MAX_NUM_INTERACTIONS) {
                /* the whole row fits: no overflow check per candidate */
                int *slot = base + cnt * NBR_TILE;
#pragma unroll kSearchUnroll
                for (int j = jb; j < je; j++) {
                    const bool hit = search_hit(si, thr_i, __ldg(&s.srch[j]));
                    if (hit) *slot = j;   /* predicated store, no divergent branch */
                    slot += hit ? NBR_TILE : 0;
                    cnt += hit ? 1 : 0;
                }
            } else {
                for (int j = jb; j < je; j++) {
                    if (search_hit(si, thr_i, __ldg(&s.srch[j]))) {
                        if (cnt < MAX_NUM_INTERACTIONS) base[cnt * NBR_TILE] = j;
                        cnt++;
                    }
                }
            }
        }
    }
    if (cnt > MAX_NUM_INTERACTIONS) {
        /* more survivors than list slots: decide with the exact test (the reference asserts on the exact count) */
        cnt = neighbours_exact(s, d, t, k, pi, st);
        if (cnt >= MAX_NUM_INTERACTIONS) {
            atomicMin(&flags[0], s.perm[k]);
            cnt = MAX_NUM_INTERACTIONS - 1;
        }
    }
    s.noi[t] = cnt;   /* list slots in use; the exact count replaces it in the LIST_VALIDATE pass */
}

/* How a pair loop treats the list it walks:
 *   LIST_EXACT     entries are exact neighbours (an earlier pass validated the list)
 *   LIST_VALIDATE  first walk after the search: apply the exact FP64 membership test, compact the
 *                  list in place, store the exact count, flag an overflow
 *   LIST_CHECK     apply the exact test but leave the list alone (a pass that does not visit every particle) */
enum { LIST_EXACT = 0, LIST_VALIDATE = 1, LIST_CHECK = 2 };

__device__ __forceinline__ bool pair_is_neighbour(double r2, double h2_i, const Rec4 &pj)
{
    return r2 < h2_i && r2 < __dmul_rn(pj.w, pj.w);
}

/* walk a list only to validate it (particles whose pair loop is skipped) */
__device__ __noinline__ int validate_only(const Sorted &s, int t, int k, const Rec4 &pi, int nslots)
{
    const double h2 = __dmul_rn(pi.w, pi.w);
    int cnt = 0;
    for (int q = 0; q < nslots; q++) {
        const int j = s.nbr[NBR_SLOT(t, q)];
        const Rec4 pj = ld_rec(&s.pos4[j]);
        double dx, dy, dz;
        const double r2 = pair_d2(pi, pj, dx, dy, dz);
        if (j == k || !pair_is_neighbour(r2, h2, pj)) continue;
        if (cnt != q) s.nbr[NBR_SLOT(t, cnt)] = j;
        cnt++;
    }
    return cnt;
}

__device__ __forceinline__ int finish_validate(const Sorted &s, int t, int k, int cnt, int *flags)
{
    if (cnt >= MAX_NUM_INTERACTIONS) { /* the reference asserts here (src/tree.cu:917) */
        atomicMin(&flags[0], s.perm[k]);
        cnt = MAX_NUM_INTERACTIONS - 1;
    }
    s.noi[t] = cnt;
    return cnt;
}

/* max and sum of the exact interaction counts (stats; one atomic pair per block) */
__global__ void __launch_bounds__(256) k_list_stats(Sorted s, int *flags)
{
    int mx = 0, sum = 0;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < s.n; k += gridDim.x * blockDim.x) {
        const int c = s.noi[k];
        mx = max(mx, c);
        sum += c;
    }
    mx = __reduce_max_sync(FULL_MASK, mx);
    sum = __reduce_add_sync(FULL_MASK, sum);
    __shared__ int smx[8], ssum[8];
    if ((threadIdx.x & 31) == 0) { smx[threadIdx.x >> 5] = mx; ssum[threadIdx.x >> 5] = sum; }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long tot = 0;
        for (int w = 0; w < 8; w++) { mx = max(mx, smx[w]); tot += ssum[w]; }
        atomicMax(&flags[1], mx);
        atomicAdd(reinterpret_cast<unsigned long long *>(flags + 2), (unsigned long long)tot);
    }
}

/* Everything a pair loop gathers about neighbour j.  The loops are software-pipelined: the list index
 * is fetched two iterations ahead and the records one iteration ahead, so the dependent chain
 * index -> records (two L2 round trips) overlaps the FP64 work of the current pair
 * (profiles/r01_ncu_full_*_v1: 61 % of the stall samples of k_forces were long-scoreboard waits). */
struct PairRecs {
    Rec4 p, v, g;
#if SOLID
    Rec4 t[TEN_RECS];
#endif
};

__device__ __forceinline__ void load_tensor_recs(const Sorted &s, int j, PairRecs &r)
{
#if SOLID
#pragma unroll
    for (int c = 0; c < TEN_RECS; c++) r.t[c] = ld_rec(&s.ten[(size_t)j * TEN_RECS + c]);
#endif
}

__device__ __forceinline__ void load_force_recs(const Sorted &s, int j, PairRecs &r)
{
    r.p = ld_rec(&s.pos4[j]);
    r.v = ld_rec(&s.vel4[j]);
    r.g = ld_rec(&s.gas4[j]);
}

/* ------------------------------------------------------------------ k_density */
template <int MODE>
__global__ void __launch_bounds__(128)
k_density(Sorted s, b200sph_view v, double *rho_sorted, int n_targets, int *flags)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_targets) return;
    const int k = t;
    const int matId = s.mat[k];
    const int i = s.perm[k];
    const Rec4 pi = ld_rec(&s.pos4[k]);
    const int nslots = s.noi[t];
    if (s.abort && *s.abort) return;
    if (s.halo_sums_external && i >= s.n_owned) return;   /* the owner sends this copy's density (b200sph_rhs_eval_stage) */
#if INTEGRATE_DENSITY
    if (mat_ignored(matId) || c_mat[matId].density_via_kernel_sum < 1) {
        rho_sorted[k] = v.p.rho[i];
        if (MODE == LIST_VALIDATE) finish_validate(s, t, k, validate_only(s, t, k, pi, nslots), flags);
        return;
    }
#endif
    if (mat_ignored(matId)) {
        const double rho = ld_rec(&s.vel4[k]).w * cubic_spline_w(0.0, 1.0 / pi.w);
        rho_sorted[k] = rho;
        v.p.rho[i] = rho;
        if (MODE == LIST_VALIDATE) finish_validate(s, t, k, validate_only(s, t, k, pi, nslots), flags);
        return;
    }
    const double hinv_i = 1.0 / pi.w;
    const double h2_i = __dmul_rn(pi.w, pi.w);
    double rho = ld_rec(&s.vel4[k]).w * cubic_spline_w(0.0, hinv_i);
    int cnt = 0;
    int j_next = 0, j_next2 = 0;
    Rec4 pj_next = pi;
    double mj_next = 0.0;
    if (nslots > 0) {
        j_next = s.nbr[NBR_SLOT(t, 0)];
        j_next2 = s.nbr[NBR_SLOT(t, min(1, nslots - 1))];
        pj_next = ld_rec(&s.pos4[j_next]);
        mj_next = s.vel4[j_next].w;
    }
PAIR_UNROLL
    for (int q = 0; q < nslots; q++) {
        const int j = j_next;
        const Rec4 pj = pj_next;
        const double mj = mj_next;
        j_next = j_next2;
        j_next2 = s.nbr[NBR_SLOT(t, min(q + 2, nslots - 1))];
        pj_next = ld_rec(&s.pos4[j_next]);
        mj_next = s.vel4[j_next].w;
        double dx, dy, dz, W, g;
        const double r2 = pair_d2(pi, pj, dx, dy, dz);
        if (MODE != LIST_EXACT && (j == k || !pair_is_neighbour(r2, h2_i, pj))) continue;
        if (MODE == LIST_VALIDATE) {
            if (cnt != q) s.nbr[NBR_SLOT(t, cnt)] = j;
            cnt++;
        }
        if (s.any_eos_ignore && mat_ignored(s.mat[j])) continue;
#if AVERAGE_KERNELS
        /* W = (W(h_i) + W(h_j))/2.  The reference evaluates the second kernel with the
         * smoothing length of the particle whose index equals the LOOP COUNTER
         * (src/density.cu:130-138); identical whenever h is uniform, which holds for every
         * config that sets AVERAGE_KERNELS (fixed sml per material). */
        cubic_spline(r2, hinv_i, W, g);
        if (pj.w != pi.w) {
            double Wj;
            cubic_spline(r2, 1.0 / pj.w, Wj, g);
            W = 0.5 * (W + Wj);
        }
#elif VARIABLE_SML || INTEGRATE_SML
        cubic_spline(r2, 1.0 / (0.5 * (pi.w + pj.w)), W, g);
#else
        cubic_spline(r2, hinv_i, W, g);
#endif
        rho = fma(mj, W, rho);
    }
    if (MODE == LIST_VALIDATE) finish_validate(s, t, k, cnt, flags);
    rho_sorted[k] = rho;
    v.p.rho[i] = rho;
}

/* ------------------------------------------------------------------ k_pointwise */
__device__ inline double eos_soundspeed(const MatParams &M, double rho, double e, double p_old, double alpha_jutzi, double cs_in)
{
    const double limit = M.cs_limit;
    switch (M.eos) {
        case EOS_TYPE_POLYTROPIC_GAS: return sqrt(M.poly_K * pow(rho, M.poly_gamma - 1.0));
        case EOS_TYPE_IDEAL_GAS: return sqrt(M.poly_gamma * p_old / rho);
        case EOS_TYPE_ISOTHERMAL_GAS: return M.iso_cs;
        case EOS_TYPE_MURNAGHAN: {
            const double cs_sq = M.bulk / M.rho0 * pow(rho / M.rho0, M.n - 1.0);
            return cs_sq < limit * limit ? limit : sqrt(cs_sq);
        }
        case EOS_TYPE_TILLOTSON: {
            const double cs_sq = tillotson_cs2(M, rho, e, p_old);
            return cs_sq < limit * limit ? limit : sqrt(cs_sq);
        }
        case EOS_TYPE_ANEOS: {
            if (rho <= 0.0) return limit;
            const TableCell c = aneos_locate(M, rho, e);
            const double cs = c.ideal_gas ? sqrt(M.aneos_gamma * (M.aneos_gamma - 1.0) * e) : aneos_bilinear(M, c_aneos.cs, c);
            return cs < limit ? limit : cs;
        }
#if PALPHA_POROSITY
        case EOS_TYPE_JUTZI_MURNAGHAN:
            if (M.pj_alpha_0 > 1.0) return M.cs_solid + (M.cs_porous - M.cs_solid) * (alpha_jutzi - 1.0) / (M.pj_alpha_0 - 1.0);
            return M.cs_solid;
        case EOS_TYPE_JUTZI_ANEOS: {
            if (rho <= 0.0) return limit;
            const TableCell c = aneos_locate(M, rho, e);
            double cs = c.ideal_gas ? sqrt(M.aneos_gamma * (M.aneos_gamma - 1.0) * e) : aneos_bilinear(M, c_aneos.cs, c);
            if (cs > M.cs_porous) cs = cs + (M.cs_porous - cs) * (alpha_jutzi - 1.0) / (M.pj_alpha_0 - 1.0);
            return cs < limit ? limit : cs;
        }
        case EOS_TYPE_JUTZI: {
            /* evaluated with the bulk density rho, not alpha*rho (src/soundspeed.cu:208-209) */
            const double cs_sq = tillotson_cs2(M, rho, e, p_old);
            if (cs_sq > M.cs_porous * M.cs_porous) {
                double cs = sqrt(cs_sq);
                if (M.pj_alpha_0 > 1.0) cs = cs + (M.cs_porous - cs) * (alpha_jutzi - 1.0) / (M.pj_alpha_0 - 1.0);
                else cs = M.cs_solid;
                return cs < limit ? limit : cs;
            }
            return cs_sq < limit * limit ? limit : sqrt(cs_sq);
        }
#endif
        default: return cs_in; /* constant sound speed set by initializeSoundspeed */
    }
}

struct PorousOut {
    double dalphadp, dalphadrho, f, delpdelrho, delpdele, alpha;
};

__device__ inline double eos_pressure(const MatParams &M, double rho, double e, double cs, double alpha_in, PorousOut &po)
{
    (void)alpha_in; (void)po;
    switch (M.eos) {
        case EOS_TYPE_POLYTROPIC_GAS: return M.poly_K * pow(rho, M.poly_gamma);
        case EOS_TYPE_IDEAL_GAS: return (M.poly_gamma - 1.0) * rho * e;
        case EOS_TYPE_ISOTHERMAL_GAS: return cs * cs * rho;
        case EOS_TYPE_MURNAGHAN: {
            const double eta = rho / M.rho0;
            return eta < M.rho_limit ? 0.0 : (M.bulk / M.n) * (pow(eta, M.n) - 1.0);
        }
        case EOS_TYPE_TILLOTSON: {
            double d1, d2;
            double pr = tillotson_p(M, rho, e, false, d1, d2);
            if (e > 1e2 * M.till_Ecv && rho / M.till_rho0 < 1.0) pr = (M.poly_gamma - 1.0) * rho * e;
            return pr;
        }
        case EOS_TYPE_ANEOS: {
            if (rho <= 0.0) return 0.0;
            const TableCell c = aneos_locate(M, rho, e);
            return c.ideal_gas ? (M.aneos_gamma - 1.0) * rho * e : aneos_bilinear(M, c_aneos.p, c);
        }
#if PALPHA_POROSITY
        case EOS_TYPE_JUTZI:
        case EOS_TYPE_JUTZI_ANEOS:
        case EOS_TYPE_JUTZI_MURNAGHAN: {
            /* p = p_solid(alpha rho, e) / alpha with the crush curve alpha(p), src/pressure.cu:204-449 */
            const double al = alpha_in;
            double psolid;
            if (M.eos == EOS_TYPE_JUTZI) {
                psolid = tillotson_p(M, rho * al, e, true, po.delpdele, po.delpdelrho);
            } else if (M.eos == EOS_TYPE_JUTZI_ANEOS) {
                /* tabulated matrix pressure with its slopes (src/pressure.cu:313-363; in-table branch of
                 * bilinear_interpolation_from_linearized_plus_derivatives, src/aneos.cu:509-531 -- the call site clamps
                 * the cell indices first, so a point outside the table is extrapolated with the edge cell's slopes) */
                if (rho <= 0.0) {
                    psolid = 0.0; po.delpdelrho = 0.0; po.delpdele = 0.0;
                } else {
                    const TableCell c = aneos_locate(M, al * rho, e);
                    if (c.ideal_gas) {
                        psolid = (M.aneos_gamma - 1.0) * rho * al * e;
                        po.delpdelrho = (M.aneos_gamma - 1.0) * e;
                        po.delpdele = (M.aneos_gamma - 1.0) * rho * al;
                    } else {
                        const double *t = c_aneos.p + M.aneos_matrix_id;
                        const double *rt = c_aneos.rho + M.aneos_rho_id, *et = c_aneos.e + M.aneos_e_id;
                        const int ne = M.aneos_n_e;
                        const double dxg = rt[c.ix + 1] - rt[c.ix], dyg = et[c.iy + 1] - et[c.iy];
                        const double delta_x = c.nx * dxg, delta_y = c.ny * dyg;
                        const double k_a = (t[(c.ix + 1) * ne + c.iy] - t[c.ix * ne + c.iy]) / dxg;
                        const double k_b = (t[(c.ix + 1) * ne + c.iy + 1] - t[c.ix * ne + c.iy + 1]) / dxg;
                        const double a2 = t[c.ix * ne + c.iy] + delta_x * k_a, b2 = t[c.ix * ne + c.iy + 1] + delta_x * k_b;
                        po.delpdelrho = k_a + delta_y * (k_b - k_a) / dyg;
                        po.delpdele = (b2 - a2) / dyg;
                        psolid = a2 + delta_y * po.delpdele;
                    }
                }
            } else {
                const double eta = rho * al / M.rho0;
                po.delpdele = 0.0;
                if (eta < M.rho_limit) {
                    psolid = 0.0;
                    po.delpdelrho = 0.0;
                } else {
                    psolid = M.bulk / M.n * (pow(eta, M.n) - 1.0);
                    po.delpdelrho = M.bulk / M.rho0 * pow(eta, M.n - 1.0);
                }
            }
            const double pr = psolid / al;
            const double p_e = M.pj_p_elastic, p_t = M.pj_p_transition, p_s = M.pj_p_compacted, a0 = M.pj_alpha_0;
            double dadp = 0.0;
            if (M.crushcurve_style == 0) {
                if (pr > p_e && pr < p_s) dadp = -2.0 * (a0 - 1.0) * (p_s - pr) / sq(p_s - p_e);
            } else if (M.crushcurve_style == 1) {
                const double k = (a0 - 1.0) / (M.pj_alpha_e - 1.0);
                if (pr > p_e && pr < p_t)
                    dadp = -k * (M.pj_alpha_e - M.pj_alpha_t) * M.pj_n1 * (pow(p_t - pr, M.pj_n1 - 1.0) / pow(p_t - p_e, M.pj_n1)) -
                           k * (M.pj_alpha_t - 1.0) * M.pj_n2 * (pow(p_s - pr, M.pj_n2 - 1.0) / pow(p_s - p_e, M.pj_n2));
                else if (pr >= p_t && pr < p_s)
                    dadp = -k * (M.pj_alpha_t - 1.0) * M.pj_n2 * (pow(p_s - pr, M.pj_n2 - 1.0) / pow(p_s - p_e, M.pj_n2));
            } else if (M.crushcurve_style == 2) {
                /* experimental crush curve of Blum et al. 2023 with the reference's built-in constants
                 * (src/pressure.cu:386-410): alpha = (P0/p)^(1/x + b p/x) + 1/phi_max */
                const double P0 = 0.044e6, x = 8.915, bb = 7e-4 * 1e-6;
                if (pr > 1.0) {
                    const double ex = 1.0 / x + bb / x * pr;
                    dadp = pow(P0 / pr, ex) * (bb / x * log(P0 / pr) - P0 * ex / (P0 * pr));
                }
            } else if (M.crushcurve_style == 3) {
                /* Malamud 2023 (src/pressure.cu:411-423) */
                if (pr > 1e2) dadp = -0.19341714781149988 / (-pr * sq(0.084 * log(pr) - 0.14736544595161893));
            } else if (M.crushcurve_style == 4) {
                /* Malamud 2023, blue curve of their figure 2c: VFF = 0.41 p^0.09, p in MPa (src/pressure.cu:424-440) */
                const double p_el = 1e6 * pow(1.0 / (0.41 * a0), 1.0 / 0.09);
                if (pr > p_el) dadp = -0.7611296717250695 * pow(pr, -1.09);
            }
            po.dalphadp = dadp;
            po.dalphadrho = ((pr / (rho * rho) * po.delpdele + al * po.delpdelrho) * dadp) / (al + dadp * (pr - rho * po.delpdelrho));
            po.f = 1.0 + po.dalphadrho * rho / al;
            po.alpha = al;
            if (al <= 1.0) {
                po.f = 1.0;
                po.alpha = 1.0;
                po.dalphadp = 0.0;
                po.dalphadrho = 0.0;
            }
            return pr;
        }
#endif
        default: return 0.0;
    }
}

__global__ void __launch_bounds__(128)
k_pointwise(Sorted s, b200sph_view v, const double *rho_sorted, int use_rho_sorted)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= s.n) return;
    const int i = s.perm[k];
    const int matId = s.mat[k];
    const b200sph_particle_arrays &p = v.p;
    const b200sph_particle_arrays &pr = v.p_rhs;
    (void)pr;   /* only the solid switch sets write p_rhs scratch here */
    if (s.abort && *s.abort) return;
    const double m = s.vel4[k].w;
    double rho = (use_rho_sorted && !(s.halo_sums_external && i >= s.n_owned)) ? rho_sorted[k] : p.rho[i];

    if (matId < 0) { /* deactivated particle: never a neighbour, keep the records finite */
        st_rec(&s.gas4[k], Rec4{0.0, 0.0, rho, 0.0});
        return;
    }
    const MatParams &M = c_mat[matId];
    const double e = p.e ? p.e[i] : 0.0;
#if PALPHA_POROSITY
    const double alpha_in = p.alpha_jutzi[i];
#else
    const double alpha_in = 1.0;
#endif
    /* sound speed first, with the pressure left by the previous call (src/rhs.cu:398 before :458) */
    const double cs = eos_soundspeed(M, rho, e, p.p[i], alpha_in, p.cs[i]);
    p.cs[i] = cs;
    /* A material whose eos.type is IGNORE: calculatePressure and damageLimit skip its particles
     * (src/pressure.cu:41, src/damage.cu:41) -- p, d and damage_total stay what they were -- while symmetrizeStress,
     * plasticityModel and set_stress_tensor have no such test and run on them with the stored p and damage_total
     * (src/timeintegration.cu:116-130, src/plasticity.cu:117-388, src/stress.cu:50-157).  Their pair sums are skipped. */
    const bool eos_ignored = (M.eos == EOS_TYPE_IGNORE);
#if !SOLID
    if (eos_ignored) {
        st_rec(&s.gas4[k], Rec4{0.0, cs, rho, m / rho});
        return;
    }
#endif
    PorousOut po;
    double pres = eos_ignored ? p.p[i] : eos_pressure(M, rho, e, cs, alpha_in, po);
#if PALPHA_POROSITY
    if (eos_ignored) {
    } else if (M.eos == EOS_TYPE_JUTZI || M.eos == EOS_TYPE_JUTZI_MURNAGHAN || M.eos == EOS_TYPE_JUTZI_ANEOS) {
        p.dalphadp[i] = po.dalphadp;
        p.dalphadrho[i] = po.dalphadrho;
        p.f[i] = po.f;
        p.delpdelrho[i] = po.delpdelrho;
        p.delpdele[i] = po.delpdele;
        if (alpha_in <= 1.0) p.alpha_jutzi[i] = 1.0;
    } else {
        p.alpha_jutzi_old[i] = alpha_in;
    }
#endif
#if REAL_HYDRO
    if (pres < 0.0 && !eos_ignored) pres = 0.0;
#endif

#if SOLID
    double S[DIM][DIM];
#pragma unroll
    for (int a = 0; a < DIM; a++)
#pragma unroll
        for (int b = 0; b < DIM; b++) S[a][b] = p.S[(size_t)i * DD + a * DIM + b];
#if FRAGMENTATION || B200_PLASTICITY
    /* symmetrizeStress, src/timeintegration.cu:116-130 */
#pragma unroll
    for (int a = 0; a < DIM; a++)
#pragma unroll
        for (int b = 0; b < a; b++) {
            const double val = 0.5 * (S[a][b] + S[b][a]);
            S[a][b] = val;
            S[b][a] = val;
        }
#endif
#if FRAGMENTATION
    /* damageLimit, src/damage.cu:33-82 */
    double damage;
    if (eos_ignored) {
        damage = p.damage_total[i];
        if (damage > 1.0) damage = 1.0;
        if (damage < 0.0) damage = 0.0;
    } else {
        double dmg = p.d[i], dmg_max = 1.0;
        const int nof = p.numFlaws[i], noaf = p.numActiveFlaws[i];
        if (dmg < 0.0) dmg = 0.0;
        if (nof > 0) dmg_max = pow((double)noaf / (double)nof, 1.0 / DIM);
        if (dmg > dmg_max) dmg = dmg_max;
        p.d[i] = dmg;
#if PALPHA_POROSITY
        double dpor = p.damage_porjutzi[i];
        if (dpor > 1.0) { dpor = 1.0; p.damage_porjutzi[i] = 1.0; }
        else if (dpor < 0.0) { dpor = 0.0; p.damage_porjutzi[i] = 0.0; }
        damage = pow(dmg, (double)DIM) + pow(dpor, (double)DIM);
        if (damage > 1.0) damage = 1.0;
#else
        damage = pow(dmg, (double)DIM);
#endif
        p.damage_total[i] = damage;
        if (damage > 1.0) damage = 1.0;
        if (damage < 0.0) damage = 0.0;
    }
#endif
#if B200_PLASTICITY
    /* plasticityModel, src/plasticity.cu:117-388 */
    {
        double J2 = 0.0, mises_f = 1.0;
#pragma unroll
        for (int a = 0; a < DIM; a++)
#pragma unroll
            for (int b = 0; b < DIM; b++) J2 = fma(S[a][b], S[a][b], J2);
        J2 *= 0.5;
#if COLLINS_PLASTICITY
        const double y_0 = M.cohesion, y_M = M.yield_stress, mu_i = M.friction;
        double y_i = y_0, y;
        if (pres > 0.0) y_i += mu_i * pres / (1.0 + mu_i * pres / (y_M - y_0));
#if COLLINS_PLASTICITY_INCLUDE_MELT_ENERGY
        if (e >= M.melt_energy) y_i = 0.0;
        else if (e > 0.0) y_i *= (1.0 - e / M.melt_energy);
#endif
#if FRAGMENTATION
        {
            const double y_0_d = M.cohesion_damaged;
            double y_d;
            if (pres > 0.0) y_d = y_0_d + M.friction_damaged * pres;
            else if (pres > -y_0_d) y_d = y_0_d + pres;
            else y_d = 0.0;
            if (y_d < 0.0) y_d = 0.0;
            y = (1.0 - damage) * y_i + damage * y_d;
            if (y > y_i) y = y_i;
            /* cap on negative pressure release by damage (applied once, here) */
            if (pres < -y_0_d) pres = ((1.0 - damage) * pres > -y_0_d) ? -y_0_d : (1.0 - damage) * pres;
        }
#else
        y = y_i;
#endif
        if (J2 > 0.0) mises_f = y / sqrt(J2);
#else /* VON_MISES_PLASTICITY */
        if (J2 > 0.0) mises_f = M.yield_stress * M.yield_stress / (3.0 * J2);
#endif
        mises_f = fmin(mises_f, 1.0);
        if (mises_f < 0.0) mises_f = 0.0;
        pr.plastic_f[i] = mises_f;
#pragma unroll
        for (int a = 0; a < DIM; a++)
#pragma unroll
            for (int b = 0; b < DIM; b++) S[a][b] *= mises_f;
    }
#endif
#if FRAGMENTATION || B200_PLASTICITY
#pragma unroll
    for (int a = 0; a < DIM; a++)
#pragma unroll
        for (int b = 0; b < DIM; b++) p.S[(size_t)i * DD + a * DIM + b] = S[a][b];
#endif
    /* set_stress_tensor, src/stress.cu:50-157 */
    double ptmp = pres;
#if !COLLINS_PLASTICITY && FRAGMENTATION
    if (pres < 0.0) ptmp = (1.0 - damage) * pres;
#endif
    const double irho2 = 1.0 / (rho * rho);
    double ten[4 * TEN_RECS];
#pragma unroll
    for (int c = 0; c < 4 * TEN_RECS; c++) ten[c] = 0.0;
#if ARTIFICIAL_STRESS
    double sigma[DIM][DIM];   /* kept in registers for the eigen-decomposition below */
#endif
#pragma unroll
    for (int a = 0; a < DIM; a++)
#pragma unroll
        for (int b = 0; b < DIM; b++) {
#if FRAGMENTATION && DAMAGE_ACTS_ON_S
            double sg = (1.0 - damage) * S[a][b];
#else
            double sg = S[a][b];
#endif
            if (a == b) sg -= ptmp;
#if ARTIFICIAL_STRESS
            sigma[a][b] = sg;
#endif
            pr.sigma[(size_t)i * DD + a * DIM + b] = sg;
            ten[ten_sig(a, b)] = sg * irho2;
        }
#if ARTIFICIAL_STRESS
    /* compute_artificial_stress, src/artificial_stress.cu:34-100: R = -eps * sigma_+ in principal axes */
    {
        double ev[DIM], V[DIM][DIM];
        sym_eigen(sigma, ev, V);
#pragma unroll
        for (int a = 0; a < DIM; a++)
#pragma unroll
            for (int b = 0; b < DIM; b++) {
                double r = 0.0;
#pragma unroll
                for (int c = 0; c < DIM; c++) {
                    const double rc = ev[c] > 0.0 ? -M.epsilon_stress * ev[c] : 0.0;
                    /* R = V^T diag V as the reference multiplies (rotation^T * (diag * rotation)) */
                    r += V[c][a] * rc * V[c][b];
                }
                pr.R[(size_t)i * DD + a * DIM + b] = r;
                if (a <= b) ten[ten_r(a, b)] = r * irho2;
            }
    }
#endif
    p.p[i] = pres;
    st_rec(&s.gas4[k], Rec4{irho2, cs, rho, m / rho});
#pragma unroll
    for (int r = 0; r < TEN_RECS; r++)
        st_rec(&s.ten[(size_t)k * TEN_RECS + r], Rec4{ten[4 * r], ten[4 * r + 1], ten[4 * r + 2], ten[4 * r + 3]});
#else /* HYDRO */
    p.p[i] = pres;
    st_rec(&s.gas4[k], Rec4{pres / (rho * rho), cs, rho, m / rho});
#endif
}

#if TENSORIAL_CORRECTION
/* ------------------------------------------------------------------ k_correction */
template <int MODE>
__global__ void __launch_bounds__(128)
k_correction(Sorted s, b200sph_view v, int n_targets, int *flags)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_targets) return;
    const int k = t;
    const int i = s.perm[k];
    const Rec4 pi = ld_rec(&s.pos4[k]);
    const int nslots = s.noi[t];
    if (s.halo_sums_external && i >= s.n_owned) return;   /* the owner sends this copy's matrix; k_import_correction stores it */
    double C[DIM][DIM];
#pragma unroll
    for (int a = 0; a < DIM; a++)
#pragma unroll
        for (int b = 0; b < DIM; b++) C[a][b] = (a == b) ? 1.0 : 0.0;
    if (!mat_ignored(s.mat[k])) {
        const double hinv = 1.0 / pi.w;   /* h_i, not the pair mean (src/kernel.cu:637) */
        const double h2_i = __dmul_rn(pi.w, pi.w);
        double A[DIM][DIM];
#pragma unroll
        for (int a = 0; a < DIM; a++)
#pragma unroll
            for (int b = 0; b < DIM; b++) A[a][b] = 0.0;
        int cnt = 0;
        int j_next = 0, j_next2 = 0;
        Rec4 pj_next = pi;
        double vol_next = 0.0;
        if (nslots > 0) {
            j_next = s.nbr[NBR_SLOT(t, 0)];
            j_next2 = s.nbr[NBR_SLOT(t, min(1, nslots - 1))];
            pj_next = ld_rec(&s.pos4[j_next]);
            vol_next = s.gas4[j_next].w;
        }
PAIR_UNROLL
        for (int q = 0; q < nslots; q++) {
            const int j = j_next;
            const Rec4 pj = pj_next;
            const double vol_j = vol_next;   /* m_j / rho_j */
            j_next = j_next2;
            j_next2 = s.nbr[NBR_SLOT(t, min(q + 2, nslots - 1))];
            pj_next = ld_rec(&s.pos4[j_next]);
            vol_next = s.gas4[j_next].w;
            double dr[3], W, g;
            const double r2 = pair_d2(pi, pj, dr[0], dr[1], dr[2]);
            if (MODE != LIST_EXACT && (j == k || !pair_is_neighbour(r2, h2_i, pj))) continue;
            if (MODE == LIST_VALIDATE) {
                if (cnt != q) s.nbr[NBR_SLOT(t, cnt)] = j;
                cnt++;
            }
            if (s.any_eos_ignore && mat_ignored(s.mat[j])) continue;
#if AVERAGE_KERNELS
            cubic_spline(r2, hinv, W, g);
            if (pj.w != pi.w) {
                double Wj, gj;
                cubic_spline(r2, 1.0 / pj.w, Wj, gj);
                g = 0.5 * (g + gj);
            }
#else
            cubic_spline(r2, hinv, W, g);
#endif
            const double w = vol_j * g;   /* (m_j/rho_j) * dW/dr / r */
            /* A is symmetric: accumulate the upper triangle only */
#pragma unroll
            for (int a = 0; a < DIM; a++) {
                const double wa = -w * dr[a];
#pragma unroll
                for (int b = a; b < DIM; b++) A[a][b] = fma(wa, dr[b], A[a][b]);
            }
        }
        if (MODE == LIST_VALIDATE) finish_validate(s, t, k, cnt, flags);
#pragma unroll
        for (int a = 0; a < DIM; a++)
#pragma unroll
            for (int b = 0; b < a; b++) A[a][b] = A[b][a];
        sym_pinv(A, C);
#if DIM == 2
        const double det = C[0][0] * C[1][1] - C[0][1] * C[1][0];
#else
        const double det = C[0][0] * (C[1][1] * C[2][2] - C[1][2] * C[2][1]) - C[0][1] * (C[1][0] * C[2][2] - C[1][2] * C[2][0]) +
                           C[0][2] * (C[1][0] * C[2][1] - C[1][1] * C[2][0]);
#endif
        double max_entry = 0.0;
#pragma unroll
        for (int a = 0; a < DIM; a++)
#pragma unroll
            for (int b = 0; b < DIM; b++) max_entry = fmax(max_entry, fabs(C[a][b]));
        if (fabs(det) < 0.2 || fabs(det) > 5.0 || max_entry > 5.0) {
#pragma unroll
            for (int a = 0; a < DIM; a++)
#pragma unroll
                for (int b = 0; b < DIM; b++) C[a][b] = (a == b) ? 1.0 : 0.0;
        }
    } else if (MODE == LIST_VALIDATE) {
        finish_validate(s, t, k, validate_only(s, t, k, pi, nslots), flags);
    }
    double *ten = reinterpret_cast<double *>(s.ten + (size_t)k * TEN_RECS);
#pragma unroll
    for (int a = 0; a < DIM; a++)
#pragma unroll
        for (int b = 0; b < DIM; b++) {
            if (a <= b) ten[ten_c(a, b)] = C[a][b];
            v.p_rhs.tensorialCorrectionMatrix[(size_t)i * DD + a * DIM + b] = C[a][b];
        }
}
#endif

#if TENSORIAL_CORRECTION
/* Multi-GPU, neighbour-sum exchange: the correction matrices of the halo copies arrived in the caller's rows
 * (tensorialCorrectionMatrix, row-major DIM x DIM); the force loop reads them from the packed sorted records. */
__global__ void k_import_correction(Sorted s, b200sph_view v)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= s.n) return;
    const int i = s.perm[k];
    if (i < s.n_owned) return;
    double *ten = reinterpret_cast<double *>(s.ten + (size_t)k * TEN_RECS);
    const double *C = v.p_rhs.tensorialCorrectionMatrix + (size_t)i * DD;
#pragma unroll
    for (int a = 0; a < DIM; a++)
#pragma unroll
        for (int b = a; b < DIM; b++) ten[ten_c(a, b)] = C[a * DIM + b];
}
#endif

/* ------------------------------------------------------------------ k_forces */
__device__ __forceinline__ double int_power(double x, int n)
{
    double r = 1.0, b = x;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (n & (1 << k)) r *= b;
        b *= b;
    }
    return r;
}

/* Register budget of the force kernel, measured on a B200 (gpurun_out v4, round 1): the 2-D solid loop gains
 * 11 % from a 128-register cap (4 blocks of 128 threads per SM), the 3-D solid loop loses 50 % to spills. */
#ifndef B200_FORCES_MIN_BLOCKS
#if SOLID && DIM == 2
#define B200_FORCES_MIN_BLOCKS 4
#else
#define B200_FORCES_MIN_BLOCKS 1
#endif
#endif
#if B200_FORCES_MIN_BLOCKS > 1
#define FORCES_BOUNDS __launch_bounds__(128, B200_FORCES_MIN_BLOCKS)
#else
#define FORCES_BOUNDS __launch_bounds__(128)   /* not (128, 1): that form makes ptxas take 190 registers for the 3-D solid loop */
#endif
template <int MODE>
__global__ void FORCES_BOUNDS
k_forces(Sorted s, b200sph_view v, int n_targets, int *flags)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_targets) return;
    const int k = t;
    const int i = s.perm[k];
    const int matId = s.mat[k];
    const b200sph_particle_arrays &p = v.p;
    const b200sph_particle_arrays &pr = v.p_rhs;
    (void)pr;
    if (i >= s.n_owned) return;   /* halo copy: its owner computes the rates */
    if (s.abort && *s.abort) return;
    const Rec4 pi = ld_rec(&s.pos4[k]);
    const Rec4 vi = ld_rec(&s.vel4[k]);
    const int nslots = s.noi[t];
    int noi = nslots;

    const bool active = !(matId == BOUNDARY_PARTICLE_ID || mat_ignored(matId)) && i < v.n_real;
    double acc[3] = {0.0, 0.0, 0.0}, drhodt = 0.0, dedt = 0.0, dhdt = 0.0, muijmax = 0.0;
    (void)dedt; (void)dhdt; (void)muijmax;
#if SOLID
    double vgrad[DIM][DIM];
#pragma unroll
    for (int a = 0; a < DIM; a++)
#pragma unroll
        for (int b = 0; b < DIM; b++) vgrad[a][b] = 0.0;
    double sig_i[DIM][DIM];
#if TENSORIAL_CORRECTION
    double Ci[DIM][DIM];
#endif
#if ARTIFICIAL_STRESS
    double Ri[DIM][DIM];
#endif
#endif

    if (active && nslots > 0) {
        const MatParams &M = c_mat[matId];
        const Rec4 gi = ld_rec(&s.gas4[k]);
        const double h2_i = __dmul_rn(pi.w, pi.w);
        int cnt = 0;
        (void)h2_i; (void)cnt;
        const double rho_i = gi.z;
        (void)rho_i;
#if ARTIFICIAL_VISCOSITY
        const double av_alpha = M.av_alpha, av_beta = M.av_beta;
#endif
#if SOLID
        {
            double ti[4 * TEN_RECS];
#pragma unroll
            for (int r = 0; r < TEN_RECS; r++) {
                const Rec4 t = ld_rec(&s.ten[(size_t)k * TEN_RECS + r]);
                ti[4 * r] = t.x; ti[4 * r + 1] = t.y; ti[4 * r + 2] = t.z; ti[4 * r + 3] = t.w;
            }
#pragma unroll
            for (int a = 0; a < DIM; a++)
#pragma unroll
                for (int b = 0; b < DIM; b++) {
                    sig_i[a][b] = ti[ten_sig(a, b)];
#if TENSORIAL_CORRECTION
                    Ci[a][b] = ti[ten_c(a, b)];
#endif
#if ARTIFICIAL_STRESS
                    Ri[a][b] = ti[ten_r(a, b)];
#endif
                }
        }
        const double m_over_rho_i_unit = 1.0 / rho_i;   /* strain rate uses m_j / rho_i (src/internal_forces.cu:476) */
#endif
#if !(VARIABLE_SML || INTEGRATE_SML)
        const double hinv_fixed = 1.0 / pi.w;
#endif
#if ARTIFICIAL_STRESS
        const double w_ref_dist = M.mean_particle_distance;
        const double w_ref_same_h = cubic_spline_w(w_ref_dist, 1.0 / pi.w);
        /* Monaghan's exponent is a small integer in every shipped material.cfg (n = 4): repeated multiplication
         * instead of pow(); any other value takes the general path */
        const int art_int_exp = (M.exponent_tensor >= 1.0 && M.exponent_tensor <= 8.0 && M.exponent_tensor == floor(M.exponent_tensor))
                                    ? (int)M.exponent_tensor : 0;
#endif
        int j_next = s.nbr[NBR_SLOT(t, 0)];
        int j_next2 = s.nbr[NBR_SLOT(t, min(1, nslots - 1))];
        PairRecs nxt;
        load_force_recs(s, j_next, nxt);
PAIR_UNROLL
        for (int q = 0; q < nslots; q++) {
            const int j = j_next;
            PairRecs cur = nxt;
            load_tensor_recs(s, j, cur);
            j_next = j_next2;
            j_next2 = s.nbr[NBR_SLOT(t, min(q + 2, nslots - 1))];
            load_force_recs(s, j_next, nxt);   /* past the end this re-reads the last neighbour (harmless) */
            const Rec4 &pj = cur.p;
            double dr[3], dv[3], W, g;
            const double r2 = pair_d2(pi, pj, dr[0], dr[1], dr[2]);
            if (MODE != LIST_EXACT && (j == k || !pair_is_neighbour(r2, h2_i, pj))) continue;
            if (MODE == LIST_VALIDATE) {
                if (cnt != q) s.nbr[NBR_SLOT(t, cnt)] = j;
                cnt++;
            }
            if (s.any_eos_ignore && mat_ignored(s.mat[j])) continue;
            const Rec4 &vj = cur.v;
            const Rec4 &gj = cur.g;
#if SOLID
            double tj[4 * TEN_RECS];
#pragma unroll
            for (int r = 0; r < TEN_RECS; r++) {
                tj[4 * r] = cur.t[r].x; tj[4 * r + 1] = cur.t[r].y; tj[4 * r + 2] = cur.t[r].z; tj[4 * r + 3] = cur.t[r].w;
            }
#endif
            dv[0] = vi.x - vj.x; dv[1] = vi.y - vj.y; dv[2] = vi.z - vj.z;
#if VARIABLE_SML || INTEGRATE_SML
            const double hbar = 0.5 * (pi.w + pj.w);
            const double hinv = 1.0 / hbar;
#else
            /* fixed h; with AVERAGE_KERNELS and no SHEPARD_CORRECTION the reference also uses h_i
             * only (the averaging lines are compiled out, src/internal_forces.cu:328-346) */
            const double hinv = hinv_fixed;
#endif
            cubic_spline(r2, hinv, W, g);
            double gw[DIM];   /* plain kernel gradient */
#pragma unroll
            for (int a = 0; a < DIM; a++) gw[a] = g * dr[a];
            const double mj = vj.w;

#if TENSORIAL_CORRECTION
            double gci[DIM], gcj[DIM], gsym[DIM];
#pragma unroll
            for (int a = 0; a < DIM; a++) {
                double tci = 0.0, tcj = 0.0;
#pragma unroll
                for (int b = 0; b < DIM; b++) {
                    tci = fma(Ci[a][b], gw[b], tci);
                    tcj = fma(tj[ten_c(a, b)], gw[b], tcj);
                }
                gci[a] = tci;
                gcj[a] = tcj;
                gsym[a] = 0.5 * (tci + tcj);
            }
#else
            const double *gsym = gw;
#endif
            double vvnablaW = 0.0;
#pragma unroll
            for (int a = 0; a < DIM; a++) vvnablaW = fma(dv[a], gsym[a], vvnablaW);

#if SOLID
            /* strain rate and rotation rate, edot_ab = 1/2 (d_b v_a + d_a v_b) */
            {
                /* accumulate the velocity gradient L_ab = sum w dv_a grad_b; edot = L + L^T and
                 * rdot = L - L^T are formed once after the loop (9 accumulators instead of 18) */
                const double w = -0.5 * mj * m_over_rho_i_unit;
#pragma unroll
                for (int a = 0; a < DIM; a++) {
                    const double wa = w * dv[a];
#pragma unroll
                    for (int b = 0; b < DIM; b++) vgrad[a][b] = fma(wa, gsym[b], vgrad[a][b]);
                }
            }
#endif
            double pij = 0.0;
#if ARTIFICIAL_VISCOSITY
            {
                double vr = 0.0;
#pragma unroll
                for (int a = 0; a < DIM; a++) vr = fma(dv[a], dr[a], vr);
                if (vr < 0.0) {
                    const double csbar = 0.5 * (gi.y + gj.y);
                    const double smooth = 0.5 * (pi.w + pj.w);
                    /* mu = h vr / (r^2 + 0.01 h^2) and Pi = (beta mu - alpha c) mu / rho_bar with ONE division:
                     * 1/(A B) gives 1/A = B/(A B) and 1/B = A/(A B) (rounding-level difference to two divisions) */
                    const double den = fma(smooth * smooth, 1e-2, r2);
                    const double rhobar = 0.5 * (gi.z + gj.z);
                    const double inv = 1.0 / (den * rhobar);
                    const double mu = smooth * vr * (inv * rhobar);
                    muijmax = fmax(muijmax, mu);
                    pij = (av_beta * mu - av_alpha * csbar) * mu * (inv * den);
                }
            }
#endif
#if SOLID
            {
                double aj[DIM];
#pragma unroll
                for (int a = 0; a < DIM; a++) {
                    double t = 0.0;
#pragma unroll
                    for (int b = 0; b < DIM; b++) {
#if TENSORIAL_CORRECTION
                        t = fma(tj[ten_sig(a, b)], gcj[b], t);
                        t = fma(sig_i[a][b], gci[b], t);
#else
                        t = fma(tj[ten_sig(a, b)] + sig_i[a][b], gw[b], t);
#endif
                    }
                    aj[a] = mj * t;
                }
#if ARTIFICIAL_STRESS
                /* Monaghan (2000) tensile-instability fix: (W(r)/W(dp))^n * (R_i/rho_i^2 + R_j/rho_j^2) */
                {
                    const double hb = 0.5 * (pi.w + pj.w), hbinv = 1.0 / hb;
                    const double r = sqrt(r2);
                    /* W(mean particle distance) only depends on h_bar: hoisted for partners with h_j = h_i */
                    const double w_ref = (pj.w == pi.w) ? w_ref_same_h : cubic_spline_w(w_ref_dist, hbinv);
                    const double ratio = cubic_spline_w(r, hbinv) / w_ref;
                    const double artf = (art_int_exp > 0) ? int_power(ratio, art_int_exp) : pow(ratio, M.exponent_tensor);
#pragma unroll
                    for (int a = 0; a < DIM; a++) {
                        double t = 0.0;
#pragma unroll
                        for (int b = 0; b < DIM; b++) {
#if TENSORIAL_CORRECTION
                            t = fma(Ri[a][b], gci[b], t);
                            t = fma(tj[ten_r(a, b)], gcj[b], t);
#else
                            t = fma(Ri[a][b] + tj[ten_r(a, b)], gw[b], t);
#endif
                        }
                        const double art = mj * artf * t;
                        acc[a] += art;
#if INTEGRATE_ENERGY && TENSORIAL_CORRECTION
                        dedt = fma(-0.5 * art, dv[a], dedt);
#endif
                    }
                }
#endif
#pragma unroll
                for (int a = 0; a < DIM; a++) acc[a] += aj[a];
#if INTEGRATE_ENERGY
                /* pairwise-conservative heating: 1/2 m_j (sigma_i/rho_i^2 grad_i + sigma_j/rho_j^2 grad_j) . dv */
                {
                    double t = 0.0;
#pragma unroll
                    for (int a = 0; a < DIM; a++) t = fma(aj[a], dv[a], t);
                    dedt = fma(0.5, t, dedt);
                }
#endif
            }
#else /* HYDRO */
            {
                const double w = -mj * (gi.x + gj.x);
                double t = 0.0;
#pragma unroll
                for (int a = 0; a < DIM; a++) {
                    const double aj = w * gw[a];
                    acc[a] += aj;
                    t = fma(aj, dv[a], t);
                }
#if INTEGRATE_ENERGY
                dedt = fma(-0.5, t, dedt);
#endif
            }
#endif
#if ARTIFICIAL_VISCOSITY
            {
                const double w = -mj * pij;
#pragma unroll
                for (int a = 0; a < DIM; a++) acc[a] = fma(w, gsym[a], acc[a]);
#if INTEGRATE_ENERGY
                if (!v.is_relaxation_run) dedt = fma(0.5 * mj * pij, vvnablaW, dedt);
#endif
            }
#endif
            drhodt = fma(gi.z * gj.w, vvnablaW, drhodt);            /* rho_i/rho_j m_j v.gradW */
#if INTEGRATE_SML
            dhdt = fma(-(1.0 / DIM) * pi.w * gj.w, vvnablaW, dhdt);
#endif
        }
        if (MODE == LIST_VALIDATE) noi = finish_validate(s, t, k, cnt, flags);
    } else if (MODE == LIST_VALIDATE) {
        noi = finish_validate(s, t, k, validate_only(s, t, k, pi, nslots), flags);
    }
    p.noi[i] = noi;

    /* ---------------- per-particle epilogue: everything the integrators read, in caller order */
    if (!active) {
        /* zero_all_derivatives + boundary hooks for deactivated / virtual particles */
        p.ax[i] = 0.0; p.dxdt[i] = 0.0;
#if DIM > 1
        p.ay[i] = 0.0; p.dydt[i] = 0.0;
#endif
#if DIM > 2
        p.az[i] = 0.0; p.dzdt[i] = 0.0;
#endif
        p.drhodt[i] = 0.0;
#if INTEGRATE_ENERGY
        p.dedt[i] = 0.0;
#endif
#if INTEGRATE_SML
        p.dhdt[i] = 0.0;
#endif
#if SOLID
#pragma unroll
        for (int a = 0; a < DD; a++) p.dSdt[(size_t)i * DD + a] = 0.0;
#endif
#if FRAGMENTATION
        p.dddt[i] = 0.0;
#endif
        return;
    }

    const MatParams &M = c_mat[matId];
    p.ax[i] = acc[0]; p.dxdt[i] = vi.x;
#if DIM > 1
    p.ay[i] = acc[1]; p.dydt[i] = vi.y;
#endif
#if DIM > 2
    p.az[i] = acc[2]; p.dzdt[i] = vi.z;
#endif
#if INTEGRATE_DENSITY
    if (M.density_via_kernel_sum) drhodt = 0.0;
#endif
    /* BoundaryConditionsAfterRHS floors (src/boundary.cu:311-324) act on the stored rate only */
    if (s.gas4[k].z < M.density_floor) {
        p.rho[i] = M.density_floor;
        p.drhodt[i] = 0.0;
    } else {
        p.drhodt[i] = drhodt;
    }
#if INTEGRATE_ENERGY
    p.dedt[i] = dedt;
#endif
#if INTEGRATE_SML
    p.dhdt[i] = dhdt;
#endif
#if PALPHA_POROSITY
    double dalphadt = 0.0, alpha_now = p.alpha_jutzi[i];
    if (noi > 0 && (M.eos == EOS_TYPE_JUTZI || M.eos == EOS_TYPE_JUTZI_MURNAGHAN || M.eos == EOS_TYPE_JUTZI_ANEOS)) {
        if (alpha_now <= 1.0) {
            alpha_now = 1.0;
            p.alpha_jutzi[i] = 1.0;
        } else {
            const double dadp = p.dalphadp[i], dpdr = p.delpdelrho[i];
#if INTEGRATE_ENERGY
            dalphadt = ((dedt * p.delpdele[i] + alpha_now * drhodt * dpdr) * dadp) / (alpha_now + dadp * (p.p[i] - s.gas4[k].z * dpdr));
#else
            dalphadt = ((alpha_now * drhodt * dpdr) * dadp) / (alpha_now + dadp * (p.p[i] - s.gas4[k].z * dpdr));
#endif
            if (dalphadt > 0.0) dalphadt = 0.0;
        }
    }
    p.dalphadt[i] = dalphadt;
#endif
#if SOLID
    if (noi < 1) {
#pragma unroll
        for (int a = 0; a < DD; a++) p.dSdt[(size_t)i * DD + a] = 0.0;
#if FRAGMENTATION
        p.dddt[i] = 0.0;
#if PALPHA_POROSITY
        p.ddamage_porjutzidt[i] = 0.0;
#endif
#endif
        return;
    }
    {
        const double shear = M.shear, bulk = M.bulk, young = M.young;
        (void)bulk;
        double edot[DIM][DIM], rdot[DIM][DIM];
#pragma unroll
        for (int a = 0; a < DIM; a++)
#pragma unroll
            for (int b = 0; b < DIM; b++) {
                edot[a][b] = vgrad[a][b] + vgrad[b][a];
                rdot[a][b] = vgrad[a][b] - vgrad[b][a];
            }
        double S[DIM][DIM];
#pragma unroll
        for (int a = 0; a < DIM; a++)
#pragma unroll
            for (int b = 0; b < DIM; b++) S[a][b] = p.S[(size_t)i * DD + a * DIM + b];
        double tr = 0.0;
#pragma unroll
        for (int a = 0; a < DIM; a++) tr += edot[a][a];
        const double pf = 1.0 - pr.plastic_f[i];
        double K2 = 0.0;
#pragma unroll
        for (int a = 0; a < DIM; a++)
#pragma unroll
            for (int b = 0; b < DIM; b++) {
                /* Hooke + Jaumann rotation terms (src/internal_forces.cu:1028-1050) */
                double ds = 2.0 * shear * edot[a][b];
                double ep = pf * edot[a][b];
                if (a == b) {
                    ds -= 2.0 * shear * tr / 3.0;
                    ep -= pf * tr / 3.0;
                }
#pragma unroll
                for (int c = 0; c < DIM; c++) {
                    ds = fma(S[a][c], rdot[b][c], ds);
                    ds = fma(S[b][c], rdot[a][c], ds);
                }
#if PALPHA_POROSITY && STRESS_PALPHA_POROSITY
                if (M.eos == EOS_TYPE_JUTZI || M.eos == EOS_TYPE_JUTZI_MURNAGHAN || M.eos == EOS_TYPE_JUTZI_ANEOS)
                    ds = p.f[i] / alpha_now * ds - 1.0 / (alpha_now * alpha_now) * S[a][b] * dalphadt;
#endif
                p.dSdt[(size_t)i * DD + a * DIM + b] = ds;
                K2 = fma(ep, ep, K2);
            }
        p.edotp[i] = sqrt(2.0 / 3.0 * K2);
#if ARTIFICIAL_VISCOSITY
        p.muijmax[i] = muijmax;
#endif
        /* largest principal stress -> local scalar strain (Grady-Kipp) */
        double sigma[DIM][DIM];
        const double rho2 = s.gas4[k].z * s.gas4[k].z;
#pragma unroll
        for (int a = 0; a < DIM; a++)
#pragma unroll
            for (int b = 0; b < DIM; b++) sigma[a][b] = pr.sigma[(size_t)i * DD + a * DIM + b];
        (void)rho2;
        const double tensile_max = sym_max_eigenvalue(sigma);
        double local_strain = tensile_max / young;
#if FRAGMENTATION
        {
            const double di_tensile = pow(p.d[i], (double)DIM);
            if (di_tensile < 1.0) {
                local_strain = tensile_max / ((1.0 - di_tensile) * young);
                const double c_g = 0.4 * sqrt((bulk + 4.0 * shear * (1.0 - di_tensile) / 3.0) / s.gas4[k].z);
                int n_active = 0;
                const int nf = p.numFlaws[i];
                const double *fl = pr.flaws + (size_t)i * v.max_num_flaws;
                for (int f = 0; f < nf; f++) n_active += (fl[f] < local_strain) ? 1 : 0;
                p.numActiveFlaws[i] = max(n_active, p.numActiveFlaws[i]);
                p.dddt[i] = n_active * c_g / pi.w;
            } else {
                local_strain = 0.0;
                p.numActiveFlaws[i] = p.numFlaws[i];
                p.dddt[i] = 0.0;
                p.d[i] = 1.0;
            }
#if PALPHA_POROSITY
            double ddp = 0.0;
            if (M.eos == EOS_TYPE_JUTZI || M.eos == EOS_TYPE_JUTZI_MURNAGHAN || M.eos == EOS_TYPE_JUTZI_ANEOS) {
                const double deld = 0.01, a0 = M.pj_alpha_0;
                if (a0 > 1.0)
                    ddp = -1.0 / DIM * pow(1.0 - (alpha_now - 1.0) / (a0 - 1.0) + deld, 1.0 / DIM - 1.0) /
                          (pow(1.0 + deld, 1.0 / DIM) - pow(deld, 1.0 / DIM)) * 1.0 / (a0 - 1.0) * dalphadt;
            }
            p.ddamage_porjutzidt[i] = ddp;
#endif
        }
#endif
        p.local_strain[i] = local_strain;
    }
#endif /* SOLID */
}

/* ------------------------------------------------------------------ export of neighbour lists */
__global__ void k_export_interactions(Sorted s, int *out, int max_per_row)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= s.n) return;
    const int i = s.perm[t];
    const int noi = s.noi[t];
    int *row = out + (size_t)i * max_per_row;
    for (int q = 0; q < max_per_row; q++) row[q] = (q < noi) ? s.perm[s.nbr[NBR_SLOT(t, q)]] : -1;
}

/* cold calls */
__global__ void k_pressure_only(b200sph_view v)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.n) return;
    const int matId = v.p_rhs.materialId[i];
    if (matId < 0 || c_mat[matId].eos == EOS_TYPE_IGNORE) return;
    const MatParams &M = c_mat[matId];
    const b200sph_particle_arrays &p = v.p;
#if PALPHA_POROSITY
    const double alpha_in = p.alpha_jutzi[i];
#else
    const double alpha_in = 1.0;
#endif
    PorousOut po;
    double pres = eos_pressure(M, p.rho[i], p.e ? p.e[i] : 0.0, p.cs[i], alpha_in, po);
#if PALPHA_POROSITY
    if (M.eos == EOS_TYPE_JUTZI || M.eos == EOS_TYPE_JUTZI_MURNAGHAN || M.eos == EOS_TYPE_JUTZI_ANEOS) {
        p.dalphadp[i] = po.dalphadp;
        p.dalphadrho[i] = po.dalphadrho;
        p.f[i] = po.f;
        p.delpdelrho[i] = po.delpdelrho;
        p.delpdele[i] = po.delpdele;
        if (alpha_in <= 1.0) p.alpha_jutzi[i] = 1.0;
    } else {
        p.alpha_jutzi_old[i] = alpha_in;
    }
#endif
#if REAL_HYDRO
    if (pres < 0.0) pres = 0.0;
#endif
    p.p[i] = pres;
}

__global__ void k_init_soundspeed(b200sph_view v)
{
    /* initializeSoundspeed, src/soundspeed.cu:292-322 */
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.n) return;
    const int matId = v.p_rhs.materialId[i];
    if (matId < 0) return;
    const MatParams &M = c_mat[matId];
    double *cs = v.p.cs;
    switch (M.eos) {
        case EOS_TYPE_POLYTROPIC_GAS: cs[i] = 0.0; break;
        case EOS_TYPE_ISOTHERMAL_GAS: cs[i] = 203.0; break;
        case EOS_TYPE_TILLOTSON: cs[i] = sqrt(M.bulk / M.till_rho0); break;
        case EOS_TYPE_ANEOS: cs[i] = M.aneos_bulk_cs; break;
        case EOS_TYPE_MURNAGHAN: cs[i] = sqrt(M.bulk / M.rho0); break;
        case EOS_TYPE_JUTZI: case EOS_TYPE_JUTZI_ANEOS: case EOS_TYPE_JUTZI_MURNAGHAN: cs[i] = M.cs_porous; break;
        case EOS_TYPE_REGOLITH: cs[i] = 500.0; break;
        default: break;
    }
}

#if FRAGMENTATION
__global__ void k_damage_limit(b200sph_view v)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.n) return;
    const int matId = v.p_rhs.materialId[i];
    if (mat_ignored(matId)) return;
    const b200sph_particle_arrays &p = v.p;
    double dmg = p.d[i], dmg_max = 1.0;
    const int nof = p.numFlaws[i], noaf = p.numActiveFlaws[i];
    if (dmg < 0.0) dmg = 0.0;
    if (nof > 0) dmg_max = pow((double)noaf / (double)nof, 1.0 / DIM);
    if (dmg > dmg_max) dmg = dmg_max;
    p.d[i] = dmg;
#if PALPHA_POROSITY
    double dpor = p.damage_porjutzi[i];
    if (dpor > 1.0) { dpor = 1.0; p.damage_porjutzi[i] = 1.0; }
    else if (dpor < 0.0) { dpor = 0.0; p.damage_porjutzi[i] = 0.0; }
    double tot = pow(dmg, (double)DIM) + pow(dpor, (double)DIM);
    if (tot > 1.0) tot = 1.0;
    p.damage_total[i] = tot;
#else
    p.damage_total[i] = pow(dmg, (double)DIM);
#endif
}
#endif

/* ------------------------------------------------------------------ host orchestration */
#define CU(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            snprintf(h->err, sizeof(h->err), "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return B200SPH_ERR_CUDA;                                                                 \
        }                                                                                            \
    } while (0)

static inline int blocks_for(int n, int threads) { return (n + threads - 1) / threads; }

int gravity_tree_create(b200sph_handle *h);
void gravity_tree_destroy(b200sph_handle *h);
int gravity_eval(b200sph_handle *h, const b200sph_view &v, int *launches);

static int rhs_check_view(b200sph_handle *h, const b200sph_view *view)
{
    if (!h || !view) return B200SPH_ERR_BAD_ARGUMENT;
    const b200sph_view &v = *view;
    if (v.n <= 0 || v.n > h->n_max) {
        snprintf(h->err, sizeof(h->err), "n = %d outside (0, n_max = %d]", v.n, h->n_max);
        return B200SPH_ERR_BAD_ARGUMENT;
    }
    if (!h->materials_set) {
        snprintf(h->err, sizeof(h->err), "b200sph_set_materials() has not been called");
        return B200SPH_ERR_BAD_ARGUMENT;
    }
    if (!v.p.x || !v.p.vx || !v.p.m || !v.p.h || !v.p.rho || !v.p.p || !v.p.cs || !v.p.ax || !v.p.dxdt || !v.p.drhodt ||
        !v.p.noi || !v.p_rhs.materialId) {
        snprintf(h->err, sizeof(h->err), "view is missing a mandatory array (x, vx, m, h, rho, p, cs, ax, dxdt, drhodt, noi, materialId)");
        return B200SPH_ERR_BAD_ARGUMENT;
    }
    return B200SPH_OK;
}

/* stage 0: boundary hooks, cell sort, neighbour search, kernel-sum density */
static int rhs_stage_search(b200sph_handle *h, const b200sph_view &v)
{
    cudaStream_t st = h->stream;
    Sorted &s = h->s;
    s.n = v.n;
    s.n_owned = (h->n_owned > 0 && h->n_owned < v.n) ? h->n_owned : v.n;
    s.halo_sums_external = (h->halo_sums_external && s.n_owned < v.n) ? 1 : 0;
    s.abort = h->abort_flag;
    const int n = v.n, T = 128;
    int launches = 0;

    CU(cudaEventRecord(h->ev[0], st));
    int init_flags[5] = {0x7fffffff, 0, 0, 0, 0};   /* [4]: gravity walk ran out of stack */
    CU(cudaMemcpyAsync(h->d_flags, init_flags, sizeof(init_flags), cudaMemcpyHostToDevice, st));
    {
        const int blocks = min(blocks_for(n, PREP_THREADS), h->n_sm * 4);
        double3 glo = make_double3(h->global_lo[0], h->global_lo[1], h->global_lo[2]);
        double3 ghi = make_double3(h->global_hi[0], h->global_hi[1], h->global_hi[2]);
        k_prepare<<<blocks, PREP_THREADS, 0, st>>>(v, h->block_partials, h->block_counter, h->d_domain, h->max_cells,
                                                   h->have_global_domain, glo, ghi, 1);
        launches++;
    }
    k_cell_keys<<<blocks_for(n, 256), 256, 0, st>>>(v, h->d_domain, h->keys_in, h->idx_in);
    launches++;
    /* library plumbing (not counted in kernel_launches) */
    CU(cub::DeviceRadixSort::SortPairs(h->cub_tmp, h->cub_tmp_bytes, h->keys_in, s.keys, h->idx_in, s.perm, n, 0, h->sort_bits, st));
    k_cell_start<<<blocks_for(n + 1, 256), 256, 0, st>>>(s.keys, n, h->d_domain, s.cell_start);   /* whole warps: no early exit inside */
    k_gather<<<blocks_for(n, 256), 256, 0, st>>>(v, s, h->d_domain);
    launches += 2;
    CU(cudaEventRecord(h->ev[1], st));
    /* The pointwise chain (EOS, plasticity, stress: HBM-bound) needs nothing from the search (issue-bound) unless the
     * density is a kernel sum: the two run side by side on a second stream.  Not for the host-buffer call, whose late
     * inputs and early outputs are ordered around k_pointwise on the main stream. */
    h->pointwise_early = 0;
    if (h->overlap_pointwise && !h->kernel_sum_density && !h->hook_wait_before_pointwise && !h->hook_after_pointwise) {
        if (!h->aux_stream) {
            CU(cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
        }
        CU(cudaEventRecord(h->ev_fork, st));
        CU(cudaStreamWaitEvent(h->aux_stream, h->ev_fork, 0));
        k_pointwise<<<blocks_for(n, T), T, 0, h->aux_stream>>>(s, v, h->rho_sorted, 0);
        CU(cudaEventRecord(h->ev_join, h->aux_stream));
        h->pointwise_early = 1;
        launches++;
    }
    {
        /* halo copies whose neighbour sums nobody reads skip the search (multi-GPU; boxes from b200sph_halo_set_domains);
         * with the neighbour-sum exchange no copy is searched for at all */
        const HaloDomains *hd = (h->halo && s.n_owned < n) ? ((HaloState *)h->halo)->dev : nullptr;
        const int halo_sums = (h->kernel_sum_density || TENSORIAL_CORRECTION) ? 1 : 0;
        k_neighbours<<<blocks_for(n, T), T, 0, st>>>(s, h->d_domain, n, h->d_flags, hd, halo_sums);
    }
    launches++;
    CU(cudaEventRecord(h->ev[2], st));

    /* The search leaves FP32-pre-filtered lists; the first pair loop that visits EVERY particle applies
     * the exact test and compacts them (LIST_VALIDATE), later loops trust them (LIST_EXACT).
     * kernel-sum density: always without INTEGRATE_DENSITY, else only for materials that ask for it
     * (then it does not visit every particle and only checks, LIST_CHECK). */
    h->lists_validated = 0;
    if (h->kernel_sum_density) {
#if INTEGRATE_DENSITY
        k_density<LIST_CHECK><<<blocks_for(n, T), T, h->pad_smem, st>>>(s, v, h->rho_sorted, n, h->d_flags);
#else
        k_density<LIST_VALIDATE><<<blocks_for(n, T), T, h->pad_smem, st>>>(s, v, h->rho_sorted, n, h->d_flags);
        h->lists_validated = 1;
#endif
        launches++;
    }
    CU(cudaEventRecord(h->ev[3], st));
    h->stage_launches = launches;
    return B200SPH_OK;
}

/* stage 1: the pointwise chain and the tensorial correction */
static int rhs_stage_pointwise(b200sph_handle *h, const b200sph_view &v)
{
    cudaStream_t st = h->stream;
    Sorted &s = h->s;
    const int n = v.n, T = 128;
    /* host-buffer entry point: the inputs only the pointwise chain and the pair loops read arrive on the
     * copy stream while the search runs; the state k_pointwise finalises leaves while the pair loops run */
    if (h->hook_wait_before_pointwise) CU(cudaStreamWaitEvent(st, h->hook_wait_before_pointwise, 0));
    if (h->pointwise_early) {
        CU(cudaStreamWaitEvent(st, h->ev_join, 0));   /* ran beside the search */
    } else {
        k_pointwise<<<blocks_for(n, T), T, 0, st>>>(s, v, h->rho_sorted, h->kernel_sum_density);
        h->stage_launches++;
    }
    if (h->hook_after_pointwise) h->hook_after_pointwise(h, h->hook_ctx);
    CU(cudaEventRecord(h->ev[4], st));
#if TENSORIAL_CORRECTION
    if (h->lists_validated) k_correction<LIST_EXACT><<<blocks_for(n, T), T, 0, st>>>(s, v, n, h->d_flags);
    else k_correction<LIST_VALIDATE><<<blocks_for(n, T), T, 0, st>>>(s, v, n, h->d_flags);
    h->lists_validated = 1;
    h->stage_launches++;
#endif
    CU(cudaEventRecord(h->ev[5], st));
    return B200SPH_OK;
}

/* stage 2: pair forces, gravity, end-of-call checks (the one host synchronisation of an evaluation) */
static int rhs_stage_forces(b200sph_handle *h, const b200sph_view &v, int *offender)
{
    cudaStream_t st = h->stream;
    Sorted &s = h->s;
    const int n = v.n;
    int launches = h->stage_launches;
#if TENSORIAL_CORRECTION
    if (s.halo_sums_external) {
        k_import_correction<<<blocks_for(n, 256), 256, 0, st>>>(s, v);
        launches++;
    }
#endif
    const int TF = h->forces_threads;
    const size_t forces_smem = (size_t)h->pad_smem;
    if (h->lists_validated) k_forces<LIST_EXACT><<<blocks_for(n, TF), TF, forces_smem, st>>>(s, v, n, h->d_flags);
    else k_forces<LIST_VALIDATE><<<blocks_for(n, TF), TF, forces_smem, st>>>(s, v, n, h->d_flags);
    k_list_stats<<<min(blocks_for(n, 256), h->n_sm * 4), 256, 0, st>>>(s, h->d_flags);
    launches += 2;
    CU(cudaEventRecord(h->ev[6], st));

    h->stats.gravity_recomputed = 0;
    if (v.selfgravity) {
        int rc = gravity_eval(h, v, &launches);
        if (rc) return rc;
    }
    CU(cudaEventRecord(h->ev[7], st));

    int flags[5], aborted = 0;
    CU(cudaMemcpyAsync(flags, h->d_flags, sizeof(flags), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&h->h_domain, h->d_domain, sizeof(Domain), cudaMemcpyDeviceToHost, st));
    if (h->abort_flag) CU(cudaMemcpyAsync(&aborted, h->abort_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());

    b200sph_stats &S = h->stats;
    S.kernel_launches = launches;
    S.n_cells = h->h_domain.n_cells;
    S.cell_size = h->h_domain.cell;
    S.max_noi = flags[1];
    S.total_noi = (int64_t)(((unsigned long long)(unsigned int)flags[3] << 32) | (unsigned int)flags[2]);
    cudaEventElapsedTime(&S.ms_total, h->ev[0], h->ev[7]);
    cudaEventElapsedTime(&S.ms_sort, h->ev[0], h->ev[1]);
    cudaEventElapsedTime(&S.ms_neighbours, h->ev[1], h->ev[2]);
    cudaEventElapsedTime(&S.ms_density, h->ev[2], h->ev[3]);
    cudaEventElapsedTime(&S.ms_pointwise, h->ev[3], h->ev[4]);
    cudaEventElapsedTime(&S.ms_correction, h->ev[4], h->ev[5]);
    cudaEventElapsedTime(&S.ms_forces, h->ev[5], h->ev[6]);
    cudaEventElapsedTime(&S.ms_gravity, h->ev[6], h->ev[7]);
    S.ms_scatter = 0.0f;

    if (aborted) {
        snprintf(h->err, sizeof(h->err), "evaluation abandoned: the caller's abort flag was set (stale halo plan); no state was modified");
        return B200SPH_ERR_ABORTED;
    }
    if (h->h_domain.nonfinite) {
        snprintf(h->err, sizeof(h->err), "non-finite particle coordinate or smoothing length (NaN/Inf) among the %d particles", v.n);
        return B200SPH_ERR_NONFINITE;
    }
    if (flags[4] != 0) {
        snprintf(h->err, sizeof(h->err), "gravity tree walk exceeded its traversal stack (tree deeper than the library supports)");
        return B200SPH_ERR_UNSUPPORTED;
    }
    if (flags[0] != 0x7fffffff) {
        if (offender) *offender = flags[0];
        snprintf(h->err, sizeof(h->err), "particle %d has >= MAX_NUM_INTERACTIONS = %d interaction partners (reference: assert in src/tree.cu:917)",
                 flags[0], MAX_NUM_INTERACTIONS);
        return B200SPH_ERR_TOO_MANY_INTERACTIONS;
    }
    if (offender) *offender = -1;
    return B200SPH_OK;
}

extern "C" int b200sph_rhs_eval(b200sph_handle *h, const b200sph_view *view, int *offender)
{
    int rc = rhs_check_view(h, view);
    if (rc) return rc;
    CU(cudaSetDevice(h->device));
    if ((rc = rhs_stage_search(h, *view)) != 0) return rc;
    if ((rc = rhs_stage_pointwise(h, *view)) != 0) return rc;
    return rhs_stage_forces(h, *view, offender);
}

/* The same evaluation in three stream-ordered stages for a multi-GPU host that completes the neighbour SUMS of halo
 * copies itself: after stage 0 the owners' kernel-sum densities are in view->p.rho, after stage 1
 * their correction matrices in view->p_rhs.tensorialCorrectionMatrix; the host moves those of the particles it sent
 * into the matching halo rows of the receivers (stream-ordered, no synchronisation) and goes on.  Stages 0 and 1 only
 * enqueue work; stage 2 ends with the evaluation's one synchronisation and its checks. */
extern "C" int b200sph_rhs_eval_stage(b200sph_handle *h, const b200sph_view *view, int stage, int *pending_sum, int *offender)
{
    int rc = rhs_check_view(h, view);
    if (rc) return rc;
    CU(cudaSetDevice(h->device));
    if (pending_sum) *pending_sum = 0;
    const bool external = h->halo_sums_external && h->n_owned > 0 && h->n_owned < view->n;
    switch (stage) {
        case 0:
            rc = rhs_stage_search(h, *view);
            if (pending_sum && external && h->kernel_sum_density) *pending_sum = B200SPH_SUM_DENSITY;
            return rc;
        case 1:
            rc = rhs_stage_pointwise(h, *view);
            if (pending_sum && external && TENSORIAL_CORRECTION) *pending_sum = B200SPH_SUM_CORRECTION;
            return rc;
        case 2:
            return rhs_stage_forces(h, *view, offender);
        default:
            return B200SPH_ERR_BAD_ARGUMENT;
    }
}

extern "C" int b200sph_export_interactions(b200sph_handle *h, int *interactions, int max_per_row)
{
    if (!h || !interactions || max_per_row <= 0 || h->s.n <= 0) return B200SPH_ERR_BAD_ARGUMENT;
    CU(cudaSetDevice(h->device));
    k_export_interactions<<<blocks_for(h->s.n, 128), 128, 0, h->stream>>>(h->s, interactions, max_per_row);
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    return B200SPH_OK;
}

extern "C" int b200sph_pressure(b200sph_handle *h, const b200sph_view *view)
{
    if (!h || !view || !h->materials_set) return B200SPH_ERR_BAD_ARGUMENT;
    CU(cudaSetDevice(h->device));
    k_pressure_only<<<blocks_for(view->n, 256), 256, 0, h->stream>>>(*view);
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    return B200SPH_OK;
}

extern "C" int b200sph_init_soundspeed(b200sph_handle *h, const b200sph_view *view)
{
    if (!h || !view || !h->materials_set) return B200SPH_ERR_BAD_ARGUMENT;
    CU(cudaSetDevice(h->device));
    k_init_soundspeed<<<blocks_for(view->n, 256), 256, 0, h->stream>>>(*view);
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    return B200SPH_OK;
}

extern "C" int b200sph_damage_limit(b200sph_handle *h, const b200sph_view *view)
{
    if (!h || !view || !h->materials_set) return B200SPH_ERR_BAD_ARGUMENT;
#if FRAGMENTATION
    CU(cudaSetDevice(h->device));
    k_damage_limit<<<blocks_for(view->n, 256), 256, 0, h->stream>>>(*view);
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
#endif
    return B200SPH_OK;
}

/* ------------------------------------------------------------------ persistent cell order
 * The library never needed the caller's buffers in any particular order -- it sorts into scratch -- but every kernel
 * that touches the caller's arrays does so through `perm`, and when the caller's order is unrelated to the cell order
 * (the reference keeps the order of the input file for the whole run, src/memory_handling.cu) those are scattered
 * 8-byte accesses: k_pointwise of the impact moves 737 MB at 2.4 TB/s effective (profiles/r01_ncu_full_impact_v4).
 * b200sph_reorder() puts p_device, the rk buffers and the immutables themselves into cell order, once in a while at a
 * step boundary; afterwards perm is close to the identity and the same kernels run coalesced. */
template <typename T>
__global__ void k_permute_rows(T *dst, const T *src, const int *perm, int n, int per)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)n * per) return;
    const int k = (int)(e / per), c = (int)(e - (size_t)k * per);
    dst[e] = src[(size_t)perm[k] * per + c];
}

struct ReorderItem {
    void *ptr;
    int per;       /* values per particle */
    int bytes;     /* 8 (double) or 4 (int) */
};

static void reorder_collect(const b200sph_particle_arrays &a, int max_num_flaws, ReorderItem *items, int *count, int cap)
{
    auto add = [&](void *ptr, int per, int bytes) {
        if (!ptr) return;
        for (int k = 0; k < *count; k++)
            if (items[k].ptr == ptr) return;   /* p_rhs aliases p in the reference (both are p_device) */
        if (*count < cap) items[(*count)++] = ReorderItem{ptr, per, bytes};
    };
    double *scalars[] = {a.x, a.y, a.z, a.vx, a.vy, a.vz, a.dxdt, a.dydt, a.dzdt, a.ax, a.ay, a.az, a.g_ax, a.g_ay, a.g_az,
                         a.g_local_cellsize, a.g_x, a.g_y, a.g_z, a.m, a.h, a.h0, a.dhdt, a.rho, a.drhodt, a.p, a.e, a.dedt,
                         a.local_strain, a.ep, a.edotp, a.plastic_f, a.d, a.damage_total, a.dddt, a.damage_porjutzi,
                         a.ddamage_porjutzidt, a.muijmax, a.pold, a.alpha_jutzi, a.alpha_jutzi_old, a.dalphadt, a.dalphadp,
                         a.dalphadrho, a.f, a.delpdelrho, a.delpdele, a.cs};
    for (double *ptr : scalars) add(ptr, 1, 8);
    double *tensors[] = {a.S, a.dSdt, a.sigma, a.R, a.tensorialCorrectionMatrix};
    for (double *ptr : tensors) add(ptr, DD, 8);
    add(a.flaws, max_num_flaws, 8);
    int *ints[] = {a.numFlaws, a.numActiveFlaws, a.noi, a.materialId, a.depth};
    for (int *ptr : ints) add(ptr, 1, 4);
}

extern "C" int b200sph_reorder(b200sph_handle *h, const b200sph_view *view, const b200sph_particle_arrays *extra, int n_extra,
                               int *perm_out)
{
    if (!h || !view || n_extra < 0 || (n_extra > 0 && !extra)) return B200SPH_ERR_BAD_ARGUMENT;
    const b200sph_view &v = *view;
    if (v.n <= 0 || v.n > h->n_max || !h->materials_set || !v.p.x || !v.p.h || !v.p_rhs.materialId) return B200SPH_ERR_BAD_ARGUMENT;
    if (h->n_owned > 0 && h->n_owned < v.n) {
        snprintf(h->err, sizeof(h->err), "b200sph_reorder: not with halo copies appended (reorder the owned particles before the exchange)");
        return B200SPH_ERR_BAD_ARGUMENT;
    }
    CU(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    Sorted &s = h->s;
    const int n = v.n;
    ReorderItem items[3 * 64];
    int count = 0;
    reorder_collect(v.p, v.max_num_flaws, items, &count, 3 * 64);
    reorder_collect(v.p_rhs, v.max_num_flaws, items, &count, 3 * 64);
    for (int k = 0; k < n_extra; k++) reorder_collect(extra[k], v.max_num_flaws, items, &count, 3 * 64);
    /* the neighbour lists are void after a reorder: their storage is the staging buffer */
    const size_t stage_bytes = ((size_t)(h->n_max + NBR_TILE - 1) / NBR_TILE) * NBR_TILE * (size_t)MAX_NUM_INTERACTIONS * sizeof(int);
    for (int k = 0; k < count; k++)
        if ((size_t)n * items[k].per * items[k].bytes > stage_bytes) {
            snprintf(h->err, sizeof(h->err), "b200sph_reorder: a member with %d values per particle does not fit the staging buffer", items[k].per);
            return B200SPH_ERR_BAD_ARGUMENT;
        }
    {
        const int blocks = min(blocks_for(n, PREP_THREADS), h->n_sm * 4);
        double3 glo = make_double3(h->global_lo[0], h->global_lo[1], h->global_lo[2]);
        double3 ghi = make_double3(h->global_hi[0], h->global_hi[1], h->global_hi[2]);
        k_prepare<<<blocks, PREP_THREADS, 0, st>>>(v, h->block_partials, h->block_counter, h->d_domain, h->max_cells,
                                                   h->have_global_domain, glo, ghi, 0);
    }
    k_cell_keys<<<blocks_for(n, 256), 256, 0, st>>>(v, h->d_domain, h->keys_in, h->idx_in);
    CU(cub::DeviceRadixSort::SortPairs(h->cub_tmp, h->cub_tmp_bytes, h->keys_in, s.keys, h->idx_in, s.perm, n, 0, h->sort_bits, st));
    for (int k = 0; k < count; k++) {
        const size_t elems = (size_t)n * items[k].per;
        const int blocks = (int)((elems + 255) / 256);
        if (items[k].bytes == 8)
            k_permute_rows<double><<<blocks, 256, 0, st>>>((double *)s.nbr, (const double *)items[k].ptr, s.perm, n, items[k].per);
        else
            k_permute_rows<int><<<blocks, 256, 0, st>>>((int *)s.nbr, (const int *)items[k].ptr, s.perm, n, items[k].per);
        CU(cudaMemcpyAsync(items[k].ptr, s.nbr, elems * items[k].bytes, cudaMemcpyDeviceToDevice, st));
    }
    if (perm_out) CU(cudaMemcpyAsync(perm_out, s.perm, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemsetAsync(s.noi, 0, sizeof(int) * (size_t)n, st));
    CU(cudaMemcpyAsync(&h->h_domain, h->d_domain, sizeof(Domain), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    if (h->h_domain.nonfinite) {
        snprintf(h->err, sizeof(h->err), "non-finite particle coordinate or smoothing length (NaN/Inf) among the %d particles", n);
        return B200SPH_ERR_NONFINITE;
    }
    h->s.n = 0;   /* b200sph_export_interactions has nothing to export until the next evaluation */
    return B200SPH_OK;
}

/* upload of the material tables into constant memory (used by capi.cu) */
int upload_materials(b200sph_handle *h, const MatParams *host, int n, const AneosTables &tables)
{
    CU(cudaMemcpyToSymbol(c_mat, host, sizeof(MatParams) * n));
    CU(cudaMemcpyToSymbol(c_aneos, &tables, sizeof(AneosTables)));
    return 0;
}

int sort_temp_bytes(int n_max, int bits, size_t *bytes)
{
    int *k = nullptr;
    return cub::DeviceRadixSort::SortPairs(nullptr, *bytes, k, k, k, k, n_max, 0, bits) == cudaSuccess ? 0 : -1;
}
