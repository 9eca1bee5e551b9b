/*
 * halo.cu -- device side of the multi-GPU halo exchange (the reference is single-GPU).
 *
 * Every rank's domain is a union of axis-aligned octree boxes (the cells of its Morton key range,
 * miluphcuda_b200/multigpu.py).  Per evaluation of the right-hand side a rank
 *
 *   h_box_hmax  finds the largest smoothing length inside each of ITS boxes (all-gathered by the host),
 *   h_mask      marks, for every owned particle, the ranks that need a copy: distance to one of the rank's
 *               boxes below h_k + extra(box)  (extra = that box's largest h for a two-level halo, 0 for one level),
 *   h_scan      turns the per-block, per-rank counts into offsets (deterministic, ascending particle order),
 *   h_write     writes the send list, grouped by destination rank,
 *   h_pack      gathers the state of the listed particles into one row-major FP64 send buffer,
 *   h_unpack    scatters received rows behind the owned particles.
 *
 * All of it is stream-ordered on the handle's stream and never synchronises: the host reads back only the
 * per-rank counts it needs as NCCL split sizes.
 */
#include "rhs_internal.h"

#include <stdio.h>
#include <string.h>

#define HALO_THREADS 256

#define HCU(call)                                                                                    \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            snprintf(h->err, sizeof(h->err), "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return B200SPH_ERR_CUDA;                                                                 \
        }                                                                                            \
    } while (0)

void halo_state_destroy(b200sph_handle *h)
{
    HaloState *st = (HaloState *)h->halo;
    if (!st) return;
    cudaFree(st->dev); cudaFree(st->mask); cudaFree(st->blk_counts);
    free(st);
    h->halo = nullptr;
}

extern "C" int b200sph_halo_set_domains(b200sph_handle *h, const double *boxes, const int *box_rank, int n_boxes, int n_ranks, int my_rank)
{
    if (!h || !boxes || !box_rank || n_boxes <= 0 || n_ranks <= 0 || my_rank < 0 || my_rank >= n_ranks) return B200SPH_ERR_BAD_ARGUMENT;
    if (n_boxes > HALO_MAX_BOXES || n_ranks > HALO_MAX_RANKS) {
        snprintf(h->err, sizeof(h->err), "halo domains: %d boxes / %d ranks exceed the limits %d / %d", n_boxes, n_ranks, HALO_MAX_BOXES,
                 HALO_MAX_RANKS);
        return B200SPH_ERR_UNSUPPORTED;
    }
    HCU(cudaSetDevice(h->device));
    HaloState *st = (HaloState *)h->halo;
    if (!st) {
        st = (HaloState *)calloc(1, sizeof(HaloState));
        if (!st) return B200SPH_ERR_BAD_ARGUMENT;
        h->halo = st;
        HCU(cudaMalloc((void **)&st->dev, sizeof(HaloDomains)));
    }
    HaloDomains &d = st->host;
    memset(&d, 0, sizeof(d));
    d.n_boxes = n_boxes; d.n_ranks = n_ranks; d.my_rank = my_rank; d.my_first = -1;
    d.list_reach_scale = 1.0; d.list_skin = 0.0;
    int per_rank[HALO_MAX_RANKS] = {0};
    for (int b = 0; b < n_boxes; b++) {
        const int r = box_rank[b];
        if (r < 0 || r >= n_ranks || (b > 0 && r < box_rank[b - 1])) {
            snprintf(h->err, sizeof(h->err), "halo domains: box_rank must be non-decreasing and inside [0, %d)", n_ranks);
            return B200SPH_ERR_BAD_ARGUMENT;
        }
        for (int a = 0; a < 3; a++) { d.lo[b][a] = boxes[6 * b + a]; d.hi[b][a] = boxes[6 * b + 3 + a]; }
        d.rank[b] = r;
        d.local[b] = per_rank[r]++;
        if (r == my_rank && d.my_first < 0) d.my_first = b;
    }
    d.my_count = per_rank[my_rank];
    if (d.my_first < 0) d.my_first = 0;
    HCU(cudaMemcpy(st->dev, &d, sizeof(HaloDomains), cudaMemcpyHostToDevice));
    return B200SPH_OK;
}

extern "C" int b200sph_halo_set_list_margin(b200sph_handle *h, double reach_scale, double skin)
{
    if (!h || !h->halo || !(reach_scale >= 1.0) || !(skin >= 0.0)) return B200SPH_ERR_BAD_ARGUMENT;
    HaloState *st = (HaloState *)h->halo;
    HCU(cudaSetDevice(h->device));
    st->host.list_reach_scale = reach_scale;
    st->host.list_skin = skin;
    HCU(cudaMemcpyAsync(&st->dev->list_reach_scale, &st->host.list_reach_scale, 2 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    return B200SPH_OK;
}

extern "C" int b200sph_set_halo_sums(b200sph_handle *h, int external)
{
    if (!h) return B200SPH_ERR_BAD_ARGUMENT;
    h->halo_sums_external = external ? 1 : 0;
    return B200SPH_OK;
}

extern "C" int b200sph_set_abort_flag(b200sph_handle *h, const int *device_flag)
{
    if (!h) return B200SPH_ERR_BAD_ARGUMENT;
    h->abort_flag = device_flag;
    return B200SPH_OK;
}

static int halo_scratch(b200sph_handle *h, HaloState *st, int n)
{
    const int n_blocks = (n + HALO_THREADS - 1) / HALO_THREADS;
    if (n > st->mask_capacity) {
        cudaFree(st->mask);
        st->mask = nullptr;
        HCU(cudaMalloc((void **)&st->mask, sizeof(unsigned long long) * (size_t)n));
        st->mask_capacity = n;
    }
    const int need = n_blocks * st->host.n_ranks + st->host.n_ranks + 8;
    if (need > st->blk_capacity) {
        cudaFree(st->blk_counts);
        st->blk_counts = nullptr;
        HCU(cudaMalloc((void **)&st->blk_counts, sizeof(int) * (size_t)need));
        st->blk_capacity = need;
    }
    return 0;
}

/* ------------------------------------------------------------------ largest h per own box */
__global__ void __launch_bounds__(HALO_THREADS)
h_box_hmax(const double *x, const double *y, const double *z, const double *sml, int n, const HaloDomains *dom, unsigned long long *hmax_bits)
{
    /* block-local maxima first: a million same-address global atomics would serialise in L2 */
    __shared__ unsigned long long sh_max[HALO_MAX_BOXES];
    const int first = dom->my_first, count = dom->my_count;
    for (int b = threadIdx.x; b < count; b += blockDim.x) sh_max[b] = 0ull;
    __syncthreads();
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) {
        const double p[3] = {x[k], (DIM > 1 && y) ? y[k] : 0.0, (DIM > 2 && z) ? z[k] : 0.0};
        /* positive doubles order like their bit patterns */
        const unsigned long long bits = (unsigned long long)__double_as_longlong(sml[k]);
        for (int b = 0; b < count; b++) {
            bool inside = true;
#pragma unroll
            for (int a = 0; a < DIM; a++) inside = inside && p[a] >= dom->lo[first + b][a] && p[a] <= dom->hi[first + b][a];
            if (inside && bits > sh_max[b]) atomicMax(&sh_max[b], bits);
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < count; b += blockDim.x)
        if (sh_max[b] != 0ull) atomicMax(&hmax_bits[b], sh_max[b]);
}

extern "C" int b200sph_halo_box_hmax(b200sph_handle *h, const double *x, const double *y, const double *z, const double *sml, int n,
                                     double *hmax_out, int hmax_len)
{
    HaloState *st = h ? (HaloState *)h->halo : nullptr;
    if (!st || !x || !sml || !hmax_out || n < 0 || hmax_len < st->host.my_count) return B200SPH_ERR_BAD_ARGUMENT;
    HCU(cudaSetDevice(h->device));
    HCU(cudaMemsetAsync(hmax_out, 0, sizeof(double) * hmax_len, h->stream));
    if (n > 0)
        h_box_hmax<<<(n + HALO_THREADS - 1) / HALO_THREADS, HALO_THREADS, 0, h->stream>>>(x, y, z, sml, n, st->dev,
                                                                                           reinterpret_cast<unsigned long long *>(hmax_out));
    HCU(cudaGetLastError());
    return B200SPH_OK;
}

/* ------------------------------------------------------------------ who needs which particle
 * Two-stage test: a rank is looked at only when the particle is within reach of the bounding box of ALL its
 * boxes (with the largest `extra` among them); then its boxes are tested one by one.  Interior particles --
 * the vast majority -- cost n_ranks - 1 tests instead of one per box of every rank. */
__global__ void __launch_bounds__(HALO_THREADS)
h_mask(const double *x, const double *y, const double *z, const double *sml, int n, const HaloDomains *dom, const double *extra,
       int extra_stride, double reach_scale, double skin, unsigned long long *mask_out, int *blk_counts, int n_blocks)
{
    extern __shared__ double sh_box[];   /* n_boxes x {lo[3], hi[3], extra}, n_ranks x {lo[3], hi[3], extra}, then int tables */
    const int n_boxes = dom->n_boxes, n_ranks = dom->n_ranks, my_rank = dom->my_rank;
    double *bx = sh_box;
    double *rb = sh_box + 7 * n_boxes;
    int *first = reinterpret_cast<int *>(rb + 7 * n_ranks);   /* first box of rank r (n_boxes for a rank without boxes) */
    for (int r = threadIdx.x; r <= n_ranks; r += blockDim.x) first[r] = n_boxes;
    __syncthreads();
    for (int b = threadIdx.x; b < n_boxes; b += blockDim.x) {
        const int r = dom->rank[b];
        for (int a = 0; a < 3; a++) {
            bx[7 * b + a] = dom->lo[b][a];
            bx[7 * b + 3 + a] = dom->hi[b][a];
        }
        bx[7 * b + 6] = extra ? extra[r * extra_stride + dom->local[b]] : 0.0;
        if (dom->local[b] == 0) first[r] = b;   /* box_rank is non-decreasing: a rank's boxes are contiguous */
    }
    __syncthreads();
    for (int r = threadIdx.x; r < n_ranks; r += blockDim.x) {
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300}, ex = 0.0;
        for (int b = first[r]; b < n_boxes && dom->rank[b] == r; b++) {
            for (int a = 0; a < 3; a++) {
                lo[a] = fmin(lo[a], bx[7 * b + a]);
                hi[a] = fmax(hi[a], bx[7 * b + 3 + a]);
            }
            ex = fmax(ex, bx[7 * b + 6]);
        }
        for (int a = 0; a < 3; a++) { rb[7 * r + a] = lo[a]; rb[7 * r + 3 + a] = hi[a]; }
        rb[7 * r + 6] = ex;
    }
    __syncthreads();
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long mask = 0ull;
    if (k < n) {
        const double p[3] = {x[k], (DIM > 1 && y) ? y[k] : 0.0, (DIM > 2 && z) ? z[k] : 0.0};
        const double hk = sml[k];
        for (int r = 0; r < n_ranks; r++) {
            if (r == my_rank) continue;
            double d2 = 0.0;
#pragma unroll
            for (int a = 0; a < DIM; a++) {
                const double g = fmax(fmax(rb[7 * r + a] - p[a], p[a] - rb[7 * r + 3 + a]), 0.0);
                d2 = fma(g, g, d2);
            }
            const double far = (hk + rb[7 * r + 6]) * reach_scale * (1.0 + 1e-9) + skin;
            if (!(d2 < far * far)) continue;
            for (int b = first[r]; b < n_boxes && dom->rank[b] == r; b++) {
                d2 = 0.0;
#pragma unroll
                for (int a = 0; a < DIM; a++) {
                    const double g = fmax(fmax(bx[7 * b + a] - p[a], p[a] - bx[7 * b + 3 + a]), 0.0);
                    d2 = fma(g, g, d2);
                }
                const double reach = (hk + bx[7 * b + 6]) * reach_scale * (1.0 + 1e-9) + skin;
                if (d2 < reach * reach) {
                    mask |= 1ull << r;
                    break;
                }
            }
        }
        mask_out[k] = mask;
    }
    for (int r = 0; r < n_ranks; r++) {
        const int c = __syncthreads_count((int)((mask >> r) & 1ull));
        if (threadIdx.x == 0) blk_counts[r * n_blocks + blockIdx.x] = c;
    }
}

/* exclusive scan of the per-block counts, rank after rank; counts_out[r] = particles for rank r,
 * counts_out[n_ranks] = 1 if the send list does not fit */
__global__ void __launch_bounds__(1024)
h_scan(int *blk_counts, int n_blocks, int n_ranks, int idx_capacity, int *counts_out)
{
    __shared__ int warp_tot[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = 0; r < n_ranks; r++) {
        const int start = carry;
        int *c = blk_counts + (size_t)r * n_blocks;
        for (int base = 0; base < n_blocks; base += 1024) {
            const int i = base + threadIdx.x;
            const int v = (i < n_blocks) ? c[i] : 0;
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) warp_tot[warp] = incl;
            __syncthreads();
            int woff = 0;
            for (int w = 0; w < warp; w++) woff += warp_tot[w];
            const int base_off = carry;
            if (i < n_blocks) c[i] = base_off + woff + incl - v;
            __syncthreads();
            if (threadIdx.x == 1023) carry = base_off + woff + incl;
            __syncthreads();
        }
        if (threadIdx.x == 0) counts_out[r] = carry - start;
        __syncthreads();
    }
    if (threadIdx.x == 0) counts_out[n_ranks] = (carry > idx_capacity) ? 1 : 0;
}

__global__ void __launch_bounds__(HALO_THREADS)
h_write(const unsigned long long *mask, int n, int n_ranks, const int *blk_offsets, int n_blocks, int idx_capacity, int *idx_out)
{
    __shared__ int warp_tot[HALO_THREADS / 32];
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long m = (k < n) ? mask[k] : 0ull;
    for (int r = 0; r < n_ranks; r++) {
        const bool flag = (m >> r) & 1ull;
        const unsigned int bal = __ballot_sync(0xffffffffu, flag);
        if (lane == 0) warp_tot[warp] = __popc(bal);
        __syncthreads();
        if (flag) {
            int off = blk_offsets[r * n_blocks + blockIdx.x] + __popc(bal & ((1u << lane) - 1u));
            for (int w = 0; w < warp; w++) off += warp_tot[w];
            if (off < idx_capacity) idx_out[off] = k;
        }
        __syncthreads();
    }
}

extern "C" int b200sph_halo_select_plan(b200sph_handle *h, const double *x, const double *y, const double *z, const double *sml, int n,
                                        const double *extra, int extra_stride, double reach_scale, double skin, int *idx_out,
                                        int idx_capacity, int *counts_out);

extern "C" int b200sph_halo_select(b200sph_handle *h, const double *x, const double *y, const double *z, const double *sml, int n,
                                   const double *extra, int extra_stride, int *idx_out, int idx_capacity, int *counts_out)
{
    return b200sph_halo_select_plan(h, x, y, z, sml, n, extra, extra_stride, 1.0, 0.0, idx_out, idx_capacity, counts_out);
}

extern "C" int b200sph_halo_select_plan(b200sph_handle *h, const double *x, const double *y, const double *z, const double *sml, int n,
                                        const double *extra, int extra_stride, double reach_scale, double skin, int *idx_out,
                                        int idx_capacity, int *counts_out)
{
    if (!(reach_scale >= 1.0) || !(skin >= 0.0)) return B200SPH_ERR_BAD_ARGUMENT;
    HaloState *st = h ? (HaloState *)h->halo : nullptr;
    if (!st || !x || !sml || !idx_out || !counts_out || n <= 0) return B200SPH_ERR_BAD_ARGUMENT;
    HCU(cudaSetDevice(h->device));
    if (halo_scratch(h, st, n)) return B200SPH_ERR_CUDA;
    const int n_blocks = (n + HALO_THREADS - 1) / HALO_THREADS;
    const HaloDomains &d = st->host;
    const size_t smem = (size_t)(d.n_boxes + d.n_ranks) * 7 * sizeof(double) + (size_t)(d.n_ranks + 1) * sizeof(int);
    if (smem > 48 * 1024) HCU(cudaFuncSetAttribute(h_mask, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    h_mask<<<n_blocks, HALO_THREADS, smem, h->stream>>>(x, y, z, sml, n, st->dev, extra, extra_stride, reach_scale, skin, st->mask,
                                                        st->blk_counts, n_blocks);
    h_scan<<<1, 1024, 0, h->stream>>>(st->blk_counts, n_blocks, d.n_ranks, idx_capacity, counts_out);
    h_write<<<n_blocks, HALO_THREADS, 0, h->stream>>>(st->mask, n, d.n_ranks, st->blk_counts, n_blocks, idx_capacity, idx_out);
    HCU(cudaGetLastError());
    return B200SPH_OK;
}

/* ------------------------------------------------------------------ is a send plan still valid?
 * A plan (send list + counts) built with reach (h_k + extra) * (1 + growth) + 2 * max_move stays complete while no
 * particle has moved further than max_move from where it was when the plan was built and no smoothing length
 * has grown by more than `growth`.  One flag for all particles; the host all-reduces it over the ranks. */
__global__ void __launch_bounds__(HALO_THREADS)
h_plan_check(const double *x, const double *y, const double *z, const double *sml, const double *x0, const double *y0, const double *z0,
             const double *sml0, int n, double max_move2, double growth, int *flag)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    bool bad = false;
    if (k < n) {
        double d2 = (x[k] - x0[k]) * (x[k] - x0[k]);
        if (DIM > 1 && y) d2 += (y[k] - y0[k]) * (y[k] - y0[k]);
        if (DIM > 2 && z) d2 += (z[k] - z0[k]) * (z[k] - z0[k]);
        bad = !(d2 <= max_move2) || !(sml[k] <= sml0[k] * (1.0 + growth));   /* NaNs count as a violation */
    }
    if (__syncthreads_or((int)bad) && threadIdx.x == 0) atomicOr(flag, 1);
}

extern "C" int b200sph_halo_plan_check(b200sph_handle *h, const double *x, const double *y, const double *z, const double *sml,
                                       const double *x0, const double *y0, const double *z0, const double *sml0, int n,
                                       double max_move, double growth, int *flag_out)
{
    if (!h || !x || !sml || !x0 || !sml0 || !flag_out || n < 0 || !(max_move >= 0.0) || !(growth >= 0.0)) return B200SPH_ERR_BAD_ARGUMENT;
    HCU(cudaSetDevice(h->device));
    HCU(cudaMemsetAsync(flag_out, 0, sizeof(int), h->stream));
    if (n > 0)
        h_plan_check<<<(n + HALO_THREADS - 1) / HALO_THREADS, HALO_THREADS, 0, h->stream>>>(x, y, z, sml, x0, y0, z0, sml0, n,
                                                                                             max_move * max_move, growth, flag_out);
    HCU(cudaGetLastError());
    return B200SPH_OK;
}

/* ------------------------------------------------------------------ pack / unpack */
#define HALO_MAX_FIELDS 96
struct HaloFields {
    void *data[HALO_MAX_FIELDS];
    int per[HALO_MAX_FIELDS];       /* values per particle */
    int kind[HALO_MAX_FIELDS];      /* 0 double, 1 int32, 2 int32 zero-filled on unpack (not transported) */
    int col[HALO_MAX_FIELDS + 1];   /* first column in a row */
    int n_fields, width;
};

__global__ void __launch_bounds__(256)
h_pack(HaloFields f, const int *idx, int n_rows, double *out)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n_rows * f.width) return;
    const int row = (int)(t / f.width), c = (int)(t % f.width);
    int q = 0;
    while (c >= f.col[q + 1]) q++;
    const int comp = c - f.col[q];
    const size_t src = (size_t)idx[row] * f.per[q] + comp;
    out[t] = (f.kind[q] == 0) ? reinterpret_cast<const double *>(f.data[q])[src] : (double)reinterpret_cast<const int *>(f.data[q])[src];
}

__global__ void __launch_bounds__(256)
h_unpack(HaloFields f, const double *in, int n_rows, int first_row)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n_rows * f.width) return;
    const int row = (int)(t / f.width), c = (int)(t % f.width);
    int q = 0;
    while (c >= f.col[q + 1]) q++;
    const int comp = c - f.col[q];
    const size_t dst = (size_t)(first_row + row) * f.per[q] + comp;
    if (f.kind[q] == 0) reinterpret_cast<double *>(f.data[q])[dst] = in[t];
    else reinterpret_cast<int *>(f.data[q])[dst] = (int)in[t];
}

__global__ void h_zero_rows(int *data, int per, int n_rows, int first_row)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_rows * per) data[(size_t)first_row * per + t] = 0;
}

static int halo_fields(b200sph_handle *h, const b200sph_halo_field *fields, int n_fields, HaloFields &f)
{
    if (!fields || n_fields <= 0 || n_fields > HALO_MAX_FIELDS) {
        snprintf(h->err, sizeof(h->err), "halo pack: %d fields (limit %d)", n_fields, HALO_MAX_FIELDS);
        return B200SPH_ERR_BAD_ARGUMENT;
    }
    memset(&f, 0, sizeof(f));
    int n = 0, col = 0;
    for (int q = 0; q < n_fields; q++) {
        if (!fields[q].data || fields[q].per <= 0 || fields[q].kind == 2) continue;
        f.data[n] = fields[q].data; f.per[n] = fields[q].per; f.kind[n] = fields[q].kind; f.col[n] = col;
        col += fields[q].per;
        n++;
    }
    f.col[n] = col;
    f.n_fields = n;
    f.width = col;
    return 0;
}

extern "C" int b200sph_halo_row_width(const b200sph_halo_field *fields, int n_fields)
{
    int w = 0;
    for (int q = 0; fields && q < n_fields; q++)
        if (fields[q].data && fields[q].per > 0 && fields[q].kind != 2) w += fields[q].per;
    return w;
}

/* ------------------------------------------------------------------ pack / unpack, column-major per rank
 * Row-major rows make every warp touch `width` different arrays at once (153 us for 48 MB on the impact case,
 * a tenth of the HBM rate).  Here the block of rows going to (coming from) one rank is stored column by column:
 *     buffer = [rank 0: col 0 rows.. | col 1 rows.. | ...][rank 1: ...]
 * so a warp walks ONE member array along ascending particle indices and writes one contiguous run.  The blocks
 * of different ranks stay contiguous, which is all all_to_all needs (split sizes = rows * width). */
__device__ __forceinline__ int rank_of_row(const int *prefix, int n_ranks, int g)
{
    int r = 0;
    while (r + 1 < n_ranks && g >= prefix[r + 1]) r++;
    return r;
}

__global__ void __launch_bounds__(256)
h_pack_cols(HaloFields f, const int *idx, const int *counts, int n_ranks, int n_rows, double *out)
{
    __shared__ int prefix[HALO_MAX_RANKS + 1];
    if (threadIdx.x == 0) {
        int run = 0;
        for (int r = 0; r < n_ranks; r++) { prefix[r] = run; run += counts[r]; }
        prefix[n_ranks] = run;
    }
    __syncthreads();
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n_rows * f.width) return;
    const int c = (int)(t / n_rows), g = (int)(t % n_rows);
    int q = 0;
    while (c >= f.col[q + 1]) q++;
    const int comp = c - f.col[q];
    const int r = rank_of_row(prefix, n_ranks, g);
    const int cnt = prefix[r + 1] - prefix[r];
    const size_t src = (size_t)idx[g] * f.per[q] + comp;
    const double v = (f.kind[q] == 0) ? reinterpret_cast<const double *>(f.data[q])[src] : (double)reinterpret_cast<const int *>(f.data[q])[src];
    out[(size_t)prefix[r] * f.width + (size_t)c * cnt + (g - prefix[r])] = v;
}

__global__ void __launch_bounds__(256)
h_unpack_cols(HaloFields f, const double *in, const int *counts, int n_ranks, int n_rows, int first_row)
{
    __shared__ int prefix[HALO_MAX_RANKS + 1];
    if (threadIdx.x == 0) {
        int run = 0;
        for (int r = 0; r < n_ranks; r++) { prefix[r] = run; run += counts[r]; }
        prefix[n_ranks] = run;
    }
    __syncthreads();
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n_rows * f.width) return;
    const int c = (int)(t / n_rows), g = (int)(t % n_rows);
    int q = 0;
    while (c >= f.col[q + 1]) q++;
    const int comp = c - f.col[q];
    const int r = rank_of_row(prefix, n_ranks, g);
    const int cnt = prefix[r + 1] - prefix[r];
    const double v = in[(size_t)prefix[r] * f.width + (size_t)c * cnt + (g - prefix[r])];
    const size_t dst = (size_t)(first_row + g) * f.per[q] + comp;
    if (f.kind[q] == 0) reinterpret_cast<double *>(f.data[q])[dst] = v;
    else reinterpret_cast<int *>(f.data[q])[dst] = (int)v;
}

extern "C" int b200sph_halo_pack_by_rank(b200sph_handle *h, const b200sph_halo_field *fields, int n_fields, const int *idx,
                                         const int *counts, int n_ranks, int n_rows, double *out)
{
    if (!h || !idx || !out || !counts || n_rows < 0 || n_ranks <= 0 || n_ranks > HALO_MAX_RANKS) return B200SPH_ERR_BAD_ARGUMENT;
    HaloFields f;
    if (int rc = halo_fields(h, fields, n_fields, f)) return rc;
    if (n_rows == 0 || f.width == 0) return B200SPH_OK;
    HCU(cudaSetDevice(h->device));
    const long long total = (long long)n_rows * f.width;
    h_pack_cols<<<(unsigned int)((total + 255) / 256), 256, 0, h->stream>>>(f, idx, counts, n_ranks, n_rows, out);
    HCU(cudaGetLastError());
    return B200SPH_OK;
}

extern "C" int b200sph_halo_unpack_by_rank(b200sph_handle *h, const b200sph_halo_field *fields, int n_fields, const double *in,
                                           const int *counts, int n_ranks, int n_rows, int first_row)
{
    if (!h || !in || !counts || n_rows < 0 || first_row < 0 || n_ranks <= 0 || n_ranks > HALO_MAX_RANKS) return B200SPH_ERR_BAD_ARGUMENT;
    HaloFields f;
    if (int rc = halo_fields(h, fields, n_fields, f)) return rc;
    if (n_rows == 0) return B200SPH_OK;
    HCU(cudaSetDevice(h->device));
    if (f.width > 0) {
        const long long total = (long long)n_rows * f.width;
        h_unpack_cols<<<(unsigned int)((total + 255) / 256), 256, 0, h->stream>>>(f, in, counts, n_ranks, n_rows, first_row);
    }
    for (int q = 0; q < n_fields; q++)
        if (fields[q].data && fields[q].kind == 2 && fields[q].per > 0)
            h_zero_rows<<<(n_rows * fields[q].per + 255) / 256, 256, 0, h->stream>>>((int *)fields[q].data, fields[q].per, n_rows, first_row);
    HCU(cudaGetLastError());
    return B200SPH_OK;
}

extern "C" int b200sph_halo_pack(b200sph_handle *h, const b200sph_halo_field *fields, int n_fields, const int *idx, int n_rows, double *out)
{
    if (!h || !idx || !out || n_rows < 0) return B200SPH_ERR_BAD_ARGUMENT;
    HaloFields f;
    if (int rc = halo_fields(h, fields, n_fields, f)) return rc;
    if (n_rows == 0 || f.width == 0) return B200SPH_OK;
    HCU(cudaSetDevice(h->device));
    const long long total = (long long)n_rows * f.width;
    h_pack<<<(unsigned int)((total + 255) / 256), 256, 0, h->stream>>>(f, idx, n_rows, out);
    HCU(cudaGetLastError());
    return B200SPH_OK;
}

extern "C" int b200sph_halo_unpack(b200sph_handle *h, const b200sph_halo_field *fields, int n_fields, const double *in, int n_rows, int first_row)
{
    if (!h || !in || n_rows < 0 || first_row < 0) return B200SPH_ERR_BAD_ARGUMENT;
    HaloFields f;
    if (int rc = halo_fields(h, fields, n_fields, f)) return rc;
    if (n_rows == 0) return B200SPH_OK;
    HCU(cudaSetDevice(h->device));
    if (f.width > 0) {
        const long long total = (long long)n_rows * f.width;
        h_unpack<<<(unsigned int)((total + 255) / 256), 256, 0, h->stream>>>(f, in, n_rows, first_row);
    }
    for (int q = 0; q < n_fields; q++)
        if (fields[q].data && fields[q].kind == 2 && fields[q].per > 0)
            h_zero_rows<<<(n_rows * fields[q].per + 255) / 256, 256, 0, h->stream>>>((int *)fields[q].data, fields[q].per, n_rows, first_row);
    HCU(cudaGetLastError());
    return B200SPH_OK;
}
