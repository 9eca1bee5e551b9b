/*
 * mg.cu -- the multi-GPU host in C++ over NCCL (the reference is single-GPU).
 *
 * One process (or thread) per GPU.  A b200sph_mg object wraps one b200sph handle and one NCCL communicator and does,
 * behind the C-ABI, everything a C host needs to evaluate the right-hand side of a particle set spread over the GPUs
 * of one box:
 *
 *   b200sph_mg_decompose   global bounding cube (all-reduce), histogram of the particles over the octree cells of one
 *                          level along the Morton curve (all-reduce), contiguous cell ranges of equal particle count or
 *                          equal work (sum of interaction counts), every rank's range as a union of aligned boxes
 *   b200sph_mg_migrate     full particle records move to the rank that owns their cell (grouped ncclSend/ncclRecv);
 *                          the stayers are compacted, the arrivals appended -- the repartition step of a long run
 *   b200sph_mg_rhs_eval    reusable one-level halo send plan (csrc/halo.cu) -> state rows over NVLink -> the staged
 *                          evaluation with the neighbour-sum exchange between its stages; the plan's verdict is a device
 *                          flag (all-reduced in-stream) that doubles as the evaluation's abort flag
 *
 * NCCL is bound at run time (dlopen of libnccl.so.2): a Python host that already loaded PyTorch's NCCL shares that
 * copy, a C host picks up the system library; single-GPU users of libb200sph never load it.
 */
#include "rhs_internal.h"

#include <cub/cub.cuh>
#include <dlfcn.h>
#include <nccl.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <vector>

#define MG_THREADS 256

struct NcclApi {
    void *lib;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)(void);
    ncclResult_t (*GroupEnd)(void);
    const char *(*GetErrorString)(ncclResult_t);
};
static NcclApi g_nccl;

static int nccl_bind(char *err, size_t errlen)
{
    if (g_nccl.lib) return 0;
    void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) {
        snprintf(err, errlen, "cannot load libnccl.so.2: %s", dlerror());
        return -1;
    }
#define BIND(name)                                                                        \
    do {                                                                                  \
        *(void **)(&g_nccl.name) = dlsym(lib, "nccl" #name);                              \
        if (!g_nccl.name) {                                                               \
            snprintf(err, errlen, "libnccl lacks the symbol nccl" #name);                 \
            return -1;                                                                    \
        }                                                                                 \
    } while (0)
    BIND(GetUniqueId); BIND(CommInitRank); BIND(CommDestroy); BIND(AllReduce); BIND(AllGather); BIND(Send); BIND(Recv);
    BIND(GroupStart); BIND(GroupEnd); BIND(GetErrorString);
#undef BIND
    g_nccl.lib = lib;
    return 0;
}

struct b200sph_mg {
    b200sph_handle *h;
    ncclComm_t comm;
    int rank, world, dim, level;
    double cube_lo[3], cube_span;                 /* bounding CUBE of the global particle set */
    std::vector<long long> cuts;                  /* world + 1 Morton cell ids */
    long long *d_cuts;
    unsigned long long *d_hist;                   /* 2^(dim*level) bins */
    double *d_red;                                /* small reduction scratch (8 doubles) */
    /* send plan */
    int have_plan, plan_n_owned, plan_builds, stale_plans;
    std::vector<int> send_counts, recv_counts;
    int n_send, n_recv;
    int *d_idx, idx_capacity, *d_counts, *d_all_counts, *d_send_counts, *d_recv_counts, *d_flag;
    double *d_send, *d_recv;
    size_t send_cap, recv_cap;
    double *d_snap[4];
    int snap_cap;
    double max_move, growth;
    /* migration scratch */
    int *d_dest, *d_dest_sorted, *d_order_in, *d_order;
    void *d_sort_tmp;
    size_t sort_tmp_bytes;
    int mig_cap;
    /* gravity sources (replicated tree) */
    double *d_grav;
    size_t grav_cap;
    b200sph_mg_stats stats;
    char err[512];
};

#define MCU(call)                                                                                     \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            snprintf(mg->err, sizeof(mg->err), "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return B200SPH_ERR_CUDA;                                                                  \
        }                                                                                             \
    } while (0)
#define MNC(call)                                                                                     \
    do {                                                                                              \
        ncclResult_t r_ = (call);                                                                     \
        if (r_ != ncclSuccess) {                                                                      \
            snprintf(mg->err, sizeof(mg->err), "%s:%d: %s: %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); \
            return B200SPH_ERR_CUDA;                                                                  \
        }                                                                                             \
    } while (0)
#define MRC(call)                                                                                     \
    do {                                                                                              \
        const int rc_ = (call);                                                                       \
        if (rc_) {                                                                                    \
            snprintf(mg->err, sizeof(mg->err), "%s", b200sph_last_error(mg->h));                      \
            return rc_;                                                                               \
        }                                                                                             \
    } while (0)

/* ------------------------------------------------------------------ device helpers */
__device__ __forceinline__ unsigned long long mg_order_bits(double v)
{
    /* doubles -> unsigned integers with the same order */
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double mg_from_order_bits(unsigned long long u)
{
    const unsigned long long b = (u & 0x8000000000000000ull) ? (u & 0x7fffffffffffffffull) : ~u;
    double v;
    memcpy(&v, &b, sizeof(v));
    return v;
}

/* [0..2] max of -x (i.e. -min), [3..5] max of x, [6] max of -h (i.e. -hmin): one all-reduce(max) does it all */
__global__ void mg_extrema(const double *x, const double *y, const double *z, const double *h, int n, unsigned long long *out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double c[3] = {x[i], y ? y[i] : 0.0, z ? z[i] : 0.0};
    for (int a = 0; a < 3; a++) {
        atomicMax(&out[a], mg_order_bits(-c[a]));
        atomicMax(&out[3 + a], mg_order_bits(c[a]));
    }
    atomicMax(&out[6], mg_order_bits(-h[i]));
}

__device__ __forceinline__ long long mg_cell_id(const double c[3], const double lo[3], double span, int dim, int level)
{
    const int g = 1 << level;
    long long id = 0;
    int q[3];
    for (int a = 0; a < dim; a++) {
        int v = (int)((c[a] - lo[a]) / span * g);
        q[a] = v < 0 ? 0 : (v >= g ? g - 1 : v);
    }
    for (int b = 0; b < level; b++)
        for (int a = 0; a < dim; a++) id |= (long long)((q[a] >> b) & 1) << (dim * b + a);
    return id;
}

struct MgCube {
    double lo[3], span;
    int dim, level;
};

__global__ void mg_histogram(const double *x, const double *y, const double *z, const int *weight, int n, MgCube cube,
                             unsigned long long *hist)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double c[3] = {x[i], y ? y[i] : 0.0, z ? z[i] : 0.0};
    atomicAdd(&hist[mg_cell_id(c, cube.lo, cube.span, cube.dim, cube.level)], (unsigned long long)(weight ? 1 + weight[i] : 1));
}

/* destination of every held particle: its owner, or `world` when it stays */
__global__ void mg_destinations(const double *x, const double *y, const double *z, int n, MgCube cube, const long long *cuts, int world,
                                int me, int *dest, int *order, int *counts)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double c[3] = {x[i], y ? y[i] : 0.0, z ? z[i] : 0.0};
    const long long id = mg_cell_id(c, cube.lo, cube.span, cube.dim, cube.level);
    int lo = 0, hi = world;   /* largest r with cuts[r] <= id */
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (cuts[mid] <= id) lo = mid; else hi = mid;
    }
    const int d = (lo == me) ? world : lo;
    dest[i] = d;
    order[i] = i;
    atomicAdd(&counts[d], 1);
}

/* ------------------------------------------------------------------ life time */
extern "C" int b200sph_mg_unique_id(void *id128, char *err, size_t errlen)
{
    char local[256];
    if (!id128) return B200SPH_ERR_BAD_ARGUMENT;
    if (nccl_bind(err ? err : local, err ? errlen : sizeof(local))) return B200SPH_ERR_UNSUPPORTED;
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return B200SPH_ERR_CUDA;
    memcpy(id128, &id, sizeof(id));
    return B200SPH_OK;
}

extern "C" const char *b200sph_mg_last_error(const b200sph_mg *mg) { return mg ? mg->err : "no multi-GPU object"; }

extern "C" int b200sph_mg_destroy(b200sph_mg *mg)
{
    if (!mg) return B200SPH_OK;
    cudaSetDevice(mg->h->device);
    b200sph_set_abort_flag(mg->h, nullptr);
    if (mg->comm) g_nccl.CommDestroy(mg->comm);
    cudaFree(mg->d_cuts); cudaFree(mg->d_hist); cudaFree(mg->d_red); cudaFree(mg->d_idx); cudaFree(mg->d_counts);
    cudaFree(mg->d_all_counts); cudaFree(mg->d_send_counts); cudaFree(mg->d_recv_counts); cudaFree(mg->d_flag);
    cudaFree(mg->d_send); cudaFree(mg->d_recv);
    for (int k = 0; k < 4; k++) cudaFree(mg->d_snap[k]);
    cudaFree(mg->d_dest); cudaFree(mg->d_dest_sorted); cudaFree(mg->d_order_in); cudaFree(mg->d_order); cudaFree(mg->d_sort_tmp);
    cudaFree(mg->d_grav);
    delete mg;
    return B200SPH_OK;
}

extern "C" int b200sph_mg_create(b200sph_mg **out, b200sph_handle *h, int rank, int world, const void *id128)
{
    if (!out || !h || !id128 || world < 1 || world > HALO_MAX_RANKS || rank < 0 || rank >= world) return B200SPH_ERR_BAD_ARGUMENT;
    *out = nullptr;
    if (nccl_bind(h->err, sizeof(h->err))) return B200SPH_ERR_UNSUPPORTED;
    b200sph_mg *mg = new b200sph_mg();
    memset(&mg->stats, 0, sizeof(mg->stats));
    mg->h = h; mg->rank = rank; mg->world = world; mg->comm = nullptr;
    mg->dim = DIM;
    mg->level = (DIM == 3) ? 6 : (DIM == 2 ? 9 : 15);   /* cells of the decomposition: 2^level per axis */
    mg->d_cuts = nullptr; mg->d_hist = nullptr; mg->d_red = nullptr;
    mg->have_plan = 0; mg->plan_n_owned = -1; mg->plan_builds = 0; mg->stale_plans = 0;
    mg->n_send = mg->n_recv = 0;
    mg->d_idx = mg->d_counts = mg->d_all_counts = mg->d_send_counts = mg->d_recv_counts = mg->d_flag = nullptr;
    mg->idx_capacity = 0;
    mg->d_send = mg->d_recv = nullptr; mg->send_cap = mg->recv_cap = 0;
    for (int k = 0; k < 4; k++) mg->d_snap[k] = nullptr;
    mg->snap_cap = 0;
    mg->d_dest = mg->d_dest_sorted = mg->d_order_in = mg->d_order = nullptr;
    mg->d_sort_tmp = nullptr; mg->sort_tmp_bytes = 0; mg->mig_cap = 0;
    mg->d_grav = nullptr; mg->grav_cap = 0;
    mg->err[0] = 0;
    *out = mg;
    MCU(cudaSetDevice(h->device));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    MNC(g_nccl.CommInitRank(&mg->comm, world, id, rank));
    const size_t bins = (size_t)1 << (mg->dim * mg->level);
    MCU(cudaMalloc((void **)&mg->d_hist, bins * sizeof(unsigned long long)));
    MCU(cudaMalloc((void **)&mg->d_cuts, (world + 1) * sizeof(long long)));
    MCU(cudaMalloc((void **)&mg->d_red, 16 * sizeof(double)));
    MCU(cudaMalloc((void **)&mg->d_counts, (world + 2) * sizeof(int)));
    MCU(cudaMalloc((void **)&mg->d_all_counts, (size_t)world * (world + 2) * sizeof(int)));
    MCU(cudaMalloc((void **)&mg->d_send_counts, world * sizeof(int)));
    MCU(cudaMalloc((void **)&mg->d_recv_counts, world * sizeof(int)));
    MCU(cudaMalloc((void **)&mg->d_flag, sizeof(int)));
    MCU(cudaMemset(mg->d_flag, 0, sizeof(int)));
    return B200SPH_OK;
}

/* ------------------------------------------------------------------ decomposition */
static void mg_cell_box(const b200sph_mg *mg, long long cell, int level, double *box6)
{
    int q[3] = {0, 0, 0};
    for (int b = 0; b < level; b++)
        for (int a = 0; a < mg->dim; a++) q[a] |= (int)((cell >> (mg->dim * b + a)) & 1) << b;
    const double size = mg->cube_span / (double)(1 << level);
    for (int a = 0; a < 3; a++) {
        box6[a] = (a < mg->dim) ? mg->cube_lo[a] + q[a] * size : 0.0;
        box6[3 + a] = (a < mg->dim) ? mg->cube_lo[a] + (q[a] + 1) * size : 0.0;
    }
}

extern "C" int b200sph_mg_decompose(b200sph_mg *mg, const b200sph_view *view, int n_held, int weight_by_interactions)
{
    if (!mg || !view || n_held < 0) return B200SPH_ERR_BAD_ARGUMENT;
    b200sph_handle *h = mg->h;
    MCU(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const b200sph_particle_arrays &p = view->p;
    const int G = (n_held + MG_THREADS - 1) / MG_THREADS;
    /* global bounding cube */
    unsigned long long *d_ext = (unsigned long long *)mg->d_red;
    MCU(cudaMemsetAsync(d_ext, 0, 8 * sizeof(unsigned long long), st));
    if (n_held > 0) mg_extrema<<<G, MG_THREADS, 0, st>>>(p.x, DIM > 1 ? p.y : nullptr, DIM > 2 ? p.z : nullptr, p.h, n_held, d_ext);
    MNC(g_nccl.AllReduce(d_ext, d_ext, 8, ncclUint64, ncclMax, mg->comm, st));
    unsigned long long ext[8];
    MCU(cudaMemcpyAsync(ext, d_ext, sizeof(ext), cudaMemcpyDeviceToHost, st));
    MCU(cudaStreamSynchronize(st));
    double lo[3], hi[3], span = 0.0;
    for (int a = 0; a < 3; a++) {
        lo[a] = -mg_from_order_bits(ext[a]);
        hi[a] = mg_from_order_bits(ext[3 + a]);
        if (a < mg->dim) span = std::max(span, hi[a] - lo[a]);
    }
    span = span > 0.0 ? span * (1.0 + 1e-12) : 1.0;
    for (int a = 0; a < 3; a++) mg->cube_lo[a] = (a < mg->dim) ? 0.5 * (lo[a] + hi[a]) - 0.5 * span : 0.0;
    mg->cube_span = span;
    /* histogram over the Morton cells, weighted by 1 (+ noi: equal WORK instead of equal counts) */
    const size_t bins = (size_t)1 << (mg->dim * mg->level);
    MgCube cube;
    for (int a = 0; a < 3; a++) cube.lo[a] = mg->cube_lo[a];
    cube.span = span; cube.dim = mg->dim; cube.level = mg->level;
    MCU(cudaMemsetAsync(mg->d_hist, 0, bins * sizeof(unsigned long long), st));
    if (n_held > 0)
        mg_histogram<<<G, MG_THREADS, 0, st>>>(p.x, DIM > 1 ? p.y : nullptr, DIM > 2 ? p.z : nullptr,
                                               weight_by_interactions ? p.noi : nullptr, n_held, cube, mg->d_hist);
    MNC(g_nccl.AllReduce(mg->d_hist, mg->d_hist, bins, ncclUint64, ncclSum, mg->comm, st));
    std::vector<unsigned long long> hist(bins);
    MCU(cudaMemcpyAsync(hist.data(), mg->d_hist, bins * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    MCU(cudaStreamSynchronize(st));
    /* cuts: rank r gets cells [cuts[r], cuts[r+1]) with (nearly) equal weight */
    std::vector<unsigned long long> csum(bins + 1, 0ull);
    for (size_t c = 0; c < bins; c++) csum[c + 1] = csum[c] + hist[c];
    const double total = (double)csum[bins];
    mg->cuts.assign(mg->world + 1, 0);
    for (int r = 1; r < mg->world; r++) {
        const double target = total * r / mg->world;
        size_t c = std::lower_bound(csum.begin(), csum.end(), (unsigned long long)(target + 0.5)) - csum.begin();
        if (c > 0 && std::abs((double)csum[c - 1] - target) <= std::abs((double)csum[std::min(c, bins)] - target)) c--;
        mg->cuts[r] = std::min<long long>(std::max<long long>((long long)c, mg->cuts[r - 1]), (long long)bins);
    }
    mg->cuts[mg->world] = (long long)bins;
    MCU(cudaMemcpyAsync(mg->d_cuts, mg->cuts.data(), (mg->world + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
    /* every rank's cell range as a union of aligned octree boxes */
    std::vector<double> boxes;
    std::vector<int> box_rank;
    const long long fan = 1ll << mg->dim;
    for (int r = 0; r < mg->world; r++) {
        long long c = mg->cuts[r];
        const long long c1 = mg->cuts[r + 1];
        while (c < c1) {
            int up = 0;
            long long block = 1;
            while (up < mg->level && c % (block * fan) == 0 && c + block * fan <= c1) { block *= fan; up++; }
            double b6[6];
            mg_cell_box(mg, c / block, mg->level - up, b6);
            boxes.insert(boxes.end(), b6, b6 + 6);
            box_rank.push_back(r);
            c += block;
        }
    }
    MRC(b200sph_halo_set_domains(h, boxes.data(), box_rank.data(), (int)box_rank.size(), mg->world, mg->rank));
    mg->have_plan = 0;
    mg->stats.n_boxes = (int)box_rank.size();
    MCU(cudaStreamSynchronize(st));
    return B200SPH_OK;
}

/* ------------------------------------------------------------------ field lists */
static int mg_state_fields(const b200sph_view *v, b200sph_halo_field *f)
{
    /* what a neighbour contributes through: inputs of the pointwise chain and of the pair loops */
    const b200sph_particle_arrays &p = v->p, &r = v->p_rhs;
    int n = 0;
    auto add = [&](void *ptr, int per, int kind) { if (ptr) f[n++] = b200sph_halo_field{ptr, per, kind}; };
    add(p.x, 1, 0); add(p.y, 1, 0); add(p.z, 1, 0); add(p.vx, 1, 0); add(p.vy, 1, 0); add(p.vz, 1, 0);
    add(p.m, 1, 0); add(p.h, 1, 0); add(r.h0, 1, 0); add(p.rho, 1, 0); add(p.e, 1, 0); add(p.p, 1, 0); add(p.cs, 1, 0);
    add(r.materialId, 1, 1); add(p.S, DD, 0); add(p.d, 1, 0); add(p.damage_porjutzi, 1, 0); add(p.alpha_jutzi, 1, 0);
    add(p.numFlaws, 1, 1); add(p.numActiveFlaws, 1, 1);
    return n;
}

static int mg_all_fields(const b200sph_particle_arrays *sets, int n_sets, int max_num_flaws, b200sph_halo_field *f, int cap)
{
    int n = 0;
    auto add = [&](void *ptr, int per, int kind) {
        if (!ptr) return;
        for (int k = 0; k < n; k++)
            if (f[k].data == ptr) return;
        if (n < cap) f[n++] = b200sph_halo_field{ptr, per, kind};
    };
    for (int s = 0; s < n_sets; s++) {
        const b200sph_particle_arrays &a = sets[s];
        double *scalars[] = {a.x, a.y, a.z, a.vx, a.vy, a.vz, a.dxdt, a.dydt, a.dzdt, a.ax, a.ay, a.az, a.g_ax, a.g_ay, a.g_az,
                             a.g_local_cellsize, a.g_x, a.g_y, a.g_z, a.m, a.h, a.h0, a.dhdt, a.rho, a.drhodt, a.p, a.e, a.dedt,
                             a.local_strain, a.ep, a.edotp, a.plastic_f, a.d, a.damage_total, a.dddt, a.damage_porjutzi,
                             a.ddamage_porjutzidt, a.muijmax, a.pold, a.alpha_jutzi, a.alpha_jutzi_old, a.dalphadt, a.dalphadp,
                             a.dalphadrho, a.f, a.delpdelrho, a.delpdele, a.cs};
        for (double *ptr : scalars) add(ptr, 1, 0);
        double *tensors[] = {a.S, a.dSdt, a.sigma, a.R, a.tensorialCorrectionMatrix};
        for (double *ptr : tensors) add(ptr, DD, 0);
        add(a.flaws, max_num_flaws, 0);
        int *ints[] = {a.numFlaws, a.numActiveFlaws, a.noi, a.materialId, a.depth};
        for (int *ptr : ints) add(ptr, 1, 1);
    }
    return n;
}

static int mg_grow(b200sph_mg *mg, double **buf, size_t *cap, size_t need)
{
    if (*cap >= need) return 0;
    cudaFree(*buf);
    *buf = nullptr;
    *cap = 0;
    const size_t want = need + need / 4 + 1024;
    MCU(cudaMalloc((void **)buf, want * sizeof(double)));
    *cap = want;
    return 0;
}

/* rows of every rank's block travel column by column (b200sph_halo_pack_by_rank); one grouped send/recv per peer */
static int mg_exchange_rows(b200sph_mg *mg, const double *send, double *recv, const std::vector<int> &sc, const std::vector<int> &rc, int width)
{
    cudaStream_t st = mg->h->stream;
    size_t so = 0, ro = 0;
    MNC(g_nccl.GroupStart());
    for (int r = 0; r < mg->world; r++) {
        if (sc[r] > 0) MNC(g_nccl.Send(send + so, (size_t)sc[r] * width, ncclDouble, r, mg->comm, st));
        if (rc[r] > 0) MNC(g_nccl.Recv(recv + ro, (size_t)rc[r] * width, ncclDouble, r, mg->comm, st));
        so += (size_t)sc[r] * width;
        ro += (size_t)rc[r] * width;
    }
    MNC(g_nccl.GroupEnd());
    return 0;
}

/* ------------------------------------------------------------------ migration */
template <typename T>
__global__ void mg_gather_rows(T *dst, const T *src, const int *idx, int n, int per)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)n * per) return;
    const int k = (int)(e / per), c = (int)(e - (size_t)k * per);
    dst[e] = src[(size_t)idx[k] * per + c];
}

extern "C" int b200sph_mg_migrate(b200sph_mg *mg, const b200sph_view *view, const b200sph_particle_arrays *extra, int n_extra,
                                  int n_held, int capacity, int *n_held_out)
{
    if (!mg || !view || n_held < 0 || capacity < n_held || !n_held_out || n_extra < 0 || n_extra > 6) return B200SPH_ERR_BAD_ARGUMENT;
    if ((int)mg->cuts.size() != mg->world + 1) {
        snprintf(mg->err, sizeof(mg->err), "b200sph_mg_migrate: call b200sph_mg_decompose first");
        return B200SPH_ERR_BAD_ARGUMENT;
    }
    b200sph_handle *h = mg->h;
    MCU(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const int W = mg->world;
    if (mg->mig_cap < capacity) {
        cudaFree(mg->d_dest); cudaFree(mg->d_dest_sorted); cudaFree(mg->d_order_in); cudaFree(mg->d_order); cudaFree(mg->d_sort_tmp);
        mg->d_dest = mg->d_dest_sorted = mg->d_order_in = mg->d_order = nullptr;
        mg->d_sort_tmp = nullptr;
        MCU(cudaMalloc((void **)&mg->d_dest, capacity * sizeof(int)));
        MCU(cudaMalloc((void **)&mg->d_dest_sorted, capacity * sizeof(int)));
        MCU(cudaMalloc((void **)&mg->d_order_in, capacity * sizeof(int)));
        MCU(cudaMalloc((void **)&mg->d_order, capacity * sizeof(int)));
        mg->sort_tmp_bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, mg->sort_tmp_bytes, mg->d_dest, mg->d_dest_sorted, mg->d_order_in, mg->d_order, capacity, 0, 8);
        MCU(cudaMalloc(&mg->d_sort_tmp, mg->sort_tmp_bytes + 16));
        mg->mig_cap = capacity;
    }
    const b200sph_particle_arrays &p = view->p;
    MgCube cube;
    for (int a = 0; a < 3; a++) cube.lo[a] = mg->cube_lo[a];
    cube.span = mg->cube_span; cube.dim = mg->dim; cube.level = mg->level;
    /* destinations, then a stable sort by destination: migrants grouped by rank first, the stayers (ascending) last */
    MCU(cudaMemsetAsync(mg->d_counts, 0, (W + 2) * sizeof(int), st));
    if (n_held > 0) {
        mg_destinations<<<(n_held + MG_THREADS - 1) / MG_THREADS, MG_THREADS, 0, st>>>(
            p.x, DIM > 1 ? p.y : nullptr, DIM > 2 ? p.z : nullptr, n_held, cube, mg->d_cuts, W, mg->rank, mg->d_dest, mg->d_order_in,
            mg->d_counts);
        MCU(cub::DeviceRadixSort::SortPairs(mg->d_sort_tmp, mg->sort_tmp_bytes, mg->d_dest, mg->d_dest_sorted, mg->d_order_in, mg->d_order,
                                            n_held, 0, 8, st));
    }
    MNC(g_nccl.AllGather(mg->d_counts, mg->d_all_counts, W + 2, ncclInt32, mg->comm, st));
    std::vector<int> table((size_t)W * (W + 2));
    MCU(cudaMemcpyAsync(table.data(), mg->d_all_counts, table.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
    MCU(cudaStreamSynchronize(st));
    std::vector<int> sc(W), rc(W);
    int n_out = 0, n_in = 0;
    for (int r = 0; r < W; r++) {
        sc[r] = table[(size_t)mg->rank * (W + 2) + r];
        rc[r] = table[(size_t)r * (W + 2) + mg->rank];
        n_out += sc[r];
        n_in += rc[r];
    }
    const int n_stay = n_held - n_out;
    if (n_stay + n_in > capacity) {
        snprintf(mg->err, sizeof(mg->err), "b200sph_mg_migrate: %d staying + %d arriving particles exceed the capacity %d", n_stay, n_in, capacity);
        return B200SPH_ERR_BAD_ARGUMENT;
    }
    long long moved = n_out;
    /* every member of p, p_rhs and the extra buffers (rk_device[3]) travels */
    std::vector<b200sph_particle_arrays> sets;
    sets.push_back(view->p);
    sets.push_back(view->p_rhs);
    for (int k = 0; k < n_extra; k++) sets.push_back(extra[k]);
    b200sph_halo_field fields[6 * 64];
    const int nf = mg_all_fields(sets.data(), (int)sets.size(), view->max_num_flaws, fields, 6 * 64);
    MCU(cudaMemcpyAsync(mg->d_send_counts, sc.data(), W * sizeof(int), cudaMemcpyHostToDevice, st));
    MCU(cudaMemcpyAsync(mg->d_recv_counts, rc.data(), W * sizeof(int), cudaMemcpyHostToDevice, st));
    /* the members travel in groups of at most MG_FIELD_GROUP (the pack kernels take their member table by value);
     * every group is packed and sent before anything is overwritten, its arrivals wait in their own buffer */
    const int MG_FIELD_GROUP = 64;
    const int n_groups = (nf + MG_FIELD_GROUP - 1) / MG_FIELD_GROUP;
    std::vector<int> g_width(n_groups), g_first(n_groups), g_count(n_groups);
    size_t recv_total = 0, send_max = 0;
    for (int g = 0; g < n_groups; g++) {
        g_first[g] = g * MG_FIELD_GROUP;
        g_count[g] = std::min(MG_FIELD_GROUP, nf - g_first[g]);
        g_width[g] = b200sph_halo_row_width(fields + g_first[g], g_count[g]);
        recv_total += (size_t)n_in * g_width[g];
        send_max = std::max(send_max, (size_t)n_out * g_width[g]);
    }
    if (mg_grow(mg, &mg->d_send, &mg->send_cap, send_max)) return B200SPH_ERR_CUDA;
    if (mg_grow(mg, &mg->d_recv, &mg->recv_cap, recv_total)) return B200SPH_ERR_CUDA;
    {
        size_t roff = 0;
        for (int g = 0; g < n_groups; g++) {
            if (n_out > 0)
                MRC(b200sph_halo_pack_by_rank(h, fields + g_first[g], g_count[g], mg->d_order, mg->d_send_counts, W, n_out, mg->d_send));
            if (mg_exchange_rows(mg, mg->d_send, mg->d_recv + roff, sc, rc, g_width[g])) return B200SPH_ERR_CUDA;
            roff += (size_t)n_in * g_width[g];
        }
    }
    /* compact the stayers (their ascending indices are the tail of the sorted order), member by member through the
     * neighbour-list storage (void after a migration anyway), then append the arrivals */
    if (n_out > 0 && n_stay > 0) {
        const int *stay = mg->d_order + n_out;
        for (int k = 0; k < nf; k++) {
            const size_t elems = (size_t)n_stay * fields[k].per;
            const int blocks = (int)((elems + 255) / 256);
            if (fields[k].kind == 0) {
                mg_gather_rows<double><<<blocks, 256, 0, st>>>((double *)h->s.nbr, (const double *)fields[k].data, stay, n_stay, fields[k].per);
                MCU(cudaMemcpyAsync(fields[k].data, h->s.nbr, elems * sizeof(double), cudaMemcpyDeviceToDevice, st));
            } else {
                mg_gather_rows<int><<<blocks, 256, 0, st>>>((int *)h->s.nbr, (const int *)fields[k].data, stay, n_stay, fields[k].per);
                MCU(cudaMemcpyAsync(fields[k].data, h->s.nbr, elems * sizeof(int), cudaMemcpyDeviceToDevice, st));
            }
        }
    }
    if (n_in > 0) {
        size_t roff = 0;
        for (int g = 0; g < n_groups; g++) {
            MRC(b200sph_halo_unpack_by_rank(h, fields + g_first[g], g_count[g], mg->d_recv + roff, mg->d_recv_counts, W, n_in, n_stay));
            roff += (size_t)n_in * g_width[g];
        }
    }
    MCU(cudaStreamSynchronize(st));
    MCU(cudaGetLastError());
    *n_held_out = n_stay + n_in;
    mg->have_plan = 0;
    mg->stats.migrated_out = moved;
    mg->stats.migrated_in = n_in;
    h->s.n = 0;
    return B200SPH_OK;
}

/* ------------------------------------------------------------------ halo plan */
static int mg_build_plan(b200sph_mg *mg, const b200sph_view *view, int n_owned, int capacity, int h_evolves)
{
    b200sph_handle *h = mg->h;
    cudaStream_t st = h->stream;
    const int W = mg->world;
    const b200sph_particle_arrays &p = view->p;
    mg->growth = h_evolves ? 0.02 : 0.0;
    /* D = 0.15 x the smallest smoothing length anywhere: the same number on every rank */
    unsigned long long *d_ext = (unsigned long long *)mg->d_red;
    MCU(cudaMemsetAsync(d_ext, 0, 8 * sizeof(unsigned long long), st));
    if (n_owned > 0)
        mg_extrema<<<(n_owned + MG_THREADS - 1) / MG_THREADS, MG_THREADS, 0, st>>>(p.x, DIM > 1 ? p.y : nullptr, DIM > 2 ? p.z : nullptr, p.h,
                                                                                    n_owned, d_ext);
    MNC(g_nccl.AllReduce(d_ext, d_ext, 8, ncclUint64, ncclMax, mg->comm, st));
    unsigned long long ext[8];
    MCU(cudaMemcpyAsync(ext, d_ext, sizeof(ext), cudaMemcpyDeviceToHost, st));
    MCU(cudaStreamSynchronize(st));
    mg->max_move = 0.15 * (-mg_from_order_bits(ext[6]));
    if (mg->idx_capacity < capacity) {
        cudaFree(mg->d_idx);
        mg->d_idx = nullptr;
        MCU(cudaMalloc((void **)&mg->d_idx, (size_t)capacity * sizeof(int)));
        mg->idx_capacity = capacity;
    }
    /* one halo level: the neighbour sums of the copies are delivered by their owners (b200sph_rhs_eval_stage) */
    MRC(b200sph_halo_select_plan(h, p.x, DIM > 1 ? p.y : nullptr, DIM > 2 ? p.z : nullptr, p.h, n_owned, nullptr, 0, 1.0 + mg->growth,
                                 2.0 * mg->max_move, mg->d_idx, mg->idx_capacity, mg->d_counts));
    MNC(g_nccl.AllGather(mg->d_counts, mg->d_all_counts, W + 1, ncclInt32, mg->comm, st));
    std::vector<int> table((size_t)W * (W + 1));
    MCU(cudaMemcpyAsync(table.data(), mg->d_all_counts, table.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
    MCU(cudaStreamSynchronize(st));
    mg->send_counts.assign(W, 0);
    mg->recv_counts.assign(W, 0);
    mg->n_send = mg->n_recv = 0;
    for (int r = 0; r < W; r++) {
        if (table[(size_t)r * (W + 1) + W]) {
            snprintf(mg->err, sizeof(mg->err), "halo send list of rank %d does not fit its index buffer", r);
            return B200SPH_ERR_BAD_ARGUMENT;
        }
        mg->send_counts[r] = table[(size_t)mg->rank * (W + 1) + r];
        mg->recv_counts[r] = table[(size_t)r * (W + 1) + mg->rank];
        mg->n_send += mg->send_counts[r];
        mg->n_recv += mg->recv_counts[r];
    }
    if (n_owned + mg->n_recv > capacity) {
        snprintf(mg->err, sizeof(mg->err), "halo of %d particles does not fit: capacity %d, owned %d", mg->n_recv, capacity, n_owned);
        return B200SPH_ERR_BAD_ARGUMENT;
    }
    MCU(cudaMemcpyAsync(mg->d_send_counts, mg->send_counts.data(), W * sizeof(int), cudaMemcpyHostToDevice, st));
    MCU(cudaMemcpyAsync(mg->d_recv_counts, mg->recv_counts.data(), W * sizeof(int), cudaMemcpyHostToDevice, st));
    if (mg->snap_cap < n_owned) {
        for (int k = 0; k < 4; k++) {
            cudaFree(mg->d_snap[k]);
            mg->d_snap[k] = nullptr;
            MCU(cudaMalloc((void **)&mg->d_snap[k], (size_t)(n_owned + n_owned / 8 + 64) * sizeof(double)));
        }
        mg->snap_cap = n_owned + n_owned / 8 + 64;
    }
    const double *src[4] = {p.x, DIM > 1 ? p.y : nullptr, DIM > 2 ? p.z : nullptr, p.h};
    for (int k = 0; k < 4; k++)
        if (src[k]) MCU(cudaMemcpyAsync(mg->d_snap[k], src[k], (size_t)n_owned * sizeof(double), cudaMemcpyDeviceToDevice, st));
    MCU(cudaMemsetAsync(mg->d_flag, 0, sizeof(int), st));
    mg->have_plan = 1;
    mg->plan_n_owned = n_owned;
    mg->plan_builds++;
    return 0;
}

static int mg_move(b200sph_mg *mg, const b200sph_halo_field *fields, int nf, int n_owned)
{
    b200sph_handle *h = mg->h;
    const int width = b200sph_halo_row_width(fields, nf);
    if (mg_grow(mg, &mg->d_send, &mg->send_cap, (size_t)mg->n_send * width)) return B200SPH_ERR_CUDA;
    if (mg_grow(mg, &mg->d_recv, &mg->recv_cap, (size_t)mg->n_recv * width)) return B200SPH_ERR_CUDA;
    if (mg->n_send > 0) MRC(b200sph_halo_pack_by_rank(h, fields, nf, mg->d_idx, mg->d_send_counts, mg->world, mg->n_send, mg->d_send));
    if (mg_exchange_rows(mg, mg->d_send, mg->d_recv, mg->send_counts, mg->recv_counts, width)) return B200SPH_ERR_CUDA;
    if (mg->n_recv > 0) MRC(b200sph_halo_unpack_by_rank(h, fields, nf, mg->d_recv, mg->d_recv_counts, mg->world, mg->n_recv, n_owned));
    mg->stats.halo_bytes_sent += (long long)mg->n_send * width * 8;
    return 0;
}

/* x, y, z, m of every rank's owned particles, rank after rank (replicated gravity tree, b200sph_set_gravity_sources) */
static int mg_gravity_sources(b200sph_mg *mg, const b200sph_view *view, int n_owned)
{
    b200sph_handle *h = mg->h;
    cudaStream_t st = h->stream;
    const int W = mg->world;
    int *d_n = mg->d_counts;
    MCU(cudaMemcpyAsync(d_n, &n_owned, sizeof(int), cudaMemcpyHostToDevice, st));
    MNC(g_nccl.AllGather(d_n, mg->d_all_counts, 1, ncclInt32, mg->comm, st));
    std::vector<int> counts(W);
    MCU(cudaMemcpyAsync(counts.data(), mg->d_all_counts, W * sizeof(int), cudaMemcpyDeviceToHost, st));
    MCU(cudaStreamSynchronize(st));
    int c_max = 0, n_total = 0, own_begin = 0;
    for (int r = 0; r < W; r++) {
        c_max = std::max(c_max, counts[r]);
        if (r < mg->rank) own_begin += counts[r];
        n_total += counts[r];
    }
    /* layout: 4 padded pieces of c_max per rank (all-gather wants equal sizes), then 4 compact arrays of n_total */
    const size_t padded = (size_t)4 * c_max, need = padded * (W + 1) + (size_t)4 * n_total;
    if (mg_grow(mg, &mg->d_grav, &mg->grav_cap, need)) return B200SPH_ERR_CUDA;
    double *mine = mg->d_grav, *all = mg->d_grav + padded, *flat = all + padded * W;
    const double *src[4] = {view->p.x, DIM > 1 ? view->p.y : nullptr, DIM > 2 ? view->p.z : nullptr, view->p.m};
    for (int k = 0; k < 4; k++) {
        if (src[k]) MCU(cudaMemcpyAsync(mine + (size_t)k * c_max, src[k], (size_t)n_owned * sizeof(double), cudaMemcpyDeviceToDevice, st));
        else MCU(cudaMemsetAsync(mine + (size_t)k * c_max, 0, (size_t)n_owned * sizeof(double), st));
    }
    MNC(g_nccl.AllGather(mine, all, padded, ncclDouble, mg->comm, st));
    size_t off = 0;
    for (int r = 0; r < W; r++) {
        for (int k = 0; k < 4; k++)
            MCU(cudaMemcpyAsync(flat + (size_t)k * n_total + off, all + padded * r + (size_t)k * c_max, (size_t)counts[r] * sizeof(double),
                                cudaMemcpyDeviceToDevice, st));
        off += counts[r];
    }
    MRC(b200sph_set_gravity_sources(h, flat, DIM > 1 ? flat + n_total : nullptr, DIM > 2 ? flat + 2 * (size_t)n_total : nullptr,
                                    flat + 3 * (size_t)n_total, n_total, own_begin));
    return 0;
}

/* The right-hand side of this rank's n_owned particles (rows [0, n_owned) of `view`, whose arrays have room for
 * `capacity` rows: the halo copies go behind the owned rows).  Rates are produced for the owned rows. */
extern "C" int b200sph_mg_rhs_eval(b200sph_mg *mg, const b200sph_view *view, int n_owned, int capacity, int *n_total_out, int *offender)
{
    if (!mg || !view || n_owned <= 0 || capacity < n_owned) return B200SPH_ERR_BAD_ARGUMENT;
    b200sph_handle *h = mg->h;
    MCU(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    b200sph_halo_field state[32], sum_rho[1], sum_c[1];
    const int n_state = mg_state_fields(view, state);
    const int h_evolves = (VARIABLE_SML || INTEGRATE_SML) ? 1 : 0;
    const b200sph_particle_arrays &p = view->p;
    sum_rho[0] = b200sph_halo_field{p.rho, 1, 0};
    sum_c[0] = b200sph_halo_field{view->p_rhs.tensorialCorrectionMatrix, DD, 0};

    for (int attempt = 0; attempt < 2; attempt++) {
        if (mg->world > 1) {
            if (!mg->have_plan || mg->plan_n_owned != n_owned || attempt > 0) {
                if (mg_build_plan(mg, view, n_owned, capacity, h_evolves)) return B200SPH_ERR_BAD_ARGUMENT;
            } else {
                /* is the plan still good?  the all-reduced verdict stays on the device: it is the evaluation's abort flag */
                MRC(b200sph_halo_plan_check(h, p.x, DIM > 1 ? p.y : nullptr, DIM > 2 ? p.z : nullptr, p.h, mg->d_snap[0], mg->d_snap[1],
                                            mg->d_snap[2], mg->d_snap[3], n_owned, mg->max_move, mg->growth, mg->d_flag));
                MNC(g_nccl.AllReduce(mg->d_flag, mg->d_flag, 1, ncclInt32, ncclMax, mg->comm, st));
            }
            if (mg_move(mg, state, n_state, n_owned)) return B200SPH_ERR_CUDA;
            if (view->selfgravity && mg_gravity_sources(mg, view, n_owned)) return B200SPH_ERR_CUDA;
        }
        const int n_total = n_owned + ((mg->world > 1) ? mg->n_recv : 0);
        b200sph_view v = *view;
        v.n = n_total;
        v.n_real = n_total;
        MRC(b200sph_set_owned(h, n_owned));
        MRC(b200sph_set_halo_sums(h, mg->world > 1));
        MRC(b200sph_set_abort_flag(h, mg->world > 1 ? mg->d_flag : nullptr));
        int rc = 0;
        for (int stage = 0; stage < 3 && rc == 0; stage++) {
            int pending = 0;
            rc = b200sph_rhs_eval_stage(h, &v, stage, &pending, offender);
            if (rc == 0 && pending == B200SPH_SUM_DENSITY && mg_move(mg, sum_rho, 1, n_owned)) return B200SPH_ERR_CUDA;
            if (rc == 0 && pending == B200SPH_SUM_CORRECTION && mg_move(mg, sum_c, 1, n_owned)) return B200SPH_ERR_CUDA;
            if (rc == 0 && pending) mg->stats.sum_exchanges++;
        }
        if (n_total_out) *n_total_out = n_total;
        mg->stats.n_halo = n_total - n_owned;
        mg->stats.plan_builds = mg->plan_builds;
        mg->stats.stale_plans = mg->stale_plans;
        if (rc == B200SPH_ERR_ABORTED && attempt == 0) {
            mg->stale_plans++;   /* every rank saw the same all-reduced flag: all of them come back here */
            continue;
        }
        if (rc) snprintf(mg->err, sizeof(mg->err), "%s", b200sph_last_error(h));
        return rc;
    }
    return B200SPH_ERR_ABORTED;
}

/* ------------------------------------------------------------------ rk2_adaptive over several GPUs
 * b200sph_rk2_step / b200sph_rk2_advance (integrate.cu) with the right-hand side going through b200sph_mg_rhs_eval and
 * the step-size reductions all-reduced over the ranks ("global reductions moved out of single-GPU kernels":
 * limitTimestepCourant/Damage -> min, checkError -> max).  Every rank takes the same steps.  All arrays of `view` and of
 * the rk buffers have room for `capacity` rows; view->n is ignored in favour of n_owned. */
struct MgRkCtx {
    b200sph_mg *mg;
    int n_owned, capacity;
};

static int mg_rk_rhs(void *ctx, const b200sph_view *bound, int *offender)
{
    MgRkCtx *c = (MgRkCtx *)ctx;
    return b200sph_mg_rhs_eval(c->mg, bound, c->n_owned, c->capacity, nullptr, offender);
}

static int mg_rk_allreduce(void *ctx, double *dev_values, int n, int is_min)
{
    b200sph_mg *mg = ((MgRkCtx *)ctx)->mg;
    MNC(g_nccl.AllReduce(dev_values, dev_values, n, ncclDouble, is_min ? ncclMin : ncclMax, mg->comm, mg->h->stream));
    return 0;
}

extern "C" int b200sph_mg_rk2_advance(b200sph_mg *mg, const b200sph_view *view, const b200sph_particle_arrays rk[3],
                                      const b200sph_rk2_params *prm, double t_end, b200sph_rk2_state *state, int n_owned, int capacity,
                                      int *offender)
{
    if (!mg || !view || !rk || !prm || !state || n_owned <= 0 || capacity < n_owned) return B200SPH_ERR_BAD_ARGUMENT;
    b200sph_handle *h = mg->h;
    MgRkCtx ctx = {mg, n_owned, capacity};
    b200sph_view v = *view;
    v.n = n_owned;
    v.n_real = n_owned;
    h->rk_rhs_hook = mg_rk_rhs;
    h->rk_allreduce = (mg->world > 1) ? mg_rk_allreduce : nullptr;
    h->rk_hook_ctx = &ctx;
    /* the cold calls of the output path (damageLimit) see owned rows only */
    MRC(b200sph_set_owned(h, 0));
    const int rc = b200sph_rk2_advance(h, &v, rk, prm, t_end, state, offender);
    h->rk_rhs_hook = nullptr;
    h->rk_allreduce = nullptr;
    h->rk_hook_ctx = nullptr;
    if (rc && !mg->err[0]) snprintf(mg->err, sizeof(mg->err), "%s", b200sph_last_error(h));
    return rc;
}

extern "C" int b200sph_mg_get_stats(const b200sph_mg *mg, b200sph_mg_stats *out)
{
    if (!mg || !out) return B200SPH_ERR_BAD_ARGUMENT;
    *out = mg->stats;
    return B200SPH_OK;
}
