/*
 * integrate.cu -- the embedded Runge-Kutta 2/3 integrator with adaptive step size on the device.
 *
 * Replaces, for the switch sets in scope, what miluphcuda's rk2Adaptive() does around its three rightHandSide() calls
 * per step (reference: src/rk2adaptive.cu:197-349 the step, :521-695 limitTimestep*, :700-1130 integrate*Step,
 * :1134-1482 checkError; src/memory_handling.cu:253-522 the device-to-device copies between p_device and rk_device[3]):
 *
 *   reference per accepted step                           here
 *   ~25 cudaMemcpy D2D  p -> rk[FIRST]                    k_rk_copy_vars            1 launch
 *   limitTimestepCourant/Forces/Damage (3 launches,       k_rk_limit_remember       1 launch: the three minima AND the
 *     one block does the final min each) + ~45 cudaMemcpy                            copy rk[FIRST] -> rk[START]
 *     D2D  rk[FIRST] -> rk[START]
 *   integrateFirstStep / SecondStep                       k_rk_first / k_rk_second  1 launch each
 *   integrateThirdStep + checkError (2 launches, final    k_rk_third_check          1 launch: update, error norms, new dt
 *     max by one block) + 6 cudaMemcpyFromSymbol
 *   ~45 cudaMemcpy D2D on a rejected step                 k_rk_restore              1 launch
 *
 * All kernels are streaming passes over the caller's buffers (HBM-bound); every buffer keeps the reference's
 * meaning: rk[0] = RKSTART, rk[1] = RKFIRST, rk[2] = RKSECOND (include/timeintegration.h:129-131), `p` = p_device.
 * The RK2_* switches of include/rk2adaptive.h:39-71 are runtime fields of b200sph_rk2_params with the shipped values.
 */
#include "rhs_internal.h"

#include <float.h>
#include <math.h>
#include <stdio.h>

#define RK_B21 0.5
#define RK_B31 (-1.0)
#define RK_B32 2.0
#define RK_C1 1.0
#define RK_C2 4.0
#define RK_C3 1.0
#define RK_THREADS 256
#define RK_NRED 8

struct RkScalars {          /* device-resident step state, read back by the host after the reducing kernels */
    double dt;              /* current step size (the reference's __device__ dt) */
    double dt_new;          /* dtNewErrorCheck */
    int error_small_enough;
    int pad;
    double err[RK_NRED];    /* 0 position, 1 velocity, 2 density, 3 energy, 4 alpha change, 5 pressure change */
    double limit[4];        /* 0 Courant, 1 forces, 2 damage (1e100 when not limiting) */
};

struct RkBuffers {
    b200sph_particle_arrays p, start, first, second;
};

__device__ __forceinline__ double rk_warp_min(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double rk_warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

/* block reduction of NV values (min when is_min, else max) + last-block combine; returns true in thread 0 of the
 * last block, with the combined values in out[] */
template <int NV>
__device__ __forceinline__ bool rk_reduce(double (&vals)[NV], bool is_min, double *partials, unsigned int *counter, double (&out)[NV])
{
    __shared__ double sh[RK_THREADS / 32][NV];
    __shared__ bool last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; k++) vals[k] = is_min ? rk_warp_min(vals[k]) : rk_warp_max(vals[k]);
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < NV; k++) sh[warp][k] = vals[k];
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < RK_THREADS / 32; w++)
#pragma unroll
            for (int k = 0; k < NV; k++) sh[0][k] = is_min ? fmin(sh[0][k], sh[w][k]) : fmax(sh[0][k], sh[w][k]);
#pragma unroll
        for (int k = 0; k < NV; k++) partials[blockIdx.x * NV + k] = sh[0][k];
        __threadfence();
        last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return false;
    __threadfence();
    double r[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) r[k] = is_min ? 1e300 : 0.0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x)
#pragma unroll
        for (int k = 0; k < NV; k++) {
            const double q = __ldcg(partials + b * NV + k);
            r[k] = is_min ? fmin(r[k], q) : fmax(r[k], q);
        }
#pragma unroll
    for (int k = 0; k < NV; k++) r[k] = is_min ? rk_warp_min(r[k]) : rk_warp_max(r[k]);
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < NV; k++) sh[warp][k] = r[k];
    __syncthreads();
    if (threadIdx.x != 0) return false;
#pragma unroll
    for (int k = 0; k < NV; k++) {
        out[k] = sh[0][k];
        for (int w = 1; w < RK_THREADS / 32; w++) out[k] = is_min ? fmin(out[k], sh[w][k]) : fmax(out[k], sh[w][k]);
    }
    *counter = 0;
    return true;
}

/* ---- the field groups of src/memory_handling.cu:291-522 as per-particle copies ---- */
__device__ __forceinline__ void rk_copy_variables(const b200sph_particle_arrays &d, const b200sph_particle_arrays &s, int i)
{
    d.x[i] = s.x[i]; d.vx[i] = s.vx[i];
#if DIM > 1
    d.y[i] = s.y[i]; d.vy[i] = s.vy[i];
#endif
#if DIM > 2
    d.z[i] = s.z[i]; d.vz[i] = s.vz[i];
#endif
    d.rho[i] = s.rho[i];
    d.h[i] = s.h[i];
#if INTEGRATE_ENERGY
    d.e[i] = s.e[i];
#endif
#if PALPHA_POROSITY
    d.alpha_jutzi[i] = s.alpha_jutzi[i];
    d.alpha_jutzi_old[i] = s.alpha_jutzi[i];   /* sic: the old value is the source's CURRENT one (memory_handling.cu:445) */
    d.dalphadp[i] = s.dalphadp[i]; d.dalphadrho[i] = s.dalphadrho[i];
    d.delpdelrho[i] = s.delpdelrho[i]; d.delpdele[i] = s.delpdele[i];
    d.f[i] = s.f[i]; d.p[i] = s.p[i]; d.pold[i] = s.pold[i];
#if FRAGMENTATION
    d.damage_porjutzi[i] = s.damage_porjutzi[i];
#endif
#endif
#if SOLID
#pragma unroll
    for (int c = 0; c < DD; c++) d.S[(size_t)i * DD + c] = s.S[(size_t)i * DD + c];
    d.ep[i] = s.ep[i];
#endif
#if FRAGMENTATION
    d.d[i] = s.d[i];
    d.damage_total[i] = s.damage_total[i];
    d.numActiveFlaws[i] = s.numActiveFlaws[i];
#endif
}

__device__ __forceinline__ void rk_copy_derivatives(const b200sph_particle_arrays &d, const b200sph_particle_arrays &s, int i)
{
    d.ax[i] = s.ax[i]; d.dxdt[i] = s.dxdt[i];
    if (d.g_ax && s.g_ax) d.g_ax[i] = s.g_ax[i];
#if DIM > 1
    d.ay[i] = s.ay[i]; d.dydt[i] = s.dydt[i];
    if (d.g_ay && s.g_ay) d.g_ay[i] = s.g_ay[i];
#endif
#if DIM > 2
    d.az[i] = s.az[i]; d.dzdt[i] = s.dzdt[i];
    if (d.g_az && s.g_az) d.g_az[i] = s.g_az[i];
#endif
    d.drhodt[i] = s.drhodt[i];
#if INTEGRATE_SML
    d.dhdt[i] = s.dhdt[i];
#endif
#if PALPHA_POROSITY
    d.dalphadt[i] = s.dalphadt[i];
#if FRAGMENTATION
    d.ddamage_porjutzidt[i] = s.ddamage_porjutzidt[i];
#endif
#endif
#if INTEGRATE_ENERGY
    d.dedt[i] = s.dedt[i];
#endif
#if SOLID
#pragma unroll
    for (int c = 0; c < DD; c++) d.dSdt[(size_t)i * DD + c] = s.dSdt[(size_t)i * DD + c];
    d.edotp[i] = s.edotp[i];
#endif
#if FRAGMENTATION
    d.dddt[i] = s.dddt[i];
    d.numActiveFlaws[i] = s.numActiveFlaws[i];
#endif
}

/* copy_particles_immutables_device_to_device, src/memory_handling.cu:373-392 (once, when the buffers are set up) */
__global__ void k_rk_init(RkBuffers b, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    b200sph_particle_arrays *dst[3] = {&b.start, &b.first, &b.second};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        dst[k]->m[i] = b.p.m[i];
        dst[k]->h[i] = b.p.h[i];
        dst[k]->cs[i] = b.p.cs[i];
#if FRAGMENTATION
        dst[k]->numFlaws[i] = b.p.numFlaws[i];
#endif
    }
}

__global__ void k_rk_copy_vars(b200sph_particle_arrays dst, b200sph_particle_arrays src, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rk_copy_variables(dst, src, i);
}

/* limitTimestepCourant / Forces / Damage (src/rk2adaptive.cu:521-695) on the buffer the first right-hand side was
 * evaluated in (rk[FIRST]), fused with "remember values of first step" (rk[START] <- rk[FIRST], variables and
 * derivatives, src/rk2adaptive.cu:262-271) */
__global__ void __launch_bounds__(RK_THREADS)
k_rk_limit_remember(RkBuffers b, int n, int use_courant, int use_forces, int use_damage, double max_damage_change, RkScalars *sc,
                    double *partials, unsigned int *counter)
{
    double v[3] = {1e100, 1e100, 1e100};
    const b200sph_particle_arrays &q = b.first;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (use_courant && q.noi[i] > 0) v[0] = fmin(v[0], q.h[i] / q.cs[i]);
        if (use_forces) {
            double t = q.ax[i] * q.ax[i];
#if DIM > 1
            t += q.ay[i] * q.ay[i];
#endif
#if DIM > 2
            t += q.az[i] * q.az[i];
#endif
            if (t > 0.0) v[1] = fmin(v[1], sqrt(q.h[i] / sqrt(t)));
        }
#if FRAGMENTATION
        if (use_damage && q.dddt[i] > 0.0) {
            double t = 0.7 * (q.d[i] + max_damage_change) / q.dddt[i];
            t = fmin(t, max_damage_change / q.dddt[i]);
            v[2] = fmin(t, v[2]);
        }
#endif
        rk_copy_variables(b.start, q, i);
        rk_copy_derivatives(b.start, q, i);
    }
    double out[3];
    if (!rk_reduce<3>(v, true, partials, counter, out)) return;
    /* this rank's minima; several GPUs all-reduce them (min) before k_rk_apply_limits */
    sc->limit[0] = out[0]; sc->limit[1] = out[1]; sc->limit[2] = out[2];
}

__global__ void k_rk_apply_limits(RkScalars *sc, int use_courant, int use_forces, int use_damage, double courant_fact, double forces_fact)
{
    double dt = sc->dt;
    const double c = sc->limit[0] * courant_fact, f = sc->limit[1] * forces_fact, d = sc->limit[2];
    if (use_courant && c < dt && c > 0.0) dt = c;
    if (use_forces && f < dt && f > 0.0) dt = f;
    if (use_damage && d < dt && d > 0.0) dt = d;
    sc->dt = dt;
    sc->limit[0] = c; sc->limit[1] = f;
}

/* integrateFirstStep, src/rk2adaptive.cu:700-842: rk[FIRST] = rk[START] + dt B21 k1 */
__global__ void k_rk_first(RkBuffers b, int n, const RkScalars *sc)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double dt = sc->dt;
    const b200sph_particle_arrays &s = b.start, &f = b.first;
#if INTEGRATE_DENSITY
    f.rho[i] = s.rho[i] + dt * RK_B21 * s.drhodt[i];
#endif
#if INTEGRATE_SML
    f.h[i] = s.h[i] + dt * RK_B21 * s.dhdt[i];
#else
    f.h[i] = s.h[i];
#endif
#if INTEGRATE_ENERGY
    f.e[i] = s.e[i] + dt * RK_B21 * s.dedt[i];
#endif
#if FRAGMENTATION
    f.d[i] = s.d[i] + dt * RK_B21 * s.dddt[i];
    f.numActiveFlaws[i] = s.numActiveFlaws[i];
#if PALPHA_POROSITY
    f.damage_porjutzi[i] = s.damage_porjutzi[i] + dt * RK_B21 * s.ddamage_porjutzidt[i];
#endif
#endif
#if SOLID
#pragma unroll
    for (int c = 0; c < DD; c++) f.S[(size_t)i * DD + c] = s.S[(size_t)i * DD + c] + dt * RK_B21 * s.dSdt[(size_t)i * DD + c];
    f.ep[i] = s.ep[i] + dt * RK_B21 * s.edotp[i];
#endif
#if PALPHA_POROSITY
    f.alpha_jutzi[i] = s.alpha_jutzi[i] + dt * RK_B21 * s.dalphadt[i];
    f.pold[i] = f.p[i];   /* pressure at the begin of the step, compared with the one at its end */
#endif
    f.x[i] = s.x[i] + dt * RK_B21 * s.dxdt[i];
    f.vx[i] = s.vx[i] + dt * RK_B21 * s.ax[i];
#if DIM > 1
    f.y[i] = s.y[i] + dt * RK_B21 * s.dydt[i];
    f.vy[i] = s.vy[i] + dt * RK_B21 * s.ay[i];
#endif
#if DIM > 2
    f.z[i] = s.z[i] + dt * RK_B21 * s.dzdt[i];
    f.vz[i] = s.vz[i] + dt * RK_B21 * s.az[i];
#endif
}

/* integrateSecondStep, src/rk2adaptive.cu:845-950: rk[SECOND] = rk[START] + dt (B31 k1 + B32 k2); with self-gravity
 * the stored g_a of rk[FIRST] goes along (src/rk2adaptive.cu:300-302: needed when the walk is skipped under -g) */
__global__ void k_rk_second(RkBuffers b, int n, const RkScalars *sc, int copy_gravity)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double dt = sc->dt;
    const b200sph_particle_arrays &s = b.start, &f = b.first, &d = b.second;
#if INTEGRATE_DENSITY
    d.rho[i] = s.rho[i] + dt * (RK_B31 * s.drhodt[i] + RK_B32 * f.drhodt[i]);
#endif
#if INTEGRATE_SML
    d.h[i] = s.h[i] + dt * (RK_B31 * s.dhdt[i] + RK_B32 * f.dhdt[i]);
#else
    d.h[i] = s.h[i];
#endif
#if INTEGRATE_ENERGY
    d.e[i] = s.e[i] + dt * (RK_B31 * s.dedt[i] + RK_B32 * f.dedt[i]);
#endif
#if FRAGMENTATION
    d.d[i] = s.d[i] + dt * (RK_B31 * s.dddt[i] + RK_B32 * f.dddt[i]);
    d.numActiveFlaws[i] = f.numActiveFlaws[i];
#if PALPHA_POROSITY
    d.damage_porjutzi[i] = s.damage_porjutzi[i] + dt * (RK_B31 * s.ddamage_porjutzidt[i] + RK_B32 * f.ddamage_porjutzidt[i]);
#endif
#endif
#if PALPHA_POROSITY
    d.alpha_jutzi[i] = s.alpha_jutzi[i] + dt * (RK_B31 * s.dalphadt[i] + RK_B32 * f.dalphadt[i]);
    d.pold[i] = f.pold[i];
#endif
#if SOLID
#pragma unroll
    for (int c = 0; c < DD; c++)
        d.S[(size_t)i * DD + c] = s.S[(size_t)i * DD + c] + dt * (RK_B31 * s.dSdt[(size_t)i * DD + c] + RK_B32 * f.dSdt[(size_t)i * DD + c]);
    d.ep[i] = s.ep[i] + dt * (RK_B31 * s.edotp[i] + RK_B32 * f.edotp[i]);
#endif
    d.vx[i] = s.vx[i] + dt * (RK_B31 * s.ax[i] + RK_B32 * f.ax[i]);
    d.x[i] = s.x[i] + dt * (RK_B31 * s.dxdt[i] + RK_B32 * f.dxdt[i]);
    if (copy_gravity && d.g_ax) d.g_ax[i] = f.g_ax[i];
#if DIM > 1
    d.vy[i] = s.vy[i] + dt * (RK_B31 * s.ay[i] + RK_B32 * f.ay[i]);
    d.y[i] = s.y[i] + dt * (RK_B31 * s.dydt[i] + RK_B32 * f.dydt[i]);
    if (copy_gravity && d.g_ay) d.g_ay[i] = f.g_ay[i];
#endif
#if DIM > 2
    d.vz[i] = s.vz[i] + dt * (RK_B31 * s.az[i] + RK_B32 * f.az[i]);
    d.z[i] = s.z[i] + dt * (RK_B31 * s.dzdt[i] + RK_B32 * f.dzdt[i]);
    if (copy_gravity && d.g_az) d.g_az[i] = f.g_az[i];
#endif
}

#define RK_THIRD(dst, member, rate)                                                                                           \
    do {                                                                                                                      \
        const double k_ = RK_C1 * s.rate[i] + RK_C2 * f.rate[i] + RK_C3 * d.rate[i];                                         \
        dst.member[i] = s.member[i] + dt / 6.0 * k_;                                                                          \
        dst.rate[i] = 1. / 6. * k_;                                                                                           \
    } while (0)

/* integrateThirdStep + checkError, src/rk2adaptive.cu:953-1130 and :1134-1482, one pass */
__global__ void __launch_bounds__(RK_THREADS)
k_rk_third_check(RkBuffers b, const int *materialId, int n, b200sph_rk2_params prm, RkScalars *sc, double *partials, unsigned int *counter)
{
    const double dt = sc->dt;
    const b200sph_particle_arrays &p = b.p, &s = b.start, &f = b.first, &d = b.second;
    double e[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#if INTEGRATE_DENSITY
        RK_THIRD(p, rho, drhodt);
#else
        p.rho[i] = d.rho[i];
#endif
#if INTEGRATE_SML
        RK_THIRD(p, h, dhdt);
#else
        p.h[i] = d.h[i];
#endif
#if INTEGRATE_ENERGY
        RK_THIRD(p, e, dedt);
#endif
#if PALPHA_POROSITY
        const double dp = d.p[i] - s.p[i];
#endif
#if FRAGMENTATION
        RK_THIRD(p, d, dddt);
#if PALPHA_POROSITY
        if (dp > 0.0) RK_THIRD(p, damage_porjutzi, ddamage_porjutzidt);
        else p.damage_porjutzi[i] = s.damage_porjutzi[i];
#endif
#endif
#if PALPHA_POROSITY
        if (dp > 0.0) RK_THIRD(p, alpha_jutzi, dalphadt);
        else p.alpha_jutzi[i] = s.alpha_jutzi[i];
#endif
#if SOLID
#pragma unroll
        for (int c = 0; c < DD; c++) {
            const size_t o = (size_t)i * DD + c;
            const double k_ = RK_C1 * s.dSdt[o] + RK_C2 * f.dSdt[o] + RK_C3 * d.dSdt[o];
            p.S[o] = s.S[o] + dt / 6.0 * k_;
            p.dSdt[o] = 1. / 6. * k_;
        }
        {
            const double k_ = RK_C1 * s.edotp[i] + RK_C2 * f.edotp[i] + RK_C3 * d.edotp[i];
            p.ep[i] = s.ep[i] + dt / 6.0 * k_;
            p.edotp[i] = 1. / 6. * k_;
        }
#endif
        RK_THIRD(p, vx, ax);
        if (p.g_ax) p.g_ax[i] = 1. / 6.0 * (RK_C1 * s.g_ax[i] + RK_C2 * f.g_ax[i] + RK_C3 * d.g_ax[i]);
        p.x[i] = s.x[i] + dt / 6.0 * (RK_C1 * s.dxdt[i] + RK_C2 * f.dxdt[i] + RK_C3 * d.dxdt[i]);
#if DIM > 1
        RK_THIRD(p, vy, ay);
        if (p.g_ay) p.g_ay[i] = 1. / 6.0 * (RK_C1 * s.g_ay[i] + RK_C2 * f.g_ay[i] + RK_C3 * d.g_ay[i]);
        p.y[i] = s.y[i] + dt / 6.0 * (RK_C1 * s.dydt[i] + RK_C2 * f.dydt[i] + RK_C3 * d.dydt[i]);
#endif
#if DIM > 2
        RK_THIRD(p, vz, az);
        if (p.g_az) p.g_az[i] = 1. / 6.0 * (RK_C1 * s.g_az[i] + RK_C2 * f.g_az[i] + RK_C3 * d.g_az[i]);
        p.z[i] = s.z[i] + dt / 6.0 * (RK_C1 * s.dzdt[i] + RK_C2 * f.dzdt[i] + RK_C3 * d.dzdt[i]);
#endif
        /* remember some more values */
        p.noi[i] = d.noi[i];
        p.p[i] = d.p[i];
#if PALPHA_POROSITY
        p.pold[i] = d.p[i];
#endif
        p.cs[i] = d.cs[i];
#if FRAGMENTATION
        p.numActiveFlaws[i] = d.numActiveFlaws[i];
#endif
#if SOLID
        p.local_strain[i] = d.local_strain[i];
#endif

        /* ---- checkError: difference between the embedded second- and third-order results (Oxley 1999) ---- */
        if (materialId[i] == EOS_TYPE_IGNORE) continue;
        const double min_pos = s.h[i] * prm.location_safety;
        {
            const double t = dt * (f.dxdt[i] / 3.0 - (s.dxdt[i] + d.dxdt[i]) / 6.0);
            const double den = fabs(s.x[i]) + fabs(dt * s.dxdt[i]);
            if (den > min_pos) e[0] = fmax(e[0], fabs(t) / den);
        }
#if DIM > 1
        {
            const double t = dt * (f.dydt[i] / 3.0 - (s.dydt[i] + d.dydt[i]) / 6.0);
            const double den = fabs(s.y[i]) + fabs(dt * s.dydt[i]);
            if (den > min_pos) e[0] = fmax(e[0], fabs(t) / den);
        }
#endif
#if DIM > 2
        {
            const double t = dt * (f.dzdt[i] / 3.0 - (s.dzdt[i] + d.dzdt[i]) / 6.0);
            const double den = fabs(s.z[i]) + fabs(dt * s.dzdt[i]);
            if (den > min_pos) e[0] = fmax(e[0], fabs(t) / den);
        }
#endif
        if (prm.use_velocity_error) {
            {
                const double t = dt * (f.ax[i] / 3.0 - (s.ax[i] + d.ax[i]) / 6.0);
                const double den = fabs(s.vx[i]) + fabs(dt * s.ax[i]);
                if (den > prm.min_vel_change) e[1] = fmax(e[1], fabs(t) / den);
            }
#if DIM > 1
            {
                const double t = dt * (f.ay[i] / 3.0 - (s.ay[i] + d.ay[i]) / 6.0);
                const double den = fabs(s.vy[i]) + fabs(dt * s.ay[i]);
                if (den > prm.min_vel_change) e[1] = fmax(e[1], fabs(t) / den);
            }
#endif
#if DIM > 2
            {
                const double t = dt * (f.az[i] / 3.0 - (s.az[i] + d.az[i]) / 6.0);
                const double den = fabs(s.vz[i]) + fabs(dt * s.az[i]);
                if (den > prm.min_vel_change) e[1] = fmax(e[1], fabs(t) / den);
            }
#endif
        }
#if INTEGRATE_DENSITY
        if (prm.use_density_error) {
            const double t = dt * (f.drhodt[i] / 3.0 - (s.drhodt[i] + d.drhodt[i]) / 6.0);
            e[2] = fmax(e[2], fabs(t) / (fabs(s.rho[i]) + fabs(dt * s.drhodt[i]) + prm.tiny_density));
        }
#endif
#if INTEGRATE_ENERGY
        if (prm.use_energy_error) {
            const int eos = c_mat[materialId[i]].eos;
            const bool has_energy = eos == EOS_TYPE_TILLOTSON || eos == EOS_TYPE_JUTZI || eos == EOS_TYPE_JUTZI_ANEOS || eos == EOS_TYPE_SIRONO ||
                                    eos == EOS_TYPE_EPSILON || eos == EOS_TYPE_ANEOS || eos == EOS_TYPE_IDEAL_GAS;
            if (has_energy) {
                const double t = dt * (f.dedt[i] / 3.0 - (s.dedt[i] + d.dedt[i]) / 6.0);
                e[3] = fmax(e[3], fabs(t) / (fabs(s.e[i]) + fabs(dt * s.dedt[i]) + prm.tiny_energy));
            }
        }
#endif
#if PALPHA_POROSITY
        if (prm.limit_alpha_change) e[4] = fmax(e[4], fabs(s.alpha_jutzi_old[i] - p.alpha_jutzi[i]));
        if (prm.limit_pressure_change) e[5] = fmax(e[5], fabs(f.p[i] - d.p[i]));
#endif
    }
    double out[6];
    if (!rk_reduce<6>(e, false, partials, counter, out)) return;
    /* this rank's maxima; several GPUs all-reduce them (max) before k_rk_finish_check */
#pragma unroll
    for (int k = 0; k < 6; k++) sc->err[k] = out[k];
}

__global__ void k_rk_finish_check(RkScalars *sc, b200sph_rk2_params prm)
{
    const double dt = sc->dt;
    double out[6];
    for (int k = 0; k < 6; k++) out[k] = sc->err[k];
    double tmp = out[0];
    if (prm.use_velocity_error) tmp = fmax(tmp, out[1]);
    if (prm.use_density_error) tmp = fmax(tmp, out[2]);
    if (prm.use_energy_error) tmp = fmax(tmp, out[3]);
    tmp /= prm.rk_epsrel;
#if PALPHA_POROSITY
    if (prm.limit_pressure_change) tmp = fmax(tmp, out[5] / prm.max_pressure_change);
    if (prm.limit_alpha_change) tmp = fmax(tmp, out[4] / prm.max_alpha_change);
#endif
    double dt_new;
    if (tmp > 1.0) {
        sc->error_small_enough = 0;
        dt_new = fmax(0.1 * dt, dt * prm.timestep_safety * pow(tmp, -0.25));
    } else {
        sc->error_small_enough = 1;
        dt_new = dt * prm.timestep_safety * pow(tmp, -0.3);
        if (dt_new > 5.0 * dt) dt_new = 5.0 * dt;
        if (dt_new < dt) dt_new = dt;
    }
    sc->dt_new = dt_new;
}

/* a rejected step: rk[FIRST] <- rk[START], variables and derivatives (src/rk2adaptive.cu:477-484) */
__global__ void k_rk_restore(RkBuffers b, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    rk_copy_variables(b.first, b.start, i);
    rk_copy_derivatives(b.first, b.start, i);
}

/* ------------------------------------------------------------------ host side */
#define RCU(call)                                                                                    \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            snprintf(h->err, sizeof(h->err), "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return B200SPH_ERR_CUDA;                                                                 \
        }                                                                                            \
    } while (0)

extern "C" int b200sph_rk2_default_params(b200sph_rk2_params *prm)
{
    if (!prm) return B200SPH_ERR_BAD_ARGUMENT;
    /* include/rk2adaptive.h:39-71 and include/timeintegration.h:41-43 as shipped */
    prm->rk_epsrel = 1e-5;           /* -Q default, src/miluph.cu:625 */
    prm->dt_max = 0.0;               /* 0: the output interval */
    prm->first_dt = 0.0;
    prm->use_courant_limit = 1;
    prm->use_forces_limit = 0;
    prm->use_damage_limit = 1;
    prm->use_velocity_error = 0;
    prm->use_density_error = 1;
    prm->use_energy_error = 0;
    prm->limit_pressure_change = 0;
    prm->limit_alpha_change = 1;
    prm->courant_fact = 0.4;
    prm->forces_fact = 0.2;
    prm->location_safety = 0.1;
    prm->min_vel_change = 10.0;
    prm->tiny_density = 1e-2;
    prm->tiny_energy = 10.0;
    prm->timestep_safety = 0.9;
    prm->smallest_dt_allowed = 1e-16;
    prm->max_damage_change = 0.15;
    prm->max_alpha_change = 1e-2;
    prm->max_pressure_change = 1e100;
    return B200SPH_OK;
}

static int rk_scratch(b200sph_handle *h)
{
    if (h->rk_scalars) return 0;
    RCU(cudaMalloc(&h->rk_scalars, sizeof(RkScalars)));
    RCU(cudaMemset(h->rk_scalars, 0, sizeof(RkScalars)));
    RCU(cudaMalloc((void **)&h->rk_partials, sizeof(double) * RK_NRED * (size_t)h->n_sm * 8));
    RCU(cudaMalloc((void **)&h->rk_counter, sizeof(unsigned int)));
    RCU(cudaMemset(h->rk_counter, 0, sizeof(unsigned int)));
    return 0;
}

/* the right-hand side the integrator evaluates: b200sph_rhs_eval, or the multi-GPU host's exchange + evaluation (mg.cu) */
static int rk_rhs(b200sph_handle *h, const b200sph_view *bound, int *offender)
{
    if (h->rk_rhs_hook) return h->rk_rhs_hook(h->rk_hook_ctx, bound, offender);
    return b200sph_rhs_eval(h, bound, offender);
}

static bool rk_buffers_ok(const b200sph_particle_arrays &a)
{
    return a.x && a.vx && a.ax && a.dxdt && a.m && a.h && a.rho && a.drhodt && a.p && a.cs && a.noi;
}

extern "C" int b200sph_rk2_init(b200sph_handle *h, const b200sph_view *view, const b200sph_particle_arrays rk[3])
{
    if (!h || !view || !rk) return B200SPH_ERR_BAD_ARGUMENT;
    RCU(cudaSetDevice(h->device));
    for (int k = 0; k < 3; k++)
        if (!rk_buffers_ok(rk[k])) {
            snprintf(h->err, sizeof(h->err), "rk buffer %d is missing a mandatory array", k);
            return B200SPH_ERR_BAD_ARGUMENT;
        }
    if (rk_scratch(h)) return B200SPH_ERR_CUDA;
    RkBuffers b = {view->p, rk[0], rk[1], rk[2]};
    k_rk_init<<<(view->n + 255) / 256, 256, 0, h->stream>>>(b, view->n);
    RCU(cudaStreamSynchronize(h->stream));
    RCU(cudaGetLastError());
    return B200SPH_OK;
}

/* One ACCEPTED step of rk2_adaptive (the body of `while (currentTime < endTime)`, src/rk2adaptive.cu:197-485):
 * state->dt is the step to try; on return state holds the advanced time, the step taken and the next step to try. */
extern "C" int b200sph_rk2_step(b200sph_handle *h, const b200sph_view *view, const b200sph_particle_arrays rk[3],
                                const b200sph_rk2_params *prm, double t_end, b200sph_rk2_state *state, int *offender)
{
    if (!h || !view || !rk || !prm || !state) return B200SPH_ERR_BAD_ARGUMENT;
    RCU(cudaSetDevice(h->device));
    if (rk_scratch(h)) return B200SPH_ERR_CUDA;
    if (!(state->dt > 0.0)) {
        snprintf(h->err, sizeof(h->err), "rk2_step: state->dt = %g is not a positive step size", state->dt);
        return B200SPH_ERR_BAD_ARGUMENT;
    }
    const int n = view->n;
    cudaStream_t st = h->stream;
    RkScalars *sc = (RkScalars *)h->rk_scalars;
    RkBuffers b = {view->p, rk[0], rk[1], rk[2]};
    const int G = (n + RK_THREADS - 1) / RK_THREADS;
    const int GR = min(G, h->n_sm * 8);
    const double dt_max = (prm->dt_max > 0.0) ? prm->dt_max : DBL_MAX;
    RkScalars hs;
    b200sph_view v1 = *view, v2 = *view;
    v1.p = rk[1];
    v2.p = rk[2];
    int rc;

    double dt_host = state->dt;
    RCU(cudaMemcpyAsync(&sc->dt, &dt_host, sizeof(double), cudaMemcpyHostToDevice, st));
    k_rk_copy_vars<<<G, RK_THREADS, 0, st>>>(rk[1], view->p, n);
    if ((rc = rk_rhs(h, &v1, offender)) != 0) return rc;
    state->rhs_calls++;
    k_rk_limit_remember<<<GR, RK_THREADS, 0, st>>>(b, n, prm->use_courant_limit, prm->use_forces_limit, prm->use_damage_limit,
                                                 prm->max_damage_change, sc, h->rk_partials, h->rk_counter);
    if (h->rk_allreduce && (rc = h->rk_allreduce(h->rk_hook_ctx, sc->limit, 3, 1)) != 0) return rc;
    k_rk_apply_limits<<<1, 1, 0, st>>>(sc, prm->use_courant_limit, prm->use_forces_limit, prm->use_damage_limit, prm->courant_fact,
                                       prm->forces_fact);
    for (;;) {
        k_rk_first<<<G, RK_THREADS, 0, st>>>(b, n, sc);
        RCU(cudaMemcpyAsync(&hs, sc, sizeof(RkScalars), cudaMemcpyDeviceToHost, st));
        RCU(cudaStreamSynchronize(st));
        dt_host = hs.dt;
        if (dt_host < prm->smallest_dt_allowed && !state->approaching_output_time) {
            snprintf(h->err, sizeof(h->err), "timestep %e is below SMALLEST_DT_ALLOWED (src/rk2adaptive.cu:281-284)", dt_host);
            return B200SPH_ERR_BAD_ARGUMENT;
        }
        if ((rc = rk_rhs(h, &v1, offender)) != 0) return rc;
        k_rk_second<<<G, RK_THREADS, 0, st>>>(b, n, sc, view->selfgravity);
        if ((rc = rk_rhs(h, &v2, offender)) != 0) return rc;
        state->rhs_calls += 2;
        k_rk_third_check<<<GR, RK_THREADS, 0, st>>>(b, view->p_rhs.materialId, n, *prm, sc, h->rk_partials, h->rk_counter);
        if (h->rk_allreduce && (rc = h->rk_allreduce(h->rk_hook_ctx, sc->err, 6, 0)) != 0) return rc;
        k_rk_finish_check<<<1, 1, 0, st>>>(sc, *prm);
        RCU(cudaMemcpyAsync(&hs, sc, sizeof(RkScalars), cudaMemcpyDeviceToHost, st));
        RCU(cudaStreamSynchronize(st));
        RCU(cudaGetLastError());
        double dt_suggested = hs.dt_new;
        for (int k = 0; k < 6; k++) state->err[k] = hs.err[k];
        if (hs.error_small_enough) {
            state->t += dt_host;
            state->dt_done = dt_host;
            state->accepted++;
        } else {
            state->rejected++;
        }
        if (dt_suggested > dt_max) dt_suggested = dt_max;
        if (state->t + dt_suggested > t_end) {
            dt_host = t_end - state->t;
            state->approaching_output_time = 1;
        } else {
            dt_host = dt_suggested;
        }
        state->dt_suggested = dt_suggested;
        state->dt = dt_host;
        RCU(cudaMemcpyAsync(&sc->dt, &dt_host, sizeof(double), cudaMemcpyHostToDevice, st));
        if (hs.error_small_enough) break;
        k_rk_restore<<<G, RK_THREADS, 0, st>>>(b, n);
    }
    RCU(cudaStreamSynchronize(st));
    return B200SPH_OK;
}

/* One output interval: steps until t_end like the reference's loop (src/rk2adaptive.cu:145-485), then damageLimit as
 * before every output (src/rk2adaptive.cu:464-469). */
extern "C" int b200sph_rk2_advance(b200sph_handle *h, const b200sph_view *view, const b200sph_particle_arrays rk[3],
                                   const b200sph_rk2_params *prm, double t_end, b200sph_rk2_state *state, int *offender)
{
    if (!h || !view || !rk || !prm || !state) return B200SPH_ERR_BAD_ARGUMENT;
    const double interval = t_end - state->t;
    if (!(interval > 0.0)) return B200SPH_ERR_BAD_ARGUMENT;
    const double dt_max = (prm->dt_max > 0.0) ? prm->dt_max : interval;
    state->approaching_output_time = 0;
    if (state->intervals == 0) {
        /* first dt of the run (src/rk2adaptive.cu:153-163) */
        if (prm->first_dt > 0.0 && interval > prm->first_dt) state->dt = state->dt_suggested = prm->first_dt;
        else if (dt_max < interval) state->dt = state->dt_suggested = dt_max;
        else state->dt = state->dt_suggested = interval;
    } else {
        state->dt = state->dt_suggested;
        if (state->dt < prm->smallest_dt_allowed) state->dt = 1.1 * prm->smallest_dt_allowed;
        if (state->dt > interval) state->dt = interval;
    }
    state->intervals++;
    while (state->t < t_end) {
        const int rc = b200sph_rk2_step(h, view, rk, prm, t_end, state, offender);
        if (rc) return rc;
    }
    return b200sph_damage_limit(h, view);
}

/* ------------------------------------------------------------------ conserved quantities on the device
 * The reference copies the whole particle set to the host before every output and sums mass, energies, linear and
 * angular momentum and the barycentre there in one thread (src/io.cu:1661-1838) -- the numbers of conserved_quantities.log
 * (src/io.cu:1980-2017).  Here: one streaming pass, deterministic (fixed block order), 13 sums read back (104 bytes). */
#define CQ_NV 13
__global__ void __launch_bounds__(RK_THREADS)
k_conserved(b200sph_view v, double *partials, unsigned int *counter, double *out)
{
    const b200sph_particle_arrays &p = v.p;
    double a[CQ_NV];
#pragma unroll
    for (int k = 0; k < CQ_NV; k++) a[k] = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < v.n; i += gridDim.x * blockDim.x) {
        if (v.p_rhs.materialId[i] == EOS_TYPE_IGNORE) {
            a[12] += 1.0;
            continue;
        }
        const double m = p.m[i];
        const double x = p.x[i], vx = p.vx[i];
        const double y = (DIM > 1) ? p.y[i] : 0.0, vy = (DIM > 1) ? p.vy[i] : 0.0;
        const double z = (DIM > 2) ? p.z[i] : 0.0, vz = (DIM > 2) ? p.vz[i] : 0.0;
        a[0] += m;
        a[1] += m * (vx * vx + vy * vy + vz * vz);
#if INTEGRATE_ENERGY
        a[2] += m * p.e[i];
#endif
        a[3] += m * vx; a[4] += m * vy; a[5] += m * vz;
#if DIM == 2
        a[6] += m * (x * vy - y * vx);
#elif DIM > 2
        a[6] += m * (y * vz - z * vy);
        a[7] += m * (z * vx - x * vz);
        a[8] += m * (x * vy - y * vx);
#endif
        a[9] += m * x; a[10] += m * y; a[11] += m * z;
    }
    __shared__ double sh[RK_THREADS / 32][CQ_NV];
    __shared__ bool last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < CQ_NV; k++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a[k] += __shfl_xor_sync(0xffffffffu, a[k], o);
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < CQ_NV; k++) sh[warp][k] = a[k];
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 0; k < CQ_NV; k++) {
            double t = 0.0;
            for (int w = 0; w < RK_THREADS / 32; w++) t += sh[w][k];
            partials[blockIdx.x * CQ_NV + k] = t;
        }
        __threadfence();
        last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!last || threadIdx.x >= CQ_NV) return;
    __threadfence();
    double t = 0.0;
    for (unsigned int b = 0; b < gridDim.x; b++) t += __ldcg(partials + b * CQ_NV + threadIdx.x);   /* block order: reproducible */
    out[threadIdx.x] = t;
    __syncwarp((1u << CQ_NV) - 1u);
    if (threadIdx.x == 0) *counter = 0;
}

extern "C" int b200sph_conserved_quantities(b200sph_handle *h, const b200sph_view *view, b200sph_conserved *out)
{
    if (!h || !view || !out || view->n <= 0 || !view->p.x || !view->p.vx || !view->p.m || !view->p_rhs.materialId) return B200SPH_ERR_BAD_ARGUMENT;
    RCU(cudaSetDevice(h->device));
    if (rk_scratch(h)) return B200SPH_ERR_CUDA;
    const int G = min((view->n + RK_THREADS - 1) / RK_THREADS, h->n_sm * 4);
    double *d_out = (double *)h->rk_scalars;   /* sizeof(RkScalars) >= 13 doubles */
    static_assert(sizeof(RkScalars) >= CQ_NV * sizeof(double), "scratch too small");
    static_assert(CQ_NV <= 2 * RK_NRED, "partials are sized for 2 * RK_NRED values per block over 8 n_sm blocks");
    k_conserved<<<G, RK_THREADS, 0, h->stream>>>(*view, h->rk_partials, h->rk_counter, d_out);
    double r[CQ_NV];
    RCU(cudaMemcpyAsync(r, d_out, sizeof(r), cudaMemcpyDeviceToHost, h->stream));
    RCU(cudaStreamSynchronize(h->stream));
    RCU(cudaGetLastError());
    out->mass = r[0];
    out->e_kin = 0.5 * r[1];
    out->e_int = r[2];
    for (int k = 0; k < 3; k++) {
        out->p[k] = r[3 + k];
        out->L[k] = r[6 + k];
        out->bary_pos[k] = r[0] > 0.0 ? r[9 + k] / r[0] : 0.0;
        out->bary_vel[k] = r[0] > 0.0 ? r[3 + k] / r[0] : 0.0;
    }
    out->p_abs = sqrt(r[3] * r[3] + r[4] * r[4] + r[5] * r[5]);
    out->L_abs = sqrt(r[6] * r[6] + r[7] * r[7] + r[8] * r[8]);
    out->n_ignored = (int)r[12];
    return B200SPH_OK;
}
