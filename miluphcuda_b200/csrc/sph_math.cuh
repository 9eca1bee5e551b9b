/*
 * sph_math.cuh -- device-side building blocks of the B200 SPH right-hand side:
 * the cubic B-spline kernel, the small symmetric eigen-solvers and the
 * equations of state.  All arithmetic is FP64.
 *
 * Written for throughput on sm_100a: no function pointers (the reference calls
 * its kernel through `__device__ SPH_kernel kernel`, src/kernel.cu:56-62), one
 * rsqrt instead of sqrt + two divisions per pair, normalisation constants
 * folded per particle where h is fixed.
 */
#ifndef B200SPH_MATH_CUH
#define B200SPH_MATH_CUH

#include "switches.h"

#define B200_MAX_MATERIALS 16
#define DD (DIM * DIM)

/* one entry per material id, broadcast from constant memory
 * (reference: one global array per property, include/config_parameter.h:180-290) */
struct MatParams {
    int eos, density_via_kernel_sum, crushcurve_style, aneos_n_rho, aneos_n_e, aneos_rho_id, aneos_e_id, aneos_matrix_id;
    double sml, f_sml_min, f_sml_max, av_alpha, av_beta;
    double poly_K, poly_gamma, iso_cs, bulk, shear, young, yield_stress;
    double rho0, n, rho_limit, cs_limit;
    double till_rho0, till_A, till_B, till_E0, till_Eiv, till_Ecv, till_a, till_b, till_alpha, till_beta;
    double cohesion, cohesion_damaged, friction, friction_damaged, melt_energy;
    double density_floor, energy_floor;
    double exponent_tensor, epsilon_stress, mean_particle_distance;
    double pj_p_elastic, pj_p_transition, pj_p_compacted, pj_alpha_0, pj_alpha_e, pj_alpha_t, pj_n1, pj_n2;
    double cs_porous, cs_solid;
    double aneos_bulk_cs, aneos_gamma;
};

/* the library is a single translation unit (libb200sph.cu), so the symbols are defined here */
static __constant__ MatParams c_mat[B200_MAX_MATERIALS];

struct AneosTables {
    const double *rho, *e, *p, *cs;
};
static __constant__ AneosTables c_aneos;

__device__ __forceinline__ bool mat_ignored(int matId)
{
    return matId == EOS_TYPE_IGNORE || c_mat[matId].eos == EOS_TYPE_IGNORE;
}

/* ------------------------------------------------------------------ SPH kernel
 * Cubic B-spline with support radius h (q = r/h <= 1), reference src/kernel.cu:112-153:
 *   W = f (6q^3 - 6q^2 + 1)            q <= 1/2
 *   W = 2 f (1-q)^3                    1/2 < q <= 1
 *   f = 4/(3h), 40/(7 pi h^2), 8/(pi h^3)  for DIM = 1, 2, 3
 * Returns W and g = (dW/dr)/r so that grad W = g * dr. */
__device__ __forceinline__ double kernel_norm(double hinv)
{
#if DIM == 1
    return (4.0 / 3.0) * hinv;
#elif DIM == 2
    return (40.0 / (7.0 * M_PI)) * hinv * hinv;
#else
    return (8.0 / M_PI) * hinv * hinv * hinv;
#endif
}

/* Measured on a B200 (gpurun_out v1/q2, round 1): pinning the reciprocal square root in a register cuts the
 * solid pair loops by 30 % (impact k_forces 1.70 -> 1.20 ms: the FP64 pipe was the busiest unit there), but
 * slows the hydro loops by 15 % (sedov k_forces 0.364 -> 0.420 ms: they are bound by L1TEX wavefronts and the
 * leaner code raises occupancy and with it the L1 miss rate).  Hence per switch set. */
#ifndef B200_RSQRT_OPAQUE
#define B200_RSQRT_OPAQUE SOLID
#endif
__device__ __forceinline__ double pair_rsqrt(double x) { return rsqrt(x); }

__device__ __forceinline__ void cubic_spline(double r2, double hinv, double &W, double &g)
{
    const double f = kernel_norm(hinv);
    double rinv = pair_rsqrt(r2);
    /* opaque to the optimiser: nvcc otherwise re-evaluates the reciprocal square root (MUFU + 5 FP64
     * instructions + slow-path check) in each branch below instead of keeping it in a register */
#if B200_RSQRT_OPAQUE
    asm volatile("" : "+d"(rinv));
#endif
    const double r = r2 * rinv;
    const double q = r * hinv;
    if (q > 1.0) {
        W = 0.0;
        g = 0.0;
    } else if (q > 0.5) {
        const double t = 1.0 - q;
        W = 2.0 * f * t * t * t;
        g = -6.0 * f * hinv * t * t * rinv;
    } else {
        W = f * fma(6.0 * q * q, q - 1.0, 1.0);
        /* dW/dr / r = 6 f/h (3q^2 - 2q) / r = 6 f/h^2 (3q - 2) */
        g = 6.0 * f * hinv * hinv * fma(3.0, q, -2.0);
    }
}

/* W only, r given (self term and artificial-stress reference distance) */
__device__ __forceinline__ double cubic_spline_w(double r, double hinv)
{
    const double f = kernel_norm(hinv);
    const double q = r * hinv;
    if (q > 1.0) return 0.0;
    if (q > 0.5) {
        const double t = 1.0 - q;
        return 2.0 * f * t * t * t;
    }
    return f * fma(6.0 * q * q, q - 1.0, 1.0);
}

#if DIM > 1
/* ------------------------------------------------------------------ Jacobi
 * Same pivot rule and stopping criteria as the reference so that converged
 * results agree to rounding (src/linalg.cu:107-127 pivot with ">=", :177-234
 * all eigenvalues until max|offdiag| <= 1e-10, :245-294 largest eigenvalue with
 * at least 5 rotations).  Index juggling is done on unrolled registers. */
struct SymPivot {
    int e, f;
    double mx;
};

__device__ __forceinline__ SymPivot jacobi_pivot(const double (&M)[DIM][DIM])
{
    SymPivot pv;
    pv.e = 0; pv.f = 0; pv.mx = 0.0;
#pragma unroll
    for (int i = 0; i < DIM; i++)
#pragma unroll
        for (int j = 0; j < DIM; j++) {
            if (i == j) continue;
            const double a = fabs(M[i][j]);
            if (a >= pv.mx) { pv.mx = a; pv.e = i; pv.f = j; }
        }
    return pv;
}

__device__ __forceinline__ void jacobi_angle(const double (&D)[DIM][DIM], int e, int f, double &c, double &s)
{
    const double thta = (D[f][f] - D[e][e]) / (2.0 * D[e][f]);
    double t = 1.0 / (fabs(thta) + sqrt(fma(thta, thta, 1.0)));
    if (thta < 0.0) t = -t;
    c = rsqrt(fma(t, t, 1.0));
    s = t * c;
}

__device__ __forceinline__ void jacobi_rotate(double (&m)[DIM][DIM], double c, double s, int e, int f)
{
    const double mee = m[e][e], mff = m[f][f], mef = m[e][f];
    double ne[DIM], nf[DIM];
#pragma unroll
    for (int i = 0; i < DIM; i++) {
        ne[i] = c * m[i][e] - s * m[i][f];
        nf[i] = c * m[i][f] + s * m[i][e];
    }
#pragma unroll
    for (int i = 0; i < DIM; i++) {
        if (i == e || i == f) continue;
        m[e][i] = ne[i]; m[i][e] = ne[i];
        m[f][i] = nf[i]; m[i][f] = nf[i];
    }
    m[e][e] = c * c * mee + s * s * mff - 2.0 * s * c * mef;
    m[f][f] = c * c * mff + s * s * mee + 2.0 * s * c * mef;
    const double off = (c * c - s * s) * mef + s * c * (mee - mff);
    m[e][f] = off;
    m[f][e] = off;
}

/* eigenvalues ev[] and eigenvectors (columns of V) of a symmetric matrix */
__device__ inline void sym_eigen(const double (&M)[DIM][DIM], double (&ev)[DIM], double (&V)[DIM][DIM])
{
    double D[DIM][DIM];
#pragma unroll
    for (int i = 0; i < DIM; i++)
#pragma unroll
        for (int j = 0; j < DIM; j++) {
            D[i][j] = M[i][j];
            V[i][j] = (i == j) ? 1.0 : 0.0;
        }
    SymPivot pv;
    int guard = 0;
    do {
        pv = jacobi_pivot(D);
        if (pv.mx > 0.0) {
            double c, s;
            jacobi_angle(D, pv.e, pv.f, c, s);
            jacobi_rotate(D, c, s, pv.e, pv.f);
            /* V <- V * A with A_ee = A_ff = c, A_ef = s, A_fe = -s */
#pragma unroll
            for (int i = 0; i < DIM; i++) {
                const double ve = V[i][pv.e], vf = V[i][pv.f];
                V[i][pv.e] = c * ve - s * vf;
                V[i][pv.f] = s * ve + c * vf;
            }
        }
    } while (pv.mx > 1e-10 && ++guard < 200);
#pragma unroll
    for (int i = 0; i < DIM; i++) ev[i] = D[i][i];
}

__device__ inline double sym_max_eigenvalue(const double (&M)[DIM][DIM])
{
    double D[DIM][DIM];
#pragma unroll
    for (int i = 0; i < DIM; i++)
#pragma unroll
        for (int j = 0; j < DIM; j++) D[i][j] = M[i][j];
    SymPivot pv;
    int nit = 0;
    do {
        nit++;
        pv = jacobi_pivot(D);
        if (pv.mx > 0.0) {
            double c, s;
            jacobi_angle(D, pv.e, pv.f, c, s);
            jacobi_rotate(D, c, s, pv.e, pv.f);
        }
    } while ((pv.mx > 1e-10 || nit < 5) && nit < 200);
    double best = D[0][0];
#pragma unroll
    for (int i = 1; i < DIM; i++) best = fmax(best, D[i][i]);
    return best;
}

/* pseudo-inverse of a symmetric matrix through its eigen-decomposition, dropping
 * eigenvalues below 1e-6 * max|ev| (reference: src/linalg.cu:351-431) */
__device__ inline void sym_pinv(const double (&A)[DIM][DIM], double (&P)[DIM][DIM])
{
    double ev[DIM], V[DIM][DIM];
    sym_eigen(A, ev, V);
    double smax = 0.0;
#pragma unroll
    for (int k = 0; k < DIM; k++) smax = fmax(smax, fabs(ev[k]));
    const double thr = 1e-6 * smax;
#pragma unroll
    for (int i = 0; i < DIM; i++)
#pragma unroll
        for (int j = 0; j < DIM; j++) P[i][j] = 0.0;
#pragma unroll
    for (int k = 0; k < DIM; k++) {
        if (fabs(ev[k]) > thr) {
            const double iv = 1.0 / ev[k];
#pragma unroll
            for (int i = 0; i < DIM; i++)
#pragma unroll
                for (int j = 0; j < DIM; j++) P[i][j] += iv * V[i][k] * V[j][k];
        }
    }
}
#endif /* DIM > 1 */

/* ------------------------------------------------------------------ equations of state */
__device__ __forceinline__ double sq(double x) { return x * x; }

/* Tillotson c_s^2 with the caller's (previous) pressure, reference src/soundspeed.cu:76-107 */
__device__ inline double tillotson_cs2(const MatParams &M, double rho, double e, double pressure)
{
    const double eta = rho / M.till_rho0;
    const double omega0 = e / (M.till_E0 * eta * eta) + 1.0;
    const double mu = eta - 1.0;
    const double z = (1.0 - eta) / eta;
    const double iw2 = 1.0 / (omega0 * omega0);
    if (eta >= 0.0 || e < M.till_Eiv) {
        if (pressure < 0.0 || eta < M.rho_limit) pressure = 0.0;
        return M.till_a * e + (M.till_b * e) * iw2 * (3.0 * omega0 - 2.0) + (M.till_A + 2.0 * M.till_B * mu) / rho +
               pressure / (rho * rho) * (M.till_a * rho + M.till_b * rho * iw2);
    }
    const double ez = exp(-M.till_beta * z * z);
    const double Ge = M.till_a + M.till_b / omega0 * ez;
    const double cs_e = (Ge + 1.0) * pressure / rho +
                        M.till_A / rho * exp(-(M.till_alpha * z + M.till_beta * z * z)) * (1.0 + mu) / (eta * eta) *
                            (M.till_alpha + 2.0 * M.till_beta * z - eta) +
                        M.till_b * rho * e * iw2 / (eta * eta) * ez * (2.0 * M.till_beta * z * omega0 / M.till_rho0 + 1.0) /
                            (M.till_E0 * rho) * (2.0 * e - pressure / rho);
    if (e > M.till_Ecv) return cs_e;
    if (pressure < 0.0 || eta < M.rho_limit) pressure = 0.0;
    const double cs_c = M.till_a * e + (M.till_b * e) * iw2 * (3.0 * omega0 - 2.0) + (M.till_A + 2.0 * M.till_B * mu) / rho +
                        pressure / (rho * rho) * (M.till_a * rho + M.till_b * rho * iw2);
    const double y = (e - M.till_Eiv) / (M.till_Ecv - M.till_Eiv);
    return cs_e * (1.0 - y) + cs_c * y;
}

/* Tillotson pressure at matrix density rho (reference src/pressure.cu:61-97).  The p-alpha
 * variant (src/pressure.cu:234-298) differs in three comparisons; `porous` selects them and
 * the analytic derivatives dp/de, dp/drho are returned for it. */
__device__ inline double tillotson_p(const MatParams &M, double rho, double e, bool porous, double &dpde, double &dpdrho)
{
    const double r0 = M.till_rho0, eta = rho / r0, mu = eta - 1.0;
    const double a = M.till_a, b = M.till_b, A = M.till_A, B = M.till_B, E0 = M.till_E0;
    const double Eiv = M.till_Eiv, Ecv = M.till_Ecv;
    dpde = 0.0;
    dpdrho = 0.0;
    if (eta < M.rho_limit && e < Ecv) return 0.0;
    const double w = e / (eta * eta * E0) + 1.0;
    const bool cold = porous ? (e < Eiv || eta >= 1.0) : (e <= Eiv || eta >= 1.0);
    const bool hot = porous ? (e > Ecv && eta < 1.0) : (e >= Ecv && eta >= 0.0);
    const bool mid = porous ? (e > Eiv && eta < 1.0) : (e > Eiv && e < Ecv);
    double pc = 0.0, ph = 0.0, dpc_de = 0.0, dpc_dr = 0.0, dph_de = 0.0, dph_dr = 0.0;
    if (cold || mid) {
        pc = (a + b / w) * rho * e + A * mu + B * mu * mu;
        if (porous) {
            dpc_de = a * rho + rho * b / (w * w);
            dpc_dr = a * e + e * b * (1.0 + 3.0 * e / (E0 * eta * eta)) / (w * w) + A / r0 + 2.0 * B / r0 * mu;
        }
    }
    if (!cold && (hot || mid)) {
        const double zz = r0 / rho - 1.0;
        const double eb = exp(-M.till_beta * zz), ea = exp(-M.till_alpha * zz * zz);
        ph = a * rho * e + (b * rho * e / w + A * mu * eb) * ea;
        if (porous) {
            dph_de = a * rho + rho * b / (w * w) * ea;
            dph_dr = a * e + ea * (2.0 * M.till_alpha * r0 / (rho * rho) * zz * (b * rho * e / w + A * mu * eb) +
                                   b * e * (1.0 + 3.0 * e / (E0 * eta * eta)) / (w * w) +
                                   A * eb * (1.0 / r0 + M.till_beta / rho - M.till_beta * r0 / (rho * rho)));
        }
    }
    if (cold) {
        dpde = dpc_de;
        dpdrho = dpc_dr;
        return pc;
    }
    if (hot) {
        dpde = dph_de;
        dpdrho = dph_dr;
        return ph;
    }
    if (mid) {
        if (porous) {
            /* as written in the reference (src/pressure.cu:273-291), including its operator grouping */
            dpde = ((ph - pc) + (e - Eiv) * a * rho + rho * b / (w * w) * exp(-M.till_alpha * sq(r0 / rho - 1.0)) +
                    (Ecv - e) * a * rho + rho * b / (w * w)) / (Ecv - Eiv);
            dpdrho = (dph_dr * (e - Eiv) + dpc_dr * (Ecv - e)) / (Ecv - Eiv);
            return ((e - Eiv) * ph + (Ecv - e) * pc) / (Ecv - Eiv);
        }
        return (pc * (Ecv - e) + ph * (e - Eiv)) / (Ecv - Eiv);
    }
    return 0.0;
}

/* tabulated EOS: bisection + bilinear interpolation, reference src/aneos.cu:232-252, 282-383 */
__device__ inline int table_index(double x, const double *arr, int n)
{
    if (x < arr[0] || x >= arr[n - 1]) return -1;
    int i1 = 0, i2 = n - 1;
    do {
        const int i = (i1 + i2) >> 1;
        if (arr[i] <= x) i1 = i; else i2 = i;
    } while (i2 - i1 > 1);
    return i1;
}

struct TableCell {
    int ix, iy;
    double nx, ny;   /* normalised offsets inside the cell (may lie outside [0,1] after clamping) */
    bool ideal_gas;  /* e above the table: ideal-gas fallback */
};

__device__ inline TableCell aneos_locate(const MatParams &M, double rho, double e)
{
    const double *rt = c_aneos.rho + M.aneos_rho_id;
    const double *et = c_aneos.e + M.aneos_e_id;
    TableCell c;
    c.ideal_gas = false;
    c.ix = table_index(rho, rt, M.aneos_n_rho);
    if (c.ix < 0) c.ix = (rho < rt[0]) ? 0 : M.aneos_n_rho - 2;
    c.iy = table_index(e, et, M.aneos_n_e);
    if (c.iy < 0 && e >= et[M.aneos_n_e - 1]) {
        c.ideal_gas = true;
        c.iy = 0;
    } else if (c.iy < 0) {
        c.iy = 0;
        e = et[0];
    }
    c.nx = (rho - rt[c.ix]) / (rt[c.ix + 1] - rt[c.ix]);
    c.ny = (e - et[c.iy]) / (et[c.iy + 1] - et[c.iy]);
    return c;
}

__device__ inline double aneos_bilinear(const MatParams &M, const double *table, const TableCell &c)
{
    const double *t = table + M.aneos_matrix_id;
    const int ne = M.aneos_n_e;
    const double t00 = t[c.ix * ne + c.iy], t10 = t[(c.ix + 1) * ne + c.iy];
    const double t01 = t[c.ix * ne + c.iy + 1], t11 = t[(c.ix + 1) * ne + c.iy + 1];
    const double a = t00 + c.nx * (t10 - t00);
    const double b = t01 + c.nx * (t11 - t01);
    return a + c.ny * (b - a);
}

#endif
