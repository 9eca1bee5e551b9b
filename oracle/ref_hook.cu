/*
 * ref_hook.cu -- TEST INFRASTRUCTURE, not part of the product.
 *
 * Replaces the reference's src/euler.cu translation unit when the unmodified
 * reference sources are compiled from /root/reference into oracle/_ref/
 * (see oracle/build_ref.sh).  The reference's main() still parses its own
 * command line, reads material.cfg and the ASCII input file, allocates its
 * own buffers and calls `integrator()`; selecting `-I euler` lands here
 * instead of in the reference's Euler loop (reference: src/miluph.cu:1026-1029,
 * src/timeintegration.cu:244-249).
 *
 * What it does with the reference's own rightHandSide() (include/rhs.h:30):
 *   1. pins the one undefined input of the first call: p_device.p
 *      is zeroed (and g_a* for -g runs), so c_s at call 1 is evaluated with p = 0;
 *   2. dumps the input state, calls rightHandSide() once, dumps every output,
 *      calls it a second time, dumps again  (REF_DUMP=<file prefix>);
 *   3. optionally times REF_TIMED calls after REF_WARMUP warm-up calls with a
 *      cudaEvent pair around each call and prints one summary line
 *      "REF_TIMING n=<N> calls=<K> ms_per_call=<t> updates_per_s=<u>".
 *
 * Dump container: repeated records { char name[32]; int32 dtype (0=f64,1=i32);
 * int64 count; payload }.
 */
#include "miluph.h"
#include "timeintegration.h"
#include "rhs.h"
#include "parameter.h"
#include "memory_handling.h"

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

extern int flag_force_gravity_calc;
extern int gravity_index;

static void hook_write(FILE *f, const char *name, int dtype, int64_t count, const void *dev)
{
    char tag[32];
    size_t bytes = (size_t)count * (dtype == 0 ? sizeof(double) : sizeof(int));
    void *host = malloc(bytes ? bytes : 1);
    memset(tag, 0, sizeof(tag));
    strncpy(tag, name, sizeof(tag) - 1);
    if (bytes) cudaVerify(cudaMemcpy(host, dev, bytes, cudaMemcpyDeviceToHost));
    fwrite(tag, 1, sizeof(tag), f);
    fwrite(&dtype, sizeof(int), 1, f);
    fwrite(&count, sizeof(int64_t), 1, f);
    fwrite(host, 1, bytes, f);
    free(host);
}

#define DUMP_F64(field, cnt) hook_write(f, #field, 0, (int64_t)(cnt), p_device.field)
#define DUMP_I32(field, cnt) hook_write(f, #field, 1, (int64_t)(cnt), p_device.field)

static void hook_dump(const char *prefix, const char *stage, int with_lists)
{
    char fname[1024];
    FILE *f;
    const int64_t N = numberOfParticles;
    snprintf(fname, sizeof(fname), "%s.%s.bin", prefix, stage);
    if ((f = fopen(fname, "wb")) == NULL) {
        fprintf(stderr, "ref_hook: cannot open %s\n", fname);
        exit(1);
    }
    cudaVerify(cudaDeviceSynchronize());

    DUMP_F64(x, N); DUMP_F64(vx, N); DUMP_F64(ax, N); DUMP_F64(dxdt, N);
#if DIM > 1
    DUMP_F64(y, N); DUMP_F64(vy, N); DUMP_F64(ay, N); DUMP_F64(dydt, N);
#endif
#if DIM > 2
    DUMP_F64(z, N); DUMP_F64(vz, N); DUMP_F64(az, N); DUMP_F64(dzdt, N);
#endif
    DUMP_F64(m, N); DUMP_F64(h, N); DUMP_F64(h0, N); DUMP_F64(rho, N); DUMP_F64(e, N);
    DUMP_F64(p, N); DUMP_F64(cs, N); DUMP_F64(drhodt, N);
    DUMP_I32(materialId, N); DUMP_I32(noi, N); DUMP_I32(depth, N);
    if (param.selfgravity) {
        DUMP_F64(g_ax, N);
#if DIM > 1
        DUMP_F64(g_ay, N);
#endif
#if DIM > 2
        DUMP_F64(g_az, N);
#endif
    }
#if INTEGRATE_ENERGY
    DUMP_F64(dedt, N);
#endif
#if INTEGRATE_SML
    DUMP_F64(dhdt, N);
#endif
#if ARTIFICIAL_VISCOSITY
    DUMP_F64(muijmax, N);
#endif
#if SOLID
    DUMP_F64(S, N * DIM * DIM); DUMP_F64(dSdt, N * DIM * DIM); DUMP_F64(sigma, N * DIM * DIM);
    DUMP_F64(local_strain, N); DUMP_F64(edotp, N); DUMP_F64(plastic_f, N); DUMP_F64(ep, N);
#endif
#if TENSORIAL_CORRECTION
    DUMP_F64(tensorialCorrectionMatrix, N * DIM * DIM);
#endif
#if ARTIFICIAL_STRESS
    DUMP_F64(R, N * DIM * DIM);
#endif
#if FRAGMENTATION
    DUMP_F64(d, N); DUMP_F64(damage_total, N); DUMP_F64(dddt, N);
    DUMP_I32(numFlaws, N); DUMP_I32(numActiveFlaws, N);
    DUMP_F64(flaws, N * MAX_NUM_FLAWS);
# if PALPHA_POROSITY
    DUMP_F64(damage_porjutzi, N); DUMP_F64(ddamage_porjutzidt, N);
# endif
#endif
#if PALPHA_POROSITY
    DUMP_F64(pold, N); DUMP_F64(alpha_jutzi, N); DUMP_F64(alpha_jutzi_old, N); DUMP_F64(dalphadt, N);
    DUMP_F64(dalphadp, N); DUMP_F64(dalphadrho, N); DUMP_F64(f, N);
    DUMP_F64(delpdelrho, N); DUMP_F64(delpdele, N);
#endif
    if (with_lists) {
        int maxni = MAX_NUM_INTERACTIONS;
        char tag[32];
        int dtype = 1;
        int64_t one = 1;
        memset(tag, 0, sizeof(tag));
        strncpy(tag, "max_num_interactions", sizeof(tag) - 1);
        fwrite(tag, 1, sizeof(tag), f);
        fwrite(&dtype, sizeof(int), 1, f);
        fwrite(&one, sizeof(int64_t), 1, f);
        fwrite(&maxni, sizeof(int), 1, f);
        hook_write(f, "interactions", 1, N * (int64_t)MAX_NUM_INTERACTIONS, interactions);
    }
    fclose(f);
}

void euler()
{
    const char *prefix = getenv("REF_DUMP");
    const char *s_timed = getenv("REF_TIMED");
    const char *s_warm = getenv("REF_WARMUP");
    const char *s_lists = getenv("REF_DUMP_LISTS");
    int timed = s_timed ? atoi(s_timed) : 0;
    int warm = s_warm ? atoi(s_warm) : 3;
    int with_lists = s_lists ? atoi(s_lists) : 1;
    const size_t nbytes = (size_t)numberOfParticles * sizeof(double);
    int k;

    /* p and p_rhs are bound to p_device by initIntegration() */
    cudaVerify(cudaMemset(p_device.p, 0, nbytes));
    if (param.selfgravity) {
        cudaVerify(cudaMemset(p_device.g_ax, 0, nbytes));
#if DIM > 1
        cudaVerify(cudaMemset(p_device.g_ay, 0, nbytes));
#endif
#if DIM > 2
        cudaVerify(cudaMemset(p_device.g_az, 0, nbytes));
#endif
    }
    cudaVerify(cudaDeviceSynchronize());

    if (prefix) {
        hook_dump(prefix, "in", 0);
        rightHandSide();
        hook_dump(prefix, "out1", with_lists);
        rightHandSide();
        hook_dump(prefix, "out2", 0);
    }

    if (timed > 0) {
        cudaEvent_t t0, t1;
        float ms, total = 0.0f, best = 1e30f;
        cudaEventCreate(&t0);
        cudaEventCreate(&t1);
        for (k = 0; k < warm; k++) rightHandSide();
        cudaVerify(cudaDeviceSynchronize());
        for (k = 0; k < timed; k++) {
            cudaEventRecord(t0, 0);
            rightHandSide();
            cudaEventRecord(t1, 0);
            cudaEventSynchronize(t1);
            cudaEventElapsedTime(&ms, t0, t1);
            total += ms;
            if (ms < best) best = ms;
        }
        fprintf(stdout, "REF_TIMING n=%d calls=%d warmup=%d ms_per_call=%.6f best_ms=%.6f updates_per_s=%.6e\n",
                numberOfParticles, timed, warm, total / timed, best,
                (double)numberOfParticles * timed / (total * 1e-3));
        cudaEventDestroy(t0);
        cudaEventDestroy(t1);
    }
    fflush(stdout);
    /* leave before endIntegration() joins an I/O thread that was never started */
    cudaDeviceReset();
    exit(0);
}
