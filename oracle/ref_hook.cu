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
 *   0. REF_EVOLVE=1: before all of that, the reference's OWN rk2Adaptive() (src/rk2adaptive.cu:60-520)
 *      integrates the input over the command line's -n / -t / -M / -Q, so that dump and timing see
 *      a state after >= 20 accepted steps (SURVEY 8d second measurement point) instead of the
 *      pristine step-0 input.  p is then the reference's own (no H1 pin needed).
 *      REF_EVOLVE=pc does the same with predictor_corrector() (monaghan_pc).
 *
 * Dump container: repeated records { char name[32]; int32 dtype (0=f64,1=i32);
 * int64 count; payload }.  REF_DUMP_LISTS=2 stores the neighbour lists compacted
 * ("nbr_idx": the first noi[i] entries of every row, row after row) instead of the
 * dense N x MAX_NUM_INTERACTIONS array (1-2 GB at 10^6 particles).
 */
#include "miluph.h"
#include "timeintegration.h"
#include "rhs.h"
#include "parameter.h"
#include "memory_handling.h"
#include "pressure.h"

#include <stdint.h>
#include <stdio.h>
#include <time.h>
#include <stdlib.h>
#include <string.h>

extern int flag_force_gravity_calc;
extern int gravity_index;
extern pthread_t fileIOthread;

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
    if (with_lists == 2) {
        /* compacted rows, copied in slabs of 65536 particles */
        const int64_t slab = 65536;
        int *noi_h = (int *)malloc(sizeof(int) * (size_t)N);
        int *rows = (int *)malloc(sizeof(int) * (size_t)slab * MAX_NUM_INTERACTIONS);
        int64_t total = 0, i0, i;
        char tag[32];
        int dtype = 1;
        cudaVerify(cudaMemcpy(noi_h, p_device.noi, sizeof(int) * (size_t)N, cudaMemcpyDeviceToHost));
        for (i = 0; i < N; i++) total += noi_h[i];
        memset(tag, 0, sizeof(tag));
        strncpy(tag, "nbr_idx", sizeof(tag) - 1);
        fwrite(tag, 1, sizeof(tag), f);
        fwrite(&dtype, sizeof(int), 1, f);
        fwrite(&total, sizeof(int64_t), 1, f);
        for (i0 = 0; i0 < N; i0 += slab) {
            const int64_t cnt = (N - i0 < slab) ? N - i0 : slab;
            cudaVerify(cudaMemcpy(rows, interactions + i0 * MAX_NUM_INTERACTIONS, sizeof(int) * (size_t)cnt * MAX_NUM_INTERACTIONS,
                                  cudaMemcpyDeviceToHost));
            for (i = 0; i < cnt; i++) fwrite(rows + i * MAX_NUM_INTERACTIONS, sizeof(int), (size_t)noi_h[i0 + i], f);
        }
        free(rows);
        free(noi_h);
    } else if (with_lists) {
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

    const char *s_evolve = getenv("REF_EVOLVE");
    const int evolve = (s_evolve && s_evolve[0] && s_evolve[0] != '0') ? 1 : 0;

    if (evolve) {
        /* the reference's own integrator over -n / -t; afterwards p is bound to p_device again
         * (src/rk2adaptive.cu:320) and holds the integrated state */
        /* no particle file at the end of the integration (write_particles_to_file falls back to ASCII when neither
         * format is chosen, src/io.cu:1567-1570; HDF5 is compiled out in this build), only the small .info / log files */
        param.ascii_output = FALSE;
        param.hdf5output = TRUE;
        struct timespec ts0, ts1;   /* wall clock of the integrator call itself (tools/integrator_speed.py) */
        cudaVerify(cudaDeviceSynchronize());
        clock_gettime(CLOCK_MONOTONIC, &ts0);
        if (0 == strcmp(s_evolve, "pc")) {
            param.integrator_type = MONAGHAN_PC;
            predictor_corrector();
        } else {
            param.integrator_type = RK2_ADAPTIVE;
            rk2Adaptive();
        }
        cudaVerify(cudaDeviceSynchronize());
        clock_gettime(CLOCK_MONOTONIC, &ts1);
        fprintf(stdout, "REF_EVOLVE_WALL_MS=%.3f\n", (ts1.tv_sec - ts0.tv_sec) * 1e3 + (ts1.tv_nsec - ts0.tv_nsec) * 1e-6);
        if (currentDiskIO) pthread_join(fileIOthread, NULL);   /* the writer thread of the last output step */
        cudaVerify(cudaMemcpyToSymbol(p, &p_device, sizeof(struct Particle)));
        fprintf(stdout, "REF_EVOLVED t=%.17e\n", currentTime);
    }
    {
        /* REF_DEACTIVATE=<stride>: particles stride/2, stride/2 + stride, ... become deactivated the way a run
         * deactivates them (materialId = EOS_TYPE_IGNORE, src/boundary.cu:179-180, src/rk2adaptive.cu:1474) */
        const char *s_deact = getenv("REF_DEACTIVATE");
        const int stride = s_deact ? atoi(s_deact) : 0;
        if (stride > 0) {
            int *mat_h = (int *)malloc(sizeof(int) * (size_t)numberOfParticles);
            cudaVerify(cudaMemcpy(mat_h, p_device.materialId, sizeof(int) * (size_t)numberOfParticles, cudaMemcpyDeviceToHost));
            for (k = stride / 2; k < numberOfParticles; k += stride) mat_h[k] = EOS_TYPE_IGNORE;
            cudaVerify(cudaMemcpy(p_device.materialId, mat_h, sizeof(int) * (size_t)numberOfParticles, cudaMemcpyHostToDevice));
            free(mat_h);
        }
    }
    /* p and p_rhs are bound to p_device by initIntegration() */
    if (!evolve) cudaVerify(cudaMemset(p_device.p, 0, nbytes));
    if (param.selfgravity && !evolve) {
        cudaVerify(cudaMemset(p_device.g_ax, 0, nbytes));
#if DIM > 1
        cudaVerify(cudaMemset(p_device.g_ay, 0, nbytes));
#endif
#if DIM > 2
        cudaVerify(cudaMemset(p_device.g_az, 0, nbytes));
#endif
    }
    cudaVerify(cudaDeviceSynchronize());

    if (prefix && (getenv("REF_SEQ") || getenv("REF_DUMP_STATE_ONLY"))) {
        hook_dump(prefix, "in", 0);   /* the state only (REF_DUMP_STATE_ONLY: input preparation for bench.py) */
    } else if (prefix) {
        hook_dump(prefix, "in", 0);
        rightHandSide();
        hook_dump(prefix, "out1", with_lists);
        rightHandSide();
        hook_dump(prefix, "out2", 0);
    }

    {
        /* REF_SEQ=<K>: K consecutive rightHandSide() calls (the -g bookkeeping of src/rhs.cu:752-813 spans calls:
         * every 10th call and whenever > 0.1 % of the particles left their cell the walk runs, otherwise the
         * stored g_a is re-added); between calls every particle drifts by 0.002 h in x, and after call REF_SEQ_SHIFT_AT every
         * 50th particle jumps by 3 h.
         * Accelerations after every call go to <prefix>.seq<k>.bin. */
        const char *s_seq = getenv("REF_SEQ");
        const char *s_at = getenv("REF_SEQ_SHIFT_AT");
        const int nseq = s_seq ? atoi(s_seq) : 0, shift_at = s_at ? atoi(s_at) : -1;
        for (k = 0; prefix && k < nseq; k++) {
            char fname[1024];
            FILE *f;
            const int64_t N = numberOfParticles;
            rightHandSide();
            cudaVerify(cudaDeviceSynchronize());
            snprintf(fname, sizeof(fname), "%s.seq%d.bin", prefix, k);
            if ((f = fopen(fname, "wb")) == NULL) exit(1);
            DUMP_F64(ax, N);
#if DIM > 1
            DUMP_F64(ay, N);
#endif
#if DIM > 2
            DUMP_F64(az, N);
#endif
            if (param.selfgravity) {
                DUMP_F64(g_ax, N);
#if DIM > 1
                DUMP_F64(g_ay, N);
#endif
#if DIM > 2
                DUMP_F64(g_az, N);
#endif
            }
            DUMP_I32(noi, N);
            fclose(f);
            {
                /* every particle drifts by 0.002 h per call (far below a cell: the stored g_a is re-added, and it differs
                 * measurably from a fresh walk); after call shift_at every 50th particle jumps by 3 h */
                double *xh = (double *)malloc(nbytes), *hh = (double *)malloc(nbytes);
                int i;
                cudaVerify(cudaMemcpy(xh, p_device.x, nbytes, cudaMemcpyDeviceToHost));
                cudaVerify(cudaMemcpy(hh, p_device.h, nbytes, cudaMemcpyDeviceToHost));
                for (i = 0; i < numberOfParticles; i++) xh[i] += 0.002 * hh[i];
                if (k == shift_at)
                    for (i = 0; i < numberOfParticles; i += 50) xh[i] += 3.0 * hh[i];
                cudaVerify(cudaMemcpy(p_device.x, xh, nbytes, cudaMemcpyHostToDevice));
                free(xh); free(hh);
            }
        }
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
