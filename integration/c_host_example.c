/*
 * c_host_example.c -- a host in plain C on the C-ABI of libb200sph (include/b200sph.h), no Python, no CUDA code of its own.
 *
 * What a C host such as miluphcuda's main() does around the hot path: parse material.cfg, create a handle for the
 * switch set it was compiled for, hand over its particle buffers and call the right-hand side.  Here the buffers are
 * HOST arrays (b200sph_rhs_eval_host copies in and out); integration/rhs_b200.cu is the variant that binds the
 * reference's own DEVICE buffers.  Built and run by tests/test_c_host_example.py against libb200sph_sedov.so:
 *
 *     gcc -std=c99 -Iinclude integration/c_host_example.c -Lmiluphcuda_b200/lib -lb200sph_sedov -lm -o c_host_example
 *
 * The particle set is a periodic-looking block of an ideal gas on a cubic lattice with a hot centre (a small Sedov
 * problem); the program checks what must hold for any correct evaluation -- symmetric neighbour counts in the bulk and a
 * vanishing total force (the pair forces are antisymmetric) -- prints one line and returns 0 on success.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200sph.h"

#define CHECK(call)                                                                             \
    do {                                                                                        \
        int rc_ = (call);                                                                       \
        if (rc_ != B200SPH_OK) {                                                                \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, b200sph_last_error(handle));    \
            return 1;                                                                           \
        }                                                                                       \
    } while (0)

int main(int argc, char **argv)
{
    const int side = (argc > 1) ? atoi(argv[1]) : 30;
    const int n = side * side * side;
    const double delta = 1.0 / side, sml = 2.2 * delta;
    const char *cfg_path = (argc > 2) ? argv[2] : "c_host_material.cfg";
    b200sph_handle *handle = NULL;
    b200sph_materials *mat = NULL;
    b200sph_view view;
    double grav = 0.0, fsum[3] = {0, 0, 0}, fabs_sum = 0.0;
    char err[512];
    int64_t h2d = 0, d2h = 0;
    int i, k, offender = -1, noi_min = 1 << 30, noi_max = 0;

    FILE *f = fopen(cfg_path, "w");
    if (!f) return 1;
    fprintf(f, "materials = (\n  {\n    ID = 0\n    name = \"gas\"\n    sml = %.17e\n"
               "    artificial_viscosity = { alpha = 1.0; beta = 2.0; };\n"
               "    eos = {\n      type = 9\n      polytropic_gamma = 1.4\n    };\n  }\n);\n", sml);
    fclose(f);
    if (b200sph_materials_load(cfg_path, &mat, &grav, err, sizeof err) != B200SPH_OK) {
        fprintf(stderr, "material.cfg: %s\n", err);
        return 1;
    }
    if (b200sph_switch_value("DIM") != 3 || b200sph_switch_value("HYDRO") != 1) {
        fprintf(stderr, "this example expects the sedov switch set (3-D hydro), got %s\n", b200sph_config_name());
        return 1;
    }
    CHECK(b200sph_create(&handle, n, 0, b200sph_switch_hash()));
    CHECK(b200sph_set_materials(handle, mat));

    memset(&view, 0, sizeof view);
    view.n = view.n_real = n;
    view.max_num_flaws = 1;
    view.theta = 0.5;
    view.grav_const = grav;
#define ALLOC_D(field) view.p.field = (double *)calloc((size_t)n, sizeof(double))
    ALLOC_D(x); ALLOC_D(y); ALLOC_D(z); ALLOC_D(vx); ALLOC_D(vy); ALLOC_D(vz); ALLOC_D(dxdt); ALLOC_D(dydt); ALLOC_D(dzdt);
    ALLOC_D(ax); ALLOC_D(ay); ALLOC_D(az); ALLOC_D(m); ALLOC_D(h); ALLOC_D(rho); ALLOC_D(drhodt); ALLOC_D(p); ALLOC_D(e); ALLOC_D(dedt);
    ALLOC_D(cs); ALLOC_D(muijmax);
    view.p.noi = (int *)calloc((size_t)n, sizeof(int));
    view.p.depth = (int *)calloc((size_t)n, sizeof(int));
    view.p.h0 = (double *)calloc((size_t)n, sizeof(double));
    view.p.materialId = (int *)calloc((size_t)n, sizeof(int));
    view.p_rhs = view.p;   /* the reference's p_rhs is p_device: the same buffers */
    for (i = 0; i < n; i++) {
        const int ix = i % side, iy = (i / side) % side, iz = i / (side * side);
        const double x = (ix + 0.5) * delta - 0.5, y = (iy + 0.5) * delta - 0.5, z = (iz + 0.5) * delta - 0.5;
        const double r = sqrt(x * x + y * y + z * z);
        view.p.x[i] = x; view.p.y[i] = y; view.p.z[i] = z;
        view.p.m[i] = delta * delta * delta;
        view.p.h[i] = view.p.h0[i] = sml;
        view.p.e[i] = 1e-3 + ((r < 3.0 * delta) ? 1.0 : 0.0);
    }
    /* two evaluations: the sound speed of the second sees the pressure of the first */
    for (k = 0; k < 2; k++) CHECK(b200sph_rhs_eval_host(handle, &view, &offender, &h2d, &d2h));

    for (i = 0; i < n; i++) {
        fsum[0] += view.p.m[i] * view.p.ax[i];
        fsum[1] += view.p.m[i] * view.p.ay[i];
        fsum[2] += view.p.m[i] * view.p.az[i];
        fabs_sum += view.p.m[i] * (fabs(view.p.ax[i]) + fabs(view.p.ay[i]) + fabs(view.p.az[i]));
        if (view.p.noi[i] < noi_min) noi_min = view.p.noi[i];
        if (view.p.noi[i] > noi_max) noi_max = view.p.noi[i];
    }
    {
        b200sph_stats st;
        const double resid = (fabs(fsum[0]) + fabs(fsum[1]) + fabs(fsum[2])) / (fabs_sum > 0.0 ? fabs_sum : 1.0);
        CHECK(b200sph_get_stats(handle, &st));
        printf("C_HOST n=%d launches=%d ms=%.3f noi=[%d,%d] total_force_residual=%.3e h2d=%lld d2h=%lld\n", n, st.kernel_launches,
               st.ms_total, noi_min, noi_max, resid, (long long)h2d, (long long)d2h);
        b200sph_destroy(handle);
        b200sph_materials_free(mat);
        if (!(fabs_sum > 0.0) || resid > 1e-10 || noi_max < 30 || noi_max > 80 || st.kernel_launches < 5) return 2;
    }
    return 0;
}
