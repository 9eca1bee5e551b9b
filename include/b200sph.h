/*
 * b200sph.h -- C-ABI of the B200-native SPH right-hand side.
 *
 * This is the drop-in boundary for miluphcuda's per-step hot path: everything
 * the reference's `void rightHandSide(void)` (reference: include/rhs.h:30,
 * src/rhs.cu:143-861) launches is replaced by `b200sph_rhs_eval()`.  The
 * reference passes its arguments through globals: the `__constant__ struct
 * Particle p` bound by the integrator right before each call (e.g.
 * src/rk2adaptive.cu:219,289,308), `p_rhs` (always `p_device`,
 * src/timeintegration.cu:200-201) and the `mat*` material arrays
 * (include/config_parameter.h:40-290).  Here they are explicit: plain structs
 * of pointers and scalars, no C++/torch types.
 *
 * One shared library is built per compile-time switch set, exactly like the
 * reference binary is built against one parameter.h (`libb200sph_<config>.so`);
 * `b200sph_switch_hash()` lets the caller check both sides agree.
 *
 * All entry points return 0 on success, a negative value for CUDA/runtime
 * errors and a positive value for model errors (reference error convention:
 * include/cuda_utils.h:29-48 exits; src/tree.cu:917 asserts):
 *   1  B200SPH_ERR_TOO_MANY_INTERACTIONS  (out-param: first offending particle)
 *   2  B200SPH_ERR_BAD_ARGUMENT
 *   3  B200SPH_ERR_SWITCH_MISMATCH
 *   4  B200SPH_ERR_UNSUPPORTED            (switch / EOS outside the hot-path scope)
 *   6  B200SPH_ERR_ABORTED                (multi-GPU: the device flag of b200sph_set_abort_flag() was raised, e.g. by a
 *                                          stale halo plan; the evaluation left the caller's state untouched)
 *   5  B200SPH_ERR_NONFINITE              (a particle coordinate or smoothing length is NaN/Inf; the reference
 *                                          runs out of tree nodes in that situation, src/tree.cu:190-193)
 * `b200sph_last_error()` returns a human-readable description.
 */
#ifndef B200SPH_H
#define B200SPH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200SPH_ABI_VERSION 1

#define B200SPH_OK 0
#define B200SPH_ERR_TOO_MANY_INTERACTIONS 1
#define B200SPH_ERR_BAD_ARGUMENT 2
#define B200SPH_ERR_SWITCH_MISMATCH 3
#define B200SPH_ERR_UNSUPPORTED 4
#define B200SPH_ERR_NONFINITE 5
#define B200SPH_ERR_ABORTED 6    /* the caller's abort flag was set: nothing was modified, evaluate again */
#define B200SPH_ERR_CUDA (-1)

/* Field-for-field mirror of the in-scope members of the reference's
 * `struct Particle` (reference: include/miluph.h:67-275), with a layout that
 * does not depend on the switch set: members a switch set does not have are
 * NULL.  Scalars are arrays of n doubles/ints in the CALLER's particle order;
 * tensors are n*DIM*DIM doubles, row-major per particle
 * (src/timeintegration.cu:109-112); flaws are n*maxNumFlaws doubles.
 * x,y,z,m may be the reference's tree-sized arrays; only [0,n) is touched. */
typedef struct b200sph_particle_arrays {
    double *x, *y, *z;
    double *vx, *vy, *vz;
    double *dxdt, *dydt, *dzdt;
    double *ax, *ay, *az;
    double *g_ax, *g_ay, *g_az;
    double *g_local_cellsize, *g_x, *g_y, *g_z;
    double *m, *h, *h0, *dhdt;
    double *rho, *drhodt, *p, *e, *dedt;
    double *S, *dSdt, *local_strain, *ep, *edotp, *plastic_f, *sigma;
    double *R;
    double *d, *damage_total, *dddt;
    int *numFlaws, *numActiveFlaws;
    double *flaws;
    double *damage_porjutzi, *ddamage_porjutzidt;
    double *muijmax;
    double *pold, *alpha_jutzi, *alpha_jutzi_old, *dalphadt, *dalphadp, *dalphadrho, *f, *delpdelrho, *delpdele;
    double *tensorialCorrectionMatrix;
    double *cs;
    int *noi, *materialId, *depth;
} b200sph_particle_arrays;

/* What one rightHandSide() call sees (reference globals: src/miluph.cu:50-63,
 * include/miluph.h:366-421, src/timeintegration.cu:52-53,80-81). */
typedef struct b200sph_view {
    int n;                      /* numberOfParticles */
    int n_real;                 /* numberOfRealParticles (== n without ghost particles) */
    int max_num_flaws;          /* maxNumFlaws_host */
    int selfgravity;            /* param.selfgravity  (-s) */
    int decouplegravity;        /* param.decouplegravity (-g) */
    int is_relaxation_run;      /* isRelaxationRun */
    double theta;               /* treeTheta (-a) */
    double grav_const;          /* gravConst (material.cfg global.c_gravity) */
    b200sph_particle_arrays p;      /* the bound state+derivative buffer (`p`) */
    b200sph_particle_arrays p_rhs;  /* `p_rhs` == p_device: materialId, h0, flaws, sigma,
                                       tensorialCorrectionMatrix, plastic_f, R, g_x/y/z, g_local_cellsize */
} b200sph_view;

/* Per-material tables, one entry per material ID, named after the reference's
 * device symbols (include/config_parameter.h:180-290; filled from material.cfg
 * in src/config_parameter.cu:357-878).  Pointers may be host or device memory
 * (copied with cudaMemcpyDefault); unused tables may be NULL (read as 0). */
typedef struct b200sph_materials {
    int n_materials;
    const int *matEOS;
    const double *matSml;
    const double *mat_f_sml_min, *mat_f_sml_max;
    const double *matAlpha, *matBeta;
    const double *matPolytropicK, *matPolytropicGamma, *matIsothermalSoundSpeed;
    const double *matBulkmodulus, *matShearmodulus, *matYoungModulus, *matYieldStress;
    const double *matRho0, *matN, *matRhoLimit, *matcsLimit;
    const double *matTillRho0, *matTillA, *matTillB, *matTillE0, *matTillEiv, *matTillEcv;
    const double *matTilla, *matTillb, *matTillAlpha, *matTillBeta;
    const double *matCohesion, *matCohesionDamaged, *matInternalFriction, *matInternalFrictionDamaged;
    const double *matMeltEnergy;
    const double *matDensityFloor, *matEnergyFloor;
    const int *matdensity_via_kernel_sum;
    const double *matexponent_tensor, *matepsilon_stress, *matmean_particle_distance;
    const double *matporjutzi_p_elastic, *matporjutzi_p_transition, *matporjutzi_p_compacted;
    const double *matporjutzi_alpha_0, *matporjutzi_alpha_e, *matporjutzi_alpha_t;
    const double *matporjutzi_n1, *matporjutzi_n2;
    const double *matcs_porous, *matcs_solid;
    const int *matcrushcurve_style;
    /* tabulated EOS in ANEOS format (reference: src/aneos.cu:119-181; include/config_parameter.h:140-163):
     * per material start offsets into the concatenated axes / linearised [i_rho*n_e + i_e] tables, -1 if not ANEOS */
    const int *aneos_n_rho, *aneos_n_e, *aneos_rho_id, *aneos_e_id, *aneos_matrix_id;
    const double *aneos_rho, *aneos_e, *aneos_p, *aneos_cs;
    const double *aneos_bulk_cs, *aneos_gamma;
    int64_t aneos_rho_len, aneos_e_len, aneos_matrix_len;
} b200sph_materials;

/* Launch/timing counters of the most recent b200sph_rhs_eval*(). */
typedef struct b200sph_stats {
    int kernel_launches;        /* kernels of this library launched by the call */
    int n_cells;                /* cells of the search grid */
    int max_noi;                /* max number of interactions found */
    int64_t total_noi;          /* sum of noi */
    double cell_size;
    float ms_total;             /* device time of the call (cudaEvent pair on the library stream) */
    float ms_sort, ms_neighbours, ms_density, ms_pointwise, ms_correction, ms_forces, ms_gravity, ms_scatter;
    int gravity_recomputed;     /* 1 if the Barnes-Hut walk ran, 0 if the stored g_a was re-added */
} b200sph_stats;

typedef struct b200sph_handle b200sph_handle;

int b200sph_abi_version(void);
/* name of the switch set this library was built for ("sedov", "impact", ...) */
const char *b200sph_config_name(void);
/* FNV-1a hash over "NAME=value;" of every switch in miluphcuda_b200/csrc/switches.h */
uint64_t b200sph_switch_hash(void);
/* value of one compile-time switch by its parameter.h name, -999 if unknown */
int b200sph_switch_value(const char *name);

/* Owns every scratch buffer (sorted SoA copies, cell grid, neighbour lists, tree); no allocation
 * happens inside rhs_eval (reference: scratch is cudaMalloc'ed per call, src/rhs.cu:293,760). */
int b200sph_create(b200sph_handle **out, int n_max, int device, uint64_t expected_switch_hash);
int b200sph_destroy(b200sph_handle *h);
const char *b200sph_last_error(const b200sph_handle *h);

/* Host-side reader of material.cfg (libconfig format) -> tables with the reference's keys,
 * defaults and derived values (replaces the parsing half of transferMaterialsToGPU(),
 * src/config_parameter.cu:346-878; ANEOS tables as src/aneos.cu:119-181).  The returned
 * object owns its arrays (host memory); release with b200sph_materials_free(). */
int b200sph_materials_load(const char *cfg_path, b200sph_materials **out, double *grav_const,
                           char *err, size_t errlen);
void b200sph_materials_free(b200sph_materials *m);

/* replaces transferMaterialsToGPU()'s uploads (src/config_parameter.cu:880-1400) */
int b200sph_set_materials(b200sph_handle *h, const b200sph_materials *mat);

/* The hot path.  Pointers in `view` are DEVICE pointers.  The call is complete on return
 * (the reference synchronises after every kernel, e.g. src/rhs.cu:182-840).
 * `offender` (may be NULL) receives the caller index of the first particle that exceeded
 * MAX_NUM_INTERACTIONS when the return value is B200SPH_ERR_TOO_MANY_INTERACTIONS. */
int b200sph_rhs_eval(b200sph_handle *h, const b200sph_view *view, int *offender);

/* Same call with HOST pointers in `view`: state fields are copied host->device, the RHS runs,
 * and every field the integrators/writer read back is copied device->host
 * (reference: copy_particle_data_to_device, src/memory_handling.cu:1108;
 * copyToHostAndWriteToFile, src/io.cu:3066-3162).  Byte counts are returned in h2d/d2h. */
int b200sph_rhs_eval_host(b200sph_handle *h, const b200sph_view *host_view, int *offender,
                          int64_t *h2d_bytes, int64_t *d2h_bytes);

/* Options of the host-buffer call (default 0: every input is uploaded and every output member is
 * read back on every call).
 *   CACHE_IMMUTABLES  m, h0, materialId, numFlaws and flaws are uploaded on the first call only and
 *                     reused while the same host arrays and particle count are passed -- the reference
 *                     uploads them once per run as well (allocate_immutables, src/memory_handling.cu:111-122;
 *                     copy_particles_immutables_device_to_device, src/memory_handling.cu:373-392).
 *                     Calling b200sph_host_options() again drops the cached copy.
 *   SKIP_SCRATCH      sigma, R, plastic_f and tensorialCorrectionMatrix (p_rhs scratch that no integrator
 *                     or writer reads) are not copied back. */
#define B200SPH_HOST_CACHE_IMMUTABLES 1
#define B200SPH_HOST_SKIP_SCRATCH 2
int b200sph_host_options(b200sph_handle *h, int options);

/* Cold calls the reference makes outside rightHandSide():
 * calculatePressure by the writer and the PC integrators (src/io.cu:3039, src/predictor_corrector.cu:829),
 * damageLimit at output (src/rk2adaptive.cu:468), initializeSoundspeed at start (src/timeintegration.cu:206). */
int b200sph_pressure(b200sph_handle *h, const b200sph_view *view);
int b200sph_damage_limit(b200sph_handle *h, const b200sph_view *view);
int b200sph_init_soundspeed(b200sph_handle *h, const b200sph_view *view);

/* Persistent cell order (the reference keeps the input file's order for the whole run,
 * src/memory_handling.cu, and pays for it with uncoalesced accesses in every kernel).  Puts every non-NULL member of
 * view->p, view->p_rhs and of the n_extra further buffers (the integrator's rk_device[3]) into the order of the search
 * cells, all with the SAME permutation: new[k] = old[perm[k]].  Call it at a step boundary, every few hundred steps;
 * perm_out (device, n ints, may be NULL) receives the permutation so that a writer can restore the input order
 * (src/io.cu).  Results of later evaluations are the same particles' results, relabelled. */
int b200sph_reorder(b200sph_handle *h, const b200sph_view *view, const b200sph_particle_arrays *extra, int n_extra,
                    int *perm_out);

/* ---- rk2_adaptive on the device: the embedded Runge-Kutta 2/3 step of src/rk2adaptive.cu:197-485
 * with its update kernels (integrateFirst/Second/ThirdStep :700-1130), reductions (limitTimestepCourant/Forces/Damage
 * :521-695, checkError :1134-1482) and the device-to-device copies between p_device and rk_device[3]
 * (src/memory_handling.cu:253-522), fused into six streaming kernels around three b200sph_rhs_eval() calls.
 * Buffers keep the reference's meaning: view->p = p_device, rk[0] = rk_device[RKSTART], rk[1] = [RKFIRST],
 * rk[2] = [RKSECOND]; all device pointers, owned by the caller. ---- */
typedef struct b200sph_rk2_params {
    double rk_epsrel;               /* -Q, param.rk_epsrel */
    double dt_max;                  /* -M, param.maxtimestep; 0 = the output interval */
    double first_dt;                /* -F, param.firsttimestep; 0 = unset */
    /* the RK2_* compile-time switches of include/rk2adaptive.h:39-71, with the shipped values as defaults */
    int use_courant_limit, use_forces_limit, use_damage_limit;
    int use_velocity_error, use_density_error, use_energy_error;
    int limit_pressure_change, limit_alpha_change;
    double courant_fact, forces_fact;           /* COURANT_FACT, FORCES_FACT (include/timeintegration.h:41-43) */
    double location_safety, min_vel_change, tiny_density, tiny_energy, timestep_safety, smallest_dt_allowed;
    double max_damage_change, max_alpha_change, max_pressure_change;
} b200sph_rk2_params;

typedef struct b200sph_rk2_state {
    double t;                       /* currentTime */
    double dt;                      /* step the next attempt uses (dt_host) */
    double dt_suggested;            /* what the error control suggested last */
    double dt_done;                 /* size of the last accepted step */
    int accepted, rejected, rhs_calls, intervals, approaching_output_time;
    double err[6];                  /* last max errors: position, velocity, density, energy, alpha change, pressure change */
} b200sph_rk2_state;

int b200sph_rk2_default_params(b200sph_rk2_params *prm);
/* copy_particles_immutables_device_to_device (src/memory_handling.cu:373-392): m, h, cs, numFlaws of p_device into the
 * three rk buffers; call once after the buffers are allocated (src/rk2adaptive.cu:116-124). */
int b200sph_rk2_init(b200sph_handle *h, const b200sph_view *view, const b200sph_particle_arrays rk[3]);
/* one ACCEPTED step (rejected attempts are repeated inside); state->dt is the step to try */
int b200sph_rk2_step(b200sph_handle *h, const b200sph_view *view, const b200sph_particle_arrays rk[3],
                     const b200sph_rk2_params *prm, double t_end, b200sph_rk2_state *state, int *offender);
/* one output interval [state->t, t_end]: first-step / continuing-step rules of src/rk2adaptive.cu:153-171, steps until
 * t_end, damageLimit before the output (src/rk2adaptive.cu:464-469).  Zero-initialise `state` before the first call. */
int b200sph_rk2_advance(b200sph_handle *h, const b200sph_view *view, const b200sph_particle_arrays rk[3],
                        const b200sph_rk2_params *prm, double t_end, b200sph_rk2_state *state, int *offender);

/* The numbers of conserved_quantities.log (src/io.cu:1661-1838, 1980-2017), summed on the device in one pass instead
 * of on the host after copying every array back.  Deactivated particles are not counted. */
typedef struct b200sph_conserved {
    double mass, e_kin, e_int;
    double p_abs, p[3];             /* linear momentum */
    double L_abs, L[3];             /* angular momentum about the origin (2-D: L[0] is the z component, as the reference stores it) */
    double bary_pos[3], bary_vel[3];
    int n_ignored;
} b200sph_conserved;
int b200sph_conserved_quantities(b200sph_handle *h, const b200sph_view *view, b200sph_conserved *out);

/* Neighbour lists of the last rhs_eval in the caller's indexing, for parity checks and for the
 * reference writer's /number_of_interactions: row i holds noi[i] neighbour ids (unordered).
 * `interactions` is a device buffer of n*max_per_row ints (the reference's dense layout,
 * src/tree.cu:866), unused slots are set to -1 (src/rhs.cu:115-118). */
int b200sph_export_interactions(b200sph_handle *h, int *interactions, int max_per_row);

int b200sph_get_stats(const b200sph_handle *h, b200sph_stats *out);

/* Run on the caller's CUDA stream (a cudaStream_t passed as void*; NULL = the legacy default
 * stream the reference uses for everything) instead of the library's own stream. */
int b200sph_set_stream(b200sph_handle *h, void *cuda_stream);

/* Multi-GPU.  Each rank owns a contiguous slab of the global search grid; halo
 * particles received from neighbouring ranks are appended after the owned ones ([n_owned, n)).
 * rhs_eval computes all rates for owned particles only; pointwise quantities are recomputed on
 * halo copies.  n_owned == 0 or == n means single-GPU. */
int b200sph_set_owned(b200sph_handle *h, int n_owned);
/* Global bounding box (allreduced by the host) so every rank builds the same cells/tree. */
int b200sph_set_global_domain(b200sph_handle *h, const double lo[3], const double hi[3]);
/* ---- halo exchange, device side (csrc/halo.cu).  Every rank's domain is a set of axis-aligned boxes (the
 * octree cells of its Morton key range).  All calls are stream-ordered on the handle's stream and do not
 * synchronise; pointers are device pointers unless stated otherwise. ---- */

/* boxes: n_boxes x 6 doubles {lo_x, lo_y, lo_z, hi_x, hi_y, hi_z}, box_rank[b] = owner of box b, non-decreasing
 * (HOST arrays); my_rank's boxes are the caller's own domain.  Call once per decomposition. */
int b200sph_halo_set_domains(b200sph_handle *h, const double *boxes, const int *box_rank, int n_boxes, int n_ranks, int my_rank);
/* hmax_out[j] <- largest sml among the n particles lying inside the caller's j-th box (0 if none);
 * hmax_len >= number of own boxes, the tail is zeroed (so equal-sized pieces can be all-gathered). */
int b200sph_halo_box_hmax(b200sph_handle *h, const double *x, const double *y, const double *z, const double *sml, int n,
                          double *hmax_out, int hmax_len);
/* Particle k is needed by rank r when its distance to one of r's boxes b is below sml[k] + extra(b), where
 * extra(b) = extra[r * extra_stride + j] for r's j-th box (the all-gather of every rank's hmax_out: two-level
 * halo) or 0 when extra == NULL (one-level halo).  idx_out receives the indices of the particles to send,
 * grouped by destination rank in rank order, ascending inside a group; counts_out[r] (n_ranks + 1 ints) the
 * group sizes, counts_out[n_ranks] != 0 if idx_capacity was too small. */
int b200sph_halo_select(b200sph_handle *h, const double *x, const double *y, const double *z, const double *sml, int n,
                        const double *extra, int extra_stride, int *idx_out, int idx_capacity, int *counts_out);
/* Selection for a REUSABLE plan: reach = (h_k + extra) * reach_scale + skin.  A list built with
 * reach_scale = 1 + growth and skin = 2 * max_move stays complete while b200sph_halo_plan_check() reports no violation. */
int b200sph_halo_select_plan(b200sph_handle *h, const double *x, const double *y, const double *z, const double *sml, int n,
                             const double *extra, int extra_stride, double reach_scale, double skin, int *idx_out,
                             int idx_capacity, int *counts_out);
/* *flag_out (device int) = 1 if a particle is further than max_move from its position (x0, y0, z0) at plan time or
 * its smoothing length exceeds sml0 * (1 + growth), else 0.  Stream-ordered, no host synchronisation. */
int b200sph_halo_plan_check(b200sph_handle *h, const double *x, const double *y, const double *z, const double *sml,
                            const double *x0, const double *y0, const double *z0, const double *sml0, int n,
                            double max_move, double growth, int *flag_out);
/* One member of the reference's struct Particle taking part in the exchange: `per` values per particle;
 * kind 0 = double, 1 = int32 (transported as double), 2 = int32 that is not transported but zeroed on the
 * received rows (numFlaws, numActiveFlaws: the flaw lists of halo copies are never read). */
typedef struct b200sph_halo_field {
    void *data;
    int per;
    int kind;
} b200sph_halo_field;
/* doubles per packed row */
int b200sph_halo_row_width(const b200sph_halo_field *fields, int n_fields);
/* out[row * width + col] <- state of particle idx[row] (fields is a HOST array of n_fields descriptors) */
int b200sph_halo_pack(b200sph_handle *h, const b200sph_halo_field *fields, int n_fields, const int *idx, int n_rows, double *out);
/* particle rows [first_row, first_row + n_rows) <- in[row * width + col] */
int b200sph_halo_unpack(b200sph_handle *h, const b200sph_halo_field *fields, int n_fields, const double *in, int n_rows, int first_row);
/* The same with the block of every rank stored column by column (a warp then walks one member array):
 * counts[r] (device, n_ranks ints) = rows going to / coming from rank r, in rank order; a rank's block starts at
 * (rows of lower ranks) * width and holds its columns one after the other. */
int b200sph_halo_pack_by_rank(b200sph_handle *h, const b200sph_halo_field *fields, int n_fields, const int *idx,
                              const int *counts, int n_ranks, int n_rows, double *out);
int b200sph_halo_unpack_by_rank(b200sph_handle *h, const b200sph_halo_field *fields, int n_fields, const double *in,
                                const int *counts, int n_ranks, int n_rows, int first_row);
/* ---- neighbour-sum exchange.  With external halo sums a halo copy needs no neighbours of its own:
 * its kernel-sum density and its tensorial correction matrix are computed by its owner and delivered between the
 * stages of the evaluation, so ONE halo level suffices (half the copies of the two-level halo, and none of them is
 * searched for).  b200sph_rhs_eval_stage(.., 0, ..) runs hooks, sort, search and the density sum of the owned
 * particles and reports B200SPH_SUM_DENSITY when the owners' view->p.rho must now be delivered into the halo rows;
 * stage 1 runs the pointwise chain and the correction matrices (B200SPH_SUM_CORRECTION: tensorialCorrectionMatrix);
 * stage 2 the pair forces, gravity and the end-of-call checks.  Stages 0 and 1 only enqueue work on the handle's
 * stream: a stream-ordered exchange (b200sph_halo_pack_by_rank -> NCCL -> b200sph_halo_unpack_by_rank) needs no host
 * synchronisation in between. ---- */
#define B200SPH_SUM_DENSITY 1
#define B200SPH_SUM_CORRECTION 2
int b200sph_set_halo_sums(b200sph_handle *h, int external);
int b200sph_rhs_eval_stage(b200sph_handle *h, const b200sph_view *view, int stage, int *pending_sum, int *offender);
/* Device flag (an int in device memory, may be NULL to clear) that every state-modifying kernel of an evaluation reads
 * first: when it is non-zero the evaluation does nothing and returns B200SPH_ERR_ABORTED.  The multi-GPU host points
 * it at the all-reduced verdict of b200sph_halo_plan_check(): a stale send plan is then discovered at the end-of-call
 * synchronisation that exists anyway, instead of by a host wait before every evaluation. */
int b200sph_set_abort_flag(b200sph_handle *h, const int *device_flag);
/* margins of a reusable send plan for the copies that still need their own lists (two-level halo without the
 * neighbour-sum exchange): reach = h * reach_scale + skin, the same numbers the plan was selected with */
int b200sph_halo_set_list_margin(b200sph_handle *h, double reach_scale, double skin);

/* Multi-GPU self-gravity with a replicated tree.  x,y,z,m (device pointers, n_sources doubles each; y/z may
 * be NULL below DIM 2/3) describe the WHOLE particle set in a rank-independent order, normally the
 * all-gather of every rank's owned particles; the caller's owned particles are the block
 * [own_begin, own_begin + n_owned) of it and correspond to view.p[0, n_owned).  Every rank builds the same
 * root cube and cells from the same data (the reference's geometry, src/tree.cu:1071-1086) and
 * walks the tree for its own particles only (src/gravity.cu:382-499).  The pointers must stay valid during
 * b200sph_rhs_eval; n_sources = 0 returns to single-GPU behaviour. */
int b200sph_set_gravity_sources(b200sph_handle *h, const double *x, const double *y, const double *z, const double *m,
                                int n_sources, int own_begin);

/* ---- the multi-GPU host behind the C-ABI (csrc/mg.cu): one b200sph_mg per GPU/process, NCCL bound at run time.
 * A C host that holds a particle set spread over the GPUs of one box calls
 *     rank 0: b200sph_mg_unique_id(id)  -> hands the 128 bytes to every rank (MPI_Bcast, a file, ...)
 *     all   : b200sph_mg_create(&mg, handle, rank, world, id)
 *             b200sph_mg_decompose(mg, view, n_held, by_work)   Morton-curve domains of equal count / equal sum(noi)
 *             b200sph_mg_migrate(mg, view, rk, 3, n_held, capacity, &n_owned)   records move to their owners
 *     per rightHandSide():
 *             b200sph_mg_rhs_eval(mg, view, n_owned, capacity, &n_total, &offender)
 * and repeats decompose + migrate every few hundred steps, at a step boundary (the halo plan tolerates a drift of
 * 0.15 h_min out of the own boxes, no more).  All arrays of `view` (and of the extra buffers) have room for `capacity`
 * rows; rows [0, n_owned) are the rank's own particles, the halo copies are written behind them. ---- */
typedef struct b200sph_mg b200sph_mg;
typedef struct b200sph_mg_stats {
    int n_boxes, n_halo, plan_builds, stale_plans, sum_exchanges;
    long long migrated_out, migrated_in, halo_bytes_sent;
} b200sph_mg_stats;
int b200sph_mg_unique_id(void *id128, char *err, size_t errlen);
int b200sph_mg_create(b200sph_mg **out, b200sph_handle *h, int rank, int world, const void *id128);
int b200sph_mg_destroy(b200sph_mg *mg);
const char *b200sph_mg_last_error(const b200sph_mg *mg);
int b200sph_mg_decompose(b200sph_mg *mg, const b200sph_view *view, int n_held, int weight_by_interactions);
int b200sph_mg_migrate(b200sph_mg *mg, const b200sph_view *view, const b200sph_particle_arrays *extra, int n_extra, int n_held,
                       int capacity, int *n_held_out);
int b200sph_mg_rhs_eval(b200sph_mg *mg, const b200sph_view *view, int n_owned, int capacity, int *n_total_out, int *offender);
/* rk2_adaptive over the GPUs of one box: b200sph_rk2_advance with the evaluation going through b200sph_mg_rhs_eval and
 * the step-size reductions (limitTimestepCourant/Damage: min; checkError: max, src/rk2adaptive.cu:521-695,1134-1482)
 * all-reduced over the ranks, so that every rank takes the same steps.  Call b200sph_rk2_init for the owned rows first. */
int b200sph_mg_rk2_advance(b200sph_mg *mg, const b200sph_view *view, const b200sph_particle_arrays rk[3],
                           const b200sph_rk2_params *prm, double t_end, b200sph_rk2_state *state, int n_owned, int capacity,
                           int *offender);
int b200sph_mg_get_stats(const b200sph_mg *mg, b200sph_mg_stats *out);

#ifdef __cplusplus
}
#endif
#endif /* B200SPH_H */
