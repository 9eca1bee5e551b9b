/*
 * switches.h -- compile-time switch set of one libb200sph_<config>.so.
 *
 * The reference is configured by editing include/parameter.h and recompiling
 * (reference: include/parameter.h, include/checks.h:27-29).  This library keeps
 * that contract: it is compiled with -I<dir containing the scenario's
 * parameter.h>; every reference switch that the file does not mention
 * defaults to 0 here.  Switches that select code outside the hot-path scope
 * ("OUT OF SCOPE") are rejected at compile time instead of
 * being silently ignored.
 */
#ifndef B200SPH_SWITCHES_H
#define B200SPH_SWITCHES_H

#include "parameter.h"

#ifndef DIM
#error parameter.h must define DIM
#endif
#ifndef MAX_NUM_INTERACTIONS
#error parameter.h must define MAX_NUM_INTERACTIONS
#endif

#define B200SPH_SWITCH_LIST(X) \
    X(DIM) X(SOLID) X(HYDRO) X(REAL_HYDRO) X(INTEGRATE_ENERGY) X(INTEGRATE_DENSITY) \
    X(FRAGMENTATION) X(DAMAGE_ACTS_ON_S) X(SPH_EQU_VERSION) X(ARTIFICIAL_STRESS) \
    X(ARTIFICIAL_VISCOSITY) X(TENSORIAL_CORRECTION) X(VON_MISES_PLASTICITY) \
    X(COLLINS_PLASTICITY) X(COLLINS_PLASTICITY_INCLUDE_MELT_ENERGY) X(PALPHA_POROSITY) \
    X(STRESS_PALPHA_POROSITY) X(VARIABLE_SML) X(INTEGRATE_SML) \
    X(READ_INITIAL_SML_FROM_PARTICLE_FILE) X(AVERAGE_KERNELS) X(MAX_NUM_INTERACTIONS) \
    X(MAX_NUM_FLAWS) X(BOUNDARY_PARTICLE_ID)

#ifndef SOLID
#define SOLID 0
#endif
#ifndef HYDRO
#define HYDRO 0
#endif
#ifndef REAL_HYDRO
#define REAL_HYDRO 0
#endif
#ifndef INTEGRATE_ENERGY
#define INTEGRATE_ENERGY 0
#endif
#ifndef INTEGRATE_DENSITY
#define INTEGRATE_DENSITY 0
#endif
#ifndef FRAGMENTATION
#define FRAGMENTATION 0
#endif
#ifndef DAMAGE_ACTS_ON_S
#define DAMAGE_ACTS_ON_S 0
#endif
#ifndef SPH_EQU_VERSION
#define SPH_EQU_VERSION 1
#endif
#ifndef ARTIFICIAL_STRESS
#define ARTIFICIAL_STRESS 0
#endif
#ifndef ARTIFICIAL_VISCOSITY
#define ARTIFICIAL_VISCOSITY 0
#endif
#ifndef TENSORIAL_CORRECTION
#define TENSORIAL_CORRECTION 0
#endif
#ifndef VON_MISES_PLASTICITY
#define VON_MISES_PLASTICITY 0
#endif
#ifndef COLLINS_PLASTICITY
#define COLLINS_PLASTICITY 0
#endif
#ifndef COLLINS_PLASTICITY_INCLUDE_MELT_ENERGY
#define COLLINS_PLASTICITY_INCLUDE_MELT_ENERGY 0
#endif
#ifndef PALPHA_POROSITY
#define PALPHA_POROSITY 0
#endif
#ifndef STRESS_PALPHA_POROSITY
#define STRESS_PALPHA_POROSITY 0
#endif
#ifndef VARIABLE_SML
#define VARIABLE_SML 0
#endif
#ifndef INTEGRATE_SML
#define INTEGRATE_SML 0
#endif
#ifndef READ_INITIAL_SML_FROM_PARTICLE_FILE
#define READ_INITIAL_SML_FROM_PARTICLE_FILE 0
#endif
#ifndef AVERAGE_KERNELS
#define AVERAGE_KERNELS 0
#endif
#ifndef MAX_NUM_FLAWS
#define MAX_NUM_FLAWS 1
#endif
#ifndef BOUNDARY_PARTICLE_ID
#define BOUNDARY_PARTICLE_ID -1
#endif

/* derived, as in include/checks.h:27-29 */
#define B200_PLASTICITY (VON_MISES_PLASTICITY || COLLINS_PLASTICITY)

/* ---- scope guard: reference switches this library does not implement ---- */
#if (SOLID && HYDRO) || (!SOLID && !HYDRO)
#error Choose either SOLID or HYDRO in parameter.h (include/checks.h:33-35).
#endif
#if DIM < 1 || DIM > 3
#error DIM must be 1, 2 or 3.
#endif
#if SPH_EQU_VERSION != 1
#error Only SPH_EQU_VERSION 1 is in the hot-path scope (all scored configs use it).
#endif
#if defined(GRAVITATING_POINT_MASSES) && GRAVITATING_POINT_MASSES
#error GRAVITATING_POINT_MASSES is out of scope (0 in all scored configs).
#endif
#if defined(NAVIER_STOKES) && NAVIER_STOKES
#error NAVIER_STOKES is out of scope.
#endif
#if defined(XSPH) && XSPH
#error XSPH is out of scope.
#endif
#if defined(SHEPARD_CORRECTION) && SHEPARD_CORRECTION
#error SHEPARD_CORRECTION is out of scope.
#endif
#if defined(SML_CORRECTION) && SML_CORRECTION
#error SML_CORRECTION is out of scope.
#endif
#if defined(GHOST_BOUNDARIES) && GHOST_BOUNDARIES
#error GHOST_BOUNDARIES is out of scope.
#endif
#if defined(SIRONO_POROSITY) && SIRONO_POROSITY
#error SIRONO_POROSITY is out of scope.
#endif
#if defined(EPSALPHA_POROSITY) && EPSALPHA_POROSITY
#error EPSALPHA_POROSITY is out of scope.
#endif
#if defined(JC_PLASTICITY) && JC_PLASTICITY
#error JC_PLASTICITY is out of scope.
#endif
#if (defined(MOHR_COULOMB_PLASTICITY) && MOHR_COULOMB_PLASTICITY) || (defined(DRUCKER_PRAGER_PLASTICITY) && DRUCKER_PRAGER_PLASTICITY) || (defined(COLLINS_PLASTICITY_SIMPLE) && COLLINS_PLASTICITY_SIMPLE)
#error Only COLLINS_PLASTICITY and VON_MISES_PLASTICITY are in scope.
#endif
#if (defined(BALSARA_SWITCH) && BALSARA_SWITCH) || (defined(INVISCID_SPH) && INVISCID_SPH)
#error BALSARA_SWITCH / INVISCID_SPH are out of scope.
#endif
#if (defined(FIXED_NOI) && FIXED_NOI) || (defined(DEAL_WITH_TOO_MANY_INTERACTIONS) && DEAL_WITH_TOO_MANY_INTERACTIONS) || (defined(TOO_MANY_INTERACTIONS_KILL_PARTICLE) && TOO_MANY_INTERACTIONS_KILL_PARTICLE)
#error FIXED_NOI / DEAL_WITH_TOO_MANY_INTERACTIONS / TOO_MANY_INTERACTIONS_KILL_PARTICLE are out of scope.
#endif
#if (defined(PURE_REGOLITH) && PURE_REGOLITH) || (defined(VISCOUS_REGOLITH) && VISCOUS_REGOLITH)
#error regolith models are out of scope.
#endif
#if FRAGMENTATION && !SOLID
#error FRAGMENTATION needs SOLID.
#endif
#if VON_MISES_PLASTICITY && COLLINS_PLASTICITY
#error You cannot choose VON_MISES_PLASTICITY and COLLINS_PLASTICITY at the same time (include/checks.h:38-40).
#endif

#ifndef B200SPH_CONFIG_NAME
#define B200SPH_CONFIG_NAME "custom"
#endif

/* EOS ids, reference: include/pressure.h:31-48 (skipped when compiled next to the reference's own headers) */
#ifndef B200SPH_NO_EOS_ENUM
enum {
    EOS_TYPE_ACCRETED = -2, EOS_TYPE_IGNORE = -1, EOS_TYPE_POLYTROPIC_GAS = 0, EOS_TYPE_MURNAGHAN = 1,
    EOS_TYPE_TILLOTSON = 2, EOS_TYPE_ISOTHERMAL_GAS = 3, EOS_TYPE_REGOLITH = 4, EOS_TYPE_JUTZI = 5,
    EOS_TYPE_JUTZI_MURNAGHAN = 6, EOS_TYPE_ANEOS = 7, EOS_TYPE_VISCOUS_REGOLITH = 8, EOS_TYPE_IDEAL_GAS = 9,
    EOS_TYPE_SIRONO = 10, EOS_TYPE_EPSILON = 11, EOS_TYPE_LOCALLY_ISOTHERMAL_GAS = 12, EOS_TYPE_JUTZI_ANEOS = 13
};
#endif

#endif
