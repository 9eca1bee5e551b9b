/* Compile-time switch set for the "impact" scenario, in the reference's own
 * parameter.h vocabulary (reference: examples/impact/parameter.h).  Only switches that are
 * non-zero or sized are listed; miluphcuda_b200/csrc/switches.h defaults every
 * other reference switch to 0 and rejects combinations outside the hot-path scope. */
#ifndef _PARAMETER_H
#define _PARAMETER_H
#define DIM 3
#define SOLID 1
#define INTEGRATE_ENERGY 1
#define INTEGRATE_DENSITY 1
#define FRAGMENTATION 1
#define SPH_EQU_VERSION 1
#define ARTIFICIAL_VISCOSITY 1
#define TENSORIAL_CORRECTION 1
#define COLLINS_PLASTICITY 1
#define PALPHA_POROSITY 1
#define STRESS_PALPHA_POROSITY 1
#define VARIABLE_SML 1
#define INTEGRATE_SML 1
#define READ_INITIAL_SML_FROM_PARTICLE_FILE 1
#define MAX_NUM_INTERACTIONS 512
#define MAX_NUM_FLAWS 32
#define BOUNDARY_PARTICLE_ID -1
#endif
