/* Compile-time switch set for the "giant_solid" scenario, in the reference's own
 * parameter.h vocabulary (reference: examples/giant_collisions/solid/parameter.h).  Only switches that are
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
#define MAX_NUM_INTERACTIONS 800
#define MAX_NUM_FLAWS 50
#define BOUNDARY_PARTICLE_ID -1
#endif
