/* Compile-time switch set for the "rings" scenario, in the reference's own
 * parameter.h vocabulary (reference: test_cases/colliding_rings/parameter.h).  Only switches that are
 * non-zero or sized are listed; miluphcuda_b200/csrc/switches.h defaults every
 * other reference switch to 0 and rejects combinations outside the hot-path scope. */
#ifndef _PARAMETER_H
#define _PARAMETER_H
#define DIM 2
#define SOLID 1
#define INTEGRATE_DENSITY 1
#define SPH_EQU_VERSION 1
#define ARTIFICIAL_STRESS 1
#define ARTIFICIAL_VISCOSITY 1
#define TENSORIAL_CORRECTION 1
#define MAX_NUM_INTERACTIONS 256
#define MAX_NUM_FLAWS 1
#define BOUNDARY_PARTICLE_ID -1
#endif
