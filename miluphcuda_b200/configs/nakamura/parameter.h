/* Compile-time switch set for the "nakamura" scenario (Nakamura & Fujiwara 1991 impact experiment: basalt sphere
 * with Grady-Kipp damage acting on S and von Mises plasticity), in the reference's own parameter.h vocabulary
 * (reference: test_cases/nakamura/parameter.h).  Only switches that are non-zero or sized are listed;
 * miluphcuda_b200/csrc/switches.h defaults every other reference switch to 0. */
#ifndef _PARAMETER_H
#define _PARAMETER_H
#define DIM 3
#define SOLID 1
#define INTEGRATE_ENERGY 1
#define INTEGRATE_DENSITY 1
#define FRAGMENTATION 1
#define DAMAGE_ACTS_ON_S 1
#define SPH_EQU_VERSION 1
#define ARTIFICIAL_VISCOSITY 1
#define TENSORIAL_CORRECTION 1
#define VON_MISES_PLASTICITY 1
#define MAX_NUM_INTERACTIONS 200
#define MAX_NUM_FLAWS 40
#define BOUNDARY_PARTICLE_ID -1
#endif
