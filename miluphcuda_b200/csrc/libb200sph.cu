/*
 * libb200sph.cu -- single translation unit of libb200sph_<config>.so.
 *
 * The material tables live in __constant__ memory that every kernel reads; without
 * relocatable device code a constant symbol is private to its translation unit, so the
 * kernels, the gravity walk and the C-ABI glue are compiled together.
 */
#include "rhs_kernels.cu"
#include "gravity.cu"
#include "halo.cu"
#include "integrate.cu"
#include "mg.cu"
#include "capi.cu"
