#!/bin/bash
# Round 2, call L: wall clock of the time integration at 10^6 particles, reference rk2Adaptive() vs b200sph_rk2_advance.
set -u
mkdir -p gpurun_out/integrator
timeout 500 python tools/integrator_speed.py impact:1000000 2>&1 | tail -n 4 | cut -c1-1500
