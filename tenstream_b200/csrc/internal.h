// internal.h -- symbols shared between solver.cu and f2c.cu (not part of the public ABI)
#pragma once
#include "../../include/pprts_b200.h"
int pprts_b200_internal_opp_probe(pprts_b200_solver *s, int mode, const float sample[4], int le, int ln, double *out);
int pprts_b200_internal_force_angles(pprts_b200_solver *s, double symmetry_phi, double theta);
int pprts_b200_internal_lut_ranges(const pprts_b200_lut *lut, float ranges[6][2], int *Nd, int *Nf);
