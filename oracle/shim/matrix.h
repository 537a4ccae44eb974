/* matrix.h -- part of the MEX shim (test infrastructure only); everything lives in mex.h */
#ifndef EGDST_SHIM_MATRIX_H
#define EGDST_SHIM_MATRIX_H
#include "mex.h"
#endif
