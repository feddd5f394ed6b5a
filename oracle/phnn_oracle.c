/*
 * oracle/phnn_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see phnn_oracle_impl.h).
 * Builds the float32 (_f32) and float64 (_f64) instances of the CPU restatement.
 * Build: make -C oracle   ->  oracle/libphnn_oracle.so
 */
#include <math.h>
#include <stdlib.h>
#include <stddef.h>

#define REAL float
#define SUF _f32
#include "phnn_oracle_impl.h"
#undef REAL
#undef SUF

#define REAL double
#define SUF _f64
#include "phnn_oracle_impl.h"
#undef REAL
#undef SUF

#ifdef _OPENMP
#include <omp.h>
int phnn_oracle_max_threads(void) { return omp_get_max_threads(); }
void phnn_oracle_set_threads(int t) { if (t > 0) omp_set_num_threads(t); }
#else
int phnn_oracle_max_threads(void) { return 1; }
void phnn_oracle_set_threads(int t) { (void)t; }
#endif
