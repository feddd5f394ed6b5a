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

/* Cart-pole plant of the closed loop: CartPoleSimulator.step (src/cartpole_simulator.py:63-112), float64,
 * explicit Euler, termination |x| > 10 or |theta| > 0.5.  state [B,4] in/out, u [B] float32 (the controller's
 * dtype), done [B] out (may be NULL).  Test infrastructure, as the rest of oracle/.                      */
int phnn_oracle_plant_step(double *state, const float *u, double dt, long B, int *done) {
    const double gravity = 9.8, masscart = 1.0, masspole = 0.1, length = 0.5;
    const double polemass_length = masspole * length, total_mass = masspole + masscart;
    for (long b = 0; b < B; ++b) {
        double *s = state + 4 * b;
        const double force = (double)u[b];
        double x = s[0], th = s[1], xd = s[2], thd = s[3];
        const double c = cos(th), sn = sin(th);
        const double temp = (force + polemass_length * thd * thd * sn) / total_mass;
        const double thacc = (gravity * sn - c * temp) / (length * (4.0 / 3.0 - masspole * c * c / total_mass));
        const double xacc = temp - polemass_length * thacc * c / total_mass;
        x = x + dt * xd;
        th = th + dt * thd;
        xd = xd + dt * xacc;
        thd = thd + dt * thacc;
        s[0] = x; s[1] = th; s[2] = xd; s[3] = thd;
        if (done) done[b] = (fabs(x) > 10.0 || fabs(th) > 0.5) ? 1 : 0;
    }
    return 0;
}
