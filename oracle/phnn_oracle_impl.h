/*
 * oracle/phnn_oracle_impl.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the pHNN-MPC hot path of Peilun-Tommy-Li/pHNN-MPC, written from the
 * reference's behaviour (not its text).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this.  The CUDA product path never calls it.
 *
 * This header is included twice by phnn_oracle.c, once with REAL=float (the reference's own
 * arithmetic type) and once with REAL=double (tie-breaker for tolerance questions).
 *
 * Parity pin: the reference ships no tests of its own ("parity unpinned" by the reference);
 * this restatement is pinned against outputs of the reference itself, generated in the build
 * container by tests/golden/make_golden.py and committed under tests/golden/.
 *
 * Reference anchors (all paths under /root/reference):
 *   MLP                       src/NN.py:6-40
 *   pHNN.forward              src/pHNN.py:52-100
 *   pHNN_Canonical.forward    src/pHNN_canonical.py:172-273
 *   CartPoleMassMatrix        src/mass_matrix.py:270-362
 *   coordinate transforms     src/coordinate_transforms.py:20-85,114-130
 *   euler_step / rk4_step     src/integrators.py:13-84
 *   rollouts                  src/integrators.py:128-258
 *   MPCController             src/mpc_controller.py:75-209
 *   MPCControllerCanonical    src/mpc_controller_canonical.py:91-228
 *   Adam                      torch/optim/adam.py _single_tensor_adam (torch 2.11, CPU default)
 *
 * The reference obtains dH/dx and the cost gradient with torch.autograd (double backward);
 * here both are written out analytically and verified against
 * autograd by the golden vectors.
 */

#ifndef REAL
#error "define REAL and SUF before including"
#endif

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)

#define MAXN 8
#define MAXM 4

typedef struct {
    int kind;            /* 0: pHNN (src/pHNN.py); 1: pHNN_Canonical (src/pHNN_canonical.py) */
    int n, m, h;         /* state dim, input dim, H_net hidden width (two hidden layers)      */
    int hr;              /* R_net hidden width (kind 0), single hidden layer                  */
    int hg;              /* G_net hidden width, 0 => fixed G                                  */
    const REAL *W1, *b1, *W2, *b2, *W3, *b3;   /* H_net  [h,n] [h] [h,h] [h] [1,h] [1]        */
    const REAL *Wr1, *br1, *Wr2, *br2;         /* R_net  [hr,n] [hr] [n*n,hr] [n*n]           */
    const REAL *Wg1, *bg1, *Wg2, *bg2;         /* G_net  [hg,n] [hg] [n*m,hg] [n*m]           */
    const REAL *J;       /* [n,n]: kind 0 raw parameter (J - J^T is used); kind 1 the buffer  */
    const REAL *G;       /* [n,m] fixed input matrix                                          */
    REAL ma, mb, mc;     /* kind 1: a = exp(log_a)+1e-3, b, c = exp(log_c)+1e-3               */
    const REAL *rdiag;   /* kind 1: softplus(R_diag_raw)+1e-4, [n]                            */
    int mconst;          /* kind 1: 1 = constant mass matrix [[a,b],[b,c]] (MassMatrixNetwork 'constant') */
} FN(oracle_model);

typedef struct {
    const REAL *Q;       /* [n,n] state weight (full matrix; controllers pass diagonals)      */
    const REAL *R;       /* [m,m] control weight                                              */
    const REAL *xt;      /* [n] target                                                        */
    int has_u_bounds;    /* clamp inside the differentiated graph                             */
    REAL u_min, u_max;
    const REAL *x_min;   /* optional [n] soft state bounds (NULL = absent)                    */
    const REAL *x_max;
    REAL barrier_w;      /* 1000 in src/mpc_controller.py:97                                  */
} FN(oracle_cost);

/* activations of one dynamics evaluation, kept for the analytic VJP */
typedef struct {
    REAL *a1, *a2, *g1, *r1, *ag;     /* [h] [h] [h] [hr] [hg]                                */
    REAL z[MAXN];                     /* H_net input (x, or canonical z)                      */
    REAL g[MAXN];                     /* dH/dz                                                */
    REAL S[MAXN * MAXN];              /* symmetrised R_net output                             */
    REAL G[MAXN * MAXM];              /* input matrix actually used                           */
    REAL beta, D, sth;                /* canonical: b cos(theta), determinant, sin(theta)     */
    REAL p[2], pdot[2];               /* canonical momenta and their rates                    */
} FN(eval_cache);

static int FN(cache_alloc)(const FN(oracle_model) * M, FN(eval_cache) * c) {
    size_t tot = (size_t)3 * M->h + M->hr + M->hg + 8;
    REAL *buf = (REAL *)calloc(tot, sizeof(REAL));
    c->a1 = buf;
    if (!buf) return -1;
    c->a2 = c->a1 + M->h;
    c->g1 = c->a2 + M->h;
    c->r1 = c->g1 + M->h;
    c->ag = c->r1 + M->hr;
    return 0;
}
static void FN(cache_free)(FN(eval_cache) * c) { free(c->a1); c->a1 = NULL; }

static inline REAL FN(tanh_)(REAL v) { return (REAL)tanh((double)v); }

/* H(z) and dH/dz for the two-hidden-layer tanh MLP (src/NN.py:16-25 with Tanh, bias).       */
static REAL FN(hnet)(const FN(oracle_model) * M, const REAL *z, FN(eval_cache) * c) {
    const int n = M->n, h = M->h;
    for (int k = 0; k < h; ++k) {
        REAL s = M->b1[k];
        for (int i = 0; i < n; ++i) s += M->W1[k * n + i] * z[i];
        c->a1[k] = FN(tanh_)(s);
    }
    REAL H = M->b3[0];
    for (int j = 0; j < h; ++j) {
        REAL s = M->b2[j];
        const REAL *w = M->W2 + (size_t)j * h;
        for (int k = 0; k < h; ++k) s += w[k] * c->a1[k];
        c->a2[j] = FN(tanh_)(s);
        H += M->W3[j] * c->a2[j];
    }
    /* reverse sweep: what autograd.grad(H.sum(), x) evaluates (src/pHNN.py:73)              */
    for (int k = 0; k < h; ++k) c->g1[k] = 0;
    for (int j = 0; j < h; ++j) {
        REAL d2 = (1 - c->a2[j] * c->a2[j]) * M->W3[j];
        const REAL *w = M->W2 + (size_t)j * h;
        for (int k = 0; k < h; ++k) c->g1[k] += w[k] * d2;
    }
    for (int i = 0; i < n; ++i) c->g[i] = 0;
    for (int k = 0; k < h; ++k) {
        REAL d1 = (1 - c->a1[k] * c->a1[k]) * c->g1[k];
        for (int i = 0; i < n; ++i) c->g[i] += M->W1[k * n + i] * d1;
    }
    for (int i = 0; i < n; ++i) c->z[i] = z[i];
    return H;
}

/* (d2H/dz2) w without a tape: forward-over-reverse through the same MLP.                     */
static void FN(hnet_hvp)(const FN(oracle_model) * M, const FN(eval_cache) * c, const REAL *w,
                         REAL *out, REAL *scratch /* 2h */) {
    const int n = M->n, h = M->h;
    REAL *da1 = scratch, *e2 = scratch + h;
    for (int k = 0; k < h; ++k) {
        REAL s = 0;
        for (int i = 0; i < n; ++i) s += M->W1[k * n + i] * w[i];
        da1[k] = (1 - c->a1[k] * c->a1[k]) * s;
    }
    for (int j = 0; j < h; ++j) {
        REAL s = 0;
        const REAL *wr = M->W2 + (size_t)j * h;
        for (int k = 0; k < h; ++k) s += wr[k] * da1[k];
        REAL da2 = (1 - c->a2[j] * c->a2[j]) * s;
        e2[j] = -2 * c->a2[j] * da2 * M->W3[j];        /* d(s2)/dt * w3                       */
    }
    for (int i = 0; i < n; ++i) out[i] = 0;
    for (int k = 0; k < h; ++k) {
        REAL dg1 = 0;
        for (int j = 0; j < h; ++j) dg1 += M->W2[(size_t)j * h + k] * e2[j];
        REAL s1 = 1 - c->a1[k] * c->a1[k];
        REAL ds1 = -2 * c->a1[k] * da1[k];
        REAL t = ds1 * c->g1[k] + s1 * dg1;
        for (int i = 0; i < n; ++i) out[i] += M->W1[k * n + i] * t;
    }
}

/* One dynamics evaluation dx = f(x,u), H.  Fills the cache for FN(f_vjp).                    */
static void FN(f_eval)(const FN(oracle_model) * M, const REAL *x, const REAL *u, REAL *dx, REAL *Hout,
                       FN(eval_cache) * c) {
    const int n = M->n, m = M->m;
    if (M->kind == 0) {
        REAL H = FN(hnet)(M, x, c);
        /* R_net (one hidden layer) -> n*n raw entries (src/pHNN.py:77)                       */
        REAL Rraw[MAXN * MAXN];
        for (int k = 0; k < M->hr; ++k) {
            REAL s = M->br1[k];
            for (int i = 0; i < n; ++i) s += M->Wr1[k * n + i] * x[i];
            c->r1[k] = FN(tanh_)(s);
        }
        for (int e = 0; e < n * n; ++e) {
            REAL s = M->br2[e];
            for (int k = 0; k < M->hr; ++k) s += M->Wr2[(size_t)e * M->hr + k] * c->r1[k];
            Rraw[e] = s;
        }
        /* S = (Rraw + Rraw^T)/2, R = S S^T (src/pHNN.py:79-80)                               */
        for (int a = 0; a < n; ++a)
            for (int b = 0; b < n; ++b) c->S[a * n + b] = (Rraw[a * n + b] + Rraw[b * n + a]) / 2;
        /* input matrix (src/pHNN.py:86-92)                                                   */
        if (M->hg > 0) {
            for (int k = 0; k < M->hg; ++k) {
                REAL s = M->bg1[k];
                for (int i = 0; i < n; ++i) s += M->Wg1[k * n + i] * x[i];
                c->ag[k] = FN(tanh_)(s);
            }
            for (int e = 0; e < n * m; ++e) {
                REAL s = M->bg2[e];
                for (int k = 0; k < M->hg; ++k) s += M->Wg2[(size_t)e * M->hg + k] * c->ag[k];
                c->G[e] = s;
            }
        } else {
            for (int e = 0; e < n * m; ++e) c->G[e] = M->G[e];
        }
        /* dx = (J - J^T - R) dH + G u (src/pHNN.py:83,97; note: no 1/2 on J - J^T)           */
        for (int a = 0; a < n; ++a) {
            REAL s = 0;
            for (int b = 0; b < n; ++b) {
                REAL Rab = 0;
                for (int k = 0; k < n; ++k) Rab += c->S[a * n + k] * c->S[b * n + k];
                REAL Aab = (M->J[a * n + b] - M->J[b * n + a]) - Rab;
                s += Aab * c->g[b];
            }
            REAL gu = 0;
            for (int j = 0; j < m; ++j) gu += c->G[a * m + j] * u[j];
            dx[a] = s + gu;
        }
        *Hout = H;
    } else {
        /* canonical, cart-pole mass matrix, q_dim = 2 (src/pHNN_canonical.py:172-273)        */
        const REAL a = M->ma, b = M->mb, cc = M->mc;
        const REAL th = x[1];
        const REAL beta = M->mconst ? b : b * (REAL)cos((double)th);   /* constant M: mass_matrix.py:130-147 */
        REAL z[4];
        z[0] = x[0];
        z[1] = x[1];
        z[2] = a * x[2] + beta * x[3];        /* p = M(q) qdot (coordinate_transforms.py:38-39) */
        z[3] = beta * x[2] + cc * x[3];
        REAL H = FN(hnet)(M, z, c);
        REAL dz[4];
        for (int r = 0; r < 4; ++r) {
            REAL s = 0;
            for (int k = 0; k < 4; ++k) {
                REAL A = M->J[r * 4 + k] - (r == k ? M->rdiag[r] : (REAL)0);
                s += A * c->g[k];
            }
            REAL gu = 0;
            for (int j = 0; j < m; ++j) gu += M->G[r * m + j] * u[j];
            dz[r] = s + gu;
        }
        /* cart-pole: mass_matrix.py:350-353; constant M: exact inverse, mass_matrix.py:190-200 */
        const REAL D = a * cc - beta * beta + (M->mconst ? (REAL)0 : (REAL)1e-6);
        const REAL n11 = cc / D, n12 = -beta / D, n22 = a / D;
        c->beta = beta; c->D = D; c->sth = M->mconst ? (REAL)0 : (REAL)sin((double)th);
        c->p[0] = z[2]; c->p[1] = z[3];
        c->pdot[0] = dz[2]; c->pdot[1] = dz[3];
        for (int e = 0; e < n * m; ++e) c->G[e] = M->G[e];
        dx[0] = n11 * z[2] + n12 * z[3];          /* qdot = M^-1 p                             */
        dx[1] = n12 * z[2] + n22 * z[3];
        dx[2] = n11 * dz[2] + n12 * dz[3];        /* qddot ~= M^-1 pdot (dM/dq neglected)      */
        dx[3] = n12 * dz[2] + n22 * dz[3];
        *Hout = H;
    }
}

/* Vector-Jacobian product of f at the cached point: xbar = (df/dx)^T v, ubar = (df/du)^T v.  */
static void FN(f_vjp)(const FN(oracle_model) * M, const FN(eval_cache) * c, const REAL *x, const REAL *u,
                      const REAL *v, REAL *xbar, REAL *ubar, REAL *scratch /* 2h + hr + hg */) {
    const int n = M->n, m = M->m;
    if (M->kind == 0) {
        /* s = S v, t = S g ; w = A^T v = (J-J^T)^T v - S s                                   */
        REAL s[MAXN], t[MAXN], w[MAXN];
        for (int a = 0; a < n; ++a) {
            REAL ss = 0, tt = 0;
            for (int b = 0; b < n; ++b) { ss += c->S[a * n + b] * v[b]; tt += c->S[a * n + b] * c->g[b]; }
            s[a] = ss; t[a] = tt;
        }
        for (int a = 0; a < n; ++a) {
            REAL acc = 0;
            for (int b = 0; b < n; ++b) acc += (M->J[b * n + a] - M->J[a * n + b]) * v[b];
            REAL Ss = 0;
            for (int b = 0; b < n; ++b) Ss += c->S[a * n + b] * s[b];
            w[a] = acc - Ss;
        }
        FN(hnet_hvp)(M, c, w, xbar, scratch);
        /* through R_net: Rbar_raw = -sym(v t^T + g s^T)                                       */
        REAL *rb = scratch;   /* reuse: [hr] */
        for (int k = 0; k < M->hr; ++k) rb[k] = 0;
        for (int a = 0; a < n; ++a)
            for (int b = 0; b < n; ++b) {
                REAL Rb = -(v[a] * t[b] + c->g[a] * s[b] + v[b] * t[a] + c->g[b] * s[a]) / 2;
                const REAL *wr = M->Wr2 + (size_t)(a * n + b) * M->hr;
                for (int k = 0; k < M->hr; ++k) rb[k] += wr[k] * Rb;
            }
        for (int k = 0; k < M->hr; ++k) {
            REAL zb = rb[k] * (1 - c->r1[k] * c->r1[k]);
            for (int i = 0; i < n; ++i) xbar[i] += M->Wr1[k * n + i] * zb;
        }
        if (M->hg > 0) {
            REAL *gb = scratch;
            for (int k = 0; k < M->hg; ++k) gb[k] = 0;
            for (int a = 0; a < n; ++a)
                for (int j = 0; j < m; ++j) {
                    REAL Gb = v[a] * u[j];
                    const REAL *wg = M->Wg2 + (size_t)(a * m + j) * M->hg;
                    for (int k = 0; k < M->hg; ++k) gb[k] += wg[k] * Gb;
                }
            for (int k = 0; k < M->hg; ++k) {
                REAL zb = gb[k] * (1 - c->ag[k] * c->ag[k]);
                for (int i = 0; i < n; ++i) xbar[i] += M->Wg1[k * n + i] * zb;
            }
        }
        for (int j = 0; j < m; ++j) {
            REAL acc = 0;
            for (int a = 0; a < n; ++a) acc += c->G[a * m + j] * v[a];
            ubar[j] = acc;
        }
    } else {
        const REAL a = M->ma, b = M->mb, cc = M->mc;
        const REAL beta = c->beta, D = c->D;
        const REAL n11 = cc / D, n12 = -beta / D, n22 = a / D;
        /* cotangents of p (through qdot = N p) and pdot (through qddot = N pdot)             */
        REAL pb[2], pdb[2];
        pb[0] = n11 * v[0] + n12 * v[1];
        pb[1] = n12 * v[0] + n22 * v[1];
        pdb[0] = n11 * v[2] + n12 * v[3];
        pdb[1] = n12 * v[2] + n22 * v[3];
        /* d/dtheta of N: beta' = -b sin, D' = -2 beta beta'                                  */
        const REAL dbeta = -b * c->sth;
        const REAL dD = -2 * beta * dbeta;
        const REAL dn11 = -cc / (D * D) * dD;
        const REAL dn12 = -dbeta / D + beta / (D * D) * dD;
        const REAL dn22 = -a / (D * D) * dD;
        REAL thbar = 0;
        thbar += v[0] * (dn11 * c->p[0] + dn12 * c->p[1]) + v[1] * (dn12 * c->p[0] + dn22 * c->p[1]);
        thbar += v[2] * (dn11 * c->pdot[0] + dn12 * c->pdot[1]) + v[3] * (dn12 * c->pdot[0] + dn22 * c->pdot[1]);
        /* pdot = rows 2,3 of (J - diag r) g + G u                                            */
        REAL dzb[4] = {0, 0, pdb[0], pdb[1]};
        REAL gbar[4];
        for (int k = 0; k < 4; ++k) {
            REAL acc = 0;
            for (int r = 0; r < 4; ++r) {
                REAL A = M->J[r * 4 + k] - (r == k ? M->rdiag[r] : (REAL)0);
                acc += A * dzb[r];
            }
            gbar[k] = acc;
        }
        for (int j = 0; j < m; ++j) {
            REAL acc = 0;
            for (int r = 0; r < 4; ++r) acc += M->G[r * m + j] * dzb[r];
            ubar[j] = acc;
        }
        REAL zb[4];
        FN(hnet_hvp)(M, c, gbar, zb, scratch);
        zb[2] += pb[0];
        zb[3] += pb[1];
        /* z = [q, M(theta) qdot]                                                             */
        thbar += dbeta * (zb[2] * x[3] + zb[3] * x[2]);
        xbar[0] = zb[0];
        xbar[1] = zb[1] + thbar;
        xbar[2] = a * zb[2] + beta * zb[3];
        xbar[3] = beta * zb[2] + cc * zb[3];
    }
}

/* ---- exported entry points ------------------------------------------------------------- */

int FN(phnn_oracle_forward)(const FN(oracle_model) * M, const REAL *x, const REAL *u, REAL *dx, REAL *H,
                            long B) {
    int err = 0;
#pragma omp parallel
    {
        FN(eval_cache) c;
        if (FN(cache_alloc)(M, &c)) {
#pragma omp atomic write
            err = -1;
        } else {
#pragma omp for schedule(static)
            for (long b = 0; b < B; ++b)
                FN(f_eval)(M, x + b * M->n, u + b * M->m, dx + b * M->n, H + b, &c);
            FN(cache_free)(&c);
        }
    }
    return err;
}

int FN(phnn_oracle_vjp)(const FN(oracle_model) * M, const REAL *x, const REAL *u, const REAL *v, REAL *xbar,
                        REAL *ubar, long B) {
    int err = 0;
#pragma omp parallel
    {
        FN(eval_cache) c;
        REAL *scr = (REAL *)malloc(sizeof(REAL) * (size_t)(2 * M->h + M->hr + M->hg + 8));
        if (FN(cache_alloc)(M, &c) || !scr) {
#pragma omp atomic write
            err = -1;
        } else {
#pragma omp for schedule(static)
            for (long b = 0; b < B; ++b) {
                REAL dx[MAXN], H;
                FN(f_eval)(M, x + b * M->n, u + b * M->m, dx, &H, &c);
                FN(f_vjp)(M, &c, x + b * M->n, u + b * M->m, v + b * M->n, xbar + b * M->n, ubar + b * M->m, scr);
            }
            FN(cache_free)(&c);
        }
        free(scr);
    }
    return err;
}

/* one integrator step; integrator 0 = euler (integrators.py:13-36), 1 = rk4 (:39-84).
 * If caches != NULL they receive the (up to 4) stage caches and ystage the stage states.     */
static void FN(step)(const FN(oracle_model) * M, const REAL *y, const REAL *u, double dt, int integrator,
                     REAL *ynext, REAL *H0, FN(eval_cache) * caches, REAL *ystage) {
    const int n = M->n;
    const REAL dtf = (REAL)dt, dt2 = (REAL)(dt / 2), dt6 = (REAL)(dt / 6.0);
    REAL k1[MAXN], k2[MAXN], k3[MAXN], k4[MAXN], yt[MAXN], Htmp;
    if (integrator == 0) {
        FN(f_eval)(M, y, u, k1, H0, &caches[0]);
        if (ystage) for (int i = 0; i < n; ++i) ystage[i] = y[i];
        for (int i = 0; i < n; ++i) ynext[i] = y[i] + dtf * k1[i];
        return;
    }
    FN(f_eval)(M, y, u, k1, H0, &caches[0]);
    if (ystage) for (int i = 0; i < n; ++i) ystage[i] = y[i];
    for (int i = 0; i < n; ++i) yt[i] = y[i] + dt2 * k1[i];
    FN(f_eval)(M, yt, u, k2, &Htmp, &caches[1]);
    if (ystage) for (int i = 0; i < n; ++i) ystage[n + i] = yt[i];
    for (int i = 0; i < n; ++i) yt[i] = y[i] + dt2 * k2[i];
    FN(f_eval)(M, yt, u, k3, &Htmp, &caches[2]);
    if (ystage) for (int i = 0; i < n; ++i) ystage[2 * n + i] = yt[i];
    for (int i = 0; i < n; ++i) yt[i] = y[i] + dtf * k3[i];
    FN(f_eval)(M, yt, u, k4, &Htmp, &caches[3]);
    if (ystage) for (int i = 0; i < n; ++i) ystage[3 * n + i] = yt[i];
    for (int i = 0; i < n; ++i) ynext[i] = y[i] + dt6 * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
}

/* energy_mode: 0 none; 1 = rollout_trajectory_differentiable(return_energies=True), i.e.
 * [H(y0), H(y0), H(y1), ..., H(y_{T-1})] (integrators.py:233-244); 2 = rollout_trajectory,
 * i.e. [H(y0), ..., H(y_T)] (:161-186).                                                      */
int FN(phnn_oracle_rollout)(const FN(oracle_model) * M, const REAL *x0, const REAL *U, REAL *traj,
                            REAL *energies, long B, int T, double dt, int integrator, int energy_mode) {
    if (integrator != 0 && integrator != 1) return -2;
    int err = 0;
    const int n = M->n, m = M->m;
#pragma omp parallel
    {
        FN(eval_cache) c[4];
        int bad = 0;
        for (int s = 0; s < 4; ++s) bad |= FN(cache_alloc)(M, &c[s]);
        if (bad) {
#pragma omp atomic write
            err = -1;
        } else {
#pragma omp for schedule(static)
            for (long b = 0; b < B; ++b) {
                REAL y[MAXN], yn[MAXN], H, dxs[MAXN];
                REAL *tr = traj + (size_t)b * (T + 1) * n;
                REAL *en = energies ? energies + (size_t)b * (T + 1) : NULL;
                for (int i = 0; i < n; ++i) { y[i] = x0[b * n + i]; tr[i] = y[i]; }
                if (en && T > 0) {
                    FN(f_eval)(M, y, U + (size_t)b * T * m, dxs, &H, &c[0]);
                    en[0] = H;
                }
                for (int t = 0; t < T; ++t) {
                    const REAL *u = U + ((size_t)b * T + t) * m;
                    FN(step)(M, y, u, dt, integrator, yn, &H, c, NULL);
                    for (int i = 0; i < n; ++i) { y[i] = yn[i]; tr[(t + 1) * n + i] = y[i]; }
                    if (en) {
                        if (energy_mode == 1) en[t + 1] = H;
                        else { FN(f_eval)(M, y, u, dxs, &H, &c[0]); en[t + 1] = H; }
                    }
                }
            }
        }
        for (int s = 0; s < 4; ++s) if (c[s].a1) FN(cache_free)(&c[s]);
    }
    return err;
}

/* stage cost l(x) = e^T Q e + w sum relu(xmin-x)^2 + w sum relu(x-xmax)^2 and its gradient   */
static REAL FN(state_cost)(const FN(oracle_cost) * C, int n, const REAL *x, REAL *grad) {
    REAL e[MAXN], Qe[MAXN], cost = 0;
    for (int i = 0; i < n; ++i) e[i] = x[i] - C->xt[i];
    for (int i = 0; i < n; ++i) {
        REAL s = 0, st = 0;
        for (int j = 0; j < n; ++j) { s += C->Q[i * n + j] * e[j]; st += C->Q[j * n + i] * e[j]; }
        Qe[i] = s;
        cost += e[i] * s;
        if (grad) grad[i] = s + st;
    }
    (void)Qe;
    if (C->x_min)
        for (int i = 0; i < n; ++i) {
            REAL viol = C->x_min[i] - x[i];
            if (viol > 0) { cost += C->barrier_w * viol * viol; if (grad) grad[i] -= 2 * C->barrier_w * viol; }
        }
    if (C->x_max)
        for (int i = 0; i < n; ++i) {
            REAL viol = x[i] - C->x_max[i];
            if (viol > 0) { cost += C->barrier_w * viol * viol; if (grad) grad[i] += 2 * C->barrier_w * viol; }
        }
    return cost;
}

static inline REAL FN(clampu)(const FN(oracle_cost) * C, REAL u) {
    if (!C->has_u_bounds) return u;
    return u < C->u_min ? C->u_min : (u > C->u_max ? C->u_max : u);
}

/* cost and dJ/dU for one instance.  work: xs [(H+1)*n], ystage [4n], caches[4], scratch.     */
static REAL FN(cost_grad_one)(const FN(oracle_model) * M, const FN(oracle_cost) * C, const REAL *x0,
                              const REAL *U, REAL *dJdU, REAL *xs, int H, double dt, int integrator,
                              FN(eval_cache) * c, REAL *scr, REAL *uc /* [H*m] clamped */) {
    const int n = M->n, m = M->m;
    const REAL dtf = (REAL)dt, dt2 = (REAL)(dt / 2), dt6 = (REAL)(dt / 6.0), dt3 = (REAL)(dt / 3.0);
    REAL Hdummy;
    for (int t = 0; t < H; ++t)
        for (int j = 0; j < m; ++j) uc[t * m + j] = FN(clampu)(C, U[t * m + j]);
    for (int i = 0; i < n; ++i) xs[i] = x0[i];
    REAL cost = 0;
    for (int t = 0; t < H; ++t) {
        cost += FN(state_cost)(C, n, xs + t * n, NULL);
        FN(step)(M, xs + t * n, uc + t * m, dt, integrator, xs + (t + 1) * n, &Hdummy, c, NULL);
    }
    cost += FN(state_cost)(C, n, xs + H * n, NULL);
    for (int t = 0; t < H; ++t)
        for (int i = 0; i < m; ++i) {
            REAL s = 0;
            for (int j = 0; j < m; ++j) s += C->R[i * m + j] * uc[t * m + j];
            cost += uc[t * m + i] * s;
        }
    if (!dJdU) return cost;
    /* reverse-time discrete adjoint                                   */
    REAL lam[MAXN], gl[MAXN];
    FN(state_cost)(C, n, xs + H * n, lam);
    for (int t = H - 1; t >= 0; --t) {
        const REAL *y = xs + t * n;
        const REAL *u = uc + t * m;
        REAL ub[MAXM], xb[MAXN], ubt[MAXM];
        for (int j = 0; j < m; ++j) ub[j] = 0;
        if (integrator == 0) {
            REAL k1[MAXN], v[MAXN];
            FN(f_eval)(M, y, u, k1, &Hdummy, &c[0]);
            for (int i = 0; i < n; ++i) v[i] = dtf * lam[i];
            FN(f_vjp)(M, &c[0], y, u, v, xb, ubt, scr);
            for (int i = 0; i < n; ++i) lam[i] += xb[i];
            for (int j = 0; j < m; ++j) ub[j] += ubt[j];
        } else {
            REAL ynext[MAXN], ys[4 * MAXN], kb[MAXN], ysum[MAXN];
            FN(step)(M, y, u, dt, 1, ynext, &Hdummy, c, ys);
            for (int i = 0; i < n; ++i) ysum[i] = 0;
            /* stage 4 */
            for (int i = 0; i < n; ++i) kb[i] = dt6 * lam[i];
            FN(f_vjp)(M, &c[3], ys + 3 * n, u, kb, xb, ubt, scr);
            for (int i = 0; i < n; ++i) ysum[i] += xb[i];
            for (int j = 0; j < m; ++j) ub[j] += ubt[j];
            /* stage 3 */
            for (int i = 0; i < n; ++i) kb[i] = dt3 * lam[i] + dtf * xb[i];
            FN(f_vjp)(M, &c[2], ys + 2 * n, u, kb, xb, ubt, scr);
            for (int i = 0; i < n; ++i) ysum[i] += xb[i];
            for (int j = 0; j < m; ++j) ub[j] += ubt[j];
            /* stage 2 */
            for (int i = 0; i < n; ++i) kb[i] = dt3 * lam[i] + dt2 * xb[i];
            FN(f_vjp)(M, &c[1], ys + 1 * n, u, kb, xb, ubt, scr);
            for (int i = 0; i < n; ++i) ysum[i] += xb[i];
            for (int j = 0; j < m; ++j) ub[j] += ubt[j];
            /* stage 1 */
            for (int i = 0; i < n; ++i) kb[i] = dt6 * lam[i] + dt2 * xb[i];
            FN(f_vjp)(M, &c[0], ys, u, kb, xb, ubt, scr);
            for (int i = 0; i < n; ++i) ysum[i] += xb[i];
            for (int j = 0; j < m; ++j) ub[j] += ubt[j];
            for (int i = 0; i < n; ++i) lam[i] += ysum[i];
        }
        FN(state_cost)(C, n, y, gl);
        for (int i = 0; i < n; ++i) lam[i] += gl[i];
        for (int i = 0; i < m; ++i) {
            REAL s = 0;
            for (int j = 0; j < m; ++j) s += (C->R[i * m + j] + C->R[j * m + i]) * u[j];
            REAL g = ub[i] + s;
            /* clamp sits inside the graph: zero gradient outside [u_min,u_max]
             * (src/mpc_controller.py:181, src/mpc_controller_canonical.py:196)                */
            if (C->has_u_bounds) {
                REAL raw = U[t * m + i];
                if (!(raw >= C->u_min && raw <= C->u_max)) g = 0;
            }
            dJdU[t * m + i] = g;
        }
    }
    return cost;
}

int FN(phnn_oracle_cost_grad)(const FN(oracle_model) * M, const FN(oracle_cost) * C, const REAL *x0,
                              const REAL *U, REAL *cost, REAL *dJdU, REAL *traj, long B, int H, double dt,
                              int integrator) {
    if (integrator != 0 && integrator != 1) return -2;
    int err = 0;
    const int n = M->n, m = M->m;
#pragma omp parallel
    {
        FN(eval_cache) c[4];
        int bad = 0;
        for (int s = 0; s < 4; ++s) bad |= FN(cache_alloc)(M, &c[s]);
        REAL *scr = (REAL *)malloc(sizeof(REAL) * (size_t)(2 * M->h + M->hr + M->hg + 8));
        REAL *xs = (REAL *)malloc(sizeof(REAL) * (size_t)(H + 1) * n);
        REAL *uc = (REAL *)malloc(sizeof(REAL) * (size_t)(H * m + 1));
        if (bad || !scr || !xs || !uc) {
#pragma omp atomic write
            err = -1;
        } else {
#pragma omp for schedule(static)
            for (long b = 0; b < B; ++b) {
                cost[b] = FN(cost_grad_one)(M, C, x0 + b * n, U + (size_t)b * H * m,
                                            dJdU ? dJdU + (size_t)b * H * m : NULL, xs, H, dt, integrator, c,
                                            scr, uc);
                if (traj)
                    for (int e = 0; e < (H + 1) * n; ++e) traj[(size_t)b * (H + 1) * n + e] = xs[e];
            }
        }
        for (int s = 0; s < 4; ++s) if (c[s].a1) FN(cache_free)(&c[s]);
        free(scr); free(xs); free(uc);
    }
    return err;
}

/* Full solve: iters x { clamp -> rollout -> cost -> adjoint -> Adam }.
 * return_mode 0: MPCController.compute_control (src/mpc_controller.py:143-209): U after the
 *   last Adam step, clamped.  1: MPCControllerCanonical.optimize_control
 *   (src/mpc_controller_canonical.py:163-228): the clamped pre-step iterate with the lowest
 *   cost (strict <), and that cost.
 * U_inout [B,H,m]: initial guess in, result out.  cost_hist [iters,B] optional.             */
int FN(phnn_oracle_mpc_solve)(const FN(oracle_model) * M, const FN(oracle_cost) * C, const REAL *x0,
                              REAL *U_inout, REAL *cost_hist, REAL *best_cost, long B, int H, double dt,
                              int integrator, double lr, double beta1, double beta2, double eps, int iters,
                              int return_mode) {
    if (integrator != 0 && integrator != 1) return -2;
    int err = 0;
    const int n = M->n, m = M->m;
    const int L = H * m;
#pragma omp parallel
    {
        FN(eval_cache) c[4];
        int bad = 0;
        for (int s = 0; s < 4; ++s) bad |= FN(cache_alloc)(M, &c[s]);
        REAL *scr = (REAL *)malloc(sizeof(REAL) * (size_t)(2 * M->h + M->hr + M->hg + 8));
        REAL *xs = (REAL *)malloc(sizeof(REAL) * (size_t)(H + 1) * n);
        REAL *buf = (REAL *)malloc(sizeof(REAL) * (size_t)(6 * L + 1));
        if (bad || !scr || !xs || !buf) {
#pragma omp atomic write
            err = -1;
        } else {
            REAL *uc = buf, *g = buf + L, *mo = buf + 2 * L, *vo = buf + 3 * L, *ub = buf + 4 * L, *u = buf + 5 * L;
#pragma omp for schedule(dynamic, 4)
            for (long b = 0; b < B; ++b) {
                for (int e = 0; e < L; ++e) { u[e] = U_inout[(size_t)b * L + e]; mo[e] = 0; vo[e] = 0; ub[e] = FN(clampu)(C, u[e]); }
                REAL best = (REAL)INFINITY;
                for (int it = 1; it <= iters; ++it) {
                    REAL cost = FN(cost_grad_one)(M, C, x0 + b * n, u, g, xs, H, dt, integrator, c, scr, uc);
                    if (cost_hist) cost_hist[(size_t)(it - 1) * B + b] = cost;
                    if (cost < best) { best = cost; for (int e = 0; e < L; ++e) ub[e] = uc[e]; }
                    /* torch.optim.Adam, single-tensor path, defaults (no amsgrad/weight decay) */
                    const double bc1 = 1.0 - pow(beta1, (double)it);
                    const double bc2 = 1.0 - pow(beta2, (double)it);
                    const REAL step_size = (REAL)(lr / bc1);
                    const REAL bc2s = (REAL)sqrt(bc2);
                    const REAL w1 = (REAL)(1.0 - beta1), b2f = (REAL)beta2, w2 = (REAL)(1.0 - beta2), epsf = (REAL)eps;
                    for (int e = 0; e < L; ++e) {
                        mo[e] = mo[e] + w1 * (g[e] - mo[e]);
                        vo[e] = vo[e] * b2f + w2 * g[e] * g[e];
                        REAL den = (REAL)sqrt((double)vo[e]) / bc2s + epsf;
                        u[e] = u[e] + (-step_size * mo[e]) / den;
                    }
                }
                if (return_mode == 0) {
                    for (int e = 0; e < L; ++e) U_inout[(size_t)b * L + e] = FN(clampu)(C, u[e]);
                    if (best_cost) best_cost[b] = best;
                } else {
                    for (int e = 0; e < L; ++e) U_inout[(size_t)b * L + e] = ub[e];
                    if (best_cost) best_cost[b] = best;
                }
            }
        }
        for (int s = 0; s < 4; ++s) if (c[s].a1) FN(cache_free)(&c[s]);
        free(scr); free(xs); free(buf);
    }
    return err;
}

#undef MAXN
#undef MAXM
#undef FN
#undef CAT
#undef CAT_
