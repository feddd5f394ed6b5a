/*
 * include/phnn_mpc.h -- C ABI of libphnn_mpc.so, the B200 (sm_100a) implementation of the
 * pHNN-MPC hot path.  Plain pointers and sizes only; no torch types.
 *
 * The reference (Peilun-Tommy-Li/pHNN-MPC) is pure Python and has no FFI of its own: the
 * boundary it exposes is the Python module surface its drivers import
 * (scripts/run_cartpole_mpc.py:21-24, scripts/run_mpc_canonical.py:18-22).  Each entry point
 * below names the reference function whose arithmetic it replaces; the ctypes binding and the
 * drop-in Python modules that sit on top are in phnn_mpc_b200/ (see INTEGRATION.md).
 *
 * Conventions
 *   - every function returns int: 0 ok, <0 bad argument (PHNN_E_*), >0 a cudaError_t value;
 *     phnn_last_error() returns a thread-local message for the last non-zero return.
 *   - all data pointers are DEVICE pointers to float32, row-major, on the pack's device.
 *   - calls are asynchronous on `stream` (a cudaStream_t passed as void*); the library keeps
 *     no global mutable state; a pack changes only through phnn_pack_set_option and may be shared.
 *   - instances (rows of the leading B dimension) are independent.
 */
#ifndef PHNN_MPC_H
#define PHNN_MPC_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PHNN_KIND_PHNN 0      /* src/pHNN.py            dx = (J - J^T - R(x)) dH + G u          */
#define PHNN_KIND_CANONICAL 1 /* src/pHNN_canonical.py  z=[q,M(q)qd], dy = [M^-1 p, M^-1 pdot]  */

#define PHNN_EULER 0 /* src/integrators.py:13-36 */
#define PHNN_RK4 1   /* src/integrators.py:39-84 */

#define PHNN_E_ARG (-1)         /* null pointer / bad size                                      */
#define PHNN_E_UNSUPPORTED (-2) /* model shape has no kernel instantiation                      */
#define PHNN_E_INTEGRATOR (-3)  /* unknown integrator (reference raises ValueError)             */
#define PHNN_E_WORKSPACE (-4)   /* workspace too small                                          */

/* Host-side description of one model, in the reference's state_dict layout
 * (nn.Linear weight = [out,in] row-major).  All pointers are HOST float32.                     */
typedef struct phnn_model_desc {
    int kind;                                  /* PHNN_KIND_*                                   */
    int n, m;                                  /* state_dim, input_dim (m must be 1)            */
    int h;                                     /* hidden width: H_net [h,h]; R_net/G_net [h]     */
    int learned_G;                             /* 1: G_net present (src/pHNN.py:34-38)           */
    const float *W1, *b1, *W2, *b2, *W3, *b3;  /* H_net.net.{0,2,4}: [h,n] [h] [h,h] [h] [1,h] [1] */
    const float *Wr1, *br1, *Wr2, *br2;        /* R_net.net.{0,2}:   [h,n] [h] [n*n,h] [n*n]      */
    const float *Wg1, *bg1, *Wg2, *bg2;        /* G_net.net.{0,2}:   [h,n] [h] [n*m,h] [n*m]      */
    const float *J;                            /* [n,n] parameter (kind 0) / buffer (kind 1)     */
    const float *G;                            /* [n,m] G_fixed (kind 0) / G (kind 1)            */
    float mass_a, mass_b, mass_c;              /* kind 1: exp(log_a)+1e-3, b, exp(log_c)+1e-3    */
    const float *r_diag;                       /* kind 1: softplus(R_diag_raw)+1e-4, [n]         */
    int mass_const;                            /* kind 1: 1 = constant mass matrix M = [[a,b],[b,c]] = L L^T
                                                  (MassMatrixNetwork 'constant', src/mass_matrix.py:130-147,190-200:
                                                  no cos(theta), exact inverse); 0 = cart-pole M(theta)              */
} phnn_model_desc;

/* Horizon cost of the controllers (src/mpc_controller.py:75-114,
 * src/mpc_controller_canonical.py:91-120).  All pointers are HOST float32.                     */
typedef struct phnn_cost_desc {
    const float *Q;        /* [n,n]                                                             */
    const float *R;        /* [m,m]                                                             */
    const float *x_target; /* [n]                                                               */
    int has_u_bounds;      /* clamp(u, u_min, u_max) inside the differentiated graph            */
    float u_min, u_max;
    const float *x_min;    /* optional [n] soft bounds, NULL if absent                          */
    const float *x_max;
    float barrier_weight;  /* 1000 in the reference                                             */
} phnn_cost_desc;

typedef struct phnn_pack phnn_pack; /* opaque: packed weights resident in HBM */

const char *phnn_last_error(void);
int phnn_version(void);

/* Pack the weights of one model onto `device` (replaces nothing in the reference: it is the
 * device-side image of pHNN.__init__/load_state_dict, src/pHNN.py:13-38).                      */
int phnn_pack_create(const phnn_model_desc *desc, int device, phnn_pack **out);
int phnn_pack_destroy(phnn_pack *pack);
int phnn_pack_dims(const phnn_pack *pack, int *kind, int *n, int *m, int *h);

/* Kernel selection knobs (no reference counterpart).
 *   "tensor_mode"      0: FP32-FMA kernel only; 4 (default where built): second-generation tcgen05 kernel, three FP16
 *                      hi/lo products with operand A in tensor memory (FP32-level accuracy); 2: first-generation
 *                      kernel, TF32 product + one BF16 correction product (FP32-level accuracy); 3: 3xTF32 error
 *                      compensation (FP32-level accuracy, 1.5x the tensor work of 2); 1: plain TF32 (looser);
 *                      5: the second-generation kernel with ONE FP16 product per algorithmic product (operands
 *                      rounded to FP16, a third of the tensor work of 4; looser: 2e-4 on costs, 1e-3 on dJ/dU).
 *   "tensor_min_batch" smallest B routed to the tcgen05 kernel (default 1: it beats the FP32-FMA kernel
 *                      at every batch size; small batches go to the latency kernel first).
 *   "tensor_fwd_min_batch" n = 2 pHNN models (fixed or learned G, hidden 64: the pendulum model): smallest B of a
 *                      forward-only job (phnn_forward, phnn_rollout, phnn_cost_grad without dJdU) routed to the
 *                      forward-only instantiation of the second-generation tcgen05 kernel (default 10 x SM count by
 *                      measured crossover against the latency kernel; 0 disables it; needs "tensor_mode" 4).
 *   "tensor_fwd_sparse" the same kernel runs jobs of up to 64 instances per SM on 64-instance tiles (TMEM lanes 0..15 of
 *                      every quadrant; twice the SMs, 1.5 instead of 1.9 ms for 100 RK4 steps): 1 (default) / 0.
 *   "tensor_pair"      tensor_mode 4 solve jobs with an even number of 128-instance tiles run as clusters of two CTAs on
 *                      the two SMs of a TPC: every tensor product is one tcgen05.mma.cta_group::2 of M = 256 over the two
 *                      tiles of the pair, each CTA staging half of every weight tile (half the tensor-core operand-B reads
 *                      and half the L2 -> SM weight stream per SM).  Same results as the single-CTA launch (bit-identical
 *                      for models without an R_net; R_net sums added in a different order otherwise, <= 2e-6).
 *                      0 (default: measured +0.5 % on the power-capped full job, -1.7 % on an unthrottled slice) / 1.
 *   "latency_max_batch" largest B routed to the latency kernel (one thread per hidden unit, up to 8 instances
 *                      per CTA; default 4-96 x SM count by measured crossover, where built: hidden width <= 128;
 *                      0 disables it).                                                                                   */
int phnn_pack_set_option(phnn_pack *pack, const char *key, long value);
long phnn_pack_get_option(const phnn_pack *pack, const char *key);

/* dx[B,n], H[B] = model(x[B,n], u[B,m])          pHNN.forward src/pHNN.py:52-100,
 *                                                pHNN_Canonical.forward src/pHNN_canonical.py:172-273 */
int phnn_forward(const phnn_pack *pack, const float *x, const float *u, float *dx, float *H, long B,
                 void *stream);

/* xbar[B,n], ubar[B,m] = (d dx / d(x,u))^T v     what autograd evaluates for the reference's
 * double backward through torch.autograd.grad (src/pHNN.py:73).                                */
int phnn_vjp(const phnn_pack *pack, const float *x, const float *u, const float *v, float *xbar,
             float *ubar, long B, void *stream);

/* traj[B,T+1,n] (and energies[B,T+1] if non-NULL) from x0[B,n], U[B,T,m].
 * energy_mode 1: rollout_trajectory_differentiable(return_energies=True) ordering
 *                [H(y0),H(y0),H(y1),..,H(y_{T-1})]  src/integrators.py:192-258
 * energy_mode 2: rollout_trajectory ordering [H(y0),..,H(y_T)]  src/integrators.py:128-189     */
int phnn_rollout(const phnn_pack *pack, const float *x0, const float *U, float *traj, float *energies,
                 long B, int T, double dt, int integrator, int energy_mode, void *stream);

/* Bytes of scratch phnn_cost_grad / phnn_mpc_solve need for (B, T) under the pack's current options: stage-state
 * checkpoints, Adam moments and best controls per instance; for a batch routed to the tcgen05 kernel also its
 * scheduler words and the activation tape (one region of T*S*3*h*128*4 bytes per SM, S = 4 for RK4).  The
 * contents need not be preserved between calls.                                                 */
size_t phnn_workspace_bytes(const phnn_pack *pack, long B, int T, int integrator);

/* cost[B] and dJdU[B,T,m] (NULL: cost only) of the horizon cost at U[B,T,m]
 * (MPCController.rollout_dynamics+compute_cost+backward, src/mpc_controller.py:75-141,176-194;
 *  MPCControllerCanonical.rollout+compute_cost, src/mpc_controller_canonical.py:91-161).        */
int phnn_cost_grad(const phnn_pack *pack, const phnn_cost_desc *cost_desc, const float *x0, const float *U,
                   float *cost, float *dJdU, float *traj, long B, int T, double dt, int integrator,
                   void *workspace, size_t workspace_bytes, void *stream);

/* The whole solve: iters x { clamp, rollout, cost, adjoint, Adam step } in ONE launch.
 * U_inout[B,T,m]: initial guess in, result out.
 * return_mode 0: controls after the last Adam step, clamped (MPCController.compute_control,
 *                src/mpc_controller.py:143-209)
 * return_mode 1: clamped pre-step iterate of lowest cost (MPCControllerCanonical.optimize_control,
 *                src/mpc_controller_canonical.py:163-228)
 * cost_hist[iters,B] and best_cost[B] may be NULL.                                              */
int phnn_mpc_solve(const phnn_pack *pack, const phnn_cost_desc *cost_desc, const float *x0, float *U_inout,
                   float *cost_hist, float *best_cost, long B, int T, double dt, int integrator, double lr,
                   double beta1, double beta2, double eps, int iters, int return_mode, void *workspace,
                   size_t workspace_bytes, void *stream);

/* ---- training: gradients with respect to the WEIGHTS ---------------------------------
 * The reference trains by unrolling the model over a data sequence and back-propagating a loss of the predicted
 * trajectory through autograd's double backward (scripts/train_cartpole_phnn.py:108-178,
 * scripts/train_cartpole_phnn_canonical.py:83-196).  phnn_rollout_vjp is the vector-Jacobian product of
 * phnn_rollout: given gtraj = dL/dtraj [B,T+1,n] it returns dL/dx0 [B,n], dL/dU [B,T,m] and dL/dtheta for every
 * parameter, in the reference's state_dict layouts (nn.Linear weight = [out,in]).  One fused launch runs the rollout
 * and the discrete adjoint and emits, per evaluation, the per-hidden-unit factors of the parameter cotangents
 * (Wbar2 += delta2 (x) da1 + e2 (x) a1, ...); a second kernel contracts them over (instance, evaluation).
 * All pointers are DEVICE float32; NULL entries are skipped.  Gradients are OVERWRITTEN, not accumulated.
 * Built for the latency-kernel shapes (hidden width <= 128: every shipped training config).                     */
typedef struct phnn_param_grads {
    float *W1, *b1, *W2, *b2, *W3;       /* H_net.net.{0,2,4}: [h,n] [h] [h,h] [h] [1,h]  (b3 gets no gradient)   */
    float *Wr1, *br1, *Wr2, *br2;        /* R_net.net.{0,2}:   [h,n] [h] [n*n,h] [n*n]                            */
    float *Wg1, *bg1, *Wg2, *bg2;        /* G_net.net.{0,2}:   [h,n] [h] [n*m,h] [n*m]                            */
    float *J;                            /* [n,n] (kind 0; a buffer without gradient in kind 1)                    */
    float *r_diag;                       /* kind 1: dL/d r_diag [n] (chain through softplus is the caller's)       */
} phnn_param_grads;
size_t phnn_rollout_vjp_workspace_bytes(const phnn_pack *pack, long B, int T, int integrator);
int phnn_rollout_vjp(const phnn_pack *pack, const float *x0, const float *U, const float *gtraj, float *dx0, float *dU,
                     const phnn_param_grads *grads, long B, int T, double dt, int integrator, void *workspace,
                     size_t workspace_bytes, void *stream);

/* ---- multi-GPU: fused result exchange -----------------------------------------
 * Instances are sharded across ranks with no data-path collective; the only exchange is the final
 * gather of U* / best cost.  Instead of a separate NCCL all_gather the solve kernel itself stores every finished
 * 128-instance tile into the result buffers of ALL ranks (peer memory over NVLink), overlapped with the tiles still
 * being solved.  Result buffers are allocated by phnn_peer_alloc (cudaMalloc + CUDA IPC handle), the 64-byte handles
 * are exchanged by the host code (torch.distributed), and every rank opens the others' with phnn_peer_open.        */
#define PHNN_MAX_PEERS 8
typedef struct phnn_peer_desc {
    int n;                        /* ranks (1..PHNN_MAX_PEERS), own buffer included                        */
    long long offset;             /* global index of this rank's first instance                            */
    float *U[PHNN_MAX_PEERS];     /* DEVICE pointers (local or peer-mapped): [B_total, T, m] on every rank  */
    float *cost[PHNN_MAX_PEERS];  /* [B_total] best cost on every rank, or NULL                            */
} phnn_peer_desc;
int phnn_peer_alloc(size_t bytes, int device, void **dptr, void *handle64);
int phnn_peer_open(const void *handle64, int device, void **dptr);
int phnn_peer_close(void *dptr, int device);
int phnn_peer_free(void *dptr, int device);
/* phnn_mpc_solve + the fused exchange (peers == NULL: identical to phnn_mpc_solve).  Only batches routed to the
 * tcgen05 kernels support it (PHNN_E_UNSUPPORTED otherwise).  The caller separates consecutive solves that target
 * the same result buffers with a cross-rank barrier (phnn_mpc_b200/peer.py alternates two buffers).               */
int phnn_mpc_solve_peer(const phnn_pack *pack, const phnn_cost_desc *cost_desc, const float *x0, float *U_inout,
                        float *cost_hist, float *best_cost, long B, int T, double dt, int integrator, double lr,
                        double beta1, double beta2, double eps, int iters, int return_mode, void *workspace,
                        size_t workspace_bytes, const phnn_peer_desc *peers, void *stream);

/* ---- batched closed loop on the device (the callers either side of the solve) -------------------- */

/* Device buffers of B concurrent cart-pole episodes.                                              */
typedef struct phnn_episode {
    double *state;            /* [B,4] float64 plant state [x, theta, x_dot, theta_dot], in/out          */
    double *traj;             /* [B,steps+1,4] or NULL                                                  */
    float *controls;          /* [B,steps] or NULL                                                      */
    int *done_step;           /* [B] -1 while running, else the number of steps run when the plant ended */
    int *stable_start;        /* [B] -1 or the step at which the current in-tolerance streak began       */
    float *stable_duration;   /* [B] seconds                                                            */
    int *stability_achieved;  /* [B] 0/1                                                                */
    int steps;                /* episode length                                                         */
} phnn_episode;

/* One plant step for every running episode: stability bookkeeping of run_mpc_control
 * (scripts/run_cartpole_mpc.py:138-159) on the current state, then CartPoleSimulator.step
 * (src/cartpole_simulator.py:63-112) with force u[b*u_stride]; target/tol are HOST [4].             */
int phnn_plant_step(const phnn_episode *ep, const float *u, long u_stride, int step, double dt, const double *target,
                    const double *tol, double min_duration, long B, void *stream);
/* x0 = float32(state) as the controllers cast it (src/mpc_controller.py:160-161); if traj is non-NULL
 * also records state as traj[:,0,:].                                                               */
int phnn_state_to_f32(const double *state, float *x0, double *traj, int steps, long B, void *stream);
/* warm start: out[b,:] = [U[b,1:], 0] (src/mpc_controller_canonical.py:252-255); out != U.          */
int phnn_shift_controls(const float *U, float *out, long B, int H, void *stream);

/* Measurement utility (not part of the replaced path): launches a pure FFMA kernel of `blocks`
 * x 256 threads and returns the FLOPs it executes in *flops; bench.py times it with CUDA events
 * to get this GPU's sustained FP32-FMA rate, the roofline denominator of the FP32 path.        */
int phnn_ffma_probe(float *d_out, int iters, int blocks, void *stream, double *flops);
/* Same for the TF32 tcgen05.mma issue rate (n_mma 128x256x8 MMAs per CTA on resident operands): the
 * tensor-pipe denominator bench.py reports beside the bf16 peak of MEASURED_PEAKS.json.            */
int phnn_tf32_probe(float *d_out, int n_mma, int blocks, void *stream, double *flops);

#ifdef __cplusplus
}
#endif
#endif /* PHNN_MPC_H */
