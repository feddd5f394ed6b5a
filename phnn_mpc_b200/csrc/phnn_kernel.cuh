// phnn_kernel.cuh -- the fused pHNN-MPC kernel for sm_100a (FP32-FMA path).
//
// One launch runs a whole job for every instance it owns: a dynamics evaluation, a horizon
// rollout, a cost+gradient evaluation, or the complete MPC solve
//     iters x { clamp -> Euler/RK4 rollout -> cost -> reverse-time adjoint -> Adam step }
// replacing the Python loops of the reference (src/mpc_controller.py:143-209,
// src/mpc_controller_canonical.py:163-228, src/integrators.py:13-258) and its autograd tape
// (src/pHNN.py:73) with analytic dH/dx and Hessian-vector products.
//
// Decomposition
//   * a GROUP is 32 instances; lane l of every warp of the group owns instance l and carries
//     its state (x, adjoint, cost, ...) in registers for the whole job.  The per-instance 4x4
//     algebra is evaluated redundantly by all warps of the group, so no broadcast is needed.
//   * the h x h layers are a dense contraction [32 inst, h] x [h, h]; a warp computes a
//     32 x 64 tile with 8x8 register micro-tiles on the FP32 pipe; NWG = h/64 warps cover h.
//   * W2^T and W2 (the two operand orders the forward and reverse sweeps need) are streamed
//     from L2 by one producer warp with cp.async.bulk (TMA bulk copy, SASS UBLKCP) into a
//     ring of shared-memory stages guarded by full/empty mbarriers; every consumer warp of the
//     CTA reuses each stage.  The small layers (n->h, h->n*n, biases) are bulk-copied to
//     shared memory once per CTA.
//   * reductions over the hidden dimension use warp shuffles (transpose-reduce over the 8
//     lanes that share an instance block) and one shared-memory exchange across the NWG warps.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace phnn {

constexpr int GI = 32;   // instances per group (one lane each)
constexpr int NST = 4;   // ring stages
constexpr int MAX_CONSUMER_WARPS = 8;
constexpr int MAX_THREADS = (MAX_CONSUMER_WARPS + 1) * 32;

enum { MODE_FORWARD = 0, MODE_VJP = 1, MODE_ROLLOUT = 2, MODE_COSTGRAD = 3, MODE_SOLVE = 4, MODE_PARAMGRAD = 5 };
// MODE_PARAMGRAD (training): rollout, then the discrete adjoint driven by a caller-supplied cotangent
// of the trajectory (gtraj [B,T+1,n]) instead of the MPC cost; every adjoint evaluation EMITS the per-hidden-unit
// factors of the parameter cotangents (phnn_lat_kernel.cuh), which phnn_capi.cu contracts over (instance, evaluation)
// rows into dL/dtheta.  Outputs dL/dx0 in out0 [B,n] and dL/dU in dJdU [B,T].
constexpr int EMIT_NARR = 11;    // per-row arrays of h floats: a1, da1, delta2, e2, delta1, zbar1, s2*dz2, r1, dr, ag, dg
constexpr int EMIT_SMALL = 48;   // per-row small record: z@0 w@4 v@8 g@12 y@16 Rbar_raw@20 (16) pdb*g@36 (2) v*u@40 (4)
enum { MK_PHNN = 0, MK_PHNN_GNET = 1, MK_CANON = 2 };

template <int MK_, int NS_, int HID_>
struct Shape {
    static constexpr int MK = MK_, NS = NS_, HID = HID_;
    static constexpr int NWG = HID / 64;  // warps per group
    static constexpr int NN = NS * NS;
    static constexpr int KT = (HID >= 256) ? 8 : 16;  // k rows per ring stage
    static constexpr bool HAS_R = (MK != MK_CANON);
    static constexpr bool HAS_GNET = (MK == MK_PHNN_GNET);
    // layout of the small-weight blob (floats); identical in HBM and in shared memory
    static constexpr int O_W1 = 0;                                    // [HID][NS]
    static constexpr int O_B1 = O_W1 + HID * NS;                      // [HID]
    static constexpr int O_B2 = O_B1 + HID;                           // [HID]
    static constexpr int O_W3 = O_B2 + HID;                           // [HID]
    static constexpr int O_WR1 = O_W3 + HID;                          // [HID][NS]
    static constexpr int O_BR1 = O_WR1 + (HAS_R ? HID * NS : 0);      // [HID]
    static constexpr int O_WR2 = O_BR1 + (HAS_R ? HID : 0);           // [NN][HID]
    static constexpr int O_WG1 = O_WR2 + (HAS_R ? NN * HID : 0);      // [HID][NS]
    static constexpr int O_BG1 = O_WG1 + (HAS_GNET ? HID * NS : 0);   // [HID]
    static constexpr int O_WG2 = O_BG1 + (HAS_GNET ? HID : 0);        // [NS][HID]
    static constexpr int SMALL = (O_WG2 + (HAS_GNET ? NS * HID : 0) + 31) / 32 * 32;
    // per-lane partial-sum slots exchanged between the warps of a group
    static constexpr int SL_R = 0;
    static constexpr int SL_G = SL_R + (HAS_R ? NN : 0);
    static constexpr int SL_H = SL_G + (HAS_GNET ? NS : 0);
    static constexpr int SL_DH = SL_H + 1;
    static constexpr int SL_X = SL_DH + NS;
    static constexpr int PW = (SL_X + NS) | 1;  // odd stride: conflict-free scalar access
    static constexpr int NAUX = (HAS_R ? NN : 0) + (HAS_GNET ? NS : 0);
    // per-group shared memory (floats)
    static constexpr int G_BUFA = 0;
    static constexpr int G_BUFB = HID * GI;
    static constexpr int G_PART = 2 * HID * GI;
    static constexpr int G_RBAR = G_PART + NWG * 32 * PW;
    static constexpr int G_FLOATS = (G_RBAR + (NAUX > 0 ? NAUX : 1) * GI + 31) / 32 * 32;
    static constexpr int RING = NST * KT * HID;
    static constexpr int STAGE_BYTES = KT * HID * 4;
    static constexpr int MAX_NG = MAX_CONSUMER_WARPS / NWG;
    __host__ __device__ static constexpr size_t smem_bytes(int ng) {
        return 128 + sizeof(float) * (size_t)(SMALL + RING + ng * G_FLOATS);
    }
};

struct KParams {
    // packed weights (HBM)
    const float* wsmall;
    const float* wbig;  // [W2^T | W2], each [HID][HID]
    const unsigned char* wtc;  // tcgen05 path: pre-swizzled hi/lo K-blocks of W2 and W2^T
    const float* wsmall_tc;    // tcgen05 path: per-hidden-unit records of the small layers
    int tc_split;              // 3 = 3xTF32 (FP32-level accuracy), 1 = plain TF32
    // second-generation tcgen05 kernel (phnn_tc16_kernel.cuh, tensor_mode 4): FP16 hi/lo operands, A in tensor memory
    const unsigned char* wtc16;  // pre-swizzled K-blocks of W2 and W2^T, rows of [b_hi (32 fp16) | b_lo (32 fp16)], scaled by S_B
    const float* wsmall16;       // field-major small-layer records (Tc16Shape)
    float s16[8];                // exact power-of-two scales: [0] 1/(S_a S_B), [1] 1/(S_delta S_B), [2] 1/S_B, [3] 1/(S_e S_B),
                                 // [4] S_a, [5] 1/S_delta, [6] S_e S_B, [7] 1/(2^9 S_RW) (R_net output sums on mma.sync)
    int wexp16;                  // per-instance adjoint scale: max |w'| lands in [2^wexp16, 2^(wexp16+1))
    int rexp16;                  // S_RW = 2^rexp16, the scale of the R_net output-layer fragments
    // model constants
    float Jm[16];  // MK 0/1: J - J^T ; MK 2: the canonical J buffer
    float Gv[4];
    float b3, ma, mb, mc;
    int mconst;  // canonical: 1 = constant mass matrix [[ma, mb], [mb, mc]] (no cos(theta), exact inverse)
    float br2[16];
    float bsym[12];  // tcgen05 path: (br2[ab] + br2[ba])/2 packed 00 01 02 03 11 12 13 22 23 33
    float bg2[4];
    float rdiag[4];
    // cost
    float Q[16], Qs[16];  // Q and Q + Q^T
    float Rw;             // R[0][0] (m = 1)
    float xt[4], xmin[4], xmax[4];
    int has_ub, has_xmin, has_xmax;
    float umin, umax, bw;
    // job
    int mode, T, S, energy_mode, iters, return_mode, want_grad, ng;
    long long B;
    float dt, dt2, dt3, dt6;
    double lr, beta1, beta2, eps;
    // io (device)
    const float* x0;   // [B,n]
    const float* uin;  // FORWARD/VJP: u [B]; ROLLOUT/COSTGRAD: U [B,T]
    const float* vin;  // VJP: v [B,n]
    float* U;          // SOLVE: U in/out [B,T]
    float* out0;       // FORWARD dx | VJP xbar | ROLLOUT/COSTGRAD traj (nullable)
    float* out1;       // FORWARD H | VJP ubar | ROLLOUT energies (nullable)
    float* cost;       // COSTGRAD cost [B] | SOLVE best_cost (nullable)
    float* dJdU;       // COSTGRAD [B,T] (nullable)
    float* cost_hist;  // SOLVE [iters,B] (nullable)
    float* ws;         // workspace
    int* sched;        // work-stealing solve (tcgen05 kernel): [0] unit counter, [1 + tile] iterations completed
    long long tiles;   // number of 128-instance tiles (tcgen05 kernel)
    float* tape;       // tcgen05 kernel: activation tape, [grid][T*S][3][h][128] floats (jobs with an adjoint)
    float* scratch;    // tcgen05 kernels: per-CTA stage states [T*S][n][128] + R_net sums / grad H [T*S][16][128] of the
                       // (tile, iteration) unit in flight; nullptr: the stage states live in the tile's workspace
    long long* dbg;    // optional profiling output (PHNN_TC_PROFILE builds)
    // fused result exchange (multi-GPU solve): after the last iteration of a tile the CTA stores the tile's controls
    // and best costs straight into the result buffers of every rank (peer memory over NVLink)
    int npeer;             // 0: no exchange
    long long peer_off;    // global index of this rank's first instance
    float* peerU[8];       // [B_total, T] on every rank (own buffer included)
    float* peerC[8];       // [B_total] best cost, nullable
    // MODE_PARAMGRAD
    const float* gtraj;    // [B,T+1,n] cotangent of the trajectory
    float* emit;           // [EMIT_NARR][rows][h] then [rows][EMIT_SMALL], rows = B*T*S
    long long emit_rows;
};

// floats of workspace per tile of TW instances: stage states [T*S][NS][TW], Adam m, v and best
// controls [T][TW] (lane-interleaved so the accesses of a warp coalesce)
// (+ `extra` floats per instance: the tcgen05 kernel keeps the R_net output sums and grad H of every forward
// evaluation there, see tc_ws_extra; its activation tape is a separate per-CTA region, see phnn_capi.cu)
__host__ __device__ inline size_t ws_floats_per_tile(int NS, int T, int S, int TW, int extra = 0) {
    return (size_t)TW * ((size_t)T * S * NS + 3 * (size_t)T + 1 + (size_t)extra);
}

// extra workspace floats per instance of the tcgen05 kernel: 16 per evaluation for the symmetrised R_net sums
// (10) and grad H (4) the forward sweep leaves for the adjoint
__host__ __device__ inline int tc_ws_extra(int /*h*/, int T, int S) { return 16 * T * S; }
// What has to outlive a (tile, iteration) unit: Adam m, v, best controls [T][TW] and the best cost [TW].  The
// tcgen05 kernels keep only this per tile; stage states, R_net sums and grad H are consumed by the reverse sweep of the
// same unit and live in a per-CTA scratch region (KParams::scratch), so 1 M instances x H = 200 fit one GPU.
__host__ __device__ inline size_t ws_persist_floats_per_tile(int T, int TW) { return (size_t)TW * (3 * (size_t)T + 1); }
__host__ __device__ inline size_t tc_scratch_floats_per_cta(int NS, int T, int S) { return (size_t)T * S * (NS + 16) * 128; }

// One unit of work of a job: iteration `it` (1-based) of tile `tile`.  The static schedule hands a
// CTA the iterations of its own tile in order; the work-stealing schedule of the tcgen05 kernel
// hands out (tile, iteration) pairs from a global counter so that batches whose tile count is not
// a multiple of the SM count do not idle SMs in the last wave.
struct Unit {
    long long tile;
    int it;
};
struct StaticSched {
    long long tile;
    int n_outer, it;
    __device__ __forceinline__ bool next(Unit& u) {
        if (it >= n_outer) return false;
        u.tile = tile;
        u.it = ++it;
        return true;
    }
    template <class ENG>
    __device__ __forceinline__ void done(ENG& c, const Unit&) {
        // the storing thread's update of U[.,0] must be visible to the instance's other owners
        // before they re-read it at the top of the next forward sweep
        c.gbar();
    }
    static constexpr bool kStateInWorkspace = false;
    __device__ __forceinline__ long long tile0() const { return tile; }
};

// ---------------------------------------------------------------------------------------
// PTX helpers: mbarrier, bulk copy, named barriers, MUFU
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void group_bar(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// tanh: the accurate libdevice tanhf (branch-free on sm_100a: a degree-7 odd polynomial below
// |x| = 0.6, 1 - 2/(exp(2x)+1) with MUFU.EX2/MUFU.RCP above; ~14 instructions, ~1 ulp).
// tanh.approx.f32 (2^-11) would break the 1e-5 per-step tolerance, and the
// 5-instruction 1 - 2/(exp(2x)+1) form alone (PHNN_TANH_FAST, ~2e-7 absolute) measured 2x the
// horizon error of tanhf on the 100-step pendulum rollout for ~4% less time, so it is opt-in.
__device__ __forceinline__ float tanh_acc(float x) {
#ifdef PHNN_TANH_FAST
    const float e = ex2_approx(x * 2.8853900817779268f);
    const float r = rcp_approx(e + 1.0f);
    return fmaf(-2.0f, r, 1.0f);
#else
    return tanhf(x);
#endif
}

// Sum V values per lane over the 8 lanes that share (lane>>3): afterwards v[0..V/8) of lane
// `lo` hold the totals of chunk `lo` (original indices [lo*V/8, (lo+1)*V/8)).
template <int V, int BIT>
__device__ __forceinline__ void lane8_step(float* v, int lo) {
    constexpr int HALF = V / 2;
    const bool up = (lo & BIT) != 0;
#pragma unroll
    for (int i = 0; i < HALF; ++i) {
        const float send = up ? v[i] : v[i + HALF];
        const float keep = up ? v[i + HALF] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, BIT);
    }
}
template <int V>
__device__ __forceinline__ void lane8_reduce(float* v, int lo) {
    lane8_step<V, 4>(v, lo);
    lane8_step<V / 2, 2>(v, lo);
    lane8_step<V / 4, 1>(v, lo);
}

// ---------------------------------------------------------------------------------------
// per-thread context
// ---------------------------------------------------------------------------------------
// Dynamic shared memory of the CTA.  Everything is addressed as (this symbol + offset) so the
// compiler emits shared-space LDS/STS with 32-bit addresses instead of generic LD/ST (the first
// ncu capture showed LD.E.128 in the product loop and ~20% long-scoreboard stalls on them).
extern __shared__ __align__(1024) unsigned char phnn_smem[];

template <class SH> struct Ctx;
template <class SH>
__device__ void eval_fwd(Ctx<SH>& c, const KParams& p, const float (&y)[SH::NS], float u, float (&f)[SH::NS], float& Hval);
template <class SH>
__device__ void eval_vjp(Ctx<SH>& c, const KParams& p, const float (&y)[SH::NS], float u, const float (&v)[SH::NS],
                         float (&xbar)[SH::NS], float& ubar);

template <class SH>
struct Ctx {
    static constexpr int NS = SH::NS;
    static constexpr int TW = GI;  // instances per workspace tile
    __device__ static int ws_extra(const KParams&) { return 0; }
    __device__ __forceinline__ void set_eval(int) {}
    uint32_t goff;  // float offset of this group's region
    int lane, li, lo, wg, barid, wcol, chunk;
    uint32_t tile;  // ring tiles consumed so far
    bool ready;     // the next ring stage was already seen full by an early poll
    bool store;     // this warp performs the group's global stores

    __device__ __forceinline__ float* smf() const { return reinterpret_cast<float*>(phnn_smem + 128); }
    __device__ __forceinline__ const float* wsm() const { return smf(); }
    __device__ __forceinline__ const float* ring() const { return smf() + SH::SMALL; }
    __device__ __forceinline__ float* bufA() const { return smf() + goff + SH::G_BUFA; }
    __device__ __forceinline__ float* bufB() const { return smf() + goff + SH::G_BUFB; }
    __device__ __forceinline__ float* part() const { return smf() + goff + SH::G_PART; }
    __device__ __forceinline__ float* rbar() const { return smf() + goff + SH::G_RBAR; }
    __device__ __forceinline__ uint64_t* full() const { return reinterpret_cast<uint64_t*>(phnn_smem); }
    __device__ __forceinline__ uint64_t* empty() const { return full() + NST; }

    __device__ __forceinline__ void gbar() const {
        if (SH::NWG == 1) __syncwarp();
        else group_bar(barid, SH::NWG * 32);
    }
    // hidden unit handled in micro-tile column o
    __device__ __forceinline__ int kown(int o) const { return wcol + ((o >> 2) << 5) + (o & 3); }
    __device__ __forceinline__ float* row_own(float* buf, int k) const { return buf + k * GI + (chunk << 3); }
    __device__ __forceinline__ void begin_unit(const KParams&, long long) {}
    __device__ __forceinline__ float* unit_scratch() const { return nullptr; }
    __device__ __forceinline__ void peer_store(const KParams&, long long) {}
    __device__ __forceinline__ void eval_fwd(const KParams& p, const float (&y)[NS], float u, float (&f)[NS], float& H) {
        phnn::eval_fwd(*this, p, y, u, f, H);
    }
    __device__ __forceinline__ void eval_vjp(const KParams& p, const float (&y)[NS], float u, const float (&v)[NS],
                                             float (&xbar)[NS], float& ubar) {
        phnn::eval_vjp(*this, p, y, u, v, xbar, ubar);
    }
};

// acc[r][o] += sum_k lhs[k][8*li + r] * W[k][kown(o)] over one full sweep of HID rows taken
// from the ring (HID/KT stages).
template <class SH>
__device__ __forceinline__ void product(Ctx<SH>& c, const float* __restrict__ lhs, float (&acc)[8][8]) {
    constexpr int HID = SH::HID, KT = SH::KT;
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int o = 0; o < 8; ++o) acc[r][o] = 0.f;
#pragma unroll 1
    for (int t = 0; t < HID / KT; ++t) {
        const uint32_t s = c.tile % NST;
        const uint32_t ph = (c.tile / NST) & 1u;
        if (!c.ready) mbar_wait(&c.full()[s], ph);
        // poll the next stage now; the answer is consumed after this stage's FFMAs, which hides
        // the ~90-cycle try_wait latency the first capture showed at the top of every stage
        {
            const uint32_t s2 = (c.tile + 1) % NST, ph2 = ((c.tile + 1) / NST) & 1u;
            c.ready = mbar_try(&c.full()[s2], ph2);
        }
        const float* __restrict__ wst = c.ring() + s * (KT * HID) + c.wcol;
        const int k0 = t * KT;
#pragma unroll
        for (int kk = 0; kk < KT; ++kk) {
            const int k = k0 + kk;
            const float4* ap = reinterpret_cast<const float4*>(lhs + k * GI + ((c.li ^ ((k >> 2) & 3)) << 3));
            const float4 a0 = ap[0], a1 = ap[1];
            const float4* wp = reinterpret_cast<const float4*>(wst + kk * HID);
            const float4 w0 = wp[0], w1 = wp[8];
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int o = 0; o < 8; ++o) acc[r][o] = fmaf(a[r], w[o], acc[r][o]);
        }
        __syncwarp();
        if (c.lane == 0) mbar_arrive(&c.empty()[s]);
        ++c.tile;
    }
}

template <int NS>
__device__ __forceinline__ void gather8(const float (&v)[NS], int li, float (&out)[8][NS]) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int i = 0; i < NS; ++i) out[r][i] = __shfl_sync(0xffffffffu, v[i], (li << 3) + r);
}

template <int NS>
__device__ __forceinline__ float dotn(const float* w, const float (&x)[NS], float b) {
    float s = b;
#pragma unroll
    for (int i = 0; i < NS; ++i) s = fmaf(w[i], x[i], s);
    return s;
}

__device__ __forceinline__ void st8(float* p, const float (&v)[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
    const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
// 8 weights W[row][kown(0..7)] of a [rows][HID] matrix in smem
template <class SH>
__device__ __forceinline__ void ldw8(const Ctx<SH>& c, const float* mat_row, float (&w)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(mat_row + c.wcol);
    const float4 b = *reinterpret_cast<const float4*>(mat_row + c.wcol + 32);
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
}

// canonical coordinates of the cart-pole mass matrix (src/mass_matrix.py:270-362)
struct Canon {
    float beta, D, n11, n12, n22, sth;
};
__device__ __forceinline__ Canon canon_of(const KParams& p, float theta) {
    Canon q;
    float s, cth;
    if (p.mconst) {
        // constant mass matrix (MassMatrixNetwork 'constant', src/mass_matrix.py:130-147): M = [[a, b], [b, c]], exact inverse
        s = 0.f;
        cth = 1.f;
    } else {
        sincosf(theta, &s, &cth);
    }
    q.sth = s;
    q.beta = p.mb * cth;
    q.D = p.ma * p.mc - q.beta * q.beta + (p.mconst ? 0.f : 1e-6f);
    q.n11 = p.mc / q.D;
    q.n12 = -q.beta / q.D;
    q.n22 = p.ma / q.D;
    return q;
}

// S = (Rraw + Rraw^T)/2 from the exchanged partial sums (src/pHNN.py:79)
template <class SH>
__device__ __forceinline__ void load_S(const Ctx<SH>& c, const KParams& p, float (&S)[SH::NS][SH::NS]) {
    constexpr int NS = SH::NS, NN = SH::NN;
    float Rraw[NN];
#pragma unroll
    for (int e = 0; e < NN; ++e) {
        float s = p.br2[e];
#pragma unroll
        for (int w = 0; w < SH::NWG; ++w) s += c.part()[(w * 32 + c.lane) * SH::PW + SH::SL_R + e];
        Rraw[e] = s;
    }
#pragma unroll
    for (int a = 0; a < NS; ++a)
#pragma unroll
        for (int b = 0; b < NS; ++b) S[a][b] = (Rraw[a * NS + b] + Rraw[b * NS + a]) * 0.5f;
}

template <class SH, int SLOT, int N>
__device__ __forceinline__ void load_part(const Ctx<SH>& c, float (&out)[N]) {
#pragma unroll
    for (int e = 0; e < N; ++e) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < SH::NWG; ++w) s += c.part()[(w * 32 + c.lane) * SH::PW + SLOT + e];
        out[e] = s;
    }
}

// layer 1 of H_net for the thread's 8 hidden units x 8 instances -> bufA ; aux nets' hidden
// layers and their output partial sums -> part
template <class SH>
__device__ __forceinline__ void layer1_and_aux(Ctx<SH>& c, const float (&zi)[8][SH::NS], const float (&yi)[8][SH::NS]) {
    constexpr int NS = SH::NS, NN = SH::NN, HID = SH::HID;
    float r1[SH::HAS_R ? 8 : 1][8];
    float ag[SH::HAS_GNET ? 8 : 1][8];
#pragma unroll
    for (int o = 0; o < 8; ++o) {
        const int k = c.kown(o);
        float w1[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) w1[i] = c.wsm()[SH::O_W1 + k * NS + i];
        const float b1 = c.wsm()[SH::O_B1 + k];
        float av[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) av[r] = tanh_acc(dotn<NS>(w1, zi[r], b1));
        st8(c.row_own(c.bufA(), k), av);
        if constexpr (SH::HAS_R) {
            float wr[NS];
#pragma unroll
            for (int i = 0; i < NS; ++i) wr[i] = c.wsm()[SH::O_WR1 + k * NS + i];
            const float br = c.wsm()[SH::O_BR1 + k];
#pragma unroll
            for (int r = 0; r < 8; ++r) r1[r][o] = tanh_acc(dotn<NS>(wr, yi[r], br));
        }
        if constexpr (SH::HAS_GNET) {
            float wg_[NS];
#pragma unroll
            for (int i = 0; i < NS; ++i) wg_[i] = c.wsm()[SH::O_WG1 + k * NS + i];
            const float bg = c.wsm()[SH::O_BG1 + k];
#pragma unroll
            for (int r = 0; r < 8; ++r) ag[r][o] = tanh_acc(dotn<NS>(wg_, yi[r], bg));
        }
    }
    if constexpr (SH::HAS_R) {
        constexpr int CH = NN < 8 ? NN : 8;
#pragma unroll
        for (int pass = 0; pass < NN / CH; ++pass) {
            float P[8 * CH];
#pragma unroll
            for (int cc = 0; cc < CH; ++cc) {
                float w[8];
                ldw8(c, c.wsm() + SH::O_WR2 + (pass * CH + cc) * HID, w);
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    float s = 0.f;
#pragma unroll
                    for (int o = 0; o < 8; ++o) s = fmaf(w[o], r1[r][o], s);
                    P[r * CH + cc] = s;
                }
            }
            lane8_reduce<8 * CH>(P, c.lo);
#pragma unroll
            for (int cc = 0; cc < CH; ++cc) c.part()[(c.wg * 32 + c.lane) * SH::PW + SH::SL_R + pass * CH + cc] = P[cc];
        }
    }
    if constexpr (SH::HAS_GNET) {
        float P[8 * NS];
#pragma unroll
        for (int a = 0; a < NS; ++a) {
            float w[8];
            ldw8(c, c.wsm() + SH::O_WG2 + a * HID, w);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                float s = 0.f;
#pragma unroll
                for (int o = 0; o < 8; ++o) s = fmaf(w[o], ag[r][o], s);
                P[r * NS + a] = s;
            }
        }
        lane8_reduce<8 * NS>(P, c.lo);
#pragma unroll
        for (int a = 0; a < NS; ++a) c.part()[(c.wg * 32 + c.lane) * SH::PW + SH::SL_G + a] = P[a];
    }
}

// ---------------------------------------------------------------------------------------
// dynamics evaluation f(y,u), H(y) for the 32 instances of the group
//   pHNN.forward            src/pHNN.py:52-100
//   pHNN_Canonical.forward  src/pHNN_canonical.py:172-273
// ---------------------------------------------------------------------------------------
template <class SH>
__device__ __noinline__ void eval_fwd(Ctx<SH>& c, const KParams& p, const float (&y)[SH::NS], float u,
                                      float (&f)[SH::NS], float& Hval) {
    constexpr int NS = SH::NS;
    float z[NS];
    Canon cq = {};
    if constexpr (SH::MK == MK_CANON) {
        cq = canon_of(p, y[1]);
        z[0] = y[0];
        z[1] = y[1];
        z[2] = p.ma * y[2] + cq.beta * y[3];  // p = M(q) qdot
        z[3] = cq.beta * y[2] + p.mc * y[3];
    } else {
#pragma unroll
        for (int i = 0; i < NS; ++i) z[i] = y[i];
    }
    {
        float zi[8][NS];
        gather8<NS>(z, c.li, zi);
        layer1_and_aux(c, zi, zi);
    }
    c.gbar();  // F1: bufA, aux partials visible
    float acc[8][8];
    product(c, c.bufA(), acc);  // z2 = W2 a1
    {
        float hp[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) hp[r] = 0.f;
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            const int j = c.kown(o);
            const float b2 = c.wsm()[SH::O_B2 + j], w3 = c.wsm()[SH::O_W3 + j];
            float d2[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float a2 = tanh_acc(acc[r][o] + b2);
                hp[r] = fmaf(w3, a2, hp[r]);
                d2[r] = fmaf(-a2, a2, 1.f) * w3;
            }
            st8(c.row_own(c.bufB(), j), d2);
        }
        lane8_reduce<8>(hp, c.lo);
        c.part()[(c.wg * 32 + c.lane) * SH::PW + SH::SL_H] = hp[0];
    }
    c.gbar();             // F2: bufB visible
    product(c, c.bufB(), acc);  // g1 = W2^T delta2
    {
        float gp[8 * NS];
#pragma unroll
        for (int e = 0; e < 8 * NS; ++e) gp[e] = 0.f;
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            const int k = c.kown(o);
            float a1[8];
            ld8(c.row_own(c.bufA(), k), a1);
            float w1[NS];
#pragma unroll
            for (int i = 0; i < NS; ++i) w1[i] = c.wsm()[SH::O_W1 + k * NS + i];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float d1 = fmaf(-a1[r], a1[r], 1.f) * acc[r][o];
#pragma unroll
                for (int i = 0; i < NS; ++i) gp[r * NS + i] = fmaf(w1[i], d1, gp[r * NS + i]);
            }
        }
        lane8_reduce<8 * NS>(gp, c.lo);
#pragma unroll
        for (int i = 0; i < NS; ++i) c.part()[(c.wg * 32 + c.lane) * SH::PW + SH::SL_DH + i] = gp[i];
    }
    c.gbar();  // F3: all partials visible
    float g[NS], hs[1];
    load_part<SH, SH::SL_DH, NS>(c, g);
    load_part<SH, SH::SL_H, 1>(c, hs);
    Hval = hs[0] + p.b3;
    if constexpr (SH::MK == MK_CANON) {
        float pd[2];
#pragma unroll
        for (int r = 2; r < 4; ++r) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) s = fmaf(p.Jm[r * 4 + k] - (r == k ? p.rdiag[r] : 0.f), g[k], s);
            pd[r - 2] = s + p.Gv[r] * u;
        }
        f[0] = cq.n11 * z[2] + cq.n12 * z[3];
        f[1] = cq.n12 * z[2] + cq.n22 * z[3];
        f[2] = cq.n11 * pd[0] + cq.n12 * pd[1];
        f[3] = cq.n12 * pd[0] + cq.n22 * pd[1];
    } else {
        float S[NS][NS];
        load_S(c, p, S);
        float G[NS];
        if constexpr (SH::HAS_GNET) {
            load_part<SH, SH::SL_G, NS>(c, G);
#pragma unroll
            for (int a = 0; a < NS; ++a) G[a] += p.bg2[a];
        } else {
#pragma unroll
            for (int a = 0; a < NS; ++a) G[a] = p.Gv[a];
        }
#pragma unroll
        for (int a = 0; a < NS; ++a) {
            float s = 0.f;
#pragma unroll
            for (int b = 0; b < NS; ++b) {
                float Rab = 0.f;
#pragma unroll
                for (int k = 0; k < NS; ++k) Rab = fmaf(S[a][k], S[b][k], Rab);
                s = fmaf(p.Jm[a * NS + b] - Rab, g[b], s);
            }
            f[a] = s + G[a] * u;
        }
    }
    c.gbar();  // F4: partial slots may be rewritten
}

// ---------------------------------------------------------------------------------------
// vector-Jacobian product of f at (y,u): xbar = (df/dy)^T v, ubar = (df/du)^T v.
// Recomputes the activations (no tape) and applies the Hessian-vector product of H_net
// (SURVEY.md Appendix A).
// ---------------------------------------------------------------------------------------
template <class SH>
__device__ __noinline__ void eval_vjp(Ctx<SH>& c, const KParams& p, const float (&y)[SH::NS], float u,
                                      const float (&v)[SH::NS], float (&xbar)[SH::NS], float& ubar) {
    constexpr int NS = SH::NS, NN = SH::NN, HID = SH::HID;
    float z[NS], w[NS];
    Canon cq = {};
    float pb[2] = {0.f, 0.f}, pdb[2] = {0.f, 0.f};
    if constexpr (SH::MK == MK_CANON) {
        cq = canon_of(p, y[1]);
        z[0] = y[0];
        z[1] = y[1];
        z[2] = p.ma * y[2] + cq.beta * y[3];
        z[3] = cq.beta * y[2] + p.mc * y[3];
        pb[0] = cq.n11 * v[0] + cq.n12 * v[1];
        pb[1] = cq.n12 * v[0] + cq.n22 * v[1];
        pdb[0] = cq.n11 * v[2] + cq.n12 * v[3];
        pdb[1] = cq.n12 * v[2] + cq.n22 * v[3];
        // w = (J - diag r)^T [0,0,pdb]
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float s = 0.f;
#pragma unroll
            for (int r = 2; r < 4; ++r) s = fmaf(p.Jm[r * 4 + k] - (r == k ? p.rdiag[r] : 0.f), pdb[r - 2], s);
            w[k] = s;
        }
    } else {
#pragma unroll
        for (int i = 0; i < NS; ++i) z[i] = y[i];
    }
    {
        float zi[8][NS];
        gather8<NS>(z, c.li, zi);
        layer1_and_aux(c, zi, zi);
    }
    c.gbar();  // A1
    if constexpr (SH::MK != MK_CANON) {
        float S[NS][NS], sv[NS];
        load_S(c, p, S);
#pragma unroll
        for (int a = 0; a < NS; ++a) {
            float s = 0.f;
#pragma unroll
            for (int b = 0; b < NS; ++b) s = fmaf(S[a][b], v[b], s);
            sv[a] = s;
        }
        // w = A^T v = (J - J^T)^T v - S (S v)
#pragma unroll
        for (int a = 0; a < NS; ++a) {
            float s = 0.f;
#pragma unroll
            for (int b = 0; b < NS; ++b) s = fmaf(p.Jm[b * NS + a], v[b], s);
#pragma unroll
            for (int b = 0; b < NS; ++b) s = fmaf(-S[a][b], sv[b], s);
            w[a] = s;
        }
    }
    float acc[8][8], keep[8][8];
    product(c, c.bufA(), acc);  // z2 = W2 a1
#pragma unroll
    for (int o = 0; o < 8; ++o) {
        const int j = c.kown(o);
        const float b2 = c.wsm()[SH::O_B2 + j], w3 = c.wsm()[SH::O_W3 + j];
        float d2[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const float a2 = tanh_acc(acc[r][o] + b2);
            keep[r][o] = a2;
            d2[r] = fmaf(-a2, a2, 1.f) * w3;
        }
        st8(c.row_own(c.bufB(), j), d2);
    }
    c.gbar();  // A2: every warp is done reading a1 from bufA
    float wi[8][NS];
    gather8<NS>(w, c.li, wi);
#pragma unroll
    for (int o = 0; o < 8; ++o) {  // da1 = (1 - a1^2) * (W1 w) -> bufA (own elements)
        const int k = c.kown(o);
        float a1[8];
        ld8(c.row_own(c.bufA(), k), a1);
        float w1[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) w1[i] = c.wsm()[SH::O_W1 + k * NS + i];
#pragma unroll
        for (int r = 0; r < 8; ++r) a1[r] = fmaf(-a1[r], a1[r], 1.f) * dotn<NS>(w1, wi[r], 0.f);
        st8(c.row_own(c.bufA(), k), a1);
    }
    c.gbar();  // A3
    product(c, c.bufA(), acc);  // dz2 = W2 da1
#pragma unroll
    for (int o = 0; o < 8; ++o) {  // e2 = d(s2)/dt * w3 = -2 a2 da2 w3
        const float w3 = c.wsm()[SH::O_W3 + c.kown(o)];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const float a2 = keep[r][o];
            const float da2 = fmaf(-a2, a2, 1.f) * acc[r][o];
            keep[r][o] = -2.f * a2 * da2 * w3;
        }
    }
    product(c, c.bufB(), acc);  // g1 = W2^T delta2
    c.gbar();                 // A4: every warp is done reading delta2 from bufB
#pragma unroll
    for (int o = 0; o < 8; ++o) {
        float e2[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) { e2[r] = keep[r][o]; keep[r][o] = acc[r][o]; }
        st8(c.row_own(c.bufB(), c.kown(o)), e2);
    }
    c.gbar();  // A5
    product(c, c.bufB(), acc);  // dg1 = W2^T e2
    // final elementwise stage: xbar_H partials and dH partials
    float xp[8 * NS];
    float zi[8][NS];
    gather8<NS>(z, c.li, zi);
    {
        float gp[8 * NS];
#pragma unroll
        for (int e = 0; e < 8 * NS; ++e) { gp[e] = 0.f; xp[e] = 0.f; }
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            const int k = c.kown(o);
            float w1[NS];
#pragma unroll
            for (int i = 0; i < NS; ++i) w1[i] = c.wsm()[SH::O_W1 + k * NS + i];
            const float b1 = c.wsm()[SH::O_B1 + k];
            float da1[8];
            ld8(c.row_own(c.bufA(), k), da1);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float a1 = tanh_acc(dotn<NS>(w1, zi[r], b1));
                const float s1 = fmaf(-a1, a1, 1.f);
                const float ds1 = -2.f * a1 * da1[r];
                const float g1 = keep[r][o];
                const float t = fmaf(ds1, g1, s1 * acc[r][o]);
                const float d1 = s1 * g1;
#pragma unroll
                for (int i = 0; i < NS; ++i) {
                    xp[r * NS + i] = fmaf(w1[i], t, xp[r * NS + i]);
                    gp[r * NS + i] = fmaf(w1[i], d1, gp[r * NS + i]);
                }
            }
        }
        lane8_reduce<8 * NS>(gp, c.lo);
#pragma unroll
        for (int i = 0; i < NS; ++i) c.part()[(c.wg * 32 + c.lane) * SH::PW + SH::SL_DH + i] = gp[i];
    }
    float G[NS];
    if constexpr (SH::MK != MK_CANON) {
        c.gbar();  // A6: dH partials visible
        float g[NS];
        load_part<SH, SH::SL_DH, NS>(c, g);
        float S[NS][NS], sv[NS], tg[NS];
        load_S(c, p, S);  // the R_net partial sums are still in place
#pragma unroll
        for (int a = 0; a < NS; ++a) {
            float s = 0.f, s2 = 0.f;
#pragma unroll
            for (int b = 0; b < NS; ++b) { s = fmaf(S[a][b], g[b], s); s2 = fmaf(S[a][b], v[b], s2); }
            tg[a] = s;
            sv[a] = s2;
        }
        if constexpr (SH::HAS_GNET) {
            load_part<SH, SH::SL_G, NS>(c, G);
#pragma unroll
            for (int a = 0; a < NS; ++a) G[a] += p.bg2[a];
        } else {
#pragma unroll
            for (int a = 0; a < NS; ++a) G[a] = p.Gv[a];
        }
        if (c.wg == 0) {
            // cotangent of the raw R_net output: -sym(v t^T + g s^T)
#pragma unroll
            for (int a = 0; a < NS; ++a)
#pragma unroll
                for (int b = 0; b < NS; ++b)
                    c.rbar()[(a * NS + b) * GI + c.lane] =
                        -0.5f * (v[a] * tg[b] + g[a] * sv[b] + v[b] * tg[a] + g[b] * sv[a]);
            if constexpr (SH::HAS_GNET) {
#pragma unroll
                for (int a = 0; a < NS; ++a) c.rbar()[(NN + a) * GI + c.lane] = v[a] * u;
            }
        }
        c.gbar();  // A7: rbar visible
        // aux nets backward: hidden cotangent = Wr2^T Rbar (K = NN), times tanh', through Wr1^T
        float rb[8][8];
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int o = 0; o < 8; ++o) rb[r][o] = 0.f;
#pragma unroll
        for (int e = 0; e < NN; ++e) {
            float a[8], wv[8];
            ld8(c.rbar() + e * GI + (c.li << 3), a);
            ldw8(c, c.wsm() + SH::O_WR2 + e * HID, wv);
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int o = 0; o < 8; ++o) rb[r][o] = fmaf(a[r], wv[o], rb[r][o]);
        }
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            const int k = c.kown(o);
            float wr[NS];
#pragma unroll
            for (int i = 0; i < NS; ++i) wr[i] = c.wsm()[SH::O_WR1 + k * NS + i];
            const float br = c.wsm()[SH::O_BR1 + k];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float r1 = tanh_acc(dotn<NS>(wr, zi[r], br));
                const float zb = rb[r][o] * fmaf(-r1, r1, 1.f);
#pragma unroll
                for (int i = 0; i < NS; ++i) xp[r * NS + i] = fmaf(wr[i], zb, xp[r * NS + i]);
            }
        }
        if constexpr (SH::HAS_GNET) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int o = 0; o < 8; ++o) rb[r][o] = 0.f;
#pragma unroll
            for (int e = 0; e < NS; ++e) {
                float a[8], wv[8];
                ld8(c.rbar() + (NN + e) * GI + (c.li << 3), a);
                ldw8(c, c.wsm() + SH::O_WG2 + e * HID, wv);
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int o = 0; o < 8; ++o) rb[r][o] = fmaf(a[r], wv[o], rb[r][o]);
            }
#pragma unroll
            for (int o = 0; o < 8; ++o) {
                const int k = c.kown(o);
                float wr[NS];
#pragma unroll
                for (int i = 0; i < NS; ++i) wr[i] = c.wsm()[SH::O_WG1 + k * NS + i];
                const float br = c.wsm()[SH::O_BG1 + k];
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const float a = tanh_acc(dotn<NS>(wr, zi[r], br));
                    const float zb = rb[r][o] * fmaf(-a, a, 1.f);
#pragma unroll
                    for (int i = 0; i < NS; ++i) xp[r * NS + i] = fmaf(wr[i], zb, xp[r * NS + i]);
                }
            }
        }
    }
    lane8_reduce<8 * NS>(xp, c.lo);
#pragma unroll
    for (int i = 0; i < NS; ++i) c.part()[(c.wg * 32 + c.lane) * SH::PW + SH::SL_X + i] = xp[i];
    c.gbar();  // A8: xbar (and, canonical, dH) partials visible
    float zb[NS];
    load_part<SH, SH::SL_X, NS>(c, zb);
    if constexpr (SH::MK == MK_CANON) {
        float g[NS];
        load_part<SH, SH::SL_DH, NS>(c, g);
        float pd[2];
#pragma unroll
        for (int r = 2; r < 4; ++r) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) s = fmaf(p.Jm[r * 4 + k] - (r == k ? p.rdiag[r] : 0.f), g[k], s);
            pd[r - 2] = s + p.Gv[r] * u;
        }
        const float dbeta = -p.mb * cq.sth;
        const float dD = -2.f * cq.beta * dbeta;
        const float iD2 = 1.f / (cq.D * cq.D);
        const float dn11 = -p.mc * iD2 * dD;
        const float dn12 = -dbeta / cq.D + cq.beta * iD2 * dD;
        const float dn22 = -p.ma * iD2 * dD;
        float thbar = v[0] * (dn11 * z[2] + dn12 * z[3]) + v[1] * (dn12 * z[2] + dn22 * z[3]);
        thbar += v[2] * (dn11 * pd[0] + dn12 * pd[1]) + v[3] * (dn12 * pd[0] + dn22 * pd[1]);
        zb[2] += pb[0];
        zb[3] += pb[1];
        thbar += dbeta * (zb[2] * y[3] + zb[3] * y[2]);
        xbar[0] = zb[0];
        xbar[1] = zb[1] + thbar;
        xbar[2] = p.ma * zb[2] + cq.beta * zb[3];
        xbar[3] = cq.beta * zb[2] + p.mc * zb[3];
        ubar = p.Gv[2] * pdb[0] + p.Gv[3] * pdb[1];
    } else {
#pragma unroll
        for (int i = 0; i < NS; ++i) xbar[i] = zb[i];
        float s = 0.f;
#pragma unroll
        for (int a = 0; a < NS; ++a) s = fmaf(G[a], v[a], s);
        ubar = s;
    }
    c.gbar();  // A9
}

// ---------------------------------------------------------------------------------------
// horizon cost pieces (src/mpc_controller.py:75-114, src/mpc_controller_canonical.py:91-120)
// ---------------------------------------------------------------------------------------
template <int NS>
__device__ __forceinline__ float state_cost(const KParams& p, const float (&x)[NS], float* grad) {
    float e[NS], cost = 0.f;
#pragma unroll
    for (int i = 0; i < NS; ++i) e[i] = x[i] - p.xt[i];
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        float s = 0.f, st = 0.f;
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            s = fmaf(p.Q[i * NS + j], e[j], s);
            st = fmaf(p.Qs[i * NS + j], e[j], st);
        }
        cost = fmaf(e[i], s, cost);
        if (grad) grad[i] = st;
    }
    if (p.has_xmin) {
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            const float viol = p.xmin[i] - x[i];
            if (viol > 0.f) {
                cost = fmaf(p.bw * viol, viol, cost);
                if (grad) grad[i] -= 2.f * p.bw * viol;
            }
        }
    }
    if (p.has_xmax) {
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            const float viol = x[i] - p.xmax[i];
            if (viol > 0.f) {
                cost = fmaf(p.bw * viol, viol, cost);
                if (grad) grad[i] += 2.f * p.bw * viol;
            }
        }
    }
    return cost;
}

__device__ __forceinline__ float clampu(const KParams& p, float u) {
    return p.has_ub ? fminf(fmaxf(u, p.umin), p.umax) : u;
}

// ---------------------------------------------------------------------------------------
// One job for one instance, shared by the FP32-FMA and the tcgen05 kernels.  ENG supplies the
// collective dynamics evaluations (eval_fwd / eval_vjp run by all threads of a tile together),
// gbar() (a barrier among the threads that co-own an instance) and `store` (exactly one of the
// co-owning threads performs the global stores).  tile/slot locate the instance and its
// lane-interleaved workspace.
// ---------------------------------------------------------------------------------------
template <class ENG, class SCHED>
__device__ __forceinline__ void run_job(ENG& c, const KParams& p, SCHED& sched, const int slot) {
    constexpr int NS = ENG::NS, TW = ENG::TW;
    Unit unit;
    float best_reg = __int_as_float(0x7f800000);
#pragma unroll 1
  while (sched.next(unit)) {
    const long long tile = unit.tile;
    const int it = unit.it;
    c.begin_unit(p, tile);
    const long long b = tile * TW + slot;  // my instance
    const bool valid = b < p.B;
    const bool st = valid && c.store;
    const int T = p.T, S = p.S;  // MODE_FORWARD / MODE_VJP arrive with T = S = 1
    const int E = T * S;

    float x0[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) x0[i] = valid ? p.x0[b * NS + i] : 0.f;

    // workspace of this tile: stage states [E][NS][TW] (or the engine's per-CTA scratch), Adam m/v and best controls [T][TW]
    float* const scratch = c.unit_scratch();
    const size_t tile_floats = scratch ? ws_persist_floats_per_tile(T, TW) : ws_floats_per_tile(NS, T, S, TW, ENG::ws_extra(p));
    float* wsg = p.ws ? p.ws + (size_t)tile * tile_floats : nullptr;
    float* ckpt = scratch ? scratch : wsg;
    float* adam_m = wsg ? wsg + (scratch ? (size_t)0 : (size_t)E * NS * TW) : nullptr;
    float* adam_v = adam_m ? adam_m + (size_t)T * TW : nullptr;
    float* ubest = adam_v ? adam_v + (size_t)T * TW : nullptr;
    float* bestws = ubest ? ubest + (size_t)T * TW : nullptr;
    const bool solve = (p.mode == MODE_SOLVE);
    const bool one_vjp = (p.mode == MODE_VJP);
    const bool pgrad = (p.mode == MODE_PARAMGRAD);
    const bool need_adj = solve || one_vjp || pgrad || (p.mode == MODE_COSTGRAD && p.want_grad);
    const float* Uread = solve ? p.U : p.uin;
    float* traj = (p.mode == MODE_ROLLOUT || p.mode == MODE_COSTGRAD) ? p.out0 : nullptr;  // (PARAMGRAD: out0 is dL/dx0)
    // rollout_trajectory's energy list needs one more evaluation, at y_T (src/integrators.py:184)
    const int t_end = one_vjp ? 0 : T + ((p.mode == MODE_ROLLOUT && p.energy_mode == 2) ? 1 : 0);

    if (solve && st && it == 1) {
        for (int t = 0; t < T; ++t) {
            adam_m[t * TW + slot] = 0.f;
            adam_v[t * TW + slot] = 0.f;
            ubest[t * TW + slot] = clampu(p, p.U[b * T + t]);
        }
    }
    float best = (it == 1) ? __int_as_float(0x7f800000) : best_reg;
    if (SCHED::kStateInWorkspace) best = (it == 1 || !valid) ? __int_as_float(0x7f800000) : __ldcg(bestws + slot);

    {
        // ---------------- forward sweep ----------------
        float x[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) x[i] = x0[i];
        float cost = 0.f;
#pragma unroll 1
        for (int t = 0; t < t_end; ++t) {
            const bool tail = (t == T);  // energy-only evaluation at the final state
            const float uraw = (valid && !tail) ? __ldcg(Uread + b * T + t) : 0.f;
            const float u = clampu(p, uraw);
            if (p.mode == MODE_COSTGRAD || solve) {
                cost += state_cost<NS>(p, x, nullptr);
                cost = fmaf(p.Rw * u, u, cost);
            }
            if (traj && st) {
#pragma unroll
                for (int i = 0; i < NS; ++i) traj[(b * (T + 1) + t) * NS + i] = x[i];
            }
            float k[NS], ksum[NS], ys[NS], Hv;
#pragma unroll
            for (int i = 0; i < NS; ++i) { ys[i] = x[i]; ksum[i] = 0.f; }
#pragma unroll 1
            for (int s = 0; s < S; ++s) {
                if (need_adj && st) {
#pragma unroll
                    for (int i = 0; i < NS; ++i) ckpt[((size_t)(t * S + s) * NS + i) * TW + slot] = ys[i];
                }
                c.set_eval(t * S + s);
                c.eval_fwd(p, ys, u, k, Hv);
                if (p.mode == MODE_FORWARD) {
                    if (st) {
#pragma unroll
                        for (int i = 0; i < NS; ++i) p.out0[b * NS + i] = k[i];
                        p.out1[b] = Hv;
                    }
                    return;
                }
                if (s == 0 && p.mode == MODE_ROLLOUT && p.out1 && st) {
                    // H(y_t): differentiable-rollout ordering puts it at index t+1 (and 0)
                    if (p.energy_mode == 1) {
                        p.out1[b * (T + 1) + t + 1] = Hv;
                        if (t == 0) p.out1[b * (T + 1)] = Hv;
                    } else {
                        p.out1[b * (T + 1) + t] = Hv;
                    }
                }
                if (tail) break;
                if (S == 1) {
#pragma unroll
                    for (int i = 0; i < NS; ++i) x[i] = fmaf(p.dt, k[i], x[i]);
                } else {
                    // classical RK4, u held over the step (src/integrators.py:66-82)
                    if (s == 0) {
#pragma unroll
                        for (int i = 0; i < NS; ++i) { ksum[i] = k[i]; ys[i] = fmaf(p.dt2, k[i], x[i]); }
                    } else if (s == 1) {
#pragma unroll
                        for (int i = 0; i < NS; ++i) { ksum[i] = fmaf(2.f, k[i], ksum[i]); ys[i] = fmaf(p.dt2, k[i], x[i]); }
                    } else if (s == 2) {
#pragma unroll
                        for (int i = 0; i < NS; ++i) { ksum[i] = fmaf(2.f, k[i], ksum[i]); ys[i] = fmaf(p.dt, k[i], x[i]); }
                    } else {
#pragma unroll
                        for (int i = 0; i < NS; ++i) x[i] = fmaf(p.dt6, ksum[i] + k[i], x[i]);
                    }
                }
            }
        }
        if (p.mode == MODE_COSTGRAD || solve) cost += state_cost<NS>(p, x, nullptr);
        if (traj && st && t_end == T) {
#pragma unroll
            for (int i = 0; i < NS; ++i) traj[(b * (T + 1) + T) * NS + i] = x[i];
        }
        if (p.mode == MODE_ROLLOUT) return;
        if (p.mode == MODE_COSTGRAD && st) p.cost[b] = cost;
        if (solve && st && p.cost_hist) p.cost_hist[(size_t)(it - 1) * p.B + b] = cost;
        if (!need_adj) return;

        // ---------------- reverse sweep (discrete adjoint) ----------------
        const bool improved = cost < best;
        if (improved) best = cost;
        float step_size = 0.f, bc2s = 1.f;
        if (solve) {
            // torch.optim.Adam (single-tensor path): bias corrections in double
            const double bc1 = 1.0 - pow(p.beta1, (double)it);
            const double bc2 = 1.0 - pow(p.beta2, (double)it);
            step_size = (float)(p.lr / bc1);
            bc2s = (float)sqrt(bc2);
        }
        const float w1 = (float)(1.0 - p.beta1), b2f = (float)p.beta2, w2 = (float)(1.0 - p.beta2), epsf = (float)p.eps;

        float lam[NS];
        if (one_vjp) {
#pragma unroll
            for (int i = 0; i < NS; ++i) lam[i] = valid ? p.vin[b * NS + i] : 0.f;
        } else if (pgrad) {
#pragma unroll
            for (int i = 0; i < NS; ++i) lam[i] = valid ? p.gtraj[(b * (T + 1) + T) * NS + i] : 0.f;
        } else {
            state_cost<NS>(p, x, lam);
        }
#pragma unroll 1
        for (int t = T - 1; t >= 0; --t) {
            const float uraw = valid ? __ldcg(Uread + b * T + t) : 0.f;
            const float u = one_vjp ? uraw : clampu(p, uraw);
            float ubsum = 0.f, y[NS], xb[NS], ub, kb[NS], ysum[NS];
#pragma unroll
            for (int i = 0; i < NS; ++i) { ysum[i] = 0.f; xb[i] = 0.f; }
#pragma unroll 1
            for (int s = S - 1; s >= 0; --s) {
                if (one_vjp) {
#pragma unroll
                    for (int i = 0; i < NS; ++i) { y[i] = x0[i]; kb[i] = lam[i]; }
                } else {
#pragma unroll
                    for (int i = 0; i < NS; ++i)
                        y[i] = valid ? __ldcg(ckpt + ((size_t)(t * S + s) * NS + i) * TW + slot) : 0.f;
                    if (S == 1) {
#pragma unroll
                        for (int i = 0; i < NS; ++i) kb[i] = p.dt * lam[i];
                    } else if (s == 3) {
#pragma unroll
                        for (int i = 0; i < NS; ++i) kb[i] = p.dt6 * lam[i];
                    } else if (s == 2) {
#pragma unroll
                        for (int i = 0; i < NS; ++i) kb[i] = fmaf(p.dt3, lam[i], p.dt * xb[i]);
                    } else if (s == 1) {
#pragma unroll
                        for (int i = 0; i < NS; ++i) kb[i] = fmaf(p.dt3, lam[i], p.dt2 * xb[i]);
                    } else {
#pragma unroll
                        for (int i = 0; i < NS; ++i) kb[i] = fmaf(p.dt6, lam[i], p.dt2 * xb[i]);
                    }
                }
                c.set_eval(t * S + s);
                c.eval_vjp(p, y, u, kb, xb, ub);
                if (one_vjp) {
                    if (st) {
#pragma unroll
                        for (int i = 0; i < NS; ++i) p.out0[b * NS + i] = xb[i];
                        p.out1[b] = ub;
                    }
                    return;
                }
                ubsum += ub;
#pragma unroll
                for (int i = 0; i < NS; ++i) ysum[i] += xb[i];
            }
            // y now holds x_t (stage 0 state)
            float gl[NS];
            if (pgrad) {
#pragma unroll
                for (int i = 0; i < NS; ++i) gl[i] = valid ? p.gtraj[(b * (T + 1) + t) * NS + i] : 0.f;
            } else {
                state_cost<NS>(p, y, gl);
            }
#pragma unroll
            for (int i = 0; i < NS; ++i) lam[i] += ysum[i] + gl[i];
            float g = pgrad ? ubsum : fmaf(2.f * p.Rw, u, ubsum);
            if (p.has_ub && !(uraw >= p.umin && uraw <= p.umax)) g = 0.f;  // clamp is inside the graph
            if (p.mode == MODE_COSTGRAD || pgrad) {
                if (st && p.dJdU) p.dJdU[b * T + t] = g;
            } else if (st) {
                if (improved) ubest[t * TW + slot] = u;
                float m = __ldcg(adam_m + t * TW + slot), vv = __ldcg(adam_v + t * TW + slot);
                m = m + w1 * (g - m);
                vv = vv * b2f + w2 * g * g;
                adam_m[t * TW + slot] = m;
                adam_v[t * TW + slot] = vv;
                const float den = sqrtf(vv) / bc2s + epsf;
                p.U[b * T + t] = uraw + (-step_size * m) / den;
            }
        }
        if (pgrad && st && p.out0) {
#pragma unroll
            for (int i = 0; i < NS; ++i) p.out0[b * NS + i] = lam[i];
        }
        best_reg = best;
        if (SCHED::kStateInWorkspace && st) bestws[slot] = best;
    }
    if (solve && st && it == p.iters) {
        for (int t = 0; t < T; ++t)
            p.U[b * T + t] = (p.return_mode == 0) ? clampu(p, __ldcg(p.U + b * T + t)) : __ldcg(ubest + t * TW + slot);
        if (p.cost) p.cost[b] = best;
    }
    sched.done(c, unit);
    if (solve && it == p.iters && p.npeer > 0) c.peer_store(p, tile);
  }
  // a solve with zero iterations still clamps / copies the initial guess (src/mpc_controller.py:203-207)
  if (p.mode == MODE_SOLVE && p.iters == 0) {
      const long long tile0 = sched.tile0();
      const long long b = tile0 * TW + slot;
      if (tile0 >= 0 && b < p.B && c.store) {
          for (int t = 0; t < p.T; ++t) p.U[b * p.T + t] = clampu(p, p.U[b * p.T + t]);
          if (p.cost) p.cost[b] = __int_as_float(0x7f800000);
      }
  }
}

// ---------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------
template <int MK, int NS, int HID>
__global__ void __launch_bounds__(MAX_THREADS, 1) phnn_kernel(const __grid_constant__ KParams p) {
    using SH = Shape<MK, NS, HID>;
    uint64_t* bars = reinterpret_cast<uint64_t*>(phnn_smem);  // [0,NST) full, [NST,2NST) empty, [2NST] small
    float* sm = reinterpret_cast<float*>(phnn_smem + 128);
    float* wsm = sm;
    float* ring = sm + SH::SMALL;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ncw = p.ng * SH::NWG;  // consumer warps
    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(&bars[s], 1);
            mbar_init(&bars[NST + s], ncw);
        }
        mbar_init(&bars[2 * NST], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    // per-job evaluation schedule (shared by the producer and the consumers)
    const int E = p.T * p.S;
    int n_outer = 1, nfwd = 0, nadj = 0;
    switch (p.mode) {
        case MODE_FORWARD: nfwd = 1; break;
        case MODE_VJP: nadj = 1; break;
        case MODE_ROLLOUT: nfwd = E + (p.energy_mode == 2 ? 1 : 0); break;
        case MODE_COSTGRAD: nfwd = E; nadj = p.want_grad ? E : 0; break;
        default: n_outer = p.iters; nfwd = E; nadj = E; break;
    }

    if (warp == ncw) {
        // ===== producer warp: TMA bulk copies of the weight tiles =====
        if (lane == 0) {
            mbar_expect_tx(&bars[2 * NST], SH::SMALL * 4);
            bulk_g2s(wsm, p.wsmall, SH::SMALL * 4, &bars[2 * NST]);
            uint32_t t = 0;
            auto sweep = [&](const float* src) {
#pragma unroll 1
                for (int i = 0; i < HID / SH::KT; ++i) {
                    const uint32_t s = t % NST, ph = (t / NST) & 1u;
                    mbar_wait(&bars[NST + s], ph ^ 1u);
                    mbar_expect_tx(&bars[s], SH::STAGE_BYTES);
                    bulk_g2s(ring + s * (SH::KT * HID), src + (size_t)i * SH::KT * HID, SH::STAGE_BYTES, &bars[s]);
                    ++t;
                }
            };
            const float* W2T = p.wbig;
            const float* W2 = p.wbig + (size_t)HID * HID;
#pragma unroll 1
            for (int it = 0; it < n_outer; ++it) {
#pragma unroll 1
                for (int e = 0; e < nfwd; ++e) { sweep(W2T); sweep(W2); }
#pragma unroll 1
                for (int e = 0; e < nadj; ++e) { sweep(W2T); sweep(W2T); sweep(W2); sweep(W2); }
            }
        }
        return;
    }

    // ===== consumer warps =====
    Ctx<SH> c;
    const int grp = warp / SH::NWG;
    c.lane = lane;
    c.li = lane >> 3;
    c.lo = lane & 7;
    c.wg = warp % SH::NWG;
    c.barid = 1 + grp;
    c.wcol = 64 * c.wg + 4 * c.lo;
    c.chunk = c.li ^ (c.lo & 3);
    c.tile = 0;
    c.ready = false;
    c.store = (c.wg == 0);
    c.goff = SH::SMALL + SH::RING + grp * SH::G_FLOATS;
    mbar_wait(&bars[2 * NST], 0);  // small weights landed

    StaticSched sched{(long long)blockIdx.x * p.ng + grp, n_outer, 0};
    run_job(c, p, sched, lane);
}

}  // namespace phnn
