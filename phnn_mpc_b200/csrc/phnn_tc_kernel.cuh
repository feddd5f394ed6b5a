// phnn_tc_kernel.cuh -- tcgen05 / TMEM version of the fused pHNN-MPC kernel (sm_100a).
//
// Same jobs and same per-instance arithmetic as phnn_kernel.cuh (run_job is shared); only the
// two h x h layers of H_net -- the dense contraction [128 instances, h] x [h, h] that carries
// >90 % of the FLOPs -- move to the 5th-generation tensor cores:
//
//   * tile = 128 instances = UMMA M.  Accumulators (z2 / g1, then dz2 / dg1 in the adjoint)
//     live in TMEM: 2 x h fp32 columns, all 512 columns at h = 256.
//   * operand A (activations) is produced by the element threads, 32 hidden units (one
//     128-byte swizzle row of tf32) at a time, straight into a 2-slot shared-memory ring in
//     the K-major SWIZZLE_128B layout the UMMA descriptor expects; the MMA of one product
//     overlaps with the elementwise work that produces the next K-block, so layer outputs
//     never exist as whole [128, h] matrices.
//   * operand B (W2 for z2/dz2, W2^T for g1/dg1) is pre-swizzled on the host and streamed from
//     L2 by a producer warp with cp.async.bulk into a 3-entry ring.
//   * FP32 accuracy on TF32 tensor cores by error compensation (3xTF32): x = hi + lo with
//     hi = cvt.rna.tf32(x); A*B ~= Ahi*Bhi + Alo*Bhi + Ahi*Blo, accumulated in FP32 in TMEM
//     (measured 2e-6 relative on a K=256 product vs 6e-7 for an FP32 FMA chain and 3e-4 for
//     plain TF32; tools/tc_probe.cu).  split = 1 selects plain TF32 (looser, stated tolerance).
//   * kinds: pHNN with fixed G (R_net on the element threads, its output layer folded with the
//     symmetrisation to 10 columns) and the canonical pHNN (mass-matrix transforms per thread).
//   * thread = instance: element thread (row, half) owns 16 of every 32 hidden units of one
//     instance (tcgen05.ld 32x32b gives a thread its own TMEM lane), so every reduction over
//     the hidden dimension is a private register loop; the two halves of an instance meet in
//     one small shared-memory exchange per reduction.
//   * the forward sweep of a cost/solve job keeps a tape: a1, a2 and g1 = W2^T(s2 * w3) of every
//     evaluation go to a per-CTA region of the workspace (HBM; 3 h floats per instance and
//     evaluation), together with the R_net sums and grad H.  The adjoint evaluation then needs
//     only the two Hessian-vector products (dz2 = W2 da1, dg1 = W2^T e2) instead of recomputing
//     z2 and g1 first: 4 tensor products per (forward, adjoint) pair instead of 6.  The producer
//     warp prefetches the tape of the next adjoint evaluation into L2 (cp.async.bulk.prefetch).
//   * element work that does not feed the tensor pipe is folded into the phases that do (their
//     K-block loops wait for the MMA most of the time): the g1 half of xbar and half of the R_net
//     backward chain ride in the da1 loop, the other half of the R_net chain in the e2 loop.
//
// Roles: warps 0-7 element threads (256), warp 8 MMA issuer (one lane), warp 9 weight producer.
#pragma once
#include "phnn_kernel.cuh"

namespace phnn {

template <int MK_, int NS_, int HID_>
struct TcShape {
    static_assert((MK_ == MK_PHNN || MK_ == MK_CANON) && NS_ == 4, "tensor-core path: cart-pole pHNN (fixed G) and canonical pHNN, n = 4");
    static constexpr int MK = MK_, NS = NS_, HID = HID_, NN = NS * NS;
    static constexpr bool HAS_R = (MK != MK_CANON);
    static constexpr int TM = 128;         // instances per tile (UMMA M)
    static constexpr int NKB = HID / 32;   // K-blocks of 32 tf32 (one 128-byte swizzle row)
    static constexpr int A_TILE = TM * 128;   // bytes of one A K-block (hi or lo)
    static constexpr int B_TILE = HID * 128;  // bytes of one B K-block (hi or lo)
    static constexpr int NBE = 3;             // B ring entries
    static constexpr int TMEM_COLS = (2 * HID <= 32) ? 32 : (2 * HID <= 64) ? 64 : (2 * HID <= 128) ? 128 : (2 * HID <= 256) ? 256 : 512;
    // small weights (floats): recA[k] = {W1[k][0..3], b1, b2, w3, br1}, recB[k] = Wr1[k][0..3],
    // recC[k] = the 10 symmetrised R_net output weights (Wr2[ab][k] + Wr2[ba][k])/2, a <= b, + 2 pad
    static constexpr int NSYM = 10, RC = 12;
    static constexpr int O_RA = 0;
    static constexpr int O_RB = O_RA + HID * 8;
    static constexpr int O_RC = O_RB + (HAS_R ? HID * 4 : 0);
    static constexpr int SMALL = O_RC + (HAS_R ? HID * RC : 0);
    static constexpr int XW = 12 + 1 + NS;    // floats per thread in the pair exchange (S partial, H, dH)
    // shared memory map (bytes)
    static constexpr int OFF_A = 1024;                        // 2 slots x (hi, lo)
    static constexpr int OFF_B = OFF_A + 4 * A_TILE;          // NBE entries
    static constexpr int OFF_SMALL = OFF_B + NBE * B_TILE;
    static constexpr int OFF_XCH = OFF_SMALL + SMALL * 4;
    static constexpr int SMEM_BYTES = OFF_XCH + XW * 256 * 4;
    static_assert(SMEM_BYTES <= 232448, "shared memory budget");
    // barrier indices
    static constexpr int B_AFULL = 0, B_AEMPTY = 2, B_BFULL = 4, B_BEMPTY = 4 + NBE, B_ACC = 4 + 2 * NBE, B_SMALL = B_ACC + 2;
    static constexpr int THREADS = 320;
};

// big-blob layout for the tensor path: [P: 0 = W2 (B[n=j][k]), 1 = W2^T (B[n=k][K=j])][kb][hi|lo][B_TILE]
__host__ __device__ inline int sw128_off(int r, int c) {
    return (r >> 3) * 1024 + (r & 7) * 128 + ((((c >> 2) ^ (r & 7)) & 7) << 4) + (c & 3) * 4;
}

// ---- tcgen05 helpers ---------------------------------------------------------------------
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    // K-major, SWIZZLE_128B, 8-row groups 1024 B apart, descriptor version 1 (validated by tools/tc_probe.cu)
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
#ifdef PHNN_TC_EXP_NOMMA  // timing experiment (wrong results): element-side time alone
    return;
#endif
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// round-to-nearest TF32 of a finite float in two integer ops (cvt.rna.tf32.f32 expands to four on
// sm_100a because of its Inf/NaN guard; activations here are finite)
__device__ __forceinline__ float tf32_rn(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
// tanh for the tensor-core path: 1 - 2/(exp(2x)+1) with MUFU.EX2 + MUFU.RCP (absolute error ~1.5e-7),
// switched to x - x^3/3 below |x| = 0.04 where the exp form would lose relative accuracy.  10
// instructions instead of tanhf's 14; its error is below that of the 3xTF32 products (2e-6) that
// consume the activations, so the FP32 tolerances of the parity tests still hold (they run with it).
__device__ __forceinline__ float tanh_tc(float x) {
#ifdef PHNN_TC_TANH_LIBM
    return tanhf(x);
#else
    const float e = ex2_approx(x * 2.8853900817779268f);
    const float r = rcp_approx(e + 1.0f);
    const float big = fmaf(-2.0f, r, 1.0f);
    const float x2 = x * x;
    const float small = x * fmaf(x2, -0.33333334f, 1.0f);
    return fabsf(x) < 0.04f ? small : big;
#endif
}
// keeps the compiler from hoisting the next chunk's loads above this point: without it ptxas
// front-loads a whole K-block of weight loads and then serialises the tanh chains on one register
__device__ __forceinline__ void sched_fence() { asm volatile("" ::: "memory"); }
// asynchronous TMEM load of 16 consecutive columns of this thread's lane; the registers are
// valid only after tmem_wait16 (which also ties them to the wait for the compiler)
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait16(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}
// visit the NKB 32-column blocks of an accumulator with the next block's load in flight while
// the current one is processed: body(block index, 16 values of this thread's half)
template <int NKB, class F>
__device__ __forceinline__ void for_acc_blocks(uint32_t tacc, F&& body) {
    uint32_t ra[16], rb[16];
    tmem_ld16_issue(tacc, ra);
#pragma unroll 1
    for (int b = 0; b < NKB; b += 2) {
        tmem_wait16(ra);
        tmem_ld16_issue(tacc + (b + 1) * 32, rb);
        body(b, ra);
        tmem_wait16(rb);
        if (b + 2 < NKB) tmem_ld16_issue(tacc + (b + 2) * 32, ra);
        body(b + 1, rb);
    }
}
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity, unsigned ns) {
    while (!mbar_try(bar, parity)) __nanosleep(ns);
}
__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float dot4(const float4& w, const float (&x)[4], float b) {
    return fmaf(w.w, x[3], fmaf(w.z, x[2], fmaf(w.y, x[1], fmaf(w.x, x[0], b))));
}

#ifdef PHNN_TC_PROFILE
#define TCP_MARK(c, i) do { long long t_ = clock64(); (c).prof[i] += t_ - (c).tlast; (c).tlast = t_; } while (0)
#else
#define TCP_MARK(c, i) do { } while (0)
#endif

// scheduling fence granularity in 4-unit chunks: 4 = one fence per K-block (measured 4.8 % faster than 1)
#ifndef PHNN_TC_PROD_SLEEP
#define PHNN_TC_PROD_SLEEP 128
#endif
#ifndef PHNN_TC_MMA_SLEEP
#define PHNN_TC_MMA_SLEEP 128
#endif
#ifndef PHNN_TC_FENCE_EVERY
#define PHNN_TC_FENCE_EVERY 4
#endif

template <class SH> struct TcCtx;
template <class SH>
__device__ __forceinline__ void tc_eval_fwd(TcCtx<SH>& c, const KParams& p, const float (&y)[4], float u, float (&f)[4], float& Hval);
template <class SH>
__device__ __forceinline__ void tc_eval_vjp(TcCtx<SH>& c, const KParams& p, const float (&y)[4], float u, const float (&v)[4],
                            float (&xbar)[4], float& ubar);

template <class SH>
struct TcCtx {
    static constexpr int NS = SH::NS;
    static constexpr int TW = SH::TM;
    __device__ static int ws_extra(const KParams& p) { return tc_ws_extra(SH::HID, p.T, p.S); }
    __device__ __forceinline__ void set_eval(int e) { ev = e; }
    int row, hf, lane, barid;
    uint32_t tlane;  // TMEM base address with this warp's lane quadrant
    uint32_t ablk;   // A K-blocks produced so far
    uint32_t qdone;  // products whose accumulator this thread has waited for
    int split;       // 3 = 3xTF32, 1 = plain TF32
    float* sck;      // per tile: R_net sums [0,10) and grad H [10,14) of every forward evaluation, [T*S][16][128]
    int ev;          // index of the evaluation in flight (t * S + s)
    float* tape;     // per CTA: a2, a1, g1 of every forward evaluation of the unit in flight, each [NKB][2][4][128]
                     // float4 (a warp's 32 rows read/write 512 contiguous bytes); nullptr when no adjoint follows
    bool store;
#ifdef PHNN_TC_PROFILE
    long long prof[16];
    long long tlast;
    long long await[8];  // a_begin wait cycles per producing phase
    long long sub[4];    // a_end: fences | syncwarp+arrive ; tmem wait
    int aphase;
#endif

    __device__ __forceinline__ uint64_t* bars() const { return reinterpret_cast<uint64_t*>(phnn_smem); }
    __device__ __forceinline__ const float* small() const { return reinterpret_cast<const float*>(phnn_smem + SH::OFF_SMALL); }
    __device__ __forceinline__ float* xch() const { return reinterpret_cast<float*>(phnn_smem + SH::OFF_XCH); }
    __device__ __forceinline__ void gbar() const { group_bar(barid, 64); }

    // ---- A-operand ring (element threads are the producers) ----
    __device__ __forceinline__ int a_begin() {
        const int slot = ablk & 1;
#ifdef PHNN_TC_PROFILE
        const long long t0 = clock64();
#endif
        mbar_wait(&bars()[SH::B_AEMPTY + slot], ((ablk >> 1) & 1u) ^ 1u);
#ifdef PHNN_TC_PROFILE
        const long long t1 = clock64();
        await[aphase] += t1 - t0;
#endif
        return slot;
    }
    // four consecutive hidden units (chunk q of this thread's 16) of the current K-block
    __device__ __forceinline__ void a_put4(int slot, int q, const float (&v)[4]) const {
        const int ch = hf * 4 + q;
        const int off = (row >> 3) * 1024 + (row & 7) * 128 + ((ch ^ (row & 7)) << 4);
        unsigned char* hi = phnn_smem + SH::OFF_A + (slot * 2) * SH::A_TILE + off;
        float4 h = make_float4(tf32_rn(v[0]), tf32_rn(v[1]), tf32_rn(v[2]), tf32_rn(v[3]));
        *reinterpret_cast<float4*>(hi) = h;
        if (split == 3)
            *reinterpret_cast<float4*>(hi + SH::A_TILE) = make_float4(v[0] - h.x, v[1] - h.y, v[2] - h.z, v[3] - h.w);
    }
    __device__ __forceinline__ void a_end(int slot) {
#ifdef PHNN_TC_PROFILE
        const long long t0 = clock64();
#endif
        tc_fence_before();                                            // earlier tcgen05.ld of this thread are ordered first
#ifndef PHNN_TC_EXP_NOFENCE  // timing experiment (unsafe): cost of the MEMBAR.ALL.CTA the proxy fence lowers to
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async proxy (UMMA)
#endif
#ifdef PHNN_TC_PROFILE
        const long long t1 = clock64();
#endif
#ifdef PHNN_TC_ARRIVE_ALL
        mbar_arrive(&bars()[SH::B_AFULL + slot]);
#else
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars()[SH::B_AFULL + slot]);
#endif
        ++ablk;
#ifdef PHNN_TC_PROFILE
        const long long t2 = clock64();
        sub[0] += t1 - t0;
        sub[1] += t2 - t1;
#endif
    }
    // wait for the next product's accumulator; returns its TMEM address for this thread's lane quadrant
    __device__ __forceinline__ uint32_t acc_wait() {
        const uint32_t q = qdone++;
        mbar_wait(&bars()[SH::B_ACC + (q & 1u)], (q >> 1) & 1u);
        tc_fence_after();
        return tlane + (q & 1u) * SH::HID + hf * 16;
    }
    // float4 slot of (tape array which: 0 a2, 1 a1, 2 g1; K-block jb; chunk q) of evaluation ev for this thread
    static constexpr size_t TAPE_ARR4 = (size_t)SH::NKB * 8 * 128;  // float4 per array
    __device__ __forceinline__ float4* tape4(int which, int jb, int q) const {
        return reinterpret_cast<float4*>(tape) + ((size_t)ev * 3 + which) * TAPE_ARR4 + ((jb * 2 + hf) * 4 + q) * 128 + row;
    }
    __device__ __forceinline__ void tape_load(int which, int jb, float4 (&v)[4]) const {
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = __ldcg(tape4(which, jb, q));
    }
    // pair exchange: returns mine + partner's for n values starting at slot s0
    template <int N>
    __device__ __forceinline__ void exchange(float (&v)[N]) {
        float* x = xch();
        const int t = hf * 128 + row;
#pragma unroll
        for (int i = 0; i < N; ++i) x[i * 256 + t] = v[i];
        gbar();
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] += x[i * 256 + (t ^ 128)];
        gbar();
    }
    __device__ __forceinline__ void begin_unit(const KParams& p, long long tile) {
        const size_t tile_floats = ws_floats_per_tile(NS, p.T, p.S, TW, ws_extra(p));
        sck = p.ws ? p.ws + (size_t)tile * tile_floats + ws_floats_per_tile(NS, p.T, p.S, TW, 0) : nullptr;
    }
    __device__ __forceinline__ void eval_fwd(const KParams& p, const float (&y)[4], float u, float (&f)[4], float& H) {
        tc_eval_fwd(*this, p, y, u, f, H);
    }
    __device__ __forceinline__ void eval_vjp(const KParams& p, const float (&y)[4], float u, const float (&v)[4],
                                             float (&xbar)[4], float& ubar) {
        tc_eval_vjp(*this, p, y, u, v, xbar, ubar);
    }
};

// index of the symmetric pair (a,b) in the 10-entry packed order 00 01 02 03 11 12 13 22 23 33
__host__ __device__ constexpr int sym_idx(int a, int b) {
    return a <= b ? (a * 4 - a * (a - 1) / 2 + (b - a)) : (b * 4 - b * (b - 1) / 2 + (a - b));
}
// S = (Rraw + Rraw^T)/2 from the packed symmetric sums (weights and biases were symmetrised on
// the host: S is linear in the R_net output, src/pHNN.py:79)
__device__ __forceinline__ void tc_make_S(const KParams& p, const float* Sp, float (&S)[4][4]) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) S[a][b] = Sp[sym_idx(a, b)] + p.bsym[sym_idx(a, b)];
}
// accumulate r * recC[k] into the 10 packed S sums
__device__ __forceinline__ void tc_acc_S(const float* rC, int k, float r, float* Sp) {
    const float4 c0 = lds4(rC + k * 12), c1 = lds4(rC + k * 12 + 4);
    const float2 c2 = *reinterpret_cast<const float2*>(rC + k * 12 + 8);
    Sp[0] = fmaf(c0.x, r, Sp[0]); Sp[1] = fmaf(c0.y, r, Sp[1]); Sp[2] = fmaf(c0.z, r, Sp[2]); Sp[3] = fmaf(c0.w, r, Sp[3]);
    Sp[4] = fmaf(c1.x, r, Sp[4]); Sp[5] = fmaf(c1.y, r, Sp[5]); Sp[6] = fmaf(c1.z, r, Sp[6]); Sp[7] = fmaf(c1.w, r, Sp[7]);
    Sp[8] = fmaf(c2.x, r, Sp[8]); Sp[9] = fmaf(c2.y, r, Sp[9]);
}

// phase A of the forward evaluation: a1 = tanh(W1 y + b1) -> A ring (product z2 = W2 a1) and the tape,
// and the R_net hidden layer with its symmetrised output sums
template <class SH>
__device__ __forceinline__ void tc_phase_a1(TcCtx<SH>& c, const float (&z)[4], const float (&y)[4], float* Sp) {
    const float* rA = c.small() + SH::O_RA;
    const float* rB = c.small() + SH::O_RB;
    const float* rC = c.small() + SH::O_RC;
#pragma unroll 1
    for (int kb = 0; kb < SH::NKB; ++kb) {
        const int slot = c.a_begin();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float av[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int k = kb * 32 + c.hf * 16 + q * 4 + e;
                const float4 w1 = lds4(rA + k * 8), m = lds4(rA + k * 8 + 4);
                av[e] = tanh_tc(dot4(w1, z, m.x));
                if constexpr (SH::HAS_R) {
                    const float r = tanh_tc(dot4(lds4(rB + k * 4), y, m.w));
                    tc_acc_S(rC, k, r, Sp);
                }
            }
            if (c.tape) *c.tape4(1, kb, q) = make_float4(av[0], av[1], av[2], av[3]);
            c.a_put4(slot, q, av);
            if (PHNN_TC_FENCE_EVERY == 1 || (q % PHNN_TC_FENCE_EVERY) == PHNN_TC_FENCE_EVERY - 1) sched_fence();
        }
        c.a_end(slot);
    }
}

// canonical model: rows 2,3 of (J - diag r) g + G u  (src/pHNN_canonical.py:227,247)
__device__ __forceinline__ void tc_canon_pdot(const KParams& p, const float* g, float u, float (&pd)[2]) {
#pragma unroll
    for (int r = 2; r < 4; ++r) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) s = fmaf(p.Jm[r * 4 + k] - (r == k ? p.rdiag[r] : 0.f), g[k], s);
        pd[r - 2] = s + p.Gv[r] * u;
    }
}

// ---------------------------------------------------------------------------------------
// f(y,u), H(y) for the tile's 128 instances (src/pHNN.py:52-100)
// ---------------------------------------------------------------------------------------
template <class SH>
__device__ __forceinline__ void tc_eval_fwd(TcCtx<SH>& c, const KParams& p, const float (&y)[4], float u, float (&f)[4],
                                         float& Hval) {
    constexpr int NKB = SH::NKB;
    const float* rA = c.small() + SH::O_RA;
    float X[SH::XW];  // [0,10) S sums, [12] H partial, [13,17) dH partial
#pragma unroll
    for (int i = 0; i < SH::XW; ++i) X[i] = 0.f;
    TCP_MARK(c, 15);
#ifdef PHNN_TC_PROFILE
    c.aphase = 0;
#endif
    float z[4];
    Canon cq = {};
    if constexpr (SH::MK == MK_CANON) {
        cq = canon_of(p, y[1]);
        z[0] = y[0]; z[1] = y[1];
        z[2] = p.ma * y[2] + cq.beta * y[3];  // p = M(q) qdot
        z[3] = cq.beta * y[2] + p.mc * y[3];
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) z[i] = y[i];
    }
    tc_phase_a1(c, z, y, X);
#ifdef PHNN_TC_PROFILE
    c.aphase = 1;
#endif
    TCP_MARK(c, 0);
    // ---- phase B: a2, H, delta2 -> product 2 (g1 = W2^T delta2) ----
    {
        const uint32_t tacc = c.acc_wait();
        TCP_MARK(c, 1);
        float Hp = 0.f;
        for_acc_blocks<NKB>(tacc, [&](int jb, const uint32_t (&zr)[16]) {
            const int slot = c.a_begin();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float dv[4], a2v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int j = jb * 32 + c.hf * 16 + q * 4 + e;
                    const float4 m = lds4(rA + j * 8 + 4);
                    const float a2 = tanh_tc(__uint_as_float(zr[q * 4 + e]) + m.y);
                    a2v[e] = a2;
                    Hp = fmaf(m.z, a2, Hp);
                    dv[e] = fmaf(-a2, a2, 1.f) * m.z;
                }
                if (c.tape) *c.tape4(0, jb, q) = make_float4(a2v[0], a2v[1], a2v[2], a2v[3]);
                c.a_put4(slot, q, dv);
                if (PHNN_TC_FENCE_EVERY == 1 || (q % PHNN_TC_FENCE_EVERY) == PHNN_TC_FENCE_EVERY - 1) sched_fence();
            }
            c.a_end(slot);
        });
        X[12] = Hp;
    }
    TCP_MARK(c, 2);
    // ---- phase C: dH = W1^T (s1 * g1) ----
    {
        const uint32_t tacc = c.acc_wait();
        TCP_MARK(c, 3);
        float g0 = 0.f, g1s = 0.f, g2 = 0.f, g3 = 0.f;
        for_acc_blocks<NKB>(tacc, [&](int kb, const uint32_t (&gr)[16]) {
            if (c.tape) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    *c.tape4(2, kb, q) = make_float4(__uint_as_float(gr[q * 4]), __uint_as_float(gr[q * 4 + 1]),
                                                     __uint_as_float(gr[q * 4 + 2]), __uint_as_float(gr[q * 4 + 3]));
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int k = kb * 32 + c.hf * 16 + i;
                const float4 w1 = lds4(rA + k * 8);
                const float a1 = tanh_tc(dot4(w1, z, rA[k * 8 + 4]));
                if ((i & 3) == 3) sched_fence();
                const float d1 = fmaf(-a1, a1, 1.f) * __uint_as_float(gr[i]);
                g0 = fmaf(w1.x, d1, g0); g1s = fmaf(w1.y, d1, g1s); g2 = fmaf(w1.z, d1, g2); g3 = fmaf(w1.w, d1, g3);
            }
        });
        X[13] = g0; X[14] = g1s; X[15] = g2; X[16] = g3;
        tc_fence_before();
    }
    TCP_MARK(c, 4);
    c.exchange(X);
    TCP_MARK(c, 5);
    if (c.tape && c.store) {
        // the adjoint evaluation at this stage state reuses the R_net sums and grad H
        if constexpr (SH::HAS_R) {
#pragma unroll
            for (int i = 0; i < SH::NSYM; ++i) c.sck[((size_t)c.ev * 16 + i) * 128 + c.row] = X[i];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) c.sck[((size_t)c.ev * 16 + 10 + i) * 128 + c.row] = X[13 + i];
    }
    Hval = X[12] + p.b3;
    if constexpr (SH::MK == MK_CANON) {
        float pd[2];
        tc_canon_pdot(p, X + 13, u, pd);
        f[0] = cq.n11 * z[2] + cq.n12 * z[3];  // qdot = M^-1 p
        f[1] = cq.n12 * z[2] + cq.n22 * z[3];
        f[2] = cq.n11 * pd[0] + cq.n12 * pd[1];  // qddot ~= M^-1 pdot
        f[3] = cq.n12 * pd[0] + cq.n22 * pd[1];
    } else {
        float S[4][4];
        tc_make_S(p, X, S);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            float s = 0.f;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                float Rab = 0.f;
#pragma unroll
                for (int k = 0; k < 4; ++k) Rab = fmaf(S[a][k], S[b][k], Rab);
                s = fmaf(p.Jm[a * 4 + b] - Rab, X[13 + b], s);
            }
            f[a] = s + p.Gv[a] * u;
        }
    }
}

// ---------------------------------------------------------------------------------------
// xbar = (df/dy)^T v, ubar = (df/du)^T v from the taped activations of the forward evaluation
// at the same stage state and the Hessian-vector product of H_net.
// Products: dz2 = W2 da1 -> acc0, dg1 = W2^T e2 -> acc1.
// ---------------------------------------------------------------------------------------
// R_net backward chain for one hidden unit: xbar += Wr1[k]^T (1 - r^2) (Wr2sym[k] . Rb)
template <class SH>
__device__ __forceinline__ void tc_rback_unit(const TcCtx<SH>& c, int k, const float (&y)[4], const float (&Rb)[12], float (&X4)[4]) {
    const float* rA = c.small() + SH::O_RA;
    const float* rB = c.small() + SH::O_RB;
    const float* rC = c.small() + SH::O_RC;
    const float4 c0 = lds4(rC + k * 12), c1 = lds4(rC + k * 12 + 4);
    const float2 c2 = *reinterpret_cast<const float2*>(rC + k * 12 + 8);
    float rb = c0.x * Rb[0];
    rb = fmaf(c0.y, Rb[1], rb); rb = fmaf(c0.z, Rb[2], rb); rb = fmaf(c0.w, Rb[3], rb);
    rb = fmaf(c1.x, Rb[4], rb); rb = fmaf(c1.y, Rb[5], rb); rb = fmaf(c1.z, Rb[6], rb); rb = fmaf(c1.w, Rb[7], rb);
    rb = fmaf(c2.x, Rb[8], rb); rb = fmaf(c2.y, Rb[9], rb);
    const float4 wr = lds4(rB + k * 4);
    const float r1 = tanh_tc(dot4(wr, y, rA[k * 8 + 7]));
    const float zb = rb * fmaf(-r1, r1, 1.f);
    X4[0] = fmaf(wr.x, zb, X4[0]); X4[1] = fmaf(wr.y, zb, X4[1]); X4[2] = fmaf(wr.z, zb, X4[2]); X4[3] = fmaf(wr.w, zb, X4[3]);
}

template <class SH>
__device__ __forceinline__ void tc_eval_vjp(TcCtx<SH>& c, const KParams& p, const float (&y)[4], float u,
                                         const float (&v)[4], float (&xbar)[4], float& ubar) {
    constexpr int NKB = SH::NKB;
    const float* rA = c.small() + SH::O_RA;
    TCP_MARK(c, 15);
#ifdef PHNN_TC_PROFILE
    c.aphase = 2;
#endif
    float z[4], w[4], G4[4], sv[4], Rb[12];
    Canon cq = {};
    float pb[2] = {0.f, 0.f}, pdb[2] = {0.f, 0.f};
    // grad H of the forward evaluation at this stage state (ld.cg: written by the partner thread)
#pragma unroll
    for (int i = 0; i < 4; ++i) G4[i] = __ldcg(c.sck + ((size_t)c.ev * 16 + 10 + i) * 128 + c.row);
    if constexpr (SH::MK == MK_CANON) {
        cq = canon_of(p, y[1]);
        z[0] = y[0]; z[1] = y[1];
        z[2] = p.ma * y[2] + cq.beta * y[3];
        z[3] = cq.beta * y[2] + p.mc * y[3];
        pb[0] = cq.n11 * v[0] + cq.n12 * v[1];
        pb[1] = cq.n12 * v[0] + cq.n22 * v[1];
        pdb[0] = cq.n11 * v[2] + cq.n12 * v[3];
        pdb[1] = cq.n12 * v[2] + cq.n22 * v[3];
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // w = (J - diag r)^T [0, 0, pdb]
            float s = 0.f;
#pragma unroll
            for (int r = 2; r < 4; ++r) s = fmaf(p.Jm[r * 4 + k] - (r == k ? p.rdiag[r] : 0.f), pdb[r - 2], s);
            w[k] = s;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) z[i] = y[i];
    }
    if constexpr (SH::HAS_R) {
        float Sp[12], S[4][4], tg[4];
#pragma unroll
        for (int i = 0; i < SH::NSYM; ++i) Sp[i] = __ldcg(c.sck + ((size_t)c.ev * 16 + i) * 128 + c.row);
        tc_make_S(p, Sp, S);
#pragma unroll
        for (int a = 0; a < 4; ++a) sv[a] = fmaf(S[a][3], v[3], fmaf(S[a][2], v[2], fmaf(S[a][1], v[1], S[a][0] * v[0])));
#pragma unroll
        for (int a = 0; a < 4; ++a) {  // w = (J - J^T)^T v - S (S v)
            float s = 0.f;
#pragma unroll
            for (int b = 0; b < 4; ++b) s = fmaf(p.Jm[b * 4 + a], v[b], s);
#pragma unroll
            for (int b = 0; b < 4; ++b) s = fmaf(-S[a][b], sv[b], s);
            w[a] = s;
        }
        // cotangent of S: -(v t^T + g s^T) symmetrised, packed with multiplicity 2 off the diagonal
#pragma unroll
        for (int a = 0; a < 4; ++a) tg[a] = fmaf(S[a][3], G4[3], fmaf(S[a][2], G4[2], fmaf(S[a][1], G4[1], S[a][0] * G4[0])));
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = a; b < 4; ++b)
                Rb[sym_idx(a, b)] = (a == b ? -0.5f : -1.0f) * (v[a] * tg[b] + G4[a] * sv[b] + v[b] * tg[a] + G4[b] * sv[a]);
        Rb[10] = 0.f; Rb[11] = 0.f;
    }
    float X4[4] = {0.f, 0.f, 0.f, 0.f};  // xbar partial over my hidden units
    // ---- A3: da1 = s1 * (W1 w) -> product 1 (dz2 = W2 da1), with the g1 half of xbar_H (sdot1 * g1) and the
    //      first half of the R_net chain in the same loop (the loop runs at the pace of the MMA) ----
    {
        float4 an[4], gn[4];
        c.tape_load(1, 0, an);
        c.tape_load(2, 0, gn);
#pragma unroll 1
        for (int kb = 0; kb < NKB; ++kb) {
            float4 ac[4], gc[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { ac[q] = an[q]; gc[q] = gn[q]; }
            if (kb + 1 < NKB) {
                c.tape_load(1, kb + 1, an);
                c.tape_load(2, kb + 1, gn);
            }
            const int slot = c.a_begin();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float a1v[4] = {ac[q].x, ac[q].y, ac[q].z, ac[q].w};
                const float g1v[4] = {gc[q].x, gc[q].y, gc[q].z, gc[q].w};
                float av[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int k = kb * 32 + c.hf * 16 + q * 4 + e;
                    const float4 w1 = lds4(rA + k * 8);
                    const float a1 = a1v[e];
                    av[e] = fmaf(-a1, a1, 1.f) * dot4(w1, w, 0.f);
                    const float t = -2.f * a1 * av[e] * g1v[e];
                    X4[0] = fmaf(w1.x, t, X4[0]); X4[1] = fmaf(w1.y, t, X4[1]); X4[2] = fmaf(w1.z, t, X4[2]); X4[3] = fmaf(w1.w, t, X4[3]);
                }
                c.a_put4(slot, q, av);
            }
            c.a_end(slot);
            if constexpr (SH::HAS_R) {
#pragma unroll 4
                for (int i = 0; i < 8; ++i) tc_rback_unit(c, kb * 32 + c.hf * 16 + i, y, Rb, X4);
            }
        }
    }
    TCP_MARK(c, 6);
#ifdef PHNN_TC_PROFILE
    c.aphase = 5;
#endif
    // ---- B3: e2 = -2 a2 da2 w3 -> product 2 (dg1 = W2^T e2); second half of the R_net chain ----
    {
        float4 an[4];
        c.tape_load(0, 0, an);
        const uint32_t tacc = c.acc_wait();
        TCP_MARK(c, 7);
        for_acc_blocks<NKB>(tacc, [&](int jb, const uint32_t (&dz)[16]) {
            float4 a2q[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) a2q[q] = an[q];
            if (jb + 1 < NKB) c.tape_load(0, jb + 1, an);
            const int slot = c.a_begin();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float a2v[4] = {a2q[q].x, a2q[q].y, a2q[q].z, a2q[q].w};
                float ev[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int j = jb * 32 + c.hf * 16 + q * 4 + e;
                    const float w3 = rA[j * 8 + 6];
                    const float da2 = fmaf(-a2v[e], a2v[e], 1.f) * __uint_as_float(dz[q * 4 + e]);
                    ev[e] = -2.f * a2v[e] * da2 * w3;
                }
                c.a_put4(slot, q, ev);
            }
            c.a_end(slot);
            if constexpr (SH::HAS_R) {
#pragma unroll 4
                for (int i = 8; i < 16; ++i) tc_rback_unit(c, jb * 32 + c.hf * 16 + i, y, Rb, X4);
            }
        });
    }
    TCP_MARK(c, 8);
    // ---- C4: the dg1 half of xbar_H ----
    {
        float4 an[4];
        c.tape_load(1, 0, an);
        const uint32_t tacc = c.acc_wait();
        TCP_MARK(c, 9);
        for_acc_blocks<NKB>(tacc, [&](int kb, const uint32_t (&dg)[16]) {
            float4 ac[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) ac[q] = an[q];
            if (kb + 1 < NKB) c.tape_load(1, kb + 1, an);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float a1v[4] = {ac[q].x, ac[q].y, ac[q].z, ac[q].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int k = kb * 32 + c.hf * 16 + q * 4 + e;
                    const float4 w1 = lds4(rA + k * 8);
                    const float t = fmaf(-a1v[e], a1v[e], 1.f) * __uint_as_float(dg[q * 4 + e]);
                    X4[0] = fmaf(w1.x, t, X4[0]); X4[1] = fmaf(w1.y, t, X4[1]); X4[2] = fmaf(w1.z, t, X4[2]); X4[3] = fmaf(w1.w, t, X4[3]);
                }
            }
        });
        tc_fence_before();
    }
    TCP_MARK(c, 10);
    c.exchange(X4);
    TCP_MARK(c, 5);
    if constexpr (SH::MK == MK_CANON) {
        // chain through z = [q, M(theta) qdot] and M^-1(theta) (src/mass_matrix.py:310-362), SURVEY.md Appendix A
        float pd[2];
        tc_canon_pdot(p, G4, u, pd);
        const float dbeta = -p.mb * cq.sth;
        const float dD = -2.f * cq.beta * dbeta;
        const float iD2 = 1.f / (cq.D * cq.D);
        const float dn11 = -p.mc * iD2 * dD;
        const float dn12 = -dbeta / cq.D + cq.beta * iD2 * dD;
        const float dn22 = -p.ma * iD2 * dD;
        float thbar = v[0] * (dn11 * z[2] + dn12 * z[3]) + v[1] * (dn12 * z[2] + dn22 * z[3]);
        thbar += v[2] * (dn11 * pd[0] + dn12 * pd[1]) + v[3] * (dn12 * pd[0] + dn22 * pd[1]);
        float zb[4] = {X4[0], X4[1], X4[2] + pb[0], X4[3] + pb[1]};
        thbar += dbeta * (zb[2] * y[3] + zb[3] * y[2]);
        xbar[0] = zb[0];
        xbar[1] = zb[1] + thbar;
        xbar[2] = p.ma * zb[2] + cq.beta * zb[3];
        xbar[3] = cq.beta * zb[2] + p.mc * zb[3];
        ubar = p.Gv[2] * pdb[0] + p.Gv[3] * pdb[1];
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) xbar[i] = X4[i];
        ubar = fmaf(p.Gv[3], v[3], fmaf(p.Gv[2], v[2], fmaf(p.Gv[1], v[1], p.Gv[0] * v[0])));
    }
}

// Work-stealing schedule of the solve: (tile, iteration) units from a global counter.  All per-instance
// state that crosses iterations (U, Adam moments, best controls, best cost) lives in global memory and is
// read with ld.cg, so any CTA can run any iteration of any tile once the previous iteration of that tile
// has been published.  next() is called by all 320 threads of the CTA (it contains CTA barriers).
struct StealSched {
    int* counter;
    int* progress;
    long long tiles;
    int iters;
    int* slot;  // shared-memory broadcast slot
    static constexpr bool kStateInWorkspace = true;
    __device__ __forceinline__ long long tile0() const { return -1; }
    __device__ __forceinline__ int grab() {
        if (threadIdx.x == 0) {
            const int n = atomicAdd(counter, 1);
            const long long total = tiles * iters;
            if (n < total) {
                const int it = (int)(n / tiles) + 1;
                const long long tile = n % tiles;
                if (it > 1) {
                    while (*reinterpret_cast<volatile int*>(progress + tile) < it - 1) __nanosleep(256);
                    __threadfence();
                }
            }
            *slot = n < total ? n : -1;
        }
        __syncthreads();
        const int n = *slot;
        __syncthreads();
        return n;
    }
    __device__ __forceinline__ bool next(Unit& u) {
        const int n = grab();
        if (n < 0) return false;
        u.it = (int)(n / tiles) + 1;
        u.tile = n % tiles;
        return true;
    }
    template <class ENG>
    __device__ __forceinline__ void done(ENG&, const Unit& u) {
        group_bar(6, 256);  // all element threads have issued their global writes of this unit
        if (threadIdx.x == 0) {
            __threadfence();
            *reinterpret_cast<volatile int*>(progress + u.tile) = u.it;
        }
    }
};

// Static schedule of the tcgen05 kernel: CTA b runs all iterations of tiles b, b + grid, b + 2 grid, ...
// (jobs with an adjoint run on at most one CTA per SM because the tape is a per-CTA region)
struct StridedSched {
    long long first, tile, tiles;
    int stride, n_outer, it;
    static constexpr bool kStateInWorkspace = false;
    __device__ __forceinline__ long long tile0() const { return first; }
    __device__ __forceinline__ bool next(Unit& u) {
        if (it >= n_outer) {
            tile += stride;
            it = 0;
        }
        if (n_outer <= 0 || tile >= tiles) return false;
        u.tile = tile;
        u.it = ++it;
        return true;
    }
    template <class ENG>
    __device__ __forceinline__ void done(ENG& c, const Unit&) { c.gbar(); }
};

// ---------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------
template <int MK, int NS, int HID>
__global__ void __launch_bounds__(320, 1) phnn_tc_kernel(const __grid_constant__ KParams p) {
    using SH = TcShape<MK, NS, HID>;
    uint64_t* bars = reinterpret_cast<uint64_t*>(phnn_smem);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(phnn_smem + 512);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
#ifdef PHNN_TC_ARRIVE_ALL
        mbar_init(&bars[SH::B_AFULL + 0], 256);
        mbar_init(&bars[SH::B_AFULL + 1], 256);
#else
        mbar_init(&bars[SH::B_AFULL + 0], 8);
        mbar_init(&bars[SH::B_AFULL + 1], 8);
#endif
        mbar_init(&bars[SH::B_AEMPTY + 0], 1);
        mbar_init(&bars[SH::B_AEMPTY + 1], 1);
        for (int e = 0; e < SH::NBE; ++e) {
            mbar_init(&bars[SH::B_BFULL + e], 1);
            mbar_init(&bars[SH::B_BEMPTY + e], 1);
        }
        mbar_init(&bars[SH::B_ACC + 0], 1);
        mbar_init(&bars[SH::B_ACC + 1], 1);
        mbar_init(&bars[SH::B_SMALL], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                     "r"((uint32_t)SH::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tmem_ptr;

    // evaluation schedule (same derivation as phnn_kernel)
    const int E = p.T * p.S;
    int n_outer = 1, nfwd = 0, nadj = 0;
    switch (p.mode) {
        case MODE_FORWARD: nfwd = 1; break;
        case MODE_VJP: nadj = 1; break;
        case MODE_ROLLOUT: nfwd = E + (p.energy_mode == 2 ? 1 : 0); break;
        case MODE_COSTGRAD: nfwd = E; nadj = p.want_grad ? E : 0; break;
        default: n_outer = p.iters; nfwd = E; nadj = E; break;
    }
    const bool steal = (p.mode == MODE_SOLVE) && p.sched != nullptr && p.iters > 0;
    // products per schedule unit (two per evaluation, forward or adjoint): all iterations of a tile (static
    // schedule) or one solve iteration (work stealing)
    const long long nprod = (steal ? 1LL : (long long)n_outer) * (2LL * nfwd + 2LL * nadj);
    const long long per_iter = 2LL * nfwd + 2LL * nadj;
    // tiles of this CTA under the static schedule
    const long long my_tiles = p.tiles > (long long)blockIdx.x ? (p.tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const size_t tape_eval = (size_t)3 * HID * 128;  // floats per taped evaluation
    float* const tape = (p.tape && nadj > 0) ? p.tape + (size_t)blockIdx.x * E * tape_eval : nullptr;
    const int split = p.tc_split;
    StealSched ss{p.sched, p.sched + 1, p.tiles, p.iters, reinterpret_cast<int*>(phnn_smem + 768)};

    if (warp < 8) {
        // ===== element threads =====
        TcCtx<SH> c;
        c.row = threadIdx.x & 127;
        c.hf = threadIdx.x >> 7;
        c.lane = lane;
        c.barid = 1 + (warp & 3);
        c.tlane = tbase + ((uint32_t)((warp & 3) * 32) << 16);
        c.ablk = 0;
        c.qdone = 0;
        c.split = split;
        c.store = (c.hf == 0);
        c.tape = tape;
        c.sck = nullptr;
        c.ev = 0;
        mbar_wait(&bars[SH::B_SMALL], 0);
#ifdef PHNN_TC_PROFILE
        for (int i = 0; i < 16; ++i) c.prof[i] = 0;
        for (int i = 0; i < 8; ++i) c.await[i] = 0;
        for (int i = 0; i < 4; ++i) c.sub[i] = 0;
        c.aphase = 0;
        c.tlast = clock64();
#endif
        if (steal) {
            run_job(c, p, ss, c.row);
        } else {
            StridedSched sched{(long long)blockIdx.x, (long long)blockIdx.x, p.tiles, (int)gridDim.x, n_outer, 0};
            run_job(c, p, sched, c.row);
        }
        tc_fence_before();
#ifdef PHNN_TC_PROFILE
        if (blockIdx.x == 0 && (threadIdx.x == 0 || threadIdx.x == 255) && p.dbg)
        {
            for (int i = 0; i < 16; ++i) p.dbg[(threadIdx.x ? 16 : 0) + i] = c.prof[i];
            if (threadIdx.x == 0) for (int i = 0; i < 8; ++i) p.dbg[32 + i] = c.await[i];
            if (threadIdx.x == 0) for (int i = 0; i < 4; ++i) p.dbg[40 + i] = c.sub[i];
        }
#endif
    } else if (warp == 8) {
        // ===== MMA issuer =====
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(HID >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a_base = smem_u32(phnn_smem + SH::OFF_A), b_base = smem_u32(phnn_smem + SH::OFF_B);
        uint32_t ablk = 0, bent = 0;
        long long qtot = 0;  // products issued so far (accumulator / operand parity continues across units)
        for (long long unit = 0;; ++unit) {
            if (steal ? ss.grab() < 0 : unit >= my_tiles) break;
            if (lane == 0) {
#pragma unroll 1
                for (long long qq = 0; qq < nprod; ++qq, ++qtot) {
                    const uint32_t acc = tbase + (uint32_t)(qtot & 1) * HID;
#pragma unroll 1
                    for (int kb = 0; kb < SH::NKB; ++kb) {
                        const uint32_t slot = ablk & 1u;
                        mbar_wait_sleep(&bars[SH::B_AFULL + slot], (ablk >> 1) & 1u, PHNN_TC_MMA_SLEEP);
                        const uint32_t a_hi = a_base + (slot * 2) * SH::A_TILE, a_lo = a_hi + SH::A_TILE;
                        uint32_t e = bent % SH::NBE;
                        mbar_wait_sleep(&bars[SH::B_BFULL + e], (bent / SH::NBE) & 1u, PHNN_TC_MMA_SLEEP);
                        tc_fence_after();
                        uint32_t b_t = b_base + e * SH::B_TILE;
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            umma_tf32(acc, umma_desc_sw128(a_hi + ks * 32), umma_desc_sw128(b_t + ks * 32), idesc, (kb | ks) ? 1u : 0u);
                        if (split == 3) {
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)
                                umma_tf32(acc, umma_desc_sw128(a_lo + ks * 32), umma_desc_sw128(b_t + ks * 32), idesc, 1u);
                        }
                        umma_commit(&bars[SH::B_BEMPTY + e]);
                        ++bent;
                        if (split == 3) {
                            e = bent % SH::NBE;
                            mbar_wait_sleep(&bars[SH::B_BFULL + e], (bent / SH::NBE) & 1u, PHNN_TC_MMA_SLEEP);
                            tc_fence_after();
                            b_t = b_base + e * SH::B_TILE;
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)
                                umma_tf32(acc, umma_desc_sw128(a_hi + ks * 32), umma_desc_sw128(b_t + ks * 32), idesc, 1u);
                            umma_commit(&bars[SH::B_BEMPTY + e]);
                            ++bent;
                        }
                        umma_commit(&bars[SH::B_AEMPTY + slot]);
                        ++ablk;
                    }
                    umma_commit(&bars[SH::B_ACC + (qtot & 1)]);
                }
            }
            __syncwarp();
        }
    } else {
        // ===== weight producer (TMA bulk copies of pre-swizzled K-blocks) =====
        if (lane == 0) {
            mbar_expect_tx(&bars[SH::B_SMALL], SH::SMALL * 4);
            bulk_g2s(phnn_smem + SH::OFF_SMALL, p.wsmall_tc, SH::SMALL * 4, &bars[SH::B_SMALL]);
        }
        uint32_t bent = 0;
        long long qtot = 0;
        for (long long unit = 0;; ++unit) {
            if (steal ? ss.grab() < 0 : unit >= my_tiles) break;
            if (lane == 0) {
#pragma unroll 1
                for (long long qq = 0; qq < nprod; ++qq, ++qtot) {
                    if (tape && !(qq & 1)) {
                        // adjoint evaluations run from the last taped evaluation down: while evaluation e is in
                        // flight, pull the tape of e - 1 (written a whole sweep ago, so in HBM) into L2
                        const long long qi = qq % per_iter - 2LL * nfwd;
                        if (qi >= 0) {
                            const long long e_next = (long long)E - 2 - (qi >> 1);
                            if (e_next >= 0) {
                                const char* src = reinterpret_cast<const char*>(tape + (size_t)e_next * tape_eval);
                                for (int off = 0; off < (int)(tape_eval * 4); off += 65536)
                                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src + off), "r"(65536) : "memory");
                            }
                        }
                    }
                    const unsigned char* src = p.wtc + (size_t)(qtot & 1) * SH::NKB * 2 * SH::B_TILE;
#pragma unroll 1
                    for (int kb = 0; kb < SH::NKB; ++kb) {
                        for (int hl = 0; hl < (split == 3 ? 2 : 1); ++hl) {
                            const uint32_t e = bent % SH::NBE;
                            mbar_wait_sleep(&bars[SH::B_BEMPTY + e], ((bent / SH::NBE) & 1u) ^ 1u, PHNN_TC_PROD_SLEEP);
#if defined(PHNN_TC_EXP_NOB)  // timing experiment (wrong results): no weight traffic at all / none for the lo tiles
                            if (PHNN_TC_EXP_NOB == 2 || hl == 1) { mbar_arrive(&bars[SH::B_BFULL + e]); ++bent; continue; }
#endif
                            mbar_expect_tx(&bars[SH::B_BFULL + e], SH::B_TILE);
                            bulk_g2s(phnn_smem + SH::OFF_B + e * SH::B_TILE, src + (size_t)(kb * 2 + hl) * SH::B_TILE, SH::B_TILE,
                                     &bars[SH::B_BFULL + e]);
                            ++bent;
                        }
                    }
                }
            }
            __syncwarp();
        }
    }
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"((uint32_t)SH::TMEM_COLS) : "memory");
    }
}

}  // namespace phnn
