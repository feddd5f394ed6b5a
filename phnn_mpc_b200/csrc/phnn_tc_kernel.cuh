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
// Roles: warps 0 .. 4 NQ - 1 element threads (NQ per instance), then the MMA issuer warp (one lane) and the
// weight producer warp.
#pragma once
#include "phnn_kernel.cuh"

namespace phnn {

// element threads per instance: each owns 32/NQ of every 32 hidden units.  4 (16 element warps, 96 registers) was
// measured 6 % slower than 2: the per-block time of the producing loops is set by the hand-off chain, not by the
// number of warps an SM sub-partition can switch between, and the per-instance work is replicated NQ times
#ifndef PHNN_TC_NQ
#define PHNN_TC_NQ 2
#endif

template <int MK_, int NS_, int HID_>
struct TcShape {
    static_assert((MK_ == MK_PHNN || MK_ == MK_CANON) && NS_ == 4, "tensor-core path: cart-pole pHNN (fixed G) and canonical pHNN, n = 4");
    static constexpr int MK = MK_, NS = NS_, HID = HID_, NN = NS * NS;
    static constexpr bool HAS_R = (MK != MK_CANON);
    static constexpr int TM = 128;         // instances per tile (UMMA M)
    static constexpr int NQ = PHNN_TC_NQ;  // element threads per instance
    static constexpr int UP = 32 / NQ;     // hidden units of a K-block per thread
    static constexpr int CQ = UP / 4;      // 4-unit chunks (float4 of the A ring / tape) of a K-block per thread
    static constexpr int PP = UP / 2;      // pairs of a K-block per thread
    static constexpr int NEW = 4 * NQ;     // element warps
    static constexpr int NKB = HID / 32;   // K-blocks of 32 tf32 (one 128-byte swizzle row)
    static constexpr int A_TILE = TM * 128;   // bytes of one A K-block (hi or lo)
    static constexpr int B_TILE = HID * 128;  // bytes of one B K-block (hi or lo)
    static constexpr int NBE = 3;             // B ring entries
    static constexpr int TMEM_COLS = (2 * HID <= 32) ? 32 : (2 * HID <= 64) ? 64 : (2 * HID <= 128) ? 128 : (2 * HID <= 256) ? 256 : 512;
    // small weights (floats), interleaved by pairs of adjacent hidden units (P = units 2P, 2P+1; every field is
    // the pair {unit 2P, unit 2P+1}) so that the element code runs on packed FP32 pairs (FFMA2):
    //   recA[P] (20): {W1[.][0]} {W1[.][1]} | {W1[.][2]} {W1[.][3]} | {b1} {br1} | {b2} {w3} | {-2 w3} {0}
    //   recB[P] (8):  {Wr1[.][0]} {Wr1[.][1]} | {Wr1[.][2]} {Wr1[.][3]}
    //   recC[P] (20): the 10 symmetrised R_net output weights (Wr2[ab][.] + Wr2[ba][.])/2, a <= b
    static constexpr int NSYM = 10, RA = 20, RB = 8, RCP = 20;
    static constexpr int O_RA = 0;
    static constexpr int O_RB = O_RA + (HID / 2) * RA;
    static constexpr int O_RC = O_RB + (HAS_R ? (HID / 2) * RB : 0);
    static constexpr int SMALL = O_RC + (HAS_R ? (HID / 2) * RCP : 0);
    static constexpr int XW = 12 + 1 + NS;    // floats per thread in the pair exchange (S partial, H, dH)
    // shared memory map (bytes)
    static constexpr int NAS = 2;                             // A ring slots (a third slot measured no faster)
    static constexpr int OFF_A = 1024;                        // NAS slots x (hi, lo)
    static constexpr int OFF_B = OFF_A + NAS * 2 * A_TILE;    // NBE entries
    static constexpr int OFF_SMALL = OFF_B + NBE * B_TILE;
    static constexpr int OFF_XCH = OFF_SMALL + SMALL * 4;
    static constexpr int SMEM_BYTES = OFF_XCH + XW * 128 * NQ * 4;
    static_assert(SMEM_BYTES <= 232448, "shared memory budget");
    // barrier indices
    static constexpr int B_AFULL = 0, B_AEMPTY = NAS, B_BFULL = 2 * NAS, B_BEMPTY = 2 * NAS + NBE, B_ACC = 2 * NAS + 2 * NBE,
                         B_SMALL = B_ACC + 2;
    static_assert((B_SMALL + 1) * 8 <= 128, "barriers live in the first 128 bytes");
    static constexpr int THREADS = 128 * NQ + 64;
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
// same for BF16 operands (K = 16 per instruction), FP32 accumulation into the same TMEM accumulator
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
#ifdef PHNN_TC_EXP_NOMMA
    return;
#endif
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
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
#if defined(PHNN_TC16_EXP_HALFLDS) || defined(PHNN_TC16_EXP_WEAKFENCE)
__device__ __forceinline__ void sched_fence() { asm volatile(""); }  // lets the duplicate loads of the experiment merge
#else
__device__ __forceinline__ void sched_fence() { asm volatile("" ::: "memory"); }
#endif
// asynchronous TMEM load of N (16 or 8) consecutive columns of this thread's lane; the registers are
// valid only after tmem_wait (which also ties them to the wait for the compiler)
__device__ __forceinline__ void tmem_ld_issue(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_issue(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}
__device__ __forceinline__ void tmem_wait(uint32_t (&r)[8]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
                 :
                 : "memory");
}
// visit the NKB 32-column blocks of an accumulator with the next block's load in flight while
// the current one is processed: body(block index, the UP values of this thread's share of the block)
template <int NKB, int UP, class F>
__device__ __forceinline__ void for_acc_blocks(uint32_t tacc, F&& body) {
    uint32_t ra[UP], rb[UP];
    tmem_ld_issue(tacc, ra);
#pragma unroll 1
    for (int b = 0; b < NKB; b += 2) {
        tmem_wait(ra);
        tmem_ld_issue(tacc + (b + 1) * 32, rb);
        body(b, ra);
        tmem_wait(rb);
        if (b + 2 < NKB) tmem_ld_issue(tacc + (b + 2) * 32, ra);
        body(b + 1, rb);
    }
}
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity, unsigned ns) {
    while (!mbar_try(bar, parity)) __nanosleep(ns);
}
__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float2 lds2(const float* p) { return *reinterpret_cast<const float2*>(p); }
// packed FP32 pairs: fma/mul/add/sub.rn.f32x2 (SASS FFMA2 / FMUL2 / FADD2; a broadcast scalar or a negated pair
// is an operand modifier, and ptxas contracts 1 - a*a into one FFMA2)
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
#define PHNN_F32X2_BINOP(name, op)                                                                                   \
    __device__ __forceinline__ float2 name(float2 a, float2 b) {                                                     \
        float2 d;                                                                                                    \
        asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t" op                   \
            ".rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"                                                    \
            : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));                                        \
        return d;                                                                                                    \
    }
PHNN_F32X2_BINOP(mul2, "mul")
PHNN_F32X2_BINOP(add2, "add")
PHNN_F32X2_BINOP(sub2, "sub")
#undef PHNN_F32X2_BINOP
__device__ __forceinline__ float2 bc2(float x) { return make_float2(x, x); }
// {lo, hi} -> one 32-bit word of two round-to-nearest BF16 (lo at the lower address)
__device__ __forceinline__ uint32_t bf16x2_of(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ float2 xy(const float4& v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 zw(const float4& v) { return make_float2(v.z, v.w); }
// tanh_tc on a pair (same operations per element)
__device__ __forceinline__ float2 tanh_tc2(float2 x) {
#ifdef PHNN_TC_TANH_LIBM
    return make_float2(tanhf(x.x), tanhf(x.y));
#else
    const float2 t = mul2(x, bc2(2.8853900817779268f));
    const float2 d = add2(make_float2(ex2_approx(t.x), ex2_approx(t.y)), bc2(1.0f));
    const float2 big = fma2(bc2(-2.0f), make_float2(rcp_approx(d.x), rcp_approx(d.y)), bc2(1.0f));
    const float2 sm = mul2(x, fma2(mul2(x, x), bc2(-0.33333334f), bc2(1.0f)));
    return make_float2(fabsf(x.x) < 0.04f ? sm.x : big.x, fabsf(x.y) < 0.04f ? sm.y : big.y);
#endif
}

#ifdef PHNN_TC_PROFILE
// per-phase cycle counters of threads 0 and 255, kept in shared memory so that the instrumented build keeps the
// register allocation (and therefore the timing) of the production build
#define TCP_MARK(c, i) do { const unsigned t_ = (unsigned)clock(); if ((c).prof) (c).prof[i] += (unsigned long long)(t_ - (c).tlast); (c).tlast = t_; } while (0)
#else
#define TCP_MARK(c, i) do { } while (0)
#endif

#ifndef PHNN_TC_PROD_SLEEP
#define PHNN_TC_PROD_SLEEP 128
#endif
#ifndef PHNN_TC_MMA_SLEEP
#define PHNN_TC_MMA_SLEEP 128
#endif
// blocks by which the R_net work trails the hand-off of the producing loops (0: same block)
#ifndef PHNN_TC_RSKEW
#define PHNN_TC_RSKEW 2
#endif
// L2 prefetch distance of the tape, in K-block steps of the adjoint loops
#ifndef PHNN_TC_PF_AHEAD
#define PHNN_TC_PF_AHEAD 3
#endif
#ifndef PHNN_TC_RFENCE
#define PHNN_TC_RFENCE 4
#endif


// Fused result exchange of the multi-GPU solve: the 256 element threads of the CTA copy the finished tile's controls
// (a contiguous block of 128 T floats) and best costs from this rank's buffers into the result buffers of all ranks --
// plain coalesced 16-byte stores to peer-mapped addresses, travelling over NVLink while the other CTAs keep computing.
__device__ __forceinline__ void tc_peer_store_tile(const KParams& p, long long tile, int tid) {
    group_bar(5, 256);  // every row of the tile has been written by its owner thread
    const long long b0 = tile * 128;
    const long long rows = (p.B - b0) < 128 ? (p.B - b0) : 128;
    const long long nfl = rows * p.T;
    const float* src = p.U + b0 * p.T;
    const long long goff = (p.peer_off + b0) * p.T;
    const bool vec = ((goff & 3) == 0) && ((((size_t)src) & 15) == 0);
    for (int r = 0; r < p.npeer; ++r) {
        float* dst = p.peerU[r] + goff;
        if (vec && ((((size_t)dst) & 15) == 0)) {
            const long long n4 = nfl >> 2;
            for (long long i = tid; i < n4; i += 256) reinterpret_cast<float4*>(dst)[i] = __ldcg(reinterpret_cast<const float4*>(src) + i);
            for (long long i = (n4 << 2) + tid; i < nfl; i += 256) dst[i] = __ldcg(src + i);
        } else {
            for (long long i = tid; i < nfl; i += 256) dst[i] = __ldcg(src + i);
        }
        if (p.peerC[r] && p.cost && tid < rows) p.peerC[r][p.peer_off + b0 + tid] = __ldcg(p.cost + b0 + tid);
    }
}

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
    int row, qt, lane, barid;  // instance (= TMEM lane), which 1/NQ of every K-block, lane, named barrier of my instance group
    uint32_t tlane;  // TMEM base address with this warp's lane quadrant
    uint32_t ablk;   // A K-blocks produced so far
    uint32_t qdone;  // products whose accumulator this thread has waited for
    int split;       // 3 = 3xTF32, 2 = TF32 + one BF16 correction product (K = 64), 1 = plain TF32
    float* sck;      // per tile: R_net sums [0,10) and grad H [10,14) of every forward evaluation, [T*S][16][128]
    int ev;          // index of the evaluation in flight (t * S + s)
    float* tape;     // per CTA: a2, a1, g1 of every forward evaluation of the unit in flight, each [NKB][8 chunks][128]
                     // float4 (a warp's 32 rows read/write 512 contiguous bytes); nullptr when no adjoint follows
    bool store;
#ifdef PHNN_TC_PROFILE
    unsigned long long* prof;  // shared-memory counters of this thread (threads 0 and 255), else nullptr
    unsigned tlast;
#endif

    __device__ __forceinline__ uint64_t* bars() const { return reinterpret_cast<uint64_t*>(phnn_smem); }
    // (the records were also tried in the kernel parameter block / constant bank, read with LDC.64: 20 % slower)
    __device__ __forceinline__ const float* small() const { return reinterpret_cast<const float*>(phnn_smem + SH::OFF_SMALL); }
    __device__ __forceinline__ float* xch() const { return reinterpret_cast<float*>(phnn_smem + SH::OFF_XCH); }
    __device__ __forceinline__ void gbar() const { group_bar(barid, 32 * SH::NQ); }

    // ---- A-operand ring (element threads are the producers) ----
    __device__ __forceinline__ int a_begin() {
        const int slot = ablk % SH::NAS;
        mbar_wait(&bars()[SH::B_AEMPTY + slot], ((ablk / SH::NAS) & 1u) ^ 1u);
        return slot;
    }
    // four consecutive hidden units (chunk q of this thread's CQ) of the current K-block
    __device__ __forceinline__ void a_put4(int slot, int q, const float (&v)[4]) const {
        const int ch = qt * SH::CQ + q;
        const int off = (row >> 3) * 1024 + (row & 7) * 128 + ((ch ^ (row & 7)) << 4);
        unsigned char* hi = phnn_smem + SH::OFF_A + (slot * 2) * SH::A_TILE + off;
        float4 h = make_float4(tf32_rn(v[0]), tf32_rn(v[1]), tf32_rn(v[2]), tf32_rn(v[3]));
        *reinterpret_cast<float4*>(hi) = h;
        if (split == 3) {
            const float2 l01 = sub2(make_float2(v[0], v[1]), xy(h)), l23 = sub2(make_float2(v[2], v[3]), zw(h));
            *reinterpret_cast<float4*>(hi + SH::A_TILE) = make_float4(l01.x, l01.y, l23.x, l23.y);
        }
    }
    // split == 2: the correction operand of units [8 q2, 8 q2 + 8) of this thread's UP: the second half of the slot
    // is one K-major SWIZZLE_128B tile of 64 BF16 per row, [a_lo (32) | a (32)], multiplied against [b | b_lo] rows
    __device__ __forceinline__ void a_put_corr8(int slot, int q2, const float (&v0)[4], const float (&v1)[4]) const {
        unsigned char* base = phnn_smem + SH::OFF_A + (slot * 2 + 1) * SH::A_TILE + (row >> 3) * 1024 + (row & 7) * 128;
        const int ch = qt * (SH::UP / 8) + q2;  // 16-byte chunk (8 BF16) of the lo part; the hi part is 4 chunks further
        float l0[4], l1[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { l0[i] = v0[i] - tf32_rn(v0[i]); l1[i] = v1[i] - tf32_rn(v1[i]); }
        *reinterpret_cast<uint4*>(base + ((ch ^ (row & 7)) << 4)) =
            make_uint4(bf16x2_of(l0[0], l0[1]), bf16x2_of(l0[2], l0[3]), bf16x2_of(l1[0], l1[1]), bf16x2_of(l1[2], l1[3]));
        *reinterpret_cast<uint4*>(base + (((ch + 4) ^ (row & 7)) << 4)) =
            make_uint4(bf16x2_of(v0[0], v0[1]), bf16x2_of(v0[2], v0[3]), bf16x2_of(v1[0], v1[1]), bf16x2_of(v1[2], v1[3]));
    }
    // all UP values of this thread for the K-block: TF32 hi tile + the correction operand of the chosen scheme
    __device__ __forceinline__ void a_put_block(int slot, const float (&v)[SH::CQ][4]) const {
        if (split == 2) {
#pragma unroll
            for (int q = 0; q < SH::CQ; ++q) {
                const int ch = qt * SH::CQ + q;
                const int off = (row >> 3) * 1024 + (row & 7) * 128 + ((ch ^ (row & 7)) << 4);
                *reinterpret_cast<float4*>(phnn_smem + SH::OFF_A + (slot * 2) * SH::A_TILE + off) =
                    make_float4(tf32_rn(v[q][0]), tf32_rn(v[q][1]), tf32_rn(v[q][2]), tf32_rn(v[q][3]));
            }
#pragma unroll
            for (int q2 = 0; q2 < SH::CQ / 2; ++q2) a_put_corr8(slot, q2, v[2 * q2], v[2 * q2 + 1]);
        } else {
#pragma unroll
            for (int q = 0; q < SH::CQ; ++q) a_put4(slot, q, v[q]);
        }
    }
    __device__ __forceinline__ void a_end(int slot) {
        tc_fence_before();                                            // earlier tcgen05.ld of this thread are ordered first
#ifndef PHNN_TC_EXP_NOFENCE  // timing experiment (unsafe): cost of the MEMBAR.ALL.CTA the proxy fence lowers to
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async proxy (UMMA)
#endif
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars()[SH::B_AFULL + slot]);
        ++ablk;
    }
    // wait for the next product's accumulator; returns its TMEM address for this thread's lane quadrant
    __device__ __forceinline__ uint32_t acc_wait() {
        const uint32_t q = qdone++;
        mbar_wait(&bars()[SH::B_ACC + (q & 1u)], (q >> 1) & 1u);
        tc_fence_after();
        return tlane + (q & 1u) * SH::HID + qt * SH::UP;
    }
    // float4 slot of (tape array which: 0 a2, 1 a1, 2 g1; K-block jb; chunk q) of evaluation ev for this thread
    static constexpr size_t TAPE_ARR4 = (size_t)SH::NKB * 8 * 128;  // float4 per array
    __device__ __forceinline__ float4* tape4(int which, int jb, int q) const {
        return reinterpret_cast<float4*>(tape) + ((size_t)ev * 3 + which) * TAPE_ARR4 + (jb * 8 + qt * SH::CQ + q) * 128 + row;
    }
    // LAST: the final use of these lines (evict-first), otherwise they are read once more soon
    template <bool LAST>
    __device__ __forceinline__ void tape_load(int which, int jb, float4 (&v)[SH::CQ]) const {
#pragma unroll
        for (int q = 0; q < SH::CQ; ++q) v[q] = LAST ? __ldcs(tape4(which, jb, q)) : __ldcg(tape4(which, jb, q));
    }
    // sum of N values over the NQ threads of an instance, in a fixed order so that the NQ copies of the
    // instance's state stay bit-identical
    template <int N>
    __device__ __forceinline__ void exchange(float (&v)[N]) {
        float* x = xch();
        constexpr int W = 128 * SH::NQ;
#pragma unroll
        for (int i = 0; i < N; ++i) x[i * W + qt * 128 + row] = v[i];
        gbar();
#pragma unroll
        for (int i = 0; i < N; ++i) {
            float a = x[i * W + row];
#pragma unroll
            for (int j = 1; j < SH::NQ; ++j) a += x[i * W + j * 128 + row];
            v[i] = a;
        }
        gbar();
    }
    float* scratch;  // per-CTA stage states + R_net sums / grad H of the unit in flight (nullptr: no adjoint follows)
    __device__ __forceinline__ float* unit_scratch() const { return scratch; }
    __device__ __forceinline__ void peer_store(const KParams& p, long long tile) const {
        static_assert(SH::NQ == 2, "peer store assumes 256 element threads");
        tc_peer_store_tile(p, tile, (int)threadIdx.x);
    }
    __device__ __forceinline__ void begin_unit(const KParams& p, long long) {
        sck = scratch ? scratch + (size_t)p.T * p.S * NS * TW : nullptr;
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

// ---- element arithmetic on pairs of adjacent hidden units (FFMA2 / FMUL2 / FADD2) -------------------------
// The small-layer records are interleaved by pairs (TcShape), so the two units of a pair share every
// instruction: W1 x + b is four FFMA2 for two units, an accumulation over hidden units keeps an (even, odd)
// pair of partial sums that is folded once at the end.

// pre-activations of the pair: b + sum_i W[.][i] x_i ; w01 = {W[.][0] pair, W[.][1] pair}, w23 likewise
__device__ __forceinline__ float2 pair_affine(const float4& w01, const float4& w23, const float (&x)[4], float2 b) {
    float2 s = fma2(xy(w01), bc2(x[0]), b);
    s = fma2(zw(w01), bc2(x[1]), s);
    s = fma2(xy(w23), bc2(x[2]), s);
    return fma2(zw(w23), bc2(x[3]), s);
}
__device__ __forceinline__ float2 pair_linear(const float4& w01, const float4& w23, const float (&x)[4]) {
    float2 s = mul2(xy(w01), bc2(x[0]));
    s = fma2(zw(w01), bc2(x[1]), s);
    s = fma2(xy(w23), bc2(x[2]), s);
    return fma2(zw(w23), bc2(x[3]), s);
}
// acc[i] += W[.][i] pair * t pair  (the transposed product W^T t over hidden units, (even, odd) partial sums)
__device__ __forceinline__ void pair_scatter(const float4& w01, const float4& w23, float2 t, float2 (&acc)[4]) {
    acc[0] = fma2(xy(w01), t, acc[0]);
    acc[1] = fma2(zw(w01), t, acc[1]);
    acc[2] = fma2(xy(w23), t, acc[2]);
    acc[3] = fma2(zw(w23), t, acc[3]);
}
__device__ __forceinline__ float2 one_minus_sq(float2 a) { return sub2(bc2(1.f), mul2(a, a)); }  // one FFMA2

// pair index of (K-block kb, this thread's i-th pair of PP)
template <class SH>
__device__ __forceinline__ int tc_pair(const TcCtx<SH>& c, int kb, int i) { return kb * 16 + c.qt * SH::PP + i; }

// R_net hidden layer and symmetrised output sums for pairs [I0, I1) of this thread's 8 in K-block kb
template <int I0, int I1, class SH>
__device__ __forceinline__ void tc_rfwd_pairs(const TcCtx<SH>& c, int kb, const float (&y)[4], float2 (&Sp2)[10]) {
    const float* rA = c.small() + SH::O_RA;
    const float* rB = c.small() + SH::O_RB;
    const float* rC = c.small() + SH::O_RC;
#pragma unroll
    for (int i = I0; i < I1; ++i) {
        const int P = tc_pair(c, kb, i);
        const float* rb = rB + P * SH::RB;
        const float2 r = tanh_tc2(pair_affine(lds4(rb), lds4(rb + 4), y, lds2(rA + P * SH::RA + 10)));
        const float* rc = rC + P * SH::RCP;
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const float4 cc = lds4(rc + 4 * j);
            Sp2[2 * j] = fma2(xy(cc), r, Sp2[2 * j]);
            Sp2[2 * j + 1] = fma2(zw(cc), r, Sp2[2 * j + 1]);
        }
        if ((i - I0) % PHNN_TC_RFENCE == PHNN_TC_RFENCE - 1) sched_fence();
    }
}

// phase A of the forward evaluation: a1 = tanh(W1 y + b1) -> A ring (product z2 = W2 a1) and the tape,
// and part of the R_net hidden layer with its symmetrised output sums.  The R_net pairs of a K-block run after
// the block's hand-off, so the MMA starts as early as possible and the last block's tail is covered.
template <class SH>
__device__ __forceinline__ void tc_phase_a1(TcCtx<SH>& c, const float (&z)[4], const float (&y)[4], float2 (&Sp2)[10]) {
    const float* rA = c.small() + SH::O_RA;
#pragma unroll 1
    for (int kb = 0; kb < SH::NKB; ++kb) {
        // the block is computed into registers before the wait for its ring slot (which frees when the MMA of
        // block kb - 2 retires): the arithmetic overlaps the wait instead of following it
        float av[SH::CQ][4];
#pragma unroll
        for (int q = 0; q < SH::CQ; ++q) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const float* r = rA + tc_pair(c, kb, q * 2 + e) * SH::RA;
                const float2 a = tanh_tc2(pair_affine(lds4(r), lds4(r + 4), z, lds2(r + 8)));
                av[q][2 * e] = a.x; av[q][2 * e + 1] = a.y;
            }
            if (c.tape) *c.tape4(1, kb, q) = make_float4(av[q][0], av[q][1], av[q][2], av[q][3]);  // read back in phase C
        }
        const int slot = c.a_begin();
        c.a_put_block(slot, av);
        c.a_end(slot);
        // R_net pairs run PHNN_TC_RSKEW blocks behind: that many blocks (+1) of them are left after the last hand-off and
        // cover the tail of the MMA before the wait for its accumulator
        if constexpr (SH::HAS_R) {
            if (kb >= PHNN_TC_RSKEW) tc_rfwd_pairs<0, SH::PP / 2>(c, kb - PHNN_TC_RSKEW, y, Sp2);
        }
    }
    if constexpr (SH::HAS_R) {
#pragma unroll
        for (int kb = SH::NKB - PHNN_TC_RSKEW; kb < SH::NKB; ++kb) tc_rfwd_pairs<0, SH::PP / 2>(c, kb, y, Sp2);
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
    float2 Sp2[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) Sp2[i] = make_float2(0.f, 0.f);
    TCP_MARK(c, 15);
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
    tc_phase_a1(c, z, y, Sp2);
    TCP_MARK(c, 0);
    // ---- phase B: a2, H, delta2 -> product 2 (g1 = W2^T delta2) ----
    {
        const uint32_t tacc = c.acc_wait();
        TCP_MARK(c, 1);
        float2 Hp2 = make_float2(0.f, 0.f);
        for_acc_blocks<NKB, SH::UP>(tacc, [&](int jb, const uint32_t (&zr)[SH::UP]) {
            float dv[SH::CQ][4];
#pragma unroll
            for (int q = 0; q < SH::CQ; ++q) {
                float a2v[4];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float4 m = lds4(rA + tc_pair(c, jb, q * 2 + e) * SH::RA + 12);  // {b2 pair, w3 pair}
                    const float2 zz = make_float2(__uint_as_float(zr[q * 4 + 2 * e]), __uint_as_float(zr[q * 4 + 2 * e + 1]));
                    const float2 a2 = tanh_tc2(add2(zz, xy(m)));
                    Hp2 = fma2(zw(m), a2, Hp2);
                    const float2 d = mul2(one_minus_sq(a2), zw(m));
                    a2v[2 * e] = a2.x; a2v[2 * e + 1] = a2.y;
                    dv[q][2 * e] = d.x; dv[q][2 * e + 1] = d.y;
                }
                if (c.tape) __stcs(c.tape4(0, jb, q), make_float4(a2v[0], a2v[1], a2v[2], a2v[3]));
            }
            const int slot = c.a_begin();
            c.a_put_block(slot, dv);
            c.a_end(slot);
            // the rest of the R_net forward pairs rides here: this loop otherwise waits for the MMA
            if constexpr (SH::HAS_R) {
                if (jb >= PHNN_TC_RSKEW) tc_rfwd_pairs<SH::PP / 2, SH::PP>(c, jb - PHNN_TC_RSKEW, y, Sp2);
            }
        });
        if constexpr (SH::HAS_R) {
#pragma unroll
            for (int jb = NKB - PHNN_TC_RSKEW; jb < NKB; ++jb) tc_rfwd_pairs<SH::PP / 2, SH::PP>(c, jb, y, Sp2);
        }
        X[12] = Hp2.x + Hp2.y;
    }
    TCP_MARK(c, 2);
    // ---- phase C: dH = W1^T (s1 * g1) ----
    {
        float2 G2[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) G2[i] = make_float2(0.f, 0.f);
        if (c.tape) {
            // a1 comes back from the tape (written by this thread in phase A, still in L2); the first block is
            // requested before the wait for the accumulator
            float4 an[SH::CQ];
            c.tape_load<false>(1, 0, an);
            const uint32_t tacc = c.acc_wait();
            TCP_MARK(c, 3);
            for_acc_blocks<NKB, SH::UP>(tacc, [&](int kb, const uint32_t (&gr)[SH::UP]) {
                float4 ac[SH::CQ];
#pragma unroll
                for (int q = 0; q < SH::CQ; ++q) ac[q] = an[q];
                if (kb + 1 < NKB) c.tape_load<false>(1, kb + 1, an);
#pragma unroll
                for (int q = 0; q < SH::CQ; ++q) {
                    __stcs(c.tape4(2, kb, q), make_float4(__uint_as_float(gr[q * 4]), __uint_as_float(gr[q * 4 + 1]),
                                                          __uint_as_float(gr[q * 4 + 2]), __uint_as_float(gr[q * 4 + 3])));
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const float* r = rA + tc_pair(c, kb, q * 2 + e) * SH::RA;
                        const float2 a1 = e ? zw(ac[q]) : xy(ac[q]);
                        const float2 g = make_float2(__uint_as_float(gr[q * 4 + 2 * e]), __uint_as_float(gr[q * 4 + 2 * e + 1]));
                        pair_scatter(lds4(r), lds4(r + 4), mul2(one_minus_sq(a1), g), G2);
                    }
                }
            });
        } else {
            const uint32_t tacc = c.acc_wait();
            TCP_MARK(c, 3);
            for_acc_blocks<NKB, SH::UP>(tacc, [&](int kb, const uint32_t (&gr)[SH::UP]) {
#pragma unroll
                for (int i = 0; i < SH::PP; ++i) {
                    const float* r = rA + tc_pair(c, kb, i) * SH::RA;
                    const float4 w01 = lds4(r), w23 = lds4(r + 4);
                    const float2 a1 = tanh_tc2(pair_affine(w01, w23, z, lds2(r + 8)));
                    if ((i & 1) == 1) sched_fence();
                    const float2 g = make_float2(__uint_as_float(gr[2 * i]), __uint_as_float(gr[2 * i + 1]));
                    pair_scatter(w01, w23, mul2(one_minus_sq(a1), g), G2);
                }
            });
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) X[13 + i] = G2[i].x + G2[i].y;
        tc_fence_before();
    }
    TCP_MARK(c, 4);
#pragma unroll
    for (int i = 0; i < 10; ++i) X[i] = Sp2[i].x + Sp2[i].y;
    X[10] = 0.f; X[11] = 0.f;
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
// R_net backward chain for one pair of hidden units: xbar += Wr1[k]^T (1 - r^2) (Wr2sym[k] . Rb)
template <class SH>
__device__ __forceinline__ void tc_rback_pair(const TcCtx<SH>& c, int P, const float (&y)[4], const float (&Rb)[12], float2 (&X2)[4]) {
    const float* rA = c.small() + SH::O_RA;
    const float* rB = c.small() + SH::O_RB;
    const float* rc = c.small() + SH::O_RC + P * SH::RCP;
    float2 rb;
    {
        const float4 cc = lds4(rc);
        rb = fma2(zw(cc), bc2(Rb[1]), mul2(xy(cc), bc2(Rb[0])));
    }
#pragma unroll
    for (int j = 1; j < 5; ++j) {
        const float4 cc = lds4(rc + 4 * j);
        rb = fma2(xy(cc), bc2(Rb[2 * j]), rb);
        rb = fma2(zw(cc), bc2(Rb[2 * j + 1]), rb);
    }
    const float* rbw = rB + P * SH::RB;
    const float4 u01 = lds4(rbw), u23 = lds4(rbw + 4);
    const float2 r1 = tanh_tc2(pair_affine(u01, u23, y, lds2(rA + P * SH::RA + 10)));
    pair_scatter(u01, u23, mul2(rb, one_minus_sq(r1)), X2);
}

// pairs [I0, I1) of this thread's PP in K-block kb
template <int I0, int I1, class SH>
__device__ __forceinline__ void tc_rback_pairs(const TcCtx<SH>& c, int kb, const float (&y)[4], const float (&Rb)[12], float2 (&X2)[4]) {
#pragma unroll
    for (int i = I0; i < I1; ++i) {
        tc_rback_pair(c, tc_pair(c, kb, i), y, Rb, X2);
        if ((i - I0) % 2 == 1) sched_fence();
    }
}

template <class SH>
__device__ __forceinline__ void tc_eval_vjp(TcCtx<SH>& c, const KParams& p, const float (&y)[4], float u,
                                         const float (&v)[4], float (&xbar)[4], float& ubar) {
    constexpr int NKB = SH::NKB;
    const float* rA = c.small() + SH::O_RA;
    TCP_MARK(c, 15);
    float z[4], w[4], G4[4], sv[4], Rb[12];
    Canon cq = {};
    float pb[2] = {0.f, 0.f}, pdb[2] = {0.f, 0.f};
    // first tape blocks of the da1 loop, requested before the per-instance algebra below
    float4 an[SH::CQ], gn[SH::CQ];
    c.tape_load<false>(1, 0, an);
    c.tape_load<true>(2, 0, gn);
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
    // xbar partials over my hidden units as (even, odd) pairs: X2 from the R_net chain and the dg1 half of
    // xbar_H, T2 the g1 half of xbar_H without its factor -2
    float2 X2[4], T2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { X2[i] = make_float2(0.f, 0.f); T2[i] = make_float2(0.f, 0.f); }
    // ---- A3: da1 = s1 * (W1 w) -> product 1 (dz2 = W2 da1), with the g1 half of xbar_H (sdot1 * g1) and part
    //      of the R_net chain in the same loop (the loop runs at the pace of the MMA) ----
    {
#pragma unroll 1
        for (int kb = 0; kb < NKB; ++kb) {
            float4 ac[SH::CQ], gc[SH::CQ];
#pragma unroll
            for (int q = 0; q < SH::CQ; ++q) { ac[q] = an[q]; gc[q] = gn[q]; }
            if (kb + 1 < NKB) {
                c.tape_load<false>(1, kb + 1, an);
                c.tape_load<true>(2, kb + 1, gn);
            }
            float av[SH::CQ][4];
#pragma unroll
            for (int q = 0; q < SH::CQ; ++q) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float* r = rA + tc_pair(c, kb, q * 2 + e) * SH::RA;
                    const float4 w01 = lds4(r), w23 = lds4(r + 4);
                    const float2 a1 = e ? zw(ac[q]) : xy(ac[q]);
                    const float2 g1 = e ? zw(gc[q]) : xy(gc[q]);
                    const float2 da = mul2(one_minus_sq(a1), pair_linear(w01, w23, w));
                    pair_scatter(w01, w23, mul2(mul2(a1, da), g1), T2);
                    av[q][2 * e] = da.x; av[q][2 * e + 1] = da.y;
                }
            }
            const int slot = c.a_begin();
            c.a_put_block(slot, av);
            c.a_end(slot);
            if constexpr (SH::HAS_R) {
                if (kb >= PHNN_TC_RSKEW) tc_rback_pairs<0, SH::PP / 2>(c, kb - PHNN_TC_RSKEW, y, Rb, X2);
            }
        }
        if constexpr (SH::HAS_R) {
#pragma unroll
            for (int kb = NKB - PHNN_TC_RSKEW; kb < NKB; ++kb) tc_rback_pairs<0, SH::PP / 2>(c, kb, y, Rb, X2);
        }
    }
    TCP_MARK(c, 6);
    // ---- B3: e2 = -2 a2 da2 w3 -> product 2 (dg1 = W2^T e2); the rest of the R_net chain ----
    {
        float4 an[SH::CQ];
        c.tape_load<true>(0, 0, an);
        const uint32_t tacc = c.acc_wait();
        TCP_MARK(c, 7);
        for_acc_blocks<NKB, SH::UP>(tacc, [&](int jb, const uint32_t (&dz)[SH::UP]) {
            float4 a2q[SH::CQ];
#pragma unroll
            for (int q = 0; q < SH::CQ; ++q) a2q[q] = an[q];
            if (jb + 1 < NKB) c.tape_load<true>(0, jb + 1, an);
            float ev[SH::CQ][4];
#pragma unroll
            for (int q = 0; q < SH::CQ; ++q) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float2 m2w3 = lds2(rA + tc_pair(c, jb, q * 2 + e) * SH::RA + 16);  // -2 w3 pair
                    const float2 a2 = e ? zw(a2q[q]) : xy(a2q[q]);
                    const float2 dzz = make_float2(__uint_as_float(dz[q * 4 + 2 * e]), __uint_as_float(dz[q * 4 + 2 * e + 1]));
                    const float2 ee = mul2(mul2(a2, mul2(one_minus_sq(a2), dzz)), m2w3);
                    ev[q][2 * e] = ee.x; ev[q][2 * e + 1] = ee.y;
                }
            }
            const int slot = c.a_begin();
            c.a_put_block(slot, ev);
            c.a_end(slot);
            if constexpr (SH::HAS_R) {
                if (jb >= PHNN_TC_RSKEW) tc_rback_pairs<SH::PP / 2, SH::PP>(c, jb - PHNN_TC_RSKEW, y, Rb, X2);
            }
        });
        if constexpr (SH::HAS_R) {
#pragma unroll
            for (int jb = NKB - PHNN_TC_RSKEW; jb < NKB; ++jb) tc_rback_pairs<SH::PP / 2, SH::PP>(c, jb, y, Rb, X2);
        }
    }
    TCP_MARK(c, 8);
    // ---- C4: the dg1 half of xbar_H ----
    {
        float4 an[SH::CQ];
        c.tape_load<true>(1, 0, an);
        const uint32_t tacc = c.acc_wait();
        TCP_MARK(c, 9);
        for_acc_blocks<NKB, SH::UP>(tacc, [&](int kb, const uint32_t (&dg)[SH::UP]) {
            float4 ac[SH::CQ];
#pragma unroll
            for (int q = 0; q < SH::CQ; ++q) ac[q] = an[q];
            if (kb + 1 < NKB) c.tape_load<true>(1, kb + 1, an);
#pragma unroll
            for (int q = 0; q < SH::CQ; ++q) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float* r = rA + tc_pair(c, kb, q * 2 + e) * SH::RA;
                    const float2 a1 = e ? zw(ac[q]) : xy(ac[q]);
                    const float2 dgg = make_float2(__uint_as_float(dg[q * 4 + 2 * e]), __uint_as_float(dg[q * 4 + 2 * e + 1]));
                    pair_scatter(lds4(r), lds4(r + 4), mul2(one_minus_sq(a1), dgg), X2);
                }
            }
        });
        tc_fence_before();
    }
    TCP_MARK(c, 10);
    float X4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) X4[i] = fmaf(-2.f, T2[i].x + T2[i].y, X2[i].x + X2[i].y);
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
// has been published.  next() is called by all threads of the CTA (it contains CTA barriers).
struct StealSched {
    int* counter;
    int* progress;
    long long tiles;
    int iters;
    int* slot;  // shared-memory broadcast slot
    int nelem;  // element threads of the CTA
    int nsync = 0;  // threads that take part in grab(): 0 = the whole CTA, else the first nsync threads (named barrier 7)
    static constexpr bool kStateInWorkspace = true;
    __device__ __forceinline__ long long tile0() const { return -1; }
    __device__ __forceinline__ void sync_all() const {
        if (nsync) group_bar(7, nsync); else __syncthreads();
    }
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
        sync_all();
        const int n = *slot;
        sync_all();
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
        group_bar(6, nelem);  // all element threads have issued their global writes of this unit
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
__global__ void __launch_bounds__(TcShape<MK, NS, HID>::THREADS, 1) phnn_tc_kernel(const __grid_constant__ KParams p) {
    using SH = TcShape<MK, NS, HID>;
    uint64_t* bars = reinterpret_cast<uint64_t*>(phnn_smem);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(phnn_smem + 512);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int e = 0; e < SH::NAS; ++e) {
            mbar_init(&bars[SH::B_AFULL + e], SH::NEW);
            mbar_init(&bars[SH::B_AEMPTY + e], 1);
        }
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
    if (warp == SH::NEW) {
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
    StealSched ss{p.sched, p.sched + 1, p.tiles, p.iters, reinterpret_cast<int*>(phnn_smem + 768), 128 * SH::NQ};

    if (warp < SH::NEW) {
        // ===== element threads =====
        TcCtx<SH> c;
        c.row = threadIdx.x & 127;
        c.qt = threadIdx.x >> 7;
        c.lane = lane;
        c.barid = 1 + (warp & 3);
        c.tlane = tbase + ((uint32_t)((warp & 3) * 32) << 16);
        c.ablk = 0;
        c.qdone = 0;
        c.split = split;
        c.store = (c.qt == 0);
        c.tape = tape;
        c.scratch = p.scratch ? p.scratch + (size_t)blockIdx.x * tc_scratch_floats_per_cta(NS, p.T, p.S) : nullptr;
        c.sck = nullptr;
        c.ev = 0;
        mbar_wait(&bars[SH::B_SMALL], 0);
#ifdef PHNN_TC_PROFILE
        c.prof = (threadIdx.x == 0 || threadIdx.x == 128 * SH::NQ - 1) ? reinterpret_cast<unsigned long long*>(phnn_smem + 128) + (threadIdx.x ? 16 : 0) : nullptr;
        if (c.prof)
            for (int i = 0; i < 16; ++i) c.prof[i] = 0;
        c.tlast = (unsigned)clock();
#endif
        if (steal) {
            run_job(c, p, ss, c.row);
        } else {
            StridedSched sched{(long long)blockIdx.x, (long long)blockIdx.x, p.tiles, (int)gridDim.x, n_outer, 0};
            run_job(c, p, sched, c.row);
        }
        tc_fence_before();
#ifdef PHNN_TC_PROFILE
        if (blockIdx.x == 0 && c.prof && p.dbg)
            for (int i = 0; i < 16; ++i) p.dbg[(threadIdx.x ? 16 : 0) + i] = (long long)c.prof[i];
#endif
    } else if (warp == SH::NEW) {
        // ===== MMA issuer =====
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(HID >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t idesc16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(HID >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);  // BF16 x BF16 -> F32
        const uint32_t a_base = smem_u32(phnn_smem + SH::OFF_A), b_base = smem_u32(phnn_smem + SH::OFF_B);
        uint32_t ablk = 0, bent = 0;
        long long qtot = 0;  // products issued so far (accumulator / operand parity continues across units)
#ifdef PHNN_TC_PROFILE
        long long mma_wait_a = 0, mma_wait_b = 0;  // cycles the issuer waited for operand A (element threads) / B (weights)
#endif
        for (long long unit = 0;; ++unit) {
            if (steal ? ss.grab() < 0 : unit >= my_tiles) break;
            if (lane == 0) {
#pragma unroll 1
                for (long long qq = 0; qq < nprod; ++qq, ++qtot) {
                    const uint32_t acc = tbase + (uint32_t)(qtot & 1) * HID;
#pragma unroll 1
                    for (int kb = 0; kb < SH::NKB; ++kb) {
                        const uint32_t slot = ablk % SH::NAS;
#ifdef PHNN_TC_PROFILE
                        const long long t0 = clock64();
#endif
                        mbar_wait_sleep(&bars[SH::B_AFULL + slot], (ablk / SH::NAS) & 1u, PHNN_TC_MMA_SLEEP);
                        const uint32_t a_hi = a_base + (slot * 2) * SH::A_TILE, a_lo = a_hi + SH::A_TILE;
                        uint32_t e = bent % SH::NBE;
#ifdef PHNN_TC_PROFILE
                        const long long t1 = clock64();
#endif
                        mbar_wait_sleep(&bars[SH::B_BFULL + e], (bent / SH::NBE) & 1u, PHNN_TC_MMA_SLEEP);
#ifdef PHNN_TC_PROFILE
                        mma_wait_a += t1 - t0;
                        mma_wait_b += clock64() - t1;
#endif
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
                        if (split != 1) {
                            e = bent % SH::NBE;
#ifdef PHNN_TC_PROFILE
                            const long long t2 = clock64();
#endif
                            mbar_wait_sleep(&bars[SH::B_BFULL + e], (bent / SH::NBE) & 1u, PHNN_TC_MMA_SLEEP);
#ifdef PHNN_TC_PROFILE
                            mma_wait_b += clock64() - t2;
#endif
                            tc_fence_after();
                            b_t = b_base + e * SH::B_TILE;
                            if (split == 3) {
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks)
                                    umma_tf32(acc, umma_desc_sw128(a_hi + ks * 32), umma_desc_sw128(b_t + ks * 32), idesc, 1u);
                            } else {
                                // [a_lo | a] (64 BF16 per row) x [b | b_lo]: both correction terms in one K = 64 product
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks)
                                    umma_bf16(acc, umma_desc_sw128(a_lo + ks * 32), umma_desc_sw128(b_t + ks * 32), idesc16, 1u);
                            }
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
#ifdef PHNN_TC_PROFILE
        if (blockIdx.x == 0 && lane == 0 && p.dbg) { p.dbg[32] = mma_wait_a; p.dbg[33] = mma_wait_b; }
#endif
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
                    // adjoint products of evaluation e (descending): 0 -> da1 loop (reads a1, g1), 1 -> e2 loop (reads a2)
                    const long long qi = tape ? qq % per_iter - 2LL * nfwd : -1;
                    const unsigned char* src = p.wtc + (size_t)(qtot & 1) * SH::NKB * 2 * SH::B_TILE;
#pragma unroll 1
                    for (int kb = 0; kb < SH::NKB; ++kb) {
                        if (qi >= 0) {
                            // pull the tape blocks the element threads will read PF_AHEAD K-block steps from now
                            // into L2 (they were written a whole sweep ago, so they come from HBM)
                            constexpr int PF_AHEAD = PHNN_TC_PF_AHEAD;
                            int step = (int)(qi & 1) * SH::NKB + kb + PF_AHEAD;
                            long long te = (long long)E - 1 - (qi >> 1);
                            if (step >= 2 * SH::NKB) { step -= 2 * SH::NKB; --te; }
                            if (te >= 0) {
                                const float* ev0 = tape + (size_t)te * tape_eval;
                                constexpr size_t ARR = (size_t)HID * 128, BLK = 4096;  // floats per array / per K-block
                                if (step < SH::NKB) {
                                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ev0 + ARR + step * BLK), "r"(16384) : "memory");
                                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ev0 + 2 * ARR + step * BLK), "r"(16384) : "memory");
                                } else {
                                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ev0 + (step - SH::NKB) * BLK), "r"(16384) : "memory");
                                }
                            }
                        }
                        for (int hl = 0; hl < (split != 1 ? 2 : 1); ++hl) {
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
    if (warp == SH::NEW) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"((uint32_t)SH::TMEM_COLS) : "memory");
    }
}

}  // namespace phnn
