// phnn_tc16_kernel.cuh -- second-generation tcgen05 kernel of the fused pHNN-MPC path (sm_100a), tensor_mode 4.
//
// Same jobs, same per-instance arithmetic (run_job) and same four h x h tensor products per (forward, adjoint) pair
// of evaluations as phnn_tc_kernel.cuh; what changes is how the operands reach the tensor cores, because the first
// kernel was bound by the shared-memory data pipe (ncu, profiles/r01_*: LDS 35 %, tensor-core operand reads 28 %,
// STS 15 %, tape 15 % of its wavefronts) and by the L2 -> SM weight stream, not by the tensor pipe:
//
//   * FP32-level accuracy from THREE FP16 products instead of TF32 + a double-depth BF16 correction.
//     x = hi + lo with hi = fp16(x), lo = fp16(x - hi) carries 22 mantissa bits; A B ~= Ahi Bhi + Ahi Blo + Alo Bhi,
//     accumulated in FP32 in TMEM (every dropped term is <= 2^-24 of |a||b|).  FP16 has a 5-bit exponent, so every
//     operand is scaled by an exact power of two into [~2^-2, 2^15]: static scales for the weights, a1 and delta2
//     (chosen on the host from the weight norms), one scale PER INSTANCE AND EVALUATION for the adjoint operands
//     (everything downstream of the cotangent w is linear in it; the result is scaled back exactly).  6 MMAs of
//     K = 16 per 32-unit K-block instead of 8: 3/4 of the tensor time, and the weight tile of a K-block is
//     [b_hi | b_lo] = 32 KB at h = 256 instead of 64 KB: half the L2 -> SM stream and half the bulk-copy writes.
//   * OPERAND A NEVER TOUCHES SHARED MEMORY.  [a_hi (32 fp16) | a_lo (32 fp16)] of a K-block is 32 TMEM columns --
//     exactly the 32 accumulator columns the element threads have just consumed to produce it.  The products whose
//     operand is the epilogue of the previous product (delta2 from z2, e2 from dz2) are written IN PLACE over the
//     consumed accumulator block with tcgen05.st; the products fed from the instance state (a1, da1) are written into
//     the other, idle accumulator (all 8 K-blocks fit: no ring, no empty-slot wait).  tcgen05.mma reads A from TMEM
//     (the .ts form): no STS, no fence.proxy.async, no tensor-core A reads on the shared-memory pipe.
//   * mma-style FRAGMENTS instead of thread = instance.  tcgen05.ld.16x256b hands a thread two instances (TMEM lanes
//     g, g + 8) x four pairs of adjacent hidden units of every K-block, so every small-layer record (W1, b1, Wr1, Wr2,
//     w3 ... of a pair of units) loaded from shared memory serves two instances, the four lanes of a quad read four
//     consecutive records (64 contiguous bytes) instead of one warp-uniform address, and the reductions over the
//     hidden dimension finish with two shuffles inside the quad: no shared-memory exchange, no named barriers.
//     Each instance is owned by two lanes of its quad for the per-instance algebra (run_job), as before.
//
// Roles: warps 0..7 element threads, warp 8 MMA issuer (one lane), warp 9 weight producer (cp.async.bulk ring), warps
// 10..11 idle (they complete the auxiliary warpgroup whose registers go to the element warps), and for models with an
// R_net warps 12..15: the R_net warpgroup (PHNN_TC16_RWARPS below).
#pragma once
#include <type_traits>

#include "phnn_tc_kernel.cuh"

namespace phnn {

// blocks by which the R_net work trails the hand-off of the producing loops: that many blocks of it are left after the
// last hand-off and cover the tail of the MMA before the wait for its accumulator (measured: 3 beats 2 by 1 %, 1 loses 4 %)
#ifndef PHNN_TC16_RSKEW
#define PHNN_TC16_RSKEW 3
#endif

// PHNN_TC16_RMMA = 1 (experiment, NOT the default): the R_net output layer (hidden -> 10 symmetrised sums, and its
// transpose in the adjoint) as mma.sync.m16n8k16 products on the fragments the element threads already hold (the
// tcgen05.ld.16x256b layout IS the mma.sync A / D fragment layout), three FP16 hi/lo products, instead of FFMA2 over
// per-pair records from shared memory: 10 of the 27 record fields a pair of evaluations loads per unit pair become 2
// (-26 % shared-memory load wavefronts, 160 FFMA2 per K-block and warp become 24 HMMA + 8 splits).  Parity unchanged
// (profiles/r02_rnet_mma_sync_rejected.txt), but MEASURED SLOWER: 15.84 vs 13.36 ms on the profiling slice, 679 vs 623 ms
// on cfg4.  tools/hmma_probe.cu: HMMA.16816.F32 runs at 8 cycles per instruction and scheduler, 21 cycles dependent, on an
// otherwise idle SM, but next to the running tcgen05.mma stream the 1536 HMMAs per evaluation pair and SM cost ~10 SM
// cycles each (a build without them, PHNN_TC16_EXP_NOHMMA, runs 12.76 ms: the HMMAs alone are 3.35 ms): the legacy path
// serialises across the SM while the tensor cores execute tcgen05.mma.  TMEM is full at h = 256, so there is no tcgen05
// route for these side products either; they stay on the FMA pipe.
#ifndef PHNN_TC16_RMMA
#define PHNN_TC16_RMMA 0
#endif

// PHNN_TC16_RWARPS = 1 (default): the R_net work of the full shapes runs on FOUR DEDICATED WARPS (thread = instance,
// warp-uniform record loads) instead of being interleaved into the 8 element warps.  R(x) depends only on the stage state
// (forward: 10 symmetrised sums back) and on the 10 cotangents of those sums (adjoint: 4 xbar contributions back), never on
// the accumulators, so it is a side computation the element warps request through a mailbox in shared memory at the start
// of an evaluation and collect at its end.  The element warps' dependent chains -- which bound the kernel at two warps per
// scheduler -- lose ~45 % of their instructions and every scheduler gets a third warp.
#ifndef PHNN_TC16_RWARPS
#define PHNN_TC16_RWARPS 1
#endif
#ifndef PHNN_TC16_RBATCH
#define PHNN_TC16_RBATCH 2  // pairs of hidden units an R_net thread (two instances) works on at once
#endif

template <int MK_, int NS_, int HID_, bool LOWP_ = false, bool SPARSE_ = false, bool PAIR_ = false>
struct Tc16Shape {
    // PAIR (solve jobs with an even number of tiles): the kernel runs as clusters of two CTAs on the two SMs of a TPC and
    // every tensor product is ONE tcgen05.mma.cta_group::2 of M = 256 over the two tiles of the pair: each CTA keeps its own
    // 128 instances (operand A and the accumulators in its own tensor memory, element code unchanged) but stages only HALF
    // of every weight tile (its 128 of the N = 256 rows), so the tensor-core reads of operand B on the shared-memory pipe
    // and the L2 -> SM weight stream are halved per SM.  The leader CTA issues the MMAs; the other CTA's element warps
    // arrive on the leader's operand barriers through the cluster window, its otherwise idle MMA warp relays "my half of
    // the weight tile has landed", and tcgen05.commit multicasts the completion barriers to both CTAs.
    static constexpr bool PAIR = PAIR_;
    static_assert(!PAIR_ || (!SPARSE_ && NS_ == 4), "CTA pairs: full (forward + adjoint) shapes only");
    // SPARSE (forward-only shapes, jobs of at most 64 instances per SM): a tile holds 64 instances in TMEM lanes 0..15 of
    // every quadrant, and the two warps of a quadrant split the K-blocks instead of the 16-lane halves.  A forward job is
    // a chain of evaluations bound by the instructions its tile issues (ncu: schedulers 68 % busy on the 32 SMs that
    // 4096 instances occupy with 128-row tiles); half the instances per tile put it on twice the SMs with all four
    // schedulers of each still busy (TMEM lane quadrant = warp % 4 = scheduler).  The MMA still computes 128 rows.
    static constexpr bool SPARSE = SPARSE_;
    static_assert(!SPARSE_ || (NS_ != 4 && HID_ == 64), "sparse tiles: forward-only shapes with two K-blocks");
    // LOWP (tensor_mode 5): one FP16 product per algorithmic product -- operands rounded to FP16 (11-bit significand), no
    // lo halves, a third of the tensor work; looser, stated tolerance (tests/test_gpu_parity.py).  Not the default.
    static constexpr bool LOWP = LOWP_;
    // full instantiations (forward + adjoint): cart-pole pHNN with fixed G and canonical pHNN, n = 4.  Forward-only
    // instantiations (forward evaluation, rollouts, cost without gradient): n = 2 pHNN with fixed or learned G (the
    // pendulum model of BASELINE cfg2, src/pHNN.py:86-92)
    static_assert(((MK_ == MK_PHNN || MK_ == MK_CANON) && NS_ == 4) || ((MK_ == MK_PHNN || MK_ == MK_PHNN_GNET) && NS_ == 2),
                  "tensor-core path: n = 4 pHNN (fixed G) / canonical pHNN, or forward-only n = 2 pHNN");
    static constexpr int MK = MK_, NS = NS_, HID = HID_, NN = NS * NS;
    static constexpr bool FWD_ONLY = (NS_ != 4);
    static constexpr bool HAS_GNET = (MK_ == MK_PHNN_GNET);
#ifdef PHNN_TC16_EXP_NOR  // timing experiment (wrong results): no R_net work
    static constexpr bool HAS_R = false;
#else
    static constexpr bool HAS_R = (MK != MK_CANON);
#endif
    static constexpr int TM = SPARSE_ ? 64 : 128;  // instances per tile (UMMA M is 128 either way)
    static_assert(FWD_ONLY || HID / 32 >= PHNN_TC16_RSKEW, "R_net skew exceeds the number of K-blocks");
    // element warps: (TMEM lane quadrant, 16-lane half).  (Splitting the K-blocks of the forward-only shapes over twice as
    // many warps was measured SLOWER -- 2.18 vs 1.93 ms on cfg2: that chain is bound by the instructions its tile issues,
    // not by their latency.)
    static constexpr int NEW = 8;
    static constexpr int NKB = HID / 32;      // K-blocks of 32 hidden units
    static constexpr int NP = HID / 2;        // pairs of adjacent hidden units
    static constexpr int B_TILE = HID * 128;  // bytes of the weight tile of one K-block: rows of [b_hi (32 fp16) | b_lo (32 fp16)]
    static constexpr int B_TILE_CTA = PAIR_ ? B_TILE / 2 : B_TILE;  // what one CTA stages of it
#ifndef PHNN_TC16_CPF
#define PHNN_TC16_CPF 1  // blocks by which the tape loads of the short loops (grad H, dg1 half of xbar) run ahead
#endif
#ifndef PHNN_TC16_NBE
#define PHNN_TC16_NBE 5
#endif
    static constexpr int NBE = (HID >= 256 && !PAIR_) ? PHNN_TC16_NBE : 8;  // weight ring entries
    static constexpr int TMEM_COLS = (2 * HID <= 32) ? 32 : (2 * HID <= 64) ? 64 : (2 * HID <= 128) ? 128 : (2 * HID <= 256) ? 256 : 512;
    // small weights: field-major arrays of float4, one entry per pair P of adjacent hidden units (2P, 2P+1); every
    // half of a field is the pair {unit 2P, unit 2P+1}
    //   F0 {W1[.][0]} {W1[.][1]}   F1 {W1[.][2]} {W1[.][3]}   F2 {b1} {br1}   F3 {b2} {w3 * S_delta}   F4 {-2 w3 * S_e} {0}
    //   F5 {Wr1[.][0]} {Wr1[.][1]} F6 {Wr1[.][2]} {Wr1[.][3]}
    //   F7..F11 the 10 symmetrised R_net output weights (Wr2[ab][.] + Wr2[ba][.])/2, a <= b, two per field
    // n = 2 (forward only):
    //   G0 {W1[.][0]} {W1[.][1]}   G1 {b1} {b2}   G2 {w3 * S_delta} {br1}   G3 {Wr1[.][0]} {Wr1[.][1]}
    //   G4 {Wr2s[00]} {Wr2s[01]}   G5 {Wr2s[11]} {bg1}   G6 {Wg1[.][0]} {Wg1[.][1]}   G7 {Wg2[0][.]} {Wg2[1][.]}
    static constexpr int NF = FWD_ONLY ? (HAS_GNET ? 8 : 6) : (HAS_R ? 12 : 5);
    // R_net output layer on mma.sync fragments (PHNN_TC16_RMMA, full shapes with an R_net): behind the fields, the B
    // fragments of the forward sums [kb][step][n-tile][lane] and of the backward chain [kb][group][lane], one uint4 per
    // lane each (layout: fill_tc16_small in phnn_capi.cu)
    static constexpr bool RMMA = !FWD_ONLY && HAS_R && (PHNN_TC16_RMMA != 0);
    static constexpr int RFRAG = NKB * 4 * 32 * 4;  // 32-bit words of one fragment array
    static constexpr int SMALL = NF * NP * 4 + (RMMA ? 2 * RFRAG : 0);  // floats staged in shared memory
    // shared memory map (bytes): barriers in [0, 256), TMEM address at 512, scheduler slot at 768
    static constexpr int OFF_B = 1024;
    static constexpr int OFF_SMALL = OFF_B + NBE * B_TILE_CTA;
    static constexpr int OFF_XCH = OFF_SMALL + SMALL * 4;  // sparse tiles: [evaluation parity][K-block owner][64 instances][8 floats]
    // dedicated R_net warps (see PHNN_TC16_RWARPS above); CTA pairs keep the interleaved form (their scheduler barrier
    // spans the cluster)
    static constexpr bool RW = !FWD_ONLY && HAS_R && !PAIR_ && !RMMA && (PHNN_TC16_RWARPS != 0);
    // mailbox (floats): request rows [0,4) stage state, [4,14) cotangents of the sums; response rows [14,24) and [24,34):
    // the 10 sums (forward) or the 4 xbar contributions (adjoint) over the first / second half of the hidden units; one
    // row = the tile's 128 instances; then the request tag
    static constexpr int OFF_MBOX = OFF_XCH + (SPARSE_ ? 2 * 2 * 64 * 8 * 4 : 0);
    static constexpr int MBOX_ROWS = 34;
    static constexpr int SMEM_BYTES = OFF_MBOX + (RW ? MBOX_ROWS * 128 * 4 + 128 : 0);
    static_assert(SMEM_BYTES <= 232448, "shared memory budget");
    static constexpr int B_AFULL = 0, B_BFULL = NKB, B_BEMPTY = NKB + NBE, B_ACC = NKB + 2 * NBE, B_SMALL = B_ACC + 2;
    static constexpr int B_RREQ = B_SMALL + 1, B_RRESP = B_SMALL + 2;
    static_assert((B_RRESP + 1) * 8 <= 256, "barriers live in the first 256 bytes");
    // 8 element warps (two warpgroups) + one warpgroup holding the MMA issuer, the weight producer and two idle warps:
    // setmaxnreg moves registers from the latter to the element warps (the per-thread state of two instances plus the
    // fragment buffers do not fit the 168 registers a 384-thread CTA gets at launch)
    // with the R_net warpgroup: 512 threads (128 registers each at launch), element threads 184, R_net threads 96
    static constexpr int THREADS = 32 * NEW + 128 + (RW ? 128 : 0);
    static constexpr int REGS_ELEM = RW ? 184 : 224, REGS_AUX = 48, REGS_R = 96;
    static_assert(32 * NEW * REGS_ELEM + 128 * REGS_AUX + (RW ? 128 * REGS_R : 0) <= 65536, "register file");
};

// ---- tcgen05 helpers of this kernel ---------------------------------------------------------------------
// D[tmem] (+)= A[tmem] * B[smem descriptor], FP16 operands, FP32 accumulation (K = 16 per instruction)
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
// ---- CTA-pair (cta_group::2) helpers ----------------------------------------------------------------------
// D[tmem of both CTAs] (+)= A[tmem of both CTAs, 128 rows each] * B[smem descriptor, half of the N rows in each CTA]
__device__ __forceinline__ void umma_f16_ts_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
// completion of all MMAs issued so far -> the barrier at this shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// address of the same shared-memory location in CTA `rank` of the cluster (shared::cluster window)
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
// arrive on a barrier of another CTA of the cluster.  Default semantics (release at CTA scope), as CUTLASS's
// ClusterBarrier::arrive(cta_id): what the arrival publishes is tensor-memory / async-proxy state that tcgen05.wait::st +
// tcgen05.fence::before_thread_sync (or the bulk copy's own complete_tx) have already made complete.  A .release.cluster
// arrive compiles to MEMBAR.ALL.GPU + ERRBAR + CGAERRBAR in front of every arrival, i.e. it waits for the warp's tape
// stores to reach L2 at every K-block hand-off: measured +30 % on the whole solve.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
}
// Work-stealing schedule of a CTA pair: the leader pulls (pair of tiles, iteration) units from the global counter and
// hands the unit to both CTAs; CTA r of the pair runs tile 2 k + r.  Same progress protocol per tile as StealSched.
struct PairStealSched {
    int* counter;
    int* progress;
    long long tiles;  // even
    int iters;
    int* slot;
    int nelem;
    uint32_t rank;
    static constexpr bool kStateInWorkspace = true;
    __device__ __forceinline__ long long tile0() const { return -1; }
    __device__ __forceinline__ int grab() {
        const long long npairs = tiles >> 1;
        if (rank == 0 && threadIdx.x == 0) {
            const int n = atomicAdd(counter, 1);
            const long long total = npairs * iters;
            if (n < total) {
                const int it = (int)(n / npairs) + 1;
                const long long tile = 2 * (n % npairs);
                if (it > 1) {
                    while (*reinterpret_cast<volatile int*>(progress + tile) < it - 1) __nanosleep(256);
                    while (*reinterpret_cast<volatile int*>(progress + tile + 1) < it - 1) __nanosleep(256);
                    __threadfence();
                }
            }
            const int v = n < total ? n : -1;
            *slot = v;
            asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(mapa_u32(smem_u32(slot), 1)), "r"(v) : "memory");
        }
        cluster_sync_all();
        const int n = *reinterpret_cast<volatile int*>(slot);
        cluster_sync_all();
        return n;
    }
    __device__ __forceinline__ bool next(Unit& u) {
        const int n = grab();
        if (n < 0) return false;
        const long long npairs = tiles >> 1;
        u.it = (int)(n / npairs) + 1;
        u.tile = 2 * (n % npairs) + rank;
        return true;
    }
    template <class ENG>
    __device__ __forceinline__ void done(ENG&, const Unit& u) {
        group_bar(6, nelem);
        if (threadIdx.x == 0) {
            __threadfence();
            *reinterpret_cast<volatile int*>(progress + u.tile) = u.it;
        }
    }
};
// 16 TMEM lanes x 32 columns as the mma-style fragment: r[4 k + 2 rsel + e] = (lane base + lane/4 + 8 rsel, column 8 k + 2 (lane%4) + e)
__device__ __forceinline__ void tmem_ld_frag_issue(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
// 16 TMEM lanes x 32 columns of packed pairs: q[2 k + rsel] = (lane base + lane/4 + 8 rsel, column 4 k + lane%4), k = 0..7
__device__ __forceinline__ void tmem_st_pairs(uint32_t taddr, const uint32_t (&q)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.16x128b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(q[0]), "r"(q[1]), "r"(q[2]), "r"(q[3]), "r"(q[4]), "r"(q[5]), "r"(q[6]), "r"(q[7]), "r"(q[8]), "r"(q[9]), "r"(q[10]),
        "r"(q[11]), "r"(q[12]), "r"(q[13]), "r"(q[14]), "r"(q[15])
        : "memory");
}
// mbarrier wait of the single-lane roles (MMA issuer, weight producer).  PHNN_TC16_WAIT_HINT > 0 passes a suspend-time
// hint (ns) to mbarrier.try_wait: the lane is parked by the hardware until the phase completes or the hint expires,
// instead of re-polling every ~15 cycles (ncu: SYNCS.TRYWAIT + BRA + YIELD of those two lanes were 20 % of all issued
// warp instructions, taken from the issue slots of schedulers 0 and 1 which also host four of the element warps)
#ifndef PHNN_TC16_WAIT_HINT
#define PHNN_TC16_WAIT_HINT 0
#endif
__device__ __forceinline__ void mbar_wait_aux(uint64_t* bar, uint32_t parity) {
#if PHNN_TC16_WAIT_HINT > 0
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity), "r"((uint32_t)PHNN_TC16_WAIT_HINT)
            : "memory");
    } while (!ok);
#else
    mbar_wait(bar, parity);
#endif
}
// 16 TMEM lanes x 16 columns of packed pairs (the hi half of an operand block): q[2 k + rsel], k = 0..3
__device__ __forceinline__ void tmem_st_pairs_hi(uint32_t taddr, const uint32_t (&q)[8]) {
    asm volatile("tcgen05.st.sync.aligned.16x128b.x4.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(q[0]), "r"(q[1]), "r"(q[2]),
                 "r"(q[3]), "r"(q[4]), "r"(q[5]), "r"(q[6]), "r"(q[7])
                 : "memory");
}
// parity of a K-block as a type: the bodies of the block loops are instantiated for even and odd blocks, so buffers that
// are filled one block ahead (tape prefetch) ping-pong between two register sets instead of being copied
template <int P> struct Par { static constexpr int value = P; };
// visit the NKB 32-column blocks of an accumulator as fragments, the next block's load in flight
template <int NKB, class F>
__device__ __forceinline__ void for_acc_frags(uint32_t tacc, F&& body) {
    static_assert(NKB % 2 == 0, "blocks are visited in pairs");
    uint32_t ra[16], rb[16];
    tmem_ld_frag_issue(tacc, ra);
#pragma unroll 1
    for (int b = 0; b < NKB; b += 2) {
        tmem_wait(ra);
        tmem_ld_frag_issue(tacc + (b + 1) * 32, rb);
        body(b, ra, Par<0>{});
        tmem_wait(rb);
        if (b + 2 < NKB) tmem_ld_frag_issue(tacc + (b + 2) * 32, ra);
        body(b + 1, rb, Par<1>{});
    }
}
// same, in groups of four blocks: the body gets the block index modulo 4 as a type (rotating prefetch buffers that
// run two blocks ahead)
template <int NKB, class F>
__device__ __forceinline__ void for_acc_frags4(uint32_t tacc, F&& body) {
    static_assert(NKB % 4 == 0, "blocks are visited in groups of four");
    uint32_t ra[16], rb[16];
    tmem_ld_frag_issue(tacc, ra);
#pragma unroll 1
    for (int b = 0; b < NKB; b += 4) {
        tmem_wait(ra);
        tmem_ld_frag_issue(tacc + (b + 1) * 32, rb);
        body(b, ra, Par<0>{});
        tmem_wait(rb);
        tmem_ld_frag_issue(tacc + (b + 2) * 32, ra);
        body(b + 1, rb, Par<1>{});
        tmem_wait(ra);
        tmem_ld_frag_issue(tacc + (b + 3) * 32, rb);
        body(b + 2, ra, Par<2>{});
        tmem_wait(rb);
        if (b + 4 < NKB) tmem_ld_frag_issue(tacc + (b + 4) * 32, ra);
        body(b + 3, rb, Par<3>{});
    }
}
// (already scaled) pair -> fp16x2 hi and fp16x2 lo = fp16(t - hi); element .x sits in the low half (even K index)
__device__ __forceinline__ void split_f16x2(float2 t, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(t.y), "f"(t.x));
    float hx, hy;
    asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(hx), "=f"(hy) : "r"(hi));
    const float2 l = sub2(t, make_float2(hx, hy));
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(l.y), "f"(l.x));
}
__device__ __forceinline__ float2 frag2(const uint32_t (&r)[16], int k, int rsel) {
    return make_float2(__uint_as_float(r[4 * k + 2 * rsel]), __uint_as_float(r[4 * k + 2 * rsel + 1]));
}

// tanh of a pair for this kernel.  PHNN_TC16_TANH_BRANCH = 1 keeps the x - x^3/3 branch of tanh_tc2 below |x| = 0.04;
// 0 evaluates 1 - 2 / (exp(2x) + 1) everywhere: absolute error ~1.5e-7 (the FP32 rounding of a value near 1), which is
// what the products that consume the activations see anyway; half the instructions (7 instead of 14 per pair).
#ifndef PHNN_TC16_TANH_BRANCH
#define PHNN_TC16_TANH_BRANCH 0
#endif
__device__ __forceinline__ float2 tanh16(float2 x) {
#if PHNN_TC16_TANH_BRANCH
    return tanh_tc2(x);
#else
    const float2 t = mul2(x, bc2(2.8853900817779268f));
    const float2 d = add2(make_float2(ex2_approx(t.x), ex2_approx(t.y)), bc2(1.0f));
    return fma2(bc2(-2.0f), make_float2(rcp_approx(d.x), rcp_approx(d.y)), bc2(1.0f));
#endif
}

// tanh of a pair times S (an exact power of two): S - 2 S / (exp(2x) + 1), no extra instruction
__device__ __forceinline__ float2 tanh16_scaled(float2 x, float S) {
    const float2 t = mul2(x, bc2(2.8853900817779268f));
    const float2 d = add2(make_float2(ex2_approx(t.x), ex2_approx(t.y)), bc2(1.0f));
    return fma2(bc2(-2.0f * S), make_float2(rcp_approx(d.x), rcp_approx(d.y)), bc2(S));
}
// D (16 x 8, FP32) += A (16 x 16, FP16, row) * B (16 x 8, FP16, col): the warp-level tensor-core product on register
// fragments.  Thread (g = lane / 4, t = lane % 4): a0 = A[g][2t, 2t+1], a1 = A[g+8][2t, 2t+1], a2 = A[g][2t+8, 2t+9],
// a3 = A[g+8][2t+8, 2t+9]; b0 = B[2t, 2t+1][g], b1 = B[2t+8, 2t+9][g]; d0,d1 = D[g][2t, 2t+1], d2,d3 = D[g+8][2t, 2t+1]
// -- rows g / g+8 are the two instances of this thread and (2t, 2t+1) its pair of units: the tcgen05.ld.16x256b layout.
__device__ __forceinline__ void hmma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
#ifdef PHNN_TC16_EXP_NOHMMA  // timing experiment (wrong results): everything of the mma.sync path but the HMMAs themselves
    d[0] += __uint_as_float(a[0] ^ b0); d[2] += __uint_as_float(a[1] ^ b1); d[1] += __uint_as_float(a[2]); d[3] += __uint_as_float(a[3]);
    return;
#endif
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// hi/lo products of one fragment pair: A B ~= Ahi Bhi + Ahi Blo + Alo Bhi   (b = {hi(b0), hi(b1), lo(b0), lo(b1)})
__device__ __forceinline__ void hmma3(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], const uint4& b) {
    hmma16816(d, ah, b.x, b.y);
    hmma16816(d, ah, b.z, b.w);
    hmma16816(d, al, b.x, b.y);
}

template <class SH> struct Tc16Ctx;
template <class SH>
__device__ __forceinline__ void tc16_eval_fwd2(Tc16Ctx<SH>& c, const KParams& p, const float (&y)[2], float u, float (&f)[2], float& Hval);
template <class SH>
__device__ __forceinline__ void tc16_eval_fwd2s(Tc16Ctx<SH>& c, const KParams& p, const float (&y)[2], float u, float (&f)[2], float& Hval);
template <class SH>
__device__ __forceinline__ void tc16_eval_fwd(Tc16Ctx<SH>& c, const KParams& p, const float (&y)[4], float u, float (&f)[4], float& Hval);
template <class SH>
__device__ __forceinline__ void tc16_eval_vjp(Tc16Ctx<SH>& c, const KParams& p, const float (&y)[4], float u, const float (&v)[4],
                                              float (&xbar)[4], float& ubar);

template <class SH>
struct Tc16Ctx {
    static constexpr int NS = SH::NS;
    static constexpr int TW = SH::TM;
    __device__ static int ws_extra(const KParams& p) { return tc_ws_extra(SH::HID, p.T, p.S); }
    __device__ __forceinline__ void set_eval(int e) { ev = e; }
    int row;         // the instance this lane owns for the per-instance algebra (= its TMEM lane)
    int lane, cq;    // lane; position in the quad (lanes of a quad share two instances: A = TMEM lane base + lane/4, B = A + 8)
    int srcA, srcB;  // lanes whose own instance is A / B of this quad
    int tid;         // element thread index (tape layout)
    int ksel, slot, xpar;  // sparse tiles: K-block of this warp, instance index in the tile, parity of the exchange buffer
    uint32_t tl16;   // TMEM base address with the lane offset of this warp's 16-lane block
    uint32_t afull;  // CTA pairs: address of the LEADER's operand-A barriers in the cluster window
    uint32_t qdone;  // products whose accumulator this thread has waited for
    uint32_t qfeed;  // products this thread has fed (operand A written for)
    uint32_t rq;     // requests posted to the R_net warps
    float* sck;      // per tile: R_net sums [0,10) and grad H [10,14) of every forward evaluation, [T*S][16][128]
    int ev;          // index of the evaluation in flight (t * S + s)
    float* tape;     // per CTA: a2, a1, g1 of every forward evaluation of the unit in flight: [T*S][3][NKB][4][256] float4
    bool store;

    __device__ __forceinline__ uint64_t* bars() const { return reinterpret_cast<uint64_t*>(phnn_smem); }
#ifdef PHNN_TC16_EXP_RLDS  // timing experiment (wrong results): the R_net records of every pair are those of pair 0 (one broadcast address)
    static constexpr bool kExpRLds = true;
#else
    static constexpr bool kExpRLds = false;
#endif
    __device__ __forceinline__ const float4* fields() const { return reinterpret_cast<const float4*>(phnn_smem + SH::OFF_SMALL); }
    __device__ __forceinline__ const uint4* rfrag_fwd() const {
        return reinterpret_cast<const uint4*>(phnn_smem + SH::OFF_SMALL + SH::NF * SH::NP * 16) + lane;
    }
    __device__ __forceinline__ const uint4* rfrag_bwd() const { return rfrag_fwd() + SH::RFRAG / 4; }
    // ---- mailbox of the R_net warps ----
    __device__ __forceinline__ float* mbox() const { return reinterpret_cast<float*>(phnn_smem + SH::OFF_MBOX); }
    // request for this evaluation: tag 0 = forward sums of R_net at stage state y, 1 = adjoint chain (y, cotangents Rb),
    // 2 = no more work.  One owner lane per instance writes its row; every element warp arrives once.
    __device__ __forceinline__ void r_post(int tag, const float* y, const float* Rb) {
        float* m = mbox();
        if (store && tag < 2) {
#pragma unroll
            for (int i = 0; i < 4; ++i) m[i * 128 + row] = y[i];
            if (tag == 1) {
#pragma unroll
                for (int i = 0; i < 10; ++i) m[(4 + i) * 128 + row] = Rb[i];
            }
        }
        if (tid == 0) *reinterpret_cast<volatile int*>(m + SH::MBOX_ROWS * 128) = tag;
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars()[SH::B_RREQ]);
        ++rq;
    }
    // response of the last request: row i of the lane's own instance
    __device__ __forceinline__ void r_wait() const { mbar_wait(&bars()[SH::B_RRESP], (rq - 1u) & 1u); }
    __device__ __forceinline__ float r_get(int i) const { return mbox()[(14 + i) * 128 + row] + mbox()[(24 + i) * 128 + row]; }
    __device__ __forceinline__ void gbar() const { __syncwarp(); }  // the co-owners of an instance are lanes of one quad
    __device__ __forceinline__ float own(float a, float b) const { return (cq & 2) ? b : a; }
    __device__ __forceinline__ float fromA(float v) const { return __shfl_sync(0xffffffffu, v, srcA); }
    __device__ __forceinline__ float fromB(float v) const { return __shfl_sync(0xffffffffu, v, srcB); }
    // partial sums over this lane's hidden units for instances A and B -> the total of the lane's own instance
    // (both co-owners add the same two partial sums: bit-identical copies)
    __device__ __forceinline__ float quad_own_sum(float sA, float sB) const {
        const bool ob = (cq & 2) != 0;
        float r = (ob ? sB : sA) + __shfl_xor_sync(0xffffffffu, ob ? sA : sB, 2);
        r += __shfl_xor_sync(0xffffffffu, r, 1);
        return r;
    }
    // wait for the next product's accumulator; returns its column base
    __device__ __forceinline__ uint32_t acc_wait() {
        const uint32_t q = qdone++;
        mbar_wait(&bars()[SH::B_ACC + (q & 1u)], (q >> 1) & 1u);
        tc_fence_after();
        return (q & 1u) * SH::HID;
    }
    // column base of operand A of the product being fed: always the accumulator the product does NOT write -- the
    // idle one for the first product of an evaluation, the one being consumed (in place) for the second
    __device__ __forceinline__ uint32_t feed_col() const { return ((qfeed & 1u) ^ 1u) * SH::HID; }
    // operand A of K-block kb: v[rsel][k] = the pair of units (32 kb + 8 k + 2 cq, +1) of instance rsel, times `scale`
    template <bool SCALED>
    __device__ __forceinline__ void put_block(int kb, const float2 (&v)[2][4], float scale) {
        if constexpr (SH::LOWP) {
            uint32_t q[8];
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int rsel = 0; rsel < 2; ++rsel) {
                    const float2 t = SCALED ? mul2(v[rsel][k], bc2(scale)) : v[rsel][k];
                    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(q[2 * k + rsel]) : "f"(t.y), "f"(t.x));
                }
            tmem_st_pairs_hi(tl16 + feed_col() + kb * 32, q);
        } else {
            uint32_t q[16];
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int rsel = 0; rsel < 2; ++rsel)
                    split_f16x2(SCALED ? mul2(v[rsel][k], bc2(scale)) : v[rsel][k], q[2 * k + rsel], q[2 * (4 + k) + rsel]);
            tmem_st_pairs(tl16 + feed_col() + kb * 32, q);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if constexpr (SH::PAIR) {
            if (lane == 0) mbar_arrive_cluster(afull + 8u * (uint32_t)kb);
        } else {
            if (lane == 0) mbar_arrive(&bars()[SH::B_AFULL + kb]);
        }
    }
    __device__ __forceinline__ void end_feed() { ++qfeed; }
    // float4 slot of the tape: array which (0 a2, 1 a1, 2 g1), K-block kb, slot = 2 rsel + (k >> 1)
    __device__ __forceinline__ float4* tape4(int which, int kb, int slot) const {
#ifdef PHNN_TC16_EXP_TAPEL2  // timing experiment (wrong results): every evaluation uses the tape slot of evaluation 0 (L2-resident)
        return reinterpret_cast<float4*>(tape) + (((size_t)0 * 3 + which) * (SH::NKB * 4) + kb * 4 + slot) * 256 + tid;
#endif
        return reinterpret_cast<float4*>(tape) + (((size_t)ev * 3 + which) * (SH::NKB * 4) + kb * 4 + slot) * 256 + tid;
    }
    template <bool LAST>
    __device__ __forceinline__ void tape_load(int which, int kb, float4 (&v)[4]) const {
#ifdef PHNN_TC16_EXP_NOTAPE  // timing experiment (wrong results): no tape traffic
#pragma unroll
        for (int s = 0; s < 4; ++s) v[s] = make_float4(0.1f * which, 0.2f, 0.3f, 0.01f * kb);
        return;
#endif
#pragma unroll
        for (int s = 0; s < 4; ++s) v[s] = LAST ? __ldcs(tape4(which, kb, s)) : __ldcg(tape4(which, kb, s));
    }
    float* scratch;  // per-CTA stage states + R_net sums / grad H of the unit in flight (nullptr: no adjoint follows)
    __device__ __forceinline__ float* unit_scratch() const { return scratch; }
    __device__ __forceinline__ void peer_store(const KParams& p, long long tile) const { tc_peer_store_tile(p, tile, tid); }
    __device__ __forceinline__ void begin_unit(const KParams& p, long long) {
        sck = scratch ? scratch + (size_t)p.T * p.S * NS * TW : nullptr;
    }
    __device__ __forceinline__ void eval_fwd(const KParams& p, const float (&y)[NS], float u, float (&f)[NS], float& H) {
        if constexpr (SH::SPARSE) tc16_eval_fwd2s(*this, p, y, u, f, H);
        else if constexpr (SH::FWD_ONLY) tc16_eval_fwd2(*this, p, y, u, f, H);
        else tc16_eval_fwd(*this, p, y, u, f, H);
    }
    __device__ __forceinline__ void eval_vjp(const KParams& p, const float (&y)[NS], float u, const float (&v)[NS],
                                             float (&xbar)[NS], float& ubar) {
        if constexpr (SH::FWD_ONLY) {
            // forward-only instantiation: the host never routes a job with an adjoint here
            asm volatile("trap;");
#pragma unroll
            for (int i = 0; i < NS; ++i) xbar[i] = 0.f;
            ubar = 0.f;
        } else {
            tc16_eval_vjp(*this, p, y, u, v, xbar, ubar);
        }
    }
};

// pair (k of this lane's four) of K-block kb: units 32 kb + 8 k + 2 cq, +1
template <class SH>
__device__ __forceinline__ int tc16_pair(const Tc16Ctx<SH>& c, int kb, int k) {
#ifdef PHNN_TC16_EXP_HALFLDS  // timing experiment (wrong results): pairs k and k^1 use the same records, loaded once (the
                              // shared-memory volume a thread with four instances instead of two would have)
    return kb * 16 + (k & ~1) * 4 + c.cq;
#endif
    return kb * 16 + k * 4 + c.cq;
}
// the pair of units k of a tape float4 slot pair (slots 2 rsel, 2 rsel + 1 hold k = 0,1 and k = 2,3)
__device__ __forceinline__ float2 tape_pair(const float4 (&t)[4], int rsel, int k) {
    const float4& q = t[2 * rsel + (k >> 1)];
    return (k & 1) ? zw(q) : xy(q);
}
__device__ __forceinline__ float4 pack4(float2 a, float2 b) { return make_float4(a.x, a.y, b.x, b.y); }

// R_net hidden layer and symmetrised output sums for pairs [K0, K1) of this lane's four in K-block kb, both instances
template <int K0, int K1, class SH>
__device__ __forceinline__ void tc16_rfwd(const Tc16Ctx<SH>& c, int kb, const float (&yA)[4], const float (&yB)[4], float2 (&SpA)[10],
                                          float2 (&SpB)[10]) {
    const float4* F = c.fields();
#pragma unroll
    for (int k = K0; k < K1; ++k) {
        const int P = tc16_pair(c, kb, k);
        const float4 u01 = F[5 * SH::NP + P], u23 = F[6 * SH::NP + P];
        const float2 br1 = zw(F[2 * SH::NP + P]);
        const float2 rA = tanh16(pair_affine(u01, u23, yA, br1));
        const float2 rB = tanh16(pair_affine(u01, u23, yB, br1));
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const float4 cc = F[(7 + j) * SH::NP + P];
            SpA[2 * j] = fma2(xy(cc), rA, SpA[2 * j]);
            SpA[2 * j + 1] = fma2(zw(cc), rA, SpA[2 * j + 1]);
            SpB[2 * j] = fma2(xy(cc), rB, SpB[2 * j]);
            SpB[2 * j + 1] = fma2(zw(cc), rB, SpB[2 * j + 1]);
        }
        sched_fence();
    }
}

// The same on mma.sync fragments: the two groups K0, K0 + 1 of this lane's four are one K = 16 step whose operand A is
// {r(A, K0), r(B, K0), r(A, K0+1), r(B, K0+1)} as FP16 hi/lo pairs (r times 2^9) -- exactly what the thread has just
// computed; the 10 sums (padded to two n-tiles of 8) accumulate in SC: SC[4 nt + {0,1}] = sums 8 nt + 2 t + {0,1} of
// instance A, SC[4 nt + {2,3}] of instance B, complete over all hidden units (no reduction over the quad).
template <int K0, int K1, class SH>
__device__ __forceinline__ void tc16_rfwd_mma(const Tc16Ctx<SH>& c, int kb, const float (&yA)[4], const float (&yB)[4], float (&SC)[2][4]) {
    static_assert(K1 == K0 + 2 && (K0 & 1) == 0, "one K = 16 step");
    const float4* F = c.fields();
    uint32_t ah[4], al[4];
#pragma unroll
    for (int k = K0; k < K1; ++k) {
        const int P = tc16_pair(c, kb, k);
        const float4 u01 = F[5 * SH::NP + P], u23 = F[6 * SH::NP + P];
        const float2 br1 = zw(F[2 * SH::NP + P]);
        split_f16x2(tanh16_scaled(pair_affine(u01, u23, yA, br1), 512.f), ah[2 * (k - K0)], al[2 * (k - K0)]);
        split_f16x2(tanh16_scaled(pair_affine(u01, u23, yB, br1), 512.f), ah[2 * (k - K0) + 1], al[2 * (k - K0) + 1]);
    }
    const uint4* RF = c.rfrag_fwd() + (kb * 2 + (K0 >> 1)) * 2 * 32;
    hmma3(SC[0], ah, al, RF[0]);
    hmma3(SC[1], ah, al, RF[32]);
    sched_fence();
}
// the 10 sums of the lane's own instance from the accumulator fragments of its quad
template <class SH>
__device__ __forceinline__ void tc16_rsums_own(const Tc16Ctx<SH>& c, const float (&SC)[2][4], float scale, float (&Sp)[12]) {
    const int base = c.lane & ~3;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const int nt = i >> 3, src = base + ((i & 7) >> 1);
        const float a = __shfl_sync(0xffffffffu, SC[nt][i & 1], src);
        const float b = __shfl_sync(0xffffffffu, SC[nt][2 + (i & 1)], src);
        Sp[i] = c.own(a, b) * scale;
    }
}

// ---------------------------------------------------------------------------------------
// f(y,u), H(y) for the tile's 128 instances (src/pHNN.py:52-100, src/pHNN_canonical.py:172-273)
// ---------------------------------------------------------------------------------------
template <class SH>
__device__ __forceinline__ void tc16_eval_fwd(Tc16Ctx<SH>& c, const KParams& p, const float (&y)[4], float u, float (&f)[4],
                                              float& Hval) {
    constexpr int NKB = SH::NKB, NP = SH::NP;
    const float4* F = c.fields();
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
    // the two instances of this quad (for kind pHNN z = y; the canonical model has no R_net, which is what reads y)
    float zA[4], zB[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { zA[i] = c.fromA(z[i]); zB[i] = c.fromB(z[i]); }
    float2 SpA[SH::RMMA ? 1 : 10], SpB[SH::RMMA ? 1 : 10];
    float SC[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    if constexpr (!SH::RMMA) {
#pragma unroll
        for (int i = 0; i < 10; ++i) { SpA[i] = make_float2(0.f, 0.f); SpB[i] = make_float2(0.f, 0.f); }
    }
    if constexpr (SH::RW) c.r_post(0, z, nullptr);  // the R_net warps work on this stage state while the products run
    auto rfwd01 = [&](int kb) {
        if constexpr (SH::RW) return;
        else if constexpr (SH::RMMA) tc16_rfwd_mma<0, 2>(c, kb, zA, zB, SC);
        else tc16_rfwd<0, 2>(c, kb, zA, zB, SpA, SpB);
    };
    auto rfwd23 = [&](int kb) {
        if constexpr (SH::RW) return;
        else if constexpr (SH::RMMA) tc16_rfwd_mma<2, 4>(c, kb, zA, zB, SC);
        else tc16_rfwd<2, 4>(c, kb, zA, zB, SpA, SpB);
    };
    // ---- phase A: a1 = tanh(W1 z + b1) -> operand A of product 1 (z2 = W2 a1), into the idle accumulator ----
    {
#pragma unroll 1
        for (int kb = 0; kb < NKB; ++kb) {
            float2 a[2][4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int P = tc16_pair(c, kb, k);
                const float4 w01 = F[P], w23 = F[NP + P];
                const float2 b1 = xy(F[2 * NP + P]);
                a[0][k] = tanh16(pair_affine(w01, w23, zA, b1));
                a[1][k] = tanh16(pair_affine(w01, w23, zB, b1));
                if (k & 1) sched_fence();
            }
#ifndef PHNN_TC16_EXP_NOTAPE
            if (c.tape) {
#pragma unroll
                for (int s = 0; s < 4; ++s) *c.tape4(1, kb, s) = pack4(a[s >> 1][2 * (s & 1)], a[s >> 1][2 * (s & 1) + 1]);  // read back in phase C
            }
#endif
            c.template put_block<true>(kb, a, p.s16[4]);
            if constexpr (SH::HAS_R) {
                if (kb >= PHNN_TC16_RSKEW) rfwd01(kb - PHNN_TC16_RSKEW);
            }
        }
        if constexpr (SH::HAS_R) {
#pragma unroll
            for (int kb = NKB - PHNN_TC16_RSKEW; kb < NKB; ++kb) rfwd01(kb);
        }
        c.end_feed();
    }
    // ---- phase B: a2, H, delta2 -> operand A of product 2 (g1 = W2^T delta2), in place over the consumed z2 block ----
    float Hown;
    {
        const uint32_t tacc = c.acc_wait();
        const float isz = p.s16[0];
        float2 HpA = make_float2(0.f, 0.f), HpB = make_float2(0.f, 0.f);
        for_acc_frags<NKB>(c.tl16 + tacc, [&](int kb, const uint32_t (&zr)[16], auto) {
            float2 d[2][4], a2[2][4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float4 m = F[3 * NP + tc16_pair(c, kb, k)];  // {b2 pair, w3 * S_delta pair}
#pragma unroll
                for (int rsel = 0; rsel < 2; ++rsel) {
                    const float2 t = tanh16(fma2(frag2(zr, k, rsel), bc2(isz), xy(m)));
                    if (rsel) HpB = fma2(zw(m), t, HpB); else HpA = fma2(zw(m), t, HpA);
                    d[rsel][k] = mul2(one_minus_sq(t), zw(m));
                    a2[rsel][k] = t;
                }
                if (k & 1) sched_fence();
            }
#ifndef PHNN_TC16_EXP_NOTAPE
            if (c.tape) {
#pragma unroll
                for (int s = 0; s < 4; ++s) __stcs(c.tape4(0, kb, s), pack4(a2[s >> 1][2 * (s & 1)], a2[s >> 1][2 * (s & 1) + 1]));
            }
#endif
            c.template put_block<false>(kb, d, 1.0f);
            if constexpr (SH::HAS_R) {
                if (kb >= PHNN_TC16_RSKEW) rfwd23(kb - PHNN_TC16_RSKEW);
            }
        });
        if constexpr (SH::HAS_R) {
#pragma unroll
            for (int kb = NKB - PHNN_TC16_RSKEW; kb < NKB; ++kb) rfwd23(kb);
        }
        c.end_feed();
        Hown = c.quad_own_sum(HpA.x + HpA.y, HpB.x + HpB.y) * p.s16[5];
    }
    // ---- phase C: dH = W1^T (s1 * g1) ----
    float dH[4];
    {
        float2 GA[4], GB[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { GA[i] = make_float2(0.f, 0.f); GB[i] = make_float2(0.f, 0.f); }
        const float isg = p.s16[1];
        if (c.tape) {
            // a1 comes back from the tape (written by this thread in phase A, still in L2)
#if PHNN_TC16_CPF == 2
            // a1 two blocks ahead (a block of this loop is ~300 cycles, an L2 hit more than that)
            float4 tb[4][4];
            c.template tape_load<false>(1, 0, tb[0]);
            c.template tape_load<false>(1, 1, tb[1]);
            const uint32_t tacc = c.acc_wait();
            for_acc_frags4<NKB>(c.tl16 + tacc, [&](int kb, const uint32_t (&gr)[16], auto par) {
                constexpr int PI = decltype(par)::value;
                float4 (&ac)[4] = tb[PI];
                if (kb + 2 < NKB) c.template tape_load<false>(1, kb + 2, tb[(PI + 2) & 3]);
                float2 g[2][4];
#else
            float4 t0[4], t1[4];
            c.template tape_load<false>(1, 0, t0);
            const uint32_t tacc = c.acc_wait();
            for_acc_frags<NKB>(c.tl16 + tacc, [&](int kb, const uint32_t (&gr)[16], auto par) {
                float4 (&ac)[4] = decltype(par)::value ? t1 : t0;
                float4 (&an)[4] = decltype(par)::value ? t0 : t1;
                if (kb + 1 < NKB) c.template tape_load<false>(1, kb + 1, an);
                float2 g[2][4];
#endif
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int P = tc16_pair(c, kb, k);
                    const float4 w01 = F[P], w23 = F[NP + P];
#pragma unroll
                    for (int rsel = 0; rsel < 2; ++rsel) {
                        g[rsel][k] = frag2(gr, k, rsel);  // g1 * S_delta S_B: the exact power of two is applied to the totals
                        const float2 t = mul2(one_minus_sq(tape_pair(ac, rsel, k)), g[rsel][k]);
                        if (rsel) pair_scatter(w01, w23, t, GB); else pair_scatter(w01, w23, t, GA);
                    }
                }
#ifndef PHNN_TC16_EXP_NOTAPE
#pragma unroll
                for (int s = 0; s < 4; ++s) __stcs(c.tape4(2, kb, s), pack4(g[s >> 1][2 * (s & 1)], g[s >> 1][2 * (s & 1) + 1]));
#endif
            });
        } else {
            const uint32_t tacc = c.acc_wait();
            for_acc_frags<NKB>(c.tl16 + tacc, [&](int kb, const uint32_t (&gr)[16], auto) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int P = tc16_pair(c, kb, k);
                    const float4 w01 = F[P], w23 = F[NP + P];
                    const float2 b1 = xy(F[2 * NP + P]);
                    const float2 aA = tanh16(pair_affine(w01, w23, zA, b1));
                    const float2 aB = tanh16(pair_affine(w01, w23, zB, b1));
                    pair_scatter(w01, w23, mul2(one_minus_sq(aA), frag2(gr, k, 0)), GA);
                    pair_scatter(w01, w23, mul2(one_minus_sq(aB), frag2(gr, k, 1)), GB);
                    if (k & 1) sched_fence();
                }
            });
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) dH[i] = c.quad_own_sum(GA[i].x + GA[i].y, GB[i].x + GB[i].y) * isg;
        tc_fence_before();
    }
    float Sp[12];
    if constexpr (SH::RW) {
        c.r_wait();
#pragma unroll
        for (int i = 0; i < 10; ++i) Sp[i] = c.r_get(i);
    } else if constexpr (SH::RMMA) {
        tc16_rsums_own(c, SC, p.s16[7], Sp);
    } else if constexpr (SH::HAS_R) {
#pragma unroll
        for (int i = 0; i < 10; ++i) Sp[i] = c.quad_own_sum(SpA[i].x + SpA[i].y, SpB[i].x + SpB[i].y);
    }
    if (c.tape && c.store) {
        // the adjoint evaluation at this stage state reuses the R_net sums and grad H
        if constexpr (SH::HAS_R) {
#pragma unroll
            for (int i = 0; i < 10; ++i) c.sck[((size_t)c.ev * 16 + i) * 128 + c.row] = Sp[i];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) c.sck[((size_t)c.ev * 16 + 10 + i) * 128 + c.row] = dH[i];
    }
    Hval = Hown + p.b3;
    if constexpr (SH::MK == MK_CANON) {
        float pd[2];
        tc_canon_pdot(p, dH, u, pd);
        f[0] = cq.n11 * z[2] + cq.n12 * z[3];  // qdot = M^-1 p
        f[1] = cq.n12 * z[2] + cq.n22 * z[3];
        f[2] = cq.n11 * pd[0] + cq.n12 * pd[1];  // qddot ~= M^-1 pdot
        f[3] = cq.n12 * pd[0] + cq.n22 * pd[1];
    } else {
        float S[4][4];
        tc_make_S(p, Sp, S);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            float s = 0.f;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                float Rab = 0.f;
#pragma unroll
                for (int k = 0; k < 4; ++k) Rab = fmaf(S[a][k], S[b][k], Rab);
                s = fmaf(p.Jm[a * 4 + b] - Rab, dH[b], s);
            }
            f[a] = s + p.Gv[a] * u;
        }
    }
}

// ---------------------------------------------------------------------------------------
// n = 2 (forward only): pHNN with fixed or learned G (src/pHNN.py:52-100; G_net branch :86-92)
// ---------------------------------------------------------------------------------------
// tanh of a pair for the forward-only shapes: libdevice tanhf.  A forward job is a chain of evaluations of one tile
// (100 RK4 steps in cfg2) whose horizon error is what the rollout parity bound measures.  Measured on cfg2 (4096 x 100 RK4,
// distance of the trajectory from the FP64 oracle / time): tanhf 7.6e-5 / 1.93 ms; the 7-instruction tanh16 of the solve
// kernel 9.6e-5 / 1.64 ms; PHNN_TC16_FWD_TANH_POLY (odd polynomial below |x| = 0.55, 1 - 2 / (exp(2x) + 1) with MUFU.EX2
// above, 18 instructions per pair) 9.6e-5 / 1.73 ms -- the error comes from the MUFU-based branch at |x| >= 0.55, which
// tanhf evaluates with a compensated exponent.  The accurate one is the default.
__device__ __forceinline__ float2 tanh_fwd2(float2 x) {
#ifndef PHNN_TC16_FWD_TANH_POLY
    return make_float2(tanhf(x.x), tanhf(x.y));
#else
    const float2 u = mul2(x, x);
    float2 q = fma2(u, bc2(-6.615746648e-03f), bc2(2.131273902e-02f));
    q = fma2(q, u, bc2(-5.391006527e-02f));
    q = fma2(q, u, bc2(1.333311761e-01f));
    q = fma2(q, u, bc2(-3.333333204e-01f));
    const float2 sm = fma2(mul2(x, u), q, x);
    const float2 t = mul2(x, bc2(2.8853900817779268f));
    const float2 d = add2(make_float2(ex2_approx(t.x), ex2_approx(t.y)), bc2(1.0f));
    const float2 big = fma2(bc2(-2.0f), make_float2(rcp_approx(d.x), rcp_approx(d.y)), bc2(1.0f));
    return make_float2(fabsf(x.x) < 0.55f ? sm.x : big.x, fabsf(x.y) < 0.55f ? sm.y : big.y);
#endif
}
__device__ __forceinline__ float2 pair_affine2(const float4& w01, const float (&x)[2], float2 b) {
    return fma2(zw(w01), bc2(x[1]), fma2(xy(w01), bc2(x[0]), b));
}
// R_net (3 symmetrised sums 00 01 11) and G_net (2 sums) hidden layers + output partial sums for pairs [K0, K1)
template <int K0, int K1, class SH>
__device__ __forceinline__ void tc16_aux2(const Tc16Ctx<SH>& c, int kb, const float (&yA)[2], const float (&yB)[2], float2 (&SA)[5],
                                          float2 (&SB)[5]) {
    const float4* F = c.fields();
#pragma unroll
    for (int k = K0; k < K1; ++k) {
        const int P = tc16_pair(c, kb, k);
        const float4 g2 = F[2 * SH::NP + P], g3 = F[3 * SH::NP + P], g4 = F[4 * SH::NP + P], g5 = F[5 * SH::NP + P];
        const float2 rA = tanh_fwd2(pair_affine2(g3, yA, zw(g2)));
        const float2 rB = tanh_fwd2(pair_affine2(g3, yB, zw(g2)));
        SA[0] = fma2(xy(g4), rA, SA[0]); SA[1] = fma2(zw(g4), rA, SA[1]); SA[2] = fma2(xy(g5), rA, SA[2]);
        SB[0] = fma2(xy(g4), rB, SB[0]); SB[1] = fma2(zw(g4), rB, SB[1]); SB[2] = fma2(xy(g5), rB, SB[2]);
        if constexpr (SH::HAS_GNET) {
            const float4 g6 = F[6 * SH::NP + P], g7 = F[7 * SH::NP + P];
            const float2 qA = tanh_fwd2(pair_affine2(g6, yA, zw(g5)));
            const float2 qB = tanh_fwd2(pair_affine2(g6, yB, zw(g5)));
            SA[3] = fma2(xy(g7), qA, SA[3]); SA[4] = fma2(zw(g7), qA, SA[4]);
            SB[3] = fma2(xy(g7), qB, SB[3]); SB[4] = fma2(zw(g7), qB, SB[4]);
        }
        sched_fence();
    }
}
template <class SH>
__device__ __forceinline__ void tc16_eval_fwd2(Tc16Ctx<SH>& c, const KParams& p, const float (&y)[2], float u, float (&f)[2],
                                               float& Hval) {
    constexpr int NKB = SH::NKB, NP = SH::NP;
    const float4* F = c.fields();
    float zA[2], zB[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) { zA[i] = c.fromA(y[i]); zB[i] = c.fromB(y[i]); }
    float2 SA[5], SB[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) { SA[i] = make_float2(0.f, 0.f); SB[i] = make_float2(0.f, 0.f); }
    // ---- phase A: a1 = tanh(W1 z + b1) -> operand A of product 1; then the first half of the auxiliary nets, which
    //      covers the tail of the product ----
    static_assert(NKB == 2, "a1 of the whole evaluation is kept in registers between phases A and C");
    float2 a1s[NKB][2][4];
#pragma unroll
    for (int kb = 0; kb < NKB; ++kb) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int P = tc16_pair(c, kb, k);
            const float4 g0 = F[P];
            const float2 b1 = xy(F[NP + P]);
            a1s[kb][0][k] = tanh_fwd2(pair_affine2(g0, zA, b1));
            a1s[kb][1][k] = tanh_fwd2(pair_affine2(g0, zB, b1));
            if (k & 1) sched_fence();
        }
        c.template put_block<true>(kb, a1s[kb], p.s16[4]);
    }
#pragma unroll 1
    for (int kb = 0; kb < NKB; ++kb) tc16_aux2<0, 2>(c, kb, zA, zB, SA, SB);
    c.end_feed();
    // ---- phase B: a2, H, delta2 -> operand A of product 2, in place; second half of the auxiliary nets ----
    float Hown;
    {
        const uint32_t tacc = c.acc_wait();
        const float isz = p.s16[0];
        float2 HpA = make_float2(0.f, 0.f), HpB = make_float2(0.f, 0.f);
        for_acc_frags<NKB>(c.tl16 + tacc, [&](int kb, const uint32_t (&zr)[16], auto) {
            float2 d[2][4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int P = tc16_pair(c, kb, k);
                const float2 b2 = zw(F[NP + P]), w3s = xy(F[2 * NP + P]);
#pragma unroll
                for (int rsel = 0; rsel < 2; ++rsel) {
                    const float2 t = tanh_fwd2(fma2(frag2(zr, k, rsel), bc2(isz), b2));
                    if (rsel) HpB = fma2(w3s, t, HpB); else HpA = fma2(w3s, t, HpA);
                    d[rsel][k] = mul2(one_minus_sq(t), w3s);
                }
                if (k & 1) sched_fence();
            }
            c.template put_block<false>(kb, d, 1.0f);
        });
#pragma unroll 1
        for (int kb = 0; kb < NKB; ++kb) tc16_aux2<2, 4>(c, kb, zA, zB, SA, SB);
        c.end_feed();
        Hown = c.quad_own_sum(HpA.x + HpA.y, HpB.x + HpB.y) * p.s16[5];
    }
    // ---- phase C: dH = W1^T (s1 * g1), a1 still in registers (no adjoint follows, nothing is taped) ----
    float dH[2];
    {
        float2 GA[2], GB[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) { GA[i] = make_float2(0.f, 0.f); GB[i] = make_float2(0.f, 0.f); }
        const uint32_t tacc = c.acc_wait();
        for_acc_frags<NKB>(c.tl16 + tacc, [&](int kb, const uint32_t (&gr)[16], auto par) {
            constexpr int KB = decltype(par)::value;  // NKB == 2: the block index is its parity
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int P = tc16_pair(c, kb, k);
                const float4 g0 = F[P];
                const float2 tA = mul2(one_minus_sq(a1s[KB][0][k]), frag2(gr, k, 0));
                const float2 tB = mul2(one_minus_sq(a1s[KB][1][k]), frag2(gr, k, 1));
                GA[0] = fma2(xy(g0), tA, GA[0]); GA[1] = fma2(zw(g0), tA, GA[1]);
                GB[0] = fma2(xy(g0), tB, GB[0]); GB[1] = fma2(zw(g0), tB, GB[1]);
                if (k & 1) sched_fence();
            }
        });
        const float isg = p.s16[1];
#pragma unroll
        for (int i = 0; i < 2; ++i) dH[i] = c.quad_own_sum(GA[i].x + GA[i].y, GB[i].x + GB[i].y) * isg;
        tc_fence_before();
    }
    float Sp[5];
#pragma unroll
    for (int i = 0; i < (SH::HAS_GNET ? 5 : 3); ++i) Sp[i] = c.quad_own_sum(SA[i].x + SA[i].y, SB[i].x + SB[i].y);
    Hval = Hown + p.b3;
    // f = (J - J^T - S S^T) grad H + G u,  S = (R_raw + R_raw^T) / 2 (src/pHNN.py:76-100)
    float S[2][2];
    S[0][0] = Sp[0] + p.bsym[0];
    S[0][1] = S[1][0] = Sp[1] + p.bsym[1];
    S[1][1] = Sp[2] + p.bsym[2];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        float s = 0.f;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            float Rab = 0.f;
#pragma unroll
            for (int k = 0; k < 2; ++k) Rab = fmaf(S[a][k], S[b][k], Rab);
            s = fmaf(p.Jm[a * 2 + b] - Rab, dH[b], s);
        }
        const float Ga = SH::HAS_GNET ? Sp[3 + a] + p.bg2[a] : p.Gv[a];
        f[a] = s + Ga * u;
    }
}

// The same evaluation on a sparse tile: this warp owns K-block c.ksel of its quadrant's 16 instances; the sums over the
// hidden dimension are completed with the warp that owns the other K-block through shared memory (one exchange and one
// named barrier of 64 threads per evaluation; the buffers alternate, so a warp that runs ahead into the next evaluation
// does not overwrite what its partner is still reading).  Both warps finish with bit-identical totals (IEEE addition is
// commutative) and run the per-instance algebra redundantly, as the two owner lanes of an instance already do.
template <class SH>
__device__ __forceinline__ void tc16_eval_fwd2s(Tc16Ctx<SH>& c, const KParams& p, const float (&y)[2], float u, float (&f)[2],
                                                float& Hval) {
    constexpr int NP = SH::NP;
    static_assert(SH::NKB == 2, "one K-block per warp of a quadrant");
    const float4* F = c.fields();
    const int kb = c.ksel;
    float zA[2], zB[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) { zA[i] = c.fromA(y[i]); zB[i] = c.fromB(y[i]); }
    float2 SA[5], SB[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) { SA[i] = make_float2(0.f, 0.f); SB[i] = make_float2(0.f, 0.f); }
    // ---- phase A ----
    float2 a1s[2][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int P = tc16_pair(c, kb, k);
        const float4 g0 = F[P];
        const float2 b1 = xy(F[NP + P]);
        a1s[0][k] = tanh_fwd2(pair_affine2(g0, zA, b1));
        a1s[1][k] = tanh_fwd2(pair_affine2(g0, zB, b1));
        if (k & 1) sched_fence();
    }
    c.template put_block<true>(kb, a1s, p.s16[4]);
    tc16_aux2<0, 2>(c, kb, zA, zB, SA, SB);
    c.end_feed();
    // ---- phase B ----
    float2 HpA = make_float2(0.f, 0.f), HpB = make_float2(0.f, 0.f);
    {
        const uint32_t tacc = c.acc_wait();
        const float isz = p.s16[0];
        uint32_t zr[16];
        tmem_ld_frag_issue(c.tl16 + tacc + kb * 32, zr);
        tmem_wait(zr);
        float2 d[2][4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int P = tc16_pair(c, kb, k);
            const float2 b2 = zw(F[NP + P]), w3s = xy(F[2 * NP + P]);
#pragma unroll
            for (int rsel = 0; rsel < 2; ++rsel) {
                const float2 t = tanh_fwd2(fma2(frag2(zr, k, rsel), bc2(isz), b2));
                if (rsel) HpB = fma2(w3s, t, HpB); else HpA = fma2(w3s, t, HpA);
                d[rsel][k] = mul2(one_minus_sq(t), w3s);
            }
            if (k & 1) sched_fence();
        }
        c.template put_block<false>(kb, d, 1.0f);
        tc16_aux2<2, 4>(c, kb, zA, zB, SA, SB);
        c.end_feed();
    }
    // ---- phase C ----
    float2 GA[2], GB[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) { GA[i] = make_float2(0.f, 0.f); GB[i] = make_float2(0.f, 0.f); }
    {
        const uint32_t tacc = c.acc_wait();
        uint32_t gr[16];
        tmem_ld_frag_issue(c.tl16 + tacc + kb * 32, gr);
        tmem_wait(gr);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float4 g0 = F[tc16_pair(c, kb, k)];
            const float2 tA = mul2(one_minus_sq(a1s[0][k]), frag2(gr, k, 0));
            const float2 tB = mul2(one_minus_sq(a1s[1][k]), frag2(gr, k, 1));
            GA[0] = fma2(xy(g0), tA, GA[0]); GA[1] = fma2(zw(g0), tA, GA[1]);
            GB[0] = fma2(xy(g0), tB, GB[0]); GB[1] = fma2(zw(g0), tB, GB[1]);
        }
        tc_fence_before();
    }
    // ---- sums over this warp's hidden units, then over the other K-block through shared memory ----
    float v[8];
    v[0] = c.quad_own_sum(HpA.x + HpA.y, HpB.x + HpB.y);
#pragma unroll
    for (int i = 0; i < 2; ++i) v[1 + i] = c.quad_own_sum(GA[i].x + GA[i].y, GB[i].x + GB[i].y);
#pragma unroll
    for (int i = 0; i < 5; ++i) v[3 + i] = (i < (SH::HAS_GNET ? 5 : 3)) ? c.quad_own_sum(SA[i].x + SA[i].y, SB[i].x + SB[i].y) : 0.f;
    {
        float4* X = reinterpret_cast<float4*>(phnn_smem + SH::OFF_XCH);
        float4* mine = X + ((c.xpar * 2 + c.ksel) * 64 + c.slot) * 2;
        const float4* other = X + ((c.xpar * 2 + (c.ksel ^ 1)) * 64 + c.slot) * 2;
        mine[0] = make_float4(v[0], v[1], v[2], v[3]);  // both owner lanes of an instance write the same values
        mine[1] = make_float4(v[4], v[5], v[6], v[7]);
        group_bar(1 + (int)((threadIdx.x >> 5) & 3), 64);  // the two warps of this quadrant
        const float4 o0 = other[0], o1 = other[1];
        v[0] += o0.x; v[1] += o0.y; v[2] += o0.z; v[3] += o0.w;
        v[4] += o1.x; v[5] += o1.y; v[6] += o1.z; v[7] += o1.w;
        c.xpar ^= 1;
    }
    Hval = v[0] * p.s16[5] + p.b3;
    const float isg = p.s16[1];
    const float dH[2] = {v[1] * isg, v[2] * isg};
    float S[2][2];
    S[0][0] = v[3] + p.bsym[0];
    S[0][1] = S[1][0] = v[4] + p.bsym[1];
    S[1][1] = v[5] + p.bsym[2];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        float s = 0.f;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            float Rab = 0.f;
#pragma unroll
            for (int k = 0; k < 2; ++k) Rab = fmaf(S[a][k], S[b][k], Rab);
            s = fmaf(p.Jm[a * 2 + b] - Rab, dH[b], s);
        }
        const float Ga = SH::HAS_GNET ? v[6 + a] + p.bg2[a] : p.Gv[a];
        f[a] = s + Ga * u;
    }
}

// R_net backward chain for pairs [K0, K1) of K-block kb, both instances: X += Wr1[.]^T (1 - r^2) (Wr2sym[.] . Rb)
template <int K0, int K1, class SH>
__device__ __forceinline__ void tc16_rback(const Tc16Ctx<SH>& c, int kb, const float (&yA)[4], const float (&yB)[4], const float (&RbA)[10],
                                           const float (&RbB)[10], float2 (&XA)[4], float2 (&XB)[4]) {
    const float4* F = c.fields();
#pragma unroll
    for (int k = K0; k < K1; ++k) {
        const int P = tc16_pair(c, kb, k);
        float2 rbA, rbB;
        {
            const float4 cc = F[7 * SH::NP + P];
            rbA = fma2(zw(cc), bc2(RbA[1]), mul2(xy(cc), bc2(RbA[0])));
            rbB = fma2(zw(cc), bc2(RbB[1]), mul2(xy(cc), bc2(RbB[0])));
        }
#pragma unroll
        for (int j = 1; j < 5; ++j) {
            const float4 cc = F[(7 + j) * SH::NP + P];
            rbA = fma2(xy(cc), bc2(RbA[2 * j]), rbA);
            rbA = fma2(zw(cc), bc2(RbA[2 * j + 1]), rbA);
            rbB = fma2(xy(cc), bc2(RbB[2 * j]), rbB);
            rbB = fma2(zw(cc), bc2(RbB[2 * j + 1]), rbB);
        }
        const float4 u01 = F[5 * SH::NP + P], u23 = F[6 * SH::NP + P];
        const float2 br1 = zw(F[2 * SH::NP + P]);
        const float2 rA = tanh16(pair_affine(u01, u23, yA, br1));
        const float2 rB = tanh16(pair_affine(u01, u23, yB, br1));
        pair_scatter(u01, u23, mul2(rbA, one_minus_sq(rA)), XA);
        pair_scatter(u01, u23, mul2(rbB, one_minus_sq(rB)), XB);
        sched_fence();
    }
}

// The same with the output-layer transpose on mma.sync fragments: operand A = the cotangents Rb of the two instances
// (K = 10 sums padded to 16, FP16 hi/lo, scaled per instance), operand B = the weights of the 8 units of a group; the
// accumulator fragment is this thread's pair of units for its two instances, in units of 2^q S_RW (inv = the inverse).
template <int K0, int K1, class SH>
__device__ __forceinline__ void tc16_rback_mma(const Tc16Ctx<SH>& c, int kb, const float (&yA)[4], const float (&yB)[4],
                                               const uint32_t (&ah)[4], const uint32_t (&al)[4], float invA, float invB,
                                               float2 (&XA)[4], float2 (&XB)[4]) {
    const float4* F = c.fields();
    const uint4* RB = c.rfrag_bwd() + kb * 4 * 32;
#pragma unroll
    for (int k = K0; k < K1; ++k) {
        const int P = tc16_pair(c, kb, k);
        float d[4] = {0.f, 0.f, 0.f, 0.f};
        hmma3(d, ah, al, RB[k * 32]);
        const float4 u01 = F[5 * SH::NP + P], u23 = F[6 * SH::NP + P];
        const float2 br1 = zw(F[2 * SH::NP + P]);
        const float2 rA = tanh16(pair_affine(u01, u23, yA, br1));
        const float2 rB = tanh16(pair_affine(u01, u23, yB, br1));
        // inv (1 - r^2) = inv - (inv r) r
        pair_scatter(u01, u23, mul2(make_float2(d[0], d[1]), fma2(mul2(rA, bc2(-invA)), rA, bc2(invA))), XA);
        pair_scatter(u01, u23, mul2(make_float2(d[2], d[3]), fma2(mul2(rB, bc2(-invB)), rB, bc2(invB))), XB);
        sched_fence();
    }
}

// ---------------------------------------------------------------------------------------
// xbar = (df/dy)^T v, ubar = (df/du)^T v from the taped activations of the forward evaluation at the same stage state
// and the Hessian-vector product of H_net.  Products: dz2 = W2 da1, dg1 = W2^T e2.
// ---------------------------------------------------------------------------------------
template <class SH>
__device__ __forceinline__ void tc16_eval_vjp(Tc16Ctx<SH>& c, const KParams& p, const float (&y)[4], float u,
                                              const float (&v)[4], float (&xbar)[4], float& ubar) {
    constexpr int NKB = SH::NKB, NP = SH::NP;
    const float4* F = c.fields();
    float z[4], w[4], G4[4], sv[4], Rb[10];
    Canon cq = {};
    float pb[2] = {0.f, 0.f}, pdb[2] = {0.f, 0.f};
    // first tape blocks of the da1 loop, requested before the per-instance algebra below
    float4 an[4], gn[4];
    c.template tape_load<false>(1, 0, an);
    c.template tape_load<true>(2, 0, gn);
    // grad H of the forward evaluation at this stage state (ld.cg: written by the co-owner lane)
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
#pragma unroll
    for (int i = 0; i < 10; ++i) Rb[i] = 0.f;
    if constexpr (SH::HAS_R) {
        float Sp[12], S[4][4], tg[4];
#pragma unroll
        for (int i = 0; i < 10; ++i) Sp[i] = __ldcg(c.sck + ((size_t)c.ev * 16 + i) * 128 + c.row);
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
    }
    // One exact power-of-two scale per instance and evaluation brings the cotangent into the FP16 window of the
    // adjoint products (max |w'| in [2^wexp, 2^(wexp+1))): everything below is linear in (w, Rb), the result is
    // multiplied by 1 / scale at the end -- bit-identical FP32 arithmetic on the element side.
    float isc;
    {
        const float mu = fmaxf(fmaxf(fabsf(w[0]), fabsf(w[1])), fmaxf(fabsf(w[2]), fabsf(w[3])));
        int se = p.wexp16 + 254 - (int)((__float_as_uint(mu) >> 23) & 0xffu);
        se = mu > 0.f ? min(max(se, 1), 253) : 127;
        const float sc = __uint_as_float((uint32_t)se << 23);
        isc = __uint_as_float((uint32_t)(254 - se) << 23);
#pragma unroll
        for (int i = 0; i < 4; ++i) w[i] *= sc;
        // the R_net chain accumulates into the same sums as the dg1 half of xbar_H, which is kept in accumulator units
        // (dg1 * S_e S_B): bring Rb there too, so one exact scale serves the whole sum
        const float scr = sc * p.s16[6];
#pragma unroll
        for (int i = 0; i < 10; ++i) Rb[i] *= scr;
    }
    float wA[4], wB[4], yA[4], yB[4], RbA[SH::RMMA ? 1 : 10], RbB[SH::RMMA ? 1 : 10];
    uint32_t rah[4] = {0u, 0u, 0u, 0u}, ral[4] = {0u, 0u, 0u, 0u};  // Rb of the quad's two instances as mma.sync A fragments
    float rinvA = 1.f, rinvB = 1.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) { wA[i] = c.fromA(w[i]); wB[i] = c.fromB(w[i]); }
    if constexpr (SH::RW) {
        c.r_post(1, y, Rb);
    } else if constexpr (SH::HAS_R) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { yA[i] = c.fromA(y[i]); yB[i] = c.fromB(y[i]); }
        if constexpr (SH::RMMA) {
            // one more exact power of two per instance puts max |Rb| in [2^11, 2^12) (FP16 hi + lo keeps 22 bits down
            // to 2^-13 of that); the accumulator comes back times 2^q S_RW
            float mr = 0.f;
#pragma unroll
            for (int i = 0; i < 10; ++i) mr = fmaxf(mr, fabsf(Rb[i]));
            int q = 138 - (int)((__float_as_uint(mr) >> 23) & 0xffu);
            q = mr > 0.f ? min(max(q, -60), 60) : 0;
            const float rs = __uint_as_float((uint32_t)(127 + q) << 23);
            const float rinv = __uint_as_float((uint32_t)(127 - q - p.rexp16) << 23);
            uint32_t wh[5], wl[5];
#pragma unroll
            for (int tt = 0; tt < 5; ++tt) split_f16x2(make_float2(Rb[2 * tt] * rs, Rb[2 * tt + 1] * rs), wh[tt], wl[tt]);
#pragma unroll
            for (int tt = 0; tt < 5; ++tt) {
                const uint32_t xh = __shfl_sync(0xffffffffu, wh[tt], c.srcA), xl = __shfl_sync(0xffffffffu, wl[tt], c.srcA);
                const uint32_t yh = __shfl_sync(0xffffffffu, wh[tt], c.srcB), yl = __shfl_sync(0xffffffffu, wl[tt], c.srcB);
                if (tt < 4) {
                    if (c.cq == tt) { rah[0] = xh; ral[0] = xl; rah[1] = yh; ral[1] = yl; }
                } else if (c.cq == 0) {
                    rah[2] = xh; ral[2] = xl; rah[3] = yh; ral[3] = yl;
                }
            }
            rinvA = c.fromA(rinv);
            rinvB = c.fromB(rinv);
        } else {
#pragma unroll
            for (int i = 0; i < 10; ++i) { RbA[i] = c.fromA(Rb[i]); RbB[i] = c.fromB(Rb[i]); }
        }
    }
    auto rback01 = [&](int kb, float2 (&XA_)[4], float2 (&XB_)[4]) {
        if constexpr (SH::RW) return;
        else if constexpr (SH::RMMA) tc16_rback_mma<0, 2>(c, kb, yA, yB, rah, ral, rinvA, rinvB, XA_, XB_);
        else tc16_rback<0, 2>(c, kb, yA, yB, RbA, RbB, XA_, XB_);
    };
    auto rback23 = [&](int kb, float2 (&XA_)[4], float2 (&XB_)[4]) {
        if constexpr (SH::RW) return;
        else if constexpr (SH::RMMA) tc16_rback_mma<2, 4>(c, kb, yA, yB, rah, ral, rinvA, rinvB, XA_, XB_);
        else tc16_rback<2, 4>(c, kb, yA, yB, RbA, RbB, XA_, XB_);
    };
    // xbar partials over my hidden units as (even, odd) pairs: X from the R_net chain and the dg1 half of xbar_H,
    // T the g1 half of xbar_H without its factor -2
    float2 XA[4], XB[4], TA[4], TB[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { XA[i] = XB[i] = TA[i] = TB[i] = make_float2(0.f, 0.f); }
    // ---- A3: da1 = s1 * (W1 w) -> operand A of product 1 (dz2 = W2 da1), with the g1 half of xbar_H (sdot1 * g1) and
    //      half of the R_net chain in the same loop ----
    {
        float4 an1[4], gn1[4];  // second register set of the tape prefetch (an / gn hold block 0)
        auto a3_block = [&](int kb, auto par) {
            float4 (&ac)[4] = decltype(par)::value ? an1 : an;
            float4 (&gc)[4] = decltype(par)::value ? gn1 : gn;
            float4 (&ax)[4] = decltype(par)::value ? an : an1;
            float4 (&gx)[4] = decltype(par)::value ? gn : gn1;
            if (kb + 1 < NKB) {
                c.template tape_load<false>(1, kb + 1, ax);
                c.template tape_load<true>(2, kb + 1, gx);
            }
            float2 da[2][4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int P = tc16_pair(c, kb, k);
                const float4 w01 = F[P], w23 = F[NP + P];
                {
                    const float2 a1 = tape_pair(ac, 0, k);
                    da[0][k] = mul2(one_minus_sq(a1), pair_linear(w01, w23, wA));
                    pair_scatter(w01, w23, mul2(mul2(a1, da[0][k]), tape_pair(gc, 0, k)), TA);
                }
                {
                    const float2 a1 = tape_pair(ac, 1, k);
                    da[1][k] = mul2(one_minus_sq(a1), pair_linear(w01, w23, wB));
                    pair_scatter(w01, w23, mul2(mul2(a1, da[1][k]), tape_pair(gc, 1, k)), TB);
                }
            }
            c.template put_block<false>(kb, da, 1.0f);
            if constexpr (SH::HAS_R) {
                if (kb >= PHNN_TC16_RSKEW) rback01(kb - PHNN_TC16_RSKEW, XA, XB);
            }
        };
#pragma unroll 1
        for (int kb = 0; kb < NKB; kb += 2) {
            a3_block(kb, Par<0>{});
            a3_block(kb + 1, Par<1>{});
        }
        if constexpr (SH::HAS_R) {
#pragma unroll
            for (int kb = NKB - PHNN_TC16_RSKEW; kb < NKB; ++kb) rback01(kb, XA, XB);
        }
        c.end_feed();
    }
    // ---- B3: e2 = -2 a2 da2 w3 -> operand A of product 2 (dg1 = W2^T e2), in place; the rest of the R_net chain ----
    {
        float4 t0[4], t1[4];
        c.template tape_load<true>(0, 0, t0);
        const uint32_t tacc = c.acc_wait();
        for_acc_frags<NKB>(c.tl16 + tacc, [&](int kb, const uint32_t (&dz)[16], auto par) {
            float4 (&a2q)[4] = decltype(par)::value ? t1 : t0;
            float4 (&a2n)[4] = decltype(par)::value ? t0 : t1;
            if (kb + 1 < NKB) c.template tape_load<true>(0, kb + 1, a2n);
            float2 e2[2][4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 m2w3 = xy(F[4 * NP + tc16_pair(c, kb, k)]);  // -2 w3 * S_e / S_B pair (dz2 arrives times S_B)
#pragma unroll
                for (int rsel = 0; rsel < 2; ++rsel) {
                    const float2 a2 = tape_pair(a2q, rsel, k);
                    e2[rsel][k] = mul2(mul2(a2, mul2(one_minus_sq(a2), frag2(dz, k, rsel))), m2w3);
                }
            }
            c.template put_block<false>(kb, e2, 1.0f);
            if constexpr (SH::HAS_R) {
                if (kb >= PHNN_TC16_RSKEW) rback23(kb - PHNN_TC16_RSKEW, XA, XB);
            }
        });
        if constexpr (SH::HAS_R) {
#pragma unroll
            for (int kb = NKB - PHNN_TC16_RSKEW; kb < NKB; ++kb) rback23(kb, XA, XB);
        }
        c.end_feed();
    }
    // ---- C4: the dg1 half of xbar_H ----
    {
#if PHNN_TC16_CPF == 2
        float4 tb[4][4];
        c.template tape_load<true>(1, 0, tb[0]);
        c.template tape_load<true>(1, 1, tb[1]);
        const uint32_t tacc = c.acc_wait();
        for_acc_frags4<NKB>(c.tl16 + tacc, [&](int kb, const uint32_t (&dg)[16], auto par) {
            constexpr int PI = decltype(par)::value;
            float4 (&ac)[4] = tb[PI];
            if (kb + 2 < NKB) c.template tape_load<true>(1, kb + 2, tb[(PI + 2) & 3]);
#else
        float4 t0[4], t1[4];
        c.template tape_load<true>(1, 0, t0);
        const uint32_t tacc = c.acc_wait();
        for_acc_frags<NKB>(c.tl16 + tacc, [&](int kb, const uint32_t (&dg)[16], auto par) {
            float4 (&ac)[4] = decltype(par)::value ? t1 : t0;
            float4 (&a1n)[4] = decltype(par)::value ? t0 : t1;
            if (kb + 1 < NKB) c.template tape_load<true>(1, kb + 1, a1n);
#endif
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int P = tc16_pair(c, kb, k);
                const float4 w01 = F[P], w23 = F[NP + P];
                pair_scatter(w01, w23, mul2(one_minus_sq(tape_pair(ac, 0, k)), frag2(dg, k, 0)), XA);
                pair_scatter(w01, w23, mul2(one_minus_sq(tape_pair(ac, 1, k)), frag2(dg, k, 1)), XB);
            }
        });
        tc_fence_before();
    }
    // T is in units of the taped g1 (g1 * S_delta S_B), X in units of the dg1 accumulator (* S_e S_B): exact powers of two
    float X4[4];
    {
        const float ct = -2.f * p.s16[1], cx = p.s16[3];
        float xr[4] = {0.f, 0.f, 0.f, 0.f};  // the R_net chain's share, in the units of X (from the R_net warps)
        if constexpr (SH::RW) {
            c.r_wait();
#pragma unroll
            for (int i = 0; i < 4; ++i) xr[i] = cx * c.r_get(i);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
            X4[i] = (c.quad_own_sum(fmaf(ct, TA[i].x + TA[i].y, cx * (XA[i].x + XA[i].y)), fmaf(ct, TB[i].x + TB[i].y, cx * (XB[i].x + XB[i].y))) + xr[i]) * isc;
    }
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

// ---------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------
template <int MK, int NS, int HID, bool LOWP = false, bool SPARSE = false, bool PAIR = false>
__global__ void __launch_bounds__(Tc16Shape<MK, NS, HID, LOWP, SPARSE, PAIR>::THREADS, 1) phnn_tc16_kernel(const __grid_constant__ KParams p) {
    using SH = Tc16Shape<MK, NS, HID, LOWP, SPARSE, PAIR>;
    uint64_t* bars = reinterpret_cast<uint64_t*>(phnn_smem);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(phnn_smem + 512);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // CTA pairs (launched as clusters of two): rank 0 is the leader that issues the MMAs of both
    const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
    if (threadIdx.x == 0) {
        // warps that feed one K-block (of both CTAs of a pair: the other CTA's warps arrive on the leader's barriers)
        for (int e = 0; e < SH::NKB; ++e) mbar_init(&bars[SH::B_AFULL + e], SH::SPARSE ? 4 : (PAIR ? 2 * SH::NEW : SH::NEW));
        for (int e = 0; e < SH::NBE; ++e) {
            // leader of a pair: its own bulk copy + the relay of the other CTA ("my half has landed")
            mbar_init(&bars[SH::B_BFULL + e], (PAIR && crank == 0) ? 2 : 1);
            mbar_init(&bars[SH::B_BEMPTY + e], 1);
        }
        mbar_init(&bars[SH::B_ACC + 0], 1);
        mbar_init(&bars[SH::B_ACC + 1], 1);
        mbar_init(&bars[SH::B_SMALL], 1);
        if constexpr (SH::RW) {
            mbar_init(&bars[SH::B_RREQ], SH::NEW);  // every element warp posts
            mbar_init(&bars[SH::B_RRESP], 4);       // every R_net warp answers
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if constexpr (PAIR) {
        // barriers of both CTAs initialised before either allocates tensor memory or arrives remotely
        __syncthreads();
        cluster_sync_all();
        if (warp == SH::NEW) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                         "r"((uint32_t)SH::TMEM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    } else if (warp == SH::NEW) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                     "r"((uint32_t)SH::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) cluster_sync_all();
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
    const long long nprod = (steal ? 1LL : (long long)n_outer) * (2LL * nfwd + 2LL * nadj);
    const long long per_iter = 2LL * nfwd + 2LL * nadj;
    const long long my_tiles = p.tiles > (long long)blockIdx.x ? (p.tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const size_t tape_eval = (size_t)3 * HID * 128;  // floats per taped evaluation
    float* const tape = (p.tape && nadj > 0) ? p.tape + (size_t)blockIdx.x * E * tape_eval : nullptr;
    using Steal = typename std::conditional<PAIR, PairStealSched, StealSched>::type;
    Steal ss;
    ss.counter = p.sched;
    ss.progress = p.sched + 1;
    ss.tiles = p.tiles;
    ss.iters = p.iters;
    ss.slot = reinterpret_cast<int*>(phnn_smem + 768);
    ss.nelem = 32 * SH::NEW;
    if constexpr (PAIR) ss.rank = crank;
    else ss.nsync = SH::RW ? 32 * SH::NEW + 128 : 0;  // the R_net warps take no part in the scheduling

    if (warp < SH::NEW) {
        // ===== element threads =====
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(SH::REGS_ELEM));
        Tc16Ctx<SH> c;
        const int quad = warp & 3, half = SH::SPARSE ? 0 : (warp >> 2), g = lane >> 2;
        c.ksel = SH::SPARSE ? (warp >> 2) : 0;
        c.xpar = 0;
        c.slot = quad * 16 + g + 8 * ((lane >> 1) & 1);
        c.lane = lane;
        c.cq = lane & 3;
        c.srcA = lane & ~3;
        c.srcB = (lane & ~3) | 2;
        c.row = quad * 32 + half * 16 + g + 8 * ((lane >> 1) & 1);
        c.tid = threadIdx.x;
        c.tl16 = tbase + ((uint32_t)(quad * 32 + half * 16) << 16);
        c.afull = PAIR ? mapa_u32(smem_u32(&bars[SH::B_AFULL]), 0) : 0u;
        c.qdone = 0;
        c.qfeed = 0;
        c.rq = 0;
        c.store = (lane & 1) == 0 && (!SH::SPARSE || (warp >> 2) == 0);
        c.tape = tape;
        c.scratch = p.scratch ? p.scratch + (size_t)blockIdx.x * tc_scratch_floats_per_cta(NS, p.T, p.S) : nullptr;
        c.sck = nullptr;
        c.ev = 0;
        mbar_wait(&bars[SH::B_SMALL], 0);
        if (steal) {
            run_job(c, p, ss, SH::SPARSE ? c.slot : c.row);
        } else if constexpr (!PAIR) {  // pairs are launched for work-stealing solve jobs only
            StridedSched sched{(long long)blockIdx.x, (long long)blockIdx.x, p.tiles, (int)gridDim.x, n_outer, 0};
            run_job(c, p, sched, SH::SPARSE ? c.slot : c.row);
        }
        if constexpr (SH::RW) c.r_post(2, nullptr, nullptr);  // no more work for the R_net warps
        tc_fence_before();
    } else if (SH::RW && warp >= SH::NEW + 4) {
        // ===== R_net warps: thread = instance, all hidden units, records read warp-uniformly =====
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(SH::REGS_R));
        if constexpr (SH::RW) {
            constexpr int NP = SH::NP;
            const float4* F = reinterpret_cast<const float4*>(phnn_smem + SH::OFF_SMALL);
            float* m = reinterpret_cast<float*>(phnn_smem + SH::OFF_MBOX);
            // warp rw: instances 64 (rw & 1) + lane and + 32 (two per thread: every record load serves both), pairs of
            // units [64 (rw >> 1), +64) (the two halves are added by the reader)
            const int rw = warp - (SH::NEW + 4);
            const int i0 = 64 * (rw & 1) + lane, i1 = i0 + 32;
            const int PB = (NP / 2) * (rw >> 1), PE = PB + NP / 2;
            float* out = m + (14 + 10 * (rw >> 1)) * 128;
            mbar_wait(&bars[SH::B_SMALL], 0);
#pragma unroll 1
            for (uint32_t n = 0;; ++n) {
                mbar_wait(&bars[SH::B_RREQ], n & 1u);
                const int tag = *reinterpret_cast<volatile int*>(m + SH::MBOX_ROWS * 128);
                if (tag == 2) break;
                float y0[4], y1[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) { y0[i] = m[i * 128 + i0]; y1[i] = m[i * 128 + i1]; }
#ifdef PHNN_TC16_EXP_REARLY  // timing experiment (wrong results): answer first, then do the work (same resources, no latency)
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars[SH::B_RRESP]);
#endif
#ifdef PHNN_TC16_EXP_RFREE  // timing experiment (wrong results): the R_net warps answer at once
                if (true) {
#pragma unroll
                    for (int i = 0; i < 10; ++i) { out[i * 128 + i0] = 0.01f * y0[i & 3]; out[i * 128 + i1] = 0.01f * y1[i & 3]; }
                } else
#endif
                if (tag == 0) {
                    // forward: Sp[c] = sum_j Wr2sym[c][j] tanh(Wr1[j] . y + br1[j])   (src/pHNN.py:70-79)
                    // (PHNN_TC16_RBATCH pairs of units per step: one warp per scheduler has to cover the LDS -> FMA -> MUFU
                    // latencies of these chains with its own instruction-level parallelism)
                    float2 S0[10], S1[10];
#pragma unroll
                    for (int i = 0; i < 10; ++i) S0[i] = S1[i] = make_float2(0.f, 0.f);
#pragma unroll 1
                    for (int P0 = PB; P0 < PE; P0 += PHNN_TC16_RBATCH) {
                        float2 r0[PHNN_TC16_RBATCH], r1[PHNN_TC16_RBATCH];
#pragma unroll
                        for (int q = 0; q < PHNN_TC16_RBATCH; ++q) {
                            const float4 u01 = F[5 * NP + P0 + q], u23 = F[6 * NP + P0 + q];
                            const float2 br1 = zw(F[2 * NP + P0 + q]);
                            r0[q] = pair_affine(u01, u23, y0, br1);
                            r1[q] = pair_affine(u01, u23, y1, br1);
                        }
#pragma unroll
                        for (int q = 0; q < PHNN_TC16_RBATCH; ++q) { r0[q] = tanh16(r0[q]); r1[q] = tanh16(r1[q]); }
#pragma unroll
                        for (int j = 0; j < 5; ++j)
#pragma unroll
                            for (int q = 0; q < PHNN_TC16_RBATCH; ++q) {
                                const float4 cc = F[(7 + j) * NP + P0 + q];
                                S0[2 * j] = fma2(xy(cc), r0[q], S0[2 * j]);
                                S0[2 * j + 1] = fma2(zw(cc), r0[q], S0[2 * j + 1]);
                                S1[2 * j] = fma2(xy(cc), r1[q], S1[2 * j]);
                                S1[2 * j + 1] = fma2(zw(cc), r1[q], S1[2 * j + 1]);
                            }
                    }
#pragma unroll
                    for (int i = 0; i < 10; ++i) { out[i * 128 + i0] = S0[i].x + S0[i].y; out[i * 128 + i1] = S1[i].x + S1[i].y; }
                } else {
                    // adjoint: X[i] = sum_j Wr1[j][i] (1 - r_j^2) sum_c Rb[c] Wr2sym[c][j]
                    float Rb0[10], Rb1[10];
#pragma unroll
                    for (int i = 0; i < 10; ++i) { Rb0[i] = m[(4 + i) * 128 + i0]; Rb1[i] = m[(4 + i) * 128 + i1]; }
                    float2 X0[4], X1[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) X0[i] = X1[i] = make_float2(0.f, 0.f);
#pragma unroll 1
                    for (int P0 = PB; P0 < PE; P0 += PHNN_TC16_RBATCH) {
                        float2 rb0[PHNN_TC16_RBATCH], rb1[PHNN_TC16_RBATCH], r0[PHNN_TC16_RBATCH], r1[PHNN_TC16_RBATCH];
                        float4 u01[PHNN_TC16_RBATCH], u23[PHNN_TC16_RBATCH];
#pragma unroll
                        for (int q = 0; q < PHNN_TC16_RBATCH; ++q) {
                            u01[q] = F[5 * NP + P0 + q];
                            u23[q] = F[6 * NP + P0 + q];
                            const float2 br1 = zw(F[2 * NP + P0 + q]);
                            r0[q] = pair_affine(u01[q], u23[q], y0, br1);
                            r1[q] = pair_affine(u01[q], u23[q], y1, br1);
                        }
#pragma unroll
                        for (int q = 0; q < PHNN_TC16_RBATCH; ++q) { r0[q] = tanh16(r0[q]); r1[q] = tanh16(r1[q]); }
#pragma unroll
                        for (int q = 0; q < PHNN_TC16_RBATCH; ++q) {
                            const float4 cc = F[7 * NP + P0 + q];
                            rb0[q] = fma2(zw(cc), bc2(Rb0[1]), mul2(xy(cc), bc2(Rb0[0])));
                            rb1[q] = fma2(zw(cc), bc2(Rb1[1]), mul2(xy(cc), bc2(Rb1[0])));
                        }
#pragma unroll
                        for (int j = 1; j < 5; ++j)
#pragma unroll
                            for (int q = 0; q < PHNN_TC16_RBATCH; ++q) {
                                const float4 cc = F[(7 + j) * NP + P0 + q];
                                rb0[q] = fma2(xy(cc), bc2(Rb0[2 * j]), rb0[q]);
                                rb0[q] = fma2(zw(cc), bc2(Rb0[2 * j + 1]), rb0[q]);
                                rb1[q] = fma2(xy(cc), bc2(Rb1[2 * j]), rb1[q]);
                                rb1[q] = fma2(zw(cc), bc2(Rb1[2 * j + 1]), rb1[q]);
                            }
#pragma unroll
                        for (int q = 0; q < PHNN_TC16_RBATCH; ++q) {
                            pair_scatter(u01[q], u23[q], mul2(rb0[q], one_minus_sq(r0[q])), X0);
                            pair_scatter(u01[q], u23[q], mul2(rb1[q], one_minus_sq(r1[q])), X1);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) { out[i * 128 + i0] = X0[i].x + X0[i].y; out[i * 128 + i1] = X1[i].x + X1[i].y; }
                }
#ifndef PHNN_TC16_EXP_REARLY
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars[SH::B_RRESP]);
#endif
            }
        }
    } else if (warp == SH::NEW) {
        // ===== MMA issuer =====
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(SH::REGS_AUX));
        const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(HID >> 3) << 17) | ((uint32_t)((PAIR ? 256 : 128) >> 4) << 24);  // F16 x F16 -> F32
        const uint32_t b_base = smem_u32(phnn_smem + SH::OFF_B);
        uint32_t bent = 0;
        long long qtot = 0;  // products issued so far (accumulator / operand parity continues across units)
        for (long long unit = 0;; ++unit) {
            if (steal ? ss.grab() < 0 : unit >= my_tiles) break;
            if constexpr (PAIR) {
                if (crank != 0) {
                    // the non-leader's MMA warp only relays: its half of the weight tile has landed -> leader's barrier
                    if (lane == 0) {
                        const uint32_t lead_bfull = mapa_u32(smem_u32(&bars[SH::B_BFULL]), 0);
#pragma unroll 1
                        for (long long qk = 0; qk < nprod * SH::NKB; ++qk) {
                            const uint32_t e = bent % SH::NBE;
                            mbar_wait_aux(&bars[SH::B_BFULL + e], (bent / SH::NBE) & 1u);
                            mbar_arrive_cluster(lead_bfull + 8u * e);
                            ++bent;
                        }
                    }
                    __syncwarp();
                    continue;
                }
            }
            if (lane == 0) {
#pragma unroll 1
                for (long long qq = 0; qq < nprod; ++qq, ++qtot) {
                    const uint32_t par = (uint32_t)(qtot & 1);
                    const uint32_t acc = tbase + par * HID;
                    const uint32_t afeed = tbase + (par ^ 1u) * HID;
#pragma unroll 1
                    for (int kb = 0; kb < SH::NKB; ++kb) {
                        const uint32_t e = bent % SH::NBE;
                        mbar_wait_aux(&bars[SH::B_AFULL + kb], par);
                        mbar_wait_aux(&bars[SH::B_BFULL + e], (bent / SH::NBE) & 1u);
                        tc_fence_after();
                        const uint32_t b_t = b_base + e * SH::B_TILE_CTA;
                        const uint32_t a_hi = afeed + kb * 32, a_lo = a_hi + 16;
                        // a_hi b_hi + a_hi b_lo + a_lo b_hi; the weight row is [b_hi (64 B) | b_lo (64 B)], K = 16 per MMA
                        if constexpr (PAIR) {
                            umma_f16_ts_pair(acc, a_hi, umma_desc_sw128(b_t), idesc, kb ? 1u : 0u);
                            umma_f16_ts_pair(acc, a_hi + 8, umma_desc_sw128(b_t + 32), idesc, 1u);
                            if constexpr (!SH::LOWP) {
                                umma_f16_ts_pair(acc, a_hi, umma_desc_sw128(b_t + 64), idesc, 1u);
                                umma_f16_ts_pair(acc, a_hi + 8, umma_desc_sw128(b_t + 96), idesc, 1u);
                                umma_f16_ts_pair(acc, a_lo, umma_desc_sw128(b_t), idesc, 1u);
                                umma_f16_ts_pair(acc, a_lo + 8, umma_desc_sw128(b_t + 32), idesc, 1u);
                            }
                            umma_commit_pair(&bars[SH::B_BEMPTY + e]);
                        } else {
                            umma_f16_ts(acc, a_hi, umma_desc_sw128(b_t), idesc, kb ? 1u : 0u);
                            umma_f16_ts(acc, a_hi + 8, umma_desc_sw128(b_t + 32), idesc, 1u);
                            if constexpr (!SH::LOWP) {
                                umma_f16_ts(acc, a_hi, umma_desc_sw128(b_t + 64), idesc, 1u);
                                umma_f16_ts(acc, a_hi + 8, umma_desc_sw128(b_t + 96), idesc, 1u);
                                umma_f16_ts(acc, a_lo, umma_desc_sw128(b_t), idesc, 1u);
                                umma_f16_ts(acc, a_lo + 8, umma_desc_sw128(b_t + 32), idesc, 1u);
                            }
                            umma_commit(&bars[SH::B_BEMPTY + e]);
                        }
                        ++bent;
                    }
                    if constexpr (PAIR) umma_commit_pair(&bars[SH::B_ACC + par]);
                    else umma_commit(&bars[SH::B_ACC + par]);
                }
            }
            __syncwarp();
        }
    } else if (warp > SH::NEW + 1 && warp < SH::NEW + 4) {
        // ===== idle warps of the auxiliary warpgroup: give their registers away, keep the CTA-wide barriers company =====
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(SH::REGS_AUX));
        for (long long unit = 0;; ++unit)
            if (steal ? ss.grab() < 0 : true) break;
    } else {
        // ===== weight producer (bulk copies of pre-swizzled K-blocks) + L2 prefetch of the tape =====
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(SH::REGS_AUX));
        if (lane == 0) {
            mbar_expect_tx(&bars[SH::B_SMALL], SH::SMALL * 4);
            bulk_g2s(phnn_smem + SH::OFF_SMALL, p.wsmall16, SH::SMALL * 4, &bars[SH::B_SMALL]);
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
                    const unsigned char* src = p.wtc16 + (size_t)(qtot & 1) * SH::NKB * SH::B_TILE + (size_t)crank * SH::B_TILE_CTA;
#pragma unroll 1
                    for (int kb = 0; kb < SH::NKB; ++kb) {
                        if (qi >= 0) {
                            // pull the tape blocks the element threads will read PF_AHEAD K-block steps from now into L2
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
                        const uint32_t e = bent % SH::NBE;
                        mbar_wait_aux(&bars[SH::B_BEMPTY + e], ((bent / SH::NBE) & 1u) ^ 1u);
                        mbar_expect_tx(&bars[SH::B_BFULL + e], SH::B_TILE_CTA);
                        bulk_g2s(phnn_smem + SH::OFF_B + e * SH::B_TILE_CTA, src + (size_t)kb * SH::B_TILE, SH::B_TILE_CTA, &bars[SH::B_BFULL + e]);
                        ++bent;
                    }
                }
            }
            __syncwarp();
        }
    }
    __syncthreads();
    if constexpr (PAIR) {
        // the leader's MMAs read the other CTA's shared and tensor memory, its barriers receive remote arrivals: neither CTA
        // may leave (or free its tensor memory) before both are done
        cluster_sync_all();
        if (warp == SH::NEW) {
            tc_fence_after();
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"((uint32_t)SH::TMEM_COLS) : "memory");
        }
    } else if (warp == SH::NEW) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"((uint32_t)SH::TMEM_COLS) : "memory");
    }
}

}  // namespace phnn
