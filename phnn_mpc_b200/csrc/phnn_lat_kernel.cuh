// phnn_lat_kernel.cuh -- latency path for small batches (the reference's own use: one MPC instance
// per solve, scripts/run_cartpole_mpc.py:129; B up to a few hundred).
//
// One CTA per instance, one thread per hidden unit.  The batched kernels amortise every weight
// over 32 or 128 instances; with a single instance that leaves the SM almost idle (a B=1 solve
// took 72 ms on the FP32-FMA kernel: 600 sequential evaluation pairs, each paced by one warp pair
// at low issue rate).  Here thread k keeps the k-th rows of the small layers in registers, W2^T
// and W2 sit in shared memory in the k-major layout of the packed blob (thread k reads column k,
// lanes coalesce, the activation vector is a shared-memory broadcast), the two products of the
// adjoint that share an operand matrix run in one loop, and the reductions over the hidden
// dimension are a warp transpose-reduce plus one shared-memory exchange.  All per-instance
// algebra is evaluated redundantly by every thread (uniform values), so run_job is reused as is.
#pragma once
#include "phnn_kernel.cuh"

namespace phnn {

// threads per CTA of the latency kernel (512 = 8 instances at h = 64, 4 at h = 128; 128 registers per thread)
#ifndef PHNN_LAT_THREADS
#define PHNN_LAT_THREADS 512
#endif
#ifndef PHNN_LAT_MINBLOCKS
#define PHNN_LAT_MINBLOCKS 1
#endif

template <int MK_, int NS_, int HID_>
struct LatShape {
    static constexpr int MK = MK_, NS = NS_, HID = HID_, NN = NS * NS;
    static constexpr bool HAS_R = (MK != MK_CANON);
    static constexpr bool HAS_GNET = (MK == MK_PHNN_GNET);
    static constexpr int NW = HID / 32;  // warps per instance
    // instances per CTA: they share the W2^T | W2 copy in shared memory, which is what limited residency to one
    // (h = 128) or a handful (h = 64) of instances per SM; each instance keeps its own vectors, reduction
    // exchange and named barrier
    static constexpr int THREADS = PHNN_LAT_THREADS;
    static constexpr int NI = THREADS / HID;
    using FS = Shape<MK, NS, HID>;       // small-weight blob layout of the FP32 kernel
    // shared memory (floats): W2^T | W2 (2*HID*HID), then per instance 4 vectors of HID and the reduction exchange [32][NW]
    static constexpr int O_W = 0;
    static constexpr int O_V = 2 * HID * HID;
    static constexpr int PER = 4 * HID + 32 * NW;
    static constexpr int FLOATS = O_V + NI * PER;
    static constexpr size_t SMEM_BYTES = 128 + sizeof(float) * (size_t)FLOATS;
    static_assert(SMEM_BYTES <= 232448, "W2 and W2^T must fit in shared memory (h <= 128)");
};

// Sum 32 values per lane over the 32 lanes of a warp: afterwards v[0] of lane l is the warp total
// of the original v[l].
__device__ __forceinline__ void warp_transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
    for (int half = 16, bit = 16; half >= 1; half >>= 1, bit >>= 1) {
        const bool up = (lane & bit) != 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (i < half) {
                const float send = up ? v[i] : v[i + half];
                const float keep = up ? v[i + half] : v[i];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
            }
        }
    }
}

template <class SH> struct LatCtx;
template <class SH>
__device__ __noinline__ void lat_eval_fwd(LatCtx<SH>& c, const KParams& p, const float (&y)[SH::NS], float u, float (&f)[SH::NS],
                                          float& Hval);
template <class SH>
__device__ __noinline__ void lat_eval_vjp(LatCtx<SH>& c, const KParams& p, const float (&y)[SH::NS], float u,
                                          const float (&v)[SH::NS], float (&xbar)[SH::NS], float& ubar);

template <class SH>
struct LatCtx {
    static constexpr int NS = SH::NS;
    static constexpr int TW = 1;
    __device__ static int ws_extra(const KParams&) { return 0; }
    __device__ __forceinline__ void set_eval(int e) { ev = e; }
    int ev;                   // evaluation in flight (t * S + s): row of the training-mode emission
    long long inst_global;    // index of my instance in the batch
    int k, lane, warp, inst;  // hidden unit, lane, warp within the instance, instance slot of the CTA
    bool store;
    // this hidden unit's rows of the small layers
    float w1[SH::NS], b1, b2, w3;
    float wr1[SH::HAS_R ? SH::NS : 1], br1, wr2[SH::HAS_R ? SH::NN : 1];
    float wg1[SH::HAS_GNET ? SH::NS : 1], bg1, wg2[SH::HAS_GNET ? SH::NS : 1];

    __device__ __forceinline__ float* smf() const { return reinterpret_cast<float*>(phnn_smem + 128); }
    __device__ __forceinline__ const float* W2T() const { return smf() + SH::O_W; }
    __device__ __forceinline__ const float* W2() const { return smf() + SH::O_W + SH::HID * SH::HID; }
    __device__ __forceinline__ float* vec(int i) const { return smf() + SH::O_V + inst * SH::PER + i * SH::HID; }
    __device__ __forceinline__ float* part() const { return smf() + SH::O_V + inst * SH::PER + 4 * SH::HID; }
    __device__ __forceinline__ void gbar() const { group_bar(1 + inst, SH::HID); }  // the threads of my instance
    __device__ __forceinline__ void begin_unit(const KParams&, long long) {}
    __device__ __forceinline__ float* unit_scratch() const { return nullptr; }
    __device__ __forceinline__ void peer_store(const KParams&, long long) {}
    __device__ __forceinline__ void eval_fwd(const KParams& p, const float (&y)[NS], float u, float (&f)[NS], float& H) {
        lat_eval_fwd(*this, p, y, u, f, H);
    }
    __device__ __forceinline__ void eval_vjp(const KParams& p, const float (&y)[NS], float u, const float (&v)[NS],
                                             float (&xbar)[NS], float& ubar) {
        lat_eval_vjp(*this, p, y, u, v, xbar, ubar);
    }
    // block totals of the first V of 32 per-thread values; every thread receives all of them.
    // Contains one instance barrier; the caller guarantees another barrier before the next call.
    template <int V>
    __device__ __forceinline__ void block_reduce(float (&v)[32], float (&out)[V]) {
        warp_transpose_reduce32(v, lane);
        part()[lane * SH::NW + warp] = v[0];
        gbar();
#pragma unroll
        for (int i = 0; i < V; ++i) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < SH::NW; ++w) s += part()[i * SH::NW + w];
            out[i] = s;
        }
    }
    // one column of  M^T-style product:  sum_j M[j][k] * a[j]  with M k-major in shared memory
    __device__ __forceinline__ float matvec(const float* __restrict__ M, const float* __restrict__ a) const {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 8
        for (int j = 0; j < SH::HID; j += 4) {
            const float4 a4 = *reinterpret_cast<const float4*>(a + j);
            s0 = fmaf(M[(j + 0) * SH::HID + k], a4.x, s0);
            s1 = fmaf(M[(j + 1) * SH::HID + k], a4.y, s1);
            s2 = fmaf(M[(j + 2) * SH::HID + k], a4.z, s2);
            s3 = fmaf(M[(j + 3) * SH::HID + k], a4.w, s3);
        }
        return (s0 + s1) + (s2 + s3);
    }
    // the same matrix applied to two vectors in one sweep (z2 with dz2, g1 with dg1)
    __device__ __forceinline__ void matvec2(const float* __restrict__ M, const float* __restrict__ a, const float* __restrict__ b,
                                            float& ra, float& rb) const {
        float s0 = 0.f, s1 = 0.f, t0 = 0.f, t1 = 0.f;
#pragma unroll 8
        for (int j = 0; j < SH::HID; j += 2) {
            const float2 a2 = *reinterpret_cast<const float2*>(a + j);
            const float2 b2v = *reinterpret_cast<const float2*>(b + j);
            const float m0 = M[(j + 0) * SH::HID + k], m1 = M[(j + 1) * SH::HID + k];
            s0 = fmaf(m0, a2.x, s0);
            s1 = fmaf(m1, a2.y, s1);
            t0 = fmaf(m0, b2v.x, t0);
            t1 = fmaf(m1, b2v.y, t1);
        }
        ra = s0 + s1;
        rb = t0 + t1;
    }
};

template <int NS>
__device__ __forceinline__ float dotv(const float* w, const float (&x)[NS], float b) {
    float s = b;
#pragma unroll
    for (int i = 0; i < NS; ++i) s = fmaf(w[i], x[i], s);
    return s;
}

// S = (Rraw + Rraw^T)/2 from reduced totals tot[0..NN)
template <int NS>
__device__ __forceinline__ void lat_make_S(const KParams& p, const float* tot, float (&S)[NS][NS]) {
    float R[NS * NS];
#pragma unroll
    for (int e = 0; e < NS * NS; ++e) R[e] = tot[e] + p.br2[e];
#pragma unroll
    for (int a = 0; a < NS; ++a)
#pragma unroll
        for (int b = 0; b < NS; ++b) S[a][b] = (R[a * NS + b] + R[b * NS + a]) * 0.5f;
}

// ---------------------------------------------------------------------------------------
// f(y,u), H(y)   (src/pHNN.py:52-100, src/pHNN_canonical.py:172-273)
// ---------------------------------------------------------------------------------------
template <class SH>
__device__ __noinline__ void lat_eval_fwd(LatCtx<SH>& c, const KParams& p, const float (&y)[SH::NS], float u, float (&f)[SH::NS],
                                          float& Hval) {
    constexpr int NS = SH::NS, NN = SH::NN;
    // reduction slots: [0,NN) Rraw, [NN,NN+NS) Graw, [NN+NS] H, [NN+NS+1, NN+2NS+1) dH
    constexpr int SL_G = SH::HAS_R ? NN : 0, SL_H = SL_G + (SH::HAS_GNET ? NS : 0), SL_DH = SL_H + 1, NV = SL_DH + NS;
    static_assert(NV <= 32, "reduction width");
    float z[NS];
    Canon cq = {};
    if constexpr (SH::MK == MK_CANON) {
        cq = canon_of(p, y[1]);
        z[0] = y[0]; z[1] = y[1];
        z[2] = p.ma * y[2] + cq.beta * y[3];
        z[3] = cq.beta * y[2] + p.mc * y[3];
    } else {
#pragma unroll
        for (int i = 0; i < NS; ++i) z[i] = y[i];
    }
    float red[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) red[i] = 0.f;
    const float a1 = tanh_acc(dotv<NS>(c.w1, z, c.b1));
    c.vec(0)[c.k] = a1;
    if constexpr (SH::HAS_R) {
        const float r1 = tanh_acc(dotv<NS>(c.wr1, y, c.br1));
#pragma unroll
        for (int e = 0; e < NN; ++e) red[e] = c.wr2[e] * r1;
    }
    if constexpr (SH::HAS_GNET) {
        const float ag = tanh_acc(dotv<NS>(c.wg1, y, c.bg1));
#pragma unroll
        for (int a = 0; a < NS; ++a) red[SL_G + a] = c.wg2[a] * ag;
    }
    c.gbar();  // a1 visible (also separates this evaluation's exchange from the previous one's reads)
    const float a2 = tanh_acc(c.matvec(c.W2T(), c.vec(0)) + c.b2);
    red[SL_H] = c.w3 * a2;
    c.vec(1)[c.k] = fmaf(-a2, a2, 1.f) * c.w3;
    c.gbar();  // delta2 visible
    const float g1 = c.matvec(c.W2(), c.vec(1));
    const float d1 = fmaf(-a1, a1, 1.f) * g1;
#pragma unroll
    for (int i = 0; i < NS; ++i) red[SL_DH + i] = c.w1[i] * d1;
    float tot[NV];
    c.template block_reduce<NV>(red, tot);
    Hval = tot[SL_H] + p.b3;
    const float* g = tot + SL_DH;
    if constexpr (SH::MK == MK_CANON) {
        float pd[2];
#pragma unroll
        for (int r = 2; r < 4; ++r) {
            float s = 0.f;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) s = fmaf(p.Jm[r * 4 + kk] - (r == kk ? p.rdiag[r] : 0.f), g[kk], s);
            pd[r - 2] = s + p.Gv[r] * u;
        }
        f[0] = cq.n11 * z[2] + cq.n12 * z[3];
        f[1] = cq.n12 * z[2] + cq.n22 * z[3];
        f[2] = cq.n11 * pd[0] + cq.n12 * pd[1];
        f[3] = cq.n12 * pd[0] + cq.n22 * pd[1];
    } else {
        float S[NS][NS];
        lat_make_S<NS>(p, tot, S);
#pragma unroll
        for (int a = 0; a < NS; ++a) {
            float s = 0.f;
#pragma unroll
            for (int b = 0; b < NS; ++b) {
                float Rab = 0.f;
#pragma unroll
                for (int kk = 0; kk < NS; ++kk) Rab = fmaf(S[a][kk], S[b][kk], Rab);
                s = fmaf(p.Jm[a * NS + b] - Rab, g[b], s);
            }
            const float Ga = SH::HAS_GNET ? tot[SL_G + a] + p.bg2[a] : p.Gv[a];
            f[a] = s + Ga * u;
        }
    }
}

// ---------------------------------------------------------------------------------------
// xbar = (df/dy)^T v, ubar = (df/du)^T v  
// ---------------------------------------------------------------------------------------
template <class SH>
__device__ __noinline__ void lat_eval_vjp(LatCtx<SH>& c, const KParams& p, const float (&y)[SH::NS], float u,
                                          const float (&v)[SH::NS], float (&xbar)[SH::NS], float& ubar) {
    constexpr int NS = SH::NS, NN = SH::NN;
    constexpr int SL_G = SH::HAS_R ? NN : 0, NV1 = SL_G + (SH::HAS_GNET ? NS : 0);
    float z[NS], w[NS];
    Canon cq = {};
    float pb[2] = {0.f, 0.f}, pdb[2] = {0.f, 0.f};
    float S[NS][NS], sv[NS], G[NS];
    float r1 = 0.f, ag = 0.f;
    float red[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) red[i] = 0.f;
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
        for (int kk = 0; kk < 4; ++kk) {
            float s = 0.f;
#pragma unroll
            for (int r = 2; r < 4; ++r) s = fmaf(p.Jm[r * 4 + kk] - (r == kk ? p.rdiag[r] : 0.f), pdb[r - 2], s);
            w[kk] = s;
        }
    } else {
#pragma unroll
        for (int i = 0; i < NS; ++i) z[i] = y[i];
        // S needs the whole R_net output: first exchange
        r1 = tanh_acc(dotv<NS>(c.wr1, y, c.br1));
#pragma unroll
        for (int e = 0; e < NN; ++e) red[e] = c.wr2[e] * r1;
        if constexpr (SH::HAS_GNET) {
            ag = tanh_acc(dotv<NS>(c.wg1, y, c.bg1));
#pragma unroll
            for (int a = 0; a < NS; ++a) red[SL_G + a] = c.wg2[a] * ag;
        }
        c.gbar();  // the previous evaluation's exchange has been read by everyone
        float tot[NV1 > 0 ? NV1 : 1];
        c.template block_reduce<(NV1 > 0 ? NV1 : 1)>(red, tot);
        lat_make_S<NS>(p, tot, S);
#pragma unroll
        for (int a = 0; a < NS; ++a) G[a] = SH::HAS_GNET ? tot[SL_G + a] + p.bg2[a] : p.Gv[a];
#pragma unroll
        for (int a = 0; a < NS; ++a) {
            float s = 0.f;
#pragma unroll
            for (int b = 0; b < NS; ++b) s = fmaf(S[a][b], v[b], s);
            sv[a] = s;
        }
#pragma unroll
        for (int a = 0; a < NS; ++a) {  // w = (J - J^T)^T v - S (S v)
            float s = 0.f;
#pragma unroll
            for (int b = 0; b < NS; ++b) s = fmaf(p.Jm[b * NS + a], v[b], s);
#pragma unroll
            for (int b = 0; b < NS; ++b) s = fmaf(-S[a][b], sv[b], s);
            w[a] = s;
        }
    }
    const float a1 = tanh_acc(dotv<NS>(c.w1, z, c.b1));
    const float s1 = fmaf(-a1, a1, 1.f);
    const float da1 = s1 * dotv<NS>(c.w1, w, 0.f);
    c.vec(0)[c.k] = a1;
    c.vec(1)[c.k] = da1;
    c.gbar();
    float z2, dz2;
    c.matvec2(c.W2T(), c.vec(0), c.vec(1), z2, dz2);
    const float a2 = tanh_acc(z2 + c.b2);
    const float s2 = fmaf(-a2, a2, 1.f);
    c.vec(2)[c.k] = s2 * c.w3;
    c.vec(3)[c.k] = -2.f * a2 * (s2 * dz2) * c.w3;
    c.gbar();
    float g1, dg1;
    c.matvec2(c.W2(), c.vec(2), c.vec(3), g1, dg1);
    const float t = fmaf(-2.f * a1 * da1, g1, s1 * dg1);
    const float d1 = s1 * g1;
#pragma unroll
    for (int i = 0; i < 32; ++i) red[i] = 0.f;
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        red[i] = c.w1[i] * d1;       // dH
        red[NS + i] = c.w1[i] * t;   // Hessian-vector product
    }
    float tot2[2 * NS];
    c.template block_reduce<2 * NS>(red, tot2);
    // training mode (MODE_PARAMGRAD): the per-unit factors of the parameter cotangents of this evaluation,
    //   Wbar2 += delta2 (x) da1 + e2 (x) a1,  Wbar1 += delta1 (x) w + zbar1 (x) z,  bbar1 += zbar1,  bbar2 += e2,
    //   wbar3 += s2 * dz2   (reverse of the double-backward chain of grad H)
    const size_t erow = p.emit ? (size_t)c.inst_global * ((size_t)p.T * p.S) + c.ev : 0;
    float* const esmall = p.emit ? p.emit + (size_t)EMIT_NARR * p.emit_rows * SH::HID + erow * EMIT_SMALL : nullptr;
    auto earr = [&](int a) -> float& { return p.emit[((size_t)a * p.emit_rows + erow) * SH::HID + c.k]; };
    if (p.emit) {
        earr(0) = a1; earr(1) = da1; earr(2) = s2 * c.w3; earr(3) = -2.f * a2 * (s2 * dz2) * c.w3;
        earr(4) = d1; earr(5) = t; earr(6) = s2 * dz2;
        if (c.store) {
#pragma unroll
            for (int i = 0; i < NS; ++i) { esmall[i] = z[i]; esmall[4 + i] = w[i]; esmall[8 + i] = v[i]; esmall[12 + i] = tot2[i]; esmall[16 + i] = y[i]; }
        }
    }
    if constexpr (SH::MK == MK_CANON) {
        const float* g = tot2;
        if (p.emit && c.store) { esmall[36] = pdb[0] * g[2]; esmall[37] = pdb[1] * g[3]; }
        float zb[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) zb[i] = tot2[NS + i];
        float pd[2];
#pragma unroll
        for (int r = 2; r < 4; ++r) {
            float s = 0.f;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) s = fmaf(p.Jm[r * 4 + kk] - (r == kk ? p.rdiag[r] : 0.f), g[kk], s);
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
        const float* g = tot2;
        float tg[NS];
#pragma unroll
        for (int a = 0; a < NS; ++a) {
            float s = 0.f;
#pragma unroll
            for (int b = 0; b < NS; ++b) s = fmaf(S[a][b], g[b], s);
            tg[a] = s;
        }
        // aux nets backward, thread-local: hidden cotangent of my unit, times tanh', through my input row
        float rb = 0.f;
#pragma unroll
        for (int a = 0; a < NS; ++a)
#pragma unroll
            for (int b = 0; b < NS; ++b)
                rb = fmaf(c.wr2[a * NS + b], -0.5f * (v[a] * tg[b] + g[a] * sv[b] + v[b] * tg[a] + g[b] * sv[a]), rb);
        const float zb = rb * fmaf(-r1, r1, 1.f);
        float zg = 0.f;
        if constexpr (SH::HAS_GNET) {
            float gb = 0.f;
#pragma unroll
            for (int a = 0; a < NS; ++a) gb = fmaf(c.wg2[a], v[a] * u, gb);
            zg = gb * fmaf(-ag, ag, 1.f);
        }
        if (p.emit) {
            // R_net: Wbar_r2 += Rbar_raw (x) r1, bbar_r2 += Rbar_raw, Wbar_r1 += dr (x) y, bbar_r1 += dr; G_net likewise
            earr(7) = r1; earr(8) = zb;
            if constexpr (SH::HAS_GNET) { earr(9) = ag; earr(10) = zg; }
            if (c.store) {
#pragma unroll
                for (int a = 0; a < NS; ++a)
#pragma unroll
                    for (int b = 0; b < NS; ++b)
                        esmall[20 + a * NS + b] = -0.5f * (v[a] * tg[b] + g[a] * sv[b] + v[b] * tg[a] + g[b] * sv[a]);
#pragma unroll
                for (int a = 0; a < NS; ++a) esmall[40 + a] = v[a] * u;
            }
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) red[i] = 0.f;
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            red[i] = c.wr1[i] * zb;
            if constexpr (SH::HAS_GNET) red[i] = fmaf(c.wg1[i], zg, red[i]);
        }
        c.gbar();  // everyone has read the previous exchange
        float tot3[NS];
        c.template block_reduce<NS>(red, tot3);
#pragma unroll
        for (int i = 0; i < NS; ++i) xbar[i] = tot2[NS + i] + tot3[i];
        float s = 0.f;
#pragma unroll
        for (int a = 0; a < NS; ++a) s = fmaf(G[a], v[a], s);
        ubar = s;
    }
}

// ---------------------------------------------------------------------------------------
// the kernel: grid = ceil(B / ng) CTAs of up to NI instances (ng in use), HID threads per instance
// ---------------------------------------------------------------------------------------
template <int MK, int NS, int HID>
__global__ void __launch_bounds__(LatShape<MK, NS, HID>::THREADS, PHNN_LAT_MINBLOCKS) phnn_lat_kernel(const __grid_constant__ KParams p) {
    using SH = LatShape<MK, NS, HID>;
    using FS = typename SH::FS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(phnn_smem);
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // W2^T | W2 in one TMA bulk copy per half (each HID*HID*4 bytes)
        mbar_expect_tx(&bars[0], 2u * HID * HID * 4u);
        bulk_g2s(phnn_smem + 128, p.wbig, HID * HID * 4u, &bars[0]);
        bulk_g2s(phnn_smem + 128 + HID * HID * 4, p.wbig + HID * HID, HID * HID * 4u, &bars[0]);
    }
    LatCtx<SH> c;
    c.k = threadIdx.x % HID;
    c.inst = threadIdx.x / HID;
    c.lane = threadIdx.x & 31;
    c.warp = c.k >> 5;
    c.store = (c.k == 0);
    const float* ws = p.wsmall;
#pragma unroll
    for (int i = 0; i < NS; ++i) c.w1[i] = ws[FS::O_W1 + c.k * NS + i];
    c.b1 = ws[FS::O_B1 + c.k];
    c.b2 = ws[FS::O_B2 + c.k];
    c.w3 = ws[FS::O_W3 + c.k];
    c.br1 = 0.f; c.bg1 = 0.f;
    if constexpr (SH::HAS_R) {
#pragma unroll
        for (int i = 0; i < NS; ++i) c.wr1[i] = ws[FS::O_WR1 + c.k * NS + i];
        c.br1 = ws[FS::O_BR1 + c.k];
#pragma unroll
        for (int e = 0; e < SH::NN; ++e) c.wr2[e] = ws[FS::O_WR2 + e * HID + c.k];
    }
    if constexpr (SH::HAS_GNET) {
#pragma unroll
        for (int i = 0; i < NS; ++i) c.wg1[i] = ws[FS::O_WG1 + c.k * NS + i];
        c.bg1 = ws[FS::O_BG1 + c.k];
#pragma unroll
        for (int a = 0; a < NS; ++a) c.wg2[a] = ws[FS::O_WG2 + a * HID + c.k];
    }
    mbar_wait(&bars[0], 0);
    // p.ng = instance slots in use per CTA (the host spreads a small batch over the SMs before it stacks instances
    // on one: co-resident instances share the shared-memory pipe and run ~2x slower each); unused slots and slots
    // past the end of the batch leave here (every later barrier is per instance)
    const long long my_instance = (long long)blockIdx.x * p.ng + c.inst;
    if (c.inst >= p.ng || my_instance >= p.B) return;
    c.inst_global = my_instance;
    c.ev = 0;
    int n_outer = (p.mode == MODE_SOLVE) ? p.iters : 1;
    StaticSched sched{my_instance, n_outer, 0};
    run_job(c, p, sched, 0);
}

}  // namespace phnn
