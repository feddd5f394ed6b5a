// phnn_capi.cu -- C ABI (include/phnn_mpc.h) over the fused kernel in phnn_kernel.cuh.
// Host side only: weight packing, argument checks, launch configuration.  No CPU fallback:
// every entry point either launches the sm_100a kernel or returns an error.
#include "../../include/phnn_mpc.h"
#include "phnn_kernel.cuh"
#include "phnn_tc_kernel.cuh"
#include "phnn_tc16_kernel.cuh"
#include "phnn_lat_kernel.cuh"

#include <cuda_fp16.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace phnn;

// CTA pairs (tcgen05 cta_group::2) for tensor_mode 4 solve jobs: option "tensor_pair"
#ifndef PHNN_TC_PAIR_DEFAULT
#define PHNN_TC_PAIR_DEFAULT 0
#endif

struct phnn_pack {
    int abi_kind, mk, n, m, h;
    int device, num_sms;
    int tc_mode;          // 0: FP32-FMA kernel only; tcgen05 kernel when eligible: 3 = 3xTF32, 2 = TF32 + BF16 correction
                          // product, 1 = plain TF32
    long tc_min_batch;    // smallest B routed to the tcgen05 kernel
    long lat_max_batch;   // largest B routed to the one-CTA-per-instance latency kernel (0 = never)
    long tc_fwd_min_batch;  // forward-only tcgen05 shapes: smallest B of a forward job (forward / rollout / cost) routed there (0 = never)
    int tc_fwd_sparse;      // forward-only tcgen05 shapes: 64-instance tiles while they fit one per SM (1, default) or always 128 (0)
    int tc_pair;            // tensor_mode 4 solve jobs: clusters of two CTAs sharing every weight tile (tcgen05 cta_group::2)
    float* d_small;
    float* d_big;
    unsigned char* d_wtc;
    float* d_small_tc;
    unsigned char* d_wtc2;  // tensor_mode 2 weights (TF32 hi tiles + BF16 correction tiles)
    unsigned char* d_wtc16; // tensor_mode 4 weights (FP16 hi | lo tiles, scaled) and its small-layer records
    float* d_small16;
    size_t small_floats;
    KParams base;  // model constants filled in once
};

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
static int cuda_fail(cudaError_t e, const char* what) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return (int)e;
}
#define CUDA_TRY(x)                                   \
    do {                                              \
        cudaError_t e_ = (x);                         \
        if (e_ != cudaSuccess) return cuda_fail(e_, #x); \
    } while (0)

extern "C" const char* phnn_last_error(void) { return g_err; }
extern "C" int phnn_version(void) { return 100; }

// ---- dispatch over the compiled (model kind, state dim, hidden width) instantiations ----
#ifdef PHNN_DEV_CFG4  // fast experiment builds (tools/devbuild.sh): only the cfg4 shape
#define PHNN_SHAPES(X) X(MK_PHNN, 4, 256)
#define PHNN_TC_SHAPES(X) X(MK_PHNN, 4, 256)
#define PHNN_LAT_SHAPES(X) X(MK_PHNN, 4, 64)
#endif
#ifdef PHNN_DEV_CFG2  // fast experiment builds: only the pendulum shape (BASELINE cfg2)
#define PHNN_SHAPES(X) X(MK_PHNN_GNET, 2, 64)
#define PHNN_TC_SHAPES(X)
#define PHNN_LAT_SHAPES(X) X(MK_PHNN_GNET, 2, 64)
#define PHNN_TC16_FWD_SHAPES(X) X(MK_PHNN_GNET, 2, 64)
#endif
#ifndef PHNN_SHAPES
#define PHNN_SHAPES(X) \
    X(MK_PHNN, 4, 64)  \
    X(MK_PHNN, 4, 128) \
    X(MK_PHNN, 4, 256) \
    X(MK_PHNN, 2, 64)  \
    X(MK_PHNN_GNET, 2, 64) \
    X(MK_PHNN_GNET, 4, 128) \
    X(MK_CANON, 4, 64) \
    X(MK_CANON, 4, 128) \
    X(MK_CANON, 4, 256)
#endif

// shapes with a tcgen05 instantiation (cart-pole pHNN, fixed G)
#ifndef PHNN_TC_SHAPES
#define PHNN_TC_SHAPES(X) \
    X(MK_PHNN, 4, 128)    \
    X(MK_PHNN, 4, 256)    \
    X(MK_CANON, 4, 128)   \
    X(MK_CANON, 4, 256)
#endif

// forward-only instantiations of the second-generation tcgen05 kernel (forward evaluation, rollouts, cost without
// gradient): the n = 2 pHNN with fixed or learned G (pendulum model, BASELINE cfg2)
#ifndef PHNN_TC16_FWD_SHAPES
#ifdef PHNN_DEV_CFG4
#define PHNN_TC16_FWD_SHAPES(X)
#else
#define PHNN_TC16_FWD_SHAPES(X) \
    X(MK_PHNN, 2, 64)          \
    X(MK_PHNN_GNET, 2, 64)
#endif
#endif
static bool has_tc16_fwd_shape(int mk, int n, int h) {
    (void)mk; (void)n; (void)h;
#define X(MK, NS, HID) \
    if (mk == MK && n == NS && h == HID) return true;
    PHNN_TC16_FWD_SHAPES(X)
#undef X
    return false;
}

static bool has_tc_shape(int mk, int n, int h) {
#define X(MK, NS, HID) \
    if (mk == MK && n == NS && h == HID) return true;
    PHNN_TC_SHAPES(X)
#undef X
    return false;
}

// shapes with a latency-kernel instantiation (W2 and W2^T must fit in shared memory: h <= 128)
#ifndef PHNN_LAT_SHAPES
#define PHNN_LAT_SHAPES(X) \
    X(MK_PHNN, 4, 64)      \
    X(MK_PHNN, 4, 128)     \
    X(MK_PHNN, 2, 64)      \
    X(MK_PHNN_GNET, 2, 64) \
    X(MK_PHNN_GNET, 4, 128) \
    X(MK_CANON, 4, 64)     \
    X(MK_CANON, 4, 128)
#endif

static bool has_lat_shape(int mk, int n, int h) {
#define X(MK, NS, HID) \
    if (mk == MK && n == NS && h == HID) return true;
    PHNN_LAT_SHAPES(X)
#undef X
    return false;
}

static float tf32_round_host(float x) {  // cvt.rna.tf32.f32
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7F800000u) != 0x7F800000u) u += 0x1000u;
    u &= 0xFFFFE000u;
    float y;
    memcpy(&y, &u, 4);
    return y;
}

// pre-swizzled K-blocks: [P: W2 | W2^T][kb][hi|lo][h rows x 128 B]
static void fill_tc_big(const phnn_model_desc* d, std::vector<unsigned char>& out) {
    const int h = d->h, nkb = h / 32;
    const size_t tile = (size_t)h * 128;
    out.assign((size_t)2 * nkb * 2 * tile, 0);
    for (int P = 0; P < 2; ++P)
        for (int kb = 0; kb < nkb; ++kb)
            for (int n = 0; n < h; ++n)
                for (int c = 0; c < 32; ++c) {
                    const int K = kb * 32 + c;
                    const float val = (P == 0) ? d->W2[(size_t)n * h + K] : d->W2[(size_t)K * h + n];
                    const float hi = tf32_round_host(val), lo = val - hi;
                    const size_t base = ((size_t)(P * nkb + kb) * 2) * tile;
                    memcpy(&out[base + sw128_off(n, c)], &hi, 4);
                    memcpy(&out[base + tile + sw128_off(n, c)], &lo, 4);
                }
}

static uint16_t bf16_round_host(float x) {  // cvt.rn.bf16.f32 (finite inputs)
    uint32_t u;
    memcpy(&u, &x, 4);
    u += 0x7FFFu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}

// tensor_mode 2: [P: W2 | W2^T][kb][ TF32 hi tile | BF16 tile of 64 per row = (b (32) | b_lo (32)) ], each h rows x 128 B,
// K-major SWIZZLE_128B.  The BF16 tile meets the element threads' (a_lo | a) rows in one K = 64 product.
static void fill_tc_big_bf16(const phnn_model_desc* d, std::vector<unsigned char>& out) {
    const int h = d->h, nkb = h / 32;
    const size_t tile = (size_t)h * 128;
    out.assign((size_t)2 * nkb * 2 * tile, 0);
    for (int P = 0; P < 2; ++P)
        for (int kb = 0; kb < nkb; ++kb)
            for (int n = 0; n < h; ++n)
                for (int c = 0; c < 32; ++c) {
                    const int K = kb * 32 + c;
                    const float val = (P == 0) ? d->W2[(size_t)n * h + K] : d->W2[(size_t)K * h + n];
                    const float hi = tf32_round_host(val);
                    const uint16_t b16 = bf16_round_host(val), lo16 = bf16_round_host(val - hi);
                    const size_t base = ((size_t)(P * nkb + kb) * 2) * tile;
                    memcpy(&out[base + sw128_off(n, c)], &hi, 4);
                    // BF16 tile: element column j of 64 sits at byte 2 j of the 128-byte row (16-byte chunk j / 8)
                    const size_t row = base + tile + (size_t)(n >> 3) * 1024 + (n & 7) * 128;
                    const int j0 = c, j1 = 32 + c;
                    memcpy(&out[row + ((((j0 >> 3) ^ (n & 7)) & 7) << 4) + (j0 & 7) * 2], &b16, 2);
                    memcpy(&out[row + ((((j1 >> 3) ^ (n & 7)) & 7) << 4) + (j1 & 7) * 2], &lo16, 2);
                }
}

// pair-interleaved records of the small layers (layout in TcShape): every field is {unit 2P, unit 2P+1}
static void fill_tc_small(const phnn_model_desc* d, std::vector<float>& s) {
    const int h = d->h, n = d->n;   // n == 4 for every tcgen05 shape
    const bool has_r = d->kind == PHNN_KIND_PHNN;
    constexpr int RA = 20, RB = 8, RCP = 20;
    s.assign((size_t)(h / 2) * (has_r ? RA + RB + RCP : RA), 0.f);
    float* rA = s.data();
    float* rB = rA + (size_t)(h / 2) * RA;
    float* rC = rB + (size_t)(h / 2) * RB;
    for (int k = 0; k < h; ++k) {
        const int P = k >> 1, o = k & 1;
        for (int i = 0; i < n; ++i) rA[P * RA + 2 * i + o] = d->W1[k * n + i];
        rA[P * RA + 8 + o] = d->b1[k];
        rA[P * RA + 12 + o] = d->b2[k];
        rA[P * RA + 14 + o] = d->W3[k];
        rA[P * RA + 16 + o] = -2.f * d->W3[k];
        if (!has_r) continue;
        rA[P * RA + 10 + o] = d->br1[k];
        for (int i = 0; i < n; ++i) rB[P * RB + 2 * i + o] = d->Wr1[k * n + i];
        for (int a = 0; a < n; ++a)
            for (int b = a; b < n; ++b)
                rC[P * RCP + 2 * sym_idx(a, b) + o] = 0.5f * (d->Wr2[(size_t)(a * n + b) * h + k] + d->Wr2[(size_t)(b * n + a) * h + k]);
    }
}

// ---- tensor_mode 4 (phnn_tc16_kernel.cuh) ------------------------------------------------------------------
static float pow2f(int e) { return std::ldexp(1.0f, e); }
static int floor_log2(float x) {  // floor(log2 x) for finite x > 0
    int e;
    std::frexp(x, &e);
    return e - 1;
}
struct Tc16Scales {
    int eB, eA, eD, eE, wexp;  // S_B = 2^eB (weights), S_a = 2^eA (a1), S_delta = 2^eD (w3 in delta2), S_e = 2^eE (w3 in e2)
    int eRW;                   // S_RW = 2^eRW: symmetrised R_net output weights, max |Wr2sym| S_RW in [2^9, 2^10)
};
// Exact power-of-two scales that put every FP16 operand of the four products in [~2^-2, 2^15] (FP16: 11-bit
// significand, normal range 6.1e-5 .. 65504; hi + lo keeps 22 bits as long as |lo| >= 6.1e-5, i.e. |x| >= 0.25):
//   weights    max |W2| S_B in [2^9, 2^10)                       a1 in [-1, 1]: S_a = 2^9
//   delta2 = (1 - a2^2) w3: max |w3| S_delta in [2^9, 2^10)
//   da1 = s1 (W1 w'): |da1| <= rho1 max|w'|, rho1 = max row L1 norm of W1; max|w'| < 2^(wexp+1) <= 2^14 / rho1
//   e2 = -2 a2 s2 dz2 w3 S_e: |2 a2 s2| <= 0.77, |dz2| <= rho2 max|da1| (rho2 = max row L1 norm of W2): bound <= 2^15
static Tc16Scales tc16_scales(const phnn_model_desc* d) {
    const int h = d->h, n = d->n;
    float maxW2 = 0.f, maxw3 = 0.f, rho1 = 0.f, rho2 = 0.f;
    for (int j = 0; j < h; ++j) {
        float r1 = 0.f, r2 = 0.f;
        for (int i = 0; i < n; ++i) r1 += std::fabs(d->W1[j * n + i]);
        for (int k = 0; k < h; ++k) {
            const float a = std::fabs(d->W2[(size_t)j * h + k]);
            r2 += a;
            maxW2 = std::fmax(maxW2, a);
        }
        rho1 = std::fmax(rho1, r1);
        rho2 = std::fmax(rho2, r2);
        maxw3 = std::fmax(maxw3, std::fabs(d->W3[j]));
    }
    maxW2 = std::fmax(maxW2, 1e-30f); maxw3 = std::fmax(maxw3, 1e-30f);
    rho1 = std::fmax(rho1, 1e-30f); rho2 = std::fmax(rho2, 1e-30f);
    Tc16Scales s;
    s.eB = 9 - floor_log2(maxW2);
    s.eA = 9;
    s.eD = 9 - floor_log2(maxw3);
    s.wexp = 13 - (floor_log2(rho1) + 1);                 // rho1 2^(wexp+1) <= 2^14
    const float bound = 0.77f * maxw3 * rho2 * rho1 * pow2f(s.wexp + 1);
    s.eE = 15 - (floor_log2(bound) + 1);                  // bound 2^eE <= 2^15
    s.eRW = 0;
    if (d->kind == PHNN_KIND_PHNN && d->Wr2) {
        float maxR = 1e-30f;
        for (int a = 0; a < n; ++a)
            for (int b = a; b < n; ++b)
                for (int k = 0; k < h; ++k)
                    maxR = std::fmax(maxR, std::fabs(0.5f * (d->Wr2[(size_t)(a * n + b) * h + k] + d->Wr2[(size_t)(b * n + a) * h + k])));
        s.eRW = 9 - floor_log2(maxR);
    }
    return s;
}

// [P: W2 | W2^T][kb][h rows x 128 B], row n = [b_hi (32 fp16) | b_lo (32 fp16)] of B[n][32 kb ..] * S_B, K-major SWIZZLE_128B
static void fill_tc16_big(const phnn_model_desc* d, const Tc16Scales& sc, std::vector<unsigned char>& out) {
    const int h = d->h, nkb = h / 32;
    const size_t tile = (size_t)h * 128;
    const float SB = pow2f(sc.eB);
    out.assign((size_t)2 * nkb * tile, 0);
    for (int P = 0; P < 2; ++P)
        for (int kb = 0; kb < nkb; ++kb)
            for (int n = 0; n < h; ++n)
                for (int c = 0; c < 32; ++c) {
                    const int K = kb * 32 + c;
                    const float val = SB * ((P == 0) ? d->W2[(size_t)n * h + K] : d->W2[(size_t)K * h + n]);
                    const __half hi = __float2half_rn(val);
                    const __half lo = __float2half_rn(val - __half2float(hi));
                    const size_t row = ((size_t)(P * nkb + kb)) * tile + (size_t)(n >> 3) * 1024 + (n & 7) * 128;
                    const int j0 = c, j1 = 32 + c;  // 16-bit column j sits at byte 2 j of the row (16-byte chunk j / 8, swizzled)
                    memcpy(&out[row + ((((j0 >> 3) ^ (n & 7)) & 7) << 4) + (j0 & 7) * 2], &hi, 2);
                    memcpy(&out[row + ((((j1 >> 3) ^ (n & 7)) & 7) << 4) + (j1 & 7) * 2], &lo, 2);
                }
}

// field-major records (layout in Tc16Shape): field f, pair P -> float4 at (f * h/2 + P)
static void fill_tc16_small(const phnn_model_desc* d, const Tc16Scales& sc, std::vector<float>& s) {
    const int h = d->h, n = d->n, np = h / 2;
    const bool has_r = d->kind == PHNN_KIND_PHNN;
    // R_net models: behind the fields, the symmetrised output layer as mma.sync.m16n8k16 B fragments (FP16 hi | lo, times
    // S_RW), one uint4 {hi(b0,b1), hi(b2,b3), lo(b0,b1), lo(b2,b3)} per lane:
    //   forward  [kb][step][n-tile][lane]: B[kk][c] = Wr2sym[c = 8 nt + lane/4][unit 32 kb + 16 step + kk], kk = 2 t + {0,1,8,9}
    //   backward [kb][group][lane]:        B[c][u]  = Wr2sym[c = 2 t + {0,1,8,9}][unit 32 kb + 8 group + lane/4]      (t = lane % 4)
    const int nkb = h / 32;
    const size_t nfield = (size_t)(has_r ? 12 : 5) * np * 4;
    s.assign(nfield + (has_r ? (size_t)nkb * 4 * 32 * 4 * 2 : 0), 0.f);
    auto at = [&](int f, int P, int half, int o) -> float& { return s[((size_t)f * np + P) * 4 + half * 2 + o]; };
    if (has_r) {
        const float SR = pow2f(sc.eRW);
        auto wsym = [&](int c, int k) -> float {
            if (c >= 10) return 0.f;
            int a = 0, b = 0;
            for (int aa = 0; aa < n; ++aa)
                for (int bb = aa; bb < n; ++bb)
                    if (sym_idx(aa, bb) == c) { a = aa; b = bb; }
            return SR * 0.5f * (d->Wr2[(size_t)(a * n + b) * h + k] + d->Wr2[(size_t)(b * n + a) * h + k]);
        };
        auto pack = [&](float lo_el, float hi_el, uint32_t& whi, uint32_t& wlo) {  // lo_el in the low 16 bits
            const __half h0 = __float2half_rn(lo_el), h1 = __float2half_rn(hi_el);
            const __half l0 = __float2half_rn(lo_el - __half2float(h0)), l1 = __float2half_rn(hi_el - __half2float(h1));
            unsigned short u0, u1, v0, v1;
            memcpy(&u0, &h0, 2); memcpy(&u1, &h1, 2); memcpy(&v0, &l0, 2); memcpy(&v1, &l1, 2);
            whi = (uint32_t)u0 | ((uint32_t)u1 << 16);
            wlo = (uint32_t)v0 | ((uint32_t)v1 << 16);
        };
        uint32_t* rf = reinterpret_cast<uint32_t*>(s.data() + nfield);
        uint32_t* rb = rf + (size_t)nkb * 4 * 32 * 4;
        for (int kb = 0; kb < nkb; ++kb)
            for (int lane = 0; lane < 32; ++lane) {
                const int g = lane >> 2, t = lane & 3;
                for (int step = 0; step < 2; ++step)
                    for (int nt = 0; nt < 2; ++nt) {
                        uint32_t* w = rf + ((((size_t)kb * 2 + step) * 2 + nt) * 32 + lane) * 4;
                        const int c = 8 * nt + g, u0 = 32 * kb + 16 * step + 2 * t;
                        pack(wsym(c, u0), wsym(c, u0 + 1), w[0], w[2]);
                        pack(wsym(c, u0 + 8), wsym(c, u0 + 9), w[1], w[3]);
                    }
                for (int grp = 0; grp < 4; ++grp) {
                    uint32_t* w = rb + (((size_t)kb * 4 + grp) * 32 + lane) * 4;
                    const int u = 32 * kb + 8 * grp + g;
                    pack(wsym(2 * t, u), wsym(2 * t + 1, u), w[0], w[2]);
                    pack(wsym(2 * t + 8, u), wsym(2 * t + 9, u), w[1], w[3]);
                }
            }
    }
    const float SD = pow2f(sc.eD), SE = pow2f(sc.eE - sc.eB);  // e2 is formed from the dz2 accumulator, which carries S_B
    for (int k = 0; k < h; ++k) {
        const int P = k >> 1, o = k & 1;
        at(0, P, 0, o) = d->W1[k * n + 0]; at(0, P, 1, o) = d->W1[k * n + 1];
        at(1, P, 0, o) = d->W1[k * n + 2]; at(1, P, 1, o) = d->W1[k * n + 3];
        at(2, P, 0, o) = d->b1[k];
        at(3, P, 0, o) = d->b2[k];         at(3, P, 1, o) = d->W3[k] * SD;
        at(4, P, 0, o) = -2.f * d->W3[k] * SE;
        if (!has_r) continue;
        at(2, P, 1, o) = d->br1[k];
        at(5, P, 0, o) = d->Wr1[k * n + 0]; at(5, P, 1, o) = d->Wr1[k * n + 1];
        at(6, P, 0, o) = d->Wr1[k * n + 2]; at(6, P, 1, o) = d->Wr1[k * n + 3];
        for (int a = 0; a < n; ++a)
            for (int b = a; b < n; ++b) {
                const int i = sym_idx(a, b);
                at(7 + i / 2, P, i & 1, o) = 0.5f * (d->Wr2[(size_t)(a * n + b) * h + k] + d->Wr2[(size_t)(b * n + a) * h + k]);
            }
    }
}

// n = 2 forward-only records (layout in Tc16Shape): field f, pair P -> float4 at (f * h/2 + P)
static void fill_tc16_small_n2(const phnn_model_desc* d, const Tc16Scales& sc, std::vector<float>& s) {
    const int h = d->h, np = h / 2;
    const bool gnet = d->learned_G != 0;
    s.assign((size_t)(gnet ? 8 : 6) * np * 4, 0.f);
    auto at = [&](int f, int P, int half, int o) -> float& { return s[((size_t)f * np + P) * 4 + half * 2 + o]; };
    const float SD = pow2f(sc.eD);
    for (int k = 0; k < h; ++k) {
        const int P = k >> 1, o = k & 1;
        at(0, P, 0, o) = d->W1[k * 2 + 0];  at(0, P, 1, o) = d->W1[k * 2 + 1];
        at(1, P, 0, o) = d->b1[k];          at(1, P, 1, o) = d->b2[k];
        at(2, P, 0, o) = d->W3[k] * SD;     at(2, P, 1, o) = d->br1[k];
        at(3, P, 0, o) = d->Wr1[k * 2 + 0]; at(3, P, 1, o) = d->Wr1[k * 2 + 1];
        at(4, P, 0, o) = d->Wr2[(size_t)0 * h + k];
        at(4, P, 1, o) = 0.5f * (d->Wr2[(size_t)1 * h + k] + d->Wr2[(size_t)2 * h + k]);
        at(5, P, 0, o) = d->Wr2[(size_t)3 * h + k];
        if (gnet) {
            at(5, P, 1, o) = d->bg1[k];
            at(6, P, 0, o) = d->Wg1[k * 2 + 0]; at(6, P, 1, o) = d->Wg1[k * 2 + 1];
            at(7, P, 0, o) = d->Wg2[(size_t)0 * h + k]; at(7, P, 1, o) = d->Wg2[(size_t)1 * h + k];
        }
    }
}

template <class SH>
static void fill_small(const phnn_model_desc* d, std::vector<float>& s) {
    constexpr int NS = SH::NS, HID = SH::HID, NN = SH::NN;
    s.assign(SH::SMALL, 0.f);
    memcpy(&s[SH::O_W1], d->W1, sizeof(float) * HID * NS);
    memcpy(&s[SH::O_B1], d->b1, sizeof(float) * HID);
    memcpy(&s[SH::O_B2], d->b2, sizeof(float) * HID);
    memcpy(&s[SH::O_W3], d->W3, sizeof(float) * HID);
    if (SH::HAS_R) {
        memcpy(&s[SH::O_WR1], d->Wr1, sizeof(float) * HID * NS);
        memcpy(&s[SH::O_BR1], d->br1, sizeof(float) * HID);
        memcpy(&s[SH::O_WR2], d->Wr2, sizeof(float) * NN * HID);
    }
    if (SH::HAS_GNET) {
        memcpy(&s[SH::O_WG1], d->Wg1, sizeof(float) * HID * NS);
        memcpy(&s[SH::O_BG1], d->bg1, sizeof(float) * HID);
        memcpy(&s[SH::O_WG2], d->Wg2, sizeof(float) * NS * HID);
    }
}

static bool shape_small(int mk, int n, int h, const phnn_model_desc* d, std::vector<float>& s) {
#define X(MK, NS, HID)                                \
    if (mk == MK && n == NS && h == HID) {            \
        fill_small<Shape<MK, NS, HID>>(d, s);         \
        return true;                                  \
    }
    PHNN_SHAPES(X)
#undef X
    return false;
}

// The kernels of one model shape need more dynamic shared memory than the 48 KB default: raise the limit once per
// (kernel, device) when a pack for that shape is created on the device, not on every launch.
static cudaError_t set_smem_limits(int mk, int n, int h) {
    cudaError_t e = cudaSuccess;
#define X(MK, NS, HID)                                                                                                  \
    if (e == cudaSuccess && mk == MK && n == NS && h == HID)                                                            \
        e = cudaFuncSetAttribute(phnn_kernel<MK, NS, HID>, cudaFuncAttributeMaxDynamicSharedMemorySize,                 \
                                 (int)Shape<MK, NS, HID>::smem_bytes(Shape<MK, NS, HID>::MAX_NG));
    PHNN_SHAPES(X)
#undef X
#define X(MK, NS, HID)                                                                                                  \
    if (e == cudaSuccess && mk == MK && n == NS && h == HID)                                                            \
        e = cudaFuncSetAttribute(phnn_tc_kernel<MK, NS, HID>, cudaFuncAttributeMaxDynamicSharedMemorySize,              \
                                 (int)TcShape<MK, NS, HID>::SMEM_BYTES);
    PHNN_TC_SHAPES(X)
#undef X
#define X(MK, NS, HID)                                                                                                  \
    if (e == cudaSuccess && mk == MK && n == NS && h == HID)                                                            \
        e = cudaFuncSetAttribute(phnn_tc16_kernel<MK, NS, HID>, cudaFuncAttributeMaxDynamicSharedMemorySize,            \
                                 (int)Tc16Shape<MK, NS, HID>::SMEM_BYTES);
    PHNN_TC_SHAPES(X)
#undef X
#define X(MK, NS, HID)                                                                                                  \
    if (e == cudaSuccess && mk == MK && n == NS && h == HID)                                                            \
        e = cudaFuncSetAttribute(phnn_tc16_kernel<MK, NS, HID, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                 (int)Tc16Shape<MK, NS, HID>::SMEM_BYTES);
    PHNN_TC_SHAPES(X)
#undef X
#define X(MK, NS, HID)                                                                                                  \
    if (e == cudaSuccess && mk == MK && n == NS && h == HID)                                                            \
        e = cudaFuncSetAttribute(phnn_tc16_kernel<MK, NS, HID, false, false, true>,                                     \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize,                                            \
                                 (int)Tc16Shape<MK, NS, HID, false, false, true>::SMEM_BYTES);
    PHNN_TC_SHAPES(X)
#undef X
#define X(MK, NS, HID)                                                                                                  \
    if (e == cudaSuccess && mk == MK && n == NS && h == HID)                                                            \
        e = cudaFuncSetAttribute(phnn_tc16_kernel<MK, NS, HID>, cudaFuncAttributeMaxDynamicSharedMemorySize,            \
                                 (int)Tc16Shape<MK, NS, HID>::SMEM_BYTES);
    PHNN_TC16_FWD_SHAPES(X)
#undef X
#define X(MK, NS, HID)                                                                                                  \
    if (e == cudaSuccess && mk == MK && n == NS && h == HID)                                                            \
        e = cudaFuncSetAttribute(phnn_tc16_kernel<MK, NS, HID, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                 (int)Tc16Shape<MK, NS, HID, false, true>::SMEM_BYTES);
    PHNN_TC16_FWD_SHAPES(X)
#undef X
#define X(MK, NS, HID)                                                                                                  \
    if (e == cudaSuccess && mk == MK && n == NS && h == HID)                                                            \
        e = cudaFuncSetAttribute(phnn_lat_kernel<MK, NS, HID>, cudaFuncAttributeMaxDynamicSharedMemorySize,             \
                                 (int)LatShape<MK, NS, HID>::SMEM_BYTES);
    PHNN_LAT_SHAPES(X)
#undef X
    return e;
}

extern "C" int phnn_pack_create(const phnn_model_desc* d, int device, phnn_pack** out) {
    if (!d || !out) return fail(PHNN_E_ARG, "phnn_pack_create: null argument");
    *out = nullptr;
    if (d->m != 1) return fail(PHNN_E_UNSUPPORTED, "input_dim m=%d unsupported (kernels are built for m=1)", d->m);
    if (!d->W1 || !d->b1 || !d->W2 || !d->b2 || !d->W3 || !d->b3 || !d->J)
        return fail(PHNN_E_ARG, "phnn_pack_create: H_net / J pointers must be set");
    int mk;
    if (d->kind == PHNN_KIND_PHNN) {
        if (!d->Wr1 || !d->br1 || !d->Wr2 || !d->br2) return fail(PHNN_E_ARG, "R_net pointers must be set");
        if (d->learned_G) {
            if (!d->Wg1 || !d->bg1 || !d->Wg2 || !d->bg2) return fail(PHNN_E_ARG, "G_net pointers must be set");
            mk = MK_PHNN_GNET;
        } else {
            if (!d->G) return fail(PHNN_E_ARG, "fixed G must be set");
            mk = MK_PHNN;
        }
    } else if (d->kind == PHNN_KIND_CANONICAL) {
        if (!d->G || !d->r_diag) return fail(PHNN_E_ARG, "canonical model needs G and r_diag");
        if (d->learned_G) return fail(PHNN_E_UNSUPPORTED, "pHNN_Canonical requires fixed_G=True");
        mk = MK_CANON;
    } else {
        return fail(PHNN_E_ARG, "unknown model kind %d", d->kind);
    }
    std::vector<float> small;
    if (!shape_small(mk, d->n, d->h, d, small))
        return fail(PHNN_E_UNSUPPORTED, "no kernel for kind=%d learned_G=%d n=%d h=%d", d->kind, d->learned_G, d->n,
                    d->h);
    const int n = d->n, h = d->h;
    std::vector<float> big((size_t)2 * h * h);
    for (int j = 0; j < h; ++j)
        for (int k = 0; k < h; ++k) {
            big[(size_t)k * h + j] = d->W2[(size_t)j * h + k];                  // W2^T: [k][j]
            big[(size_t)h * h + (size_t)j * h + k] = d->W2[(size_t)j * h + k];  // W2:   [j][k]
        }
    phnn_pack* pk = new phnn_pack();
    memset(pk, 0, sizeof(*pk));
    pk->abi_kind = d->kind; pk->mk = mk; pk->n = n; pk->m = d->m; pk->h = h; pk->device = device;
    int prev = 0;
    cudaError_t e = cudaGetDevice(&prev);
    if (e == cudaSuccess) e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&pk->num_sms, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) e = set_smem_limits(mk, d->n, d->h);
    if (e == cudaSuccess) e = cudaMalloc(&pk->d_small, small.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&pk->d_big, big.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(pk->d_small, small.data(), small.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(pk->d_big, big.data(), big.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && has_tc_shape(mk, n, h)) {
        std::vector<unsigned char> wtc;
        std::vector<float> stc;
        fill_tc_big(d, wtc);
        fill_tc_small(d, stc);
        std::vector<unsigned char> wtc2;
        fill_tc_big_bf16(d, wtc2);
        e = cudaMalloc(&pk->d_wtc, wtc.size());
        if (e == cudaSuccess) e = cudaMalloc(&pk->d_wtc2, wtc2.size());
        if (e == cudaSuccess) e = cudaMemcpy(pk->d_wtc2, wtc2.data(), wtc2.size(), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMalloc(&pk->d_small_tc, stc.size() * sizeof(float));
        if (e == cudaSuccess) e = cudaMemcpy(pk->d_wtc, wtc.data(), wtc.size(), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(pk->d_small_tc, stc.data(), stc.size() * sizeof(float), cudaMemcpyHostToDevice);
        std::vector<unsigned char> w16;
        std::vector<float> s16;
        const Tc16Scales sc = tc16_scales(d);
        fill_tc16_big(d, sc, w16);
        fill_tc16_small(d, sc, s16);
        if (e == cudaSuccess) e = cudaMalloc(&pk->d_wtc16, w16.size());
        if (e == cudaSuccess) e = cudaMemcpy(pk->d_wtc16, w16.data(), w16.size(), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMalloc(&pk->d_small16, s16.size() * sizeof(float));
        if (e == cudaSuccess) e = cudaMemcpy(pk->d_small16, s16.data(), s16.size() * sizeof(float), cudaMemcpyHostToDevice);
        pk->base.s16[0] = pow2f(-(sc.eA + sc.eB));
        pk->base.s16[1] = pow2f(-(sc.eD + sc.eB));
        pk->base.s16[2] = pow2f(-sc.eB);
        pk->base.s16[3] = pow2f(-(sc.eE + sc.eB));
        pk->base.s16[4] = pow2f(sc.eA);
        pk->base.s16[5] = pow2f(-sc.eD);
        pk->base.s16[6] = pow2f(sc.eE + sc.eB);
        pk->base.s16[7] = pow2f(-(9 + sc.eRW));
        pk->base.wexp16 = sc.wexp;
        pk->base.rexp16 = sc.eRW;
        // default: the second-generation kernel (three FP16 products, operand A in tensor memory): FP32-level accuracy,
        // measured 1.2-1.3x the first-generation kernel's default (mode 2: TF32 + BF16 correction product)
        pk->tc_mode = 4;
        pk->tc_min_batch = 1;  // measured: the tcgen05 kernel beats the FP32-FMA kernel at every batch size (tools/gpu_crossover.py)
    }
    if (e == cudaSuccess && has_tc16_fwd_shape(mk, n, h)) {
        std::vector<unsigned char> w16;
        std::vector<float> s16;
        const Tc16Scales sc = tc16_scales(d);
        fill_tc16_big(d, sc, w16);
        fill_tc16_small_n2(d, sc, s16);
        e = cudaMalloc(&pk->d_wtc16, w16.size());
        if (e == cudaSuccess) e = cudaMemcpy(pk->d_wtc16, w16.data(), w16.size(), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMalloc(&pk->d_small16, s16.size() * sizeof(float));
        if (e == cudaSuccess) e = cudaMemcpy(pk->d_small16, s16.data(), s16.size() * sizeof(float), cudaMemcpyHostToDevice);
        pk->base.s16[0] = pow2f(-(sc.eA + sc.eB));
        pk->base.s16[1] = pow2f(-(sc.eD + sc.eB));
        pk->base.s16[2] = pow2f(-sc.eB);
        pk->base.s16[4] = pow2f(sc.eA);
        pk->base.s16[5] = pow2f(-sc.eD);
        pk->tc_mode = 4;
        // one tile per CTA runs its evaluations one after the other, so a forward job takes the same time whatever the
        // number of tiles up to one per SM: 1.5 ms for 100 RK4 steps on 64-instance tiles (up to 64 instances per SM),
        // 1.9 ms on 128-instance tiles.  The latency kernel (8 instances per CTA, 1.0 ms per wave of 8 per SM) is faster
        // below ~10 instances per SM (tools/gpu_cfg2_tc.py, measured on B200)
        pk->tc_fwd_min_batch = 10L * pk->num_sms;
        pk->tc_fwd_sparse = 1;
    }
    pk->tc_pair = PHNN_TC_PAIR_DEFAULT;
    cudaSetDevice(prev);
    if (e != cudaSuccess) {
        cudaFree(pk->d_small);
        cudaFree(pk->d_big);
        cudaFree(pk->d_wtc);
        cudaFree(pk->d_wtc2);
        cudaFree(pk->d_small_tc);
        cudaFree(pk->d_wtc16);
        cudaFree(pk->d_small16);
        delete pk;
        return cuda_fail(e, "phnn_pack_create");
    }
    pk->small_floats = small.size();
    // crossover measured on B200 (tools/gpu_crossover.py) with up to 8 (h = 64) / 4 (h = 128) instances per CTA: the
    // latency kernel wins up to ~100 instances per SM at h = 64, ~34 per SM at h = 128 against the FP32-FMA kernel and
    // up to ONE WAVE of its 4-instance CTAs against the second-generation tcgen05 kernel (h = 128, 50 Euler iterations
    // of H = 10: 5.7 ms at 592 instances, 11.0 ms from 593 on, tcgen05 7.7 ms for any batch up to one tile per SM)
    pk->lat_max_batch = !has_lat_shape(mk, n, h) ? 0 : (h <= 64 ? 96L : (pk->d_wtc ? 4L : 32L)) * pk->num_sms;
    KParams& P = pk->base;
    P.wsmall = pk->d_small;
    P.wbig = pk->d_big;
    P.wtc = pk->d_wtc;
    P.wsmall_tc = pk->d_small_tc;
    P.wtc16 = pk->d_wtc16;
    P.wsmall16 = pk->d_small16;
    for (int a = 0; a < n; ++a)
        for (int b = 0; b < n; ++b)
            P.Jm[a * n + b] = (mk == MK_CANON) ? d->J[a * n + b] : (d->J[a * n + b] - d->J[b * n + a]);
    for (int a = 0; a < n; ++a) P.Gv[a] = d->G ? d->G[a] : 0.f;
    P.b3 = d->b3[0];
    if (mk != MK_CANON) {
        for (int e2 = 0; e2 < n * n; ++e2) P.br2[e2] = d->br2[e2];
        if (n == 4)
            for (int a = 0; a < 4; ++a)
                for (int b = a; b < 4; ++b) P.bsym[sym_idx(a, b)] = 0.5f * (d->br2[a * 4 + b] + d->br2[b * 4 + a]);
        if (n == 2) {  // packed 00 01 11 (forward-only tcgen05 shapes)
            P.bsym[0] = d->br2[0];
            P.bsym[1] = 0.5f * (d->br2[1] + d->br2[2]);
            P.bsym[2] = d->br2[3];
        }
    }
    if (mk == MK_PHNN_GNET)
        for (int a = 0; a < n; ++a) P.bg2[a] = d->bg2[a];
    if (mk == MK_CANON) {
        P.ma = d->mass_a; P.mb = d->mass_b; P.mc = d->mass_c;
        P.mconst = d->mass_const != 0;
        for (int a = 0; a < n; ++a) P.rdiag[a] = d->r_diag[a];
    }
    *out = pk;
    return 0;
}

extern "C" int phnn_pack_destroy(phnn_pack* pk) {
    if (!pk) return 0;
    cudaFree(pk->d_small);
    cudaFree(pk->d_big);
    cudaFree(pk->d_wtc);
    cudaFree(pk->d_wtc2);
    cudaFree(pk->d_small_tc);
    cudaFree(pk->d_wtc16);
    cudaFree(pk->d_small16);
    delete pk;
    return 0;
}

extern "C" int phnn_pack_set_option(phnn_pack* pk, const char* key, long value) {
    if (!pk || !key) return fail(PHNN_E_ARG, "phnn_pack_set_option: null argument");
    if (!strcmp(key, "tensor_mode")) {
        if (value < 0 || value > 5) return fail(PHNN_E_ARG, "tensor_mode must be 0, 1, 2, 3, 4 or 5");
        if (value != 0 && !pk->d_wtc && !(value == 4 && pk->d_wtc16))
            return fail(PHNN_E_UNSUPPORTED, "no tcgen05 kernel for this model shape");
        pk->tc_mode = (int)value;
        return 0;
    }
    if (!strcmp(key, "tensor_min_batch")) {
        pk->tc_min_batch = value;
        return 0;
    }
    if (!strcmp(key, "tensor_fwd_min_batch")) {
        if (value > 0 && !has_tc16_fwd_shape(pk->mk, pk->n, pk->h))
            return fail(PHNN_E_UNSUPPORTED, "no forward-only tcgen05 kernel for this model shape");
        pk->tc_fwd_min_batch = value;
        return 0;
    }
    if (!strcmp(key, "tensor_fwd_sparse")) {
        pk->tc_fwd_sparse = value != 0;
        return 0;
    }
    if (!strcmp(key, "tensor_pair")) {
        pk->tc_pair = value != 0;
        return 0;
    }
    if (!strcmp(key, "latency_max_batch")) {
        if (value > 0 && !has_lat_shape(pk->mk, pk->n, pk->h)) return fail(PHNN_E_UNSUPPORTED, "no latency kernel for this model shape");
        pk->lat_max_batch = value;
        return 0;
    }
    return fail(PHNN_E_ARG, "unknown option %s", key);
}

#ifdef PHNN_TC_PROFILE
// profiling builds only (tools/gpu_tc_phases.py): device buffer (34 x int64) the instrumented kernel fills with
// per-phase cycle counts.  The production library has no such hook and no global mutable state.
static long long* g_dbg = nullptr;
extern "C" void phnn_debug_set_buffer(void* p) { g_dbg = (long long*)p; }
#endif

extern "C" long phnn_pack_get_option(const phnn_pack* pk, const char* key) {
    if (!pk || !key) return -1;
    if (!strcmp(key, "tensor_mode")) return pk->tc_mode;
    if (!strcmp(key, "tensor_min_batch")) return pk->tc_min_batch;
    if (!strcmp(key, "latency_max_batch")) return pk->lat_max_batch;
    if (!strcmp(key, "tensor_fwd_min_batch")) return pk->tc_fwd_min_batch;
    if (!strcmp(key, "tensor_fwd_sparse")) return pk->tc_fwd_sparse;
    if (!strcmp(key, "tensor_pair")) return pk->tc_pair;
    return -1;
}

extern "C" int phnn_pack_dims(const phnn_pack* pk, int* kind, int* n, int* m, int* h) {
    if (!pk) return fail(PHNN_E_ARG, "null pack");
    if (kind) *kind = pk->abi_kind;
    if (n) *n = pk->n;
    if (m) *m = pk->m;
    if (h) *h = pk->h;
    return 0;
}

template <class SH>
static int launch_shape(const phnn_pack* pk, KParams& P, cudaStream_t stream) {
    const long long groups = (P.B + GI - 1) / GI;
    long long ng = (groups + pk->num_sms - 1) / pk->num_sms;
    if (ng < 1) ng = 1;
    if (ng > SH::MAX_NG) ng = SH::MAX_NG;
    P.ng = (int)ng;
    const long long grid = (groups + ng - 1) / ng;
    const size_t smem = SH::smem_bytes((int)ng);
    auto kern = phnn_kernel<SH::MK, SH::NS, SH::HID>;  // dynamic shared-memory limit: set once in phnn_pack_create
    const int threads = ((int)ng * SH::NWG + 1) * 32;
    kern<<<(unsigned)grid, threads, smem, stream>>>(P);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

template <class SH>
static int launch_lat_shape(const phnn_pack* pk, KParams& P, cudaStream_t stream) {
    auto kern = phnn_lat_kernel<SH::MK, SH::NS, SH::HID>;
    // instances per CTA: one per SM first, stacked (up to NI, sharing the weights in shared memory) only when the batch
    // exceeds the SM count
    long long per = (P.B + pk->num_sms - 1) / pk->num_sms;
    if (per < 1) per = 1;
    if (per > SH::NI) per = SH::NI;
    P.ng = (int)per;
    kern<<<(unsigned)((P.B + per - 1) / per), SH::THREADS, SH::SMEM_BYTES, stream>>>(P);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// Workspace of a job with an adjoint on the tcgen05 kernels (floats unless noted):
//   [tiles x 128 x (3 T + 1)]                       what outlives a (tile, iteration) unit: Adam moments, best controls, best cost
//   [tiles + 4 ints, padded to 128 bytes]           work-stealing scheduler words
//   [grid x T*S x 3 x h x 128]                      activation tape, one region per CTA (grid = min(tiles, SMs))
//   [grid x T*S x (n + 16) x 128]                   per-CTA stage states + R_net sums / grad H of the unit in flight
struct TcWorkspace {
    size_t tile_floats, sched_off, tape_off, scratch_off, bytes;
    long long tiles, grid;
};
static TcWorkspace tc_workspace(const phnn_pack* pk, long long B, int T, int S) {
    TcWorkspace w;
    w.tiles = (B + 127) / 128;
    w.grid = w.tiles < pk->num_sms ? w.tiles : pk->num_sms;
    w.tile_floats = ws_persist_floats_per_tile(T, 128);
    w.sched_off = (size_t)w.tiles * w.tile_floats * sizeof(float);
    w.tape_off = (w.sched_off + sizeof(int) * (size_t)(w.tiles + 4) + 127) / 128 * 128;
    w.scratch_off = w.tape_off + (size_t)w.grid * T * S * 3 * pk->h * 128 * sizeof(float);
    w.bytes = w.scratch_off + (size_t)w.grid * tc_scratch_floats_per_cta(pk->n, T, S) * sizeof(float);
    return w;
}
// the FP32-FMA kernel (32-instance tiles) and the latency kernel (single-instance tiles) keep the stage states per tile
static size_t fp32_workspace_bytes(const phnn_pack* pk, long long B, int T, int S) {
    const size_t inst = (size_t)((B + 31) / 32) * 32;
    return (inst * ((size_t)T * S * pk->n + 3 * (size_t)T + 1) * sizeof(float) + 127) / 128 * 128;
}

template <class SH>
static int launch_tc_shape(const phnn_pack* pk, KParams& P, cudaStream_t stream) {
    static_assert((3 * SH::HID * 128 * 4) % 65536 == 0, "tape prefetch granularity");
    const long long tiles = (P.B + SH::TM - 1) / SH::TM;
    using SH16 = Tc16Shape<SH::MK, SH::NS, SH::HID>;
    const bool gen2 = pk->tc_mode >= 4;  // FP16 operands, A in tensor memory (phnn_tc16_kernel.cuh): 4 = hi/lo, three products; 5 = one product
    P.tc_split = pk->tc_mode;  // 1 plain TF32, 2 TF32 + BF16 correction product, 3 3xTF32
    if (pk->tc_mode == 2) P.wtc = pk->d_wtc2;
    P.ng = 1;
#ifdef PHNN_TC_PROFILE
    P.dbg = g_dbg;
#else
    P.dbg = nullptr;
#endif
    P.tiles = tiles;
    P.sched = nullptr;
    P.tape = nullptr;
    P.scratch = nullptr;
    long long grid = tiles;
#ifdef PHNN_TC_PROFILE
    static const bool no_steal = getenv("PHNN_NO_STEAL") != nullptr;  // experiment switch (profiling builds only)
#else
    constexpr bool no_steal = false;
#endif
    const bool adjoint = (P.mode == MODE_SOLVE && P.iters > 0) || (P.mode == MODE_COSTGRAD && P.want_grad);
    if (adjoint) {
        // the forward sweep tapes its activations into a per-CTA region: at most one CTA per SM, each running
        // its tiles one after the other (static stride) or pulling (tile, iteration) units (work-stealing solve)
        if (!P.ws) return fail(PHNN_E_WORKSPACE, "tcgen05 path: a job with an adjoint needs the workspace");
        const TcWorkspace w = tc_workspace(pk, P.B, P.T, P.S);
        grid = w.grid;
        P.tape = reinterpret_cast<float*>(reinterpret_cast<char*>(P.ws) + w.tape_off);
        P.scratch = reinterpret_cast<float*>(reinterpret_cast<char*>(P.ws) + w.scratch_off);
        if (P.mode == MODE_SOLVE && !no_steal) {
            P.sched = reinterpret_cast<int*>(reinterpret_cast<char*>(P.ws) + w.sched_off);
            CUDA_TRY(cudaMemsetAsync(P.sched, 0, sizeof(int) * (size_t)(tiles + 1), stream));
        }
    }
    // work-stealing solve jobs with an even number of tiles on at least two SMs: CTA pairs (clusters of two along x)
    if (gen2 && pk->tc_mode == 4 && pk->tc_pair && P.sched && tiles % 2 == 0 && grid >= 2) {
        using SHP = Tc16Shape<SH::MK, SH::NS, SH::HID, false, false, true>;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(grid & ~1LL));
        cfg.blockDim = dim3(SHP::THREADS);
        cfg.dynamicSmemBytes = SHP::SMEM_BYTES;
        cfg.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        CUDA_TRY(cudaLaunchKernelEx(&cfg, phnn_tc16_kernel<SH::MK, SH::NS, SH::HID, false, false, true>, P));
    } else if (gen2 && pk->tc_mode == 5)
        phnn_tc16_kernel<SH::MK, SH::NS, SH::HID, true><<<(unsigned)grid, SH16::THREADS, SH16::SMEM_BYTES, stream>>>(P);
    else if (gen2)
        phnn_tc16_kernel<SH::MK, SH::NS, SH::HID><<<(unsigned)grid, SH16::THREADS, SH16::SMEM_BYTES, stream>>>(P);
    else
        phnn_tc_kernel<SH::MK, SH::NS, SH::HID><<<(unsigned)grid, SH::THREADS, SH::SMEM_BYTES, stream>>>(P);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// forward-only shapes on the second-generation tcgen05 kernel: one CTA per 128-instance tile, nothing taped
template <class SH16>
static int launch_tc16_fwd_shape(const phnn_pack* pk, KParams& P, cudaStream_t stream) {
    P.tc_split = 4;
    P.ng = 1;
    P.dbg = nullptr;
    P.sched = nullptr;
    P.tape = nullptr;
    P.scratch = nullptr;
    // 64-instance (sparse) tiles while they still fit one per SM: twice the SMs for the same job, half the instruction
    // stream per tile; 128-instance tiles beyond that
    using SP = Tc16Shape<SH16::MK, SH16::NS, SH16::HID, false, true>;
    const long long tiles64 = (P.B + SP::TM - 1) / SP::TM;
    if (pk->tc_fwd_sparse && tiles64 <= pk->num_sms) {
        P.tiles = tiles64;
        phnn_tc16_kernel<SH16::MK, SH16::NS, SH16::HID, false, true><<<(unsigned)P.tiles, SP::THREADS, SP::SMEM_BYTES, stream>>>(P);
        CUDA_TRY(cudaGetLastError());
        return 0;
    }
    P.tiles = (P.B + SH16::TM - 1) / SH16::TM;
    phnn_tc16_kernel<SH16::MK, SH16::NS, SH16::HID><<<(unsigned)P.tiles, SH16::THREADS, SH16::SMEM_BYTES, stream>>>(P);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
static bool forward_only_job(const KParams& P) {
    return P.mode == MODE_FORWARD || P.mode == MODE_ROLLOUT || (P.mode == MODE_COSTGRAD && !P.want_grad);
}

static int launch(const phnn_pack* pk, KParams& P, void* stream) {
    if (P.B <= 0) return 0;
    int prev = 0;
    CUDA_TRY(cudaGetDevice(&prev));
    if (prev != pk->device) CUDA_TRY(cudaSetDevice(pk->device));
    int rc = fail(PHNN_E_UNSUPPORTED, "no kernel instantiation");
    // forward jobs of the n = 2 models: tensor cores from a few instances per SM upwards
    if (pk->tc_mode == 4 && !pk->d_wtc && pk->d_wtc16 && pk->tc_fwd_min_batch > 0 && P.B >= pk->tc_fwd_min_batch &&
        forward_only_job(P) && has_tc16_fwd_shape(pk->mk, pk->n, pk->h)) {
#define X(MK, NS, HID) \
    if (pk->mk == MK && pk->n == NS && pk->h == HID) rc = launch_tc16_fwd_shape<Tc16Shape<MK, NS, HID>>(pk, P, (cudaStream_t)stream);
        PHNN_TC16_FWD_SHAPES(X)
#undef X
        if (prev != pk->device) cudaSetDevice(prev);
        return rc;
    }
    // small batches: one CTA per instance (latency path)
    if (pk->lat_max_batch > 0 && P.B <= pk->lat_max_batch && has_lat_shape(pk->mk, pk->n, pk->h)) {
#define X(MK, NS, HID) \
    if (pk->mk == MK && pk->n == NS && pk->h == HID) rc = launch_lat_shape<LatShape<MK, NS, HID>>(pk, P, (cudaStream_t)stream);
        PHNN_LAT_SHAPES(X)
#undef X
        if (prev != pk->device) cudaSetDevice(prev);
        return rc;
    }
    // tcgen05 path only when batch x hidden width make the layer products a real dense contraction
    // (phnn_vjp carries no workspace for the tcgen05 kernel's activation stash: it stays on the FP32 kernel)
    if (pk->tc_mode != 0 && pk->d_wtc && P.B >= pk->tc_min_batch && P.mode != MODE_VJP) {
#define X(MK, NS, HID) \
    if (pk->mk == MK && pk->n == NS && pk->h == HID) rc = launch_tc_shape<TcShape<MK, NS, HID>>(pk, P, (cudaStream_t)stream);
        PHNN_TC_SHAPES(X)
#undef X
        if (prev != pk->device) cudaSetDevice(prev);
        return rc;
    }
#define X(MK, NS, HID) \
    if (pk->mk == MK && pk->n == NS && pk->h == HID) rc = launch_shape<Shape<MK, NS, HID>>(pk, P, (cudaStream_t)stream);
    PHNN_SHAPES(X)
#undef X
    if (prev != pk->device) cudaSetDevice(prev);
    return rc;
}

static int set_integrator(KParams& P, int integrator, double dt) {
    if (integrator != PHNN_EULER && integrator != PHNN_RK4)
        return fail(PHNN_E_INTEGRATOR, "Unknown integrator: %d", integrator);
    P.S = integrator == PHNN_RK4 ? 4 : 1;
    // the reference multiplies Python doubles dt, dt/2, dt/6.0 into float32 tensors
    P.dt = (float)dt;
    P.dt2 = (float)(dt / 2);
    P.dt3 = (float)(dt / 3.0);
    P.dt6 = (float)(dt / 6.0);
    return 0;
}

static int set_cost(const phnn_pack* pk, KParams& P, const phnn_cost_desc* c) {
    if (!c || !c->Q || !c->R || !c->x_target) return fail(PHNN_E_ARG, "cost: Q, R, x_target must be set");
    const int n = pk->n;
    for (int a = 0; a < n; ++a)
        for (int b = 0; b < n; ++b) {
            P.Q[a * n + b] = c->Q[a * n + b];
            P.Qs[a * n + b] = c->Q[a * n + b] + c->Q[b * n + a];
        }
    P.Rw = c->R[0];
    for (int a = 0; a < n; ++a) P.xt[a] = c->x_target[a];
    P.has_ub = c->has_u_bounds;
    P.umin = c->u_min;
    P.umax = c->u_max;
    P.has_xmin = c->x_min != nullptr;
    P.has_xmax = c->x_max != nullptr;
    for (int a = 0; a < n; ++a) {
        P.xmin[a] = c->x_min ? c->x_min[a] : 0.f;
        P.xmax[a] = c->x_max ? c->x_max[a] : 0.f;
    }
    P.bw = c->barrier_weight;
    return 0;
}

extern "C" int phnn_forward(const phnn_pack* pk, const float* x, const float* u, float* dx, float* H, long B,
                            void* stream) {
    if (pk && B == 0) return 0;
    if (!pk || !x || !u || !dx || !H || B < 0) return fail(PHNN_E_ARG, "phnn_forward: bad argument");
    KParams P = pk->base;
    P.mode = MODE_FORWARD; P.B = B; P.T = 1; P.S = 1; P.iters = 1;
    P.x0 = x; P.uin = u; P.out0 = dx; P.out1 = H;
    return launch(pk, P, stream);
}

extern "C" int phnn_vjp(const phnn_pack* pk, const float* x, const float* u, const float* v, float* xbar,
                        float* ubar, long B, void* stream) {
    if (pk && B == 0) return 0;
    if (!pk || !x || !u || !v || !xbar || !ubar || B < 0) return fail(PHNN_E_ARG, "phnn_vjp: bad argument");
    KParams P = pk->base;
    P.mode = MODE_VJP; P.B = B; P.T = 1; P.S = 1; P.iters = 1;
    P.x0 = x; P.uin = u; P.vin = v; P.out0 = xbar; P.out1 = ubar;
    return launch(pk, P, stream);
}

extern "C" int phnn_rollout(const phnn_pack* pk, const float* x0, const float* U, float* traj, float* energies,
                            long B, int T, double dt, int integrator, int energy_mode, void* stream) {
    if (integrator != PHNN_EULER && integrator != PHNN_RK4)
        return fail(PHNN_E_INTEGRATOR, "Unknown integrator: %d", integrator);
    if (pk && B == 0) return 0;
    if (!pk || !x0 || (!U && T > 0) || B < 0 || T < 0) return fail(PHNN_E_ARG, "phnn_rollout: bad argument");
    if (energies && energy_mode != 1 && energy_mode != 2) return fail(PHNN_E_ARG, "energy_mode must be 1 or 2");
    KParams P = pk->base;
    int rc = set_integrator(P, integrator, dt);
    if (rc) return rc;
    P.mode = MODE_ROLLOUT; P.B = B; P.T = T; P.iters = 1;
    P.energy_mode = energies ? energy_mode : 0;
    P.x0 = x0; P.uin = U; P.out0 = traj; P.out1 = energies;
    return launch(pk, P, stream);
}

extern "C" size_t phnn_workspace_bytes(const phnn_pack* pk, long B, int T, int integrator) {
    if (!pk || B <= 0 || T <= 0) return 0;
    const int S = integrator == PHNN_RK4 ? 4 : 1;
    // the tcgen05 kernels' layout when this batch is routed there, otherwise the per-tile layout of the FP32-FMA /
    // latency kernels
    const bool lat = pk->lat_max_batch > 0 && B <= pk->lat_max_batch && has_lat_shape(pk->mk, pk->n, pk->h);
    const bool tc = !lat && pk->tc_mode != 0 && pk->d_wtc && B >= pk->tc_min_batch;
    return tc ? tc_workspace(pk, B, T, S).bytes : fp32_workspace_bytes(pk, B, T, S);
}

extern "C" int phnn_cost_grad(const phnn_pack* pk, const phnn_cost_desc* cd, const float* x0, const float* U,
                              float* cost, float* dJdU, float* traj, long B, int T, double dt, int integrator,
                              void* workspace, size_t workspace_bytes, void* stream) {
    if (pk && B == 0) return 0;
    if (!pk || !x0 || !U || !cost || B < 0 || T <= 0) return fail(PHNN_E_ARG, "phnn_cost_grad: bad argument");
    KParams P = pk->base;
    int rc = set_integrator(P, integrator, dt);
    if (rc) return rc;
    rc = set_cost(pk, P, cd);
    if (rc) return rc;
    if (dJdU && (!workspace || workspace_bytes < phnn_workspace_bytes(pk, B, T, integrator)))
        return fail(PHNN_E_WORKSPACE, "phnn_cost_grad: workspace too small (%zu < %zu)", workspace_bytes,
                    phnn_workspace_bytes(pk, B, T, integrator));
    P.mode = MODE_COSTGRAD; P.B = B; P.T = T; P.iters = 1; P.want_grad = dJdU != nullptr;
    P.x0 = x0; P.uin = U; P.cost = cost; P.dJdU = dJdU; P.out0 = traj; P.ws = (float*)workspace;
    return launch(pk, P, stream);
}

extern "C" int phnn_mpc_solve_peer(const phnn_pack* pk, const phnn_cost_desc* cd, const float* x0, float* U_inout,
                                   float* cost_hist, float* best_cost, long B, int T, double dt, int integrator, double lr,
                                   double beta1, double beta2, double eps, int iters, int return_mode, void* workspace,
                                   size_t workspace_bytes, const phnn_peer_desc* peers, void* stream) {
    if (pk && B == 0) return 0;
    if (!pk || !x0 || !U_inout || B < 0 || T <= 0 || iters < 0)
        return fail(PHNN_E_ARG, "phnn_mpc_solve: bad argument");
    if (return_mode != 0 && return_mode != 1) return fail(PHNN_E_ARG, "return_mode must be 0 (last) or 1 (best)");
    KParams P = pk->base;
    int rc = set_integrator(P, integrator, dt);
    if (rc) return rc;
    rc = set_cost(pk, P, cd);
    if (rc) return rc;
    if (!workspace || workspace_bytes < phnn_workspace_bytes(pk, B, T, integrator))
        return fail(PHNN_E_WORKSPACE, "phnn_mpc_solve: workspace too small (%zu < %zu)", workspace_bytes,
                    phnn_workspace_bytes(pk, B, T, integrator));
    P.mode = MODE_SOLVE; P.B = B; P.T = T; P.iters = iters; P.return_mode = return_mode; P.want_grad = 1;
    P.lr = lr; P.beta1 = beta1; P.beta2 = beta2; P.eps = eps;
    P.x0 = x0; P.U = U_inout; P.cost_hist = cost_hist; P.cost = best_cost; P.ws = (float*)workspace;
    P.npeer = 0;
    if (peers) {
        if (peers->n < 1 || peers->n > PHNN_MAX_PEERS || peers->offset < 0)
            return fail(PHNN_E_ARG, "phnn_mpc_solve_peer: 1..%d peers, offset >= 0", PHNN_MAX_PEERS);
        const bool lat = pk->lat_max_batch > 0 && B <= pk->lat_max_batch && has_lat_shape(pk->mk, pk->n, pk->h);
        const bool tc = !lat && pk->tc_mode != 0 && pk->d_wtc && B >= pk->tc_min_batch;
        if (!tc || iters == 0)
            return fail(PHNN_E_UNSUPPORTED, "phnn_mpc_solve_peer: the fused exchange needs a batch routed to a tcgen05 kernel and iters > 0");
        P.npeer = peers->n;
        P.peer_off = peers->offset;
        for (int r = 0; r < peers->n; ++r) {
            if (!peers->U[r]) return fail(PHNN_E_ARG, "phnn_mpc_solve_peer: peer %d has no result buffer", r);
            P.peerU[r] = peers->U[r];
            P.peerC[r] = best_cost ? peers->cost[r] : nullptr;
        }
    }
    return launch(pk, P, stream);
}

extern "C" int phnn_mpc_solve(const phnn_pack* pk, const phnn_cost_desc* cd, const float* x0, float* U_inout,
                              float* cost_hist, float* best_cost, long B, int T, double dt, int integrator, double lr,
                              double beta1, double beta2, double eps, int iters, int return_mode, void* workspace,
                              size_t workspace_bytes, void* stream) {
    return phnn_mpc_solve_peer(pk, cd, x0, U_inout, cost_hist, best_cost, B, T, dt, integrator, lr, beta1, beta2, eps, iters,
                               return_mode, workspace, workspace_bytes, nullptr, stream);
}

// =========================================================================================
// Training mode: dL/dtheta of a loss of the rolled-out trajectory
// =========================================================================================
// C[i][j] (ldc) = alpha * sum_r A[r*lda + i] * B[r*ldb + j]  (B == nullptr: column sums of A, q = 1), i < p, j < q.
// One CTA per 32 x 32 tile of C and per slice of the rows; slices meet with atomicAdd (C is zeroed first).
__global__ void __launch_bounds__(256) atb_kernel(const float* __restrict__ A, int lda, int p, const float* __restrict__ B, int ldb, int q,
                                                  long long rows, float* __restrict__ C, int ldc, float alpha) {
    __shared__ float sA[32][33], sB[32][33];
    const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // ty 0..7
    const long long per = (rows + gridDim.z - 1) / gridDim.z;
    const long long r0 = (long long)blockIdx.z * per, r1 = (r0 + per < rows) ? r0 + per : rows;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};  // C[i0 + ty + 8 e][j0 + tx]
    for (long long rb = r0; rb < r1; rb += 32) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const long long r = rb + ty + 8 * e;
            sA[ty + 8 * e][tx] = (r < r1 && i0 + tx < p) ? A[r * lda + i0 + tx] : 0.f;
            sB[ty + 8 * e][tx] = (r < r1 && j0 + tx < q) ? (B ? B[r * ldb + j0 + tx] : 1.f) : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int r = 0; r < 32; ++r) {
            const float b = sB[r][tx];
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[e] = fmaf(sA[r][ty + 8 * e], b, acc[e]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int i = i0 + ty + 8 * e, j = j0 + tx;
        if (i < p && j < q) atomicAdd(&C[(size_t)i * ldc + j], alpha * acc[e]);
    }
}

static int atb(const float* A, int lda, int p, const float* B, int ldb, int q, long long rows, float* C, int ldc, float alpha,
               cudaStream_t st) {
    if (!C || rows <= 0) return 0;
    long long slices = (rows + 2047) / 2048;
    if (slices > 64) slices = 64;
    dim3 grid((p + 31) / 32, (q + 31) / 32, (unsigned)slices);
    atb_kernel<<<grid, 256, 0, st>>>(A, lda, p, B, ldb, q, rows, C, ldc, alpha);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

struct VjpWorkspace {
    size_t state_bytes, emit_off, bytes;
    long long rows;
};
static VjpWorkspace vjp_workspace(const phnn_pack* pk, long long B, int T, int S) {
    VjpWorkspace w;
    w.rows = B * T * S;
    w.state_bytes = fp32_workspace_bytes(pk, B, T, S);
    w.emit_off = w.state_bytes;
    w.bytes = w.emit_off + ((size_t)EMIT_NARR * w.rows * pk->h + (size_t)w.rows * EMIT_SMALL) * sizeof(float);
    return w;
}

extern "C" size_t phnn_rollout_vjp_workspace_bytes(const phnn_pack* pk, long B, int T, int integrator) {
    if (!pk || B <= 0 || T <= 0) return 0;
    return vjp_workspace(pk, B, T, integrator == PHNN_RK4 ? 4 : 1).bytes;
}

extern "C" int phnn_rollout_vjp(const phnn_pack* pk, const float* x0, const float* U, const float* gtraj, float* dx0, float* dU,
                                const phnn_param_grads* gr, long B, int T, double dt, int integrator, void* workspace,
                                size_t workspace_bytes, void* stream) {
    if (pk && B == 0) return 0;
    if (!pk || !x0 || !U || !gtraj || B < 0 || T <= 0) return fail(PHNN_E_ARG, "phnn_rollout_vjp: bad argument");
    if (!has_lat_shape(pk->mk, pk->n, pk->h))
        return fail(PHNN_E_UNSUPPORTED, "phnn_rollout_vjp: the training mode is built for the latency-kernel shapes (hidden width <= 128)");
    KParams P = pk->base;
    int rc = set_integrator(P, integrator, dt);
    if (rc) return rc;
    const VjpWorkspace w = vjp_workspace(pk, B, T, P.S);
    if (!workspace || workspace_bytes < w.bytes)
        return fail(PHNN_E_WORKSPACE, "phnn_rollout_vjp: workspace too small (%zu < %zu)", workspace_bytes, w.bytes);
    cudaStream_t st = (cudaStream_t)stream;
    int prev = 0;
    CUDA_TRY(cudaGetDevice(&prev));
    if (prev != pk->device) CUDA_TRY(cudaSetDevice(pk->device));
    P.mode = MODE_PARAMGRAD; P.B = B; P.T = T; P.iters = 1; P.want_grad = 1;
    P.has_ub = 0; P.has_xmin = 0; P.has_xmax = 0; P.Rw = 0.f;
    P.x0 = x0; P.uin = U; P.gtraj = gtraj; P.out0 = dx0; P.dJdU = dU; P.ws = (float*)workspace;
    P.emit = gr ? reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + w.emit_off) : nullptr;
    P.emit_rows = w.rows;
    rc = fail(PHNN_E_UNSUPPORTED, "no latency-kernel instantiation");
#define X(MK, NS, HID) \
    if (pk->mk == MK && pk->n == NS && pk->h == HID) rc = launch_lat_shape<LatShape<MK, NS, HID>>(pk, P, st);
    PHNN_LAT_SHAPES(X)
#undef X
    if (rc == 0 && gr) {
        const int h = pk->h, n = pk->n;
        const long long R = w.rows;
        const float* E = P.emit;
        auto arr = [&](int a) { return E + (size_t)a * R * h; };
        const float* sm = E + (size_t)EMIT_NARR * R * h;
        struct Out { float* p; size_t n; } outs[] = {{gr->W1, (size_t)h * n}, {gr->b1, (size_t)h}, {gr->W2, (size_t)h * h}, {gr->b2, (size_t)h},
            {gr->W3, (size_t)h}, {gr->Wr1, (size_t)h * n}, {gr->br1, (size_t)h}, {gr->Wr2, (size_t)n * n * h}, {gr->br2, (size_t)n * n},
            {gr->Wg1, (size_t)h * n}, {gr->bg1, (size_t)h}, {gr->Wg2, (size_t)n * h}, {gr->bg2, (size_t)n}, {gr->J, (size_t)n * n},
            {gr->r_diag, (size_t)n}};
        for (auto& o : outs)
            if (o.p && rc == 0) rc = (int)cudaMemsetAsync(o.p, 0, o.n * sizeof(float), st);
        const int SW = EMIT_SMALL;
        // H_net
        if (!rc) rc = atb(arr(4), h, h, sm + 4, SW, n, R, gr->W1, n, 1.f, st);   // delta1 (x) w
        if (!rc) rc = atb(arr(5), h, h, sm + 0, SW, n, R, gr->W1, n, 1.f, st);   // zbar1 (x) z
        if (!rc) rc = atb(arr(5), h, h, nullptr, 0, 1, R, gr->b1, 1, 1.f, st);
        if (!rc) rc = atb(arr(2), h, h, arr(1), h, h, R, gr->W2, h, 1.f, st);    // delta2 (x) da1
        if (!rc) rc = atb(arr(3), h, h, arr(0), h, h, R, gr->W2, h, 1.f, st);    // e2 (x) a1
        if (!rc) rc = atb(arr(3), h, h, nullptr, 0, 1, R, gr->b2, 1, 1.f, st);
        if (!rc) rc = atb(arr(6), h, h, nullptr, 0, 1, R, gr->W3, 1, 1.f, st);
        if (pk->mk != MK_CANON) {
            if (!rc) rc = atb(arr(8), h, h, sm + 16, SW, n, R, gr->Wr1, n, 1.f, st);         // dr (x) y
            if (!rc) rc = atb(arr(8), h, h, nullptr, 0, 1, R, gr->br1, 1, 1.f, st);
            if (!rc) rc = atb(sm + 20, SW, n * n, arr(7), h, h, R, gr->Wr2, h, 1.f, st);     // Rbar_raw (x) r1
            if (!rc) rc = atb(sm + 20, SW, n * n, nullptr, 0, 1, R, gr->br2, 1, 1.f, st);
            if (!rc) rc = atb(sm + 8, SW, n, sm + 12, SW, n, R, gr->J, n, 1.f, st);          // J enters as J - J^T: v g^T - g v^T
            if (!rc) rc = atb(sm + 12, SW, n, sm + 8, SW, n, R, gr->J, n, -1.f, st);
        } else {
            if (!rc && gr->r_diag) rc = atb(sm + 36, SW, 2, nullptr, 0, 1, R, gr->r_diag + 2, 1, -1.f, st);  // rows 2,3 of diag r
        }
        if (pk->mk == MK_PHNN_GNET) {
            if (!rc) rc = atb(arr(10), h, h, sm + 16, SW, n, R, gr->Wg1, n, 1.f, st);
            if (!rc) rc = atb(arr(10), h, h, nullptr, 0, 1, R, gr->bg1, 1, 1.f, st);
            if (!rc) rc = atb(sm + 40, SW, n, arr(9), h, h, R, gr->Wg2, h, 1.f, st);
            if (!rc) rc = atb(sm + 40, SW, n, nullptr, 0, 1, R, gr->bg2, 1, 1.f, st);
        }
        if (rc > 0) cuda_fail((cudaError_t)rc, "phnn_rollout_vjp reductions");
    }
    if (prev != pk->device) cudaSetDevice(prev);
    return rc;
}

// ---- result buffers shared between the ranks of one node (CUDA IPC) -------------------------------------
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev); else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
extern "C" int phnn_peer_alloc(size_t bytes, int device, void** dptr, void* handle64) {
    if (!dptr || !handle64 || bytes == 0) return fail(PHNN_E_ARG, "phnn_peer_alloc: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    DeviceGuard g(device);
    CUDA_TRY(cudaMalloc(dptr, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, *dptr);
    if (e != cudaSuccess) {
        cudaFree(*dptr);
        *dptr = nullptr;
        return cuda_fail(e, "cudaIpcGetMemHandle");
    }
    memcpy(handle64, &h, 64);
    return 0;
}
extern "C" int phnn_peer_open(const void* handle64, int device, void** dptr) {
    if (!dptr || !handle64) return fail(PHNN_E_ARG, "phnn_peer_open: bad argument");
    DeviceGuard g(device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    CUDA_TRY(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
extern "C" int phnn_peer_close(void* dptr, int device) {
    if (!dptr) return 0;
    DeviceGuard g(device);
    CUDA_TRY(cudaIpcCloseMemHandle(dptr));
    return 0;
}
extern "C" int phnn_peer_free(void* dptr, int device) {
    if (!dptr) return 0;
    DeviceGuard g(device);
    CUDA_TRY(cudaFree(dptr));
    return 0;
}

// ---- measurement utility: sustained FP32-FMA rate of this GPU (roofline denominator for the
// FP32 path; bench.py times it with CUDA events) --------------------------------------------
__global__ void __launch_bounds__(256) ffma_probe_kernel(float* out, int iters) {
    float a[16];
    const float m = 1.0f + 1e-7f * (float)threadIdx.x, c = 1e-9f;
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = (float)(i + threadIdx.x);
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], m, c);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 12345.678f) out[0] = s;  // keeps the chain alive without a real store
}

extern "C" int phnn_ffma_probe(float* d_out, int iters, int blocks, void* stream, double* flops) {
    if (!d_out || iters <= 0 || blocks <= 0) return fail(PHNN_E_ARG, "phnn_ffma_probe: bad argument");
    ffma_probe_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_out, iters);
    CUDA_TRY(cudaGetLastError());
    if (flops) *flops = 2.0 * 16.0 * 8.0 * (double)iters * 256.0 * (double)blocks;
    return 0;
}

// ---- measurement utility: TF32 tcgen05.mma issue rate of this GPU (128x256x8 MMAs back to back on
// resident operands, all SMs); bench.py uses it as the tensor-pipe denominator of the 3xTF32 kernel ----
__global__ void __launch_bounds__(128, 1) tf32_probe_kernel(int n_mma, float* sink) {
    uint64_t* bars = reinterpret_cast<uint64_t*>(phnn_smem);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(phnn_smem + 128);
    float* A = reinterpret_cast<float*>(phnn_smem + 1024);  // 128 x 32 tf32 (16 KB) then 256 x 32 (32 KB)
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) A[i] = 0.001f * (float)(i & 255);
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tmem_ptr;
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a = smem_u32(phnn_smem + 1024), b = a + 16384;
        uint32_t ph = 0;
        for (int i = 0; i < n_mma; i += 64) {
            for (int j = 0; j < 64; ++j)
                umma_tf32(tbase + (j & 1) * 256, umma_desc_sw128(a + (j & 3) * 32), umma_desc_sw128(b + (j & 3) * 32), idesc, 1u);
            umma_commit(&bars[0]);
            mbar_wait(&bars[0], ph);
            ph ^= 1;
        }
        if (sink && n_mma < 0) sink[0] = 1.f;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512u) : "memory");
}

extern "C" int phnn_tf32_probe(float* d_out, int n_mma, int blocks, void* stream, double* flops) {
    if (!d_out || n_mma <= 0 || blocks <= 0) return fail(PHNN_E_ARG, "phnn_tf32_probe: bad argument");
    const int smem = 1024 + 16384 + 32768;
    CUDA_TRY(cudaFuncSetAttribute(tf32_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    n_mma = (n_mma + 63) / 64 * 64;
    tf32_probe_kernel<<<blocks, 128, smem, (cudaStream_t)stream>>>(n_mma, d_out);
    CUDA_TRY(cudaGetLastError());
    if (flops) *flops = 2.0 * 128.0 * 256.0 * 8.0 * (double)n_mma * (double)blocks;
    return 0;
}

// =========================================================================================
// Batched closed loop on the device: the plant step, the bookkeeping
// of the reference's driver loop and the warm-start shift, so thousands of closed-loop episodes
// advance without a host round trip per step.
// =========================================================================================
struct PlantArgs {
    phnn_episode ep;
    const float* u;
    long long u_stride;
    int step;
    double dt, min_duration;
    double target[4], tol[4];
    long long B;
};

// CartPoleSimulator.step (src/cartpole_simulator.py:63-112, float64, explicit Euler, termination
// |x| > 10 or |theta| > 0.5) preceded by the stability bookkeeping of run_mpc_control
// (scripts/run_cartpole_mpc.py:138-159), one thread per plant.
__global__ void plant_step_kernel(const PlantArgs a) {
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    double* s = a.ep.state + 4 * b;
    double x = s[0], th = s[1], xd = s[2], thd = s[3];
    const bool running = a.ep.done_step[b] < 0;
    const float uf = running ? a.u[b * a.u_stride] : 0.f;
    if (a.ep.controls) a.ep.controls[b * a.ep.steps + a.step] = uf;
    if (running) {
        // stability bookkeeping on the state the controller just saw
        const double st[4] = {x, th, xd, thd};
        bool within = true;
#pragma unroll
        for (int i = 0; i < 4; ++i) within = within && (fabs(st[i] - a.target[i]) <= a.tol[i]);
        if (within) {
            if (a.ep.stable_start[b] < 0) a.ep.stable_start[b] = a.step;
            const double dur = (double)(a.step - a.ep.stable_start[b] + 1) * a.dt;
            a.ep.stable_duration[b] = (float)dur;
            if (dur >= a.min_duration) a.ep.stability_achieved[b] = 1;
        } else {
            a.ep.stable_start[b] = -1;
            a.ep.stable_duration[b] = 0.f;
        }
        const double gravity = 9.8, masscart = 1.0, masspole = 0.1, length = 0.5;
        const double polemass_length = masspole * length, total_mass = masspole + masscart;
        const double force = (double)uf;
        const double c = cos(th), sn = sin(th);
        const double temp = (force + polemass_length * thd * thd * sn) / total_mass;
        const double thacc = (gravity * sn - c * temp) / (length * (4.0 / 3.0 - masspole * c * c / total_mass));
        const double xacc = temp - polemass_length * thacc * c / total_mass;
        x = x + a.dt * xd;
        th = th + a.dt * thd;
        xd = xd + a.dt * xacc;
        thd = thd + a.dt * thacc;
        s[0] = x; s[1] = th; s[2] = xd; s[3] = thd;
        if (fabs(x) > 10.0 || fabs(th) > 0.5) a.ep.done_step[b] = a.step + 1;
    }
    if (a.ep.traj) {
        double* t = a.ep.traj + ((size_t)b * (a.ep.steps + 1) + a.step + 1) * 4;
        t[0] = x; t[1] = th; t[2] = xd; t[3] = thd;
    }
}

extern "C" int phnn_plant_step(const phnn_episode* ep, const float* u, long u_stride, int step, double dt,
                               const double* target, const double* tol, double min_duration, long B, void* stream) {
    if (!ep || !ep->state || !ep->done_step || !ep->stable_start || !ep->stable_duration || !ep->stability_achieved || !u ||
        !target || !tol || B < 0 || step < 0 || step >= ep->steps)
        return fail(PHNN_E_ARG, "phnn_plant_step: bad argument");
    if (B == 0) return 0;
    PlantArgs a;
    a.ep = *ep; a.u = u; a.u_stride = u_stride; a.step = step; a.dt = dt; a.min_duration = min_duration; a.B = B;
    for (int i = 0; i < 4; ++i) { a.target[i] = target[i]; a.tol[i] = tol[i]; }
    plant_step_kernel<<<(unsigned)((B + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// x0[b,:] = float32(state[b,:]) -- the cast the controllers apply to the simulator's float64 state
// (src/mpc_controller.py:160-161, src/mpc_controller_canonical.py:249); also writes traj[:,0,:] when step 0
__global__ void state_to_f32_kernel(const double* __restrict__ state, float* __restrict__ x0, double* traj0, int traj_stride,
                                    long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    x0[i] = (float)state[i];
    if (traj0) traj0[(i / 4) * (size_t)traj_stride + (i % 4)] = state[i];
}
extern "C" int phnn_state_to_f32(const double* state, float* x0, double* traj, int steps, long B, void* stream) {
    if (!state || !x0 || B < 0) return fail(PHNN_E_ARG, "phnn_state_to_f32: bad argument");
    if (B == 0) return 0;
    state_to_f32_kernel<<<(unsigned)((B * 4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(state, x0, traj, (steps + 1) * 4, B * 4);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// warm start of MPCControllerCanonical.control: previous plan shifted left by one, zero appended
// (src/mpc_controller_canonical.py:252-255)
__global__ void shift_controls_kernel(const float* __restrict__ U, float* __restrict__ out, int H, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int t = (int)(i % H);
    out[i] = (t + 1 < H) ? U[i + 1] : 0.f;
}
extern "C" int phnn_shift_controls(const float* U, float* out, long B, int H, void* stream) {
    if (!U || !out || U == out || B < 0 || H <= 0) return fail(PHNN_E_ARG, "phnn_shift_controls: bad argument (in-place not allowed)");
    if (B == 0) return 0;
    shift_controls_kernel<<<(unsigned)((B * H + 255) / 256), 256, 0, (cudaStream_t)stream>>>(U, out, H, (long long)B * H);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
