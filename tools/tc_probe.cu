// tools/tc_probe.cu -- standalone probe of the tcgen05 building blocks the tensor-core path uses:
//   * A operand written by threads into a 128B-swizzled K-major shared-memory tile,
//   * B operand pre-swizzled in global memory and fetched with cp.async.bulk,
//   * tcgen05.mma kind::tf32 (M=128, N=256, K=8 per instruction), 3xTF32 split (hi*hi + lo*hi + hi*lo),
//   * accumulator in TMEM read back with tcgen05.ld 32x32b,
// checked against a double-precision CPU product, plus an issue-rate measurement.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_probe tools/tc_probe.cu && ./tc_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int M = 128, N = 256, K = 256, KB = 32;        // K-block = 32 tf32 = one 128-byte swizzle row
constexpr int NKB = K / KB;
constexpr int A_TILE = M * 128;                          // bytes of one A K-block (hi or lo)
constexpr int B_TILE = N * 128;                          // bytes of one B K-block (hi or lo)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// shared-memory matrix descriptor: K-major, 128B swizzle, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);         // start address
    d |= (uint64_t)0 << 16;                           // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                 // stride byte offset: 8 rows * 128 B
    d |= (uint64_t)1 << 46;                           // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                           // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float tf32_rn(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// byte offset of element (row r, col c in 0..31) inside a 128B-swizzled K-major tile
__host__ __device__ inline int sw128_off(int r, int c) {
    return (r >> 3) * 1024 + (r & 7) * 128 + ((((c >> 2) ^ (r & 7)) & 7) << 4) + (c & 3) * 4;
}

// smem: [0,1024) barriers + tmem ptr ; A slots: 2 x (hi,lo) ; B stages: 2 x (hi,lo)
constexpr int SMEM_BYTES = 1024 + 2 * 2 * A_TILE + 2 * 2 * B_TILE;

__global__ void __launch_bounds__(320, 1) probe(const float* __restrict__ Ag /*[M][K]*/, const unsigned char* __restrict__ Bsw /* [hi|lo][NKB][B_TILE] */,
                                               float* __restrict__ D /*[M][N]*/, int nsplit, int reps, long long* cycles) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);   // 0,1: A full ; 2,3: A empty ; 4,5: B full ; 6,7: B empty ; 8: acc full
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + 128);
    unsigned char* Asm = smem + 1024;
    unsigned char* Bsm = Asm + 4 * A_TILE;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 256); mbar_init(&bars[1], 256);
        mbar_init(&bars[2], 1); mbar_init(&bars[3], 1);
        mbar_init(&bars[4], 1); mbar_init(&bars[5], 1);
        mbar_init(&bars[6], 1); mbar_init(&bars[7], 1);
        mbar_init(&bars[8], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = *tmem_ptr;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);

    for (int rep = 0; rep < reps; ++rep) {
        if (warp < 8) {
            // ---- A writers: thread (row, half) writes 16 of the 32 columns of each K-block ----
            const int row = threadIdx.x & 127, half = threadIdx.x >> 7;
            for (int kb = 0; kb < NKB; ++kb) {
                const int it = rep * NKB + kb, slot = it & 1;
                mbar_wait(&bars[2 + slot], ((it >> 1) & 1) ^ 1);      // slot free
                unsigned char* hi = Asm + (slot * 2 + 0) * A_TILE;
                unsigned char* lo = Asm + (slot * 2 + 1) * A_TILE;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int c0 = half * 16 + q * 4;
                    const float4 a = *reinterpret_cast<const float4*>(Ag + (size_t)row * K + kb * KB + c0);
                    float4 h, l;
                    h.x = tf32_rn(a.x); h.y = tf32_rn(a.y); h.z = tf32_rn(a.z); h.w = tf32_rn(a.w);
                    l.x = a.x - h.x; l.y = a.y - h.y; l.z = a.z - h.z; l.w = a.w - h.w;
                    *reinterpret_cast<float4*>(hi + sw128_off(row, c0)) = h;
                    *reinterpret_cast<float4*>(lo + sw128_off(row, c0)) = l;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> async proxy (UMMA reads)
                mbar_arrive(&bars[slot]);
            }
        } else if (warp == 8) {
            // ---- MMA issuer ----
            if (lane == 0) {
                for (int kb = 0; kb < NKB; ++kb) {
                    const int it = rep * NKB + kb, slot = it & 1;
                    mbar_wait(&bars[slot], (it >> 1) & 1);            // A landed
                    mbar_wait(&bars[4 + slot], (it >> 1) & 1);        // B landed
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_hi = smem_u32(Asm + (slot * 2 + 0) * A_TILE), a_lo = smem_u32(Asm + (slot * 2 + 1) * A_TILE);
                    const uint32_t b_hi = smem_u32(Bsm + (slot * 2 + 0) * B_TILE), b_lo = smem_u32(Bsm + (slot * 2 + 1) * B_TILE);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint32_t o = ks * 32;   // 8 tf32 = 32 bytes along the swizzled row
                        umma_tf32(tbase, make_desc(a_hi + o), make_desc(b_hi + o), idesc, (kb | ks) ? 1u : 0u);
                        if (nsplit == 3) {
                            umma_tf32(tbase, make_desc(a_lo + o), make_desc(b_hi + o), idesc, 1u);
                            umma_tf32(tbase, make_desc(a_hi + o), make_desc(b_lo + o), idesc, 1u);
                        }
                    }
                    umma_commit(&bars[2 + slot]);   // frees the A slot when these MMAs retire
                    umma_commit(&bars[6 + slot]);   // frees the B stage
                }
                umma_commit(&bars[8]);              // accumulator complete
            }
        } else if (warp == 9) {
            // ---- TMA producer for B ----
            if (lane == 0) {
                for (int kb = 0; kb < NKB; ++kb) {
                    const int it = rep * NKB + kb, slot = it & 1;
                    mbar_wait(&bars[6 + slot], ((it >> 1) & 1) ^ 1);
                    mbar_expect_tx(&bars[4 + slot], 2 * B_TILE);
                    bulk_g2s(Bsm + (slot * 2 + 0) * B_TILE, Bsw + (size_t)(0 * NKB + kb) * B_TILE, B_TILE, &bars[4 + slot]);
                    bulk_g2s(Bsm + (slot * 2 + 1) * B_TILE, Bsw + (size_t)(1 * NKB + kb) * B_TILE, B_TILE, &bars[4 + slot]);
                }
            }
        }
        // everyone waits for the accumulator of this repetition (keeps reps serialised)
        mbar_wait(&bars[8], rep & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    // ---- epilogue: TMEM -> registers -> global ----
    if (warp < 8) {
        const int row = threadIdx.x & 127, half = threadIdx.x >> 7, q = warp & 3;
        for (int cb = 0; cb < N / 32; ++cb) {
            float v[16];
            const int col = cb * 32 + half * 16;
            tmem_ld16(tbase + ((uint32_t)(q * 32) << 16) + col, v);
#pragma unroll
            for (int i = 0; i < 16; ++i) D[(size_t)row * N + col + i] = v[i];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512) : "memory");
    (void)cycles;
}

// issue-rate probe: resident operands, many MMAs back to back
__global__ void __launch_bounds__(128, 1) rate(int n_mma, int nper, long long* cycles) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + 128);
    unsigned char* Asm = smem + 1024;
    unsigned char* Bsm = Asm + A_TILE;
    for (int i = threadIdx.x; i < (A_TILE + B_TILE) / 4; i += blockDim.x) reinterpret_cast<float*>(Asm)[i] = 0.001f * (i & 255);
    if (threadIdx.x == 0) { mbar_init(&bars[0], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = *tmem_ptr;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    if (threadIdx.x == 0) {
        const uint64_t ad = make_desc(smem_u32(Asm)), bd = make_desc(smem_u32(Bsm));
        long long t0 = clock64();
        int ph = 0;
        for (int i = 0; i < n_mma; i += nper) {
            for (int j = 0; j < nper; ++j) umma_tf32(tbase + (j & 1) * 256, ad + 2 * (j & 3), bd + 2 * (j & 3), idesc, 1u);
            umma_commit(&bars[0]);
            mbar_wait(&bars[0], ph);
            ph ^= 1;
        }
        long long t1 = clock64();
        if (blockIdx.x == 0) cycles[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512) : "memory");
}

int main() {
    std::vector<float> A((size_t)M * K), B((size_t)N * K);
    srand(1);
    for (auto& v : A) v = (float)rand() / RAND_MAX * 2 - 1;
    for (auto& v : B) v = ((float)rand() / RAND_MAX * 2 - 1) * 0.1f;
    // pre-swizzle B (hi/lo) per K-block
    std::vector<unsigned char> Bsw((size_t)2 * NKB * B_TILE);
    auto tf32 = [](float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x1000u; u &= 0xFFFFE000u; float y; memcpy(&y, &u, 4); return y; };
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) {
            float h = tf32(B[(size_t)n * K + k]), l = B[(size_t)n * K + k] - h;
            int kb = k / KB, c = k % KB;
            memcpy(&Bsw[(size_t)(0 * NKB + kb) * B_TILE + sw128_off(n, c)], &h, 4);
            memcpy(&Bsw[(size_t)(1 * NKB + kb) * B_TILE + sw128_off(n, c)], &l, 4);
        }
    float *dA, *dD; unsigned char* dB; long long* dC;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dD, (size_t)M * N * 4)); CK(cudaMalloc(&dB, Bsw.size())); CK(cudaMalloc(&dC, 64));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bsw.data(), Bsw.size(), cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    CK(cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    std::vector<double> ref((size_t)M * N);
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)A[(size_t)m * K + k] * (double)B[(size_t)n * K + k];
            ref[(size_t)m * N + n] = s;
        }
    for (int nsplit : {1, 3}) {
        CK(cudaMemset(dD, 0, (size_t)M * N * 4));
        probe<<<1, 320, SMEM_BYTES>>>(dA, dB, dD, nsplit, 1, dC);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        std::vector<float> D((size_t)M * N);
        CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
        double emax = 0, rmax = 0;
        for (size_t i = 0; i < D.size(); ++i) { emax = fmax(emax, fabs(D[i] - ref[i])); rmax = fmax(rmax, fabs(ref[i])); }
        printf("nsplit=%d  max abs err %.3e  (max |ref| %.3f, rel %.3e)  D[0]=%f ref=%f  D[last]=%f ref=%f\n", nsplit, emax, rmax, emax / rmax,
               D[0], ref[0], D.back(), ref.back());
    }
    // fp32 sequential reference error for scale
    {
        double emax = 0, rmax = 0;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n) {
                float s = 0;
                for (int k = 0; k < K; ++k) s = fmaf(A[(size_t)m * K + k], B[(size_t)n * K + k], s);
                emax = fmax(emax, fabs(s - ref[(size_t)m * N + n])); rmax = fmax(rmax, fabs(ref[(size_t)m * N + n]));
            }
        printf("fp32 fma chain: max abs err %.3e (rel %.3e)\n", emax, emax / rmax);
    }
    // multi-rep pipeline sanity + time
    {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        const int reps = 200;
        probe<<<148, 320, SMEM_BYTES>>>(dA, dB, dD, 3, reps, dC);
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        probe<<<148, 320, SMEM_BYTES>>>(dA, dB, dD, 3, reps, dC);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double macs = 148.0 * reps * 3.0 * M * N * K;
        printf("pipelined probe (A written by threads from global, B via bulk copy, 3xTF32): %.3f ms, %.1f TFLOP/s tf32 executed, %.1f TFLOP/s fp32-equivalent\n",
               ms, 2 * macs / ms / 1e9, 2 * macs / 3 / ms / 1e9);
        std::vector<float> D((size_t)M * N);
        CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
        double emax = 0;
        for (size_t i = 0; i < D.size(); ++i) emax = fmax(emax, fabs(D[i] - ref[i]));
        printf("   after %d reps x 148 CTAs: max abs err %.3e\n", reps, emax);
    }
    for (int nper : {4, 16, 64}) {
        const int n_mma = 4096;
        rate<<<148, 128, SMEM_BYTES>>>(n_mma, nper, dC);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        long long cyc; CK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost));
        printf("rate: %d MMAs (128x256x8 tf32) in batches of %d: %lld cycles -> %.1f cycles/MMA, %.0f MAC/clk/SM\n", n_mma, nper, cyc, (double)cyc / n_mma,
               (double)n_mma * M * N * 8 / cyc);
    }
    printf("done\n");
    return 0;
}
