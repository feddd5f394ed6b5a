// tools/ts_probe.cu -- hardware checks for feeding operand A of tcgen05.mma from TENSOR MEMORY (round 2):
//  (1) tcgen05.st.16x256b.x4 / .16x128b.x4 write the same (lane, column) fragment their tcgen05.ld twins read
//      (checked by reading back with 32x32b loads);
//  (2) kind::tf32 with A in TMEM: A[m][k] sits at lane m, column a0 + k (32 tf32 = 32 columns per K-block of 32);
//  (3) kind::f16 (BF16) with A in TMEM: A[m][k] sits at lane m, column a0 + k/2, half k%2 (64 BF16 = 32 columns).
// B comes from shared memory in the K-major SWIZZLE_128B layout the production kernel uses.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ts_probe tools/ts_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__host__ __device__ inline int sw128_off(int r, int c) {  // byte offset of 32-bit column c of row r (128-byte rows)
    return (r >> 3) * 1024 + (r & 7) * 128 + ((((c >> 2) ^ (r & 7)) & 7) << 4) + (c & 3) * 4;
}
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}

constexpr int N = 64;  // UMMA N

// out: [0, 128*64) D of the tf32 TS product, [128*64, 2*128*64) D of the bf16 TS product, then 2 x 128 x 32 read-backs
__global__ void __launch_bounds__(128) ts_kernel(const float* Atf, const uint32_t* Abf, const unsigned char* Btf, const unsigned char* Bbf,
                                                 float* out, uint32_t* rb) {
    extern __shared__ __align__(1024) unsigned char sm[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm);
    uint32_t* tptr = reinterpret_cast<uint32_t*>(sm + 64);
    unsigned char* sB = sm + 1024;           // N x 128 B
    unsigned char* sB2 = sm + 1024 + N * 128;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < N * 128 / 4; i += 128) {
        reinterpret_cast<uint32_t*>(sB)[i] = reinterpret_cast<const uint32_t*>(Btf)[i];
        reinterpret_cast<uint32_t*>(sB2)[i] = reinterpret_cast<const uint32_t*>(Bbf)[i];
    }
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tptr)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = *tptr;
    // column map: [0,64) D tf32, [64,128) D bf16, [128,160) A tf32 (32 cols), [160,192) A bf16 (64 bf16 = 32 cols)
    const int g = lane >> 2, c = lane & 3;
    for (int half = 0; half < 2; ++half) {
        const uint32_t tl = tbase + ((uint32_t)(warp * 32 + half * 16) << 16);
        // --- A tf32 via 16x256b.x4: reg[4k + 2 rsel + e] = A[row0 + g + 8 rsel][8 k + 2 c + e]
        uint32_t r[16];
        for (int i = 0; i < 16; ++i) {
            const int k = i >> 2, rsel = (i >> 1) & 1, e = i & 1;
            const int row = warp * 32 + half * 16 + g + 8 * rsel, col = 8 * k + 2 * c + e;
            r[i] = __float_as_uint(Atf[row * 32 + col]);
        }
        asm volatile("tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(tl + 128),
                     "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
                     "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                     : "memory");
        // --- A bf16 (packed pairs; 32-bit column j holds elements 2j, 2j+1) via 16x128b.x8: reg[2k + rsel] = col 4 k + c
        uint32_t q[16];
        for (int i = 0; i < 16; ++i) {
            const int k = i >> 1, rsel = i & 1;
            const int row = warp * 32 + half * 16 + g + 8 * rsel, col = 4 * k + c;
            q[i] = Abf[row * 32 + col];
        }
        asm volatile("tcgen05.st.sync.aligned.16x128b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(tl + 160),
                     "r"(q[0]), "r"(q[1]), "r"(q[2]), "r"(q[3]), "r"(q[4]), "r"(q[5]), "r"(q[6]), "r"(q[7]), "r"(q[8]), "r"(q[9]), "r"(q[10]),
                     "r"(q[11]), "r"(q[12]), "r"(q[13]), "r"(q[14]), "r"(q[15])
                     : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // read back with 32x32b (thread = lane): checks the store fragments
    {
        const uint32_t tl = tbase + ((uint32_t)(warp * 32) << 16);
        for (int part = 0; part < 2; ++part)
            for (int c0 = 0; c0 < 32; c0 += 8) {
                uint32_t v[8];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                             : "r"(tl + 128 + part * 32 + c0));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                for (int i = 0; i < 8; ++i) rb[(part * 128 + warp * 32 + lane) * 32 + c0 + i] = v[i];
            }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t idesc16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t bt = smem_u32(sB), bb = smem_u32(sB2);
        for (int ks = 0; ks < 4; ++ks) {  // tf32: K = 8 per instruction, A columns 8 ks .. 8 ks + 7
            const uint32_t acc = ks ? 1u : 0u;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tbase + 0),
                         "r"(tbase + 128 + ks * 8), "l"(umma_desc_sw128(bt + ks * 32)), "r"(idesc), "r"(acc)
                         : "memory");
        }
        for (int ks = 0; ks < 4; ++ks) {  // bf16: K = 16 per instruction, A columns 8 ks .. 8 ks + 7 (16 packed elements)
            const uint32_t acc = ks ? 1u : 0u;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tbase + 64),
                         "r"(tbase + 160 + ks * 8), "l"(umma_desc_sw128(bb + ks * 32)), "r"(idesc16), "r"(acc)
                         : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
    mbar_wait(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
        const uint32_t tl = tbase + ((uint32_t)(warp * 32) << 16);
        for (int part = 0; part < 2; ++part)
            for (int c0 = 0; c0 < N; c0 += 8) {
                uint32_t v[8];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                             : "r"(tl + part * 64 + c0));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                for (int i = 0; i < 8; ++i) out[(part * 128 + warp * 32 + lane) * N + c0 + i] = __uint_as_float(v[i]);
            }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(256u) : "memory");
}

static uint16_t bf16_bits(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    return (uint16_t)(u >> 16);  // values used here are exactly representable
}

int main() {
    static float Atf[128 * 32], Btf[N * 32], Abf_f[128 * 64], Bbf_f[N * 64];
    static uint32_t Abf[128 * 32];
    static unsigned char Btf_sw[N * 128], Bbf_sw[N * 128];
    for (int m = 0; m < 128; ++m)
        for (int k = 0; k < 32; ++k) Atf[m * 32 + k] = (float)(((m * 7 + k * 3) % 11) - 5);
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < 32; ++k) Btf[n * 32 + k] = (float)(((n * 5 + k * 2) % 9) - 4);
    for (int m = 0; m < 128; ++m)
        for (int k = 0; k < 64; ++k) Abf_f[m * 64 + k] = (float)(((m * 3 + k * 5) % 13) - 6);
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < 64; ++k) Bbf_f[n * 64 + k] = (float)(((n * 11 + k) % 7) - 3);
    for (int m = 0; m < 128; ++m)
        for (int j = 0; j < 32; ++j) Abf[m * 32 + j] = (uint32_t)bf16_bits(Abf_f[m * 64 + 2 * j]) | ((uint32_t)bf16_bits(Abf_f[m * 64 + 2 * j + 1]) << 16);
    for (int n = 0; n < N; ++n)
        for (int c = 0; c < 32; ++c) {
            memcpy(&Btf_sw[sw128_off(n, c)], &Btf[n * 32 + c], 4);
            const uint32_t w = (uint32_t)bf16_bits(Bbf_f[n * 64 + 2 * c]) | ((uint32_t)bf16_bits(Bbf_f[n * 64 + 2 * c + 1]) << 16);
            memcpy(&Bbf_sw[sw128_off(n, c)], &w, 4);
        }
    float *dA, *dOut;
    uint32_t *dAbf, *dRb;
    unsigned char *dBt, *dBb;
    cudaMalloc(&dA, sizeof(Atf)); cudaMalloc(&dAbf, sizeof(Abf)); cudaMalloc(&dBt, sizeof(Btf_sw)); cudaMalloc(&dBb, sizeof(Bbf_sw));
    cudaMalloc(&dOut, 2 * 128 * N * 4); cudaMalloc(&dRb, 2 * 128 * 32 * 4);
    cudaMemcpy(dA, Atf, sizeof(Atf), cudaMemcpyHostToDevice); cudaMemcpy(dAbf, Abf, sizeof(Abf), cudaMemcpyHostToDevice);
    cudaMemcpy(dBt, Btf_sw, sizeof(Btf_sw), cudaMemcpyHostToDevice); cudaMemcpy(dBb, Bbf_sw, sizeof(Bbf_sw), cudaMemcpyHostToDevice);
    const int smem = 1024 + 2 * N * 128;
    ts_kernel<<<1, 128, smem>>>(dA, dAbf, dBt, dBb, dOut, dRb);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("ts_kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
    static float out[2 * 128 * N];
    static uint32_t rb[2 * 128 * 32];
    cudaMemcpy(out, dOut, sizeof(out), cudaMemcpyDeviceToHost);
    cudaMemcpy(rb, dRb, sizeof(rb), cudaMemcpyDeviceToHost);
    int bad0 = 0, bad1 = 0, bad2 = 0, bad3 = 0;
    for (int m = 0; m < 128; ++m)
        for (int c = 0; c < 32; ++c) {
            uint32_t want;
            memcpy(&want, &Atf[m * 32 + c], 4);
            if (rb[m * 32 + c] != want) ++bad0;
            if (rb[(128 + m) * 32 + c] != Abf[m * 32 + c]) ++bad1;
        }
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            double s = 0, s2 = 0;
            for (int k = 0; k < 32; ++k) s += (double)Atf[m * 32 + k] * Btf[n * 32 + k];
            for (int k = 0; k < 64; ++k) s2 += (double)Abf_f[m * 64 + k] * Bbf_f[n * 64 + k];
            if (out[m * N + n] != (float)s) { if (bad2 < 4) printf("  tf32 TS D[%d][%d] = %g, want %g\n", m, n, out[m * N + n], s); ++bad2; }
            if (out[(128 + m) * N + n] != (float)s2) { if (bad3 < 4) printf("  bf16 TS D[%d][%d] = %g, want %g\n", m, n, out[(128 + m) * N + n], s2); ++bad3; }
        }
    printf("tcgen05.st.16x256b.x4 fragment == ld fragment: %s (%d mismatches)\n", bad0 ? "WRONG" : "CONFIRMED", bad0);
    printf("tcgen05.st.16x128b.x8 fragment reg[2k+rsel] = (row g + 8 rsel, col 4k + c): %s (%d mismatches)\n", bad1 ? "WRONG" : "CONFIRMED", bad1);
    printf("kind::tf32 with A in TMEM (lane m, column a0 + k): %s (%d mismatches)\n", bad2 ? "WRONG" : "CONFIRMED", bad2);
    printf("kind::f16 BF16 with A in TMEM (lane m, column a0 + k/2, half k%%2): %s (%d mismatches)\n", bad3 ? "WRONG" : "CONFIRMED", bad3);
    return 0;
}
