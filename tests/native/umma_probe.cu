// Stand-alone hardware probe for the primitives mlp_tc.cu relies on (test infrastructure, not product):
// one CTA computes D[128,256] = A[128,64] * B[256,64]^T with tcgen05.mma from hand-swizzled shared-memory
// operands (A: SWIZZLE_128B, B: two 32-K SWIZZLE_64B chunks), reads D back with tcgen05.ld and the host
// compares with an exact integer reference.  Several layout hypotheses are tried so that one GPU run
// tells which encoding the hardware implements.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "../../msra_practice_project_b200/csrc/umma.cuh"

using namespace b2r::umma;

__device__ __forceinline__ void mma_any(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}

struct Variant { uint32_t a_layout, a_sbo, b_layout, b_sbo, use_bulk; };

__global__ void __launch_bounds__(128, 1) probe_kernel(const uint8_t* a_img, const uint8_t* b_img, Variant v, float* d_out) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t* gen = raw + (base - smem_u32(raw));
    const uint32_t a_s = base, b_s = base + 16384, bar = base + 16384 + 32768, slot = bar + 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar + 8, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(slot, 256);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    uint32_t tmem; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
    // A by plain stores, B by plain stores or by the bulk-copy engine
    for (int i = threadIdx.x; i < 16384 / 16; i += 128) reinterpret_cast<uint4*>(gen)[i] = reinterpret_cast<const uint4*>(a_img)[i];
    if (!v.use_bulk) {
        for (int i = threadIdx.x; i < 32768 / 16; i += 128) reinterpret_cast<uint4*>(gen + 16384)[i] = reinterpret_cast<const uint4*>(b_img)[i];
    } else if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar + 8, 32768);
        bulk_g2s(b_s, b_img, 16384, bar + 8);
        bulk_g2s(b_s + 16384, b_img + 16384, 16384, bar + 8);
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (v.use_bulk) mbar_wait(bar + 8, 0);
        tc_fence_after();
        const uint32_t idesc = make_idesc_bf16(128, 256);
        for (int c = 0; c < 2; ++c)
            for (int k = 0; k < 2; ++k) {
                uint64_t ad = make_desc(a_s + c * 64 + k * 32, 16, v.a_sbo, v.a_layout);
                uint64_t bd = make_desc(b_s + c * 16384 + k * 32, 16, v.b_sbo, v.b_layout);
                mma_bf16(tmem, ad, bd, idesc, (c | k) != 0);
            }
        mma_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    tc_fence_after();
    const int r = warp * 32 + lane;
    for (int j = 0; j < 8; ++j) {
        uint32_t x[32];
        tmem_ld32(tmem + ((uint32_t)warp << 21) + j * 32, x);
        tmem_ld_wait();
        for (int e = 0; e < 32; ++e) d_out[r * 256 + j * 32 + e] = __uint_as_float(x[e]);
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

static uint16_t bf16_bits(float f);

// ---- CTA-pair probe: cluster of 2, tcgen05.mma.cta_group::2, M=256 (128 rows per CTA), N=256 (128 B rows per CTA), K=64 ----
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe2_kernel(const uint8_t* a_img /*2 x 16 KB*/, const uint8_t* b_img /*2 x 16 KB*/, float* d_out /*[256,256]*/) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t* gen = raw + (base - smem_u32(raw));
    const uint32_t rank = cluster_ctarank();
    const uint32_t a_s = base, b_s = base + 16384, bar_full = base + 32768, bar_done = bar_full + 8, slot = bar_full + 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(bar_full, 2); mbar_init(bar_done, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc_2cta(slot, 256);
    for (int i = threadIdx.x; i < 16384 / 16; i += 128) {
        reinterpret_cast<uint4*>(gen)[i] = reinterpret_cast<const uint4*>(a_img + rank * 16384)[i];
        reinterpret_cast<uint4*>(gen + 16384)[i] = reinterpret_cast<const uint4*>(b_img + rank * 16384)[i];
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    uint32_t tmem; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
    // both CTAs report "operands ready" to the leader's barrier (local arrive in the leader, remote from the peer)
    if (threadIdx.x == 0) {
        if (rank == 0) mbar_arrive_release_cluster_local(bar_full);
        else mbar_arrive_cluster(mapa(bar_full, 0));
    }
    if (rank == 0 && threadIdx.x == 0) {
        mbar_wait_cluster(bar_full, 0);
        tc_fence_after();
        const uint32_t idesc = make_idesc_bf16(256, 256);
        for (int k = 0; k < 4; ++k)
            mma_bf16_2cta(tmem, make_desc(a_s + k * 32, 16, 1024, 2), make_desc(b_s + k * 32, 16, 1024, 2), idesc, k != 0);
        mma_commit_2cta(bar_done, (uint16_t)3);
    }
    __syncwarp();
    mbar_wait(bar_done, 0);
    tc_fence_after();
    const int r = warp * 32 + lane;
    for (int j = 0; j < 8; ++j) {
        uint32_t x[32];
        tmem_ld32(tmem + ((uint32_t)warp << 21) + j * 32, x);
        tmem_ld_wait();
        for (int e = 0; e < 32; ++e) d_out[(rank * 128 + r) * 256 + j * 32 + e] = __uint_as_float(x[e]);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) tmem_dealloc_2cta(tmem, 256);
}

// ---- tf32 probe: D[128,256] = A[128,32] * B[256,32]^T with kind::tf32, operands K-major (kmaj=1) or MN-major (kmaj=0) ----
// MN-major SWIZZLE_128B tile: [mn group of 32][k row][128 B]; LBO = rows_k * 128 (next MN group), SBO = 1024 (next 8 k rows)
__global__ void __launch_bounds__(128, 1) probe3_kernel(const uint8_t* a_img, const uint8_t* b_img, int kmaj, float* d_out) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t* gen = raw + (base - smem_u32(raw));
    const uint32_t a_s = base, b_s = base + 16384, bar = base + 16384 + 32768, slot = bar + 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(slot, 256);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    uint32_t tmem; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
    for (int i = threadIdx.x; i < 16384 / 16; i += 128) reinterpret_cast<uint4*>(gen)[i] = reinterpret_cast<const uint4*>(a_img)[i];
    for (int i = threadIdx.x; i < 32768 / 16; i += 128) reinterpret_cast<uint4*>(gen + 16384)[i] = reinterpret_cast<const uint4*>(b_img)[i];
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
        tc_fence_after();
        // idesc: D f32 (1<<4), A/B tf32 (2<<7, 2<<10), majors bits 15/16, N>>3 at 17, M>>4 at 24
        uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
        if (!kmaj) idesc |= (1u << 15) | (1u << 16);
        for (int k = 0; k < 4; ++k) {      // 4 x K=8
            uint64_t ad, bd;
            if (kmaj) { ad = make_desc(a_s + k * 32, 16, 1024, 2); bd = make_desc(b_s + k * 32, 16, 1024, 2); }
            else { ad = make_desc(a_s + k * 1024, 4096, 1024, 2); bd = make_desc(b_s + k * 1024, 4096, 1024, 2); }
            mma_any(tmem, ad, bd, idesc, k != 0);
        }
        mma_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    tc_fence_after();
    const int r = warp * 32 + lane;
    for (int j = 0; j < 8; ++j) {
        uint32_t x[32];
        tmem_ld32(tmem + ((uint32_t)warp << 21) + j * 32, x);
        tmem_ld_wait();
        for (int e = 0; e < 32; ++e) d_out[r * 256 + j * 32 + e] = __uint_as_float(x[e]);
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

static int run_probe3(int kmaj) {
    const int M = 128, N = 256, K = 32;
    std::vector<float> A(M * K), B(N * K), D(M * N);
    srand(11 + kmaj);
    for (auto& x : A) x = (float)(rand() % 9 - 4);
    for (auto& x : B) x = (float)(rand() % 7 - 3);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += A[m * K + k] * B[n * K + k]; D[m * N + n] = s; }
    std::vector<uint8_t> a_img(16384), b_img(32768);
    auto put = [&](std::vector<uint8_t>& img, int mn, int k, float v) {
        uint32_t off;
        if (kmaj) off = sw128_offset(mn, k / 4) + (k % 4) * 4;                                   // row = mn, 32 k per 128 B row
        else { int g = mn / 32, w = mn % 32; off = g * 4096 + k * 128 + (((w / 4) ^ (k & 7)) << 4) + (w % 4) * 4; }
        memcpy(&img[off], &v, 4);
    };
    for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) put(a_img, m, k, A[m * K + k]);
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) put(b_img, n, k, B[n * K + k]);
    uint8_t *da, *db; float* dd;
    cudaMalloc(&da, 16384); cudaMalloc(&db, 32768); cudaMalloc(&dd, M * N * 4);
    cudaMemcpy(da, a_img.data(), 16384, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b_img.data(), 32768, cudaMemcpyHostToDevice);
    cudaMemset(dd, 0xff, M * N * 4);
    cudaFuncSetAttribute(probe3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    probe3_kernel<<<1, 128, 64 * 1024>>>(da, db, kmaj, dd);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("PROBE3 tf32 %s : CUDA error %s\n", kmaj ? "K-major" : "MN-major", cudaGetErrorString(e)); return 2; }
    std::vector<float> got(M * N);
    cudaMemcpy(got.data(), dd, M * N * 4, cudaMemcpyDeviceToHost);
    int bad = 0, first = -1;
    for (int i = 0; i < M * N; ++i) if (got[i] != D[i]) { if (first < 0) first = i; ++bad; }
    printf("PROBE3 kind::tf32 %s operands : mismatches %d / %d", kmaj ? "K-major " : "MN-major", bad, M * N);
    if (bad) printf("  first at (m=%d,n=%d) got %g want %g", first / N, first % N, got[first], D[first]);
    printf("\n%s\n", bad ? "PROBE3 FAILED" : "PROBE3 OK");
    return bad ? 1 : 0;
}

static uint16_t bf16_bits(float f);

// ---- bf16 MN-major probe (wgrad operand layout): D[128 m, N n] = sum_k A[k][m] * B[k][n], K = 64 rows.
// Operands are the tiled activation images the training kernels spill: blocks of [K rows x 64 columns] bf16, row pitch
// 128 B, 16-byte chunk c of row r at c ^ (r & 7) (the K-major SWIZZLE_128B image of a [rows x 64] tile, re-read with the
// row index as K).  Stage = half blocks of 64 rows (8 KB), M / N extend over consecutive half blocks.
__global__ void __launch_bounds__(128, 1) probe4_kernel(const uint8_t* a_img, const uint8_t* b_img, int n_cols, uint32_t lbo, uint32_t sbo,
                                                        float* d_out) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t* gen = raw + (base - smem_u32(raw));
    const uint32_t a_s = base, b_s = base + 16384, bar = base + 16384 + 32768, slot = bar + 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(slot, 256);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    uint32_t tmem; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
    for (int i = threadIdx.x; i < 16384 / 16; i += 128) reinterpret_cast<uint4*>(gen)[i] = reinterpret_cast<const uint4*>(a_img)[i];
    for (int i = threadIdx.x; i < 32768 / 16; i += 128) reinterpret_cast<uint4*>(gen + 16384)[i] = reinterpret_cast<const uint4*>(b_img)[i];
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
        tc_fence_after();
        const uint32_t idesc = make_idesc_bf16(128, (uint32_t)n_cols) | (1u << 15) | (1u << 16);     // A and B MN-major
        for (int k = 0; k < 4; ++k) {      // 4 x K=16 = two 8-row groups each
            uint64_t ad = make_desc(a_s + k * 2048, lbo, sbo, 2), bd = make_desc(b_s + k * 2048, lbo, sbo, 2);
            mma_bf16(tmem, ad, bd, idesc, k != 0);
        }
        mma_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    tc_fence_after();
    const int r = warp * 32 + lane;
    for (int j = 0; j < n_cols / 32; ++j) {
        uint32_t x[32];
        tmem_ld32(tmem + ((uint32_t)warp << 21) + j * 32, x);
        tmem_ld_wait();
        for (int e = 0; e < 32; ++e) d_out[r * 256 + j * 32 + e] = __uint_as_float(x[e]);
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

static int run_probe4(int n_cols, uint32_t lbo, uint32_t sbo) {
    const int M = 128, N = 256, K = 64;
    std::vector<float> A(K * M), B(K * N), D(M * N, 0.f);
    srand(23);
    for (auto& x : A) x = (float)(rand() % 9 - 4);
    for (auto& x : B) x = (float)(rand() % 7 - 3);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += A[k * M + m] * B[k * N + n]; D[m * N + n] = s; }
    std::vector<uint8_t> a_img(16384), b_img(32768);
    auto put = [&](std::vector<uint8_t>& img, int k, int col, float v) {
        uint16_t h = bf16_bits(v);
        uint32_t off = (uint32_t)(col / 64) * 8192u + sw128_offset((uint32_t)k, (uint32_t)((col % 64) / 8)) + (uint32_t)(col % 8) * 2u;
        memcpy(&img[off], &h, 2);
    };
    for (int k = 0; k < K; ++k) for (int m = 0; m < M; ++m) put(a_img, k, m, A[k * M + m]);
    for (int k = 0; k < K; ++k) for (int n = 0; n < N; ++n) put(b_img, k, n, B[k * N + n]);
    uint8_t *da, *db; float* dd;
    cudaMalloc(&da, 16384); cudaMalloc(&db, 32768); cudaMalloc(&dd, M * N * 4);
    cudaMemcpy(da, a_img.data(), 16384, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b_img.data(), 32768, cudaMemcpyHostToDevice);
    cudaMemset(dd, 0xff, M * N * 4);
    cudaFuncSetAttribute(probe4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    probe4_kernel<<<1, 128, 64 * 1024>>>(da, db, n_cols, lbo, sbo, dd);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("PROBE4 bf16 MN-major : CUDA error %s\n", cudaGetErrorString(e)); return 2; }
    std::vector<float> got(M * N);
    cudaMemcpy(got.data(), dd, M * N * 4, cudaMemcpyDeviceToHost);
    int bad = 0, first = -1;
    for (int m = 0; m < M; ++m) for (int n = 0; n < n_cols; ++n) if (got[m * N + n] != D[m * N + n]) { if (first < 0) first = m * N + n; ++bad; }
    printf("PROBE4 bf16 MN-major A,B (tiled activation images) N=%d LBO=%u SBO=%u : mismatches %d / %d", n_cols, lbo, sbo, bad, M * n_cols);
    if (bad) printf("  first at (m=%d,n=%d) got %g want %g", first / N, first % N, got[first], D[first]);
    printf("\n%s\n", bad ? "PROBE4 variant failed" : "PROBE4 OK");
    cudaFree(da); cudaFree(db); cudaFree(dd);
    return bad ? 1 : 0;
}

static int run_probe2() {
    const int M = 256, N = 256, K = 64;
    std::vector<float> A(M * K), B(N * K), D(M * N);
    srand(7);
    for (auto& x : A) x = (float)(rand() % 7 - 3);
    for (auto& x : B) x = (float)(rand() % 5 - 2);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += A[m * K + k] * B[n * K + k]; D[m * N + n] = s; }
    std::vector<uint8_t> a_img(32768), b_img(32768);
    for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) {
        uint16_t h = bf16_bits(A[m * K + k]);
        memcpy(&a_img[(m / 128) * 16384 + sw128_offset(m % 128, k / 8) + (k % 8) * 2], &h, 2);
    }
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) {      // hypothesis: CTA r holds B rows [128 r, 128 r + 128)
        uint16_t h = bf16_bits(B[n * K + k]);
        memcpy(&b_img[(n / 128) * 16384 + sw128_offset(n % 128, k / 8) + (k % 8) * 2], &h, 2);
    }
    uint8_t *da, *db; float* dd;
    cudaMalloc(&da, 32768); cudaMalloc(&db, 32768); cudaMalloc(&dd, M * N * 4);
    cudaMemcpy(da, a_img.data(), 32768, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b_img.data(), 32768, cudaMemcpyHostToDevice);
    cudaMemset(dd, 0xff, M * N * 4);
    cudaFuncSetAttribute(probe2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
    probe2_kernel<<<2, 128, 48 * 1024>>>(da, db, dd);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("PROBE2 cta_group::2 : CUDA error %s\n", cudaGetErrorString(e)); return 2; }
    std::vector<float> got(M * N);
    cudaMemcpy(got.data(), dd, M * N * 4, cudaMemcpyDeviceToHost);
    int bad = 0, first = -1;
    for (int i = 0; i < M * N; ++i) if (got[i] != D[i]) { if (first < 0) first = i; ++bad; }
    printf("PROBE2 cta_group::2 M=256 N=256 (A rows and B rows split by CTA rank) : mismatches %d / %d", bad, M * N);
    if (bad) {
        printf("  first at (m=%d,n=%d) got %g want %g\n", first / N, first % N, got[first], D[first]);
        // diagnose the column mapping of row 0 and row 128
        for (int m : {0, 128}) {
            printf("   row %d: D columns 0,1,2,128,129 match expected n =", m);
            for (int j : {0, 1, 2, 128, 129}) {
                int hit = -1;
                for (int n = 0; n < N; ++n) if (got[m * N + j] == D[m * N + n]) { hit = n; break; }
                printf(" %d", hit);
            }
            printf("\n");
        }
    } else printf("\n");
    printf(bad ? "PROBE2 FAILED\n" : "PROBE2 OK\n");
    return bad ? 1 : 0;
}

static uint16_t bf16_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return (uint16_t)(u >> 16); }   // exact for small ints

int main() {
    const int M = 128, N = 256, K = 64;
    std::vector<float> A(M * K), B(N * K), D(M * N);
    srand(1);
    for (auto& x : A) x = (float)(rand() % 7 - 3);
    for (auto& x : B) x = (float)(rand() % 5 - 2);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += A[m * K + k] * B[n * K + k]; D[m * N + n] = s; }
    std::vector<uint8_t> a_img(16384), b_img[2] = {std::vector<uint8_t>(32768), std::vector<uint8_t>(32768)};
    for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) {
        uint16_t h = bf16_bits(A[m * K + k]);
        memcpy(&a_img[sw128_offset(m, k / 8) + (k % 8) * 2], &h, 2);
    }
    for (int var = 0; var < 2; ++var)
        for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) {
            uint16_t h = bf16_bits(B[n * K + k]);
            int c = k / 32, kk = k % 32;
            uint32_t off = var == 0 ? sw64_offset(n, kk / 8)
                                    : (uint32_t)((n >> 3) * 512 + (n & 7) * 64 + (((kk / 8) ^ (n & 3)) << 4));
            memcpy(&b_img[var][c * 16384 + off + (kk % 8) * 2], &h, 2);
        }
    uint8_t *da, *db; float* dd;
    cudaMalloc(&da, 16384); cudaMalloc(&db, 32768); cudaMalloc(&dd, M * N * 4);
    cudaMemcpy(da, a_img.data(), 16384, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    struct { const char* name; int bvar; Variant v; } cases[] = {
        {"A=SW128/sbo1024 B=SW64(r>>1)/sbo512 plain-store", 0, {2, 1024, 4, 512, 0}},
        {"A=SW128/sbo1024 B=SW64(r>>1)/sbo512 bulk-copy  ", 0, {2, 1024, 4, 512, 1}},
        {"A=SW128/sbo1024 B=SW64(r&3)/sbo512 plain-store ", 1, {2, 1024, 4, 512, 0}},
    };
    int ok_any = 0;
    for (auto& cs : cases) {
        cudaMemcpy(db, b_img[cs.bvar].data(), 32768, cudaMemcpyHostToDevice);
        cudaMemset(dd, 0xff, M * N * 4);
        probe_kernel<<<1, 128, 64 * 1024>>>(da, db, cs.v, dd);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("PROBE %s : CUDA error %s\n", cs.name, cudaGetErrorString(e)); return 2; }
        std::vector<float> got(M * N);
        cudaMemcpy(got.data(), dd, M * N * 4, cudaMemcpyDeviceToHost);
        int bad = 0, first = -1;
        for (int i = 0; i < M * N; ++i) if (got[i] != D[i]) { if (first < 0) first = i; ++bad; }
        printf("PROBE %s : mismatches %d / %d", cs.name, bad, M * N);
        if (bad) printf("  first at (m=%d,n=%d) got %g want %g", first / N, first % N, got[first], D[first]);
        printf("\n");
        if (!bad) ok_any = 1;
    }
    printf(ok_any ? "PROBE OK\n" : "PROBE FAILED\n");
    run_probe2();
    run_probe3(1);
    run_probe3(0);
    run_probe4(256, 8192, 1024);
    run_probe4(64, 8192, 1024);
    run_probe4(256, 1024, 8192);
    return 0;
}
