// Building block for the next round's training-tile layout (first row): tcgen05.mma with MN-major bf16 operands in
// the NO-SWIZZLE core-matrix layout.  A half block (64 K-rows x 64 MN-columns) is stored [chunk c (8)][k (64)][16 B]: the 8
// MN-elements of chunk c of row k at c * 1024 + k * 16 -- what the epilogue threads would write with fully coalesced stores
// (thread = row k).  D[128 m, N n] = sum_k A[k][m] B[k][n], K = 64; M / N extend over consecutive 8 KB half blocks.
// The probe tries the LBO / SBO assignments and reports which one the hardware implements (exact integer reference).
// nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a umma_noswizzle_probe.cu -o umma_noswizzle_probe
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "../../msra_practice_project_b200/csrc/umma.cuh"

using namespace b2r::umma;

__global__ void __launch_bounds__(128, 1) probe_kernel(const uint8_t* a_img, const uint8_t* b_img, int n_cols, uint32_t lbo, uint32_t sbo, uint32_t kstep,
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
        for (int k = 0; k < 4; ++k) {      // 4 x K=16
            uint64_t ad = make_desc(a_s + k * kstep, lbo, sbo, 0), bd = make_desc(b_s + k * kstep, lbo, sbo, 0);      // layout 0 = no swizzle
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

static uint16_t bf16_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return (uint16_t)(u >> 16); }       // small integers: exact

int main() {
    const int M = 128, N = 256, K = 64;
    std::vector<float> A(K * M), B(K * N), D(M * N, 0.f);
    srand(23);
    for (auto& x : A) x = (float)(rand() % 9 - 4);
    for (auto& x : B) x = (float)(rand() % 7 - 3);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += A[k * M + m] * B[k * N + n]; D[m * N + n] = s; }
    std::vector<uint8_t> a_img(16384), b_img(32768);
    auto put = [&](std::vector<uint8_t>& img, int k, int col, float v) {
        uint16_t h = bf16_bits(v);
        uint32_t off = (uint32_t)(col / 64) * 8192u + (uint32_t)((col % 64) / 8) * 1024u + (uint32_t)k * 16u + (uint32_t)(col % 8) * 2u;
        memcpy(&img[off], &h, 2);
    };
    for (int k = 0; k < K; ++k) for (int m = 0; m < M; ++m) put(a_img, k, m, A[k * M + m]);
    for (int k = 0; k < K; ++k) for (int n = 0; n < N; ++n) put(b_img, k, n, B[k * N + n]);
    uint8_t *da, *db; float* dd;
    cudaMalloc(&da, 16384); cudaMalloc(&db, 32768); cudaMalloc(&dd, M * N * 4);
    cudaMemcpy(da, a_img.data(), 16384, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b_img.data(), 32768, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    // chunk pitch (next 8 MN elements) 1024 B, 8-row K group pitch 128 B, K = 16 per MMA = 256 B
    const uint32_t combos[4][3] = {{1024, 128, 256}, {128, 1024, 256}, {1024, 128, 128}, {128, 1024, 128}};
    int ok_any = 0;
    for (auto& c : combos) {
        cudaMemset(dd, 0xff, M * N * 4);
        probe_kernel<<<1, 128, 64 * 1024>>>(da, db, N, c[0], c[1], c[2], dd);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("LBO=%u SBO=%u kstep=%u : CUDA error %s\n", c[0], c[1], c[2], cudaGetErrorString(e)); return 2; }
        std::vector<float> got(M * N);
        cudaMemcpy(got.data(), dd, M * N * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int i = 0; i < M * N; ++i) if (got[i] != D[i]) ++bad;
        printf("no-swizzle MN-major [chunk][k][16 B]: LBO=%u SBO=%u K-step=%u B : mismatches %d / %d %s\n", c[0], c[1], c[2], bad, M * N, bad ? "" : "<-- OK");
        ok_any |= !bad;
    }
    printf(ok_any ? "PROBE OK\n" : "PROBE FAILED\n");
    return ok_any ? 0 : 1;
}
