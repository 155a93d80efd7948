// Pins the encoding the compact "post chunk" of the sine-model kernels relies on (tc_core.cuh, MapC): tcgen05.mma kind::f16 with
// K-major bf16 operands in the NO-SWIZZLE core-matrix layout [16-byte K chunk c (2)][row r][16 B] (one K = 16 step: chunk c of row r
// at c * rows * 16 + r * 16, i.e. core matrices of 8 rows x 16 B are 128 B apart along M / N and rows * 16 B apart along K), and the
// mix "A = SWIZZLE_128B block, B = no-swizzle" the training forward would use.  D[128, 128] = A[128, 16] B[128, 16]^T, exact integers.
// The probe tries both LBO / SBO assignments and reports which one the hardware implements.
// Built by __graft_entry__.build(), run by tests/test_gpu_parity.py::test_umma_kmajor_noswizzle_probe.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "../../msra_practice_project_b200/csrc/umma.cuh"

using namespace b2r::umma;

// a_mode 0: A no-swizzle at a_s (4 KB image); 1: A SWIZZLE_128B block at a_s (16 KB image, 16 K used)
__global__ void __launch_bounds__(128, 1) probe_kernel(const uint8_t* a_img, const uint8_t* b_img, int a_mode, uint32_t lbo, uint32_t sbo, float* d_out) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t* gen = raw + (base - smem_u32(raw));
    const uint32_t a_s = base, b_s = base + 16384, bar = base + 32768, slot = bar + 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(slot, 128);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    uint32_t tmem; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
    for (int i = threadIdx.x; i < 16384 / 16; i += 128) reinterpret_cast<uint4*>(gen)[i] = reinterpret_cast<const uint4*>(a_img)[i];
    for (int i = threadIdx.x; i < 4096 / 16; i += 128) reinterpret_cast<uint4*>(gen + 16384)[i] = reinterpret_cast<const uint4*>(b_img)[i];
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
        tc_fence_after();
        const uint32_t idesc = make_idesc_bf16(128, 128);            // A and B K-major
        const uint64_t ad = a_mode ? desc_sw128(a_s) : make_desc(a_s, lbo, sbo, 0);
        const uint64_t bd = make_desc(b_s, lbo, sbo, 0);
        mma_bf16(tmem, ad, bd, idesc, 0);
        mma_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    tc_fence_after();
    const int r = warp * 32 + lane;
    for (int j = 0; j < 4; ++j) {
        uint32_t x[32];
        tmem_ld32(tmem + ((uint32_t)warp << 21) + j * 32, x);
        tmem_ld_wait();
        for (int e = 0; e < 32; ++e) d_out[r * 128 + j * 32 + e] = __uint_as_float(x[e]);
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

static uint16_t bf16_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return (uint16_t)(u >> 16); }       // small integers: exact

int main() {
    const int M = 128, N = 128, K = 16;
    std::vector<float> A(M * K), B(N * K), D(M * N, 0.f);
    srand(29);
    for (auto& x : A) x = (float)(rand() % 9 - 4);
    for (auto& x : B) x = (float)(rand() % 7 - 3);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += A[m * K + k] * B[n * K + k]; D[m * N + n] = s; }
    std::vector<uint8_t> a_ns(16384, 0), a_sw(16384, 0), b_ns(4096, 0);
    for (int r = 0; r < 128; ++r)
        for (int k = 0; k < K; ++k) {
            uint16_t ha = bf16_bits(A[r * K + k]), hb = bf16_bits(B[r * K + k]);
            const uint32_t off_ns = (uint32_t)(k / 8) * 2048u + (uint32_t)r * 16u + (uint32_t)(k % 8) * 2u;
            memcpy(&a_ns[off_ns], &ha, 2);
            memcpy(&b_ns[off_ns], &hb, 2);
            memcpy(&a_sw[sw128_offset((uint32_t)r, (uint32_t)(k / 8)) + (uint32_t)(k % 8) * 2u], &ha, 2);
        }
    uint8_t *da, *das, *db; float* dd;
    cudaMalloc(&da, 16384); cudaMalloc(&das, 16384); cudaMalloc(&db, 4096); cudaMalloc(&dd, M * N * 4);
    cudaMemcpy(da, a_ns.data(), 16384, cudaMemcpyHostToDevice);
    cudaMemcpy(das, a_sw.data(), 16384, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b_ns.data(), 4096, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
    // K-direction core-matrix pitch 2048 B, M / N-direction (8-row group) pitch 128 B
    const uint32_t combos[2][2] = {{2048, 128}, {128, 2048}};
    int ok_all = 1;
    for (int a_mode = 0; a_mode < 2; ++a_mode) {
        int ok_any = 0;
        for (auto& c : combos) {
            cudaMemset(dd, 0xff, M * N * 4);
            probe_kernel<<<1, 128, 48 * 1024>>>(a_mode ? das : da, db, a_mode, c[0], c[1], dd);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("LBO=%u SBO=%u : CUDA error %s\n", c[0], c[1], cudaGetErrorString(e)); return 2; }
            std::vector<float> got(M * N);
            cudaMemcpy(got.data(), dd, M * N * 4, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int i = 0; i < M * N; ++i) if (got[i] != D[i]) ++bad;
            printf("K-major no-swizzle [chunk][row][16 B], A %s: LBO=%u SBO=%u : mismatches %d / %d %s\n", a_mode ? "SWIZZLE_128B" : "no-swizzle", c[0], c[1], bad,
                   M * N, bad ? "" : "<-- OK");
            ok_any |= !bad;
        }
        ok_all &= ok_any;
    }
    printf(ok_all ? "PROBE OK\n" : "PROBE FAILED\n");
    return ok_all ? 0 : 1;
}
