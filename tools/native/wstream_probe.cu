// How fast can ONE SM pull an L2-resident weight image into shared memory?  (the fused MLP kernels stream 16 KB chunks through a
// 3-stage ring with 1-D cp.async.bulk; measured here: what that path delivers and what the alternatives do)
//   A: 1-D cp.async.bulk, op size x ops in flight          B: 2-D tensor-map TMA (cp.async.bulk.tensor.2d), 16 KB boxes, 3 in flight
//   C: LDGSTS (cp.async 16 B per thread) by one warp / four warps, 16 KB stages, 3 in flight
// Reports bytes per SM clock and GB/s (globaltimer), 1 CTA and one CTA per SM.
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a wstream_probe.cu -o wstream_probe
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

constexpr uint32_t kRing = 196608;
// mode 0: 1-D bulk; mode 1: tensor map 2-D (op = 16 KB box); mode 2: LDGSTS by `nwarps` warps
__global__ void __launch_bounds__(160, 1) k_stream(const uint8_t* __restrict__ wbuf, uint32_t w_bytes, const __grid_constant__ CUtensorMap tm, int mode,
                                                   uint32_t op_bytes, int depth, int n_ops, int nwarps, long long* __restrict__ stats) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ __align__(8) unsigned long long bars[16];
    const uint32_t ring = (smem_u32(sm) + 1023u) & ~1023u;
    if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) mbar_init(smem_u32(&bars[i]), mode == 2 ? 32 * nwarps : 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const uint32_t n_src = w_bytes / op_bytes;
    if (mode < 2) {
        if (threadIdx.x == 0) {
            const long long t0 = clock64(); const unsigned long long g0 = gtimer();
            uint32_t par = 0;
            for (int i = 0; i < n_ops + depth; ++i) {
                const int s = i % depth;
                if (i >= depth) { mbar_wait(smem_u32(&bars[s]), (par >> s) & 1u); par ^= 1u << s; }
                if (i < n_ops) {
                    const uint32_t k = (uint32_t)(i * 7 + blockIdx.x) % n_src;
                    mbar_expect(smem_u32(&bars[s]), op_bytes);
                    if (mode == 0)
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ring + s * op_bytes),
                                     "l"(wbuf + (size_t)k * op_bytes), "r"(op_bytes), "r"(smem_u32(&bars[s])) : "memory");
                    else
                        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                                     ::"r"(ring + s * op_bytes), "l"(&tm), "r"(0), "r"((int)(k * 128)), "r"(smem_u32(&bars[s])) : "memory");
                }
            }
            stats[blockIdx.x * 4 + 0] = clock64() - t0; stats[blockIdx.x * 4 + 1] = (long long)(gtimer() - g0);
        }
    } else {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (warp < nwarps) {
            const long long t0 = clock64(); const unsigned long long g0 = gtimer();
            uint32_t par = 0;
            const uint32_t per_warp = op_bytes / nwarps;
            for (int i = 0; i < n_ops + depth; ++i) {
                const int s = i % depth;
                if (i >= depth) { mbar_wait(smem_u32(&bars[s]), (par >> s) & 1u); par ^= 1u << s; }
                if (i < n_ops) {
                    const uint32_t k = (uint32_t)(i * 7 + blockIdx.x) % n_src;
                    const uint8_t* src = wbuf + (size_t)k * op_bytes + warp * per_warp + lane * 16;
                    const uint32_t dst = ring + s * op_bytes + warp * per_warp + lane * 16;
                    for (uint32_t o = 0; o < per_warp; o += 512)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + o), "l"(src + o) : "memory");
                    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&bars[s])) : "memory");
                }
            }
            if (threadIdx.x == 0) { stats[blockIdx.x * 4 + 0] = clock64() - t0; stats[blockIdx.x * 4 + 1] = (long long)(gtimer() - g0); }
        }
    }
}

int main() {
    const uint32_t w_bytes = 74 * 16384 * 2;            // 2.4 MB: L2-resident
    uint8_t* wbuf; long long* stats;
    cudaMalloc(&wbuf, w_bytes); cudaMemset(wbuf, 1, w_bytes);
    cudaMalloc(&stats, 148 * 4 * sizeof(long long));
    void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn) { printf("no cuTensorMapEncodeTiled\n"); return 2; }
    auto encode = reinterpret_cast<CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill)>(fn);
    // the weight image as a [rows x 64] bf16 matrix (128 B per row); a box of 128 rows = 16 KB, SWIZZLE_128B
    CUtensorMap tm;
    const cuuint64_t dims[2] = {64, w_bytes / 128};
    const cuuint64_t strides[1] = {128};
    const cuuint32_t box[2] = {64, 128}, estr[2] = {1, 1};
    CUresult rc = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, wbuf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)rc); return 3; }
    const int smem = kRing + 2048;
    cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    long long h[148 * 4];
    auto run = [&](const char* name, int grid, int mode, uint32_t op_bytes, int depth, int nwarps = 1) {
        const int n_ops = (int)((64u << 20) / op_bytes);           // 64 MB per CTA
        for (int rep = 0; rep < 2; ++rep) {
            k_stream<<<grid, 160, smem>>>(wbuf, w_bytes, tm, mode, op_bytes, depth, n_ops, nwarps, stats);
            cudaDeviceSynchronize();
        }
        cudaMemcpy(h, stats, sizeof h, cudaMemcpyDeviceToHost);
        double cyc = 0, ns = 0;
        for (int b = 0; b < grid; ++b) { cyc += h[b * 4]; ns += h[b * 4 + 1]; }
        cyc /= grid; ns /= grid;
        const double bytes = (double)n_ops * op_bytes;
        printf("%-46s grid %3d | %6.1f B/clk/SM  %6.1f GB/s/SM  clk %.2f GHz | per op %6.0f clk  (%s)\n", name, grid, bytes / cyc, bytes / ns, cyc / ns,
               cyc / n_ops, cudaGetErrorString(cudaGetLastError()));
    };
    for (int grid : {1, 148}) {
        run("1-D bulk  4 KB x 12", grid, 0, 4096, 12);
        run("1-D bulk  8 KB x 6", grid, 0, 8192, 6);
        run("1-D bulk 16 KB x 1", grid, 0, 16384, 1);
        run("1-D bulk 16 KB x 2", grid, 0, 16384, 2);
        run("1-D bulk 16 KB x 3 (today's ring)", grid, 0, 16384, 3);
        run("1-D bulk 16 KB x 6", grid, 0, 16384, 6);
        run("1-D bulk 16 KB x 12", grid, 0, 16384, 12);
        run("1-D bulk 32 KB x 2", grid, 0, 32768, 2);
        run("1-D bulk 32 KB x 3", grid, 0, 32768, 3);
        run("1-D bulk 64 KB x 2", grid, 0, 65536, 2);
        run("1-D bulk 64 KB x 3", grid, 0, 65536, 3);
        run("tensor-map 2-D 16 KB box x 3", grid, 1, 16384, 3);
        run("tensor-map 2-D 16 KB box x 6", grid, 1, 16384, 6);
        run("LDGSTS 1 warp, 16 KB x 3", grid, 2, 16384, 3, 1);
        run("LDGSTS 4 warps, 16 KB x 3", grid, 2, 16384, 3, 4);
        run("LDGSTS 4 warps, 16 KB x 6", grid, 2, 16384, 6, 4);
    }
    return 0;
}
