// Per-SM bulk-copy (TMA) engine probe: what happens to L2 -> shared weight-chunk loads (16 KB, 3 in flight: the fused MLP kernels'
// weight ring) while the same SM streams 64 KB shared -> global tile copies (the training kernels' activation spill)?
// Reports, for 1 CTA and for one CTA per SM: load throughput and mean issue-to-arrival latency alone, store throughput alone, and
// both together.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a tma_mix_probe.cu -o tma_mix_probe
#include <cstdio>
#include <cstdint>
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

// out_stats[cta]: {load cycles, load latency sum, loads, store cycles, stores}
__global__ void __launch_bounds__(128, 1) k_mix(const uint8_t* __restrict__ wbuf, uint32_t w_chunks, uint8_t* __restrict__ out, size_t out_chunks,
                                                int n_loads, int n_stores, uint32_t store_bytes, long long* __restrict__ stats, int pieces, int depth, int lsu_stores, int ldgsts_loads) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ __align__(8) unsigned long long bars[8];
    const uint32_t ring = smem_u32(sm), stile = ring + 3 * 16384;
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bars[i]), ldgsts_loads ? 32 : 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    for (uint32_t i = threadIdx.x; i < 2 * 65536 / 16; i += blockDim.x) reinterpret_cast<uint4*>(sm + 3 * 16384)[i] = make_uint4(i, 2u, 3u, 4u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (ldgsts_loads && threadIdx.x < 32 && n_loads > 0) {
        // the weight ring fed by LDGSTS (cp.async 16 B per lane, completion through cp.async.mbarrier.arrive.noinc): does not use the bulk-copy engine
        const int lane = threadIdx.x;
        long long issue_t[8] = {0, 0, 0, 0, 0, 0, 0, 0}, lat = 0;
        const long long t0 = clock64();
        uint32_t ph = 0;
        for (int i = 0; i < n_loads + depth; ++i) {
            const int s = i % depth;
            if (i >= depth) { mbar_wait(smem_u32(&bars[s]), ph); lat += clock64() - issue_t[s]; if (s == depth - 1) ph ^= 1u; }
            if (i < n_loads) {
                const uint8_t* src = wbuf + (size_t)((i * 7 + blockIdx.x) % w_chunks) * 16384 + lane * 16;
                const uint32_t dst = ring + s * 16384 + lane * 16;
                issue_t[s] = clock64();
#pragma unroll 8
                for (uint32_t o = 0; o < 16384; o += 512) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + o), "l"(src + o) : "memory");
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&bars[s])) : "memory");
            }
        }
        if (lane == 0) { stats[blockIdx.x * 8 + 0] = clock64() - t0; stats[blockIdx.x * 8 + 1] = lat; stats[blockIdx.x * 8 + 2] = n_loads; }
    }
    if (!ldgsts_loads && threadIdx.x == 0 && n_loads > 0) {
        long long issue_t[8] = {0, 0, 0, 0, 0, 0, 0, 0}, lat = 0;
        const long long t0 = clock64();
        uint32_t ph = 0;
        const uint32_t stage_bytes = 49152u / depth;
        for (int i = 0; i < n_loads + depth; ++i) {
            const int s = i % depth;
            if (i >= depth) { mbar_wait(smem_u32(&bars[s]), ph); lat += clock64() - issue_t[s]; if (s == depth - 1) ph ^= 1u; }
            if (i < n_loads) {
                const uint8_t* src = wbuf + (size_t)((i * 7 + blockIdx.x) % w_chunks) * 16384;
                mbar_expect(smem_u32(&bars[s]), stage_bytes);
                issue_t[s] = clock64();
                const uint32_t pb = stage_bytes / pieces;
                for (int q = 0; q < pieces; ++q)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ring + s * stage_bytes + q * pb), "l"(src + q * pb), "r"(pb),
                                 "r"(smem_u32(&bars[s])) : "memory");
            }
        }
        stats[blockIdx.x * 8 + 0] = clock64() - t0; stats[blockIdx.x * 8 + 1] = lat; stats[blockIdx.x * 8 + 2] = n_loads;
    }
    if (lsu_stores && threadIdx.x >= 64 && n_stores > 0) {
        const int w = (threadIdx.x >> 5) - 2, lane = threadIdx.x & 31;       // 2 warps split every 64 KB copy
        const long long t0 = clock64();
        for (int k = 0; k < n_stores; ++k) {
            uint8_t* dst = out + (((size_t)k * gridDim.x + blockIdx.x) % out_chunks) * 65536;
            const uint8_t* src = sm + 3 * 16384 + (k & 1) * 65536;
#pragma unroll 8
            for (uint32_t o = (uint32_t)w * 512u + (uint32_t)lane * 16u; o < store_bytes; o += 1024u)
                *reinterpret_cast<uint4*>(dst + o) = *reinterpret_cast<const uint4*>(src + o);
        }
        if (threadIdx.x == 64) { stats[blockIdx.x * 8 + 3] = clock64() - t0; stats[blockIdx.x * 8 + 4] = n_stores; }
    }
    if (!lsu_stores && threadIdx.x == 32 && n_stores > 0) {
        const long long t0 = clock64();
        for (int k = 0; k < n_stores; ++k) {
            uint8_t* dst = out + (((size_t)k * gridDim.x + blockIdx.x) % out_chunks) * 65536;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(stile + (k & 1) * 65536), "r"(store_bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        stats[blockIdx.x * 8 + 3] = clock64() - t0; stats[blockIdx.x * 8 + 4] = n_stores;
    }
}

int main() {
    const uint32_t w_chunks = 74;                    // 1.2 MB of "weights": L2-resident
    const size_t out_bytes = (size_t)8 << 30, out_chunks = out_bytes / 65536;
    uint8_t *wbuf, *out; long long* stats;
    cudaMalloc(&wbuf, (size_t)w_chunks * 16384); cudaMemset(wbuf, 1, (size_t)w_chunks * 16384);
    cudaMalloc(&out, out_bytes); cudaMalloc(&stats, 148 * 8 * sizeof(long long));
    const int smem = 3 * 16384 + 2 * 65536 + 1024;
    cudaFuncSetAttribute(k_mix, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    long long h[148 * 8];
    auto run = [&](const char* name, int grid, int n_loads, int n_stores, uint32_t store_bytes, int pieces = 1, int depth = 3, int lsu = 0, int ldg = 0) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaMemset(stats, 0, 148 * 8 * sizeof(long long));
            k_mix<<<grid, 128, smem>>>(wbuf, w_chunks, out, out_chunks, n_loads * depth / 3, n_stores, store_bytes, stats, pieces, depth, lsu, ldg);
            cudaDeviceSynchronize();
        }
        cudaMemcpy(h, stats, sizeof h, cudaMemcpyDeviceToHost);
        double lc = 0, ll = 0, sc = 0;
        for (int b = 0; b < grid; ++b) { lc += h[b * 8]; ll += h[b * 8 + 1]; sc += h[b * 8 + 3]; }
        n_loads = n_loads * depth / 3;
        lc /= grid; ll /= grid; sc /= grid;
        printf("%-52s grid %3d | loads: %6.1f B/clk/SM, latency %7.0f clk | stores: %6.1f B/clk/SM  (%s)\n", name, grid,
               n_loads ? n_loads * (49152.0 / depth) / lc : 0.0, n_loads ? ll / n_loads : 0.0, n_stores ? n_stores * (double)store_bytes / sc : 0.0,
               cudaGetErrorString(cudaGetLastError()));
    };
    for (int grid : {1, 148}) {
        run("TMA loads only (16 KB x 3 in flight)", grid, 4000, 0, 65536);
        run("LDGSTS loads only (1 warp, 16 KB x 3)", grid, 4000, 0, 65536, 1, 3, 0, 1);
        run("TMA loads + TMA 16 KB stores", grid, 4000, 4000, 16384);
        run("LDGSTS loads + TMA 16 KB stores", grid, 4000, 4000, 16384, 1, 3, 0, 1);
        run("LDGSTS loads + TMA 64 KB stores", grid, 4000, 1000, 65536, 1, 3, 0, 1);
    }
    return 0;
}
