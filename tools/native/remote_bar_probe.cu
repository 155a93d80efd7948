// Can a plain (non-tensor) cp.async.bulk issued by the PEER CTA of a cluster signal the LEADER's mbarrier (destination in the peer's own
// shared memory, mbarrier operand = the leader's barrier address mapped with mapa)?  If so the weight ring of the CTA-pair kernels needs no
// relay hop (tc_core.cuh: relay_loop).  Prints the cycles from issue to the leader seeing the barrier complete, for the relay scheme
// (peer waits locally, then remote-arrives) and for the direct scheme.
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a remote_bar_probe.cu -o remote_bar_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r; }
__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// mode 0: relay (leader barrier count 2: own copy + relay arrive).  mode 1: direct (leader barrier count 1, expects both CTAs' bytes; the peer's
// copy names the leader's barrier).
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64, 1) k(const uint8_t* __restrict__ src, int mode, int iters, long long* __restrict__ out, uint32_t* __restrict__ check) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ __align__(8) unsigned long long bar, go;
    const uint32_t rank = ctarank(), b = smem_u32(&bar), g = smem_u32(&go), dst = smem_u32(sm);
    if (threadIdx.x == 0) { mbar_init(b, ((mode == 0 || mode == 3) && rank == 0) ? 2 : 1); mbar_init(g, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    cluster_sync();
    long long total = 0; uint32_t ph = 0; int failed = 0;
    const uint32_t bytes = 16384;
    for (int it = 0; it < iters; ++it) {
        if (threadIdx.x == 0 && !failed) {
            const uint8_t* s = src + (size_t)((it * 2 + rank) % 64) * bytes;
            const long long t0 = clock64();
            if (mode == 2) {          // local only: no pairing, the leader waits for its own copy
                mbar_expect(b, bytes);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(s), "r"(bytes), "r"(b) : "memory");
                { uint32_t sp = 0; while (!mbar_try(b, ph)) { if (++sp > (1u << 20)) { failed = 1; break; } } }
                if (rank == 0) total += clock64() - t0;
            } else if (mode == 3) {   // relay, polling with test_wait (never suspends)
                mbar_expect(b, bytes);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(s), "r"(bytes), "r"(b) : "memory");
                { uint32_t sp = 0; while (!mbar_test(b, ph)) { if (++sp > (1u << 22)) { failed = 1; break; } } }
                if (rank == 1) asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(mapa(b, 0)) : "memory");
                else total += clock64() - t0;
            } else if (mode == 0) {
                mbar_expect(b, bytes);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(s), "r"(bytes), "r"(b) : "memory");
                if (rank == 1) {
                    { uint32_t sp = 0; while (!mbar_try(b, ph)) { if (++sp > (1u << 20)) { failed = 1; break; } } }
                    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(mapa(b, 0)) : "memory");
                } else {
                    { uint32_t sp = 0; while (!mbar_try(b, ph)) { if (++sp > (1u << 20)) { failed = 1; break; } } }
                    total += clock64() - t0;
                }
            } else {
                if (rank == 0) {
                    mbar_expect(b, 2 * bytes);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(s), "r"(bytes), "r"(b) : "memory");
                    { uint32_t sp = 0; while (!mbar_try(b, ph)) { if (++sp > (1u << 20)) { failed = 1; break; } } }
                    total += clock64() - t0;
                } else {
                    // destination: own shared memory (shared::cluster address of this CTA), barrier: the leader's
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(mapa(dst, 1)), "l"(s), "r"(bytes),
                                 "r"(mapa(b, 0)) : "memory");
                }
            }
            ph ^= 1u;
        }
        __syncthreads();
        cluster_sync();            // both CTAs in step: the leader's wait has completed, so the peer's data has landed
        if (it == iters - 1 && threadIdx.x < 32) check[rank * 32 + threadIdx.x] = reinterpret_cast<uint32_t*>(sm)[threadIdx.x * 127];
    }
    if (threadIdx.x == 0 && rank == 0) out[blockIdx.x / 2] = failed ? -1 : total;
}

int main() {
    uint8_t* src; long long* out; uint32_t* check;
    cudaMalloc(&src, 64 * 16384); cudaMalloc(&out, 74 * sizeof(long long)); cudaMalloc(&check, 64 * 4);
    uint32_t* h = new uint32_t[64 * 4096];
    for (int i = 0; i < 64 * 4096; ++i) h[i] = (uint32_t)i * 2654435761u;
    cudaMemcpy(src, h, 64 * 16384, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    const int iters = 200; setvbuf(stdout, nullptr, _IONBF, 0);
    for (int grid : {2, 148}) for (int mode : {2, 0, 3}) {
        cudaMemset(check, 0, 64 * 4);
        k<<<grid, 64, 32768>>>(src, mode, iters, out, check);
        cudaError_t e = cudaDeviceSynchronize();
        long long t[74]; uint32_t c[64];
        cudaMemcpy(t, out, sizeof(long long) * (grid / 2), cudaMemcpyDeviceToHost); cudaMemcpy(c, check, sizeof c, cudaMemcpyDeviceToHost);
        double mean = 0; for (int i = 0; i < grid / 2; ++i) mean += (double)t[i] / iters; mean /= grid / 2;
        // last iteration: rank r copied chunk ((iters-1)*2 + r) % 64
        int bad = 0;
        for (int r = 0; r < 2; ++r) for (int j = 0; j < 32; ++j) if (c[r * 32 + j] != h[(size_t)(((iters - 1) * 2 + r) % 64) * 4096 + j * 127]) ++bad;
        fflush(stdout);
        printf("%s, %3d CTAs: issue -> leader sees both halves: %.0f clk   data check: %s   (%s)\n", mode == 2 ? "local only (no pairing)                        " : (mode == 3 ? "relay, test_wait polling                      " : "relay  (peer try_wait, then remote arrive)     "),
               grid, mean, bad ? "MISMATCH" : "ok", cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
    }
    return 0;
}
