// HBM write-bandwidth probe: how fast can 148 SMs write a large buffer (a) with plain coalesced 16-byte stores, (b) with the
// bulk-copy engine from shared memory (cp.async.bulk shared -> global, 64 KB per copy, two in flight per CTA) -- the path the
// training kernels' tile spills take -- and (c) cudaMemsetAsync.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a write_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k_stg(uint4* __restrict__ out, size_t n16) {
    const uint4 v = make_uint4(1u, 2u, 3u, 4u);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) out[i] = v;
}

__global__ void __launch_bounds__(128, 1) k_bulk(uint8_t* __restrict__ out, size_t bytes, uint32_t chunk) {
    extern __shared__ __align__(128) uint8_t sm[];
    for (uint32_t i = threadIdx.x; i < chunk * 2 / 16; i += blockDim.x) reinterpret_cast<uint4*>(sm)[i] = make_uint4(i, 2u, 3u, 4u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t s = (uint32_t)__cvta_generic_to_shared(sm);
        size_t n_chunks = bytes / chunk;
        int k = 0;
        for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x, ++k) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + c * chunk), "r"(s + (k & 1) * chunk), "r"(chunk) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

// the training forward's pattern: every CTA walks tiles; per tile it writes `layers` chunks, one per tensor.
// tensor_major: chunk (layer l, tile t) at l * n_tiles * chunk + t * chunk (the layout of `saved` today); else tile-major t * layers * chunk + l * chunk
__global__ void __launch_bounds__(128, 1) k_bulk_layers(uint8_t* __restrict__ out, size_t n_tiles, int layers, uint32_t chunk, int tensor_major) {
    extern __shared__ __align__(128) uint8_t sm[];
    for (uint32_t i = threadIdx.x; i < chunk * 2 / 16; i += blockDim.x) reinterpret_cast<uint4*>(sm)[i] = make_uint4(i, 2u, 3u, 4u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t s = (uint32_t)__cvta_generic_to_shared(sm);
        int k = 0;
        for (size_t t = blockIdx.x; t < n_tiles; t += gridDim.x)
            for (int l = 0; l < layers; ++l, ++k) {
                uint8_t* dst = tensor_major ? out + ((size_t)l * n_tiles + t) * chunk : out + (t * layers + l) * chunk;
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(s + (k & 1) * chunk), "r"(chunk) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

int main() {
    const size_t bytes = (size_t)8 << 30;
    uint8_t* buf; cudaMalloc(&buf, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    auto time = [&](const char* name, auto fn) {
        fn(); cudaDeviceSynchronize();
        cudaEventRecord(e0); for (int i = 0; i < 5; ++i) fn(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
        printf("%-44s %8.3f ms  %7.1f GB/s  (%s)\n", name, ms, bytes / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
    };
    time("cudaMemsetAsync", [&] { cudaMemsetAsync(buf, 0, bytes); });
    time("st.global.v4, 148 x 8 CTAs x 256 threads", [&] { k_stg<<<sms * 8, 256>>>((uint4*)buf, bytes / 16); });
    for (uint32_t chunk : {16384u, 65536u}) {
        cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chunk * 2);
        char name[96]; snprintf(name, sizeof name, "bulk s2g %u KB copies, 1 CTA / SM, 2 in flight", chunk >> 10);
        time(name, [&] { k_bulk<<<sms, 128, chunk * 2>>>(buf, bytes, chunk); });
        snprintf(name, sizeof name, "bulk s2g %u KB copies, 2 CTAs / SM", chunk >> 10);
        if (chunk * 4 <= 200 * 1024) time(name, [&] { k_bulk<<<sms * 2, 128, chunk * 2>>>(buf, bytes, chunk); });
    }
    {
        const uint32_t chunk = 65536; const int layers = 10;
        const size_t n_tiles = bytes / chunk / layers;
        cudaFuncSetAttribute(k_bulk_layers, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chunk * 2);
        time("10 tensors, tensor-major (today's `saved`)", [&] { k_bulk_layers<<<sms * 2, 128, chunk * 2>>>(buf, n_tiles, layers, chunk, 1); });
        time("10 tensors, tile-major", [&] { k_bulk_layers<<<sms * 2, 128, chunk * 2>>>(buf, n_tiles, layers, chunk, 0); });
    }
    return 0;
}
