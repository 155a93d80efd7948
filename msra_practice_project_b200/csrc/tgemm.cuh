// TF32 tensor-core GEMM (tcgen05 kind::tf32, fp32 accumulate in TMEM) with the same contract as sgemm.cuh:
//
//   C[i,j] (op)= sum_r P(i,r) * Q(j,r)          tile 128 (i) x 256 (j) x 32 (r), 256 threads, 3-stage pipeline
//
// It reads the SAME fp32 buffers as the CUDA-core path (tf32 MMAs take fp32 words and ignore the low 13
// mantissa bits), so the layer-wise training path switches between the two per call.  Operands are staged
// K-major (r contiguous) with SWIZZLE_128B by the CTA's own threads: global -> registers -> st.shared, transposing
// on the fly when the global layout is i-contiguous (dgrad's W, wgrad's g and x), then fence.proxy.async and an
// mbarrier hand-off to the MMA warp.  No tensor maps: leading dimensions such as 259 or 316 floats and column
// offsets inside concatenated buffers do not meet TMA's 16-byte stride rule.
//   forward   y = x W^T      : P = x [m][k],            Q = W [n][k]
//   dgrad     dx = g W       : P = g [m][n],            Q = W read as Q(j=k, r=n)      (transposed load)
//   wgrad     dW += g^T x    : P = g read as P(i=n, r=m), Q = x read as Q(j=k, r=m)    (both transposed, split over r,
//                                                                                       fp32 atomics)
#pragma once
#include "sgemm.cuh"
#include "umma.cuh"

namespace b2r {
namespace tg {

using namespace umma;

constexpr int TBM = 128, TBN = 256, TBK = 32, TSTAGES = 3;
constexpr uint32_t kAStage = TBM * TBK * 4;            // 16 KB
constexpr uint32_t kBStage = TBN * TBK * 4;            // 32 KB
constexpr uint32_t kStage = kAStage + kBStage;         // 48 KB
constexpr uint32_t kTgBarOff = TSTAGES * kStage;       // full[3], empty[3], acc, tmem slot
constexpr uint32_t kTgSmem = kTgBarOff + 128 + 1024;

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void st_shared_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }

// stage a tile of kRows (128 or 256) x 32 r into K-major SWIZZLE_128B shared memory at `dst`
//   kT = false: global is r-contiguous   (base[row * ld + r])
//   kT = true : global is row-contiguous (base[r * ld + row])
// Every thread first issues ALL of its global loads (kRows / 32 float4 = up to 8 x 16 B in flight), then stores.
template <bool kT, int kRows>
__device__ __forceinline__ void stage_tile(uint32_t dst, const float* __restrict__ base, long long ld, long long row0,
                                           long long row_max, long long r0, long long r_max, int vec, int tid) {
    constexpr int NV = kRows / 32;                     // float4 per thread
    float4 v[NV];
    if (!kT) {
        // 8 threads per row (one 16-byte chunk each), 32 rows per pass
        const int c16 = tid & 7;
#pragma unroll
        for (int it = 0; it < NV; ++it) {
            const int row = (tid >> 3) + it * 32;
            const long long gi = row0 + row, gr = r0 + c16 * 4;
            v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gi < row_max) {
                const float* p = base + gi * ld + gr;
                if (vec && gr + 3 < r_max) v[it] = __ldg(reinterpret_cast<const float4*>(p));
                else {
                    if (gr + 0 < r_max) v[it].x = p[0];
                    if (gr + 1 < r_max) v[it].y = p[1];
                    if (gr + 2 < r_max) v[it].z = p[2];
                    if (gr + 3 < r_max) v[it].w = p[3];
                }
            }
        }
#pragma unroll
        for (int it = 0; it < NV; ++it) {
            const int row = (tid >> 3) + it * 32;
            st_shared_v4(dst + sw128_offset((uint32_t)row, (uint32_t)c16), __float_as_uint(v[it].x), __float_as_uint(v[it].y),
                         __float_as_uint(v[it].z), __float_as_uint(v[it].w));
        }
    } else {
        // a float4 covers 4 consecutive rows at one r; kRows / 4 float4 per r, 32 r
        constexpr int PER_R = kRows / 4;
#pragma unroll
        for (int it = 0; it < NV; ++it) {
            const int f = tid + it * 256;
            const int r = f / PER_R, rq = (f - r * PER_R) * 4;
            const long long gr = r0 + r, gi = row0 + rq;
            v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gr < r_max) {
                const float* p = base + gr * ld + gi;
                if (vec && gi + 3 < row_max) v[it] = __ldg(reinterpret_cast<const float4*>(p));
                else {
                    if (gi + 0 < row_max) v[it].x = p[0];
                    if (gi + 1 < row_max) v[it].y = p[1];
                    if (gi + 2 < row_max) v[it].z = p[2];
                    if (gi + 3 < row_max) v[it].w = p[3];
                }
            }
        }
#pragma unroll
        for (int it = 0; it < NV; ++it) {
            const int f = tid + it * 256;
            const int r = f / PER_R, rq = (f - r * PER_R) * 4;
            const uint32_t c16 = (uint32_t)(r >> 2), sub = (uint32_t)(r & 3) * 4u;
            st_shared_f32(dst + sw128_offset((uint32_t)(rq + 0), c16) + sub, v[it].x);
            st_shared_f32(dst + sw128_offset((uint32_t)(rq + 1), c16) + sub, v[it].y);
            st_shared_f32(dst + sw128_offset((uint32_t)(rq + 2), c16) + sub, v[it].z);
            st_shared_f32(dst + sw128_offset((uint32_t)(rq + 3), c16) + sub, v[it].w);
        }
    }
}

template <bool kPT, bool kQT>
__global__ void __launch_bounds__(256, 1) tgemm_kernel(GemmArgs g) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t full = smem + kTgBarOff, empty = full + 8 * TSTAGES, acc_bar = empty + 8 * TSTAGES, slot = acc_bar + 8;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long i0 = (long long)blockIdx.x * TBM;
    const long long j0 = (long long)blockIdx.y * TBN;
    const long long r_begin = (long long)blockIdx.z * g.r_chunk;
    const long long r_end = min(g.R, r_begin + g.r_chunk);
    const int n_cols = (int)min((long long)TBN, (long long)g.J - j0);          // valid columns of this tile
    const uint32_t n_mma = (uint32_t)((n_cols + 15) & ~15);                      // UMMA N (multiple of 16)
    const int b_rows = n_mma > 128 ? 256 : 128;                                   // rows of Q staged per stage
    if (tid == 0) {
        for (int s = 0; s < TSTAGES; ++s) { mbar_init(full + 8 * s, 256); mbar_init(empty + 8 * s, 1); }
        mbar_init(acc_bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));

    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((n_mma >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t d_hi = desc_sw128(0);
    const int nk = (int)((r_end - r_begin + TBK - 1) / TBK);
    for (int kt = 0; kt < nk; ++kt) {
        const int stage = kt % TSTAGES, use = kt / TSTAGES;
        const uint32_t a_s = smem + (uint32_t)stage * kStage, b_s = a_s + kAStage;
        if (kt >= TSTAGES) mbar_wait(empty + 8 * stage, (uint32_t)((use - 1) & 1));   // MMAs of iteration kt - TSTAGES are done
        const long long r0 = r_begin + (long long)kt * TBK;
        stage_tile<kPT, TBM>(a_s, g.P, g.ldp, i0, g.I, r0, r_end, g.vec, tid);
        if (b_rows == 256) stage_tile<kQT, 256>(b_s, g.Q, g.ldq, j0, g.J, r0, r_end, g.vec, tid);
        else stage_tile<kQT, 128>(b_s, g.Q, g.ldq, j0, g.J, r0, r_end, g.vec, tid);
        fence_proxy_async_smem();
        mbar_arrive(full + 8 * stage);
        if (warp == 0) {
            mbar_wait(full + 8 * stage, (uint32_t)(use & 1));
            tc_fence_after();
            if (elect_one()) {
                const uint64_t ad = d_hi | (uint64_t)((a_s >> 4) & 0x3FFFu), bd = d_hi | (uint64_t)((b_s >> 4) & 0x3FFFu);
#pragma unroll
                for (int k = 0; k < 4; ++k) mma_tf32(tmem, ad + 2 * k, bd + 2 * k, idesc, (uint32_t)((kt | k) != 0));   // +32 B = next 8 K
                mma_commit(empty + 8 * stage);
                if (kt == nk - 1) mma_commit(acc_bar);
            }
            __syncwarp();
        }
    }
    // ---- epilogue: thread = row (TMEM lane), warps 0-3 take columns [0,128), warps 4-7 columns [128,256)
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    const int quad = warp & 3, chalf = warp >> 2;
    const long long i = i0 + quad * 32 + lane;
    const bool c_vec = (((uintptr_t)g.C & 15) == 0) && (g.ldc % 4 == 0);
    const bool m_vec = g.mask && (((uintptr_t)g.mask & 15) == 0) && (g.ldmask % 4 == 0);
    for (int jj = 0; jj < 4; ++jj) {
        const int c0 = chalf * 128 + jj * 32;
        if (c0 >= (int)n_mma) break;
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)quad << 21) + (uint32_t)c0, v);
        tmem_ld_wait();
        if (i >= g.I) continue;
#pragma unroll
        for (int e4 = 0; e4 < 32; e4 += 4) {
            const long long j = j0 + c0 + e4;
            if (j >= g.J) break;
            float val[4] = {__uint_as_float(v[e4]), __uint_as_float(v[e4 + 1]), __uint_as_float(v[e4 + 2]), __uint_as_float(v[e4 + 3])};
            float* c = g.C + i * g.ldc + j;
            const bool full4 = j + 3 < g.J;
            const bool vec_c = full4 && c_vec;                       // 16-byte path for C (and the mask, same geometry)
            float old[4] = {0.f, 0.f, 0.f, 0.f}, msk[4] = {1.f, 1.f, 1.f, 1.f}, bs[4] = {0.f, 0.f, 0.f, 0.f};
            const int nv = full4 ? 4 : (int)(g.J - j);
            if (g.epi != EPI_DGRAD && g.epi != EPI_ATOMIC && g.bias) {
                for (int q = 0; q < nv; ++q) bs[q] = g.bias[j + q];
            }
            if (g.epi == EPI_DGRAD) {
                if (g.accumulate) {
                    if (vec_c) { float4 t = *reinterpret_cast<const float4*>(c); old[0] = t.x; old[1] = t.y; old[2] = t.z; old[3] = t.w; }
                    else for (int q = 0; q < nv; ++q) old[q] = c[q];
                }
                if (g.mask) {
                    const float* mp = g.mask + i * g.ldmask + j;
                    if (full4 && m_vec) { float4 t = *reinterpret_cast<const float4*>(mp); msk[0] = t.x; msk[1] = t.y; msk[2] = t.z; msk[3] = t.w; }
                    else for (int q = 0; q < nv; ++q) msk[q] = mp[q];
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float x = val[q];
                switch (g.epi) {
                    case EPI_STORE: x = g.bias ? __fadd_rn(x, bs[q]) : x; break;
                    case EPI_RELU: x = fmaxf(__fadd_rn(x, bs[q]), 0.f); break;
                    case EPI_SIGMOID: x = 1.0f / (1.0f + expf(-__fadd_rn(x, bs[q]))); break;
                    case EPI_FILM_SIN: {
                        float a_lin = __fadd_rn(x, bs[q]);
                        if (g.pre && q < nv) g.pre[i * g.ldpre + j + q] = a_lin;
                        x = q >= nv ? 0.f : (g.gamma ? sinf(__fmul_rn(30.0f, __fadd_rn(__fmul_rn(g.gamma[j + q], a_lin), g.beta[j + q]))) : sinf(__fmul_rn(30.0f, a_lin)));
                        break;
                    }
                    case EPI_DGRAD: x = (x + old[q]) * (msk[q] > 0.f ? 1.0f : 0.f); break;
                    default: break;
                }
                val[q] = x;
            }
            if (g.epi == EPI_ATOMIC) { for (int q = 0; q < nv; ++q) atomicAdd(c + q, val[q]); }
            else if (vec_c) *reinterpret_cast<float4*>(c) = make_float4(val[0], val[1], val[2], val[3]);
            else for (int q = 0; q < nv; ++q) c[q] = val[q];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

template <bool kPT, bool kQT>
inline int launch_tgemm(GemmArgs g, cudaStream_t st, const char* what) {
    if (g.I == 0 || g.J == 0) return 0;
    long long gz = 1;
    if (g.epi == EPI_ATOMIC) gz = (g.R + g.r_chunk - 1) / g.r_chunk; else g.r_chunk = g.R;
    bool vec = aligned16(g.P) && aligned16(g.Q) && (g.ldp % 4 == 0) && (g.ldq % 4 == 0) && (g.r_chunk % 4 == 0);
    g.vec = vec ? 1 : 0;
    {   // per device, idempotent and cheap: no cached flag (the library keeps no mutable global state)
        int rc = cuda_result(cudaFuncSetAttribute(tgemm_kernel<kPT, kQT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTgSmem), "tgemm smem attribute");
        if (rc) return rc;
    }
    dim3 grid((unsigned)((g.I + TBM - 1) / TBM), (unsigned)((g.J + TBN - 1) / TBN), (unsigned)gz);
    tgemm_kernel<kPT, kQT><<<grid, 256, kTgSmem, st>>>(g);
    return cuda_result(cudaGetLastError(), what);
}

}  // namespace tg
}  // namespace b2r
