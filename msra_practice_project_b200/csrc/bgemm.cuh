// BF16 tensor-core GEMM (tcgen05 kind::f16, bf16 operands, fp32 accumulate in TMEM) with the same contract as sgemm.cuh /
// tgemm.cuh -- the engine of the layer-wise path when gemm_mode = 2 (FiLM-SIREN / SirenNeRF training in bf16 class):
//
//   C[i,j] (op)= sum_r P(i,r) * Q(j,r)          tile 128 (i) x 256 (j) x 64 (r), 256 threads, 2-stage pipeline, 2 CTAs / SM
//
// It reads the SAME fp32 buffers as the other two engines and converts to bf16 while staging (global -> registers ->
// cvt.rn.bf16x2 -> st.shared, SWIZZLE_128B).  No operand is ever transposed: a tile is always copied as
// [rows of the global matrix] x [64 contiguous columns] blocks, and the MMA reads a block either K-major (the rows are the
// M / N index: forward's x and W, dgrad's g) or MN-major (the rows are the reduction index: dgrad's W, wgrad's g and x) --
// the MN-major SWIZZLE_128B descriptors (LBO = 8 KB to the next 64 columns, SBO = 1 KB to the next 8 rows) are the ones
// PROBE4 of tests/native/umma_probe.cu pins.
//   forward   y = x W^T      : P = x [m][k] K-major,        Q = W [n][k] K-major
//   dgrad     dx = g W       : P = g [m][n] K-major,        Q = W [n][k] read as Q(j=k, r=n): MN-major
//   wgrad     dW += g^T x    : P = g [m][n] as P(i=n, r=m): MN-major,   Q = x [m][k] as Q(j=k, r=m): MN-major (split over r, fp32 atomics)
#pragma once
#include "sgemm.cuh"
#include "umma.cuh"

namespace b2r {
namespace bg {

using namespace umma;

constexpr int BBM = 128, BBN = 256, BBK = 64, BSTAGES = 2;
constexpr uint32_t kAStage = BBM * BBK * 2;            // 16 KB
constexpr uint32_t kBStage = BBN * BBK * 2;            // 32 KB
constexpr uint32_t kStage = kAStage + kBStage;         // 48 KB
constexpr uint32_t kBgBarOff = BSTAGES * kStage;       // full[2], empty[2], acc, tmem slot
constexpr uint32_t kBgSmem = kBgBarOff + 128 + 1024;   // 99,456 B -> 2 CTAs per SM

__device__ __forceinline__ float4 ld4(const float* __restrict__ p, long long have, int vec) {
    // up to 4 consecutive floats starting at p, `have` of them inside the matrix
    if (vec && have >= 4) return __ldg(reinterpret_cast<const float4*>(p));
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (have > 0) v.x = __ldg(p);
    if (have > 1) v.y = __ldg(p + 1);
    if (have > 2) v.z = __ldg(p + 2);
    if (have > 3) v.w = __ldg(p + 3);
    return v;
}

// stage a tile into SWIZZLE_128B bf16 blocks of [rows x 64 columns] (128 B per row) at `dst`
//   kT = false (K-major operand):  smem rows = kRows tile rows (i or j), smem columns = 64 r        (global base[row * ld + r])
//                                  blocks of 128 rows, 16 KB apart
//   kT = true  (MN-major operand): smem rows = 64 r, smem columns = kRows tile rows (i or j)        (global base[r * ld + row])
//                                  column blocks of 64, 8 KB apart
// Every thread handles kRows * 8 / 256 chunks of 8 values: all global loads are issued first, then converted and stored.
template <bool kT, int kRows>
__device__ __forceinline__ void stage_tile(uint32_t dst, const float* __restrict__ base, long long ld, long long row0,
                                           long long row_max, long long r0, long long r_max, int vec, int tid) {
    constexpr int NC = kRows * 8 / 256;                // 16-byte (8 x bf16) chunks per thread
    float4 lo[NC], hi[NC];
#pragma unroll
    for (int it = 0; it < NC; ++it) {
        const int f = tid + it * 256;
        long long grow, gcol, cols_left;
        bool row_ok;
        if (!kT) { const int row = f >> 3, c = f & 7; grow = row0 + row; gcol = r0 + c * 8; row_ok = grow < row_max; cols_left = r_max - gcol; }
        else { constexpr int CPR = kRows / 8; const int r = f / CPR, c = f - r * CPR; grow = r0 + r; gcol = row0 + c * 8; row_ok = grow < r_max; cols_left = row_max - gcol; }
        lo[it] = hi[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row_ok && cols_left > 0) {
            const float* p = base + grow * ld + gcol;
            lo[it] = ld4(p, cols_left, vec);
            hi[it] = ld4(p + 4, cols_left - 4, vec);
        }
    }
#pragma unroll
    for (int it = 0; it < NC; ++it) {
        const int f = tid + it * 256;
        uint32_t off;
        if (!kT) { const uint32_t row = (uint32_t)(f >> 3), c = (uint32_t)(f & 7); off = (row >> 7) * 16384u + sw128_offset(row & 127u, c); }
        else { constexpr int CPR = kRows / 8; const uint32_t r = (uint32_t)(f / CPR), c = (uint32_t)(f - (int)r * CPR); off = (c >> 3) * 8192u + sw128_offset(r, c & 7u); }
        st_shared_v4(dst + off, pack_bf16(lo[it].x, lo[it].y), pack_bf16(lo[it].z, lo[it].w), pack_bf16(hi[it].x, hi[it].y), pack_bf16(hi[it].z, hi[it].w));
    }
}

template <bool kPT, bool kQT>
__global__ void __launch_bounds__(256, 2) bgemm_kernel(GemmArgs g) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t full = smem + kBgBarOff, empty = full + 8 * BSTAGES, acc_bar = empty + 8 * BSTAGES, slot = acc_bar + 8;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long i0 = (long long)blockIdx.x * BBM;
    const long long j0 = (long long)blockIdx.y * BBN;
    const long long r_begin = (long long)blockIdx.z * g.r_chunk;
    const long long r_end = min(g.R, r_begin + g.r_chunk);
    const int n_cols = (int)min((long long)BBN, (long long)g.J - j0);          // valid columns of this tile
    const uint32_t n_mma = (uint32_t)((n_cols + 15) & ~15);                      // UMMA N (multiple of 16)
    const int b_rows = n_mma > 128 ? 256 : 128;                                   // rows of Q staged per stage
    if (tid == 0) {
        for (int s = 0; s < BSTAGES; ++s) { mbar_init(full + 8 * s, 256); mbar_init(empty + 8 * s, 1); }
        mbar_init(acc_bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));

    const uint32_t idesc = make_idesc_bf16(128, n_mma) | (kPT ? (1u << 15) : 0u) | (kQT ? (1u << 16) : 0u);
    const uint64_t dk = desc_sw128(0);                                            // K-major: SBO 1024
    const uint64_t dmn = make_desc(0, 8192, 1024, kLayoutSW128);                  // MN-major: LBO 8 KB (next 64 columns), SBO 1 KB
    const int nk = (int)((r_end - r_begin + BBK - 1) / BBK);
    for (int kt = 0; kt < nk; ++kt) {
        const int stage = kt % BSTAGES, use = kt / BSTAGES;
        const uint32_t a_s = smem + (uint32_t)stage * kStage, b_s = a_s + kAStage;
        if (kt >= BSTAGES) mbar_wait(empty + 8 * stage, (uint32_t)((use - 1) & 1));   // MMAs of iteration kt - BSTAGES are done
        const long long r0 = r_begin + (long long)kt * BBK;
        stage_tile<kPT, BBM>(a_s, g.P, g.ldp, i0, g.I, r0, r_end, g.vec, tid);
        if (b_rows == 256) stage_tile<kQT, 256>(b_s, g.Q, g.ldq, j0, g.J, r0, r_end, g.vec, tid);
        else stage_tile<kQT, 128>(b_s, g.Q, g.ldq, j0, g.J, r0, r_end, g.vec, tid);
        fence_proxy_async_smem();
        mbar_arrive(full + 8 * stage);
        if (warp == 0) {
            mbar_wait(full + 8 * stage, (uint32_t)(use & 1));
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {                                     // 4 x K=16
                    const uint32_t aa = kPT ? a_s + (uint32_t)k * 2048u : a_s + (uint32_t)k * 32u;
                    const uint32_t ba = kQT ? b_s + (uint32_t)k * 2048u : b_s + (uint32_t)k * 32u;
                    const uint64_t ad = (kPT ? dmn : dk) | (uint64_t)((aa >> 4) & 0x3FFFu);
                    const uint64_t bd = (kQT ? dmn : dk) | (uint64_t)((ba >> 4) & 0x3FFFu);
                    mma_bf16(tmem, ad, bd, idesc, (uint32_t)((kt | k) != 0));
                }
                mma_commit(empty + 8 * stage);
                if (kt == nk - 1) mma_commit(acc_bar);
            }
            __syncwarp();
        }
    }
    // ---- epilogue: warps 0-3 take columns [0,128), warps 4-7 columns [128,256) of their TMEM lane quadrant.  A TMEM load gives
    // thread = row; writing C that way touches 32 different 128-byte lines per instruction (8x the L2 requests of a coalesced
    // store: that, not the MMAs, bounded the first version).  Each 32 x 32 block is therefore transposed through a per-warp
    // scratch in the (now idle) stage buffers so that 8 consecutive lanes cover one 128-byte row segment of C / pre / mask.
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    const int quad = warp & 3, chalf = warp >> 2;
    const uint32_t scratch = smem + (uint32_t)warp * (32u * 144u);               // 32 rows x (32 + 4) floats
    const int sub_row = lane >> 3, cg = lane & 7;                                 // transposed domain: 4 rows x 8 column groups per pass
    const bool c_vec = (((uintptr_t)g.C & 15) == 0) && (g.ldc % 4 == 0);
    const bool m_vec = g.mask && (((uintptr_t)g.mask & 15) == 0) && (g.ldmask % 4 == 0);
    const bool p_vec = g.pre && (((uintptr_t)g.pre & 15) == 0) && (g.ldpre % 4 == 0);
    for (int jj = 0; jj < 4; ++jj) {
        const int c0 = chalf * 128 + jj * 32;
        if (c0 >= (int)n_mma) break;
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)quad << 21) + (uint32_t)c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q) st_shared_v4(scratch + (uint32_t)lane * 144u + (uint32_t)q * 16u, v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        __syncwarp();
        const long long j = j0 + c0 + cg * 4;
        const bool full4 = j + 3 < g.J;
        const int nv = j >= g.J ? 0 : (full4 ? 4 : (int)(g.J - j));
        float bs[4] = {0.f, 0.f, 0.f, 0.f}, gm[4] = {1.f, 1.f, 1.f, 1.f}, bt[4] = {0.f, 0.f, 0.f, 0.f};
        if (g.epi != EPI_DGRAD && g.epi != EPI_ATOMIC && g.bias) {
#pragma unroll
            for (int q = 0; q < 4; ++q) if (q < nv) bs[q] = g.bias[j + q];
        }
        if (g.epi == EPI_FILM_SIN && g.gamma) {
#pragma unroll
            for (int q = 0; q < 4; ++q) if (q < nv) { gm[q] = g.gamma[j + q]; bt[q] = g.beta[j + q]; }
        }
        // two batches of 4 row passes: the global reads of a batch (old C / relu mask of EPI_DGRAD) are all issued before any
        // of them is used -- the epilogue is otherwise a chain of dependent ~1 us loads
#pragma unroll
        for (int rb = 0; rb < 8; rb += 4) {
            float4 old4[4], msk4[4];
            if (g.epi == EPI_DGRAD) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const long long i = i0 + quad * 32 + (rb + u) * 4 + sub_row;
                    old4[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    msk4[u] = make_float4(1.f, 1.f, 1.f, 1.f);
                    if (i < g.I && nv > 0) {
                        const float* c = g.C + i * g.ldc + j;
                        if (g.accumulate) {
                            if (full4 && c_vec) old4[u] = *reinterpret_cast<const float4*>(c);
                            else { old4[u].x = c[0]; if (nv > 1) old4[u].y = c[1]; if (nv > 2) old4[u].z = c[2]; if (nv > 3) old4[u].w = c[3]; }
                        }
                        if (g.mask) {
                            const float* mp = g.mask + i * g.ldmask + j;
                            if (full4 && m_vec) msk4[u] = *reinterpret_cast<const float4*>(mp);
                            else { msk4[u].x = mp[0]; if (nv > 1) msk4[u].y = mp[1]; if (nv > 2) msk4[u].z = mp[2]; if (nv > 3) msk4[u].w = mp[3]; }
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int row = (rb + u) * 4 + sub_row;
                const long long i = i0 + quad * 32 + row;
                float4 t;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w)
                             : "r"(scratch + (uint32_t)row * 144u + (uint32_t)cg * 16u));
                if (i >= g.I || nv == 0) continue;
                float* c = g.C + i * g.ldc + j;
                const bool vec_c = full4 && c_vec;
                float val[4] = {t.x, t.y, t.z, t.w}, pre[4] = {0.f, 0.f, 0.f, 0.f};
                const float old[4] = {old4[u].x, old4[u].y, old4[u].z, old4[u].w}, msk[4] = {msk4[u].x, msk4[u].y, msk4[u].z, msk4[u].w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float x = val[q];
                    switch (g.epi) {
                        case EPI_STORE: x = g.bias ? __fadd_rn(x, bs[q]) : x; break;
                        case EPI_RELU: x = fmaxf(__fadd_rn(x, bs[q]), 0.f); break;
                        case EPI_SIGMOID: x = 1.0f / (1.0f + __expf(-__fadd_rn(x, bs[q]))); break;
                        case EPI_FILM_SIN: {
                            const float a_lin = __fadd_rn(x, bs[q]);
                            pre[q] = a_lin;
                            // MUFU sine: |abs error| < 1e-5 for the |arguments| < ~100 of these networks, far below the bf16 rounding of
                            // the operands this engine works with (the accurate libdevice sinf, inlined 32 times, made the kernel
                            // instruction-cache bound: 41 % of the samples were "no instruction")
                            x = g.gamma ? __sinf(__fmul_rn(30.0f, __fadd_rn(__fmul_rn(gm[q], a_lin), bt[q]))) : __sinf(__fmul_rn(30.0f, a_lin));
                            break;
                        }
                        case EPI_DGRAD: x = (x + old[q]) * (msk[q] > 0.f ? 1.0f : 0.f); break;
                        default: break;
                    }
                    val[q] = x;
                }
                if (g.epi == EPI_FILM_SIN && g.pre) {
                    float* pp = g.pre + i * g.ldpre + j;
                    if (full4 && p_vec) *reinterpret_cast<float4*>(pp) = make_float4(pre[0], pre[1], pre[2], pre[3]);
                    else { pp[0] = pre[0]; if (nv > 1) pp[1] = pre[1]; if (nv > 2) pp[2] = pre[2]; if (nv > 3) pp[3] = pre[3]; }
                }
                if (g.epi == EPI_ATOMIC) { atomicAdd(c, val[0]); if (nv > 1) atomicAdd(c + 1, val[1]); if (nv > 2) atomicAdd(c + 2, val[2]); if (nv > 3) atomicAdd(c + 3, val[3]); }
                else if (vec_c) *reinterpret_cast<float4*>(c) = make_float4(val[0], val[1], val[2], val[3]);
                else { c[0] = val[0]; if (nv > 1) c[1] = val[1]; if (nv > 2) c[2] = val[2]; if (nv > 3) c[3] = val[3]; }
            }
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

template <bool kPT, bool kQT>
inline int launch_bgemm(GemmArgs g, cudaStream_t st, const char* what) {
    if (g.I == 0 || g.J == 0) return 0;
    long long gz = 1;
    if (g.epi == EPI_ATOMIC) gz = (g.R + g.r_chunk - 1) / g.r_chunk; else g.r_chunk = g.R;
    // 16-byte loads need every touched row start and column offset 16-byte aligned
    bool vec = aligned16(g.P) && aligned16(g.Q) && (g.ldp % 4 == 0) && (g.ldq % 4 == 0);
    g.vec = vec ? 1 : 0;
    {
        int rc = cuda_result(cudaFuncSetAttribute(bgemm_kernel<kPT, kQT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBgSmem), "bgemm smem attribute");
        if (rc) return rc;
    }
    dim3 grid((unsigned)((g.I + BBM - 1) / BBM), (unsigned)((g.J + BBN - 1) / BBN), (unsigned)gz);
    bgemm_kernel<kPT, kQT><<<grid, 256, kBgSmem, st>>>(g);
    return cuda_result(cudaGetLastError(), what);
}

}  // namespace bg
}  // namespace b2r
