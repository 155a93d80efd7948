// fp32 CUDA-core GEMM used by the exact (fp32) MLP path and by its reverse mode.
//
//   C[i,j] (op)= sum_r P(i,r) * Q(j,r)          tile 128 x 128 x 16, 256 threads, 8x8 per thread
//
// Each operand is either r-contiguous (row-major [i][r]: kT=false) or i-contiguous ([r][i]: kT=true):
//   forward   y = x W^T      : P = x  [m][k] (false), Q = W [n][k] (false)
//   dgrad     dx = g W       : P = g  [m][n] (false), Q = W [n][k] read as Q(j=k, r=n)  (true)
//   wgrad     dW += g^T x    : P = g  [m][n] read as P(i=n, r=m) (true), Q = x [m][k] as Q(j=k, r=m) (true)
// The reduction of wgrad runs over millions of rows: grid.z splits it and partial tiles are
// combined with float atomics.
#pragma once
#include "common.cuh"

namespace b2r {

enum : int { EPI_STORE = 0, EPI_RELU = 1, EPI_SIGMOID = 2, EPI_FILM_SIN = 3, EPI_DGRAD = 4, EPI_ATOMIC = 5 };

struct GemmArgs {
    const float* P; long long ldp;
    const float* Q; long long ldq;
    float* C; long long ldc;
    long long I;            // rows of C
    int J;                  // columns of C
    long long R;            // reduction length
    long long r_chunk;      // reduction slice per blockIdx.z (EPI_ATOMIC), else == R
    int vec;                // 1: every operand row is 16-byte aligned and the contiguous extents are multiples of 4
    int epi;
    const float* bias;      // [J]                         (forward)
    const float* gamma;     // [J] FiLM                    (EPI_FILM_SIN; NULL = plain SIREN, gamma 1 / beta 0)
    const float* beta;      // [J]
    float* pre; long long ldpre;            // optional pre-activation copy (EPI_FILM_SIN, saved for backward)
    const float* mask; long long ldmask;    // EPI_DGRAD: multiply by (mask[i,j] > 0) when non-null
    int accumulate;                          // EPI_DGRAD: add the existing C before masking
    const int* row_lat; long long lat_stride;   // EPI_FILM_SIN (fp32 engine only): row i uses gamma / beta + row_lat[i] * lat_stride
};

constexpr int GBM = 128, GBN = 128, GBK = 16, GPAD = 4;

// A 128 (i) x 16 (r) operand tile travels global memory -> registers (fetch_tile: 2 float4 per thread) -> shared memory sm[r][i]
// (store_tile).  Splitting the two lets the main loop request the NEXT tile's global loads before it computes on the current one.
template <bool kT>
__device__ __forceinline__ void fetch_tile(const float* __restrict__ base, long long ld, long long i0, long long i_max,
                                           long long r0, long long r_max, int vec, int tid, float4 (&v)[2]) {
    if (!kT) {
        // r-contiguous: 512 float4 (i, 4 r) -> 2 per thread
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            int f = tid + it * 256;
            int i = f >> 2, rq = (f & 3) * 4;
            long long gi = i0 + i, gr = r0 + rq;
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gi < i_max) {
                const float* p = base + gi * ld + gr;
                if (vec && gr + 3 < r_max) t = *reinterpret_cast<const float4*>(p);
                else {
                    if (gr + 0 < r_max) t.x = p[0];
                    if (gr + 1 < r_max) t.y = p[1];
                    if (gr + 2 < r_max) t.z = p[2];
                    if (gr + 3 < r_max) t.w = p[3];
                }
            }
            v[it] = t;
        }
    } else {
        // i-contiguous: rows r (16) x 32 float4 along i -> 2 per thread
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            int f = tid + it * 256;
            int r = f >> 5, iq = (f & 31) * 4;
            long long gr = r0 + r, gi = i0 + iq;
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gr < r_max) {
                const float* p = base + gr * ld + gi;
                if (vec && gi + 3 < i_max) t = *reinterpret_cast<const float4*>(p);
                else {
                    if (gi + 0 < i_max) t.x = p[0];
                    if (gi + 1 < i_max) t.y = p[1];
                    if (gi + 2 < i_max) t.z = p[2];
                    if (gi + 3 < i_max) t.w = p[3];
                }
            }
            v[it] = t;
        }
    }
}

template <bool kT>
__device__ __forceinline__ void store_tile(float (*sm)[GBM + GPAD], int tid, const float4 (&v)[2]) {
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        int f = tid + it * 256;
        if (!kT) {
            int i = f >> 2, rq = (f & 3) * 4;
            sm[rq + 0][i] = v[it].x; sm[rq + 1][i] = v[it].y; sm[rq + 2][i] = v[it].z; sm[rq + 3][i] = v[it].w;
        } else {
            int r = f >> 5, iq = (f & 31) * 4;
            *reinterpret_cast<float4*>(&sm[r][iq]) = v[it];
        }
    }
}

// kPipe: software-pipelined main loop (the next tile's global loads are requested before the current tile is multiplied).  It needs 16
// more registers (147: one CTA per SM instead of two), which pays when the grid has at most ~2 CTAs per SM anyway -- the few-thousand-row
// GEMMs of b2r_mlp_f32_last_sigma, 0.61 -> 0.47 ms for 6,006 rows -- and costs ~7 % on large problems, which keep the plain loop.
template <bool kPT, bool kQT, bool kPipe>
__global__ void __launch_bounds__(256) sgemm_kernel(GemmArgs g) {
    __shared__ __align__(16) float Ps[GBK][GBM + GPAD];
    __shared__ __align__(16) float Qs[GBK][GBN + GPAD];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const long long i0 = (long long)blockIdx.x * GBM;
    const long long j0 = (long long)blockIdx.y * GBN;
    const long long r_begin = (long long)blockIdx.z * g.r_chunk;
    const long long r_end = min(g.R, r_begin + g.r_chunk);
    float acc[8][8];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;

    // software pipeline: the global loads of tile t + 1 are in flight while tile t is multiplied (the accumulation order of every output
    // element -- k ascending, one fma per k -- is unchanged, so results are bit-identical to the unpipelined loop)
    float4 pv_[2], qv_[2];
    if (kPipe) {
        fetch_tile<kPT>(g.P, g.ldp, i0, g.I, r_begin, r_end, g.vec, tid, pv_);
        fetch_tile<kQT>(g.Q, g.ldq, j0, g.J, r_begin, r_end, g.vec, tid, qv_);
    }
    for (long long r0 = r_begin; r0 < r_end; r0 += GBK) {
        if (!kPipe) {
            fetch_tile<kPT>(g.P, g.ldp, i0, g.I, r0, r_end, g.vec, tid, pv_);
            fetch_tile<kQT>(g.Q, g.ldq, j0, g.J, r0, r_end, g.vec, tid, qv_);
        }
        store_tile<kPT>(Ps, tid, pv_);
        store_tile<kQT>(Qs, tid, qv_);
        __syncthreads();
        if (kPipe && r0 + GBK < r_end) {
            fetch_tile<kPT>(g.P, g.ldp, i0, g.I, r0 + GBK, r_end, g.vec, tid, pv_);
            fetch_tile<kQT>(g.Q, g.ldq, j0, g.J, r0 + GBK, r_end, g.vec, tid, qv_);
        }
#pragma unroll
        for (int r = 0; r < GBK; ++r) {
            float4 p0 = *reinterpret_cast<const float4*>(&Ps[r][ty * 4]);
            float4 p1 = *reinterpret_cast<const float4*>(&Ps[r][64 + ty * 4]);
            float4 q0 = *reinterpret_cast<const float4*>(&Qs[r][tx * 4]);
            float4 q1 = *reinterpret_cast<const float4*>(&Qs[r][64 + tx * 4]);
            float pv[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
            float qv[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(pv[a], qv[b], acc[a][b]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int a = 0; a < 8; ++a) {
        long long i = i0 + (a < 4 ? ty * 4 + a : 64 + ty * 4 + (a - 4));
        if (i >= g.I) continue;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            long long j = j0 + (b < 4 ? tx * 4 + b : 64 + tx * 4 + (b - 4));
            if (j >= g.J) continue;
            float v = acc[a][b];
            float* c = g.C + i * g.ldc + j;
            switch (g.epi) {
                case EPI_STORE: *c = g.bias ? __fadd_rn(v, g.bias[j]) : v; break;
                case EPI_RELU: *c = fmaxf(__fadd_rn(v, g.bias[j]), 0.f); break;
                case EPI_SIGMOID: *c = 1.0f / (1.0f + expf(-__fadd_rn(v, g.bias[j]))); break;
                case EPI_FILM_SIN: {
                    float a_lin = __fadd_rn(v, g.bias[j]);
                    if (g.pre) g.pre[i * g.ldpre + j] = a_lin;
                    // sin(w0 * (gamma * x + beta)), w0 = 30 (pi_GAN/modules.py:22-25)
                    // gamma == NULL: plain SIREN layer sin(30 (W x + b)) (nerf/nerf.py:111-112)
                    const long long lo = g.row_lat ? (long long)g.row_lat[i] * g.lat_stride : 0;
                    *c = g.gamma ? sinf(__fmul_rn(30.0f, __fadd_rn(__fmul_rn(g.gamma[lo + j], a_lin), g.beta[lo + j]))) : sinf(__fmul_rn(30.0f, a_lin));
                    break;
                }
                case EPI_DGRAD: {
                    if (g.accumulate) v += *c;
                    if (g.mask && !(g.mask[i * g.ldmask + j] > 0.f)) v = 0.f;
                    *c = v;
                    break;
                }
                case EPI_ATOMIC: atomicAdd(c, v); break;
            }
        }
    }
}

inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

template <bool kPT, bool kQT>
inline int launch_sgemm(GemmArgs g, cudaStream_t st, const char* what) {
    if (g.I == 0 || g.J == 0) return 0;
    long long gz = 1;
    if (g.epi == EPI_ATOMIC) gz = (g.R + g.r_chunk - 1) / g.r_chunk; else g.r_chunk = g.R;
    // vector path: contiguous extents multiple of 4, leading dims multiple of 4, bases aligned
    bool vec = aligned16(g.P) && aligned16(g.Q) && (g.ldp % 4 == 0) && (g.ldq % 4 == 0);
    long long p_contig = kPT ? g.I : g.R, q_contig = kQT ? (long long)g.J : g.R;
    vec = vec && (p_contig % 4 == 0) && (q_contig % 4 == 0) && (g.r_chunk % 4 == 0);
    g.vec = vec ? 1 : 0;
    dim3 grid((unsigned)((g.I + GBM - 1) / GBM), (unsigned)((g.J + GBN - 1) / GBN), (unsigned)gz);
    if ((long long)grid.x * grid.y * grid.z <= 2 * 148) sgemm_kernel<kPT, kQT, true><<<grid, 256, 0, st>>>(g);
    else sgemm_kernel<kPT, kQT, false><<<grid, 256, 0, st>>>(g);
    return cuda_result(cudaGetLastError(), what);
}

}  // namespace b2r
