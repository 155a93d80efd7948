// K4 alpha compositing, forward and reverse mode.
//   raw_to_outputs  nerf/render.py:78-103   (== pi_GAN/render.py:123-148)
//   backward        autograd in the reference (nerf/train_nerf.py:167); closed form SURVEY.md A.3
//
// One warp owns one ray.  Samples are processed in chunks of 32 (lane = sample), so every raw
// load is a coalesced 512-byte float4 row and the transmittance T_k = prod_{j<k} (1-alpha_j+1e-10)
// is a warp-level product scan (5 shuffles per chunk) with a carry between chunks.  The kernel is
// HBM-bound: S*20+12 bytes in, S*4+20 bytes out per ray (S*20 bytes in / 20 out when the weights
// are not requested).  Grid = 148 SMs x 8 CTAs x 8 warps, rays are walked with a grid stride.
#include "common.cuh"

namespace b2r {

__device__ __forceinline__ float ray_norm(const float* __restrict__ d) {
    float x = d[0], y = d[1], z = d[2];
    return sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
}

// inclusive product scan over the 32 lanes
__device__ __forceinline__ float warp_scan_mul(float p, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float t = __shfl_up_sync(kFull, p, o);
        if (lane >= o) p *= t;
    }
    return p;
}

// Forward: every lane owns TWO consecutive samples of a 64-sample chunk (32 contiguous bytes of raw per lane,
// 1 KB per warp-load), so the transmittance scan costs 5 shuffle steps per 64 samples; the next chunk's loads
// are issued before the current one is reduced.  exp() is ex2.approx (|error| < 2e-7 on alpha).
template <bool kWeights>
__global__ void __launch_bounds__(256) composite_fwd_kernel(
    const float4* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ rays_d, int d_stride,
    long long n_rays, int S, float* __restrict__ rgb_out, int rgb_stride, float* __restrict__ depth_out, int depth_stride,
    float* __restrict__ acc_out, int acc_stride, float* __restrict__ w_out) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long ray = warp0; ray < n_rays; ray += n_warps) {
        const float norm = ray_norm(rays_d + ray * d_stride);
        const float4* rr = raw + ray * S;
        const float* zz = z + ray * S;
        float carry = 1.0f;
        float ar = 0.f, ag = 0.f, ab = 0.f, ad = 0.f, aa = 0.f;
        int k = 2 * lane;
        float4 r0 = k < S ? rr[k] : zero4, r1 = k + 1 < S ? rr[k + 1] : zero4;
        float z0 = k < S ? zz[k] : 0.f, z1 = k + 1 < S ? zz[k + 1] : 0.f;
        float ze = (lane == 31 && 64 < S) ? zz[64] : 0.f;          // first z of the next chunk (needed by lane 31 only)
        for (int c0 = 0; c0 < S; c0 += 64) {
            const int kn = c0 + 64 + 2 * lane;
            float4 r0n = kn < S ? rr[kn] : zero4, r1n = kn + 1 < S ? rr[kn + 1] : zero4;
            float z0n = kn < S ? zz[kn] : 0.f, z1n = kn + 1 < S ? zz[kn + 1] : 0.f;
            float zen = (lane == 31 && c0 + 128 < S) ? zz[c0 + 128] : 0.f;
            float z2 = __shfl_down_sync(kFull, z0, 1);                 // z[k+2] = next lane's first sample
            if (lane == 31) z2 = ze;
            const bool v0 = k < S, v1 = k + 1 < S;
            float d0 = (k == S - 1) ? 1e10f : __fsub_rn(z1, z0);
            float d1 = (k + 1 == S - 1) ? 1e10f : __fsub_rn(z2, z1);
            d0 = __fmul_rn(d0, norm); d1 = __fmul_rn(d1, norm);
            float a0 = v0 ? 1.0f - __expf(-r0.w * d0) : 0.f;
            float a1 = v1 ? 1.0f - __expf(-r1.w * d1) : 0.f;
            float q0 = v0 ? (1.0f - a0) + 1e-10f : 1.0f;
            float q1 = v1 ? (1.0f - a1) + 1e-10f : 1.0f;
            float p = warp_scan_mul(q0 * q1, lane);
            float excl = __shfl_up_sync(kFull, p, 1);
            if (lane == 0) excl = 1.0f;
            const float t0 = carry * excl, t1 = t0 * q0;
            const float w0 = a0 * t0, w1 = a1 * t1;
            carry *= __shfl_sync(kFull, p, 31);
            ar = fmaf(w0, r0.x, fmaf(w1, r1.x, ar)); ag = fmaf(w0, r0.y, fmaf(w1, r1.y, ag)); ab = fmaf(w0, r0.z, fmaf(w1, r1.z, ab));
            ad = fmaf(w0, z0, fmaf(w1, z1, ad)); aa += w0 + w1;
            if (kWeights) {
                if (v0) w_out[ray * S + k] = w0;
                if (v1) w_out[ray * S + k + 1] = w1;
            }
            r0 = r0n; r1 = r1n; z0 = z0n; z1 = z1n; ze = zen; k = kn;
        }
        ar = warp_sum(ar); ag = warp_sum(ag); ab = warp_sum(ab); ad = warp_sum(ad); aa = warp_sum(aa);
        if (lane == 0) {
            float bg = 1.0f - aa;                          // white background, always (render.py:101)
            float* c = rgb_out + ray * rgb_stride;
            c[0] = ar + bg; c[1] = ag + bg; c[2] = ab + bg;
            depth_out[ray * depth_stride] = ad;
            acc_out[ray * acc_stride] = aa;
        }
    }
}


// Single-chunk forward for the shapes that matter (S <= L x LANES): a group of LANES lanes (16: two rays per warp; 32: one) owns a
// ray and every lane owns L CONSECUTIVE samples -- L x 16 contiguous bytes of raw, L x 4 of z and of the weights, moved as 16-byte
// vectors -- so the transmittance is L - 1 serial products per lane plus ONE log2(LANES)-step scan per ray, and the five per-ray
// sums are reduced once.  For S = 64 this is 2 rays per warp pass and about half the warp-instructions per ray of the chunked
// kernel above (which stays as the fallback for other sizes).  Same arithmetic per sample; the product scan associates
// differently, which moves T by a few ulp (tests: <= 1e-5 against the reference's golden outputs).
// Outputs are strided so that a caller can have (rgb, depth, acc) written straight into rows of a packed [N,5] buffer
// (the multi-GPU image gather, dist.py).
template <int L, int LANES, bool kWeights>
__global__ void __launch_bounds__(256) composite_fwd_group_kernel(
    const float4* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ rays_d, int d_stride,
    long long n_rays, int S, float* __restrict__ rgb_out, int rgb_stride, float* __restrict__ depth_out, int depth_stride,
    float* __restrict__ acc_out, int acc_stride, float* __restrict__ w_out) {
    constexpr int RPW = 32 / LANES;                       // rays per warp pass
    const int lane = threadIdx.x & 31;
    const int gl = lane & (LANES - 1);                    // lane within the ray's group
    const int sub = lane / LANES;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const bool vec = (S % 4) == 0 && L % 4 == 0;          // 16-byte z loads / weight stores are aligned
    for (long long ray0 = warp0 * RPW; ray0 < n_rays; ray0 += n_warps * RPW) {
        const long long ray = ray0 + sub;
        const bool live = ray < n_rays;
        const long long rr_ = live ? ray : n_rays - 1;
        const float norm = ray_norm(rays_d + rr_ * d_stride);
        const float4* rr = raw + rr_ * S;
        const float* zz = z + rr_ * S;
        const int k0 = gl * L;
        float4 r[L];
        float zv[L + 1];
#pragma unroll
        for (int j = 0; j < L; ++j) r[j] = (k0 + j < S) ? rr[k0 + j] : make_float4(0.f, 0.f, 0.f, 0.f);
        if (vec) {
#pragma unroll
            for (int j = 0; j < L; j += 4) {
                float4 t = (k0 + j < S) ? *reinterpret_cast<const float4*>(zz + k0 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                zv[j] = t.x; zv[j + 1] = t.y; zv[j + 2] = t.z; zv[j + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < L; ++j) zv[j] = (k0 + j < S) ? zz[k0 + j] : 0.f;
        }
        zv[L] = __shfl_down_sync(kFull, zv[0], 1, LANES);       // first z of the next lane's run
        float a[L], q[L], pl = 1.0f;
#pragma unroll
        for (int j = 0; j < L; ++j) {
            const int k = k0 + j;
            const bool v = k < S;
            float d = (k == S - 1) ? 1e10f : __fsub_rn(zv[j + 1], zv[j]);
            d = __fmul_rn(d, norm);
            a[j] = v ? 1.0f - __expf(-r[j].w * d) : 0.f;
            q[j] = v ? (1.0f - a[j]) + 1e-10f : 1.0f;
            pl *= q[j];
        }
        float p = pl;                                          // inclusive product scan over the group's lanes
#pragma unroll
        for (int o = 1; o < LANES; o <<= 1) {
            float t = __shfl_up_sync(kFull, p, o, LANES);
            if (gl >= o) p *= t;
        }
        float T = __shfl_up_sync(kFull, p, 1, LANES);
        if (gl == 0) T = 1.0f;
        float ar = 0.f, ag = 0.f, ab = 0.f, ad = 0.f, aa = 0.f, w[L];
#pragma unroll
        for (int j = 0; j < L; ++j) {
            w[j] = a[j] * T;
            T *= q[j];
            ar = fmaf(w[j], r[j].x, ar); ag = fmaf(w[j], r[j].y, ag); ab = fmaf(w[j], r[j].z, ab);
            ad = fmaf(w[j], zv[j], ad); aa += w[j];
        }
        if (kWeights && live) {
            float* wo = w_out + ray * S + k0;
            if (vec) {
#pragma unroll
                for (int j = 0; j < L; j += 4)
                    if (k0 + j < S) *reinterpret_cast<float4*>(wo + j) = make_float4(w[j], w[j + 1], w[j + 2], w[j + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < L; ++j)
                    if (k0 + j < S) wo[j] = w[j];
            }
        }
#pragma unroll
        for (int o = LANES / 2; o > 0; o >>= 1) {
            ar += __shfl_xor_sync(kFull, ar, o); ag += __shfl_xor_sync(kFull, ag, o); ab += __shfl_xor_sync(kFull, ab, o);
            ad += __shfl_xor_sync(kFull, ad, o); aa += __shfl_xor_sync(kFull, aa, o);
        }
        if (gl == 0 && live) {
            const float bg = 1.0f - aa;                    // white background, always (render.py:101)
            float* c = rgb_out + ray * rgb_stride;
            c[0] = ar + bg; c[1] = ag + bg; c[2] = ab + bg;
            depth_out[ray * depth_stride] = ad;
            acc_out[ray * acc_stride] = aa;
        }
    }
}

// Reverse mode.  Pass 1 walks the chunks forward keeping (e=1-alpha, T, w, gw, dist) of this
// lane's samples in registers; pass 2 walks them backward with a suffix-sum scan of gw*w.
//   gw_k = sum_ch gC_ch (c_k,ch - 1) + gA + gD z_k ;  gc_k = w_k gC
//   R_k = sum_{j>k} gw_j w_j ;  galpha_k = gw_k T_k - R_k / q_k ;  gsigma_k = galpha_k dist_k e_k
template <int CH>
__global__ void __launch_bounds__(256) composite_bwd_kernel(
    const float4* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ rays_d, int d_stride,
    long long n_rays, int S, const float* __restrict__ g_rgb, const float* __restrict__ g_depth,
    const float* __restrict__ g_acc, float4* __restrict__ d_raw) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long ray = warp0; ray < n_rays; ray += n_warps) {
        const float norm = ray_norm(rays_d + ray * d_stride);
        const float4* rr = raw + ray * S;
        const float* zz = z + ray * S;
        const float gr = g_rgb[ray * 3 + 0], gg = g_rgb[ray * 3 + 1], gb = g_rgb[ray * 3 + 2];
        const float gd = g_depth ? g_depth[ray] : 0.f;
        const float ga = g_acc ? g_acc[ray] : 0.f;
        float e[CH], T[CH], w[CH], gw[CH], dist[CH];
        float carry = 1.0f;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            int k = c * 32 + lane;
            bool valid = k < S;
            float4 r = valid ? rr[k] : make_float4(0.f, 0.f, 0.f, 0.f);
            float zk = valid ? zz[k] : 0.f;
            float zn = k + 1 < S ? zz[k + 1] : 0.f;
            float dd = (k == S - 1) ? 1e10f : __fsub_rn(zn, zk);
            dd = __fmul_rn(dd, norm);
            float ee = valid ? expf(-__fmul_rn(r.w, dd)) : 1.0f;
            float alpha = __fsub_rn(1.0f, ee);
            float q = valid ? __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f) : 1.0f;
            float p = warp_scan_mul(q, lane);
            float excl = __shfl_up_sync(kFull, p, 1);
            if (lane == 0) excl = 1.0f;
            float t = carry * excl;
            carry *= __shfl_sync(kFull, p, 31);
            e[c] = ee; T[c] = t; w[c] = valid ? alpha * t : 0.f; dist[c] = dd;
            gw[c] = valid ? (gr * (r.x - 1.0f) + gg * (r.y - 1.0f) + gb * (r.z - 1.0f) + ga + gd * zk) : 0.f;
        }
        float carry_r = 0.f;
#pragma unroll
        for (int c = CH - 1; c >= 0; --c) {
            int k = c * 32 + lane;
            float s = gw[c] * w[c];
            float p = s;                                   // inclusive suffix sum over lanes
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                float t = __shfl_down_sync(kFull, p, o);
                if (lane + o < 32) p += t;
            }
            float R = (p - s) + carry_r;
            carry_r += __shfl_sync(kFull, p, 0);
            if (k < S) {
                float q = __fadd_rn(e[c], 1e-10f);         // 1 - alpha + 1e-10 with 1-alpha == e up to rounding
                float g_alpha = gw[c] * T[c] - R / q;
                float g_sigma = g_alpha * dist[c] * e[c];
                d_raw[ray * S + k] = make_float4(w[c] * gr, w[c] * gg, w[c] * gb, g_sigma);
            }
        }
    }
}


// Training: loss + reverse mode of the composite in ONE kernel (nerf/train_nerf.py:157-167).  The reference computes
//   loss = mean((rgb - target)^2) [+ 0.1 * mean((acc - target_alpha)^2)]      over the GLOBAL batch
// with ~40 elementwise / reduction launches between raw_to_outputs and backward(); here pass 1 recomputes the forward of the ray
// (as composite_bwd_kernel does), the upstream gradients follow from the ray's own rgb / acc
//   g_rgb = 2 (rgb - target) inv_count / 3,   g_acc = 2 alpha_weight (acc - target_alpha) inv_count,   g_depth = 0
// and pass 2 is the reverse scan.  ray_weight (nullable): 0 for the padded rays of a short last batch.
// sums[0] += sum_rays w sum_ch (rgb - target)^2, sums[1] += sum_rays w (acc - target_alpha)^2 (one atomic per CTA).
template <int CH>
__global__ void __launch_bounds__(256) composite_loss_bwd_kernel(
    const float4* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ rays_d, int d_stride,
    long long n_rays, int S, const float* __restrict__ target_rgb, const float* __restrict__ target_acc,
    const float* __restrict__ ray_weight, const float* __restrict__ inv_count_p, float alpha_weight,
    float4* __restrict__ d_raw, float* __restrict__ sums) {
    __shared__ float s_part[2][8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const float inv_count = *inv_count_p;
    float loss_rgb = 0.f, loss_acc = 0.f;
    for (long long ray = warp0; ray < n_rays; ray += n_warps) {
        const float norm = ray_norm(rays_d + ray * d_stride);
        const float4* rr = raw + ray * S;
        const float* zz = z + ray * S;
        float e[CH], T[CH], w[CH], dist[CH], cr[CH], cg[CH], cb[CH];
        float carry = 1.0f, ar = 0.f, ag = 0.f, ab = 0.f, aa = 0.f;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            int k = c * 32 + lane;
            bool valid = k < S;
            float4 r = valid ? rr[k] : make_float4(0.f, 0.f, 0.f, 0.f);
            float zk = valid ? zz[k] : 0.f;
            float zn = k + 1 < S ? zz[k + 1] : 0.f;
            float dd = (k == S - 1) ? 1e10f : __fsub_rn(zn, zk);
            dd = __fmul_rn(dd, norm);
            float ee = valid ? expf(-__fmul_rn(r.w, dd)) : 1.0f;
            float alpha = __fsub_rn(1.0f, ee);
            float q = valid ? __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f) : 1.0f;
            float p = warp_scan_mul(q, lane);
            float excl = __shfl_up_sync(kFull, p, 1);
            if (lane == 0) excl = 1.0f;
            float t = carry * excl;
            carry *= __shfl_sync(kFull, p, 31);
            e[c] = ee; T[c] = t; w[c] = valid ? alpha * t : 0.f; dist[c] = dd;
            cr[c] = r.x; cg[c] = r.y; cb[c] = r.z;
            ar = fmaf(w[c], r.x, ar); ag = fmaf(w[c], r.y, ag); ab = fmaf(w[c], r.z, ab); aa += w[c];
        }
        ar = warp_sum(ar); ag = warp_sum(ag); ab = warp_sum(ab); aa = warp_sum(aa);
        const float bg = 1.0f - aa;
        const float wt = ray_weight ? ray_weight[ray] : 1.0f;
        const float er = ar + bg - target_rgb[ray * 3 + 0], eg = ag + bg - target_rgb[ray * 3 + 1], eb = ab + bg - target_rgb[ray * 3 + 2];
        const float ea = target_acc ? aa - target_acc[ray] : 0.f;
        const float sc_rgb = 2.0f * inv_count * (1.0f / 3.0f) * wt;
        const float gr = er * sc_rgb, gg = eg * sc_rgb, gb = eb * sc_rgb;
        const float ga = target_acc ? 2.0f * alpha_weight * inv_count * wt * ea : 0.f;
        if (lane == 0) { loss_rgb += wt * (er * er + eg * eg + eb * eb); loss_acc += wt * ea * ea; }
        float carry_r = 0.f;
#pragma unroll
        for (int c = CH - 1; c >= 0; --c) {
            int k = c * 32 + lane;
            const float gw = (k < S) ? (gr * (cr[c] - 1.0f) + gg * (cg[c] - 1.0f) + gb * (cb[c] - 1.0f) + ga) : 0.f;
            float s = gw * w[c];
            float p = s;                                   // inclusive suffix sum over lanes
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                float t = __shfl_down_sync(kFull, p, o);
                if (lane + o < 32) p += t;
            }
            float R = (p - s) + carry_r;
            carry_r += __shfl_sync(kFull, p, 0);
            if (k < S) {
                float q = __fadd_rn(e[c], 1e-10f);
                float g_alpha = gw * T[c] - R / q;
                float g_sigma = g_alpha * dist[c] * e[c];
                d_raw[ray * S + k] = make_float4(w[c] * gr, w[c] * gg, w[c] * gb, g_sigma);
            }
        }
    }
    if (lane == 0) { s_part[0][wid] = loss_rgb; s_part[1][wid] = loss_acc; }
    __syncthreads();
    if (threadIdx.x < 2) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += s_part[threadIdx.x][i];
        if (t != 0.f) atomicAdd(sums + threadIdx.x, t);
    }
}

static inline unsigned ray_grid(long long n_rays) {
    long long want = (n_rays + 7) / 8;             // 8 warps per CTA
    long long cap = 148LL * 8;
    return (unsigned)(want < 1 ? 1 : (want > cap ? cap : want));
}

}  // namespace b2r

namespace b2r {
template <int L, int LANES>
static void launch_group(bool weights, unsigned grid, cudaStream_t st, const float* raw, const float* z, const float* rays_d, int d_stride,
                         long long n_rays, int S, float* rgb, int rs, float* depth, int ds, float* acc, int as, float* w) {
    if (weights)
        composite_fwd_group_kernel<L, LANES, true><<<grid, 256, 0, st>>>((const float4*)raw, z, rays_d, d_stride, n_rays, S, rgb, rs, depth, ds, acc, as, w);
    else
        composite_fwd_group_kernel<L, LANES, false><<<grid, 256, 0, st>>>((const float4*)raw, z, rays_d, d_stride, n_rays, S, rgb, rs, depth, ds, acc, as, nullptr);
}
}  // namespace b2r

extern "C" int b2r_composite_fwd_strided(const float* raw, const float* z, const float* rays_d, int d_stride,
                                         long long n_rays, int n_samples, float* rgb_out, int rgb_stride, float* depth_out,
                                         int depth_stride, float* acc_out, int acc_stride, float* weights_out, void* stream) {
    using namespace b2r;
    B2R_CHECK_ARG(raw && z && rays_d && rgb_out && depth_out && acc_out, "b2r_composite_fwd: NULL pointer");
    B2R_CHECK_ARG(n_rays >= 0 && n_samples >= 1, "b2r_composite_fwd: need n_rays >= 0, n_samples >= 1");
    B2R_CHECK_ARG(d_stride >= 3, "b2r_composite_fwd: d_stride must be >= 3");
    B2R_CHECK_ARG(rgb_stride >= 3 && depth_stride >= 1 && acc_stride >= 1, "b2r_composite_fwd: output strides must be >= 3 / 1 / 1");
    B2R_CHECK_ARG(((uintptr_t)raw & 15) == 0, "b2r_composite_fwd: raw must be 16-byte aligned");
    if (n_rays == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int S = n_samples;
    const bool w = weights_out != nullptr;
    // 16-byte z loads / weight stores of the grouped kernels need aligned rows
    const bool al = (((uintptr_t)z | (uintptr_t)weights_out) & 15) == 0;
    // S <= 128: grouped single-chunk kernel.  Longer rays keep the chunked kernel: with 8 samples per lane a warp-load touches 32
    // different 128-byte lines (measured at S = 192: 0.69 ms against 0.45 ms for the 640,000-ray frame).
    if (S <= 128 && (al || S % 4 != 0)) {
        const unsigned g2 = ray_grid((n_rays + 1) / 2), g1 = ray_grid(n_rays);
        if (S <= 64) launch_group<4, 16>(w, g2, st, raw, z, rays_d, d_stride, n_rays, S, rgb_out, rgb_stride, depth_out, depth_stride, acc_out, acc_stride, weights_out);
        else launch_group<4, 32>(w, g1, st, raw, z, rays_d, d_stride, n_rays, S, rgb_out, rgb_stride, depth_out, depth_stride, acc_out, acc_stride, weights_out);
    } else {
        unsigned grid = ray_grid(n_rays);
        if (w)
            composite_fwd_kernel<true><<<grid, 256, 0, st>>>((const float4*)raw, z, rays_d, d_stride, n_rays, S, rgb_out, rgb_stride, depth_out, depth_stride, acc_out, acc_stride, weights_out);
        else
            composite_fwd_kernel<false><<<grid, 256, 0, st>>>((const float4*)raw, z, rays_d, d_stride, n_rays, S, rgb_out, rgb_stride, depth_out, depth_stride, acc_out, acc_stride, nullptr);
    }
    B2R_LAUNCH_CHECK("b2r_composite_fwd");
    return 0;
}

extern "C" int b2r_composite_fwd(const float* raw, const float* z, const float* rays_d, int d_stride,
                                 long long n_rays, int n_samples, float* rgb_out, float* depth_out,
                                 float* acc_out, float* weights_out, void* stream) {
    return b2r_composite_fwd_strided(raw, z, rays_d, d_stride, n_rays, n_samples, rgb_out, 3, depth_out, 1, acc_out, 1, weights_out, stream);
}

extern "C" int b2r_composite_bwd(const float* raw, const float* z, const float* rays_d, int d_stride,
                                 long long n_rays, int n_samples, const float* g_rgb, const float* g_depth,
                                 const float* g_acc, float* d_raw, void* stream) {
    using namespace b2r;
    B2R_CHECK_ARG(raw && z && rays_d && g_rgb && d_raw, "b2r_composite_bwd: NULL pointer");
    B2R_CHECK_ARG(n_rays >= 0 && n_samples >= 1 && n_samples <= 512, "b2r_composite_bwd: need 1 <= n_samples <= 512");
    B2R_CHECK_ARG(d_stride >= 3, "b2r_composite_bwd: d_stride must be >= 3");
    B2R_CHECK_ARG((((uintptr_t)raw | (uintptr_t)d_raw) & 15) == 0, "b2r_composite_bwd: raw / d_raw must be 16-byte aligned");
    if (n_rays == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned grid = ray_grid(n_rays);
    int ch = (n_samples + 31) / 32;
#define B2R_BWD(CH)                                                                                          \
    composite_bwd_kernel<CH><<<grid, 256, 0, st>>>((const float4*)raw, z, rays_d, d_stride, n_rays, n_samples, \
                                                   g_rgb, g_depth, g_acc, (float4*)d_raw)
    if (ch <= 1) B2R_BWD(1);
    else if (ch <= 2) B2R_BWD(2);
    else if (ch <= 3) B2R_BWD(3);
    else if (ch <= 4) B2R_BWD(4);
    else if (ch <= 6) B2R_BWD(6);
    else if (ch <= 8) B2R_BWD(8);
    else if (ch <= 12) B2R_BWD(12);
    else B2R_BWD(16);
#undef B2R_BWD
    B2R_LAUNCH_CHECK("b2r_composite_bwd");
    return 0;
}

extern "C" int b2r_composite_loss_bwd(const float* raw, const float* z, const float* rays_d, int d_stride, long long n_rays, int n_samples,
                                      const float* target_rgb, const float* target_acc, const float* ray_weight, const float* inv_count,
                                      float alpha_weight, float* d_raw, float* sums, void* stream) {
    using namespace b2r;
    B2R_CHECK_ARG(raw && z && rays_d && target_rgb && inv_count && d_raw && sums, "b2r_composite_loss_bwd: NULL pointer");
    B2R_CHECK_ARG(n_rays >= 0 && n_samples >= 1 && n_samples <= 512, "b2r_composite_loss_bwd: need 1 <= n_samples <= 512");
    B2R_CHECK_ARG(d_stride >= 3, "b2r_composite_loss_bwd: d_stride must be >= 3");
    B2R_CHECK_ARG((((uintptr_t)raw | (uintptr_t)d_raw) & 15) == 0, "b2r_composite_loss_bwd: raw / d_raw must be 16-byte aligned");
    if (n_rays == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned grid = ray_grid(n_rays);
    int ch = (n_samples + 31) / 32;
#define B2R_LBWD(CH)                                                                                               \
    composite_loss_bwd_kernel<CH><<<grid, 256, 0, st>>>((const float4*)raw, z, rays_d, d_stride, n_rays, n_samples, \
                                                        target_rgb, target_acc, ray_weight, inv_count, alpha_weight, (float4*)d_raw, sums)
    if (ch <= 1) B2R_LBWD(1);
    else if (ch <= 2) B2R_LBWD(2);
    else if (ch <= 3) B2R_LBWD(3);
    else if (ch <= 4) B2R_LBWD(4);
    else if (ch <= 6) B2R_LBWD(6);
    else if (ch <= 8) B2R_LBWD(8);
    else if (ch <= 12) B2R_LBWD(12);
    else B2R_LBWD(16);
#undef B2R_LBWD
    B2R_LAUNCH_CHECK("b2r_composite_loss_bwd");
    return 0;
}

namespace b2r {
// loss = (sums[0] + sums[2]) inv_count / 3 + alpha_weight (sums[1] + sums[3]) inv_count      (train_nerf.py:158-166; [0..1] fine, [2..3] coarse)
// psnr = -10 log10(sums[0] inv_local / 3)                                                    (train_nerf.py:160: fine pass, this rank's rays)
__global__ void train_loss_finish_kernel(const float* __restrict__ sums, const float* __restrict__ inv_count, float alpha_weight,
                                         const float* __restrict__ inv_local, float* __restrict__ loss, float* __restrict__ psnr) {
    const float ic = *inv_count;
    *loss = (sums[0] + sums[2]) * ic * (1.0f / 3.0f) + alpha_weight * (sums[1] + sums[3]) * ic;
    *psnr = -10.0f * log10f(sums[0] * (*inv_local) * (1.0f / 3.0f));
}
}  // namespace b2r

extern "C" int b2r_train_loss_finish(const float* sums, const float* inv_count, float alpha_weight, const float* inv_local, float* loss,
                                     float* psnr, void* stream) {
    using namespace b2r;
    B2R_CHECK_ARG(sums && inv_count && inv_local && loss && psnr, "b2r_train_loss_finish: NULL pointer");
    train_loss_finish_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(sums, inv_count, alpha_weight, inv_local, loss, psnr);
    B2R_LAUNCH_CHECK("b2r_train_loss_finish");
    return 0;
}
