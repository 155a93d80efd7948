// K1 ray generation and K2 stratified coarse samples.
//   get_rays      nerf/render.py:7-23   (== pi_GAN/render.py:52-68)
//   render_rays   nerf/render.py:123-132 (linspace / mids / upper / lower / jitter lerp)
// Both are pure streaming kernels (HBM-write bound): one thread per output row / element,
// 148 SMs x several CTAs, fully coalesced stores.
#include "common.cuh"

namespace b2r {

template <typename T>
struct Cam {
    T r[9];    // c2w[:3,:3] row-major
    T t[3];    // c2w[:3,3]
};

// Each op is rounded separately (numpy evaluates mul / sum as separate float32 ufuncs, no FMA):
// dirs = [(i - W/2)/f, -(j - H/2)/f, -1];  d_k = sum_m dirs_m * R[k,m]  (left-to-right)
// one thread per output float: element e = ray*6 + c, c < 3 origin, c >= 3 direction component (coalesced stores)
__global__ void raygen_f32_kernel(Cam<float> cam, int width, float half_w, float half_h, float focal,
                                  long long begin, long long count, float* __restrict__ rays) {
    long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= count * 6) return;
    long long t = e / 6;
    int c = (int)(e - t * 6);
    float v;
    if (c < 3) v = c == 0 ? cam.t[0] : (c == 1 ? cam.t[1] : cam.t[2]);
    else {
        long long pix = begin + t;
        float i = (float)(pix % width), j = (float)(pix / width);
        float dx = __fdiv_rn(__fsub_rn(i, half_w), focal);
        float dy = __fdiv_rn(-__fsub_rn(j, half_h), focal);
        int k = c - 3;
        float r0 = k == 0 ? cam.r[0] : (k == 1 ? cam.r[3] : cam.r[6]);
        float r1 = k == 0 ? cam.r[1] : (k == 1 ? cam.r[4] : cam.r[7]);
        float r2 = k == 0 ? cam.r[2] : (k == 1 ? cam.r[5] : cam.r[8]);
        float s = __fadd_rn(__fmul_rn(dx, r0), __fmul_rn(dy, r1));
        v = __fadd_rn(s, __fmul_rn(-1.0f, r2));
    }
    rays[e] = v;
}

// np.float64 focal (pi_GAN/modules.py:127): (i - W/2) is still float32, the division and
// everything after it are float64; the caller rounds to float32 (pi_GAN/render.py:202).
__global__ void raygen_f64_kernel(Cam<double> cam, int width, float half_w, float half_h, double focal,
                                  int div_f64, long long begin, long long count, float* __restrict__ rays) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= count) return;
    long long pix = begin + t;
    float i = (float)(pix % width), j = (float)(pix / width);
    double dx, dy;
    if (div_f64) {
        dx = __ddiv_rn((double)__fsub_rn(i, half_w), focal);
        dy = __ddiv_rn((double)(-__fsub_rn(j, half_h)), focal);
    } else {   // float32 dirs (python-float focal), float64 pose
        dx = (double)__fdiv_rn(__fsub_rn(i, half_w), (float)focal);
        dy = (double)__fdiv_rn(-__fsub_rn(j, half_h), (float)focal);
    }
    double dz = -1.0;
    float* out = rays + t * 6;
    out[0] = (float)cam.t[0]; out[1] = (float)cam.t[1]; out[2] = (float)cam.t[2];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double s = __dadd_rn(__dmul_rn(dx, cam.r[3 * k + 0]), __dmul_rn(dy, cam.r[3 * k + 1]));
        out[3 + k] = (float)__dadd_rn(s, __dmul_rn(dz, cam.r[3 * k + 2]));
    }
}

// B full images from B poses that live in DEVICE memory (poses[B,12] doubles, row-major 3x4): the ray table of
// Generator.forward's per-latent loop (pi_GAN/modules.py:176-184, one random pose per latent) in one launch, with nothing
// step-dependent in the kernel arguments -- so a CUDA graph that contains it can be replayed with new poses.  Same arithmetic
// as the two kernels above (mode = compute_f64 of b2r_raygen: 0 = all float32, bit 0 = float64 division, bit 1 = float64 pose).
__global__ void raygen_poses_kernel(const double* __restrict__ poses, int width, long long rays_per_pose, long long total,
                                    float half_w, float half_h, double focal, int mode, float* __restrict__ rays) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= total) return;
    const long long b = t / rays_per_pose, pix = t - b * rays_per_pose;
    const double* P = poses + b * 12;
    float i = (float)(pix % width), j = (float)(pix / width);
    float* out = rays + t * 6;
    if (mode) {
        double dx, dy;
        if (mode & 1) {
            dx = __ddiv_rn((double)__fsub_rn(i, half_w), focal);
            dy = __ddiv_rn((double)(-__fsub_rn(j, half_h)), focal);
        } else {
            dx = (double)__fdiv_rn(__fsub_rn(i, half_w), (float)focal);
            dy = (double)__fdiv_rn(-__fsub_rn(j, half_h), (float)focal);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            out[k] = (float)P[4 * k + 3];
            double s = __dadd_rn(__dmul_rn(dx, P[4 * k + 0]), __dmul_rn(dy, P[4 * k + 1]));
            out[3 + k] = (float)__dadd_rn(s, __dmul_rn(-1.0, P[4 * k + 2]));
        }
    } else {
        const float dx = __fdiv_rn(__fsub_rn(i, half_w), (float)focal);
        const float dy = __fdiv_rn(-__fsub_rn(j, half_h), (float)focal);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            out[k] = (float)P[4 * k + 3];
            float s = __fadd_rn(__fmul_rn(dx, (float)P[4 * k + 0]), __fmul_rn(dy, (float)P[4 * k + 1]));
            out[3 + k] = __fadd_rn(s, __fmul_rn(-1.0f, (float)P[4 * k + 2]));
        }
    }
}

// z = lower + (upper - lower) * t, three separately rounded ops as in torch (render.py:132).
__global__ void stratified_kernel(const float* __restrict__ z_lin, const float* __restrict__ t_rand,
                                  long long total, int sc, float* __restrict__ z_out,
                                  float* __restrict__ mids_out) {
    extern __shared__ float sm[];
    float* lower = sm;
    float* span = sm + sc;
    for (int k = threadIdx.x; k < sc; k += blockDim.x) {
        float zk = z_lin[k];
        float lo = k == 0 ? zk : __fmul_rn(0.5f, __fadd_rn(zk, z_lin[k - 1]));
        float up = k == sc - 1 ? zk : __fmul_rn(0.5f, __fadd_rn(z_lin[k + 1], zk));
        lower[k] = lo;
        span[k] = __fsub_rn(up, lo);
        if (mids_out && blockIdx.x == 0 && k < sc - 1) mids_out[k] = up;
    }
    __syncthreads();
    long long stride = (long long)gridDim.x * blockDim.x;
    if ((sc & 3) == 0 && ((((uintptr_t)t_rand) | ((uintptr_t)z_out)) & 15) == 0) {
        // 16-byte path: 4 consecutive samples of one ray per thread-iteration
        const float4* t4 = reinterpret_cast<const float4*>(t_rand);
        float4* z4 = reinterpret_cast<float4*>(z_out);
        const long long total4 = total >> 2;
        const int sc4 = sc >> 2;
        for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total4; e += stride) {
            const int k = (int)(e % sc4) * 4;
            float4 t = t4[e], o;
            o.x = __fadd_rn(lower[k + 0], __fmul_rn(span[k + 0], t.x));
            o.y = __fadd_rn(lower[k + 1], __fmul_rn(span[k + 1], t.y));
            o.z = __fadd_rn(lower[k + 2], __fmul_rn(span[k + 2], t.z));
            o.w = __fadd_rn(lower[k + 3], __fmul_rn(span[k + 3], t.w));
            z4[e] = o;
        }
        return;
    }
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += stride) {
        int k = (int)(e % sc);
        z_out[e] = __fadd_rn(lower[k], __fmul_rn(span[k], t_rand[e]));
    }
}

}  // namespace b2r

extern "C" int b2r_raygen(const double* c2w_host, int width, int height, double focal, int compute_f64,
                          long long ray_begin, long long ray_count, float* rays_out, void* stream) {
    using namespace b2r;
    B2R_CHECK_ARG(c2w_host && rays_out, "b2r_raygen: NULL pointer");
    B2R_CHECK_ARG(width > 0 && height > 0 && focal != 0.0, "b2r_raygen: bad image geometry");
    B2R_CHECK_ARG(ray_begin >= 0 && ray_count >= 0 && ray_begin + ray_count <= (long long)width * height,
                  "b2r_raygen: ray range outside the image");
    if (ray_count == 0) return 0;
    Cam<double> cam;
    Cam<float> camf;
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) { cam.r[3 * r + c] = c2w_host[4 * r + c]; camf.r[3 * r + c] = (float)c2w_host[4 * r + c]; }
        cam.t[r] = c2w_host[4 * r + 3]; camf.t[r] = (float)c2w_host[4 * r + 3];
    }
    // width * 0.5 is a python float; numpy (NEP 50) keeps the float32 array dtype
    float half_w = (float)(width * 0.5), half_h = (float)(height * 0.5);
    int block = 256;
    long long grid = (ray_count + block - 1) / block;
    cudaStream_t st = (cudaStream_t)stream;
    if (compute_f64)
        raygen_f64_kernel<<<(unsigned)grid, block, 0, st>>>(cam, width, half_w, half_h, focal, compute_f64 & 1, ray_begin, ray_count, rays_out);
    else
        raygen_f32_kernel<<<(unsigned)((ray_count * 6 + block - 1) / block), block, 0, st>>>(camf, width, half_w, half_h, (float)focal, ray_begin,
                                                                                            ray_count, rays_out);
    B2R_LAUNCH_CHECK("b2r_raygen");
    return 0;
}

extern "C" int b2r_raygen_poses(const double* poses_dev, int n_poses, int width, int height, double focal, int compute_f64,
                                float* rays_out, void* stream) {
    using namespace b2r;
    B2R_CHECK_ARG(n_poses >= 0, "b2r_raygen_poses: negative pose count");
    if (n_poses == 0) return 0;
    B2R_CHECK_ARG(poses_dev && rays_out, "b2r_raygen_poses: NULL pointer");
    B2R_CHECK_ARG(width > 0 && height > 0 && focal != 0.0, "b2r_raygen_poses: bad image geometry");
    const long long per = (long long)width * height, total = per * n_poses;
    float half_w = (float)(width * 0.5), half_h = (float)(height * 0.5);
    raygen_poses_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(poses_dev, width, per, total, half_w, half_h, focal,
                                                                                          compute_f64 & 3, rays_out);
    B2R_LAUNCH_CHECK("b2r_raygen_poses");
    return 0;
}

extern "C" int b2r_stratified_z(const float* z_lin, const float* t_rand, long long n_rays, int n_coarse,
                                float* z_out, float* mids_out, void* stream) {
    using namespace b2r;
    B2R_CHECK_ARG(z_lin && t_rand && z_out, "b2r_stratified_z: NULL pointer");
    B2R_CHECK_ARG(n_rays >= 0 && n_coarse >= 2 && n_coarse <= 4096, "b2r_stratified_z: need n_rays >= 0, 2 <= n_coarse <= 4096");
    long long total = n_rays * n_coarse;
    int block = 256;
    long long want = (total + block - 1) / block;
    long long cap = 148LL * 8;
    unsigned grid = (unsigned)(want < 1 ? 1 : (want > cap ? cap : want));
    stratified_kernel<<<grid, block, 2 * n_coarse * sizeof(float), (cudaStream_t)stream>>>(z_lin, t_rand, total, n_coarse, z_out, mids_out);
    B2R_LAUNCH_CHECK("b2r_stratified_z");
    return 0;
}

// ---- image-space output: to8b (nerf/render.py:5) on the device -----------------------------------------------------
//   to8b = lambda x: (255 * np.clip(x, 0, 1)).astype(np.uint8)      float32 product, truncation toward zero
namespace b2r {
__global__ void __launch_bounds__(256) to8b_kernel(const float4* __restrict__ x, long long n4, long long n, uchar4* __restrict__ out) {
    auto q = [](float v) -> unsigned char { return (unsigned char)__float2uint_rz(__fmul_rn(255.0f, fminf(fmaxf(v, 0.0f), 1.0f))); };
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = x[i];
        out[i] = make_uchar4(q(v.x), q(v.y), q(v.z), q(v.w));
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n - 4 * n4)) {
        const long long i = 4 * n4 + threadIdx.x;
        reinterpret_cast<unsigned char*>(out)[i] = q(reinterpret_cast<const float*>(x)[i]);
    }
}
}  // namespace b2r

extern "C" int b2r_to8b(const float* x, long long n, unsigned char* out, void* stream) {
    using namespace b2r;
    B2R_CHECK_ARG(n >= 0, "b2r_to8b: negative size");
    if (n == 0) return 0;
    B2R_CHECK_ARG(x && out, "b2r_to8b: NULL pointer");
    B2R_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)out & 3) == 0, "b2r_to8b: x must be 16-byte and out 4-byte aligned");
    const long long n4 = n / 4;
    long long blocks = (n4 + 255) / 256;
    blocks = blocks < 1 ? 1 : (blocks > 148 * 8 ? 148 * 8 : blocks);
    to8b_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const float4*)x, n4, n, (uchar4*)out);
    B2R_LAUNCH_CHECK("b2r_to8b");
    return 0;
}
