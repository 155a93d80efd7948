// Fused Adam on the flat fp32 parameter bucket: the optimiser step of nerf/train_nerf.py:168 with the
// learning-rate schedule of :170-175, as two launches that can be captured in a CUDA graph -- every quantity that changes
// from step to step (step count, decayed learning rate, bias corrections) lives in a 4-word device state, not in kernel
// arguments.
//   torch.optim.Adam (no weight decay, no amsgrad):  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
//   p -= lr_t / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
//   train_nerf.py schedule: step t (1-based) runs with lr0 * decay_rate^((t-1) / decay_steps).
#include "common.cuh"

namespace b2r {

// state: [0] step count (int32 bits), [1] lr_t, [2] 1 - b1^t, [3] sqrt(1 - b2^t)
__global__ void adam_tick_kernel(float* __restrict__ state, float lr0, float lr_end, float decay_rate, float decay_steps, float beta1, float beta2) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int t = __float_as_int(state[0]) + 1;
    state[0] = __int_as_float(t);
    // train_nerf.py:170-175 (lr_end = 0) and pi_GAN/train.py:140-145 (floor lr_end): lr_end + (lr0 - lr_end) * rate^((t-1) / decay_steps)
    const double lr = decay_steps > 0.f ? (double)lr_end + ((double)lr0 - (double)lr_end) * pow((double)decay_rate, (double)(t - 1) / (double)decay_steps)
                                        : (double)lr0;
    state[1] = (float)lr;
    state[2] = (float)(1.0 - pow((double)beta1, (double)t));
    state[3] = (float)sqrt(1.0 - pow((double)beta2, (double)t));
}

__global__ void __launch_bounds__(256) adam_step_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m,
                                                        float4* __restrict__ v, long long n4, long long n, const float* __restrict__ state,
                                                        float beta1, float beta2, float eps, float grad_scale) {
    const float lr = state[1], bc1 = state[2], sbc2 = state[3];
    const float step_size = lr / bc1;
    auto upd = [&](float& pp, float gg, float& mm, float& vv) {
        gg *= grad_scale;
        mm = beta1 * mm + (1.0f - beta1) * gg;
        vv = beta2 * vv + (1.0f - beta2) * gg * gg;
        pp -= step_size * (mm / (sqrtf(vv) / sbc2 + eps));
    };
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 pp = p[i], gg = g[i], mm = m[i], vv = v[i];
        upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y); upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
        p[i] = pp; m[i] = mm; v[i] = vv;
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n - n4 * 4)) {      // tail (n not a multiple of 4)
        const long long i = n4 * 4 + threadIdx.x;
        float* ps = reinterpret_cast<float*>(p); const float* gs = reinterpret_cast<const float*>(g);
        float* ms = reinterpret_cast<float*>(m); float* vs = reinterpret_cast<float*>(v);
        upd(ps[i], gs[i], ms[i], vs[i]);
    }
}

}  // namespace b2r

extern "C" int b2r_adam_step_floor(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float* state,
                                   float lr0, float lr_end, float decay_rate, float decay_steps, float beta1, float beta2, float eps,
                                   float grad_scale, void* stream);

extern "C" int b2r_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float* state,
                             float lr0, float decay_rate, float decay_steps, float beta1, float beta2, float eps, float grad_scale,
                             void* stream) {
    return b2r_adam_step_floor(params, grads, exp_avg, exp_avg_sq, n, state, lr0, 0.0f, decay_rate, decay_steps, beta1, beta2, eps, grad_scale, stream);
}

extern "C" int b2r_adam_step_floor(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float* state,
                                   float lr0, float lr_end, float decay_rate, float decay_steps, float beta1, float beta2, float eps,
                                   float grad_scale, void* stream) {
    using namespace b2r;
    B2R_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && state, "b2r_adam_step: NULL pointer");
    B2R_CHECK_ARG(n >= 0, "b2r_adam_step: negative size");
    B2R_CHECK_ARG((((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq | (uintptr_t)state) & 15) == 0,
                  "b2r_adam_step: buffers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    adam_tick_kernel<<<1, 32, 0, st>>>(state, lr0, lr_end, decay_rate, decay_steps, beta1, beta2);
    B2R_LAUNCH_CHECK("b2r_adam_step (tick)");
    if (n == 0) return 0;
    const long long n4 = n / 4;
    long long blocks = (n4 + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 8) blocks = 148 * 8;
    adam_step_kernel<<<(unsigned)blocks, 256, 0, st>>>((float4*)params, (const float4*)grads, (float4*)exp_avg, (float4*)exp_avg_sq, n4, n, state,
                                                      beta1, beta2, eps, grad_scale);
    B2R_LAUNCH_CHECK("b2r_adam_step");
    return 0;
}
