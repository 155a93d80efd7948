// K8 (tensor-core path): reverse mode of the fused NeRF MLP -- the backward pass nerf/train_nerf.py:167 gets from
// autograd (loss.backward() through NeRF.forward, nerf/nerf.py:75-94), restructured for the B200:
//
//   forward  (mlp_tc.cu, nerf_tc_kernel<true>)  the inference kernel + every layer input spilled as tiled bf16 tensors;
//   dgrad    (nerf_tc_bwd_kernel, here)         ONE persistent CTA-pair kernel walking the layers backwards: the incoming
//            gradient of a 128-row sub-tile stays in shared memory as the A operand of the next tcgen05 step
//            (dX = dPre W, weights pre-packed TRANSPOSED by b2r_mlp_tc_pack_bwd and streamed by the bulk-copy engine);
//            the epilogue applies relu'(h) from the saved activation (packed bf16x2 compare + mask), adds the sigma-head
//            term, rounds to bf16, writes the next A operand in place and spills d(pre-activation) for wgrad;
//   wgrad    (nerf_tc_wgrad_kernel, here)       dW_l = dPre_l^T X_l over all rows: both operands are the spilled tiles read
//            back as MN-major SWIZZLE_128B operands (row index = K, no transposition), 64-row stages through a 3-stage
//            bulk-copy ring, fp32 accumulators in TMEM for a whole range of tiles, flushed once with float atomics;
//            four otherwise idle warps add up the bias gradients (column sums of dPre) from the staged tiles;
//   heads    (nerf_head_wgrad_kernel)           output_layer_sigma / output_layer_rgb weight + bias gradients (1 and 3 rows).
//
// Arithmetic: bf16 operands, fp32 accumulation, gradients accumulated into the caller's fp32 flat bucket (the layout the
// NCCL all-reduce and Adam use).  HBM traffic per row: 5,120 B written by the forward, 4,352 B read + 4,880 B written by
// dgrad, 10,624 B read by wgrad -- the training step is HBM-bound, not tensor-bound.
#include "tc_core.cuh"

namespace b2r {
namespace tc {

// ---- dgrad schedule -------------------------------------------------------------------------------------------------
// step 0: d g  = dPre(layers_dir.1)[128] . W_d1[:, 0:256]        (K = 128: two chunks)
// step 1: d h7 = d g . W_d0            (+ sigma-head term, relu'(h7))
// step 2, 3: layers_pos.7, layers_pos.6;  step 4: layers_pos.5[:, 60:316] (the h4 part of the skip input);
// step 5..8: layers_pos.4 .. layers_pos.1.   layers_pos.0 has no dgrad (its input is the encoding).
struct BwdSched {
    static constexpr int kSteps = 9;
    __host__ __device__ static constexpr int n_pre(int, int) { return 0; }
    __host__ __device__ static constexpr int n_h(int s, int) { return s == 0 ? 2 : 4; }
    __host__ __device__ static constexpr int n_post(int, int) { return 0; }
    __host__ __device__ static constexpr int n(int) { return 256; }
    static constexpr int kPostMmas = 1;
};
__host__ __device__ constexpr int bwd_layer(int s) { return s == 0 ? 9 : (s == 1 ? 8 : 9 - s); }
constexpr long long kBwdChunkBytes = step_base<BwdSched>(BwdSched::kSteps);           // 34 chunks x 2 halves x 16 KB
static_assert(kBwdChunkBytes == 34LL * 32768, "bwd packed chunk bytes");
constexpr int kBwdTabWSigma = 0, kBwdTabWRgb = 256, kBwdTabFloats = 640;              // w_sigma[256] | w_rgb[3][128]
constexpr long long kBwdPackedBytes = kBwdChunkBytes + kBwdTabFloats * 4;

// B operand of dgrad step s: B[n][k] = scale_L * W_L[k][n + col_off]  (n = input feature = output column of dX, k = output feature).
// kSiren: the SirenNeRF layer table, skip offset 3 (raw position first, nerf/nerf.py:158) and the factor 30 of the sine layers
// folded into the rows exactly as in the forward's packed weights (t = 30 (W x + b)  =>  dt/dx = 30 W).
template <bool kSiren>
__global__ void nerf_pack_bwd_kernel(const float* __restrict__ params, uint8_t* __restrict__ packed) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t < kBwdChunkBytes / 16) {
        int s, c, hf, row, grp;
        locate<BwdSched>(t * 16, s, c, hf, row, grp);
        LayerDesc L = kSiren ? siren_layer(bwd_layer(s)) : nerf_layer(bwd_layer(s));
        const int n = hf * 128 + row + (s == 4 ? (kSiren ? 3 : 60) : 0);
        const float scale = (kSiren && s != 1) ? 30.0f : 1.0f;                // step 1 = layers_dir.0 (linear)
        __nv_bfloat16 v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            int k = c * 64 + grp * 8 + e;
            v[e] = __float2bfloat16_rn(k < L.out ? scale * params[L.w_off + (long long)k * L.in + n] : 0.f);
        }
        uint8_t* dst = packed + step_base<BwdSched>(s) + (long long)(c * 2 + hf) * half_bytes<BwdSched>(s) +
                       sw128_offset((uint32_t)row, (uint32_t)grp);
        *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(v);
    }
    if (t < kBwdTabFloats) {
        float* tab = reinterpret_cast<float*>(packed + kBwdChunkBytes);
        int i = (int)t;
        const LayerDesc Ls = kSiren ? siren_layer(10) : nerf_layer(10), Lc = kSiren ? siren_layer(11) : nerf_layer(11);
        tab[i] = i < kBwdTabWRgb ? params[Ls.w_off + i] : params[Lc.w_off + (i - kBwdTabWRgb)];
    }
}

// One dgrad step's epilogue for this warp's half (128) of the columns.
//   MODE 0: linear (d g: layers_dir.0 has no activation);  MODE 1: + gs * w_sigma (sigma-head term), act'(h7);  MODE 2: act'(h)
// act' = relu' from the layer's relu bits (NeRF: one prefetched 16-byte word per thread) or, kSiren, cos(t) from the forward's
// thread-major byte checkpoint (two 16-byte words per 32-column group, the next group's in flight; tc_core.cuh: cos_unq).
// The result (bf16) is written in place as the next step's A operand; the spill thread copies the same tile to global memory
// for wgrad.  dptr: NeRF: this thread's relu-bit word of the layer; SirenNeRF: its first cosine word (quarter 2 * half, w = 0).
template <int MODE, bool kSiren>
__device__ __forceinline__ void bwd_epi(uint32_t t_half, uint32_t h_half, const uint32_t (&xoff)[8], const uint8_t* __restrict__ dptr,
                                        float gs, uint32_t wsig_half, uint32_t acc_bar, uint32_t& acc_phase, uint32_t done_bar, uint32_t& sp_phase) {
    uint4 mk4 = make_uint4(0u, 0u, 0u, 0u);
    uint4 cw[2];
    // cosine bytes of 32-column group jj: quarter (jj >> 1) of this half, words (jj & 1) * 2, + 1; 2048 B between words, 4 words per quarter
    auto cos_ptr = [&](int jj, int q) -> const uint8_t* { return dptr + (size_t)((jj >> 1) * 4 + (jj & 1) * 2 + q) * 2048; };
    if (MODE != 0) {
        if (kSiren) {
#pragma unroll
            for (int q = 0; q < 2; ++q) cw[q] = ldg128(cos_ptr(0, q));
        } else mk4 = ldg128(dptr);
    }
    mbar_wait_cluster(acc_bar, acc_phase);
    acc_phase ^= 1u;
    tc_fence_after();
    mbar_wait(done_bar, sp_phase);                 // the previous tile copy has read the blocks this epilogue overwrites
    sp_phase ^= 1u;
    const uint32_t mk[4] = {mk4.x, mk4.y, mk4.z, mk4.w};
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
        uint32_t v[32];
        tmem_ld32(t_half + (uint32_t)jj * 32u, v);
        uint4 cn[2];
        if (kSiren && MODE != 0 && jj < 3) {
#pragma unroll
            for (int q = 0; q < 2; ++q) cn[q] = ldg128(cos_ptr(jj + 1, q));
        }
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) f[e] = __uint_as_float(v[e]);
        if (MODE == 1) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                float4 w = lds128(wsig_half + (uint32_t)(jj * 32 + q * 4) * 4u);
                f[4 * q + 0] = fmaf(gs, w.x, f[4 * q + 0]); f[4 * q + 1] = fmaf(gs, w.y, f[4 * q + 1]);
                f[4 * q + 2] = fmaf(gs, w.z, f[4 * q + 2]); f[4 * q + 3] = fmaf(gs, w.w, f[4 * q + 3]);
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (kSiren && MODE != 0) {         // fp32 product with the checkpointed cosine, then one rounding to bf16
                const uint4 c = cw[q >> 1];
                cos_mul8(f + 8 * q, (q & 1) ? c.z : c.x, (q & 1) ? c.w : c.y);
            }
            uint32_t w0 = pack_bf16(f[8 * q + 0], f[8 * q + 1]), w1 = pack_bf16(f[8 * q + 2], f[8 * q + 3]);
            uint32_t w2 = pack_bf16(f[8 * q + 4], f[8 * q + 5]), w3 = pack_bf16(f[8 * q + 6], f[8 * q + 7]);
            if (MODE != 0 && !kSiren) {
                w0 &= mask_get(mk[jj], 4 * q + 0); w1 &= mask_get(mk[jj], 4 * q + 1);
                w2 &= mask_get(mk[jj], 4 * q + 2); w3 &= mask_get(mk[jj], 4 * q + 3);
            }
            st_shared_v4(h_half + (uint32_t)(jj >> 1) * kBlk + xoff[(jj & 1) * 4 + q], w0, w1, w2, w3);
        }
        if (kSiren && MODE != 0 && jj < 3) {
#pragma unroll
            for (int q = 0; q < 2; ++q) cw[q] = cn[q];
        }
    }
}

// =====================================================================================================================
// dgrad: d_raw -> d(pre-activation) of every layer (tiled bf16 tensors in `scratch`) + head gradients HG
// =====================================================================================================================
template <bool kSiren>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
nerf_tc_bwd_kernel(const uint8_t* __restrict__ packed, long long rows, const float4* __restrict__ raw, const float4* __restrict__ d_raw,
                   const uint8_t* __restrict__ saved, uint8_t* __restrict__ scratch) {
    extern __shared__ uint8_t smem_raw[];
    const Ctx cx = make_ctx(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const PairLoop pl(rows);
    {   // w_sigma | w_rgb -> shared memory
        const float4* tab_g = reinterpret_cast<const float4*>(packed + kBwdChunkBytes);
        for (int i = threadIdx.x; i < kBwdTabFloats / 4; i += kThreads) {
            float4 v = __ldg(tab_g + i);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(cx.smem + kTabOff + 16u * i), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
        }
    }
    const uint32_t tmem_base = tc_prologue(cx, warp);

    if (warp == 0) {
        if (lane == 0) producer_loop<BwdSched>(cx, packed, pl, BwdSched::kSteps, 0);
    } else if (warp == 1) {
        if (cx.rank == 0) mma_loop<BwdSched>(cx, tmem_base, pl, BwdSched::kSteps, 0);
        else if (lane == 0) relay_loop<BwdSched>(cx, pl, BwdSched::kSteps, 0);
    } else if (warp < kCtrlWarps) {
        if (lane == 0) {
            // ===== spill thread of sub-tile g: gradient tiles (the A operands) -> tiled tensors in `scratch` =====
            const int g = warp - 2;
            const uint32_t hreg = cx.smem + (uint32_t)g * kSubBytes + kPeBytes;
            const uint32_t ready = cx.spill_ready + 8 * g, done = cx.spill_done + 8 * g;
            const size_t n_sub = (size_t)pl.n_pairs * 4;
            uint32_t ph = 0;
            for (long long p = pl.first; p < pl.n_pairs; p += pl.stride) {
                const size_t T = (size_t)((2 * p + cx.rank) * 2 + g);
                auto tile = [&](int off, int nb) -> uint8_t* { return scratch + ((size_t)off * n_sub + T * (size_t)nb) * kBlk; };
                auto finish = [&]() { bulk_commit(); bulk_wait_read(); mbar_arrive(done); ph ^= 1u; };
                mbar_wait(ready, ph); spill_tile(tile(kScrGD1, 2), hreg, 2); finish();                 // d pre(layers_dir.1)
                mbar_wait(ready, ph); spill_tile(tile(kScrGG, 4), hreg, 4); finish();                  // d g
                for (int l = 7; l >= 0; --l) { mbar_wait(ready, ph); spill_tile(tile(scr_gh(l), 4), hreg, 4); finish(); }
            }
            bulk_wait_all();
        }
    } else {
        const int ew = warp - kCtrlWarps;
        const int g = ew >> 3, half = (ew >> 2) & 1, quad = ew & 3;
        const int r = (quad << 5) | lane;
        const uint32_t sub = cx.smem + (uint32_t)g * kSubBytes;
        const uint32_t h_base = sub + kPeBytes;
        const uint32_t t_addr = tmem_base + ((uint32_t)quad << 21) + (uint32_t)g * 256u;
        const uint32_t row_off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
        const uint32_t xr = (uint32_t)(r & 7);
        const uint32_t tab = cx.smem + kTabOff;
        const uint32_t act_local = cx.act_ready + 8 * g, act_leader = mapa(act_local, 0);
        const uint32_t acc_bar = cx.acc_full + 8 * g;
        const uint32_t ready_bar = cx.spill_ready + 8 * g, done_bar = cx.spill_done + 8 * g;
        const uint32_t t_half = t_addr + (uint32_t)half * 128u;
        const uint32_t h_half = h_base + row_off + (uint32_t)half * 2u * kBlk;
        const uint32_t wsig_half = tab + (uint32_t)(kBwdTabWSigma + half * 128) * 4u;
        uint32_t xoff[8];
#pragma unroll
        for (uint32_t c = 0; c < 8; ++c) xoff[c] = (c ^ xr) << 4;
        uint32_t acc_phase = 0, sp_phase = 0;
        bool first_tile = true;
        const size_t n_sub = (size_t)pl.n_pairs * 4;
        float4* __restrict__ hg_out = reinterpret_cast<float4*>(scratch + (size_t)kScrBlocks * n_sub * kBlk);
        for (long long p = pl.first; p < pl.n_pairs; p += pl.stride) {
            const long long row = (2 * p + cx.rank) * kRowsTile + g * kRowsSub + r;
            const bool valid = row < rows;
            const size_t T = (size_t)((2 * p + cx.rank) * 2 + g);
            auto spill_sig = [&]() { if (lane == 0) mbar_arrive(ready_bar); };
            // ---- heads: d raw -> (d rgb pre-sigmoid, d sigma pre-relu); rows past the end contribute zero everywhere
            float gc0 = 0.f, gc1 = 0.f, gc2 = 0.f, gs = 0.f;
            if (valid) {
                const float4 y = __ldg(raw + row), dy = __ldg(d_raw + row);
                gc0 = dy.x * (y.x * (1.0f - y.x));
                gc1 = dy.y * (y.y * (1.0f - y.y));
                gc2 = dy.z * (y.z * (1.0f - y.z));
                gs = y.w > 0.f ? dy.w : 0.f;
            }
            if (half == 0) hg_out[T * kRowsSub + r] = make_float4(gc0, gc1, gc2, gs);
            {
                // d h_d = (d rgb pre) . W_rgb, act'(h_d): this half produces columns half*64 .. +63 = K-block `half` of step 0
                uint2 hm = make_uint2(0u, 0u);
                if (!kSiren) hm = *reinterpret_cast<const uint2*>(saved + hdmask_off(n_sub, T, half, r));
                const uint32_t wr = tab + (uint32_t)(kBwdTabWRgb + half * 64) * 4u;
                if (!first_tile) { mbar_wait(done_bar, sp_phase); sp_phase ^= 1u; }          // previous tile's d h0 copy
                first_tile = false;
                uint4 cwq = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    // SirenNeRF: cos(t) of layers_dir.1, quarter half * 2 + (q >> 2) (32 columns each), word (q & 3) >> 1 (16 columns)
                    if (kSiren && (q & 1) == 0) cwq = ldg128(saved + siren_cos9_off(n_sub, T, half * 2 + (q >> 2), (q & 3) >> 1, r));
                    float f[8];
#pragma unroll
                    for (int hq = 0; hq < 2; ++hq) {
                        const uint32_t wa = wr + (uint32_t)(q * 8 + hq * 4) * 4u;
                        const float4 w0 = lds128(wa), w1 = lds128(wa + 512u), w2 = lds128(wa + 1024u);
                        f[4 * hq + 0] = fmaf(gc2, w2.x, fmaf(gc1, w1.x, gc0 * w0.x));
                        f[4 * hq + 1] = fmaf(gc2, w2.y, fmaf(gc1, w1.y, gc0 * w0.y));
                        f[4 * hq + 2] = fmaf(gc2, w2.z, fmaf(gc1, w1.z, gc0 * w0.z));
                        f[4 * hq + 3] = fmaf(gc2, w2.w, fmaf(gc1, w1.w, gc0 * w0.w));
                    }
                    if (kSiren) cos_mul8(f, (q & 1) ? cwq.z : cwq.x, (q & 1) ? cwq.w : cwq.y);
                    uint32_t w0 = pack_bf16(f[0], f[1]), w1 = pack_bf16(f[2], f[3]), w2 = pack_bf16(f[4], f[5]), w3 = pack_bf16(f[6], f[7]);
                    if (!kSiren) {
                        const uint32_t bits = (q >> 2) ? hm.y : hm.x;
                        const int i0 = 4 * (q & 3);
                        w0 &= mask_get(bits, i0 + 0); w1 &= mask_get(bits, i0 + 1); w2 &= mask_get(bits, i0 + 2); w3 &= mask_get(bits, i0 + 3);
                    }
                    st_shared_v4(h_base + (uint32_t)half * kBlk + row_off + xoff[q], w0, w1, w2, w3);
                }
            }
            arrive_act(act_local, act_leader, cx.rank, lane);
            spill_sig();
            // activation-derivative checkpoint of layer l for this thread: relu bits (NeRF) / first cosine word of quarter 2 * half (SirenNeRF)
            auto mptr = [&](int l) -> const uint8_t* {
                return kSiren ? saved + siren_cos_off(n_sub, l, T, half * 2, 0, r) : saved + mask_off(n_sub, l, T, half, r);
            };
            // step 0: d g (linear)
            bwd_epi<0, kSiren>(t_half, h_half, xoff, nullptr, 0.f, 0u, acc_bar, acc_phase, done_bar, sp_phase);
            arrive_act(act_local, act_leader, cx.rank, lane);
            spill_sig();
            // step 1: d h7 (+ sigma head), act'(h7)
            bwd_epi<1, kSiren>(t_half, h_half, xoff, mptr(7), gs, wsig_half, acc_bar, acc_phase, done_bar, sp_phase);
            arrive_act(act_local, act_leader, cx.rank, lane);
            spill_sig();
            // steps 2..7: d h6 .. d h1
            for (int s = 2; s < 8; ++s) {
                bwd_epi<2, kSiren>(t_half, h_half, xoff, mptr(8 - s), 0.f, 0u, acc_bar, acc_phase, done_bar, sp_phase);
                arrive_act(act_local, act_leader, cx.rank, lane);
                spill_sig();
            }
            // step 8: d h0 -> dPre0: only wgrad reads it (layers_pos.0 has no dgrad), no MMA follows
            bwd_epi<2, kSiren>(t_half, h_half, xoff, mptr(0), 0.f, 0u, acc_bar, acc_phase, done_bar, sp_phase);
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            spill_sig();
        }
    }
    tc_teardown(tmem_base, warp);
}

// =====================================================================================================================
// wgrad: dW[out, in] += sum_rows dPre[row, out] * X[row, in];  db[out] += sum_rows dPre[row, out]
// =====================================================================================================================
struct WUnit {
    int g_off, g_nb;        // gradient tensor in scratch: block offset, blocks per tile (4 = two 128-row output halves, 2 = one)
    int x_off, x_nb;        // layer-input tensor in saved: block offset, blocks per tile
    int layer;              // nerf_layer index (w_off, in, b_off)
    int col_off, n_valid;   // columns [col_off, col_off + n_valid) of dW come from this X tensor ...
    int bias;               // this unit also reduces the bias gradient
    int x_col0;             // ... from its columns [x_col0, x_col0 + n_valid)
    float scale;            // dW = scale * G^T X (30 for SirenNeRF's sine layers, whose G is the gradient wrt t = 30 (W x + b))
};
constexpr int kWUnits = 12;
__constant__ WUnit c_wunits[kWUnits] = {
    {scr_gh(0), 4, kSavPE, 1, 0, 0, 60, 1, 0, 1.0f},
    {scr_gh(1), 4, sav_h(0), 4, 1, 0, 256, 1, 0, 1.0f},
    {scr_gh(2), 4, sav_h(1), 4, 2, 0, 256, 1, 0, 1.0f},
    {scr_gh(3), 4, sav_h(2), 4, 3, 0, 256, 1, 0, 1.0f},
    {scr_gh(4), 4, sav_h(3), 4, 4, 0, 256, 1, 0, 1.0f},
    {scr_gh(5), 4, kSavPE, 1, 5, 0, 60, 0, 0, 1.0f},
    {scr_gh(5), 4, sav_h(4), 4, 5, 60, 256, 1, 0, 1.0f},
    {scr_gh(6), 4, sav_h(5), 4, 6, 0, 256, 1, 0, 1.0f},
    {scr_gh(7), 4, sav_h(6), 4, 7, 0, 256, 1, 0, 1.0f},
    {kScrGG, 4, sav_h(7), 4, 8, 0, 256, 1, 0, 1.0f},
    {kScrGD1, 2, kSavGL, 4, 9, 0, 256, 1, 0, 1.0f},
    {kScrGD1, 2, kSavDE, 1, 9, 256, 24, 0, 0, 1.0f},
};
// SirenNeRF: the aux tile [pos(3), 1, 1, dir(3)] feeds layers_pos.0, the skip layer and layers_dir.1
__constant__ WUnit c_wunits_siren[kWUnits] = {
    {scr_gh(0), 4, kSsAUX, 1, 0, 0, 3, 1, 0, 30.0f},
    {scr_gh(1), 4, ss_h(0), 4, 1, 0, 256, 1, 0, 30.0f},
    {scr_gh(2), 4, ss_h(1), 4, 2, 0, 256, 1, 0, 30.0f},
    {scr_gh(3), 4, ss_h(2), 4, 3, 0, 256, 1, 0, 30.0f},
    {scr_gh(4), 4, ss_h(3), 4, 4, 0, 256, 1, 0, 30.0f},
    {scr_gh(5), 4, kSsAUX, 1, 5, 0, 3, 0, 0, 30.0f},
    {scr_gh(5), 4, ss_h(4), 4, 5, 3, 256, 1, 0, 30.0f},
    {scr_gh(6), 4, ss_h(5), 4, 6, 0, 256, 1, 0, 30.0f},
    {scr_gh(7), 4, ss_h(6), 4, 7, 0, 256, 1, 0, 30.0f},
    {kScrGG, 4, ss_h(7), 4, 8, 0, 256, 1, 0, 1.0f},
    {kScrGD1, 2, kSsGL, 4, 9, 0, 256, 1, 0, 30.0f},
    {kScrGD1, 2, kSsAUX, 1, 9, 256, 3, 0, 5, 30.0f},
};
// FiLM-SIREN (pi_GAN/modules.py:101-118): G_l = gradient wrt the sine argument t_l of layer l (0 = input_layer, 1..7 =
// hidden_layers.0..6, 8 = hidden_layer_rgb); the aux tile [dir(3), 1, 1, pos(3)] is the input of input_layer and the direction
// columns of hidden_layer_rgb.  The units produce the gradients of the FOLDED parameters (W' = 30 gamma W, s' = 30 (gamma b + beta));
// film_grad_finish_kernel turns them into d W, d b, d gamma, d beta.
__constant__ WUnit c_wunits_film[10] = {
    {fscr_g(0), 4, kFsAUX, 1, 0, 0, 3, 1, 5, 1.0f},
    {fscr_g(1), 4, fs_h(0), 4, 1, 0, 256, 1, 0, 1.0f},
    {fscr_g(2), 4, fs_h(1), 4, 2, 0, 256, 1, 0, 1.0f},
    {fscr_g(3), 4, fs_h(2), 4, 3, 0, 256, 1, 0, 1.0f},
    {fscr_g(4), 4, fs_h(3), 4, 4, 0, 256, 1, 0, 1.0f},
    {fscr_g(5), 4, fs_h(4), 4, 5, 0, 256, 1, 0, 1.0f},
    {fscr_g(6), 4, fs_h(5), 4, 6, 0, 256, 1, 0, 1.0f},
    {fscr_g(7), 4, fs_h(6), 4, 7, 0, 256, 1, 0, 1.0f},
    {fscr_g(8), 4, fs_h(7), 4, 9, 0, 256, 1, 0, 1.0f},
    {fscr_g(8), 4, kFsAUX, 1, 9, 256, 3, 0, 0, 1.0f},
};
// model kinds of the wgrad kernel (= B2R_MODEL_*: 0 NeRF, 1 FiLM-SIREN, 2 SirenNeRF)
// kKFilmNoDir: FilmSirenNeRF(use_dir=False) -- hidden_layer_rgb has 256 inputs: the parameter offsets move and the last unit (its
// direction columns) does not exist
constexpr int kKNerf = B2R_MODEL_NERF, kKFilm = B2R_MODEL_FILM, kKSiren = B2R_MODEL_SIREN, kKFilmNoDir = 3;
template <int KIND> __host__ __device__ constexpr int n_wunits() { return KIND == kKFilm ? 10 : (KIND == kKFilmNoDir ? 9 : kWUnits); }
// half-block loads per 64-row stage, all units: 83 (NeRF, SirenNeRF), 74 / 69 (FiLM-SIREN with / without view direction)
template <int KIND> __host__ __device__ constexpr int wcost_total() {
    return KIND == kKFilm ? 5 + 7 * 8 + 8 + 5 : (KIND == kKFilmNoDir ? 5 + 7 * 8 + 8 : 5 + 4 * 8 + 5 + 8 + 2 * 8 + 8 + 6 + 3);
}
template <int KIND> __host__ __device__ constexpr LayerDesc wlayer(int i) {
    return KIND == kKFilm ? film_layer(i, true) : (KIND == kKFilmNoDir ? film_layer(i, false) : (KIND == kKSiren ? siren_layer(i) : nerf_layer(i)));
}

// Operand ring of the wgrad kernel: 192 KB cut into stages of (g_nb + x_nb) half blocks of 8 KB -- 3 stages of 64 KB for the 256 x 256
// units, up to 8 stages for the narrow ones (pos-enc / dir-enc / aux inputs: 24-40 KB per stage).  With a fixed 3-stage ring the narrow
// units had only 72-120 KB in flight per SM and were latency-bound: their CTAs finished 25-45 % after the others (per-CTA timers:
// 537 us against a median of 417 us on the 262,144-row coarse pass) and set the kernel's duration.
constexpr int kWMaxStages = 8;
constexpr uint32_t kWRing = 196608;
constexpr uint32_t kWBarOff = kWRing;
constexpr uint32_t kWSmem = kWBarOff + 256 + 1024;
constexpr int kWThreads = 192;                                           // producer, MMA issuer, 4 reduce / flush warps
__host__ __device__ constexpr int wstages(int cost) { return (int)(kWRing / ((uint32_t)cost * 8192u)) < kWMaxStages ? (int)(kWRing / ((uint32_t)cost * 8192u)) : kWMaxStages; }

struct WPiece { int u; long long t0, t1; };
// CTA b owns the slice [b, b+1) * total / grid of the cost line (units laid end to end, each n_sub tiles x cost(u)); both
// ends are rounded to tiles with the same function, so neighbouring CTAs agree on the boundary.
template <int KIND>
__device__ __forceinline__ WUnit wunit(int u) {
    return (KIND == kKFilm || KIND == kKFilmNoDir) ? c_wunits_film[u] : (KIND == kKSiren ? c_wunits_siren[u] : c_wunits[u]);
}

template <int KIND>
__device__ __forceinline__ bool wpiece(int u, long long n_sub, long long lo, long long hi, long long& base, WPiece& pc) {
    const WUnit un = wunit<KIND>(u);
    const long long cost = un.g_nb + un.x_nb;
    const long long b0 = base, b1 = base + n_sub * cost;
    base = b1;
    if (hi <= b0 || lo >= b1) return false;
    const long long a = lo > b0 ? lo : b0, b = hi < b1 ? hi : b1;
    pc.u = u;
    pc.t0 = (a - b0 + cost - 1) / cost;
    pc.t1 = (b - b0 + cost - 1) / cost;
    return pc.t0 < pc.t1;
}

template <int KIND>
__global__ void __launch_bounds__(kWThreads, 1)
nerf_tc_wgrad_kernel(const uint8_t* __restrict__ saved, const uint8_t* __restrict__ scratch, long long n_sub, float* __restrict__ d_params,
                     long long t_begin, long long t_count) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t full = smem + kWBarOff, empty = full + 8 * kWMaxStages, acc_full = empty + 8 * kWMaxStages, tmem_empty = acc_full + 8,
                   slot_addr = tmem_empty + 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kWMaxStages; ++s) { mbar_init(full + 8 * s, 1); mbar_init(empty + 8 * s, 5); }
        mbar_init(acc_full, 1);
        mbar_init(tmem_empty, 4);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(slot_addr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot_addr));

    // the tensors are laid out for n_sub sub-tiles; this launch reduces over sub-tiles [t_begin, t_begin + t_count) (one latent of a batch)
    const long long total = t_count * wcost_total<KIND>();
    const long long lo = total * blockIdx.x / gridDim.x, hi = total * (blockIdx.x + 1) / gridDim.x;

    if (warp == 0) {
        if (lane == 0) {
            // every piece starts at stage 0 of ITS stage size; `par` holds one phase bit per barrier (the number of stages differs between
            // pieces, so the barriers are not used equally often)
            uint32_t par = 0;
            long long base = 0;
            for (int u = 0; u < n_wunits<KIND>(); ++u) {
                WPiece pc;
                if (!wpiece<KIND>(u, t_count, lo, hi, base, pc)) continue;
                const WUnit un = wunit<KIND>(u);
                const uint8_t* gsrc = scratch + (size_t)un.g_off * n_sub * kBlk;
                const uint8_t* xsrc = saved + (size_t)un.x_off * n_sub * kBlk;
                const uint32_t bytes = (uint32_t)(un.g_nb + un.x_nb) * 8192u;
                const int n_st = wstages(un.g_nb + un.x_nb);
                int stage = 0;
                for (long long T = t_begin + pc.t0; T < t_begin + pc.t1; ++T) {
                    for (int hs = 0; hs < 2; ++hs) {
                        mbar_wait(empty + 8 * stage, ((par >> stage) & 1u) ^ 1u);
                        par ^= 1u << stage;
                        mbar_arrive_expect_tx(full + 8 * stage, bytes);
                        const uint32_t dst = smem + (uint32_t)stage * bytes;
                        for (int i = 0; i < un.g_nb; ++i)
                            bulk_g2s(dst + (uint32_t)i * 8192u, gsrc + ((size_t)T * un.g_nb + i) * kBlk + (size_t)hs * 8192, 8192u, full + 8 * stage);
                        for (int j = 0; j < un.x_nb; ++j)
                            bulk_g2s(dst + (uint32_t)(un.g_nb + j) * 8192u, xsrc + ((size_t)T * un.x_nb + j) * kBlk + (size_t)hs * 8192, 8192u,
                                     full + 8 * stage);
                        if (++stage == n_st) stage = 0;
                    }
                }
                // the next piece cuts the ring differently: wait until the consumers have released every stage of this one
                for (int st = 0; st < n_st; ++st) mbar_wait(empty + 8 * st, ((par >> st) & 1u) ^ 1u);
            }
        }
    } else if (warp == 1) {
        uint32_t par = 0, te_phase = 0;
        bool first_piece = true;
        long long base = 0;
        // A = dPre half blocks, B = X half blocks: MN-major SWIZZLE_128B, LBO = 8 KB (next 64 columns), SBO = 1 KB (next 8 rows of K)
        const uint64_t d_hi = make_desc(0, 8192, 1024, kLayoutSW128);
        for (int u = 0; u < n_wunits<KIND>(); ++u) {
            WPiece pc;
            if (!wpiece<KIND>(u, t_count, lo, hi, base, pc)) continue;
            const WUnit un = wunit<KIND>(u);
            const uint32_t idesc = make_idesc_bf16(128, (uint32_t)un.x_nb * 64u) | (1u << 15) | (1u << 16);
            const int n_m = un.g_nb >> 1;
            if (!first_piece) { mbar_wait(tmem_empty, te_phase); te_phase ^= 1u; tc_fence_after(); }
            first_piece = false;
            uint32_t acc = 0;
            const int n_st = wstages(un.g_nb + un.x_nb);
            const uint32_t st_bytes = (uint32_t)(un.g_nb + un.x_nb) * 8192u;
            int stage = 0;
            for (long long st = 0; st < 2 * (pc.t1 - pc.t0); ++st) {
                mbar_wait(full + 8 * stage, (par >> stage) & 1u);
                par ^= 1u << stage;
                tc_fence_after();
                const uint32_t a_s = smem + (uint32_t)stage * st_bytes, b_s = a_s + (uint32_t)un.g_nb * 8192u;
                if (elect_one()) {
                    for (int m = 0; m < n_m; ++m) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t ad = d_hi | (uint64_t)(((a_s + (uint32_t)m * 16384u + (uint32_t)k * 2048u) >> 4) & 0x3FFFu);
                            const uint64_t bd = d_hi | (uint64_t)(((b_s + (uint32_t)k * 2048u) >> 4) & 0x3FFFu);
                            mma_bf16(tmem + (uint32_t)m * 256u, ad, bd, idesc, acc | (uint32_t)k);
                        }
                    }
                    mma_commit(empty + 8 * stage);
                }
                __syncwarp();
                acc = 1;
                if (++stage == n_st) stage = 0;
            }
            if (elect_one()) mma_commit(acc_full);
            __syncwarp();
        }
    } else {
        const int w = warp - 2;                    // 0..3: G block reduced by this warp
        const int quad = warp & 3;                 // TMEM lane quadrant this warp may read
        uint32_t par = 0, af_phase = 0;
        long long base = 0;
        for (int u = 0; u < n_wunits<KIND>(); ++u) {
            WPiece pc;
            if (!wpiece<KIND>(u, t_count, lo, hi, base, pc)) continue;
            const WUnit un = wunit<KIND>(u);
            const LayerDesc L = wlayer<KIND>(un.layer);
            const bool do_bias = un.bias && w < un.g_nb;
            float s0 = 0.f, s1 = 0.f;
            const int n_st = wstages(un.g_nb + un.x_nb);
            const uint32_t st_bytes = (uint32_t)(un.g_nb + un.x_nb) * 8192u;
            int stage = 0;
            for (long long st = 0; st < 2 * (pc.t1 - pc.t0); ++st) {
                mbar_wait(full + 8 * stage, (par >> stage) & 1u);
                par ^= 1u << stage;
                if (do_bias) {
                    const uint32_t blk = smem + (uint32_t)stage * st_bytes + (uint32_t)w * 8192u + (uint32_t)(lane & 3) * 4u;
#pragma unroll 8
                    for (uint32_t row = 0; row < 64; ++row) {
                        uint32_t x;
                        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(x) : "r"(blk + row * 128u + ((((uint32_t)lane >> 2) ^ (row & 7u)) << 4)));
                        s0 += __uint_as_float(x << 16);
                        s1 += __uint_as_float(x & 0xFFFF0000u);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty + 8 * stage);
                if (++stage == n_st) stage = 0;
            }
            // ---- flush the accumulators of this piece
            mbar_wait(acc_full, af_phase);
            af_phase ^= 1u;
            tc_fence_after();
            const int n_m = un.g_nb >> 1;
            float* __restrict__ dW = d_params + L.w_off;
            for (int m = 0; m < n_m; ++m) {
                const long long o = m * 128 + quad * 32 + lane;
                for (int j = 0; j < un.x_nb * 2; ++j) {
                    uint32_t v[32];
                    tmem_ld32(tmem + ((uint32_t)quad << 21) + (uint32_t)m * 256u + (uint32_t)j * 32u, v);
                    tmem_ld_wait();
                    // the thread's 32 values are 128 contiguous bytes of dW's row o: 16-byte vector reductions where the group is whole and aligned
                    float* __restrict__ prow = dW + o * L.in + un.col_off + (j * 32 - un.x_col0);
#pragma unroll
                    for (int e = 0; e < 32; e += 4) {
                        const int col = j * 32 + e - un.x_col0;
                        if (col >= 0 && col + 3 < un.n_valid && ((reinterpret_cast<uintptr_t>(prow + e) & 15) == 0)) {
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(prow + e), "f"(un.scale * __uint_as_float(v[e])),
                                         "f"(un.scale * __uint_as_float(v[e + 1])), "f"(un.scale * __uint_as_float(v[e + 2])),
                                         "f"(un.scale * __uint_as_float(v[e + 3])) : "memory");
                        } else {
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                if (col + q >= 0 && col + q < un.n_valid) atomicAdd(prow + e + q, un.scale * __uint_as_float(v[e + q]));
                        }
                    }
                }
            }
            if (do_bias) {
                atomicAdd(d_params + L.b_off + w * 64 + 2 * lane, un.scale * s0);
                atomicAdd(d_params + L.b_off + w * 64 + 2 * lane + 1, un.scale * s1);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

// output_layer_sigma (256 -> 1, input h7) and output_layer_rgb (128 -> 3, input h_d): weight and bias gradients from the
// head gradients HG and the saved tiles.  A thread owns one 16-byte chunk (8 columns) of a tile row: warps 0-3 walk the
// rows of h7 (a warp reads one 512-byte row of the 4 blocks per step), warps 4-7 the rows of h_d (a half warp per row),
// 8 rows in flight per thread.
__device__ __forceinline__ void fma8(float (&acc)[8], float g, const uint4& x) {
    acc[0] = fmaf(g, __uint_as_float(x.x << 16), acc[0]); acc[1] = fmaf(g, __uint_as_float(x.x & 0xFFFF0000u), acc[1]);
    acc[2] = fmaf(g, __uint_as_float(x.y << 16), acc[2]); acc[3] = fmaf(g, __uint_as_float(x.y & 0xFFFF0000u), acc[3]);
    acc[4] = fmaf(g, __uint_as_float(x.z << 16), acc[4]); acc[5] = fmaf(g, __uint_as_float(x.z & 0xFFFF0000u), acc[5]);
    acc[6] = fmaf(g, __uint_as_float(x.w << 16), acc[6]); acc[7] = fmaf(g, __uint_as_float(x.w & 0xFFFF0000u), acc[7]);
}
// h7_off / hd_off: block offsets of the two input tensors in `saved`; Ls / Lc: the heads' slots in the flat parameter vector
__global__ void __launch_bounds__(256) nerf_head_wgrad_kernel(const uint8_t* __restrict__ saved, const uint8_t* __restrict__ scratch, long long n_sub,
                                                              float* __restrict__ d_params, int h7_off, int hd_off, LayerDesc Ls, LayerDesc Lc) {
    __shared__ float red_s[4][256 + 1];          // sigma: per-warp partial weight gradients (+ bias sum)
    __shared__ float red_c[8][384 + 3];          // rgb: per half-warp partials (+ 3 bias sums)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float4* __restrict__ hg = reinterpret_cast<const float4*>(scratch + (size_t)kScrBlocks * n_sub * kBlk);
    const uint32_t c = (uint32_t)lane & 7u;                 // logical chunk of the row this thread owns (columns blk*64 + c*8 ..)
    if (warp < 4) {
        const int blk = lane >> 3;
        const uint8_t* __restrict__ xt = saved + (size_t)h7_off * n_sub * kBlk + (size_t)blk * kBlk;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, bsum = 0.f;
        for (long long T = blockIdx.x; T < n_sub; T += gridDim.x) {
            const uint8_t* tile = xt + (size_t)T * 4 * kBlk;
            const float4* g = hg + T * kRowsSub;
#pragma unroll 1
            for (int i0 = 0; i0 < 32; i0 += 8) {
                uint4 x[8]; float gs[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint32_t row = (uint32_t)(warp + 4 * (i0 + i));
                    x[i] = ldg128(tile + row * 128u + ((c ^ (row & 7u)) << 4));
                    gs[i] = __ldg(&g[row].w);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) { fma8(acc, gs[i], x[i]); bsum += gs[i]; }
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) red_s[warp][blk * 64 + (int)c * 8 + j] = acc[j];
        if (lane == 0) red_s[warp][256] = bsum;
    } else {
        const int blk = (lane >> 3) & 1, sub = lane >> 4;   // half warp `sub` takes every other row
        const uint8_t* __restrict__ xt = saved + (size_t)hd_off * n_sub * kBlk + (size_t)blk * kBlk;
        float a0[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, a1[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f},
              a2[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, b0 = 0.f, b1 = 0.f, b2 = 0.f;
        for (long long T = blockIdx.x; T < n_sub; T += gridDim.x) {
            const uint8_t* tile = xt + (size_t)T * 2 * kBlk;
            const float4* g = hg + T * kRowsSub;
#pragma unroll 1
            for (int i0 = 0; i0 < 16; i0 += 4) {
                uint4 x[4]; float4 gg[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t row = (uint32_t)((warp - 4) * 2 + sub + 8 * (i0 + i));
                    x[i] = ldg128(tile + row * 128u + ((c ^ (row & 7u)) << 4));
                    gg[i] = __ldg(&g[row]);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    fma8(a0, gg[i].x, x[i]); fma8(a1, gg[i].y, x[i]); fma8(a2, gg[i].z, x[i]);
                    b0 += gg[i].x; b1 += gg[i].y; b2 += gg[i].z;
                }
            }
        }
        float* dst = red_c[(warp - 4) * 2 + sub];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int col = blk * 64 + (int)c * 8 + j;
            dst[col] = a0[j]; dst[128 + col] = a1[j]; dst[256 + col] = a2[j];
        }
        if ((lane & 15) == 0) { dst[384] = b0; dst[385] = b1; dst[386] = b2; }
    }
    __syncthreads();
    // one atomic per output and CTA
    const int t = threadIdx.x;
    {
        const LayerDesc L = Ls;
        atomicAdd(d_params + L.w_off + t, red_s[0][t] + red_s[1][t] + red_s[2][t] + red_s[3][t]);
        if (t == 0) atomicAdd(d_params + L.b_off, red_s[0][256] + red_s[1][256] + red_s[2][256] + red_s[3][256]);
    }
    {
        const LayerDesc L = Lc;
        for (int i = t; i < 387; i += 256) {
            float sum = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) sum += red_c[k][i];
            atomicAdd(i < 384 ? d_params + L.w_off + i : d_params + L.b_off + (i - 384), sum);
        }
    }
}

// =====================================================================================================================
// FiLM-SIREN reverse mode (pi_GAN/train.py:134 / synthesis.py:107: autograd through FilmSirenNeRF.forward, modules.py:101-118)
// =====================================================================================================================
// The forward (film_tc_kernel<true>, mlp_tc.cu) evaluates t_l = W'_l x + s'_l with the FiLM scale folded into the weights
// (W' = 30 gamma W, s' = 30 (gamma b + beta)) and keeps h_l = sin(t_l) as tiles and cos(t_l) thread-major.  G_l = dL/dt_l:
//   G8 = (d rgb_pre . W_rgb) cos(t8)                                   CUDA cores (K = 3), written as the first A operand
//   step 0: G7 = (G8 . W'_rgb[:, 0:256] + gs w_sigma) cos(t7)          hidden_layer_rgb
//   step s = 1..7: G_{7-s} = (G_{8-s} . W'_{hidden_layers.(7-s)}) cos(t_{7-s})
// Every G_l is spilled for wgrad (units c_wunits_film), which yields d W' and d s'; film_grad_finish_kernel recovers
//   d W = 30 gamma (.) d W',  d b = 30 gamma d s',  d gamma = 30 (sum_j W_ij d W'_ij + b_i d s'_i),  d beta = 30 d s'.
struct FilmBwdSched {
    static constexpr int kSteps = 8;
    __host__ __device__ static constexpr int n_pre(int, int) { return 0; }
    __host__ __device__ static constexpr int n_h(int, int) { return 4; }
    __host__ __device__ static constexpr int n_post(int, int) { return 0; }
    __host__ __device__ static constexpr int n(int) { return 256; }
    static constexpr int kPostMmas = 1;
};
constexpr long long kFilmBwdChunkBytes = step_base<FilmBwdSched>(FilmBwdSched::kSteps);
static_assert(kFilmBwdChunkBytes == 32LL * 32768, "FiLM bwd packed chunk bytes");
constexpr int kFBwdTabWSigma = 0, kFBwdTabWRgb = 256, kFBwdTabFloats = 1024;          // w_sigma[256] | w_rgb[3][256]
constexpr long long kFilmBwdPackedBytes = kFilmBwdChunkBytes + kFBwdTabFloats * 4;
static_assert(kFBwdTabFloats * 4 <= (int)kTabBytes, "FiLM bwd table region");
// film_layer index / film row of dgrad step s
__host__ __device__ constexpr int fbwd_layer(int s) { return s == 0 ? 9 : 8 - s; }
__host__ __device__ constexpr int fbwd_film_row(int s) { return 8 - s; }

// B[n][k] = 30 gamma_k W[k][n]: n = input feature (output column of dX), k = output feature (the forward's folded row k)
// blockIdx.y = latent: film [n_latents][9][512] -> n_latents images of kFilmBwdPackedBytes
__global__ void film_pack_bwd_kernel(const float* __restrict__ params, const float* __restrict__ film_all, uint8_t* __restrict__ packed_all, int use_dir) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const bool ud = use_dir != 0;
    const float* __restrict__ film = film_all + (size_t)blockIdx.y * B2R_FILM_PARAMS;
    uint8_t* __restrict__ packed = packed_all + (size_t)blockIdx.y * kFilmBwdPackedBytes;
    if (t < kFilmBwdChunkBytes / 16) {
        int s, c, hf, row, grp;
        locate<FilmBwdSched>(t * 16, s, c, hf, row, grp);
        const LayerDesc L = film_layer(fbwd_layer(s), ud);
        const float* __restrict__ gamma = film + fbwd_film_row(s) * 512;
        const int n = hf * 128 + row;
        __nv_bfloat16 v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int k = c * 64 + grp * 8 + e;
            v[e] = __float2bfloat16_rn(30.0f * gamma[k] * params[L.w_off + (long long)k * L.in + n]);
        }
        uint8_t* dst = packed + step_base<FilmBwdSched>(s) + (long long)(c * 2 + hf) * half_bytes<FilmBwdSched>(s) +
                       sw128_offset((uint32_t)row, (uint32_t)grp);
        *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(v);
    }
    if (t < kFBwdTabFloats) {
        float* tab = reinterpret_cast<float*>(packed + kFilmBwdChunkBytes);
        const int i = (int)t;
        tab[i] = i < kFBwdTabWRgb ? params[film_layer(8, ud).w_off + i] : params[film_layer(10, ud).w_off + (i - kFBwdTabWRgb)];
    }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
film_tc_bwd_kernel(const uint8_t* __restrict__ packed, long long rows, const float4* __restrict__ raw, const float4* __restrict__ d_raw,
                   const uint8_t* __restrict__ saved, uint8_t* __restrict__ scratch, int n_latents, long long rows_per_latent) {
    extern __shared__ uint8_t smem_raw[];
    const Ctx cx = make_ctx(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const PairLoop pl(rows);
    {   // w_sigma | w_rgb -> shared memory
        const float4* tab_g = reinterpret_cast<const float4*>(packed + kFilmBwdChunkBytes);
        for (int i = threadIdx.x; i < kFBwdTabFloats / 4; i += kThreads) {
            float4 v = __ldg(tab_g + i);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(cx.smem + kTabOff + 16u * i), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
        }
    }
    const uint32_t tmem_base = tc_prologue(cx, warp);

    if (warp == 0) {
        // batched: rows [b * rows_per_latent, (b+1) * rows_per_latent) use latent b's transposed folded weights (a tile pair never straddles)
        auto packed_of = [&](long long p) -> const uint8_t* {
            return packed + (n_latents > 1 ? (size_t)((2 * p * kRowsTile) / rows_per_latent) * kFilmBwdPackedBytes : (size_t)0);
        };
        if (lane == 0) producer_loop_fn<FilmBwdSched>(cx, packed_of, pl, FilmBwdSched::kSteps, 0);
    } else if (warp == 1) {
        if (cx.rank == 0) mma_loop<FilmBwdSched>(cx, tmem_base, pl, FilmBwdSched::kSteps, 0);
        else if (lane == 0) relay_loop<FilmBwdSched>(cx, pl, FilmBwdSched::kSteps, 0);
    } else if (warp < kCtrlWarps) {
        if (lane == 0) {
            // ===== spill thread of sub-tile g: G8, G7 .. G0 (the A operands) -> tiled tensors in `scratch` =====
            const int g = warp - 2;
            const uint32_t hreg = cx.smem + (uint32_t)g * kSubBytes + kPeBytes;
            const uint32_t ready = cx.spill_ready + 8 * g, done = cx.spill_done + 8 * g;
            const size_t n_sub = (size_t)pl.n_pairs * 4;
            uint32_t ph = 0;
            for (long long p = pl.first; p < pl.n_pairs; p += pl.stride) {
                const size_t T = (size_t)((2 * p + cx.rank) * 2 + g);
                for (int l = 8; l >= 0; --l) {
                    mbar_wait(ready, ph);
                    spill_tile(scratch + ((size_t)fscr_g(l) * n_sub + T * 4) * kBlk, hreg, 4);
                    bulk_commit(); bulk_wait_read(); mbar_arrive(done); ph ^= 1u;
                }
            }
            bulk_wait_all();
        }
    } else {
        const int ew = warp - kCtrlWarps;
        const int g = ew >> 3, half = (ew >> 2) & 1, quad = ew & 3;
        const int r = (quad << 5) | lane;
        const uint32_t sub = cx.smem + (uint32_t)g * kSubBytes;
        const uint32_t h_base = sub + kPeBytes;
        const uint32_t t_addr = tmem_base + ((uint32_t)quad << 21) + (uint32_t)g * 256u;
        const uint32_t row_off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
        const uint32_t xr = (uint32_t)(r & 7);
        const uint32_t tab = cx.smem + kTabOff;
        const uint32_t act_local = cx.act_ready + 8 * g, act_leader = mapa(act_local, 0);
        const uint32_t acc_bar = cx.acc_full + 8 * g;
        const uint32_t ready_bar = cx.spill_ready + 8 * g, done_bar = cx.spill_done + 8 * g;
        const uint32_t t_half = t_addr + (uint32_t)half * 128u;
        const uint32_t h_half = h_base + row_off + (uint32_t)half * 2u * kBlk;
        const uint32_t wsig_half = tab + (uint32_t)(kFBwdTabWSigma + half * 128) * 4u;
        uint32_t xoff[8];
#pragma unroll
        for (uint32_t c = 0; c < 8; ++c) xoff[c] = (c ^ xr) << 4;
        uint32_t acc_phase = 0, sp_phase = 0;
        bool first_tile = true;
        const size_t n_sub = (size_t)pl.n_pairs * 4;
        float4* __restrict__ hg_out = reinterpret_cast<float4*>(scratch + (size_t)kFScrBlocks * n_sub * kBlk);
        for (long long p = pl.first; p < pl.n_pairs; p += pl.stride) {
            const long long row = (2 * p + cx.rank) * kRowsTile + g * kRowsSub + r;
            const bool valid = row < rows;
            const size_t T = (size_t)((2 * p + cx.rank) * 2 + g);
            auto spill_sig = [&]() { if (lane == 0) mbar_arrive(ready_bar); };
            // ---- heads: d raw -> (d rgb pre-sigmoid, d sigma pre-relu); rows past the end contribute zero everywhere
            float gc0 = 0.f, gc1 = 0.f, gc2 = 0.f, gs = 0.f;
            if (valid) {
                const float4 y = __ldg(raw + row), dy = __ldg(d_raw + row);
                gc0 = dy.x * (y.x * (1.0f - y.x));
                gc1 = dy.y * (y.y * (1.0f - y.y));
                gc2 = dy.z * (y.z * (1.0f - y.z));
                gs = y.w > 0.f ? dy.w : 0.f;
            }
            if (half == 0) hg_out[T * kRowsSub + r] = make_float4(gc0, gc1, gc2, gs);
            {
                // G8 = (d rgb pre . W_rgb) cos(t8): this half produces columns half*128 .. +127 = K-blocks 2*half, 2*half + 1
                const uint32_t wr = tab + (uint32_t)(kFBwdTabWRgb + half * 128) * 4u;
                if (!first_tile) { mbar_wait(done_bar, sp_phase); sp_phase ^= 1u; }          // previous tile's G0 copy
                first_tile = false;
                uint4 cwq = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll 4
                for (int q = 0; q < 16; ++q) {
                    if ((q & 1) == 0) cwq = ldg128(saved + film_cos_off(n_sub, 8, T, half * 2 + (q >> 3), (q & 7) >> 1, r));
                    float f[8];
#pragma unroll
                    for (int hq = 0; hq < 2; ++hq) {
                        const uint32_t wa = wr + (uint32_t)(q * 8 + hq * 4) * 4u;
                        const float4 w0 = lds128(wa), w1 = lds128(wa + 1024u), w2 = lds128(wa + 2048u);
                        f[4 * hq + 0] = fmaf(gc2, w2.x, fmaf(gc1, w1.x, gc0 * w0.x));
                        f[4 * hq + 1] = fmaf(gc2, w2.y, fmaf(gc1, w1.y, gc0 * w0.y));
                        f[4 * hq + 2] = fmaf(gc2, w2.z, fmaf(gc1, w1.z, gc0 * w0.z));
                        f[4 * hq + 3] = fmaf(gc2, w2.w, fmaf(gc1, w1.w, gc0 * w0.w));
                    }
                    cos_mul8(f, (q & 1) ? cwq.z : cwq.x, (q & 1) ? cwq.w : cwq.y);
                    const uint32_t w0 = pack_bf16(f[0], f[1]), w1 = pack_bf16(f[2], f[3]), w2 = pack_bf16(f[4], f[5]), w3 = pack_bf16(f[6], f[7]);
                    st_shared_v4(h_half + (uint32_t)(q >> 3) * kBlk + ((((uint32_t)q & 7u) ^ xr) << 4), w0, w1, w2, w3);
                }
            }
            arrive_act(act_local, act_leader, cx.rank, lane);
            spill_sig();
            // cos(t_l) checkpoint of this thread: first word of quarter 2 * half
            auto mptr = [&](int l) -> const uint8_t* { return saved + film_cos_off(n_sub, l, T, half * 2, 0, r); };
            // step 0: G7 (+ sigma head)
            bwd_epi<1, true>(t_half, h_half, xoff, mptr(7), gs, wsig_half, acc_bar, acc_phase, done_bar, sp_phase);
            arrive_act(act_local, act_leader, cx.rank, lane);
            spill_sig();
            // steps 1..6: G6 .. G1
            for (int s = 1; s < 7; ++s) {
                bwd_epi<2, true>(t_half, h_half, xoff, mptr(7 - s), 0.f, 0u, acc_bar, acc_phase, done_bar, sp_phase);
                arrive_act(act_local, act_leader, cx.rank, lane);
                spill_sig();
            }
            // step 7: G0: only wgrad reads it (the input layer's input is the position), no MMA follows
            bwd_epi<2, true>(t_half, h_half, xoff, mptr(0), 0.f, 0u, acc_bar, acc_phase, done_bar, sp_phase);
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            spill_sig();
        }
    }
    tc_teardown(tmem_base, warp);
}

// output_layer_sigma (256 -> 1, input h7) and output_layer_rgb (256 -> 3, input = hidden_layer_rgb's output HC): weight and
// bias gradients from HG and the saved tiles.  A thread owns one 16-byte chunk (8 columns) of a 256-column row (a warp reads
// one 512-byte row of the 4 blocks per step); the 8 warps take every 8th row, 4 rows in flight per thread.
__global__ void __launch_bounds__(256) film_head_wgrad_kernel(const uint8_t* __restrict__ saved, const uint8_t* __restrict__ scratch, long long n_sub,
                                                              float* __restrict__ d_out, LayerDesc Ls, LayerDesc Lc) {
    __shared__ float red[8][1024 + 4];           // per warp: d w_sigma[256] | d w_rgb[3][256] | 4 bias sums (sigma, rgb)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float4* __restrict__ hg = reinterpret_cast<const float4*>(scratch + (size_t)kFScrBlocks * n_sub * kBlk);
    const uint32_t c = (uint32_t)lane & 7u;
    const int blk = lane >> 3;
    float as[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, a0[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f},
          a1[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, a2[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float bs = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f;
    for (long long T = blockIdx.x; T < n_sub; T += gridDim.x) {
        const uint8_t* t7 = saved + ((size_t)fs_h(7) * n_sub + (size_t)T * 4 + blk) * kBlk;
        const uint8_t* tc = saved + ((size_t)kFsHC * n_sub + (size_t)T * 4 + blk) * kBlk;
        const float4* g = hg + T * kRowsSub;
#pragma unroll 1
        for (int i0 = 0; i0 < 16; i0 += 4) {
            uint4 x7[4], xc[4]; float4 gg[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t row = (uint32_t)(warp + 8 * (i0 + i));
                const uint32_t o = row * 128u + ((c ^ (row & 7u)) << 4);
                x7[i] = ldg128(t7 + o);
                xc[i] = ldg128(tc + o);
                gg[i] = __ldg(&g[row]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                fma8(as, gg[i].w, x7[i]); fma8(a0, gg[i].x, xc[i]); fma8(a1, gg[i].y, xc[i]); fma8(a2, gg[i].z, xc[i]);
                bs += gg[i].w; b0 += gg[i].x; b1 += gg[i].y; b2 += gg[i].z;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int col = blk * 64 + (int)c * 8 + j;
        red[warp][col] = as[j]; red[warp][256 + col] = a0[j]; red[warp][512 + col] = a1[j]; red[warp][768 + col] = a2[j];
    }
    if (lane == 0) { red[warp][1024] = bs; red[warp][1025] = b0; red[warp][1026] = b1; red[warp][1027] = b2; }   // every lane saw the same rows
    __syncthreads();
    for (int i = threadIdx.x; i < 1028; i += 256) {
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) sum += red[k][i];
        float* dst = i < 256 ? d_out + Ls.w_off + i : (i < 1024 ? d_out + Lc.w_off + (i - 256) : (i == 1024 ? d_out + Ls.b_off : d_out + Lc.b_off + (i - 1025)));
        atomicAdd(dst, sum);
    }
}

// folded gradients -> parameter / FiLM gradients.  One warp per (FiLM layer, output feature i); blockIdx.y = latent (its film row
// block, its d_folded, its d_film).  d_params (shared by the latents: float atomics) and d_film are ACCUMULATED into; either may be NULL.
__global__ void __launch_bounds__(256) film_grad_finish_kernel(const float* __restrict__ params, const float* __restrict__ film_all,
                                                               const float* __restrict__ d_folded_all, float* __restrict__ d_params,
                                                               float* __restrict__ d_film_all, int use_dir) {
    const int gw = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (gw >= 9 * 256) return;
    const float* __restrict__ film = film_all + (size_t)blockIdx.y * B2R_FILM_PARAMS;
    const float* __restrict__ d_folded = d_folded_all + (size_t)blockIdx.y * (use_dir ? B2R_FILM_NUMEL : B2R_FILM_NODIR_NUMEL);
    const int fl = gw >> 8, i = gw & 255;
    const LayerDesc L = film_layer(fl == 8 ? 9 : fl, use_dir != 0);
    const float gsc = 30.0f * film[fl * 512 + i];
    const float* __restrict__ w = params + L.w_off + (long long)i * L.in;
    const float* __restrict__ dwp = d_folded + L.w_off + (long long)i * L.in;
    float acc = 0.f;
    for (int j = lane; j < L.in; j += 32) {
        const float d = dwp[j];
        acc = fmaf(w[j], d, acc);
        if (d_params) atomicAdd(d_params + L.w_off + (long long)i * L.in + j, gsc * d);
    }
    acc = warp_sum(acc);
    if (lane == 0) {
        const float ds = d_folded[L.b_off + i];
        if (d_params) atomicAdd(d_params + L.b_off + i, gsc * ds);
        if (d_film_all) {
            float* __restrict__ d_film = d_film_all + (size_t)blockIdx.y * B2R_FILM_PARAMS;
            d_film[fl * 512 + i] += 30.0f * fmaf(params[L.b_off + i], ds, acc);
            d_film[fl * 512 + 256 + i] += 30.0f * ds;
        }
    }
}

int pair_grid(long long rows, unsigned* grid);   // mlp_tc.cu

}  // namespace tc
}  // namespace b2r

static inline bool train_kind(int k) { return k == B2R_MODEL_NERF || k == B2R_MODEL_SIREN; }

extern "C" size_t b2r_mlp_tc_bwd_packed_bytes(int model_kind) {
    if (model_kind == B2R_MODEL_FILM) return (size_t)b2r::tc::kFilmBwdPackedBytes;
    return train_kind(model_kind) ? (size_t)b2r::tc::kBwdPackedBytes : 0;
}

extern "C" int b2r_mlp_tc_pack_bwd_film(const float* params, const float* film, int use_dir, int n_latents, void* packed_out, void* stream) {
    using namespace b2r;
    B2R_CHECK_ARG(params && film && packed_out, "b2r_mlp_tc_pack_bwd_film: NULL pointer");
    B2R_CHECK_ARG(((uintptr_t)packed_out & 15) == 0, "b2r_mlp_tc_pack_bwd_film: packed_out must be 16-byte aligned");
    B2R_CHECK_ARG(n_latents >= 0 && n_latents <= 65535, "b2r_mlp_tc_pack_bwd_film: n_latents out of range");
    if (n_latents == 0) return 0;
    long long threads = tc::kFilmBwdChunkBytes / 16;
    tc::film_pack_bwd_kernel<<<dim3((unsigned)((threads + 255) / 256), (unsigned)n_latents), 256, 0, (cudaStream_t)stream>>>(params, film,
                                                                                                                            (uint8_t*)packed_out, use_dir);
    B2R_LAUNCH_CHECK("b2r_mlp_tc_pack_bwd_film");
    return 0;
}

extern "C" int b2r_mlp_tc_pack_bwd(int model_kind, const float* params, void* packed_out, void* stream) {
    using namespace b2r;
    B2R_CHECK_ARG(train_kind(model_kind), "b2r_mlp_tc_pack_bwd: only the NeRF and SirenNeRF models have a fused tensor-core training path (kind %d)", model_kind);
    B2R_CHECK_ARG(params && packed_out, "b2r_mlp_tc_pack_bwd: NULL pointer");
    B2R_CHECK_ARG(((uintptr_t)packed_out & 15) == 0, "b2r_mlp_tc_pack_bwd: packed_out must be 16-byte aligned");
    long long threads = tc::kBwdChunkBytes / 16;
    if (model_kind == B2R_MODEL_SIREN)
        tc::nerf_pack_bwd_kernel<true><<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(params, (uint8_t*)packed_out);
    else
        tc::nerf_pack_bwd_kernel<false><<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(params, (uint8_t*)packed_out);
    B2R_LAUNCH_CHECK("b2r_mlp_tc_pack_bwd");
    return 0;
}

extern "C" size_t b2r_mlp_tc_train_scratch_bytes(int model_kind, long long rows) {
    if (model_kind == B2R_MODEL_FILM && rows >= 0)
        return (size_t)b2r::tc::n_sub_tiles(rows) * ((size_t)b2r::tc::kFScrBlocks * b2r::tc::kBlk + b2r::tc::kRowsSub * 16);
    if (!train_kind(model_kind) || rows < 0) return 0;
    return (size_t)b2r::tc::n_sub_tiles(rows) * ((size_t)b2r::tc::kScrBlocks * b2r::tc::kBlk + b2r::tc::kRowsSub * 16);
}

template <bool kSiren>
static int train_bwd_launch(const void* packed_bwd, long long rows, const float* raw, const float* d_raw, const void* saved, void* scratch,
                            float* d_params, cudaStream_t st) {
    using namespace b2r;
    unsigned grid = 0;
    int rc = tc::pair_grid(rows, &grid);
    if (rc) return rc;
    rc = cuda_result(cudaFuncSetAttribute(tc::nerf_tc_bwd_kernel<kSiren>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kSmemBytes), "tc bwd smem attribute");
    if (rc) return rc;
    tc::nerf_tc_bwd_kernel<kSiren><<<grid, tc::kThreads, tc::kSmemBytes, st>>>((const uint8_t*)packed_bwd, rows, (const float4*)raw, (const float4*)d_raw,
                                                                              (const uint8_t*)saved, (uint8_t*)scratch);
    B2R_LAUNCH_CHECK("b2r_mlp_tc_train_bwd (dgrad)");
    const long long n_sub = tc::n_sub_tiles(rows);
    int dev = 0, sms = 0;
    rc = cuda_result(cudaGetDevice(&dev), "cudaGetDevice");
    if (rc) return rc;
    rc = cuda_result(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "SM count");
    if (rc) return rc;
    rc = cuda_result(cudaFuncSetAttribute(tc::nerf_tc_wgrad_kernel<kSiren ? tc::kKSiren : tc::kKNerf>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kWSmem), "tc wgrad smem attribute");
    if (rc) return rc;
    const long long work = n_sub * tc::kWUnits;
    unsigned wgrid = (unsigned)(work < sms ? work : sms);
    tc::nerf_tc_wgrad_kernel<kSiren ? tc::kKSiren : tc::kKNerf><<<wgrid, tc::kWThreads, tc::kWSmem, st>>>((const uint8_t*)saved, (const uint8_t*)scratch, n_sub, d_params, 0, n_sub);
    B2R_LAUNCH_CHECK("b2r_mlp_tc_train_bwd (wgrad)");
    unsigned hgrid = (unsigned)(n_sub < 4LL * sms ? n_sub : 4LL * sms);
    tc::nerf_head_wgrad_kernel<<<hgrid, 256, 0, st>>>((const uint8_t*)saved, (const uint8_t*)scratch, n_sub, d_params,
                                                      kSiren ? tc::ss_h(7) : tc::sav_h(7), kSiren ? tc::kSsHD : tc::kSavHD,
                                                      kSiren ? siren_layer(10) : nerf_layer(10), kSiren ? siren_layer(11) : nerf_layer(11));
    B2R_LAUNCH_CHECK("b2r_mlp_tc_train_bwd (heads)");
    return 0;
}

extern "C" int b2r_mlp_tc_train_bwd(int model_kind, const void* packed_bwd, long long rows, const float* raw, const float* d_raw,
                                    const void* saved, void* scratch, size_t scratch_bytes, float* d_params, void* stream) {
    using namespace b2r;
    B2R_CHECK_ARG(train_kind(model_kind), "b2r_mlp_tc_train_bwd: only the NeRF and SirenNeRF models have a fused tensor-core training path (kind %d)", model_kind);
    B2R_CHECK_ARG(packed_bwd && raw && d_raw && saved && scratch && d_params, "b2r_mlp_tc_train_bwd: NULL pointer");
    B2R_CHECK_ARG((((uintptr_t)packed_bwd | (uintptr_t)raw | (uintptr_t)d_raw | (uintptr_t)saved | (uintptr_t)scratch | (uintptr_t)d_params) & 15) == 0,
                  "b2r_mlp_tc_train_bwd: buffers must be 16-byte aligned");
    B2R_CHECK_ARG(rows >= 0, "b2r_mlp_tc_train_bwd: negative row count");
    B2R_CHECK_ARG(scratch_bytes >= b2r_mlp_tc_train_scratch_bytes(model_kind, rows), "b2r_mlp_tc_train_bwd: scratch too small (%zu B)", scratch_bytes);
    if (rows == 0) return 0;
    if (model_kind == B2R_MODEL_SIREN) return train_bwd_launch<true>(packed_bwd, rows, raw, d_raw, saved, scratch, d_params, (cudaStream_t)stream);
    return train_bwd_launch<false>(packed_bwd, rows, raw, d_raw, saved, scratch, d_params, (cudaStream_t)stream);
}

extern "C" int b2r_mlp_tc_train_bwd_film(const void* packed_bwd, const float* params, const float* film, int use_dir, int n_latents, long long rows_per_latent,
                                         long long rows, const float* raw, const float* d_raw, const void* saved, void* scratch,
                                         size_t scratch_bytes, float* d_folded, float* d_params, float* d_film, void* stream) {
    using namespace b2r;
    B2R_CHECK_ARG(packed_bwd && params && film && raw && d_raw && saved && scratch && d_folded, "b2r_mlp_tc_train_bwd_film: NULL pointer");
    B2R_CHECK_ARG((((uintptr_t)packed_bwd | (uintptr_t)raw | (uintptr_t)d_raw | (uintptr_t)saved | (uintptr_t)scratch | (uintptr_t)d_folded) & 15) == 0,
                  "b2r_mlp_tc_train_bwd_film: buffers must be 16-byte aligned");
    B2R_CHECK_ARG(rows >= 0 && n_latents >= 1 && n_latents <= 65535, "b2r_mlp_tc_train_bwd_film: bad row / latent count");
    B2R_CHECK_ARG(n_latents == 1 || (rows_per_latent > 0 && rows_per_latent % (2 * tc::kRowsTile) == 0 && rows <= rows_per_latent * (long long)n_latents),
                  "b2r_mlp_tc_train_bwd_film: rows_per_latent (%lld) must be a positive multiple of %d covering all rows", rows_per_latent, 2 * tc::kRowsTile);
    B2R_CHECK_ARG(scratch_bytes >= b2r_mlp_tc_train_scratch_bytes(B2R_MODEL_FILM, rows), "b2r_mlp_tc_train_bwd_film: scratch too small (%zu B)", scratch_bytes);
    if (rows == 0 || (!d_params && !d_film)) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = cuda_result(cudaMemsetAsync(d_folded, 0, (size_t)n_latents * (use_dir ? B2R_FILM_NUMEL : B2R_FILM_NODIR_NUMEL) * sizeof(float), st), "cudaMemsetAsync(d_folded)");
    if (rc) return rc;
    unsigned grid = 0;
    rc = tc::pair_grid(rows, &grid);
    if (rc) return rc;
    rc = cuda_result(cudaFuncSetAttribute(tc::film_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kSmemBytes), "tc bwd smem attribute");
    if (rc) return rc;
    tc::film_tc_bwd_kernel<<<grid, tc::kThreads, tc::kSmemBytes, st>>>((const uint8_t*)packed_bwd, rows, (const float4*)raw, (const float4*)d_raw,
                                                                      (const uint8_t*)saved, (uint8_t*)scratch, n_latents, rows_per_latent);
    B2R_LAUNCH_CHECK("b2r_mlp_tc_train_bwd_film (dgrad)");
    const long long n_sub = tc::n_sub_tiles(rows);
    int dev = 0, sms = 0;
    rc = cuda_result(cudaGetDevice(&dev), "cudaGetDevice");
    if (rc) return rc;
    rc = cuda_result(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "SM count");
    if (rc) return rc;
    const bool ud = use_dir != 0;
    auto wkern = ud ? tc::nerf_tc_wgrad_kernel<tc::kKFilm> : tc::nerf_tc_wgrad_kernel<tc::kKFilmNoDir>;
    const int n_units = ud ? tc::n_wunits<tc::kKFilm>() : tc::n_wunits<tc::kKFilmNoDir>();
    const size_t numel = ud ? B2R_FILM_NUMEL : B2R_FILM_NODIR_NUMEL;
    rc = cuda_result(cudaFuncSetAttribute(wkern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kWSmem), "tc wgrad smem attribute");
    if (rc) return rc;
    // one wgrad launch per latent over its own sub-tiles (the folded weights, hence their gradients, are per latent)
    const long long sub_per_latent = n_latents > 1 ? rows_per_latent / tc::kRowsSub : n_sub;
    for (int b = 0; b < n_latents; ++b) {
        const long long t0 = (long long)b * sub_per_latent;
        const long long tn = (t0 + sub_per_latent <= n_sub ? sub_per_latent : n_sub - t0);
        if (tn <= 0) break;
        const long long work = tn * n_units;
        unsigned wgrid = (unsigned)(work < sms ? work : sms);
        wkern<<<wgrid, tc::kWThreads, tc::kWSmem, st>>>((const uint8_t*)saved, (const uint8_t*)scratch, n_sub, d_folded + (size_t)b * numel, t0, tn);
        B2R_LAUNCH_CHECK("b2r_mlp_tc_train_bwd_film (wgrad)");
    }
    if (d_params) {    // the heads are not FiLM-folded: plain gradients straight into d_params
        unsigned hgrid = (unsigned)(n_sub < 2LL * sms ? n_sub : 2LL * sms);
        tc::film_head_wgrad_kernel<<<hgrid, 256, 0, st>>>((const uint8_t*)saved, (const uint8_t*)scratch, n_sub, d_params, film_layer(8, ud),
                                                          film_layer(10, ud));
        B2R_LAUNCH_CHECK("b2r_mlp_tc_train_bwd_film (heads)");
    }
    tc::film_grad_finish_kernel<<<dim3(9 * 256 / 8, (unsigned)n_latents), 256, 0, st>>>(params, film, d_folded, d_params, d_film, use_dir);
    B2R_LAUNCH_CHECK("b2r_mlp_tc_train_bwd_film (finish)");
    return 0;
}
