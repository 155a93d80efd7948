// K3 for SirenNeRF (nerf/nerf.py:97-170; `use_siren`, nerf/train_nerf.py:89-91) on the tensor cores: the same CTA-pair
// machinery as nerf_tc_kernel / film_tc_kernel (tc_core.cuh) with SirenNeRF's layer program
//   layers_pos.0 (3 -> 256)          CUDA cores, fp32, in the input stage (K = 3 is no GEMM; its 30x-amplified argument must
//                                    not see bf16 inputs)
//   layers_pos.1..4                  tcgen05 steps 0..3, epilogue sin(30 acc + 30 b)
//   layers_pos.5 on [pos | h4]       step 4: 4 h chunks + one 16-K chunk holding the raw position (aux block)
//   layers_pos.6, 7                  steps 5, 6; output_layer_sigma (256 -> 1) rides on step 6's epilogue in fp32
//   layers_dir.0 (linear)            step 7, epilogue acc + b
//   layers_dir.1 on [g | dir]        step 8: N = 128, 4 h chunks + one 16-K chunk holding the view direction, epilogue sin;
//                                    output_layer_rgb (128 -> 3) on its epilogue in fp32
// Algorithmic work: 2 * (3*256 + 4*65536 + 259*256 + 2*65536 + 256 + 65536 + 259*128 + 384) = 1,123,840 FLOP per row.
#include "tc_core.cuh"

namespace b2r {
namespace tc {

struct SirenSched {
    static constexpr int kSteps = 9;
    __host__ __device__ static constexpr int n_pre(int, int) { return 0; }
    __host__ __device__ static constexpr int n_h(int, int) { return 4; }
    __host__ __device__ static constexpr int n_post(int s, int) { return (s == 4 || s == 8) ? 1 : 0; }
    __host__ __device__ static constexpr int n(int s) { return s == 8 ? 128 : 256; }
    static constexpr int kPostMmas = 1;                  // 3 raw inputs -> 16 K
};
constexpr long long kSirenChunkBytes = step_base<SirenSched>(SirenSched::kSteps);
static_assert(kSirenChunkBytes == (4 * 4 + 5 + 2 * 4 + 4) * 32768LL + 5 * 16384LL, "siren packed chunk bytes");
// fp32 tables: shift[9][256] (30 b for the sine steps, b for layers_dir.0) | w0[3][256] (layers_pos.0, column-major) |
//              shift0[256] | w_sigma[256] | w_rgb[3][128] | b_sigma, b_rgb[3]
constexpr int kSSh = 0, kSW0 = 2304, kST0 = 3072, kSWS = 3328, kSWR = 3584, kSBH = 3968, kSirenTabFloats = 3972;
constexpr long long kSirenPackedBytes = kSirenChunkBytes + kSirenTabFloats * 4;
static_assert(kSirenTabFloats * 4 <= (int)(kTabBytes + kPartBytes), "siren tables fit the table region");

__host__ __device__ constexpr int siren_step_layer(int s) { return s + 1; }      // steps 0..8 = siren_layer 1..9

__global__ void siren_pack_kernel(const float* __restrict__ params, uint8_t* __restrict__ packed) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t < kSirenChunkBytes / 16) {
        int s, c, hf, row, grp;
        locate<SirenSched>(t * 16, s, c, hf, row, grp);
        LayerDesc L = siren_layer(siren_step_layer(s));
        const int n = hf * (SirenSched::n(s) / 2) + row;
        __nv_bfloat16 v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            int kk = grp * 8 + e;
            int col;
            if (s == 4) col = c < 4 ? 3 + c * 64 + kk : (kk < 3 ? kk : -1);                 // [pos | h4] (nerf/nerf.py:158)
            else if (s == 8) col = c < 4 ? c * 64 + kk : (kk < 3 ? 256 + kk : -1);          // [g | dir] (nerf/nerf.py:166)
            else col = c * 64 + kk;
            v[e] = __float2bfloat16_rn(col >= 0 ? params[L.w_off + (long long)n * L.in + col] : 0.f);
        }
        uint8_t* dst = packed + step_base<SirenSched>(s) + (long long)(c * 2 + hf) * half_bytes<SirenSched>(s) +
                       sw128_offset((uint32_t)row, (uint32_t)grp);
        *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(v);
    }
    if (t < kSirenTabFloats) {
        float* tab = reinterpret_cast<float*>(packed + kSirenChunkBytes);
        int i = (int)t;
        float val;
        if (i < kSW0) {
            int s = i / 256, n = i % 256;
            LayerDesc L = siren_layer(siren_step_layer(s));
            float b = n < L.out ? params[L.b_off + n] : 0.f;
            val = s == 7 ? b : 30.0f * b;
        } else if (i < kST0) {
            int k = (i - kSW0) / 256, n = (i - kSW0) % 256;
            val = params[siren_layer(0).w_off + n * 3 + k];
        } else if (i < kSWS) val = 30.0f * params[siren_layer(0).b_off + (i - kST0)];
        else if (i < kSWR) val = params[siren_layer(10).w_off + (i - kSWS)];
        else if (i < kSBH) val = params[siren_layer(11).w_off + (i - kSWR)];
        else if (i == kSBH) val = params[siren_layer(10).b_off];
        else val = params[siren_layer(11).b_off + (i - kSBH - 1)];
        tab[i] = val;
    }
}

// One SirenNeRF step's epilogue for this warp's half of the columns.
//   MODE 0: sin(30 acc + shift) -> bf16 h;   MODE 1: same + partial sigma head;   MODE 2: acc + shift -> bf16 h (linear);
//   MODE 3: N = 128 (64 columns per half): sin(30 acc + shift) -> partial rgb head, no store
template <int MODE>
__device__ __forceinline__ void siren_epi(uint32_t t_half, uint32_t sh_half, const float* __restrict__ head, uint32_t h_half,
                                          const uint32_t (&xoff)[8], float& sigma, float& rgb0, float& rgb1, float& rgb2) {
    constexpr int NJH = MODE == 3 ? 2 : 4;
#pragma unroll
    for (int jj = 0; jj < NJH; ++jj) {
        uint32_t v[32];
        tmem_ld32(t_half + (uint32_t)jj * 32u, v);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 sh = lds128(sh_half + (uint32_t)(jj * 32 + q * 4) * 4u);
            if (MODE == 2) {
                f[4 * q + 0] = __uint_as_float(v[4 * q + 0]) + sh.x; f[4 * q + 1] = __uint_as_float(v[4 * q + 1]) + sh.y;
                f[4 * q + 2] = __uint_as_float(v[4 * q + 2]) + sh.z; f[4 * q + 3] = __uint_as_float(v[4 * q + 3]) + sh.w;
            } else {
                f[4 * q + 0] = __sinf(fmaf(__uint_as_float(v[4 * q + 0]), 30.0f, sh.x));
                f[4 * q + 1] = __sinf(fmaf(__uint_as_float(v[4 * q + 1]), 30.0f, sh.y));
                f[4 * q + 2] = __sinf(fmaf(__uint_as_float(v[4 * q + 2]), 30.0f, sh.z));
                f[4 * q + 3] = __sinf(fmaf(__uint_as_float(v[4 * q + 3]), 30.0f, sh.w));
            }
        }
        if (MODE == 1) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                float4 w = __ldg(reinterpret_cast<const float4*>(head + jj * 32) + q);
                sigma = fmaf(f[4 * q + 0], w.x, fmaf(f[4 * q + 1], w.y, fmaf(f[4 * q + 2], w.z, fmaf(f[4 * q + 3], w.w, sigma))));
            }
        }
        if (MODE == 3) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                float4 w0 = __ldg(reinterpret_cast<const float4*>(head + jj * 32) + q);
                float4 w1 = __ldg(reinterpret_cast<const float4*>(head + 128 + jj * 32) + q);
                float4 w2 = __ldg(reinterpret_cast<const float4*>(head + 256 + jj * 32) + q);
                rgb0 = fmaf(f[4 * q + 0], w0.x, fmaf(f[4 * q + 1], w0.y, fmaf(f[4 * q + 2], w0.z, fmaf(f[4 * q + 3], w0.w, rgb0))));
                rgb1 = fmaf(f[4 * q + 0], w1.x, fmaf(f[4 * q + 1], w1.y, fmaf(f[4 * q + 2], w1.z, fmaf(f[4 * q + 3], w1.w, rgb1))));
                rgb2 = fmaf(f[4 * q + 0], w2.x, fmaf(f[4 * q + 1], w2.y, fmaf(f[4 * q + 2], w2.z, fmaf(f[4 * q + 3], w2.w, rgb2))));
            }
        } else {
            const uint32_t blk = h_half + (uint32_t)(jj >> 1) * 16384u;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                st_shared_v4(blk + xoff[(jj & 1) * 4 + q], pack_bf16(f[8 * q + 0], f[8 * q + 1]), pack_bf16(f[8 * q + 2], f[8 * q + 3]),
                             pack_bf16(f[8 * q + 4], f[8 * q + 5]), pack_bf16(f[8 * q + 6], f[8 * q + 7]));
        }
    }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
siren_tc_kernel(const uint8_t* __restrict__ packed, RowSource src, long long rows, float4* __restrict__ raw_out) {
    extern __shared__ uint8_t smem_raw[];
    const Ctx cx = make_ctx(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const PairLoop pl(rows);
    const float* __restrict__ tab = reinterpret_cast<const float*>(packed + kSirenChunkBytes);
    {   // shift[9][256] (9 KB) -> shared memory; input-layer / head weights stay in global memory (L1)
        const float4* tab_g = reinterpret_cast<const float4*>(tab);
        for (int i = threadIdx.x; i < kSW0 / 4; i += kThreads) {
            float4 v = __ldg(tab_g + i);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(cx.smem + kTabOff + 16u * i), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
        }
    }
    const uint32_t tmem_base = tc_prologue(cx, warp);

    if (warp == 0) {
        if (lane == 0) producer_loop<SirenSched>(cx, packed, pl, SirenSched::kSteps, 0);
    } else if (warp == 1) {
        if (cx.rank == 0) mma_loop<SirenSched>(cx, tmem_base, pl, SirenSched::kSteps, 0);
        else if (lane == 0) relay_loop<SirenSched>(cx, pl, SirenSched::kSteps, 0);
    } else if (warp >= kCtrlWarps) {
        const int ew = warp - kCtrlWarps;
        const int g = ew >> 3, half = (ew >> 2) & 1, quad = ew & 3;
        const int r = (quad << 5) | lane;
        const uint32_t sub = cx.smem + (uint32_t)g * kSubBytes;
        const uint32_t aux = sub, h_base = sub + kPeBytes;
        const uint32_t t_addr = tmem_base + ((uint32_t)quad << 21) + (uint32_t)g * 256u;
        const uint32_t row_off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
        const uint32_t xr = (uint32_t)(r & 7);
        const uint32_t part = aux + 8192u + (uint32_t)r * 16u;          // head partials live in the aux block (rows >= 64 of it are never an operand)
        const uint32_t bar_id = 1 + g;
        const uint32_t act_local = cx.act_ready + 8 * g, act_leader = mapa(act_local, 0);
        const uint32_t acc_bar = cx.acc_full + 8 * g;
        const uint32_t t_half = t_addr + (uint32_t)half * 128u;
        const uint32_t sh_half = cx.smem + kTabOff + (uint32_t)(half * 128) * 4u;
        const uint32_t h_half = h_base + row_off + (uint32_t)half * 2u * 16384u;
        uint32_t xoff[8];
#pragma unroll
        for (uint32_t c = 0; c < 8; ++c) xoff[c] = (c ^ xr) << 4;
        uint32_t acc_phase = 0;
        for (long long p = pl.first; p < pl.n_pairs; p += pl.stride) {
            const long long row = (2 * p + cx.rank) * kRowsTile + g * kRowsSub + r;
            const bool valid = row < rows;
            float pnt[3], vdir[3];
            load_row(src, valid ? row : rows - 1, pnt, vdir);
            // ---- layers_pos.0 on CUDA cores: this half produces columns half*128 .. +127 of h0
            for (int jj = 0; jj < 4; ++jj) {
                const int j = half * 4 + jj;
                uint32_t pk[16];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int n0 = j * 32 + q * 4;
                    float4 wx = __ldg(reinterpret_cast<const float4*>(tab + kSW0 + n0));
                    float4 wy = __ldg(reinterpret_cast<const float4*>(tab + kSW0 + 256 + n0));
                    float4 wz = __ldg(reinterpret_cast<const float4*>(tab + kSW0 + 512 + n0));
                    float4 sh = __ldg(reinterpret_cast<const float4*>(tab + kST0 + n0));
                    float a0 = fmaf(wz.x, pnt[2], fmaf(wy.x, pnt[1], wx.x * pnt[0]));
                    float a1 = fmaf(wz.y, pnt[2], fmaf(wy.y, pnt[1], wx.y * pnt[0]));
                    float a2 = fmaf(wz.z, pnt[2], fmaf(wy.z, pnt[1], wx.z * pnt[0]));
                    float a3 = fmaf(wz.w, pnt[2], fmaf(wy.w, pnt[1], wx.w * pnt[0]));
                    pk[2 * q + 0] = pack_bf16(__sinf(fmaf(a0, 30.0f, sh.x)), __sinf(fmaf(a1, 30.0f, sh.y)));
                    pk[2 * q + 1] = pack_bf16(__sinf(fmaf(a2, 30.0f, sh.z)), __sinf(fmaf(a3, 30.0f, sh.w)));
                }
                const uint32_t blk = h_base + (uint32_t)(j >> 1) * 16384u + row_off;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    st_shared_v4(blk + (((uint32_t)((j & 1) * 4 + q) ^ xr) << 4), pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
            }
            if (half == 0) {
                // raw position (3 values, zero-padded to 16) -> chunks 0..1 of the aux block (layers_pos.5's extra K)
                st_shared_v4(aux + row_off + ((0u ^ xr) << 4), pack_bf16(pnt[0], pnt[1]), pack_bf16(pnt[2], 0.f), 0u, 0u);
                st_shared_v4(aux + row_off + ((1u ^ xr) << 4), 0u, 0u, 0u, 0u);
            }
            arrive_act(act_local, act_leader, cx.rank, lane);

            float sigma = 0.f, rgb0 = 0.f, rgb1 = 0.f, rgb2 = 0.f;
            auto wait_acc = [&]() {
                mbar_wait_cluster(acc_bar, acc_phase);
                acc_phase ^= 1u;
                tc_fence_after();
            };
            for (int s = 0; s < 6; ++s) {                               // layers_pos.1 .. layers_pos.6
                wait_acc();
                siren_epi<0>(t_half, sh_half + (uint32_t)s * 1024u, nullptr, h_half, xoff, sigma, rgb0, rgb1, rgb2);
                arrive_act(act_local, act_leader, cx.rank, lane);
            }
            wait_acc();                                                 // layers_pos.7 (+ sigma head)
            siren_epi<1>(t_half, sh_half + 6u * 1024u, tab + kSWS + half * 128, h_half, xoff, sigma, rgb0, rgb1, rgb2);
            arrive_act(act_local, act_leader, cx.rank, lane);
            wait_acc();                                                 // layers_dir.0 (linear) + view direction into the aux block
            siren_epi<2>(t_half, sh_half + 7u * 1024u, nullptr, h_half, xoff, sigma, rgb0, rgb1, rgb2);
            if (half == 0) st_shared_v4(aux + row_off + ((0u ^ xr) << 4), pack_bf16(vdir[0], vdir[1]), pack_bf16(vdir[2], 0.f), 0u, 0u);
            arrive_act(act_local, act_leader, cx.rank, lane);
            wait_acc();                                                 // layers_dir.1 (N = 128) + rgb head
            siren_epi<3>(tmem_base + ((uint32_t)quad << 21) + (uint32_t)g * 256u + (uint32_t)half * 64u,
                         cx.smem + kTabOff + (uint32_t)(8 * 256 + half * 64) * 4u, tab + kSWR + half * 64, 0u, xoff, sigma, rgb0, rgb1, rgb2);
            tc_fence_before();
            if (half == 1)
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(part), "f"(rgb0), "f"(rgb1), "f"(rgb2), "f"(sigma) : "memory");
            asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
            if (half == 0) {
                float4 o2 = lds128(part);
                if (valid) {
                    float4 bh = __ldg(reinterpret_cast<const float4*>(tab + kSBH));     // (b_sigma, b_rgb[3])
                    float4 o;
                    o.x = 1.0f / (1.0f + __expf(-(rgb0 + o2.x + bh.y)));
                    o.y = 1.0f / (1.0f + __expf(-(rgb1 + o2.y + bh.z)));
                    o.z = 1.0f / (1.0f + __expf(-(rgb2 + o2.z + bh.w)));
                    o.w = fmaxf(sigma + o2.w + bh.x, 0.f);
                    raw_out[row] = o;
                }
            }
            asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
        }
    }
    tc_teardown(tmem_base, warp);
}

int pair_grid(long long rows, unsigned* grid);   // mlp_tc.cu

size_t siren_packed_bytes() { return (size_t)kSirenPackedBytes; }

int siren_pack(const float* params, void* packed_out, cudaStream_t st) {
    long long threads = kSirenChunkBytes / 16;
    siren_pack_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(params, (uint8_t*)packed_out);
    B2R_LAUNCH_CHECK("b2r_mlp_tc_pack (SirenNeRF)");
    return 0;
}

int siren_fwd(const void* packed, const b2r_mlp_input* in, long long rows, float* raw_out, cudaStream_t st) {
    unsigned grid = 0;
    int rc = pair_grid(rows, &grid);
    if (rc) return rc;
    rc = cuda_result(cudaFuncSetAttribute(siren_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes), "tc smem attribute");
    if (rc) return rc;
    siren_tc_kernel<<<grid, kThreads, kSmemBytes, st>>>((const uint8_t*)packed, make_row_source(in), rows, (float4*)raw_out);
    B2R_LAUNCH_CHECK("b2r_mlp_tc_fwd (SirenNeRF)");
    return 0;
}

}  // namespace tc
}  // namespace b2r
