// K3 (tensor-core path): the whole radiance-field MLP of a tile of samples evaluated inside ONE
// persistent kernel on the Blackwell 5th-generation tensor cores.
//   run_network + NeRF.forward            nerf/render.py:59-75, nerf/nerf.py:44-49, 75-94
//   FilmSirenNeRF.forward / create_mesh   pi_GAN/modules.py:22-25, 101-118, pi_GAN/utils.py:59-91
//
// Design (one CTA per SM, 384 threads, persistent over 256-row tiles = two 128-row sub-tiles):
//   warp 0        weight producer: streams pre-swizzled bf16 weight chunks (N x 32 K, SWIZZLE_64B image,
//                 packed once by b2r_mlp_tc_pack) L2 -> shared memory with cp.async.bulk (TMA engine)
//                 through a 4-stage mbarrier ring (4 x 16 KB);
//   warp 1        tcgen05.mma issuer (one elected lane): M=128, N=256|128, K=16 bf16 MMAs, A = the
//                 sub-tile's activations in shared memory (K-major, SWIZZLE_128B), D = fp32 accumulator
//                 in tensor memory (2 x 256 columns = the two sub-tiles, ping-pong);
//   warps 4-7     epilogue of sub-tile 0, warps 8-11 epilogue of sub-tile 1 (thread = row = TMEM lane):
//                 tcgen05.ld the accumulator, + bias, ReLU (or sin(s*acc+t) for FiLM-SIREN), round to
//                 bf16 and write the next layer's A operand back into shared memory IN PLACE -- the
//                 activations never touch HBM.  The first stage of the same threads generates the
//                 sample (o + d z), its positional encoding and the view-direction encoding; the last
//                 stage computes the narrow heads (sigma: 256 -> 1, rgb: 128|256 -> 3) on CUDA cores
//                 in fp32 and writes raw[row] = (sigmoid rgb, relu sigma) as one float4.
//   The MMA issuer alternates the two sub-tiles layer by layer, so the epilogue of one overlaps the
//   MMAs of the other.  Skip connections are extra K-chunks ([pe | h] for layers_pos.5,
//   [h | dir-enc] for layers_dir.1), zero-padded to the MMA K granularity in the packed weights.
//
// Algorithmic work: 1,182,976 FLOP per NeRF row; padded: 1,187,840 (+0.4 %).
#include "common.cuh"
#include "umma.cuh"

namespace b2r {
namespace tc {

using namespace umma;

constexpr int kThreads = 384;
constexpr int kRowsSub = 128;
constexpr int kRowsTile = 256;
constexpr int kStages = 4;
constexpr uint32_t kStageBytes = 16384;                 // 256 rows x 64 B
constexpr uint32_t kPeBytes = 16384;                    // 128 rows x 64 bf16 (SW128): pos-enc / dir-enc block
constexpr uint32_t kHBytes = 65536;                     // 4 K-blocks of 128 rows x 64 bf16
constexpr uint32_t kSubBytes = kPeBytes + kHBytes;      // 80 KB per sub-tile
constexpr uint32_t kRingOff = 2 * kSubBytes;
constexpr uint32_t kBarOff = kRingOff + kStages * kStageBytes;   // 224 KB
constexpr uint32_t kSmemBytes = kBarOff + 128 + 1024;            // + barriers + alignment slack

// ---- NeRF schedule: 10 MMA steps ------------------------------------------------------------------
// step 0..7 = layers_pos.0..7, 8 = layers_dir.0, 9 = layers_dir.1.  A chunk is N x 32 K.
constexpr int kNerfSteps = 10;
__host__ __device__ constexpr int nerf_chunks(int s) { return s == 0 ? 2 : (s == 5 ? 10 : (s == 9 ? 9 : 8)); }
__host__ __device__ constexpr int nerf_n(int s) { return s == 9 ? 128 : 256; }
// byte offset of the A operand of chunk c of step s inside the sub-tile region [pe | h]
__host__ __device__ constexpr uint32_t nerf_a_off(int s, int c) {
    if (s == 0) return (uint32_t)c * 64u;
    if (s == 5) { if (c < 2) return (uint32_t)c * 64u; c -= 2; }
    if (s == 9 && c == 8) return 0u;
    return kPeBytes + (uint32_t)(c >> 1) * 16384u + (uint32_t)(c & 1) * 64u;
}
__host__ __device__ constexpr long long nerf_chunk_off(int s, int c) {
    long long off = 0;
    for (int t = 0; t < s; ++t) off += (long long)nerf_chunks(t) * nerf_n(t) * 64;
    return off + (long long)c * nerf_n(s) * 64;
}
constexpr long long kNerfChunkBytes = nerf_chunk_off(kNerfSteps, 0);      // 1,187,840
// fp32 tables after the chunks: bias[10][256] | w_sigma[256] | w_rgb[3][128] | b_sigma, b_rgb[3]
constexpr int kNerfTabBias = 0, kNerfTabWSigma = 2560, kNerfTabWRgb = 2816, kNerfTabBHead = 3200, kNerfTabFloats = 3204;
constexpr long long kNerfPackedBytes = kNerfChunkBytes + kNerfTabFloats * 4;
static_assert(kNerfChunkBytes == 1187840, "NeRF packed chunk bytes");

// ---- pack kernel: fp32 state-dict parameters -> swizzled bf16 chunk images + fp32 tables -------------
__global__ void nerf_pack_kernel(const float* __restrict__ params, uint8_t* __restrict__ packed) {
    // one thread per (step, chunk, row, 16-byte group of 8 k)
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long total = kNerfChunkBytes / 16;
    if (t < total) {
        long long byte = t * 16;
        int s = 0;
        while (s + 1 < kNerfSteps && byte >= nerf_chunk_off(s + 1, 0)) ++s;
        long long in_step = byte - nerf_chunk_off(s, 0);
        int n_rows = nerf_n(s);
        int c = (int)(in_step / (n_rows * 64));
        int rem = (int)(in_step % (n_rows * 64));
        int row = rem / 64, grp = (rem % 64) / 16;        // logical (row, 16-byte chunk) of this thread
        int layer = s;                                     // nerf_layer index: 0..7 trunk, 8 dir.0, 9 dir.1
        LayerDesc L = nerf_layer(layer);
        __nv_bfloat16 v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            int kk = grp * 8 + e;                          // 0..31 inside the chunk
            int col = -1;
            if (s == 0) { int k = c * 32 + kk; col = k < 60 ? k : -1; }
            else if (s == 5) {
                if (c < 2) { int k = c * 32 + kk; col = k < 60 ? k : -1; }
                else col = 60 + (c - 2) * 32 + kk;
            } else if (s == 9) {
                if (c < 8) col = c * 32 + kk; else col = kk < 24 ? 256 + kk : -1;
            } else col = c * 32 + kk;
            float w = col >= 0 ? params[L.w_off + (long long)row * L.in + col] : 0.f;
            v[e] = __float2bfloat16_rn(w);
        }
        uint8_t* dst = packed + nerf_chunk_off(s, c) + sw64_offset((uint32_t)row, (uint32_t)grp);
        *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(v);
    }
    if (t < kNerfTabFloats) {
        float* tab = reinterpret_cast<float*>(packed + kNerfChunkBytes);
        int i = (int)t;
        float val = 0.f;
        if (i < kNerfTabWSigma) {
            int s = i / 256, n = i % 256;
            LayerDesc L = nerf_layer(s);
            val = n < L.out ? params[L.b_off + n] : 0.f;
        } else if (i < kNerfTabWRgb) val = params[nerf_layer(10).w_off + (i - kNerfTabWSigma)];
        else if (i < kNerfTabBHead) val = params[nerf_layer(11).w_off + (i - kNerfTabWRgb)];
        else if (i == kNerfTabBHead) val = params[nerf_layer(10).b_off];
        else val = params[nerf_layer(11).b_off + (i - kNerfTabBHead - 1)];
        tab[i] = val;
    }
}

// ---- fused kernel --------------------------------------------------------------------------------------
struct Ctx {
    uint32_t smem;        // 1024-aligned shared base (shared-window address)
    uint32_t w_full, w_empty, act_ready, acc_full, tmem_slot;
};

__device__ __forceinline__ Ctx make_ctx(uint8_t* raw) {
    Ctx c;
    c.smem = (smem_u32(raw) + 1023u) & ~1023u;
    c.w_full = c.smem + kBarOff;
    c.w_empty = c.w_full + 8 * kStages;
    c.act_ready = c.w_empty + 8 * kStages;
    c.acc_full = c.act_ready + 16;
    c.tmem_slot = c.acc_full + 16;
    return c;
}

// positional encoding of 3 values with L octaves into out[6L]: [sin(2^i x)(3), cos(2^i x)(3)] per octave.
// Octave 0 uses the accurate sincosf; higher octaves use the double-angle recurrence (abs. error grows
// ~2x per octave, < 1e-4 at octave 9: far below the bf16 rounding of the operand).
template <int L>
__device__ __forceinline__ void posenc_bf16(const float x[3], float* out) {
    float s[3], c[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) sincosf(x[k], &s[k], &c[k]);
#pragma unroll
    for (int i = 0; i < L; ++i) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { out[6 * i + k] = s[k]; out[6 * i + 3 + k] = c[k]; }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float s2 = 2.0f * s[k] * c[k];
            float c2 = 1.0f - 2.0f * s[k] * s[k];
            s[k] = s2; c[k] = c2;
        }
    }
}

// write `n16` 16-byte chunks (8 bf16 each) of row r of a SW128 K-block
template <int N16>
__device__ __forceinline__ void store_row_sw128(uint32_t block_base, int r, int first_chunk, const uint32_t* packed) {
#pragma unroll
    for (int q = 0; q < N16; ++q)
        st_shared_v4(block_base + sw128_offset((uint32_t)r, (uint32_t)(first_chunk + q)),
                     packed[4 * q + 0], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
}

__global__ void __launch_bounds__(kThreads, 1) nerf_tc_kernel(const uint8_t* __restrict__ packed, RowSource src,
                                                              long long rows, float4* __restrict__ raw_out) {
    extern __shared__ uint8_t smem_raw[];
    const Ctx cx = make_ctx(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long n_tiles = (rows + kRowsTile - 1) / kRowsTile;
    const float* __restrict__ tab = reinterpret_cast<const float*>(packed + kNerfChunkBytes);

    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(cx.w_full + 8 * i, 1); mbar_init(cx.w_empty + 8 * i, 1); }
        for (int g = 0; g < 2; ++g) { mbar_init(cx.act_ready + 8 * g, kRowsSub); mbar_init(cx.acc_full + 8 * g, 1); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(cx.tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(cx.tmem_slot));

    if (warp == 0) {
        // ===== weight producer =====
        uint32_t stage = 0, phase = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int s = 0; s < kNerfSteps; ++s) {
                const int nc = nerf_chunks(s);
                const uint32_t bytes = (uint32_t)nerf_n(s) * 64u;
                for (int g = 0; g < 2; ++g) {
                    for (int c = 0; c < nc; ++c) {
                        if (lane == 0) {
                            mbar_wait(cx.w_empty + 8 * stage, phase ^ 1u);
                            mbar_arrive_expect_tx(cx.w_full + 8 * stage, bytes);
                            bulk_g2s(cx.smem + kRingOff + stage * kStageBytes, packed + nerf_chunk_off(s, 0) + (long long)c * bytes,
                                     bytes, cx.w_full + 8 * stage);
                        }
                        __syncwarp();
                        if (++stage == kStages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        uint32_t stage = 0, phase = 0, act_phase[2] = {0u, 0u};
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int s = 0; s < kNerfSteps; ++s) {
                const int nc = nerf_chunks(s);
                const uint32_t idesc = make_idesc_bf16(128, (uint32_t)nerf_n(s));
                for (int g = 0; g < 2; ++g) {
                    if (lane == 0) {
                        mbar_wait(cx.act_ready + 8 * g, act_phase[g]);
                        tc_fence_after();
                    }
                    __syncwarp();
                    act_phase[g] ^= 1u;
                    const uint32_t d_tmem = tmem_base + (uint32_t)g * 256u;
                    const uint32_t a_base = cx.smem + (uint32_t)g * kSubBytes;
                    for (int c = 0; c < nc; ++c) {
                        if (lane == 0) {
                            mbar_wait(cx.w_full + 8 * stage, phase);
                            tc_fence_after();
                            const uint32_t a_addr = a_base + nerf_a_off(s, c);
                            const uint32_t b_addr = cx.smem + kRingOff + stage * kStageBytes;
#pragma unroll
                            for (int k = 0; k < 2; ++k)
                                mma_bf16(d_tmem, desc_sw128(a_addr + 32u * k), desc_sw64(b_addr + 32u * k), idesc, (uint32_t)((c | k) != 0));
                            mma_commit(cx.w_empty + 8 * stage);
                            if (c == nc - 1) mma_commit(cx.acc_full + 8 * g);
                        }
                        __syncwarp();
                        if (++stage == kStages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp >= 4) {
        // ===== input generation + epilogue; thread = row of sub-tile g = TMEM lane =====
        const int g = (warp - 4) >> 2;
        const int r = ((warp & 3) << 5) | lane;
        const uint32_t sub = cx.smem + (uint32_t)g * kSubBytes;
        const uint32_t pe_base = sub, h_base = sub + kPeBytes;
        const uint32_t t_addr = tmem_base + ((uint32_t)(warp & 3) << 21) + (uint32_t)g * 256u;   // lane (warp%4)*32 in bits [16,32)
        uint32_t acc_phase = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const long long row = tile * kRowsTile + g * kRowsSub + r;
            const bool valid = row < rows;
            float p[3], vdir[3];
            load_row(src, valid ? row : rows - 1, p, vdir);
            {
                // positional encoding: 60 values + 4 zero pads -> 64 bf16 = 8 chunks of the pe block
                float pe[64];
                posenc_bf16<10>(p, pe);
                pe[60] = pe[61] = pe[62] = pe[63] = 0.f;
                uint32_t pk[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) pk[i] = pack_bf16(pe[2 * i], pe[2 * i + 1]);
                store_row_sw128<8>(pe_base, r, 0, pk);
            }
            fence_proxy_async_smem();
            mbar_arrive(cx.act_ready + 8 * g);

            float sigma = 0.f, rgb0 = 0.f, rgb1 = 0.f, rgb2 = 0.f;
            for (int s = 0; s < kNerfSteps; ++s) {
                mbar_wait(cx.acc_full + 8 * g, acc_phase);
                acc_phase ^= 1u;
                tc_fence_after();
                const int nj = nerf_n(s) / 32;
                const float* __restrict__ bias = tab + kNerfTabBias + s * 256;
                for (int j = 0; j < nj; ++j) {
                    uint32_t v[32];
                    tmem_ld32(t_addr + (uint32_t)j * 32u, v);
                    tmem_ld_wait();
                    float f[32];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        float4 b = __ldg(reinterpret_cast<const float4*>(bias + j * 32) + q);
                        f[4 * q + 0] = __uint_as_float(v[4 * q + 0]) + b.x;
                        f[4 * q + 1] = __uint_as_float(v[4 * q + 1]) + b.y;
                        f[4 * q + 2] = __uint_as_float(v[4 * q + 2]) + b.z;
                        f[4 * q + 3] = __uint_as_float(v[4 * q + 3]) + b.w;
                    }
                    if (s == 7) {
                        // sigma head on the fp32 activations (output_layer_sigma, nerf/nerf.py:72,88)
                        const float* __restrict__ ws = tab + kNerfTabWSigma + j * 32;
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            float4 w = __ldg(reinterpret_cast<const float4*>(ws) + q);
                            sigma = fmaf(fmaxf(f[4 * q + 0], 0.f), w.x, sigma);
                            sigma = fmaf(fmaxf(f[4 * q + 1], 0.f), w.y, sigma);
                            sigma = fmaf(fmaxf(f[4 * q + 2], 0.f), w.z, sigma);
                            sigma = fmaf(fmaxf(f[4 * q + 3], 0.f), w.w, sigma);
                        }
                    }
                    if (s == 9) {
                        // rgb head (output_layer_rgb: 128 -> 3) on the fp32 relu activations
#pragma unroll
                        for (int e = 0; e < 32; ++e) {
                            float h = fmaxf(f[e], 0.f);
                            int k = j * 32 + e;
                            rgb0 = fmaf(h, __ldg(tab + kNerfTabWRgb + k), rgb0);
                            rgb1 = fmaf(h, __ldg(tab + kNerfTabWRgb + 128 + k), rgb1);
                            rgb2 = fmaf(h, __ldg(tab + kNerfTabWRgb + 256 + k), rgb2);
                        }
                    } else {
                        uint32_t pk[16];
                        if (s == 8) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) pk[i] = pack_bf16(f[2 * i], f[2 * i + 1]);        // layers_dir.0 is linear
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i) pk[i] = pack_bf16_relu(f[2 * i], f[2 * i + 1]);
                        }
                        // columns j*32 .. j*32+31 of this layer = K of the next: K-block j/2, chunks (j&1)*4 .. +3
                        store_row_sw128<4>(h_base + (uint32_t)(j >> 1) * 16384u, r, (j & 1) * 4, pk);
                    }
                }
                if (s == 8) {
                    // view-direction encoding for layers_dir.1: 24 values + 8 zero pads -> chunks 0..3 of the pe block
                    float de[32];
                    posenc_bf16<4>(vdir, de);
#pragma unroll
                    for (int i = 24; i < 32; ++i) de[i] = 0.f;
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) pk[i] = pack_bf16(de[2 * i], de[2 * i + 1]);
                    store_row_sw128<4>(pe_base, r, 0, pk);
                }
                tc_fence_before();
                if (s < kNerfSteps - 1) {
                    fence_proxy_async_smem();
                    mbar_arrive(cx.act_ready + 8 * g);
                }
            }
            if (valid) {
                float bs = __ldg(tab + kNerfTabBHead);
                float b0 = __ldg(tab + kNerfTabBHead + 1), b1 = __ldg(tab + kNerfTabBHead + 2), b2 = __ldg(tab + kNerfTabBHead + 3);
                float4 o;
                o.x = 1.0f / (1.0f + __expf(-(rgb0 + b0)));
                o.y = 1.0f / (1.0f + __expf(-(rgb1 + b1)));
                o.z = 1.0f / (1.0f + __expf(-(rgb2 + b2)));
                o.w = fmaxf(sigma + bs, 0.f);
                raw_out[row] = o;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace tc
}  // namespace b2r

extern "C" size_t b2r_mlp_tc_packed_bytes(int model_kind) {
    if (model_kind == B2R_MODEL_NERF) return (size_t)b2r::tc::kNerfPackedBytes;
    return 0;
}

extern "C" int b2r_mlp_tc_pack(int model_kind, const float* params, const float* film, int use_dir,
                               void* packed_out, void* stream) {
    using namespace b2r;
    (void)film; (void)use_dir;
    B2R_CHECK_ARG(params && packed_out, "b2r_mlp_tc_pack: NULL pointer");
    B2R_CHECK_ARG(((uintptr_t)packed_out & 15) == 0, "b2r_mlp_tc_pack: packed_out must be 16-byte aligned");
    if (model_kind == B2R_MODEL_NERF) {
        long long threads = tc::kNerfChunkBytes / 16;
        tc::nerf_pack_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(params, (uint8_t*)packed_out);
        B2R_LAUNCH_CHECK("b2r_mlp_tc_pack");
        return 0;
    }
    return fail(-2, "b2r_mlp_tc_pack: model kind %d has no tensor-core path in this build", model_kind);
}

extern "C" int b2r_mlp_tc_fwd(int model_kind, const void* packed, const b2r_mlp_input* in, float* raw_out,
                              int sigma_only, void* stream) {
    using namespace b2r;
    B2R_CHECK_ARG(packed && raw_out, "b2r_mlp_tc_fwd: NULL pointer");
    B2R_CHECK_ARG((((uintptr_t)packed | (uintptr_t)raw_out) & 15) == 0, "b2r_mlp_tc_fwd: packed / raw_out must be 16-byte aligned");
    int rc = check_mlp_input(in);
    if (rc) return rc;
    long long rows = row_count(in);
    if (rows == 0) return 0;
    if (model_kind != B2R_MODEL_NERF || sigma_only)
        return fail(-2, "b2r_mlp_tc_fwd: model kind %d (sigma_only=%d) has no tensor-core path in this build", model_kind, sigma_only);
    int dev = 0, sms = 0;
    rc = cuda_result(cudaGetDevice(&dev), "cudaGetDevice");
    if (rc) return rc;
    rc = cuda_result(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "SM count");
    if (rc) return rc;
    rc = cuda_result(cudaFuncSetAttribute(tc::nerf_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kSmemBytes), "tc smem attribute");
    if (rc) return rc;
    long long n_tiles = (rows + tc::kRowsTile - 1) / tc::kRowsTile;
    unsigned grid = (unsigned)(n_tiles < sms ? n_tiles : sms);
    tc::nerf_tc_kernel<<<grid, tc::kThreads, tc::kSmemBytes, (cudaStream_t)stream>>>((const uint8_t*)packed, make_row_source(in), rows,
                                                                                  (float4*)raw_out);
    B2R_LAUNCH_CHECK("b2r_mlp_tc_fwd");
    return 0;
}
