// Shared device-side machinery of the fused tensor-core MLP kernels (mlp_tc.cu: inference + training forward,
// mlp_tc_train.cu: reverse mode): schedules, packed-weight layout, shared-memory map, mbarrier protocol, the weight
// producer / relay / MMA-issuer loops of a CTA pair.  See mlp_tc.cu for the design.
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace b2r {
namespace tc {

using namespace umma;

constexpr int kRowsSub = 128;
constexpr int kRowsTile = 256;                          // rows per CTA per iteration (a pair covers 512)
constexpr int kStages = 3;
constexpr uint32_t kStageBytes = 16384;                 // 128 weight rows x 64 K bf16 (this CTA's N-half)
constexpr uint32_t kPeBytes = 16384;                    // 128 rows x 64 bf16 (SW128): pos-enc / dir-enc block
constexpr uint32_t kHBytes = 65536;                     // 4 K-blocks of 128 rows x 64 bf16
constexpr uint32_t kSubBytes = kPeBytes + kHBytes;      // 80 KB per sub-tile
constexpr uint32_t kRingOff = 2 * kSubBytes;

constexpr int kCtrlWarps = 4;                           // 0 producer, 1 MMA issuer / relay (+TMEM alloc), 2-3 idle
constexpr int kEpiWarps = 16;                           // 2 sub-tiles x 2 column halves x 4 TMEM lane quadrants
constexpr int kThreads = (kCtrlWarps + kEpiWarps) * 32; // 640

// ---- schedules (chunks of 64 K) ---------------------------------------------------------------------------
// A step = one layer's MMAs for one sub-tile: n_pre chunks whose A operand is the aux block (pos-enc),
// then n_h chunks reading the K-blocks of h, then n_post chunks reading the aux block again (dir-enc /
// view direction; only kPostMmas x 16 K of it are issued).
struct NerfSched {       // step 0..7 = layers_pos.0..7, 8 = layers_dir.0, 9 = layers_dir.1
    static constexpr int kSteps = 10;
    __host__ __device__ static constexpr int n_pre(int s, int) { return (s == 0 || s == 5) ? 1 : 0; }
    __host__ __device__ static constexpr int n_h(int s, int) { return s == 0 ? 0 : 4; }
    __host__ __device__ static constexpr int n_post(int s, int) { return s == 9 ? 1 : 0; }
    __host__ __device__ static constexpr int n(int s) { return s == 9 ? 128 : 256; }
    static constexpr int kPostMmas = 2;                  // dir-enc: 24 -> 32 K
};
// FiLM-SIREN: steps 0..6 = hidden_layers.0..6, step 7 = hidden_layer_rgb ([h | dir]).  Every step has a 16-K "post" chunk read
// from the aux block [dir(3), 1, 1, 0...]: its weights carry the view-direction columns (step 7) and the FiLM SHIFT of the
// layer as two bf16 terms (hi + lo) against the two constant ones, and the weight rows are pre-multiplied by the FiLM SCALE
// (30 gamma) -- so the accumulator IS the sine's argument and the epilogue needs no table at all.
struct FilmSched {
    static constexpr int kSteps = 8;
    __host__ __device__ static constexpr int n_pre(int, int) { return 0; }
    __host__ __device__ static constexpr int n_h(int, int) { return 4; }
    __host__ __device__ static constexpr int n_post(int, int) { return 1; }
    __host__ __device__ static constexpr int n(int) { return 256; }
    static constexpr int kPostMmas = 1;                  // [dir(3), 1, 1] -> 16 K
};
template <class S>
__host__ __device__ constexpr uint32_t half_bytes(int s) { return (uint32_t)(S::n(s) / 2) * 128u; }
// packed chunks of a step (the post chunk is always stored, even when use_dir = 0 skips it)
template <class S>
__host__ __device__ constexpr int stored_chunks(int s) { return S::n_pre(s, 1) + S::n_h(s, 1) + S::n_post(s, 1); }
template <class S>
__host__ __device__ constexpr long long step_base(int s) {
    long long off = 0;
    for (int t = 0; t < s; ++t) off += 2LL * stored_chunks<S>(t) * half_bytes<S>(t);
    return off;
}

// ---- compact copy of every step's LAST h chunk + post chunk (sine models, inference kernels; MapC below) ----------------------
// A post chunk carries 16 K columns of which a SWIZZLE_128B stage image (16 KB, one bulk copy, one ring-stage cycle) uses 32 bytes
// per row.  The inference kernels of the sine models read, per step and N-half, ONE blob [h chunk 3 (SWIZZLE_128B image) | post chunk
// in the NO-SWIZZLE core-matrix layout [16-byte K chunk (2)][weight row][16 B]] with one bulk copy: 4 copies per step and sub-tile
// instead of 5.  The blobs live in an extra area behind the fp32 tables of the packed image (the training kernels keep reading the
// chunk area, which is unchanged).
template <class S>
__host__ __device__ constexpr uint32_t post_bytes(int s) { return (uint32_t)(S::n(s) / 2) * 32u; }
template <class S>
__host__ __device__ constexpr long long extra_base(int s) {
    long long off = 0;
    for (int t = 0; t < s; ++t) off += 2LL * (half_bytes<S>(t) + post_bytes<S>(t));
    return off;
}
// which (step, half, part) a 16-byte group of the extra area belongs to: c = 3 (h chunk 3: row, grp 0..7 as in `locate`) or
// c = 4 (post chunk: grp = 16-byte K chunk 0 / 1); dst = byte offset inside the extra area
template <class S>
__device__ __forceinline__ void locate_extra(long long byte, int& s, int& c, int& hf, int& row, int& grp) {
    s = 0;
    while (s + 1 < S::kSteps && byte >= extra_base<S>(s + 1)) ++s;
    const int in_step = (int)(byte - extra_base<S>(s));
    const int hb = (int)half_bytes<S>(s), pb = (int)post_bytes<S>(s);
    hf = in_step / (hb + pb);
    const int rem = in_step % (hb + pb);
    if (rem < hb) {                      // SWIZZLE_128B image: the thread's group is (row, grp) -> written at sw128_offset
        c = 3; row = rem / 128; grp = (rem % 128) / 16;
    } else {
        c = 4; grp = (rem - hb) / (pb / 2); row = ((rem - hb) % (pb / 2)) / 16;
    }
}
template <class S>
__device__ __forceinline__ long long extra_dst(int s, int c, int hf, int row, int grp) {
    const long long b = extra_base<S>(s) + (long long)hf * (half_bytes<S>(s) + post_bytes<S>(s));
    return c == 3 ? b + sw128_offset((uint32_t)row, (uint32_t)grp) : b + half_bytes<S>(s) + (long long)grp * (post_bytes<S>(s) / 2) + row * 16;
}

constexpr long long kNerfChunkBytes = step_base<NerfSched>(NerfSched::kSteps);      // 1,196,032
constexpr long long kFilmChunkBytes = step_base<FilmSched>(FilmSched::kSteps);      // 1,310,720
static_assert(kNerfChunkBytes == 1196032 && kFilmChunkBytes == 8 * 5 * 32768, "packed chunk bytes");
// NeRF fp32 tables after the chunks: bias[10][256] | w_sigma[256] | w_rgb[3][128] | b_sigma, b_rgb[3]
constexpr int kNerfTabBias = 0, kNerfTabWSigma = 2560, kNerfTabWRgb = 2816, kNerfTabBHead = 3200, kNerfTabFloats = 3204;
constexpr long long kNerfPackedBytes = kNerfChunkBytes + kNerfTabFloats * 4;
// FiLM fp32 tables (all staged in shared memory): w0[3][256] (input layer, column-major) | scale0[256] | shift0[256] |
//                   w_sigma[256] | w_rgb[3][256] | b_sigma, b_rgb[3]
constexpr int kFW0 = 0, kFS0 = 768, kFT0 = 1024, kFWS = 1280, kFWR = 1536, kFBH = 2304, kFilmTabFloats = 2308;
constexpr long long kFilmExtraOff = kFilmChunkBytes + kFilmTabFloats * 4;               // compact blobs (extra_base) behind the tables
constexpr long long kFilmPackedBytes = kFilmExtraOff + extra_base<FilmSched>(FilmSched::kSteps);
static_assert(kFilmExtraOff % 16 == 0 && kFilmPackedBytes % 16 == 0, "bulk copies need 16-byte alignment");

// ---- packed weights: which (step, chunk, half, row, 16-byte group) a byte of the chunk area belongs to ----
// one thread per 16-byte group: (step, chunk, half, row of the half, 8 consecutive k)
template <class S>
__device__ __forceinline__ void locate(long long byte, int& s, int& c, int& hf, int& row, int& grp) {
    s = 0;
    while (s + 1 < S::kSteps && byte >= step_base<S>(s + 1)) ++s;
    long long in_step = byte - step_base<S>(s);
    const int hb = (int)half_bytes<S>(s);
    int hc = (int)(in_step / hb);
    c = hc >> 1; hf = hc & 1;
    int rem = (int)(in_step % hb);
    row = rem / 128; grp = (rem % 128) / 16;
}


// ---- tiled activation tensors kept for the reverse mode (training) ------------------------------------------------
// A tensor [rows x C] bf16 is stored as blocks of [128 rows x 64 columns]: exactly the K-major SWIZZLE_128B shared-memory
// image of that block (16 KB; 16-byte chunk c of row r at c ^ (r & 7)).  The forward / dgrad kernels write a block with
// the same per-thread offsets they use for shared memory, and the wgrad kernel streams half blocks (64 rows, 8 KB
// contiguous) back with the bulk-copy engine and feeds them to tcgen05 as MN-major operands (row index = K) with no
// transposition anywhere.  Tensors are stored one after the other (tensor-major); block (T, kb) of a tensor with nb
// blocks per 128-row tile T sits at tensor_base + (T * nb + kb) * 16 KB.
constexpr uint32_t kBlk = 16384;
// forward ("saved"): PE | H0..H7 | GL (layers_dir.0 output) | DE | HD   -- block offsets in units of n_sub blocks
constexpr int kSavPE = 0, kSavH0 = 1, kSavGL = 33, kSavDE = 37, kSavHD = 38, kSavBlocks = 40;
// reverse ("scratch"): GD1 (d pre-activation of layers_dir.1) | GG (d layers_dir.0 output) | GH0..GH7 (d pre-activations of the
// trunk), then the head gradients HG[row] = (d rgb pre-sigmoid x3, d sigma pre-relu) as float4
constexpr int kScrGD1 = 0, kScrGG = 2, kScrGH0 = 6, kScrBlocks = 38;
__host__ __device__ constexpr int sav_h(int l) { return kSavH0 + 4 * l; }
__host__ __device__ constexpr int scr_gh(int l) { return kScrGH0 + 4 * l; }
// 128-row sub-tiles covered by the CTA pairs (every pair always processes 4 sub-tiles)
__host__ __device__ inline long long n_sub_tiles(long long rows) {
    long long n_tiles = (rows + kRowsTile - 1) / kRowsTile;
    return ((n_tiles + 1) / 2) * 4;
}
// relu masks (training): one bit per activation, written by the forward epilogue thread (= row) as one 16-byte word per
// layer and column half: mask tensor [8 layers][n_sub][2 halves][128 rows] uint4 after the blocks, then the h_d mask
// [n_sub][2][128] uint2.  Bit i of word jj <-> column 32 jj + 2 i of the half, bit 16 + i <-> column 32 jj + 2 i + 1.
__host__ __device__ constexpr size_t saved_bytes_per_sub() { return (size_t)kSavBlocks * kBlk + 8 * 2 * 128 * 16 + 2 * 128 * 8; }
__device__ __forceinline__ size_t mask_off(size_t n_sub, int layer, size_t T, int half, int r) {
    return (size_t)kSavBlocks * n_sub * kBlk + ((((size_t)layer * n_sub + T) * 2 + half) * 128 + r) * 16;
}
__device__ __forceinline__ size_t hdmask_off(size_t n_sub, size_t T, int half, int r) {
    return (size_t)kSavBlocks * n_sub * kBlk + (size_t)8 * n_sub * 2 * 128 * 16 + ((T * 2 + half) * 128 + r) * 8;
}
// 0xFFFF in every half whose bf16 value is > 0
__device__ __forceinline__ uint32_t relu_mask2(uint32_t h2) {
    __nv_bfloat162 a, z;
    *reinterpret_cast<uint32_t*>(&a) = h2;
    *reinterpret_cast<uint32_t*>(&z) = 0u;
    return __hgt2_mask(a, z);
}
// gather the two relu bits of packed word i of a 32-column group / expand them back to a bf16x2 AND-mask
__device__ __forceinline__ void mask_put(uint32_t& bits, uint32_t w, int i) { bits |= relu_mask2(w) & (0x00010001u << i); }
__device__ __forceinline__ uint32_t mask_get(uint32_t bits, int i) { return ((bits >> i) & 0x00010001u) * 0xFFFFu; }

// cos(t) checkpoints of the sine layers (training): ONE BYTE per activation, offset-binary fixed point
//   ub = clamp(round(128 cos t), -128, 127) + 128        (absolute error <= 2^-8, the same as a bf16 near |cos| = 1)
// stored THREAD-MAJOR -- [layer][T][quarter cq of 64 columns][word w 0..3][row r] uint4, word w = columns 16 w .. 16 w + 15 of the
// quarter -- so the 32 lanes of a warp touch 512 contiguous bytes and the forward's stores / the reverse mode's loads are fully
// coalesced (the trick that made the relu bits cheap), at 256 B per row and layer.
// four sine arguments -> four packed bytes (t0 in the low byte).  1.5 * 2^23 + 128 + 128 c rounds to an integer whose low
// mantissa byte is ub; cos is clamped below 1 so that 128 c + 128 never reaches 256.
__device__ __forceinline__ uint32_t cos_q4(float t0, float t1, float t2, float t3) {
    constexpr float kMagic = 12582912.0f + 128.0f, kTop = 0.9921875f;
    const uint32_t a = __float_as_uint(fmaf(fminf(__cosf(t0), kTop), 128.0f, kMagic)), b = __float_as_uint(fmaf(fminf(__cosf(t1), kTop), 128.0f, kMagic));
    const uint32_t c = __float_as_uint(fmaf(fminf(__cosf(t2), kTop), 128.0f, kMagic)), d = __float_as_uint(fmaf(fminf(__cosf(t3), kTop), 128.0f, kMagic));
    return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
}
// byte K of a packed word -> cos: the byte becomes mantissa bits 8..15 of a float with exponent 2^8 (256 + ub / 128), minus 257
template <int K>
__device__ __forceinline__ float cos_unq(uint32_t word) {
    return __uint_as_float(__byte_perm(word, 0x43800000u, 0x7604u | (K << 4))) - 257.0f;
}
// f[0..7] *= cos of the 8 bytes in (lo, hi)
__device__ __forceinline__ void cos_mul8(float* f, uint32_t lo, uint32_t hi) {
    f[0] *= cos_unq<0>(lo); f[1] *= cos_unq<1>(lo); f[2] *= cos_unq<2>(lo); f[3] *= cos_unq<3>(lo);
    f[4] *= cos_unq<0>(hi); f[5] *= cos_unq<1>(hi); f[6] *= cos_unq<2>(hi); f[7] *= cos_unq<3>(hi);
}
constexpr size_t kCosLayerBytes = 32768;                 // per 128-row sub-tile and 256-column layer: 4 quarters x 4 words x 2 KB

// SirenNeRF training checkpoints: tiles AUX [pos(3), 1, 1, dir(3)] | H0..H7 | GL | HD, then the cosine bytes of layers_pos.0..7
// and of layers_dir.1 (128 columns: 4 "quarters" of 32 columns = 2 words each)
constexpr int kSsAUX = 0, kSsH0 = 1, kSsGL = 33, kSsHD = 37, kSsBlocks = 39;
__host__ __device__ constexpr int ss_h(int l) { return kSsH0 + 4 * l; }
__host__ __device__ constexpr size_t siren_saved_bytes_per_sub() { return (size_t)kSsBlocks * kBlk + 8 * kCosLayerBytes + kCosLayerBytes / 2; }
__device__ __forceinline__ size_t siren_cos_off(size_t n_sub, int layer, size_t T, int cq, int w, int r) {
    return (size_t)kSsBlocks * n_sub * kBlk + ((((size_t)layer * n_sub + T) * 4 + cq) * 4 + w) * 2048 + (size_t)r * 16;
}
__device__ __forceinline__ size_t siren_cos9_off(size_t n_sub, size_t T, int cq, int w, int r) {
    return (size_t)kSsBlocks * n_sub * kBlk + (size_t)8 * n_sub * kCosLayerBytes + (((T * 4 + cq) * 2 + w) * 2048) + (size_t)r * 16;
}

// FiLM-SIREN training checkpoints: tiles AUX [dir(3), 1, 1, pos(3)] | H0..H7 (inputs of hidden_layers.0..6 and hidden_layer_rgb) |
// HC (hidden_layer_rgb output), then the cosine bytes of the nine sine layers
constexpr int kFsAUX = 0, kFsH0 = 1, kFsHC = 33, kFsBlocks = 37;
__host__ __device__ constexpr int fs_h(int l) { return kFsH0 + 4 * l; }
__host__ __device__ constexpr size_t film_saved_bytes_per_sub() { return (size_t)kFsBlocks * kBlk + 9 * kCosLayerBytes; }
__device__ __forceinline__ size_t film_cos_off(size_t n_sub, int layer, size_t T, int cq, int w, int r) {
    return (size_t)kFsBlocks * n_sub * kBlk + ((((size_t)layer * n_sub + T) * 4 + cq) * 4 + w) * 2048 + (size_t)r * 16;
}
// reverse-mode scratch of the FiLM path: G8 (d t of hidden_layer_rgb) | G7 .. G0, then the head gradients HG
constexpr int kFScrBlocks = 36;
__host__ __device__ constexpr int fscr_g(int l) { return 4 * (8 - l); }        // G8 first (written first), G0 last

__device__ __forceinline__ void stg128(uint8_t* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ldg128(const uint8_t* p) {
    uint4 v;
    asm volatile("ld.global.nc.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// Tile copy shared -> global of the training kernels' spill threads: n_blk blocks of 16 KB as PACED 16 KB bulk copies, two in flight.
// The SM's bulk-copy engine serves one operation at a time (tools/native/wstream_probe.cu, tma_mix_probe.cu): a single 64 KB copy
// holds it for ~2,000 clk, during which the weight chunks of the running MMAs (one 16 KB copy per 512 clk) queue behind it.  With
// 16 KB pieces they interleave: training forward of the fine pass 1.08 -> 1.02 ms (8 KB pieces: 1.04 ms, more engine time per byte).
// The caller still ends the tile with bulk_commit() + bulk_wait_read().
__device__ __forceinline__ void spill_tile(uint8_t* dst, uint32_t src_smem, int n_blk) {
    for (int q = 0; q < n_blk; ++q) {
        bulk_s2g(dst + (size_t)q * kBlk, src_smem + (uint32_t)q * kBlk, kBlk);
        bulk_commit();
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    }
}

// ---- shared-memory map (offsets from the 1024-aligned base; identical in both CTAs of a pair) -------------------
constexpr uint32_t kTabOff = kRingOff + kStages * kStageBytes;          // fp32 tables (NeRF bias / head weights)
constexpr uint32_t kTabBytes = kNerfTabFloats * 4;                      // 12,816
constexpr uint32_t kPartOff = kTabOff + kTabBytes;                      // head partial sums: 2 x 128 x float4
constexpr uint32_t kPartBytes = 2 * kRowsSub * 16;
constexpr uint32_t kPart2Off = kPartOff + kPartBytes;                   // NeRF last-sample check: the other column half's sum |w h|
constexpr uint32_t kPart2Bytes = 2 * kRowsSub * 4;
constexpr uint32_t kBarOff = kPart2Off + kPart2Bytes;                   // mbarriers + TMEM slot
constexpr uint32_t kSmemBytes = kBarOff + 128 + 1024;                   // + alignment slack
static_assert(kBarOff % 8 == 0 && kSmemBytes <= 232448, "shared-memory budget");

struct Ctx {
    uint32_t smem;        // 1024-aligned shared base (shared-window address)
    uint32_t w_full, w_empty, act_ready, acc_full, tmem_slot, spill_ready, spill_done;
    uint32_t rank;        // CTA rank in the pair (0 = leader: issues the MMAs)
};

__device__ __forceinline__ Ctx make_ctx(uint8_t* raw) {
    Ctx c;
    c.smem = (smem_u32(raw) + 1023u) & ~1023u;
    c.w_full = c.smem + kBarOff;
    c.w_empty = c.w_full + 8 * kStages;
    c.act_ready = c.w_empty + 8 * kStages;
    c.acc_full = c.act_ready + 16;
    c.tmem_slot = c.acc_full + 16;
    c.spill_ready = c.tmem_slot + 16;      // training kernels: tile written -> spill thread (count 8: the sub-tile's epilogue warps)
    c.spill_done = c.spill_ready + 16;     // spill thread -> epilogue warps: the bulk store has read the tile (count 1)
    c.rank = cluster_ctarank();
    return c;
}

// ---- compact shared-memory map of the sine models' INFERENCE kernels (film_tc_kernel<false>, siren_tc_kernel<false>) ----------
// Their aux operand is ONE 16-K chunk, so it is kept in the no-swizzle core-matrix layout [16-byte K chunk (2)][row (128)][16 B]
// (4 KB instead of a 16 KB SWIZZLE_128B block; chunk 1 is all zero), which frees 24 KB: the weight ring gets a FOURTH stage, and
// stage 3 has room for the blob [h chunk 3 | post chunk] (post_bytes above).  Every step has exactly 4 copies per sub-tile, so chunk
// c always lands in stage c.  Why it matters (profiles/r2_role_timers.txt): a stage cycle (wait for w_empty, bulk copy,
// relay hop, MMAs, commit) is ~1,480 clk whatever the copy's size, three stages delivered one chunk per ~493 clk, and a sine-model
// step asked for 5 chunks per 2,176 clk of MMA time (435 clk per chunk): the ring, not the tensor or MUFU pipe, set the pace.
struct MapC {
    static constexpr int kStagesC = 4;
    static constexpr uint32_t kAux = 4096;
    static constexpr uint32_t kSub = kAux + kHBytes;                              // 68 KB per sub-tile
    static constexpr uint32_t kRing = 2 * kSub;
    static constexpr uint32_t kRingBytes = kStagesC * kStageBytes + 4096;         // stage 3: + the largest post chunk
    static constexpr uint32_t kTab = kRing + kRingBytes;                          // fp32 tables (region of kTabBytes + kPartBytes)
    static constexpr uint32_t kBar = kTab + kTabBytes + kPartBytes;
    static constexpr uint32_t kSmem = kBar + 256 + 1024;
    static constexpr uint32_t kAuxLbo = 2048, kAuxSbo = 128;                      // K-direction / 8-row-group pitch (tests/native/umma_kmajor_noswizzle_probe.cu)
};
static_assert(MapC::kBar % 8 == 0 && MapC::kSmem <= 232448 && MapC::kSub % 1024 == 0 && MapC::kAux % 1024 == 0, "compact shared-memory map");

__device__ __forceinline__ Ctx make_ctx_c(uint8_t* raw) {
    Ctx c;
    c.smem = (smem_u32(raw) + 1023u) & ~1023u;
    c.w_full = c.smem + MapC::kBar;
    c.w_empty = c.w_full + 8 * MapC::kStagesC;
    c.act_ready = c.w_empty + 8 * MapC::kStagesC;
    c.acc_full = c.act_ready + 16;
    c.tmem_slot = c.acc_full + 16;
    c.spill_ready = c.tmem_slot + 16;      // (unused: the compact kernels keep nothing for a reverse mode)
    c.spill_done = c.spill_ready + 16;
    c.rank = cluster_ctarank();
    return c;
}

// ---- last-sample sign check (include/b2r.h: b2r_last_sample) -----------------------------------------------------------
// The kernels list every ray whose LAST sample's pre-relu sigma lies inside the bf16 error band; the host re-evaluates those
// rows in fp32 (b2r_mlp_f32_last_sigma).  count == nullptr: off.
struct LastFlag {
    int* count;
    int* ray_ids;
    int capacity, s;
    float rel, abs;
};
inline LastFlag make_last_flag(const b2r_last_sample* l) {
    LastFlag f{nullptr, nullptr, 0, 1, 0.f, 0.f};
    if (l) { f.count = l->count; f.ray_ids = l->ray_ids; f.capacity = l->capacity; f.s = l->samples_per_ray; f.rel = l->rel; f.abs = l->abs; }
    return f;
}
// is `row` the last sample of its ray?  ray = load_row's ray_out (rays mode: row / n_samples == row / lf.s)
__device__ __forceinline__ bool last_of_ray(const LastFlag& lf, const RowSource& src, long long row, long long& ray) {
    if (!lf.count) return false;
    if (src.rays) return row - ray * src.n_samples == src.n_samples - 1;
    ray = row / lf.s;
    return row - ray * lf.s == lf.s - 1;
}
__device__ __forceinline__ void flag_ray(const LastFlag& lf, long long ray) {
    const int slot = atomicAdd(lf.count, 1);
    if (slot < lf.capacity) lf.ray_ids[slot] = (int)ray;
}

__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// Positional encoding of 3 values with L octaves as 3L packed bf16x2 words, in the reference's order
// [sin(2^i x)(3), cos(2^i x)(3)] per octave (nerf/nerf.py:44-49): 6 values = 3 words per octave.
// Octave 0 uses the accurate sincosf; higher octaves the double-angle recurrence (abs. error grows ~2x per
// octave, < 1e-4 at octave 9: far below the bf16 rounding of the operand).
template <int L>
__device__ __forceinline__ void posenc_words(const float x[3], uint32_t* w) {
    float s[3], c[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) sincosf(x[k], &s[k], &c[k]);
#pragma unroll
    for (int i = 0; i < L; ++i) {
        w[3 * i + 0] = pack_bf16(s[0], s[1]);
        w[3 * i + 1] = pack_bf16(s[2], c[0]);
        w[3 * i + 2] = pack_bf16(c[1], c[2]);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float s2 = 2.0f * s[k] * c[k];
            float c2 = 1.0f - 2.0f * s[k] * s[k];
            s[k] = s2; c[k] = c2;
        }
    }
}

// tile pairs walked by the cluster: pair p -> tiles 2p (leader) and 2p+1 (peer)
struct PairLoop {
    long long n_pairs, first, stride;
    __device__ PairLoop(long long rows) {
        long long n_tiles = (rows + kRowsTile - 1) / kRowsTile;
        n_pairs = (n_tiles + 1) / 2;
        first = blockIdx.x >> 1;
        stride = gridDim.x >> 1;
    }
};

// weight producer: one thread per CTA streams ITS half of every chunk in schedule order (each step twice: once per sub-tile)
// base_of(p): packed chunk area used by tile pair p (one per latent in the batched FiLM mode)
template <class S, class BaseFn>
__device__ __forceinline__ void producer_loop_fn(const Ctx& cx, BaseFn base_of, const PairLoop& pl, int n_steps, int flag) {
    uint32_t stage = 0, phase = 0;
    for (long long p = pl.first; p < pl.n_pairs; p += pl.stride) {
        const uint8_t* __restrict__ packed = base_of(p);
        for (int s = 0; s < n_steps; ++s) {
            const int nc = S::n_pre(s, flag) + S::n_h(s, flag) + S::n_post(s, flag);
            const uint32_t bytes = half_bytes<S>(s);
            const uint8_t* src_w = packed + step_base<S>(s) + (size_t)cx.rank * bytes;
            for (int g = 0; g < 2; ++g) {
                for (int c = 0; c < nc; ++c) {
                    mbar_wait_cluster(cx.w_empty + 8 * stage, phase ^ 1u);
                    mbar_arrive_expect_tx(cx.w_full + 8 * stage, bytes);
                    bulk_g2s(cx.smem + kRingOff + stage * kStageBytes, src_w + (size_t)c * 2 * bytes, bytes, cx.w_full + 8 * stage);
                    if (++stage == kStages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    }
}

template <class S>
__device__ __forceinline__ void producer_loop(const Ctx& cx, const uint8_t* __restrict__ packed, const PairLoop& pl, int n_steps, int flag) {
    producer_loop_fn<S>(cx, [packed](long long) { return packed; }, pl, n_steps, flag);
}

// peer CTA: forward "my half of this stage has landed" to the leader's ring barrier (count 2 there)
template <class S>
__device__ __forceinline__ void relay_loop(const Ctx& cx, const PairLoop& pl, int n_steps, int flag) {
    uint32_t stage = 0, phase = 0;
    const uint32_t remote0 = mapa(cx.w_full, 0);
    for (long long p = pl.first; p < pl.n_pairs; p += pl.stride) {
        for (int s = 0; s < n_steps; ++s) {
            const int nc = 2 * (S::n_pre(s, flag) + S::n_h(s, flag) + S::n_post(s, flag));
            for (int c = 0; c < nc; ++c) {
                mbar_wait_cluster(cx.w_full + 8 * stage, phase);
                mbar_arrive_cluster(remote0 + 8 * stage);
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
        }
    }
}

// MMA issuer: warp 1 of the leader CTA, converged (every lane walks the schedule and polls the barriers; one elected
// lane issues), so the descriptors live in uniform registers.  Alternates the two sub-tiles step by step.
template <class S>
__device__ __forceinline__ void mma_loop(const Ctx& cx, uint32_t tmem_base, const PairLoop& pl, int n_steps, int flag) {
    uint32_t stage = 0, phase = 0, act_phase0 = 0, act_phase1 = 0;
    const uint64_t d_hi = desc_sw128(0);                          // A and B: K-major SWIZZLE_128B, zero address field
    auto issue_chunk = [&](uint32_t d_tmem, uint32_t a_addr, uint32_t idesc, uint32_t accumulate, int n_mma) {
        mbar_wait_cluster(cx.w_full + 8 * stage, phase);
        tc_fence_after();
        const uint32_t b_addr = cx.smem + kRingOff + stage * kStageBytes;
        const uint64_t ad = d_hi | (uint64_t)((a_addr >> 4) & 0x3FFFu);
        const uint64_t bd = d_hi | (uint64_t)((b_addr >> 4) & 0x3FFFu);
        if (elect_one()) {
            mma_bf16_2cta(d_tmem, ad, bd, idesc, accumulate);
            for (int k = 1; k < n_mma; ++k) mma_bf16_2cta(d_tmem, ad + 2 * k, bd + 2 * k, idesc, 1u);   // +32 B = next 16 K
            mma_commit_2cta(cx.w_empty + 8 * stage, (uint16_t)3);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
    };
    for (long long p = pl.first; p < pl.n_pairs; p += pl.stride) {
        for (int s = 0; s < n_steps; ++s) {
            const uint32_t idesc = make_idesc_bf16(256, (uint32_t)S::n(s));
            const int n_pre = S::n_pre(s, flag), n_h = S::n_h(s, flag), n_post = S::n_post(s, flag);
            for (int g = 0; g < 2; ++g) {
                if (g == 0) { mbar_wait_cluster(cx.act_ready, act_phase0); act_phase0 ^= 1u; }
                else { mbar_wait_cluster(cx.act_ready + 8, act_phase1); act_phase1 ^= 1u; }
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)g * 256u;
                const uint32_t a_base = cx.smem + (uint32_t)g * kSubBytes;
                uint32_t acc = 0;
                for (int c = 0; c < n_pre; ++c) { issue_chunk(d_tmem, a_base, idesc, acc, 4); acc = 1; }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (c < n_h) { issue_chunk(d_tmem, a_base + kPeBytes + (uint32_t)c * 16384u, idesc, acc, 4); acc = 1; }
                }
                for (int c = 0; c < n_post; ++c) issue_chunk(d_tmem, a_base, idesc, 1u, S::kPostMmas);
                if (elect_one()) mma_commit_2cta(cx.acc_full + 8 * g, (uint16_t)3);
                __syncwarp();
            }
        }
    }
}

// ---- the same three roles on the compact map (MapC): 4 stages, 4 copies per step and sub-tile, chunk c in stage c ---------------
// extra_off: byte offset of the blob area inside a packed image (behind its fp32 tables)
template <class S, class BaseFn>
__device__ __forceinline__ void producer_loop_c(const Ctx& cx, BaseFn base_of, long long extra_off, const PairLoop& pl, int n_steps) {
    uint32_t phase = 0;
    for (long long p = pl.first; p < pl.n_pairs; p += pl.stride) {
        const uint8_t* __restrict__ packed = base_of(p);
        for (int s = 0; s < n_steps; ++s) {
            const uint32_t hb = half_bytes<S>(s), pb = post_bytes<S>(s);
            const uint8_t* src_w = packed + step_base<S>(s) + (size_t)cx.rank * hb;
            const uint8_t* src_x = packed + extra_off + extra_base<S>(s) + (size_t)cx.rank * (hb + pb);
            for (int g = 0; g < 2; ++g) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint32_t bytes = c < 3 ? hb : hb + pb;
                    mbar_wait_cluster(cx.w_empty + 8 * c, phase ^ 1u);
                    mbar_arrive_expect_tx(cx.w_full + 8 * c, bytes);
                    bulk_g2s(cx.smem + MapC::kRing + (uint32_t)c * kStageBytes, c < 3 ? src_w + (size_t)c * 2 * hb : src_x, bytes, cx.w_full + 8 * c);
                }
                phase ^= 1u;
            }
        }
    }
}

__device__ __forceinline__ void relay_loop_c(const Ctx& cx, const PairLoop& pl, int n_steps) {
    uint32_t phase = 0;
    const uint32_t remote0 = mapa(cx.w_full, 0);
    for (long long p = pl.first; p < pl.n_pairs; p += pl.stride) {
        for (int sg = 0; sg < 2 * n_steps; ++sg) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                mbar_wait_cluster(cx.w_full + 8 * c, phase);
                mbar_arrive_cluster(remote0 + 8 * c);
            }
            phase ^= 1u;
        }
    }
}

template <class S>
__device__ __forceinline__ void mma_loop_c(const Ctx& cx, uint32_t tmem_base, const PairLoop& pl, int n_steps) {
    static_assert(S::kPostMmas == 1, "the compact post chunk is one K = 16 step");
    uint32_t phase = 0, act_phase0 = 0, act_phase1 = 0;
    const uint64_t d_hi = desc_sw128(0);                          // h blocks and h chunks: K-major SWIZZLE_128B
    const uint64_t a_ns = make_desc(0, MapC::kAuxLbo, MapC::kAuxSbo, 0);      // aux operand: no swizzle, chunk pitch 2 KB
    for (long long p = pl.first; p < pl.n_pairs; p += pl.stride) {
        for (int s = 0; s < n_steps; ++s) {
            const uint32_t idesc = make_idesc_bf16(256, (uint32_t)S::n(s));
            const uint32_t hb = half_bytes<S>(s);
            const uint64_t b_ns = make_desc(0, post_bytes<S>(s) / 2, 128, 0);  // post chunk: chunk pitch = weight rows x 16 B
            for (int g = 0; g < 2; ++g) {
                if (g == 0) { mbar_wait_cluster(cx.act_ready, act_phase0); act_phase0 ^= 1u; }
                else { mbar_wait_cluster(cx.act_ready + 8, act_phase1); act_phase1 ^= 1u; }
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)g * 256u;
                const uint32_t a_base = cx.smem + (uint32_t)g * MapC::kSub;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    mbar_wait_cluster(cx.w_full + 8 * c, phase);
                    tc_fence_after();
                    const uint32_t b_addr = cx.smem + MapC::kRing + (uint32_t)c * kStageBytes;
                    const uint64_t ad = d_hi | (uint64_t)(((a_base + MapC::kAux + (uint32_t)c * 16384u) >> 4) & 0x3FFFu);
                    const uint64_t bd = d_hi | (uint64_t)((b_addr >> 4) & 0x3FFFu);
                    if (elect_one()) {
                        mma_bf16_2cta(d_tmem, ad, bd, idesc, c == 0 ? 0u : 1u);
#pragma unroll
                        for (int k = 1; k < 4; ++k) mma_bf16_2cta(d_tmem, ad + 2 * k, bd + 2 * k, idesc, 1u);
                        if (c == 3)
                            mma_bf16_2cta(d_tmem, a_ns | (uint64_t)((a_base >> 4) & 0x3FFFu), b_ns | (uint64_t)(((b_addr + hb) >> 4) & 0x3FFFu), idesc, 1u);
                        mma_commit_2cta(cx.w_empty + 8 * c, (uint16_t)3);
                    }
                    __syncwarp();
                }
                phase ^= 1u;
                if (elect_one()) mma_commit_2cta(cx.acc_full + 8 * g, (uint16_t)3);
                __syncwarp();
            }
        }
    }
}

// common prologue: barriers, TMEM allocation (both CTAs), cluster rendezvous; returns the TMEM base address
// act_count: arrivals per act_ready phase (epilogue warps per sub-tile x 2 CTAs)
__device__ __forceinline__ uint32_t tc_prologue(const Ctx& cx, int warp, uint32_t act_count = 16, uint32_t spill_count = 8, int n_stages = kStages) {
    if (threadIdx.x == 0) {
        for (int i = 0; i < n_stages; ++i) { mbar_init(cx.w_full + 8 * i, cx.rank == 0 ? 2 : 1); mbar_init(cx.w_empty + 8 * i, 1); }
        for (int g = 0; g < 2; ++g) {
            mbar_init(cx.act_ready + 8 * g, act_count); mbar_init(cx.acc_full + 8 * g, 1);
            mbar_init(cx.spill_ready + 8 * g, spill_count); mbar_init(cx.spill_done + 8 * g, 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2cta(cx.tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(cx.tmem_slot));
    return tmem_base;
}
__device__ __forceinline__ void tc_teardown(uint32_t tmem_base, int warp) {
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) tmem_dealloc_2cta(tmem_base, 512);
}

// epilogue warp -> leader's act_ready[g]: this warp's rows of the next A operand are written and its TMEM reads are done
__device__ __forceinline__ void arrive_act(uint32_t act_bar_local, uint32_t act_bar_leader, uint32_t rank, int lane) {
    fence_proxy_async_smem();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
        if (rank == 0) mbar_arrive_release_cluster_local(act_bar_local);
        else mbar_arrive_cluster(act_bar_leader);
    }
}

// packed fp32x2 fma (FFMA2): {d0,d1} = {a0,a1} * {b,b} + {c0,c1}, each lane rounded like fmaf
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b, float c0, float c1) {
    asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %4};\n\tmov.b64 rc, {%5, %6};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b), "f"(c0), "f"(c1));
}

// sin of a PAIR of arguments on the FMA pipe (packed fp32x2: 9 instructions per pair, none on the MUFU / XU pipe):
// u = x / 2 pi, r = u - round(u) in [-0.5, 0.5] (magic-number rounding), sin(2 pi r) = r P(r^2) with the degree-4 minimax P
// (max error 6.2e-6 over the period, evaluated in fp32; the bf16 rounding of the result is 2e-3).  B2R_SIN_POLY_PAIRS of every 8 pairs
// of a sine epilogue take this path instead of MUFU.SIN.  Measured three times (scalar form in round 1, packed form twice in round 2,
// the last time after the weight ring and the input layer had stopped co-limiting the kernels: profiles/r2_sin_poly_ab.txt): 0 / 1 / 2 /
// 3 pairs differ by 1-3 % on the 256^3 grid and by less than the box-to-box spread on the pi-GAN batch and the SirenNeRF frame -- the
// kernels run under the board's power cap, where a sine moved to the FMA pipe is not free.  Default 0 (every sine on MUFU.SIN).
#ifndef B2R_SIN_POLY_PAIRS
#define B2R_SIN_POLY_PAIRS 0
#endif
constexpr int kSinPolyPairs = B2R_SIN_POLY_PAIRS;
__device__ __forceinline__ void sin_poly2(float x0, float x1, float& o0, float& o1) {
    asm("{\n\t.reg .b64 x, m, r, s, p, c;\n\t.reg .b32 i, g;\n\t"
        "mov.b64 x, {%2, %3};\n\t"
        "mov.b32 i, 0f3E22F983;\n\tmov.b64 c, {i, i};\n\t"                    // 1 / 2 pi
        "mov.b32 g, 0f4B400000;\n\tmov.b64 m, {g, g};\n\t"                    // 1.5 * 2^23
        "fma.rn.f32x2 r, x, c, m;\n\t"                                         // u + magic: u rounded to an integer k
        "mov.b32 i, 0fBF800000;\n\tmov.b64 s, {i, i};\n\t"
        "fma.rn.f32x2 r, r, s, m;\n\t"                                         // magic - (k + magic) = -k
        "fma.rn.f32x2 r, x, c, r;\n\t"                                         // r = u - k
        "mul.rn.f32x2 s, r, r;\n\t"
        "mov.b32 i, 0f4203202E;\n\tmov.b64 p, {i, i};\n\t"                    // 32.78142547607422
        "mov.b32 i, 0fC294F4BF;\n\tmov.b64 c, {i, i};\n\tfma.rn.f32x2 p, p, s, c;\n\t"     // -74.47801971435547
        "mov.b32 i, 0f42A2BBD0;\n\tmov.b64 c, {i, i};\n\tfma.rn.f32x2 p, p, s, c;\n\t"     // 81.3668212890625
        "mov.b32 i, 0fC225532A;\n\tmov.b64 c, {i, i};\n\tfma.rn.f32x2 p, p, s, c;\n\t"     // -41.331214904785156
        "mov.b32 i, 0f40C90ECB;\n\tmov.b64 c, {i, i};\n\tfma.rn.f32x2 p, p, s, c;\n\t"     // 6.283055782318115
        "mul.rn.f32x2 p, p, r;\n\t"
        "mov.b64 {%0, %1}, p;\n\t}"
        : "=f"(o0), "=f"(o1) : "f"(x0), "f"(x1));
}
// 16 sines of one accumulator unit: the last kSinPolyPairs pairs on the FMA pipe, the others MUFU.SIN
__device__ __forceinline__ void sin16(const uint32_t (&v)[16], float (&f)[16]) {
#pragma unroll
    for (int pr = 0; pr < 8; ++pr) {
        if (pr >= 8 - kSinPolyPairs) sin_poly2(__uint_as_float(v[2 * pr]), __uint_as_float(v[2 * pr + 1]), f[2 * pr], f[2 * pr + 1]);
        else { f[2 * pr] = __sinf(__uint_as_float(v[2 * pr])); f[2 * pr + 1] = __sinf(__uint_as_float(v[2 * pr + 1])); }
    }
}

// Input layer (K = 3) of the sine models in the INFERENCE kernels: h0[row, n] = sin(w'[0][n] p_x + w'[1][n] p_y + w'[2][n] p_z + shift[n])
// for the warp's 32 rows (quad) and 64 columns (K-block cq), written as bf16 into the SWIZZLE_128B K-block at `kblk`.
// The epilogue's thread = row mapping would make every thread read the whole table (4 x 16 B per 4 outputs, the same address in all
// lanes: a broadcast still returns 512 B per warp-load, and the stage was bound by the shared-memory return path: 160 LDS.128 per
// thread and tile).  Here a lane OWNS 4 consecutive columns (its 16 table values stay in registers), the half-warps take
// alternate rows and the row's position comes by shuffle: per 4 sines 3 SHFL + 6 FFMA2 + 1 STS.64 instead of 5 LDS.128 + 16 FFMA.
// Same fma chain per element as the thread = row form (bit-identical).  tab_w / tab_sh: shared-memory addresses of w'[3][256], shift[256];
// pnt: THIS lane's row's position (row quad * 32 + lane).
__device__ __forceinline__ void sine_input_layer(uint32_t tab_w, uint32_t tab_sh, uint32_t kblk, int cq, int quad, int lane, const float (&pnt)[3]) {
    const uint32_t half = (uint32_t)lane >> 4, j = (uint32_t)lane & 15u;
    const uint32_t n0 = ((uint32_t)cq * 64u + 4u * j) * 4u;
    const float4 wx = lds128(tab_w + n0), wy = lds128(tab_w + 1024u + n0), wz = lds128(tab_w + 2048u + n0), sh = lds128(tab_sh + n0);
    const uint32_t base = kblk + (uint32_t)quad * 4096u + half * 128u + (j & 1u) * 8u;      // 32 rows = 4 groups of 8 rows (1,024 B)
#pragma unroll
    for (int it = 0; it < 16; ++it) {
        const int src = 2 * it + (int)half;
        const float px = __shfl_sync(0xffffffffu, pnt[0], src), py = __shfl_sync(0xffffffffu, pnt[1], src), pz = __shfl_sync(0xffffffffu, pnt[2], src);
        float t0, t1, t2, t3;
        ffma2(t0, t1, wx.x, wx.y, px, sh.x, sh.y);
        ffma2(t2, t3, wx.z, wx.w, px, sh.z, sh.w);
        ffma2(t0, t1, wy.x, wy.y, py, t0, t1);
        ffma2(t2, t3, wy.z, wy.w, py, t2, t3);
        ffma2(t0, t1, wz.x, wz.y, pz, t0, t1);
        ffma2(t2, t3, wz.z, wz.w, pz, t2, t3);
        const uint32_t r7 = (uint32_t)((2 * it) & 7) + half;                                  // row & 7 of row quad * 32 + 2 it + half
        const uint32_t addr = base + (uint32_t)((2 * it) >> 3) * 1024u + (uint32_t)((2 * it) & 7) * 128u + (((j >> 1) ^ r7) << 4);
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(pack_bf16(__sinf(t0), __sinf(t1))), "r"(pack_bf16(__sinf(t2), __sinf(t3))) : "memory");
    }
}

// packed fp32x2 add (FADD2): {a0,a1} += {b0,b1}
__device__ __forceinline__ void add2(float& a0, float& a1, float b0, float b1) {
    asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%0, %1};\n\tmov.b64 rb, {%2, %3};\n\tadd.rn.f32x2 rc, ra, rb;\n\tmov.b64 {%0, %1}, rc;\n\t}"
        : "+f"(a0), "+f"(a1) : "f"(b0), "f"(b1));
}

}  // namespace tc
}  // namespace b2r
