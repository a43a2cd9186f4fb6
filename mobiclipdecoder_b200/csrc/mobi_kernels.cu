// Mobiclip reconstruction kernels for sm_100a.  "MD:n" = LibMobiclip/Codec/Mobiclip/MobiclipDecoder.cs:n.
//
// Data layout in HBM (identical to the reference's managed arrays so that its flat, unclamped
// addressing falls out for free): a picture is Stride*H luma bytes followed by Stride*H/2 chroma
// bytes, U in columns [0,Stride/2) and V in [Stride/2,Stride) of each chroma row (MD:107-108, 267-268).
// Stride padding is zero and never written.
//
// k_inter_chunk : inter macroblocks.  Persistent warps draw 16-macroblock chunks by ticket (the last pictures run by run) and
//               work them off four macroblocks at a time: reference windows arrive by TMA (the ring is one rank-3 luma
//               tensor and one rank-4 U|V tensor) into per-warp shared memory, half-pel filter on packed bytes (truncating
//               averages MD:418-456), the run's coded blocks pooled for the transform passes (eight lanes per 8x8 block),
//               saturating pack onto prediction tiles, 16-byte stores.  mobi_inter_v3.cuh / mobi_inter_split.cuh hold two
//               other complete formulations (MOBI_INTER_KERNEL=v3 / split).
// k_intra     : intra macroblocks inside P-pictures (and of steps with more I-pictures than SMs), one warp each, drawn by
//               ticket from a dependency-ordered list (longest chain of dependents first); the neighbourhood (row above incl.
//               top-right, column left, and the not-yet-decoded pixels to the right, which the reference reads as 0 from its
//               freshly allocated planes MD:107) is staged in shared memory, availability decided by coordinates, never by
//               memory contents.
// k_intra_key : I-pictures, one CTA per picture, a luma and a chroma warp per macroblock row, neighbour pixels handed on
//               through shared memory (line buffers, progress counters).
// k_bgra / k_pack_i420: output conversions.
// No tensor cores: these are 8-bit fixed-point butterflies and byte shuffles.
#include <cstdlib>
#include <cstring>
#include "mobi_kernels.h"

namespace mobi {
namespace {

// ids of the set bits of a 6-bit coded-block mask, one nibble each, lowest first (slot order of the coefficient buffers)
__constant__ uint32_t c_blklist[64];
constexpr int INTRA_WARPS = 4;

// ------------------------------------------------------------------------------------------------
// shared arithmetic
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t half4(uint32_t x) { return (x >> 1) & 0x7f7f7f7fu; }

__device__ __forceinline__ void bfly8(const int32_t in[8], int32_t out[8]) {  // MD:3452-3485
    int32_t a0 = in[0] + in[4], a1 = in[0] - in[4], a2 = in[2] + (in[6] >> 1), a3 = (in[2] >> 1) - in[6];
    int32_t e0 = a0 + a2, e3 = a0 - a2, e1 = a1 + a3, e2 = a1 - a3;
    int32_t b0 = in[1] + in[7] - in[3] - (in[3] >> 1), b1 = in[7] - in[1] + in[5] + (in[5] >> 1);
    int32_t b2 = in[5] - (in[7] + (in[7] >> 1)) - in[3], b3 = in[3] + in[5] + in[1] + (in[1] >> 1);
    int32_t o1 = b2 + (b3 >> 2), o7 = b3 - (b2 >> 2), o3 = b0 + (b1 >> 2), o5 = (b0 >> 2) - b1;
    out[0] = e0 + o7; out[7] = e0 - o7; out[1] = e1 + o5; out[6] = e1 - o5;
    out[2] = e2 + o3; out[5] = e2 - o3; out[3] = e3 + o1; out[4] = e3 - o1;
}
__device__ __forceinline__ void bfly4(const int32_t in[4], int32_t out[4]) {  // MD:3740-3747
    int32_t s = in[0] + in[2], d = in[0] - in[2], p = (in[1] >> 1) - in[3], q = in[1] + (in[3] >> 1);
    out[0] = s + q; out[3] = s - q; out[1] = d + p; out[2] = d - p;
}
__device__ __forceinline__ int clip255(int v) { return min(max(v, 0), 255); }  // Vx2MinMaxTable, MobiConst.cs:587
__device__ __forceinline__ uint32_t addclip4(uint32_t px, const int32_t* r) {  // r = transform outputs before >>6
    uint32_t o = (uint32_t)clip255((int)(px & 255u) + (r[0] >> 6));
    o |= (uint32_t)clip255((int)((px >> 8) & 255u) + (r[1] >> 6)) << 8;
    o |= (uint32_t)clip255((int)((px >> 16) & 255u) + (r[2] >> 6)) << 16;
    o |= (uint32_t)clip255((int)(px >> 24) + (r[3] >> 6)) << 24;
    return o;
}

// ------------------------------------------------------------------------------------------------
// motion compensation primitives (CopyBlock MD:418-456)
// ------------------------------------------------------------------------------------------------
// 8 (or 4) consecutive bytes at an arbitrary address, plus the same run one byte further on.
__device__ __forceinline__ void ld_row8(const uint8_t* p, uint32_t& a0, uint32_t& a1, uint32_t& b0, uint32_t& b1) {
    uintptr_t ad = (uintptr_t)p;
    uint32_t sh = ((uint32_t)ad & 3u) * 8u;
    const uint32_t* q = (const uint32_t*)(ad & ~(uintptr_t)3);
    uint32_t w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2);
    a0 = __funnelshift_r(w0, w1, sh); a1 = __funnelshift_r(w1, w2, sh);
    b0 = __funnelshift_rc(w0, w1, sh + 8u); b1 = __funnelshift_rc(w1, w2, sh + 8u);
}
__device__ __forceinline__ void ld_row4(const uint8_t* p, uint32_t& a0, uint32_t& b0) {
    uintptr_t ad = (uintptr_t)p;
    uint32_t sh = ((uint32_t)ad & 3u) * 8u;
    const uint32_t* q = (const uint32_t*)(ad & ~(uintptr_t)3);
    uint32_t w0 = __ldg(q), w1 = __ldg(q + 1);
    a0 = __funnelshift_r(w0, w1, sh);
    b0 = __funnelshift_rc(w0, w1, sh + 8u);
}
__device__ __forceinline__ void mc_row8(const uint8_t* p, int S, int phase, uint32_t& o0, uint32_t& o1) {
    uint32_t a0, a1, b0, b1;
    ld_row8(p, a0, a1, b0, b1);
    if (phase == 0) { o0 = a0; o1 = a1; return; }
    if (phase == 1) { o0 = half4(a0) + half4(b0); o1 = half4(a1) + half4(b1); return; }
    uint32_t c0, c1, d0, d1;
    ld_row8(p + S, c0, c1, d0, d1);
    if (phase == 2) { o0 = half4(a0) + half4(c0); o1 = half4(a1) + half4(c1); return; }
    o0 = half4(half4(a0) + half4(b0)) + half4(half4(c0) + half4(d0));
    o1 = half4(half4(a1) + half4(b1)) + half4(half4(c1) + half4(d1));
}
__device__ __forceinline__ uint32_t mc_row4(const uint8_t* p, int S, int phase) {
    uint32_t a0, b0;
    ld_row4(p, a0, b0);
    if (phase == 0) return a0;
    if (phase == 1) return half4(a0) + half4(b0);
    uint32_t c0, d0;
    ld_row4(p + S, c0, d0);
    if (phase == 2) return half4(a0) + half4(c0);
    return half4(half4(a0) + half4(b0)) + half4(half4(c0) + half4(d0));
}
__device__ __forceinline__ uint32_t mc_px(const uint8_t* p, int S, int phase) {
    uint32_t a = __ldg(p);
    if (phase == 0) return a;
    if (phase == 1) return (a >> 1) + (__ldg(p + 1) >> 1);
    if (phase == 2) return (a >> 1) + (__ldg(p + S) >> 1);
    return (((a >> 1) + (__ldg(p + 1) >> 1)) >> 1) + (((__ldg(p + S) >> 1) + (__ldg(p + S + 1) >> 1)) >> 1);
}
struct PartV { int mvx, mvy, ref; };
__device__ __forceinline__ PartV part_of(uint32_t x, uint32_t y) {  // the two words of a mobi_part
    PartV v;
    v.mvx = (int)(int16_t)(x >> 16);
    v.mvy = (int)(int16_t)(y & 0xFFFFu);
    v.ref = (int)((x >> 12) & 15u);
    return v;
}
// Four pixels + four residuals (transform outputs before >>6) -> saturated bytes.  cvt.pack.sat.u8.s32 packs two
// clamped values per instruction: d = sat(a) << 8 | sat(b) | c << 16.  Saturation == Vx2MinMaxTable (MobiConst.cs:587)
// on its whole domain.
__device__ __forceinline__ uint32_t addsat4(uint32_t px, int r0, int r1, int r2, int r3) {
    const int p0 = (int)(px & 255u) + (r0 >> 6), p1 = (int)((px >> 8) & 255u) + (r1 >> 6);
    const int p2 = (int)((px >> 16) & 255u) + (r2 >> 6), p3 = (int)(px >> 24) + (r3 >> 6);
    uint32_t hi, out;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(p3), "r"(p2), "r"(0));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(out) : "r"(p1), "r"(p0), "r"(hi));
    return out;
}

// ------------------------------------------------------------------------------------------------
// inter macroblocks
// ------------------------------------------------------------------------------------------------
// Reference pixels arrive by TMA: the whole ring is one rank-3 u8 tensor (Stride, 1.5*H, pictures), so a macroblock's
// 16x16+halo luma window lies inside one 32x17 box, and (rank-4 view, each row split into its U and V halves) both chroma
// windows inside one 32x2x9 box, fetched by one lane straight into shared memory (TMA wants the innermost start coordinate
// 16-byte aligned -- an unaligned start raises "illegal instruction" -- so boxes start at the window's column rounded down
// to 16 and lanes apply the remainder when they read shared memory).  This takes the gathers off the LSU path, which is
// what bounds a load-per-lane formulation (every warp-wide load touches 16 cache lines).  TMA zero-fills outside the
// tensor whereas the reference addresses its planes flat (a column < 0 or >= Stride wraps into the neighbouring row), so
// windows that leave their pixel row take a load-per-lane path instead.
constexpr uint32_t TMA_BYTES_L = 32 * 17;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}

// Same arithmetic as mc_row8 / mc_row4 on a box row held in shared memory (row pitch 32).  base is 4-byte aligned (the
// boxes are 128-byte aligned), off any byte offset: the misalignment comes from off alone, no address conversion needed.
__device__ __forceinline__ void lds_row8(const uint8_t* base, uint32_t off, uint32_t& a0, uint32_t& a1, uint32_t& b0, uint32_t& b1) {
    const uint32_t sh = (off & 3u) * 8u;
    const uint32_t* q = reinterpret_cast<const uint32_t*>(base + (off & ~3u));
    const uint32_t w0 = q[0], w1 = q[1], w2 = q[2];
    a0 = __funnelshift_r(w0, w1, sh); a1 = __funnelshift_r(w1, w2, sh);
    b0 = __funnelshift_rc(w0, w1, sh + 8u); b1 = __funnelshift_rc(w1, w2, sh + 8u);
}
__device__ __forceinline__ void tile_row8(const uint8_t* base, uint32_t off, int phase, uint32_t& o0, uint32_t& o1) {
    uint32_t a0, a1, b0, b1;
    lds_row8(base, off, a0, a1, b0, b1);
    if (phase == 0) { o0 = a0; o1 = a1; return; }
    if (phase == 1) { o0 = half4(a0) + half4(b0); o1 = half4(a1) + half4(b1); return; }
    uint32_t c0, c1, d0, d1;
    lds_row8(base, off + 32u, c0, c1, d0, d1);
    if (phase == 2) { o0 = half4(a0) + half4(c0); o1 = half4(a1) + half4(c1); return; }
    o0 = half4(half4(a0) + half4(b0)) + half4(half4(c0) + half4(d0));
    o1 = half4(half4(a1) + half4(b1)) + half4(half4(c1) + half4(d1));
}
__device__ __forceinline__ void lds_row4(const uint8_t* base, uint32_t off, uint32_t& a0, uint32_t& b0) {
    const uint32_t sh = (off & 3u) * 8u;
    const uint32_t* q = reinterpret_cast<const uint32_t*>(base + (off & ~3u));
    const uint32_t w0 = q[0], w1 = q[1];
    a0 = __funnelshift_r(w0, w1, sh);
    b0 = __funnelshift_rc(w0, w1, sh + 8u);
}
__device__ __forceinline__ uint32_t tile_row4(const uint8_t* base, uint32_t off, int phase) {
    uint32_t a0, b0;
    lds_row4(base, off, a0, b0);
    if (phase == 0) return a0;
    if (phase == 1) return half4(a0) + half4(b0);
    uint32_t c0, d0;
    lds_row4(base, off + 32u, c0, d0);
    if (phase == 2) return half4(a0) + half4(c0);
    return half4(half4(a0) + half4(b0)) + half4(half4(c0) + half4(d0));
}
__device__ __forceinline__ uint32_t tile_px(const uint8_t* t, int pitch, int phase) {
    const uint32_t a = t[0];
    if (phase == 0) return a;
    if (phase == 1) return (a >> 1) + (t[1] >> 1);
    if (phase == 2) return (a >> 1) + (t[pitch] >> 1);
    return (((a >> 1) + (t[1] >> 1)) >> 1) + (((t[pitch] >> 1) + (t[pitch + 1] >> 1)) >> 1);
}

// ------------------------------------------------------------------------------------------------
// L2 prefetch of a chunk's reference region
// ------------------------------------------------------------------------------------------------
// EXPERIMENT (MOBI_INTER_EXP=8), measured and not adopted.  One of the two things that bound the inter path (the other is
// instruction issue: DESIGN.md section 4) is the rate at which the boxes' rows come out of the memory system
// (tools/probe/tma_rate.cu: a 32x17 box costs 29 SM-cycles when its rows hit L2 and 70-115 when they come from DRAM; the kernel
// issues 1.9 boxes per macroblock and takes 165 SM-cycles per macroblock with 44 % of its sectors missing L2): every box row
// is its own sector or two, rows lie Stride bytes apart.  Most leaves point into the previous picture close to where they sit, so the
// idea was to ask L2, before a warp issues a chunk's boxes, for the whole region of picture 1 the chunk's windows can be
// expected in, as full 128-byte lines.  Result on the bench mix: DRAM reads 414 -> 517 MB per launch, L2 hit rate unchanged
// (52 %), kernel 0.263 -> 0.301 ms: the lines a chunk asks for are mostly the ones its neighbours' boxes already brought, and
// what misses are the 20 % of leaves that point into OTHER pictures, which no such region covers.
template <int LOG2S>
__device__ __forceinline__ void prefetch_rows(const uint8_t* plane, int x_lo, int x_hi, int y_lo, int y_hi, int y_max, int lane) {
    constexpr int S = 1 << LOG2S;
    x_lo = max(x_lo, 0) & ~127; x_hi = min(x_hi, S);
    y_lo = max(y_lo, 0); y_hi = min(y_hi, y_max);
    const int lines = (x_hi - x_lo + 127) >> 7, n = lines * (y_hi - y_lo);
    if (lines <= 0) return;
    for (int i = lane; i < n; i += 32) {
        const int r = i / lines, c = i - r * lines;
        asm volatile("prefetch.global.L2 [%0];" :: "l"(plane + ((size_t)(y_lo + r) << LOG2S) + x_lo + c * 128));
    }
}
// Macroblocks first .. first + count - 1 (raster order, may run over the end of a macroblock row) of a W-pixel-wide picture:
// luma rows -MARGIN .. 16 + MARGIN around each row of macroblocks, the U and V rows that go with them.
template <int LOG2S>
__device__ __forceinline__ void prefetch_chunk_region(const uint8_t* ref, int H, int mbw, int first_mbx, int first_mby, int count, int lane) {
    constexpr int S = 1 << LOG2S, MARGIN = 9;
    if (!ref) return;
    const uint8_t* const chroma = ref + ((size_t)H << LOG2S);
    int mbx = first_mbx, mby = first_mby;
    while (count > 0) {
        const int n = min(count, mbw - mbx);
        const int x0 = mbx * 16 - MARGIN, x1 = (mbx + n) * 16 + MARGIN, y0 = mby * 16 - MARGIN, y1 = mby * 16 + 16 + MARGIN;
        prefetch_rows<LOG2S>(ref, x0, x1, y0, y1, H, lane);
        prefetch_rows<LOG2S>(chroma, x0 >> 1, (x1 + 1) >> 1, y0 >> 1, (y1 + 1) >> 1, H >> 1, lane);                        // U
        prefetch_rows<LOG2S>(chroma + (S >> 1), x0 >> 1, (x1 + 1) >> 1, y0 >> 1, (y1 + 1) >> 1, H >> 1, lane);             // V
        count -= n; mbx = 0; mby++;
    }
}

// ------------------------------------------------------------------------------------------------
// inter macroblocks, chunk kernel (the default): shared-memory layout and box helpers
// ------------------------------------------------------------------------------------------------
// What k_inter spends most of its issue slots on is work every lane of the warp repeats for its one macroblock
// (descriptor decode, coordinates, eligibility, box addresses) and transform passes whose eight-lane groups are half
// empty (1.8 coded 8x8 blocks per macroblock on the bench mix, four groups per pass).  k_inter_chunk works on runs of
// RUN = 4 consecutive macroblocks per warp:
//   * set-up is lane-parallel: lane l decodes macroblock l of the chunk; whatever is needed warp-wide later travels by shuffle;
//   * the lanes that hold a run's macroblocks issue their boxes up front (one luma box, one rank-4 box holding the U and
//     the V window, per leaf), so a whole run's windows are in flight per warp;
//   * prediction goes macroblock by macroblock into per-macroblock tiles in shared memory;
//   * the coefficients of the whole run (one contiguous range of the picture's array) are dequantised together into
//     one pool of block buffers (which reuse the space of the boxes), and the transform passes walk the pool, four coded
//     blocks per pass regardless of which macroblock they belong to;
//   * the tiles leave with 16-byte stores, lane-parallel over the macroblocks again.
// Macroblocks with more than two leaves and windows that leave their pixel row (flat addressing wraps, TMA zero-fills)
// take a load-per-lane path.
constexpr uint32_t TMA_BYTES_C4 = 32 * 2 * 9;

template <int RUN>
struct RunSmem {
    static constexpr int BOX = RUN * 1280;        // RUN luma boxes (640 apart), then RUN chroma boxes (640 apart)
    static constexpr int SCR = 640;               // leaf records + partition map of the split macroblock in hand
    static constexpr int SLOTS = 6 * RUN;         // coded 8x8 blocks of a run
    static constexpr int REGION = BOX + SCR > SLOTS * 256 ? BOX + SCR : SLOTS * 256;
    union {
        uint8_t raw[REGION];
        int32_t coef[SLOTS][64];
    } u;
    static constexpr int TILE = 416;   // 384 used; 416 keeps the four tiles 32 bytes apart modulo 128: the run-parallel 16-byte reads of the final store hit distinct banks
    uint8_t tile[RUN][TILE];    // prediction (+ residual): luma 16 rows x 16, then U 8x8, V 8x8
    uint32_t slotinfo[SLOTS];   // per pooled block, in the order the passes visit them: byte offset of its pixels in tile[][] | chroma << 16 | 8x8 transform << 17 | pool slot << 18
    uint64_t bar[RUN];
    uint8_t pad[(128 - (REGION + RUN * TILE + SLOTS * 4 + RUN * 8) % 128) % 128];
};
static_assert(sizeof(RunSmem<4>) % 128 == 0, "TMA destinations must stay 128-byte aligned");
static_assert(sizeof(RunSmem<4>) * 4 + 1024 <= 233472 / 7, "up to seven 4-warp CTAs per SM (six are launched: 80 registers per thread measured 1 % faster than seven at 72)");

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 :: "r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}
// tile_row4 on the merged chroma box: rows of the U and V windows alternate, so the next pixel row is 64 bytes on.
__device__ __forceinline__ uint32_t tile_row4_p64(const uint8_t* base, uint32_t off, int phase) {
    uint32_t a0, b0;
    lds_row4(base, off, a0, b0);
    if (phase == 0) return a0;
    if (phase == 1) return half4(a0) + half4(b0);
    uint32_t c0, d0;
    lds_row4(base, off + 64u, c0, d0);
    if (phase == 2) return half4(a0) + half4(c0);
    return half4(half4(a0) + half4(b0)) + half4(half4(c0) + half4(d0));
}

// ------------------------------------------------------------------------------------------------
// inter macroblocks, persistent chunk kernel: warps draw chunks of 16 consecutive macroblocks (four runs of four)
// ------------------------------------------------------------------------------------------------
// CTAs that live for one run (about 1700 instructions per warp) lose a third of the warp slots to CTA turnover and to
// uneven warps inside a CTA, and every run starts with dependent trips to memory (job table, descriptors, leaves).  So
// the grid is sized to the machine and warps draw chunks of 16 macroblocks through an atomic ticket (the next ticket is
// requested before the current chunk is worked on).  Lanes 0..15 decode the chunk's 16 descriptors at once and fetch
// leaves 0..3 of the split ones; the four runs of the chunk then go through the steps listed above with their
// per-macroblock facts arriving by shuffle.
constexpr int CH_WARPS = 4, CH_MBS = 16;

// Word index of element (row, col) of a pooled 8x8 coefficient block.  Plain row-major puts rows r and r+4 of a block,
// and the same row of the four blocks of a pass, on the same banks (16-byte row loads 2-way, transposed stores 4-way
// conflicted: ncu source counters).  Rows 4-7 swap their 16-byte halves, and the low two row bits are XORed with the
// block's slot: row loads, transposed stores and the scatter then spread over all banks.  Every access goes through here.
__device__ __forceinline__ uint32_t pool_word(uint32_t slot, uint32_t row, uint32_t col) {
    return ((row ^ (slot & 3u)) << 3) | (col ^ ((row & 4u)));
}

// Window facts of one leaf as the prediction step wants them, 12 bits: column of the window inside its 16-byte-aligned
// box (4), luma half-pel phase (2), the same for the chroma window (4 + 2).
__device__ __forceinline__ uint32_t leaf_word(int x0, int cx0, int mvx, int mvy) {
    return (uint32_t)(x0 & 15) | (uint32_t)((mvx & 1) | ((mvy & 1) << 1)) << 4 | (uint32_t)(cx0 & 15) << 6
         | (uint32_t)(((mvx >> 1) & 1) | (((mvy >> 1) & 1) << 1)) << 10;
}
// mcw, the per-macroblock word that travels by shuffle: bits 0-11 leaf 0, 12-23 leaf 1, then
constexpr uint32_t MCW_INTER = 1u << 24, MCW_BOX = 1u << 25, MCW_TWO = 1u << 26, MCW_LR = 1u << 27, MCW_SWAP = 1u << 30;   // bits 28-29: first box slot

struct InterTail { uint32_t n_big, first_job, rpp, rpp_magic, big; };   // see the ticket decoding in k_inter_chunk (big: macroblocks per ordinary ticket, 16 or 8)

template <int LOG2S>
__global__ void __launch_bounds__(CH_WARPS * 32, 6)
k_inter_chunk(const DevJob* __restrict__ jobs, uint32_t n_chunks, uint32_t cpp, uint32_t cpp_magic, int mbw, uint32_t mbw_magic, int H,
              uint32_t* __restrict__ ticket, uint32_t ticket_base, uint32_t prefetch_on, InterTail tail,
              const __grid_constant__ CUtensorMap tm_l, const __grid_constant__ CUtensorMap tm_c4) {
    constexpr int RUN = 4;
    using Smem = RunSmem<RUN>;
    __shared__ __align__(128) Smem s_all[CH_WARPS];
    constexpr int S = 1 << LOG2S;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Smem& sm = s_all[warp];
    uint8_t* const raw = sm.u.raw;
    const size_t ysz = (size_t)S * H;
    if (lane < RUN) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&sm.bar[lane])) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    uint32_t phases = 0;   // bit i: parity of the phase barrier i completes next
    const int lrow = lane >> 1, lhalf = lane & 1;
    const int cpl = lane >> 4, crow = (lane >> 1) & 7;
    const int k16 = lane & 15, k4 = lane & 3;
    // the kernel launched behind this one (k_intra, as a programmatic dependent) may take the SMs this grid's CTAs give back while
    // the last tickets are worked off; it waits for this grid's completion before it touches a picture
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    uint32_t t = 0;
    if (lane == 0) asm volatile("atom.global.add.u32 %0, [%1], 1;" : "=r"(t) : "l"(ticket) : "memory");
    t = __shfl_sync(0xffffffffu, t, 0) - ticket_base;
    while (t < n_chunks) {
        uint32_t t_next = 0;   // requested now, looked at when this chunk is done
        if (lane == 0) asm volatile("atom.global.add.u32 %0, [%1], 1;" : "=r"(t_next) : "l"(ticket) : "memory");
        // Tickets below tail.n_big are 16-macroblock chunks of pictures 0 .. tail.first_job - 1; the rest are single runs (4
        // macroblocks) of the last pictures: when the tickets run out, a warp's last piece of work is a quarter as long, so
        // the SMs drain in a quarter of the time (the launch's tail was ~1/6 of its duration: 5.8 chunks per warp).
        uint32_t job, mbc, cnt;
        if (t < tail.n_big) { job = __umulhi(t, cpp_magic); mbc = (t - job * cpp) * tail.big; cnt = tail.big; }
        else { const uint32_t t4 = t - tail.n_big, jq = __umulhi(t4, tail.rpp_magic); job = tail.first_job + jq; mbc = (t4 - jq * tail.rpp) * 4u; cnt = 4u; }
        const DevJob& J = jobs[job];
        const uint32_t n_mb = J.n_mb;
        if (J.n_intra != n_mb) {   // an I-picture has nothing for this kernel
            // ---- lane-parallel set-up: lanes l and l + 16 look after macroblock mbc + l ----
            if (prefetch_on) {
                const int fy = (int)__umulhi(mbc, mbw_magic);
                prefetch_chunk_region<LOG2S>(J.ref[0], H, mbw, (int)mbc - fy * mbw, fy, (int)min(cnt, n_mb - mbc), lane);
            }
            const bool in = (uint32_t)k16 < cnt && mbc + k16 < n_mb;
            const uint32_t mbk = in ? mbc + k16 : n_mb - 1;
            const uint4 d = __ldg(reinterpret_cast<const uint4*>(J.mbs + mbk));
            const uint32_t* const coefs = reinterpret_cast<const uint32_t*>(J.coefs);
            const uint32_t* const qtab = J.hdr->qtab;
            if (lane < 20) asm volatile("prefetch.global.L1 [%0];" :: "l"(qtab + lane * 4));   // the picture's 80 dequantisation words
            uint8_t* const dst = J.dst;
            const bool inter = in && !(d.x & 3u);
            const uint32_t n_parts_k = (d.x >> 2) & 127u;
            const uint32_t n_coef = inter ? (d.x >> 9) & 511u : 0u;
            const uint32_t bm = inter ? (d.x >> 18) & 63u : 0u;
            const int mby = (int)__umulhi(mbk, mbw_magic), mbx = (int)mbk - mby * mbw;
            const int yoff = ((mby * 16) << LOG2S) + mbx * 16;
            // leaves 0 and 1: the inline copy of an unsplit macroblock's single leaf (info bit 28), else the records; lanes
            // 16..31 fetch leaves 2 and 3 of the same macroblock for the load-per-lane path
            const bool inl = n_parts_k == 1u && (d.x & (1u << 28));
            uint2 pp0 = make_uint2(0u, 0u), pp1 = make_uint2(0u, 0u);
            if (inter && !inl) {
                const uint32_t q = (uint32_t)(lane >> 4) * 2u;
                const uint2* pr = reinterpret_cast<const uint2*>(J.parts + d.y);
                if (q < n_parts_k) pp0 = __ldg(pr + q);
                if (q + 1u < n_parts_k) pp1 = __ldg(pr + q + 1u);
            }
            int mvx0, mvy0, ref0;
            if (inl) { mvx0 = ((int)(d.w << 18)) >> 18; mvy0 = ((int)(d.w << 4)) >> 18; ref0 = (int)(d.w >> 28); }
            else { const PartV v = part_of(pp0.x, pp0.y); mvx0 = v.mvx; mvy0 = v.mvy; ref0 = v.ref; }
            const PartV v1 = part_of(pp1.x, pp1.y);
            // boxes only when every column the windows need lies inside its own pixel row (flat addressing wraps, TMA zero-fills)
            const int x00 = mbx * 16 + (mvx0 >> 1), cx00 = mbx * 8 + (mvx0 >> 2);
            const int x01 = mbx * 16 + (v1.mvx >> 1), cx01 = mbx * 8 + (v1.mvx >> 2);
            const bool row0 = x00 >= 0 && x00 + 17 <= S && cx00 >= 0 && cx00 + 9 <= (S >> 1);
            const bool row1 = x01 >= 0 && x01 + 17 <= S && cx01 >= 0 && cx01 + 9 <= (S >> 1);
            uint32_t need = 0;   // box slots wanted: 1 unsplit, 2 split once (16x8 + 16x8 or 8x16 + 8x16: each leaf fetches the whole window at its vector)
            if (inter && lane < 16) {
                if (n_parts_k == 1u && row0) need = 1;
                else if (n_parts_k == 2u && row0 && row1) need = 2;
            }
            // first slot: running sum over the four macroblocks of the run; who does not fit takes the load-per-lane path
            uint32_t incl = need;
            { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, 1); if (k4 >= 1) incl += u; }
            { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, 2); if (k4 >= 2) incl += u; }
            const uint32_t slot0 = incl - need;
            if (incl > (uint32_t)RUN) need = 0;   // (later macroblocks of the run keep the slots they counted on)
            // an unsplit macroblock that goes down the load-per-lane path after all: its leaf record, rebuilt from the inline copy
            if (inl && need == 0u) pp0 = make_uint2((0xFu | (uint32_t)ref0 << 4) << 8 | (uint32_t)mvx0 << 16, (uint32_t)mvy0 & 0xFFFFu);
            const uint32_t mcw = leaf_word(x00, cx00, mvx0, mvy0) | leaf_word(x01, cx01, v1.mvx, v1.mvy) << 12 | (inter ? MCW_INTER : 0u)
                               | (need ? MCW_BOX : 0u) | (need == 2u ? MCW_TWO : 0u) | (((pp0.x | pp1.x) & 15u) ? MCW_LR : 0u) | ((pp0.x & 255u) ? MCW_SWAP : 0u) | (slot0 & 3u) << 28;
            int pic0 = 0, pic1 = 0;
            if (need) pic0 = (int)J.ref_pic[ref0 - 1];
            if (need == 2u) pic1 = (int)J.ref_pic[v1.ref - 1];

#pragma unroll 1
            for (int r = 0; r < CH_MBS / RUN; r++) {
                if ((uint32_t)(RUN * r) >= cnt || mbc + (uint32_t)(RUN * r) >= n_mb) break;
                const int l0 = RUN * r;              // lanes l0 .. l0+3 hold this run's macroblocks
                const bool mine = (k16 >> 2) == r;
                // everyone is done with the previous run's shared memory; its generic-proxy traffic is ordered before the boxes
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane >= l0 && lane < l0 + RUN && need) {
                    const uint32_t bar = smem_u32(&sm.bar[k4]);
                    const uint32_t s0 = (mcw >> 28) & 3u;
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(need * (TMA_BYTES_L + TMA_BYTES_C4)) : "memory");
                    tma_load_3d(smem_u32(raw + s0 * 640), &tm_l, x00 & ~15, mby * 16 + (mvy0 >> 1), pic0, bar);
                    tma_load_4d(smem_u32(raw + RUN * 640 + s0 * 640), &tm_c4, cx00 & ~15, 0, H + mby * 8 + (mvy0 >> 2), pic0, bar);
                    if (need == 2u) {
                        tma_load_3d(smem_u32(raw + (s0 + 1u) * 640), &tm_l, x01 & ~15, mby * 16 + (v1.mvy >> 1), pic1, bar);
                        tma_load_4d(smem_u32(raw + RUN * 640 + (s0 + 1u) * 640), &tm_c4, cx01 & ~15, 0, H + mby * 8 + (v1.mvy >> 2), pic1, bar);
                    }
                }
                // the run's coefficient records: one range of the picture's array; the first 64 are fetched now
                const uint32_t jmin = __reduce_min_sync(0xffffffffu, mine && n_coef ? d.z : 0xffffffffu);
                const uint32_t jmax = __reduce_max_sync(0xffffffffu, mine && n_coef ? d.z + n_coef : 0u);
                const uint32_t ntot = jmax > jmin ? jmax - jmin : 0u;
                const uint32_t* cf = coefs + (ntot ? jmin : 0u);
                uint32_t ca = 0, cb = 0;
                if ((uint32_t)lane < ntot) ca = __ldg(cf + lane);
                if ((uint32_t)lane + 32u < ntot) cb = __ldg(cf + 32 + lane);
                const uint32_t bmp = __reduce_or_sync(0xffffffffu, mine ? bm << (8 * k4) : 0u);   // the run's coded-block masks, one byte each

                // ---- prediction, one macroblock at a time, into its tile ----
#pragma unroll 1
                for (int i = 0; i < RUN; i++) {
                    const uint32_t w = __shfl_sync(0xffffffffu, mcw, l0 + i);
                    if (!(w & MCW_INTER)) continue;   // intra (k_intra's job) or past the picture's last macroblock
                    uint32_t y0, y1, c0;
                    if (w & MCW_BOX) {
                        const uint32_t bar = smem_u32(&sm.bar[i]), par = (phases >> i) & 1u;
                        phases ^= 1u << i;
                        uint32_t done, spins = 0;
                        do {
                            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar), "r"(par) : "memory");
                            if (!done && ++spins > (1u << 22)) __trap();   // a box that never arrives must not hang the device
                        } while (!done);
                        // which leaf covers this lane's pixels: split top/bottom -> by row, left/right -> by half (luma 8, chroma 4 pixels per lane)
                        uint32_t sl = 0, sc = 0;
                        if (w & MCW_TWO) {   // leaf 0 is the top / left half unless the records come in the other order
                            const uint32_t sw = (w >> 30) & 1u;
                            sl = ((w & MCW_LR) ? (uint32_t)lhalf : (uint32_t)(lrow >> 3)) ^ sw;
                            sc = ((w & MCW_LR) ? (uint32_t)lhalf : (uint32_t)(crow >> 2)) ^ sw;
                        }
                        const uint32_t s0 = (w >> 28) & 3u;
                        const uint32_t wl = sl ? w >> 12 : w, wc = sc ? w >> 12 : w;
                        tile_row8(raw + (s0 + sl) * 640, (uint32_t)(lrow * 32 + lhalf * 8) + (wl & 15u), (int)((wl >> 4) & 3u), y0, y1);
                        c0 = tile_row4_p64(raw + RUN * 640 + (s0 + sc) * 640, (uint32_t)((crow * 2 + cpl) * 32 + lhalf * 4) + ((wc >> 6) & 15u), (int)((wc >> 10) & 3u));
                    } else {
                        // load-per-lane path (CopyBlock MD:418-456 on the flat planes)
                        const uint32_t dx = __shfl_sync(0xffffffffu, d.x, l0 + i), dy = __shfl_sync(0xffffffffu, d.y, l0 + i);
                        const int yo = __shfl_sync(0xffffffffu, yoff, l0 + i);
                        const int n_parts = (int)((dx >> 2) & 127u);
                        uint2* sp = reinterpret_cast<uint2*>(raw + Smem::BOX);
                        uint8_t* map = raw + Smem::BOX + 512;
                        // leaves 0..3 were fetched with the descriptors (lanes l0+i and 16+l0+i hold them), the rest is fetched now
                        if (k16 == l0 + i) { sp[(lane >> 4) * 2] = pp0; sp[(lane >> 4) * 2 + 1] = pp1; }
                        const mobi_part* parts = J.parts + dy;
                        for (int p = 4 + lane; p < n_parts; p += 32) sp[p] = __ldg(reinterpret_cast<const uint2*>(parts + p));
                        __syncwarp();
                        uint32_t ml = 0, mc = 0;
                        if (n_parts > 1) {
                            // partition map at 2x2-pixel granularity (leaves go down to 2x2, MD:1726): lane l owns cells 2l, 2l+1 of the 8x8 grid
                            const int cy2 = lane >> 2, cx2 = (lane & 3) * 2;
                            uint32_t i0 = 0, i1 = 0;
                            for (int p = 0; p < n_parts; p++) {
                                const uint32_t pw = sp[p].x;
                                const int x2 = pw & 15, y2 = (pw >> 4) & 15, cw = 1 << ((pw >> 8) & 3), ch = 1 << ((pw >> 10) & 3);
                                const bool rowin = (unsigned)(cy2 - y2) < (unsigned)ch;
                                if (rowin && (unsigned)(cx2 - x2) < (unsigned)cw) i0 = (uint32_t)p;
                                if (rowin && (unsigned)(cx2 + 1 - x2) < (unsigned)cw) i1 = (uint32_t)p;
                            }
                            reinterpret_cast<uint16_t*>(map)[lane] = (uint16_t)(i0 | i1 << 8);
                            __syncwarp();
                            ml = *reinterpret_cast<const uint32_t*>(map + (lrow >> 1) * 8 + lhalf * 4);
                            mc = *reinterpret_cast<const uint32_t*>(map + crow * 8 + lhalf * 4);
                        }
                        const int ypix = yo + (lrow << LOG2S) + lhalf * 8;
                        const int cpix = (yo >> 1) + (cpl ? (S >> 1) : 0) + (crow << LOG2S) + lhalf * 4;
                        auto leaf = [&](uint32_t idx) { const uint2 pw = sp[idx]; return part_of(pw.x, pw.y); };
                        if (ml == (ml & 255u) * 0x01010101u) {
                            const PartV p = leaf(ml & 255u);
                            mc_row8(J.ref[p.ref - 1] + ypix + ((p.mvy >> 1) << LOG2S) + (p.mvx >> 1), S, (p.mvx & 1) | ((p.mvy & 1) << 1), y0, y1);
                        } else {
                            uint32_t o[2] = {0, 0};
#pragma unroll
                            for (int c = 0; c < 4; c++) {
                                const PartV p = leaf((ml >> (8 * c)) & 255u);
                                const uint8_t* s = J.ref[p.ref - 1] + ypix + 2 * c + ((p.mvy >> 1) << LOG2S) + (p.mvx >> 1);
                                const int ph = (p.mvx & 1) | ((p.mvy & 1) << 1);
                                const uint32_t v = mc_px(s, S, ph) | mc_px(s + 1, S, ph) << 8;
                                o[c >> 1] |= v << (16 * (c & 1));
                            }
                            y0 = o[0]; y1 = o[1];
                        }
                        if (mc == (mc & 255u) * 0x01010101u) {
                            const PartV p = leaf(mc & 255u);
                            const int cx = p.mvx >> 1, cy = p.mvy >> 1;
                            c0 = mc_row4(J.ref[p.ref - 1] + ysz + cpix + ((cy >> 1) << LOG2S) + (cx >> 1), S, (cx & 1) | ((cy & 1) << 1));
                        } else {
                            c0 = 0;
#pragma unroll
                            for (int c = 0; c < 4; c++) {
                                const PartV p = leaf((mc >> (8 * c)) & 255u);
                                const int cx = p.mvx >> 1, cy = p.mvy >> 1;
                                c0 |= mc_px(J.ref[p.ref - 1] + ysz + cpix + c + ((cy >> 1) << LOG2S) + (cx >> 1), S, (cx & 1) | ((cy & 1) << 1)) << (8 * c);
                            }
                        }
                        __syncwarp();   // everyone is done with the leaf records before the next split macroblock replaces them
                    }
                    uint8_t* tl = sm.tile[i];
                    *reinterpret_cast<uint2*>(tl + lrow * 16 + lhalf * 8) = make_uint2(y0, y1);
                    *reinterpret_cast<uint32_t*>(tl + 256 + cpl * 64 + crow * 8 + lhalf * 4) = c0;
                }
                __syncwarp();   // tiles complete; the boxes are dead, the coefficient pool may overwrite them

                if (ntot) {
                    // ---- pool slots: macroblock i's coded blocks take slots sb_i .. sb_i + popc(mask_i) - 1, lowest block first ----
                    uint32_t sbp = 0, ns = 0;
#pragma unroll
                    for (int i = 0; i < RUN; i++) { sbp |= ns << (8 * i); ns += __popc((bmp >> (8 * i)) & 63u); }
                    {
                        int4* z = reinterpret_cast<int4*>(&sm.u.coef[0][0]);
                        for (uint32_t q = lane; q < ns * 16u; q += 32u) z[q] = make_int4(0, 0, 0, 0);
                    }
                    __syncwarp();
                    // ---- dequantise into the pool (MD:3424-3429) ----
                    uint32_t m8 = 0;
                    for (uint32_t j0 = 0; j0 < ntot; j0 += 32u) {
                        const uint32_t j = j0 + (uint32_t)lane;
                        if (j < ntot) {
                            const uint32_t c = j0 == 0 ? ca : j0 == 32u ? cb : __ldg(cf + j);
                            // whose record: the parser tags every record with its macroblock's index & 3 (mobi_coef.blk bits 3-4), and
                            // a run is four macroblocks aligned to four.  Records of intra macroblocks lying inside the range (k_intra's),
                            // or naming a block their macroblock does not code, find no bit in the mask and are passed over.
                            const uint32_t own = (c >> 27) & 3u, blk = (c >> 24) & 7u;
                            const uint32_t mask = (bmp >> (8 * own)) & 63u;
                            if ((mask >> blk) & 1u) {
                                const int level = (int)(int16_t)(c & 0xFFFFu);
                                const uint32_t pos = (c >> 16) & 63u, sub = (c >> 22) & 3u, is8 = c >> 31;
                                const uint32_t wq = __ldg(qtab + (is8 ? pos : 64u + (pos & 15u)));
                                const uint32_t slot = ((sbp >> (8 * own)) & 255u) + __popc(mask & ((1u << blk) - 1u));
                                const uint32_t e = is8 ? (wq & 63u) : sub * 16u + (wq & 15u);
                                sm.u.coef[slot][pool_word(slot, e >> 3, e & 7u)] = (int)(wq >> 8) * level;
                                m8 |= is8 << (8 * own + blk);
                            }
                        }
                    }
                    m8 = __reduce_or_sync(0xffffffffu, m8);
                    {
                        // One lane per (macroblock, block) of the run.  The passes visit the blocks transformed as one 8x8 first and
                        // the ones transformed as four 4x4 after them, so that a pass rarely has to run both butterflies (60 % / 40 %
                        // on the bench mix: unsorted, five passes in six are mixed); slotinfo[] is indexed by that visiting position
                        // and carries the pool slot.
                        const uint32_t kk = (uint32_t)lane / 6u, b = (uint32_t)lane - 6u * kk;
                        const uint32_t mask = lane < 6 * RUN ? (bmp >> (8 * kk)) & 63u : 0u;
                        const bool coded = (mask >> b) & 1u;
                        const bool t8 = coded && ((m8 >> (8 * kk + b)) & 1u);
                        const uint32_t cm = __ballot_sync(0xffffffffu, coded), m8b = __ballot_sync(0xffffffffu, t8);
                        if (coded) {
                            const uint32_t lt = (1u << lane) - 1u;
                            const uint32_t pos = t8 ? __popc(m8b & lt) : __popc(m8b) + __popc(cm & ~m8b & lt);
                            const uint32_t slot = ((sbp >> (8 * kk)) & 255u) + __popc(mask & ((1u << b) - 1u));
                            const uint32_t toff = kk * (uint32_t)Smem::TILE + (b < 4u ? ((b >> 1) * 8u) * 16u + (b & 1u) * 8u : 256u + (b - 4u) * 64u);
                            sm.slotinfo[pos] = toff | (b < 4u ? 0u : 1u << 16) | (t8 ? 1u << 17 : 0u) | slot << 18;
                        }
                    }
                    __syncwarp();

                    // ---- inverse transforms: eight lanes per pooled block (one row each), four blocks per pass ----
                    const int g = lane >> 3, rr = lane & 7, i4 = rr & 3, s0 = (rr >> 2) * 2;
                    for (uint32_t base = 0; base < ns; base += 4u) {
                        const bool has = base + (uint32_t)g < ns;
                        const uint32_t info = sm.slotinfo[has ? base + (uint32_t)g : 0u];
                        const bool is8 = (info >> 17) & 1u;
                        const uint32_t slot = (info >> 18) & 31u;
                        int32_t* B = sm.u.coef[slot];
                        // this lane's two 4-word groups: row rr of an 8x8 block, or row i4 of 4x4 units s0 and s0+1 (elements
                        // s0*16 + 4*i4 .. and (s0+1)*16 + 4*i4 ..: block rows 2*s0 + (i4>>1) and 2*s0 + 2 + (i4>>1), columns 4*(i4&1) ..)
                        const uint32_t rlo = is8 ? (uint32_t)rr : (uint32_t)(2 * s0 + (i4 >> 1)), rhi = is8 ? (uint32_t)rr : rlo + 2u;
                        const uint32_t clo = is8 ? 0u : (uint32_t)(4 * (i4 & 1)), chi = is8 ? 4u : clo;
                        const uint32_t wlo = pool_word(slot, rlo, clo), whi = pool_word(slot, rhi, chi);
                        int32_t in[8], v[8];
                        if (has) {
                            const int4 lo = *reinterpret_cast<const int4*>(B + wlo);
                            const int4 hi = *reinterpret_cast<const int4*>(B + whi);
                            in[0] = lo.x; in[1] = lo.y; in[2] = lo.z; in[3] = lo.w; in[4] = hi.x; in[5] = hi.y; in[6] = hi.z; in[7] = hi.w;
                            if (is8) { if (rr == 0) in[0] += 32; bfly8(in, v); }
                            else { if (i4 == 0) { in[0] += 32; in[4] += 32; } bfly4(in, v); bfly4(in + 4, v + 4); }
                        }
                        __syncwarp();
                        if (has) {
                            if (is8) {
#pragma unroll
                                for (int q = 0; q < 8; q++) B[pool_word(slot, (uint32_t)q, (uint32_t)rr)] = v[q];   // element 8*q + rr
                            } else {
#pragma unroll
                                for (int q = 0; q < 4; q++) {   // elements s0*16 + 4*q + i4 and (s0+1)*16 + 4*q + i4
                                    const uint32_t e0 = (uint32_t)(s0 * 16 + 4 * q + i4);
                                    B[pool_word(slot, e0 >> 3, e0 & 7u)] = v[q];
                                    B[pool_word(slot, (e0 >> 3) + 2u, e0 & 7u)] = v[4 + q];
                                }
                            }
                        }
                        __syncwarp();
                        if (has) {
                            const int4 lo = *reinterpret_cast<const int4*>(B + wlo);
                            const int4 hi = *reinterpret_cast<const int4*>(B + whi);
                            in[0] = lo.x; in[1] = lo.y; in[2] = lo.z; in[3] = lo.w; in[4] = hi.x; in[5] = hi.y; in[6] = hi.z; in[7] = hi.w;
                            if (is8) bfly8(in, v); else { bfly4(in, v); bfly4(in + 4, v + 4); }
                            // either way the lane now holds the residuals of row rr, columns 0..7 of its block: add onto the prediction
                            uint8_t* tp = &sm.tile[0][0] + (info & 0xFFFFu) + rr * ((info >> 16) & 1u ? 8 : 16);
                            uint2 px = *reinterpret_cast<uint2*>(tp);
                            px.x = addsat4(px.x, v[0], v[1], v[2], v[3]);
                            px.y = addsat4(px.y, v[4], v[5], v[6], v[7]);
                            *reinterpret_cast<uint2*>(tp) = px;
                        }
                        __syncwarp();
                    }
                }

                // ---- the tiles leave, lane-parallel over the run again: 16 luma bytes / 8 chroma bytes per lane and round ----
                const int yoff_k = __shfl_sync(0xffffffffu, yoff, l0 + k4);
                const bool inter_k = (__shfl_sync(0xffffffffu, mcw, l0 + k4) & MCW_INTER) != 0u;
                if (inter_k) {
                    const uint8_t* tl = sm.tile[k4];
                    uint8_t* const py = dst + yoff_k;
                    uint8_t* const pc = dst + ysz + (yoff_k >> 1);
#pragma unroll
                    for (int row = lane / RUN; row < 16; row += 32 / RUN)
                        *reinterpret_cast<uint4*>(py + (row << LOG2S)) = *reinterpret_cast<const uint4*>(tl + row * 16);
#pragma unroll
                    for (int q = lane / RUN; q < 16; q += 32 / RUN)
                        *reinterpret_cast<uint2*>(pc + (q >> 3) * (S >> 1) + ((q & 7) << LOG2S)) = *reinterpret_cast<const uint2*>(tl + 256 + q * 8);
                }
            }
        }
        t = __shfl_sync(0xffffffffu, t_next, 0) - ticket_base;
    }
}

#include "mobi_inter_v3.cuh"
#include "mobi_inter_split.cuh"

// ------------------------------------------------------------------------------------------------
// intra macroblocks
// ------------------------------------------------------------------------------------------------
// Per-warp shared state.  Pixel tiles: luma rows y-1..y+15, columns x-4..x+27 (column x at index 4, so x-1 = 3 and
// the top-right / right-hand run x+16..x+20 = 20..24); chroma rows c-1..c+7, column c at index 4.
// The residuals of ALL of a macroblock's transform units depend on its coefficient records only, so they are computed
// before the macroblock waits for its neighbours (dequantise, transform, >> 6, all coded 8x8 blocks side by side like the
// inter kernel does) and parked in resid[]; what remains on the wavefront's critical path per block is predict + add.
struct IntraSmem {
    union {
        int32_t coef[6][64];   // coefficient blocks by 8x8 block id (0-3 luma raster, 4 U, 5 V); dead once resid[] is filled
        struct { uint8_t y[17][32]; uint8_t c[2][9][16]; } t;   // the pixel tiles reuse the space
    } u;
    int16_t resid[384];    // residuals by pixel position (luma 16x16, U 8x8, V 8x8), saturated to 16 bits (exact under the final clip)
    uint32_t qtab[80];     // the picture's dequantisation words (MD:3897-3912)
};

// Is the 4-byte word at flat luma address `flat` (a) inside the array, (b) inside the visible picture, (c) part of a
// macroblock that precedes m in decode order?  Otherwise the reference reads 0 there: fresh planes (MD:107-108).
// MB and picture edges are multiples of 16, so the answer is the same for all four bytes of an aligned word.
__device__ __forceinline__ uint32_t nb_luma4(const DevJob& J, const Geom& g, int flat, uint32_t m) {
    if (flat < 0 || flat >= g.S * g.H) return 0;
    const int row = flat >> g.log2S, col = flat & (g.S - 1);
    if (col >= g.W) return 0;
    if ((uint32_t)((row >> 4) * g.mbw + (col >> 4)) >= m) return 0;
    return __ldcg(reinterpret_cast<const uint32_t*>(J.dst + flat));
}
__device__ __forceinline__ uint32_t nb_chroma4(const DevJob& J, const Geom& g, int flat, uint32_t m) {
    if (flat < 0 || flat >= (g.S * g.H >> 1)) return 0;
    const int row = flat >> g.log2S, col = flat & (g.S - 1);
    const int pc = col < (g.S >> 1) ? col : col - (g.S >> 1);
    if (pc >= (g.W >> 1)) return 0;
    if ((uint32_t)((row >> 3) * g.mbw + (pc >> 3)) >= m) return 0;
    return __ldcg(reinterpret_cast<const uint32_t*>(J.dst + (size_t)g.S * g.H + flat));
}

// Value of pixel (x,y) of an NxN block under predictor `mode` (0,1,4..8: MD:1890-2472 / 2475-2769 closed forms).
// t points at the block's top-left pixel inside a shared tile with row pitch ts.  All reads are outside the block.
template <int N>
__device__ __forceinline__ int dir_px(const uint8_t* t, int ts, int mode, int x, int y) {
#define TT(k) ((int)t[-ts + (k)])
#define LL(k) ((int)t[(k) * ts - 1])
    switch (mode) {
    case 0: return TT(x);
    case 1: return LL(y);
    case 4: {
        const int z = x + 2 * y, k = y + (x >> 1);
        if (z > 2 * N - 3) return LL(N - 1);
        if (z == 2 * N - 3) return (LL(N - 2) + 3 * LL(N - 1) + 2) >> 2;
        if (z & 1) return (LL(k) + 2 * LL(k + 1) + LL(k + 2) + 2) >> 2;
        return (LL(k) + LL(k + 1) + 1) >> 1; }
    case 5: {
        const int z = 2 * y - x, k = y - (x >> 1);
        if (z < -1) return (TT(x - 2 * y - 1) + 2 * TT(x - 2 * y - 2) + TT(x - 2 * y - 3) + 2) >> 2;
        if (z == -1) return (LL(0) + 2 * TT(-1) + TT(0) + 2) >> 2;
        if (z & 1) return (LL(k - 2) + 2 * LL(k - 1) + LL(k) + 2) >> 2;   // LL(-1) == TT(-1) == top-left
        return (LL(k - 1) + LL(k) + 1) >> 1; }
    case 6: {
        const int z = 2 * x - y, k = x - (y >> 1);
        if (z < -1) return (LL(y - 2 * x - 1) + 2 * LL(y - 2 * x - 2) + LL(y - 2 * x - 3) + 2) >> 2;
        if (z == -1) return (LL(0) + 2 * TT(-1) + TT(0) + 2) >> 2;
        if (z & 1) return (TT(k - 2) + 2 * TT(k - 1) + TT(k) + 2) >> 2;
        return (TT(k - 1) + TT(k) + 1) >> 1; }
    case 7:
        if (x > y) return (TT(x - y - 2) + 2 * TT(x - y - 1) + TT(x - y) + 2) >> 2;
        if (x < y) return (LL(y - x - 2) + 2 * LL(y - x - 1) + LL(y - x) + 2) >> 2;
        return (TT(0) + 2 * TT(-1) + LL(0) + 2) >> 2;
    case 8: {
        const int k = x + (y >> 1);
        if (y & 1) return (TT(k) + 2 * TT(k + 1) + TT(k + 2) + 2) >> 2;
        return (TT(k) + TT(k + 1) + 1) >> 1; }
    }
    return 0;
#undef TT
#undef LL
}
// The diagonal predictors (modes 4-8) without per-lane case analysis.  Every one of them is a two- or three-tap filter sliding
// along ONE line of neighbours: the left column from the bottom up, the top-left pixel, the top row (and what lies right of
// it): E(c) = L(-1 - c) for c < 0, TL for c = 0, T(c - 1) for c > 0.  dir_px's cases are then
//   three taps  F3(c) = (E(c-1) + 2 E(c) + E(c+1) + 2) >> 2        two taps  F2(c) = (E(c) + E(c+1) + 1) >> 1
//   mode 7  F3(x - y)
//   mode 5  z = 2y - x, k = y - (x >> 1):  z < 0: F3(-z - 1);  z odd: F3(-k);  z even: F2(-k - 1)
//   mode 6  z = 2x - y, k = x - (y >> 1):  z < 0: F3(z + 1);   z odd: F3(k);   z even: F2(k)
//   mode 8  k = x + (y >> 1):              y odd: F3(k + 2);   y even: F2(k + 1)
//   mode 4  k = y + (x >> 1):              x odd: F3(-2 - k);  x even: F2(-2 - k), with E clamped at its lower end (c >= -N):
//           that clamp IS the reference's (L(N-2) + 3 L(N-1) + 2) >> 2 and its run of L(N-1) (MD:2141-2200).
// A pixel computes its position c on that line and fetches its three taps with three independent loads: no divergent paths,
// which were a quarter of the stall samples of the latency-bound I-picture kernel (branch resolution).  (Holding E one element
// per lane and fetching taps by shuffle was measured too: the load -> shuffle chain is longer.)  dir_px stays as the readable
// statement (and serves 0, 1).
template <int N>
__device__ __forceinline__ int edge_tap(const uint8_t* t, int ts, int c) {   // E(c), clamped at its lower end
    c = max(c, -N);
    return t[c < 0 ? (-1 - c) * ts - 1 : c - ts - 1];
}
template <int N>
__device__ __forceinline__ int edge_px(const uint8_t* t, int ts, int mode, int x, int y) {
    int c; bool two;
    switch (mode) {   // (uniform across the warp)
    case 4: { const int k = y + (x >> 1); c = -2 - k; two = !(x & 1); break; }
    case 5: { const int z = 2 * y - x, k = y - (x >> 1); c = z < 0 ? -z - 1 : ((z & 1) ? -k : -k - 1); two = z >= 0 && !(z & 1); break; }
    case 6: { const int z = 2 * x - y, k = x - (y >> 1); c = z < 0 ? z + 1 : k; two = z >= 0 && !(z & 1); break; }
    case 7: c = x - y; two = false; break;
    default: { const int k = x + (y >> 1); c = (y & 1) ? k + 2 : k + 1; two = !(y & 1); break; }
    }
    const int e0 = edge_tap<N>(t, ts, c - 1), e1 = edge_tap<N>(t, ts, c), e2 = edge_tap<N>(t, ts, c + 1);   // three independent loads
    return two ? (e1 + e2 + 1) >> 1 : (e0 + 2 * e1 + e2 + 2) >> 2;
}
// Unclipped plane-predictor value (sub_1167BC MD:3017 for N=16, sub_116CCC MD:3168 for 8, sub_117E98 MD:3253 for 4),
// the reference's running sums written in closed form (all arithmetic int32, wrap-around like C#).
template <int N>
__device__ __forceinline__ int plane_val(const uint8_t* t, int ts, int delta, int x, int y) {
    const int l = t[(N - 1) * ts - 1], tt = t[-ts + N - 1], T = t[-ts + x], L = t[y * ts - 1];
    const int m = ((l + tt + 1) >> 1) + delta * 2;
    if (N == 16) {
        const int gx = (m - l + 1) >> 1, gy = (m - tt + 1) >> 1;
        const int Bc = (l * 8 + (x + 1) * gx) - T * 8 + 1, A = T * 64 + (y + 1) * (Bc >> 1);
        const int step = (tt * 8 + (y + 1) * gy) - L * 8 + 1, run = L * 64 + (x + 1) * (step >> 1);
        return (A + run + 64) >> 7;
    }
    constexpr int sh = N == 8 ? 3 : 2;
    const int gx = m - l, gy = m - tt;
    const int Bc = ((l << sh) + (x + 1) * gx) - (T << sh), A = (T << (2 * sh)) + (y + 1) * Bc;
    const int step = ((tt << sh) + (y + 1) * gy) - (L << sh), run = (L << (2 * sh)) + (x + 1) * step;
    return N == 8 ? (A + run + 64) >> 7 : (A + run + 16) >> 5;
}

// Predicted pixels + their parked residuals, clipped (the "+ residual, clip" half of loc_116518 / loc_116628 MD:2898-2956).
__device__ __forceinline__ uint32_t add_res4(uint32_t px, const int16_t* rp) {
    const uint2 rr = *reinterpret_cast<const uint2*>(rp);
    const int p0 = (int)(px & 255u) + (int)(int16_t)(rr.x & 0xFFFFu), p1 = (int)((px >> 8) & 255u) + ((int)rr.x >> 16);
    const int p2 = (int)((px >> 16) & 255u) + (int)(int16_t)(rr.y & 0xFFFFu), p3 = (int)(px >> 24) + ((int)rr.y >> 16);
    uint32_t hi, out;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(p3), "r"(p2), "r"(0));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(out) : "r"(p1), "r"(p0), "r"(hi));
    return out;
}
__device__ __forceinline__ uint32_t add_res2(uint32_t px, const int16_t* rp) {
    const uint32_t rr = *reinterpret_cast<const uint32_t*>(rp);
    const int p0 = (int)(px & 255u) + (int)(int16_t)(rr & 0xFFFFu), p1 = (int)((px >> 8) & 255u) + ((int)rr >> 16);
    uint32_t out;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(out) : "r"(p1), "r"(p0), "r"(0));
    return out;
}

// One intra op: predict an NxN block inside a shared tile (t = its top-left pixel, 4-byte aligned) and, when the block is
// coded, add its residual (rp, row pitch rs) in the same pass: the lane that computed a pixel's prediction also holds its
// residual, so a block costs one store and one warp barrier (this is the wavefront's critical path).  Predictors read only
// pixels outside the block, so values are computed and stored without an intermediate barrier.
template <int N>
__device__ __forceinline__ void intra_predict(uint8_t* t, int ts, int mode, int delta, bool left_av, bool top_av, bool res, const int16_t* rp, int rs, int lane) {
    constexpr int WORDS = N * N / 4, WPR = N / 4;  // 4-pixel words in the block / per row
    if (mode == 3) {  // DC by flat-offset availability (MD:1920-2022, 2501-2580)
        unsigned sum = 0;
        if (top_av) {
#pragma unroll
            for (int k = 0; k < N; k += 4) sum = __dp4a(*reinterpret_cast<const uint32_t*>(t - ts + k), 0x01010101u, sum);
        }
        if (left_av) {
#pragma unroll
            for (int k = 0; k < N; k++) sum += t[k * ts - 1];
        }
        unsigned dc;
        if (top_av && left_av) dc = (sum + N) / (2 * N);
        else if (top_av || left_av) dc = (sum + N / 2) / N;
        else dc = 0x80;
        const uint32_t w = (dc & 255u) * 0x01010101u;
        for (int i = lane; i < WORDS; i += 32) {
            const int y = i / WPR, x = (i % WPR) * 4;
            *reinterpret_cast<uint32_t*>(t + y * ts + x) = res ? add_res4(w, rp + y * rs + x) : w;
        }
    } else if (mode == 2) {
        // the reference ORs four unclipped values into one u32 (MD:3064-3074, 3212-3219, 3314-3321): a value outside
        // 0..255 bleeds into the bytes above it, and the top byte's overflow is lost
        for (int i = lane; i < WORDS; i += 32) {
            const int y = i / WPR, x = (i % WPR) * 4;
            const uint32_t w = (uint32_t)plane_val<N>(t, ts, delta, x, y) | (uint32_t)plane_val<N>(t, ts, delta, x + 1, y) << 8 |
                               (uint32_t)plane_val<N>(t, ts, delta, x + 2, y) << 16 | (uint32_t)plane_val<N>(t, ts, delta, x + 3, y) << 24;
            *reinterpret_cast<uint32_t*>(t + y * ts + x) = (N != 16 && res) ? add_res4(w, rp + y * rs + x) : w;
        }
    } else if (N == 8) {
        // directional predictors: every value is a byte (averages of bytes), so the pixels spread over all 32 lanes --
        // two neighbours per lane, one 16-bit store -- instead of four per lane on half the warp
        const int y = lane >> 2, x = (lane & 3) * 2;
        uint32_t w;
        if (mode <= 1) w = (uint32_t)dir_px<N>(t, ts, mode, x, y) | (uint32_t)dir_px<N>(t, ts, mode, x + 1, y) << 8;
        else w = (uint32_t)edge_px<N>(t, ts, mode, x, y) | (uint32_t)edge_px<N>(t, ts, mode, x + 1, y) << 8;
        if (res) w = add_res2(w, rp + y * rs + x);
        *reinterpret_cast<uint16_t*>(t + y * ts + x) = (uint16_t)w;
    } else if (N == 4) {
        if (lane < 16) {
            const int y = lane >> 2, x = lane & 3;
            int v = mode <= 1 ? dir_px<N>(t, ts, mode, x, y) : edge_px<N>(t, ts, mode, x, y);
            if (res) v = clip255(v + (int)rp[y * rs + x]);
            t[y * ts + x] = (uint8_t)v;
        }
    }
    __syncwarp();
}

// Residuals of every transform unit of a macroblock (loc_116540 MD:2931 / sub_1166E8 MD:2958 / ReadDCTMatrix's
// dequantisation MD:3424-3429): coefficient records carry their block / sub-block tags, so all of them are scattered
// at once; eight lanes per coded 8x8 block (one row each, a block coded as four 4x4 units included), four blocks per pass.
__device__ __forceinline__ void intra_residuals(IntraSmem& sm, const uint32_t* __restrict__ cf, int n_coef, uint32_t blkmask, int lane) {
    {
        int4* z = reinterpret_cast<int4*>(&sm.u.coef[0][0]);
#pragma unroll
        for (int i = 0; i < 3; i++) z[lane + 32 * i] = make_int4(0, 0, 0, 0);
    }
    __syncwarp();
    uint32_t m8 = 0;
    for (int j = lane; j < n_coef; j += 32) {
        const uint32_t c = __ldg(cf + j);
        const int level = (int)(int16_t)(c & 0xFFFFu);
        const uint32_t pos = (c >> 16) & 63u, sub = (c >> 22) & 3u, blk = min((c >> 24) & 7u, 5u), is8 = c >> 31;
        const uint32_t w = sm.qtab[is8 ? pos : 64u + (pos & 15u)];
        if ((blkmask >> blk) & 1u) sm.u.coef[blk][is8 ? (w & 63u) : sub * 16u + (w & 15u)] = (int)(w >> 8) * level;   // (the caller may ask for luma or chroma only)
        m8 |= is8 << blk;
    }
    m8 = __reduce_or_sync(0xffffffffu, m8);
    const uint32_t list = c_blklist[blkmask];   // ids of the coded blocks, one nibble each, lowest first
    const int nblk = __popc(blkmask);
    __syncwarp();
    const int g = lane >> 3, r = lane & 7, i4 = r & 3, s0 = (r >> 2) * 2;
    for (int base = 0; base < nblk; base += 4) {
        const bool has = base + g < nblk;
        const int b = (int)((list >> (4 * (base + g))) & 7u);
        const bool is8 = (m8 >> b) & 1u;
        int32_t* B = sm.u.coef[has ? b : 0];
        int32_t in[8], v[8];
        if (has) {
            const int4 lo = *reinterpret_cast<const int4*>(B + (is8 ? 8 * r : s0 * 16 + 4 * i4));
            const int4 hi = *reinterpret_cast<const int4*>(B + (is8 ? 8 * r + 4 : (s0 + 1) * 16 + 4 * i4));
            in[0] = lo.x; in[1] = lo.y; in[2] = lo.z; in[3] = lo.w; in[4] = hi.x; in[5] = hi.y; in[6] = hi.z; in[7] = hi.w;
            if (is8) { if (r == 0) in[0] += 32; bfly8(in, v); }
            else { if (i4 == 0) { in[0] += 32; in[4] += 32; } bfly4(in, v); bfly4(in + 4, v + 4); }
        }
        __syncwarp();
        if (has) {
            if (is8) {
#pragma unroll
                for (int k = 0; k < 8; k++) B[8 * k + r] = v[k];
            } else {
#pragma unroll
                for (int k = 0; k < 4; k++) { B[s0 * 16 + 4 * k + i4] = v[k]; B[(s0 + 1) * 16 + 4 * k + i4] = v[4 + k]; }
            }
        }
        __syncwarp();
        if (has) {
            const int4 lo = *reinterpret_cast<const int4*>(B + (is8 ? 8 * r : s0 * 16 + 4 * i4));
            const int4 hi = *reinterpret_cast<const int4*>(B + (is8 ? 8 * r + 4 : (s0 + 1) * 16 + 4 * i4));
            in[0] = lo.x; in[1] = lo.y; in[2] = lo.z; in[3] = lo.w; in[4] = hi.x; in[5] = hi.y; in[6] = hi.z; in[7] = hi.w;
            if (is8) bfly8(in, v); else { bfly4(in, v); bfly4(in + 4, v + 4); }
            // either way the lane now holds row r, columns 0..7 of block b
            uint4 o;
            asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(o.x) : "r"(v[1] >> 6), "r"(v[0] >> 6));
            asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(o.y) : "r"(v[3] >> 6), "r"(v[2] >> 6));
            asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(o.z) : "r"(v[5] >> 6), "r"(v[4] >> 6));
            asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(o.w) : "r"(v[7] >> 6), "r"(v[6] >> 6));
            int16_t* rp = b < 4 ? sm.resid + ((b >> 1) * 8 + r) * 16 + (b & 1) * 8 : sm.resid + 256 + (b - 4) * 64 + r * 8;
            *reinterpret_cast<uint4*>(rp) = o;
        }
    }
    __syncwarp();
}
// A coded block without a predictor (modes 9 / 19): the residual lands on whatever the tile holds.
template <int N>
__device__ __forceinline__ void add_resid(uint8_t* t, int ts, const int16_t* rp, int rpitch, int lane) {
    constexpr int WORDS = N * N / 4, WPR = N / 4;
    if (lane < WORDS) {
        const int y = lane / WPR, x = (lane % WPR) * 4;
        uint32_t* pw = reinterpret_cast<uint32_t*>(t + y * ts + x);
        *pw = add_res4(*pw, rp + y * rpitch + x);
    }
    __syncwarp();
}

// One intra macroblock, in two halves around the wait for its neighbours.
struct IntraItem { uint32_t job, m, info, first_op, first_coef, wait; };
__device__ __forceinline__ IntraItem load_item(const IntraWork* w) {
    const uint4 w0 = __ldg(reinterpret_cast<const uint4*>(w));
    const uint2 w1 = __ldg(reinterpret_cast<const uint2*>(w) + 2);
    return IntraItem{w0.x, w0.y, w0.z, w0.w, w1.x, w1.y};
}
// Everything that does not depend on neighbouring macroblocks: the op list (returned, one op per lane), the picture's
// scale table, the residuals of all transform units, zeroed tiles.
// PLANES: 1 = luma only, 2 = chroma only, 3 = the whole macroblock (the I-picture kernel runs a luma and a chroma
// wavefront side by side: neither plane's predictors ever read the other plane).
template <bool LOAD_QTAB>
__device__ __forceinline__ uint32_t intra_prefetch(const DevJob& J, IntraSmem& sm, const IntraItem& it, int lane, const int PLANES) {
    const int n_ops = (int)((it.info >> 2) & 127u), n_coef = (int)((it.info >> 9) & 511u);
    const uint32_t blkmask = ((it.info >> 18) & 63u) & (PLANES == 1 ? 0x0Fu : PLANES == 2 ? 0x30u : 0x3Fu);
    const uint32_t myop = lane < n_ops ? __ldg(J.ops + it.first_op + lane) : 0u;
    if (LOAD_QTAB) {
        const uint32_t* qt = J.hdr->qtab;
        for (int i = lane; i < 80; i += 32) sm.qtab[i] = __ldg(qt + i);
    }
    __syncwarp();   // the previous macroblock's tiles have been written out; the scale table is in place
    if (n_coef && blkmask) intra_residuals(sm, reinterpret_cast<const uint32_t*>(J.coefs) + it.first_coef, n_coef, blkmask, lane);
    uint4* z = reinterpret_cast<uint4*>(&sm.u.t.y[0][0]);  // y and c tiles are contiguous: 832 B = 52 x 16
    for (int i = lane; i < 52; i += 32) z[i] = make_uint4(0, 0, 0, 0);
    return myop;
}
// Stage the neighbourhood from the picture in global memory (the general form: any macroblock of any picture).
__device__ __forceinline__ void intra_stage_global(const DevJob& J, const Geom& g, IntraSmem& sm, const IntraItem& it, int lane, const int PLANES) {
    const int S = g.S;
    const uint32_t m = it.m;
    const int mbx = (int)(m % (uint32_t)g.mbw), mby = (int)(m / (uint32_t)g.mbw);
    const int yoff = mby * 16 * S + mbx * 16, coff = yoff >> 1;
    // One round of independent aligned word loads.  luma: 7 words of row y-1 (columns x-4..x+23), then columns
    // x-4..x-1 of rows y..y+15; when the picture is as wide as the stride the columns right of the last macroblock
    // wrap onto real pixels of the next row, so the right-hand run x+16..x+23 of rows y..y+15 is staged too (otherwise
    // it is zero: not yet decoded / padding).
    const bool wrap = g.W == S && mbx == g.mbw - 1;
    // slot i -> (destination word in the tiles, flat source address, plane); both of a lane's loads are issued before
    // either is stored, so the staging costs one L2 round trip
    auto slot = [&](int i, uint32_t*& dstw, int& flat, bool& chroma) {
        if (i < 7) { dstw = reinterpret_cast<uint32_t*>(&sm.u.t.y[0][4 * i]); flat = yoff - S - 4 + 4 * i; chroma = false; }
        else if (i < 23) { dstw = reinterpret_cast<uint32_t*>(&sm.u.t.y[i - 6][0]); flat = yoff + (i - 7) * S - 4; chroma = false; }
        else if (i < 45) {
            const int k = i - 23, p = k >= 11 ? 1 : 0, q = k - 11 * p, base = coff + (p ? (S >> 1) : 0);
            chroma = true;
            if (q < 3) { dstw = reinterpret_cast<uint32_t*>(&sm.u.t.c[p][0][4 * q]); flat = base - S - 4 + 4 * q; }
            else { dstw = reinterpret_cast<uint32_t*>(&sm.u.t.c[p][q - 2][0]); flat = base + (q - 3) * S - 4; }
        } else {
            const int k = i - 45, r = k >> 1, h = k & 1;
            dstw = reinterpret_cast<uint32_t*>(&sm.u.t.y[1 + r][20 + 4 * h]); flat = yoff + r * S + 16 + 4 * h; chroma = false;
        }
    };
    const int n = 45 + (wrap ? 32 : 0);
    uint32_t* d0; uint32_t* d1; int f0, f1; bool c0, c1;
    slot(lane, d0, f0, c0);
    const bool two = lane + 32 < n;
    slot(two ? lane + 32 : lane, d1, f1, c1);
    const bool w0 = PLANES == 3 || c0 == (PLANES == 2), w1 = two && (PLANES == 3 || c1 == (PLANES == 2));   // this plane group's slots only
    const uint32_t v0 = !w0 ? 0u : c0 ? nb_chroma4(J, g, f0, m) : nb_luma4(J, g, f0, m);
    const uint32_t v1 = !w1 ? 0u : c1 ? nb_chroma4(J, g, f1, m) : nb_luma4(J, g, f1, m);
    if (w0) *d0 = v0;
    if (w1) *d1 = v1;
    if ((PLANES & 1) && lane + 64 < n) {   // only in the wrap case (luma)
        slot(lane + 64, d0, f0, c0);
        *d0 = nb_luma4(J, g, f0, m);
    }
    __syncwarp();
}
// Run the macroblock's ops in stream order on the staged tiles (fixed block order, each block sees the previous block's
// reconstruction: MD:1759-1880, 2776-2902).  Lane k decodes op k ONCE -- where its block sits in the tiles and in the parked
// residuals, which neighbours exist -- and the loop hands the packed result round by shuffle: the per-op decode was a quarter
// of the instructions on the wavefront's critical path when every lane repeated it for every op.  Luma ops precede chroma
// ops in stream order, so a luma-only / chroma-only caller walks its own range only.
__device__ __forceinline__ void intra_ops(const Geom& g, IntraSmem& sm, const IntraItem& it, uint32_t myop, int lane, const int PLANES, int yoff) {
    const int S = g.S;
    const int n_ops = (int)((it.info >> 2) & 127u);
    const int coff = yoff >> 1;
    uint32_t mya;   // bits 0-10 tile byte offset of the block, 11-19 residual index, 20-24 mode, 25 residual, 26 left, 27 top, 28 chroma
    {
        const uint32_t op = myop;
        const int plane = (int)((op >> 6) & 3u), x4 = (int)((op >> 8) & 3u), y4 = (int)((op >> 10) & 3u);
        uint32_t toff, roff; int off;
        if (plane == 0) { toff = (uint32_t)((1 + y4 * 4) * 32 + 4 + x4 * 4); roff = (uint32_t)(y4 * 4 * 16 + x4 * 4); off = yoff + y4 * 4 * S + x4 * 4; }
        else {
            toff = (uint32_t)(17 * 32 + (plane - 1) * 9 * 16 + (1 + y4 * 4) * 16 + 4 + x4 * 4); roff = (uint32_t)(256 + (plane - 1) * 64 + y4 * 4 * 8 + x4 * 4);
            off = coff + (plane == 2 ? (S >> 1) : 0) + y4 * 4 * S + x4 * 4;
        }
        const bool left_av = ((off - (plane == 2 ? (S >> 1) : 0)) & (S - 1)) != 0;  // MD:1923, VOffsetfix MD:1885
        const bool top_av = off >= S;                                               // MD:1924
        mya = toff | roff << 11 | (op & 31u) << 20 | ((op >> 5) & 1u) << 25 | (left_av ? 1u << 26 : 0u) | (top_av ? 1u << 27 : 0u) | (plane ? 1u << 28 : 0u);
    }
    // luma ops precede chroma ops (the parser emits them in decode order; mobi_packed_validate insists on it for caller-packed frames)
    const int n_luma = __popc(__ballot_sync(0xffffffffu, lane < n_ops && !(mya >> 28)));
    const int k0 = PLANES == 2 ? n_luma : 0, k1 = PLANES == 1 ? n_luma : n_ops;
    uint8_t* const tiles = &sm.u.t.y[0][0];   // y and c tiles are contiguous (17 x 32, then 2 x 9 x 16)
    for (int k = k0; k < k1; k++) {
        const uint32_t a = __shfl_sync(0xffffffffu, mya, k);
        const int delta = (int)(int16_t)(__shfl_sync(0xffffffffu, myop, k) >> 16);
        const int mode = (int)((a >> 20) & 31u);
        const bool res = (a >> 25) & 1u, left_av = (a >> 26) & 1u, top_av = (a >> 27) & 1u, chroma = (a >> 28) & 1u;
        uint8_t* tp = tiles + (a & 2047u);
        const int16_t* rp = sm.resid + ((a >> 11) & 511u);
        const int ts = chroma ? 16 : 32, rs = chroma ? 8 : 16;
        if (mode == 20) intra_predict<16>(tp, ts, 2, delta, left_av, top_av, false, rp, rs, lane);
        else if (mode >= 10) {
            if (mode != 19) intra_predict<4>(tp, ts, mode - 10, delta, left_av, top_av, res, rp, rs, lane);
            else if (res) add_resid<4>(tp, ts, rp, rs, lane);
        } else {
            if (mode != 9) intra_predict<8>(tp, ts, mode, delta, left_av, top_av, res, rp, rs, lane);
            else if (res) add_resid<8>(tp, ts, rp, rs, lane);
        }
    }
}
// Write the macroblock out.
__device__ __forceinline__ void intra_store(const DevJob& J, const Geom& g, IntraSmem& sm, int lane, const int PLANES, int yoff) {
    const int S = g.S;
    const int coff = yoff >> 1;
    if (PLANES & 1) {
        const int lrow = lane >> 1, lhalf = lane & 1;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(&sm.u.t.y[1 + lrow][4 + lhalf * 8]);
        *reinterpret_cast<uint2*>(J.dst + yoff + lrow * S + lhalf * 8) = make_uint2(src[0], src[1]);
    }
    if (PLANES & 2) {
        const int cpl = lane >> 4, crow = (lane >> 1) & 7, chalf = lane & 1;
        const uint32_t cv = *reinterpret_cast<const uint32_t*>(&sm.u.t.c[cpl][1 + crow][4 + chalf * 4]);
        *reinterpret_cast<uint32_t*>(J.dst + (size_t)S * g.H + coff + (cpl ? (S >> 1) : 0) + crow * S + chalf * 4) = cv;
    }
}
__device__ __forceinline__ void intra_reconstruct(const DevJob& J, const Geom& g, IntraSmem& sm, const IntraItem& it, uint32_t myop, int lane, const int PLANES) {
    const int mbx = (int)(it.m % (uint32_t)g.mbw), mby = (int)(it.m / (uint32_t)g.mbw);
    const int yoff = mby * 16 * g.S + mbx * 16;
    intra_stage_global(J, g, sm, it, lane, PLANES);
    intra_ops(g, sm, it, myop, lane, PLANES, yoff);
    intra_store(J, g, sm, lane, PLANES, yoff);
}

// Scattered intra macroblocks (those inside P-pictures): persistent warps draw tickets from a dependency-ordered
// list; completion is published as a per-macroblock stamp in global memory.
__global__ void __launch_bounds__(INTRA_WARPS * 32, 8) k_intra(const DevJob* __restrict__ jobs, const IntraWork* __restrict__ work, uint32_t n_work,
                                                           uint32_t* ticket, uint32_t ticket_base, uint32_t stamp, Geom g) {
    __shared__ __align__(16) IntraSmem s_all[INTRA_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    IntraSmem& sm = s_all[warp];
    for (;;) {
        uint32_t t = 0;
        if (lane == 0) t = atomicAdd(ticket, 1u) - ticket_base;
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= n_work) break;
        const IntraItem it = load_item(work + t);
        const DevJob& J = jobs[it.job];
        const uint32_t myop = intra_prefetch<true>(J, sm, it, lane, 3);
        // Programmatic dependent launch: this grid may start while the inter kernel before it in the stream is still draining
        // (its CTAs free their SMs one by one over the last tenth of its run).  Everything up to here read only the step's
        // uploaded arrays; the pictures are touched below, after the inter kernel has completed and flushed.
        asm volatile("griddepcontrol.wait;" ::: "memory");
        // wait for the intra neighbours whose pixels this macroblock's predictors read (host-computed mask)
        if (lane < 4 && ((it.wait >> lane) & 1u)) {
            const int nb = lane == 0 ? (int)it.m - 1 : (int)it.m - g.mbw - 2 + lane;
            // acquire: the neighbour's pixel stores (released below) are visible once its stamp is
            const uint32_t* f = J.flags + nb;
            uint32_t seen;
            for (;;) {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(f) : "memory");
                if (seen == stamp) break;
                __nanosleep(20);
            }
        }
        __syncwarp();
        intra_reconstruct(J, g, sm, it, myop, lane, 3);
        // release: every lane's pixel stores happen-before the stamp (bar.warp.sync orders the lanes' stores before
        // lane 0's release at gpu scope through cumulativity)
        __syncwarp();
        if (lane == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(J.flags + it.m), "r"(stamp) : "memory");
    }
}

// I-pictures: one CTA per picture, TWO warps per macroblock row -- one walks the row's luma, the other its chroma (no
// predictor reads across planes, so they are independent wavefronts and the per-macroblock chain of dependent steps
// shrinks to the longer of the two) -- macroblocks of a row left to right, rows r, r + KEY_ROWS, ... per warp pair.
// Every neighbour pixel a macroblock's predictors can read was produced inside this CTA, so none of them is fetched from
// the picture: the left column is the warp's own previous macroblock (kept in a register per lane), the row above comes
// from a line buffer in shared memory that each row's warp fills with the bottom pixel line of its macroblocks, and the
// few places where flat addressing (SURVEY.md 8a hazard 2) reaches further when Width == Stride -- pixel line 14 of the
// row above's last macroblock (top-left of column 0), the first eight columns of this row's first macroblock (right of the
// last column) -- are kept beside it.  A macroblock is published (per-row progress counter in shared memory) as soon as
// its line is in the buffer; its global stores follow, off the wavefront's critical path, and nobody reads them back.
// All rows of a picture live in one CTA, so every awaited row is resident: no deadlock.
constexpr int KEY_ROWS = 16, KEY_WARPS = 2 * KEY_ROWS;
struct KeyPic { uint32_t job, work_base; };
struct KeyShared {   // behind the KEY_WARPS per-warp IntraSmem blocks
    uint32_t prog[2][64];          // macroblocks finished per macroblock row (H <= 1024), luma / chroma; accessed with atomics only
    uint32_t tail[2][KEY_ROWS];    // luma: pixel line 14, last four columns of the row's last macroblock; chroma: V line 6, likewise
    uint8_t first8[KEY_ROWS][16][8];   // luma columns 0..7 of the row's first macroblock
    // then: uint8_t line_y[KEY_ROWS][S], line_c[KEY_ROWS][S] (chroma lines as in the picture: U | V)
};
__host__ __device__ inline size_t key_smem_bytes(int S) { return KEY_WARPS * sizeof(IntraSmem) + sizeof(KeyShared) + 2 * (size_t)KEY_ROWS * (size_t)S; }

// Wait until a row's progress counter (shared memory) reaches `need`.  A picture takes well under a millisecond; 2^24 polls of
// >= 20 ns each can only mean a broken dependency table, and a trap is a reported error where a hang would take the GPU along.
__device__ __forceinline__ void key_wait(uint32_t* counter, uint32_t need) {
    uint32_t polls = 0;
    while (atomicAdd(counter, 0u) < need) { __nanosleep(20); if (++polls > (1u << 24)) __trap(); }
}

// PLANES is a run-time value here on purpose: one copy of the (large) intra code serves both wavefronts, and the
// instruction cache is what a handful of latency-bound warps live on.
__device__ __forceinline__ void key_rows(const DevJob& J, const IntraWork* __restrict__ items, const Geom& g, IntraSmem& sm, KeyShared& ks, uint8_t* lines,
                                         int first_row, int lane, const int PLANES) {
    const int mbw = g.mbw, S = g.S, W = g.W;
    const bool full = W == S;                      // flat addressing wraps onto real pixels
    uint32_t* prog = ks.prog[PLANES - 1];
    uint8_t* line_all = lines + (PLANES == 2 ? KEY_ROWS * S : 0);
    for (int row = first_row; row < g.mbh; row += KEY_ROWS) {
        const int slot = row % KEY_ROWS, pslot = (row + KEY_ROWS - 1) % KEY_ROWS;
        uint8_t* my_line = line_all + slot * S;
        const uint8_t* up_line = line_all + pslot * S;
        IntraItem it_next = load_item(items + row * mbw);
        uint32_t leftw = 0;   // luma: lane k < 16 holds columns 12..15 of line k of the previous macroblock; chroma: lanes 0-7 U, 8-15 V, columns 4..7
        for (int x = 0; x < mbw; x++) {
            // the next macroblock's work item travels one iteration ahead, and its ops and coefficient records are pulled
            // into L1 while this one is reconstructed: two dependent trips to memory less on the wavefront's critical path
            const IntraItem it = it_next;
            if (x + 1 < mbw) it_next = load_item(items + row * mbw + x + 1);
            const uint32_t myop = intra_prefetch<false>(J, sm, it, lane, PLANES);
            if (lane == 0 && row > 0 && it.wait) {
                // neighbours by raster index (SURVEY.md 8a hazard 2): left of column 0 = last macroblock of the row above
                // (bit 0), top-left of column 0 = last macroblock two rows up (bit 1); top-right of the last column is
                // this row's first macroblock, which this warp finished long ago.
                uint32_t need1 = 0, need2 = 0;   // progress required of row-1 / row-2
                if (it.wait & 1u) { if (x == 0) need1 = (uint32_t)mbw; }
                if (it.wait & 2u) { if (x == 0) need2 = (uint32_t)mbw; else need1 = max(need1, (uint32_t)x); }
                if (it.wait & 4u) need1 = max(need1, (uint32_t)x + 1u);
                if ((it.wait & 8u) && x + 1 < mbw) need1 = max(need1, (uint32_t)x + 2u);
                key_wait(&prog[row - 1], need1);
                if (need2 && row > 1) key_wait(&prog[row - 2], need2);
                __threadfence_block();
            }
            __syncwarp();
            if (x + 1 < mbw) {
                const uint8_t* c1 = reinterpret_cast<const uint8_t*>(J.coefs + it_next.first_coef);
                const uint32_t bytes = ((it_next.info >> 9) & 511u) * 4u + (uint32_t)(reinterpret_cast<uintptr_t>(c1) & 127u);
                if ((uint32_t)lane * 128u < bytes) asm volatile("prefetch.global.L1 [%0];" :: "l"(c1 + lane * 128));
                if (lane == 31) asm volatile("prefetch.global.L1 [%0];" :: "l"(J.ops + it_next.first_op));
            }
            // ---- stage the neighbourhood from shared memory (what nb_luma4 / nb_chroma4 would return, resolved in place) ----
            if (PLANES == 1) {
                if (lane < 16) {          // left column, lines 0..15: columns x-4..x-1
                    uint32_t v = leftw;
                    if (x == 0) v = (full && lane == 0 && row > 0) ? *reinterpret_cast<const uint32_t*>(up_line + S - 4) : 0u;   // flat address - 4 = the line above, last columns
                    *reinterpret_cast<uint32_t*>(&sm.u.t.y[1 + lane][0]) = v;
                } else if (lane < 23) {   // the line above: columns x-4 .. x+23
                    const int col = x * 16 - 4 + 4 * (lane - 16);
                    uint32_t v = 0;
                    if (col >= S) { if (mbw > 1) v = *reinterpret_cast<const uint32_t*>(&ks.first8[slot][0][col - S]); }   // (full) wraps onto line 0 of this row's first macroblock
                    else if (row > 0) {
                        if (col < 0) { if (full) v = ks.tail[0][pslot]; }      // pixel line y-2, columns S-4..: line 14 of the row above's last macroblock
                        else if (col < W) v = *reinterpret_cast<const uint32_t*>(up_line + col);
                    }
                    *reinterpret_cast<uint32_t*>(&sm.u.t.y[0][4 * (lane - 16)]) = v;
                }
                if (full && x == mbw - 1 && mbw > 1) {   // right of the last column: lines 1..15 of this row's first macroblock, one line down
                    const int r = lane >> 1, h = lane & 1;
                    if (r < 15) *reinterpret_cast<uint32_t*>(&sm.u.t.y[1 + r][20 + 4 * h]) = *reinterpret_cast<const uint32_t*>(&ks.first8[slot][r + 1][4 * h]);
                }
            } else {
                if (lane < 16) {          // left columns of U (lanes 0-7) and V (8-15)
                    const int pl = lane >> 3, k = lane & 7;
                    uint32_t v = leftw;
                    if (x == 0) v = (full && pl == 0 && k == 0 && row > 0) ? *reinterpret_cast<const uint32_t*>(up_line + S - 4) : 0u;   // U's left at column 0: V's last columns, one line up
                    *reinterpret_cast<uint32_t*>(&sm.u.t.c[pl][1 + k][0]) = v;
                } else if (lane < 22) {   // the line above: three words per plane
                    const int pl = (lane - 16) / 3, q = (lane - 16) % 3;
                    const int fc = pl * (S >> 1) + x * 8 - 4 + 4 * q;          // column inside the S-wide chroma line (U | V)
                    uint32_t v = 0;
                    if (row > 0) {
                        if (fc < 0) { if (full) v = ks.tail[1][pslot]; }       // chroma line c-2, V's last columns
                        else if (fc < S) {
                            const int pc = fc < (S >> 1) ? fc : fc - (S >> 1);
                            if (pc < (W >> 1)) v = *reinterpret_cast<const uint32_t*>(up_line + fc);
                        }
                    }
                    *reinterpret_cast<uint32_t*>(&sm.u.t.c[pl][0][4 * q]) = v;
                }
            }
            __syncwarp();
            intra_ops(g, sm, it, myop, lane, PLANES, (row * 16) * S + x * 16);
            // ---- hand the macroblock's edges on: next macroblock (register), row below (line buffer) ----
            if (row >= KEY_ROWS && lane == 0) {
                // the slot still holds the line of row - KEY_ROWS, which row - KEY_ROWS + 1 reads up to a macroblock ahead
                const uint32_t need = (uint32_t)min(mbw, x + 2);
                key_wait(&prog[row - KEY_ROWS + 1], need);
            }
            __syncwarp();
            if (PLANES == 1) {
                if (lane < 16) leftw = *reinterpret_cast<const uint32_t*>(&sm.u.t.y[1 + lane][16]);
                else if (lane < 20) *reinterpret_cast<uint32_t*>(my_line + x * 16 + 4 * (lane - 16)) = *reinterpret_cast<const uint32_t*>(&sm.u.t.y[16][4 + 4 * (lane - 16)]);
                else if (lane == 20 && x == mbw - 1) ks.tail[0][slot] = *reinterpret_cast<const uint32_t*>(&sm.u.t.y[15][16]);
                if (x == 0) { const int r = lane >> 1, h = lane & 1; *reinterpret_cast<uint32_t*>(&ks.first8[slot][r][4 * h]) = *reinterpret_cast<const uint32_t*>(&sm.u.t.y[1 + r][4 + 4 * h]); }
            } else {
                if (lane < 16) leftw = *reinterpret_cast<const uint32_t*>(&sm.u.t.c[lane >> 3][1 + (lane & 7)][8]);
                else if (lane < 20) {
                    const int pl = (lane - 16) >> 1, j = (lane - 16) & 1;
                    *reinterpret_cast<uint32_t*>(my_line + pl * (S >> 1) + x * 8 + 4 * j) = *reinterpret_cast<const uint32_t*>(&sm.u.t.c[pl][8][4 + 4 * j]);
                } else if (lane == 20 && x == mbw - 1) ks.tail[1][slot] = *reinterpret_cast<const uint32_t*>(&sm.u.t.c[1][7][8]);
            }
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();      // the line is in shared memory before the counter moves
                atomicExch(&prog[row], (uint32_t)x + 1u);
            }
            intra_store(J, g, sm, lane, PLANES, (row * 16) * S + x * 16);
        }
    }
}

__global__ void __launch_bounds__(KEY_WARPS * 32, 1) k_intra_key(const DevJob* __restrict__ jobs, const IntraWork* __restrict__ work,
                                                                 const KeyPic* __restrict__ pics, Geom g, uint32_t* resident) {
    extern __shared__ __align__(16) uint8_t s_key_raw[];   // KEY_WARPS x IntraSmem, KeyShared, the line buffers: above the 48 KB static limit
    IntraSmem* s_all = reinterpret_cast<IntraSmem*>(s_key_raw);
    KeyShared& ks = *reinterpret_cast<KeyShared*>(s_key_raw + KEY_WARPS * sizeof(IntraSmem));
    uint8_t* lines = s_key_raw + KEY_WARPS * sizeof(IntraSmem) + sizeof(KeyShared);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) atomicAdd(resident, 1u);   // this CTA holds its SM resources now (see k_gate)
    if (threadIdx.x < 128) ks.prog[threadIdx.x >> 6][threadIdx.x & 63] = 0;
    __syncthreads();
    IntraSmem& sm = s_all[warp];
    const KeyPic pic = pics[blockIdx.x];
    const DevJob& J = jobs[pic.job];
    const IntraWork* items = work + pic.work_base;   // raster order: every macroblock of an I-picture is intra
    if ((warp >> 1) < g.mbh) {   // the picture's scale table: once per warp, not once per macroblock
        const uint32_t* qt = J.hdr->qtab;
        for (int i = lane; i < 80; i += 32) sm.qtab[i] = __ldg(qt + i);
    }
    key_rows(J, items, g, sm, ks, lines, warp >> 1, lane, 1 + (warp & 1));
}

// ------------------------------------------------------------------------------------------------
// YUV -> BGRA (MD:260-323): strict binary32, source order, no contraction (file is built with -fmad=false)
// ------------------------------------------------------------------------------------------------
// One pixel of the bitmap (MD:262-321) from its luma byte and its (averaged) chroma.
// Clamp to 0..255 and truncation (MD:312-320) are one saturating conversion: cvt.rzi.u8.f32 clamps to the destination's range.
__device__ __forceinline__ uint32_t sat_u8(float v) {
    uint32_t r;
    asm("cvt.rzi.u8.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ uint32_t pack_bgra(float B, float G, float R) {
    // (clamp + add 2^23 rounding toward zero + PRMT instead of the saturating conversions was measured: 0.182 vs 0.165 ms)
    return __byte_perm(__byte_perm(sat_u8(B), sat_u8(G), 0x0040), sat_u8(R) | 0xFF00u, 0x5410);
}
// The reference's arithmetic, binary32 in source order (MD:300-305): the statement of what the bytes must be.
__device__ __forceinline__ uint32_t bgra_px_reference(float Y2, float U, float V) {
    float R = __fadd_rn(Y2, __fmul_rn(1.420f, V)), G = __fsub_rn(__fsub_rn(Y2, __fmul_rn(0.344f, U)), __fmul_rn(0.714f, V)), B = __fadd_rn(Y2, __fmul_rn(1.772f, U));
    R = __fdiv_rn(__fmul_rn(__fsub_rn(R, 16.0f), 255.0f), 239.0f);   // (c - 16) * 255 / (255 - 16)
    G = __fdiv_rn(__fmul_rn(__fsub_rn(G, 16.0f), 255.0f), 239.0f);
    B = __fdiv_rn(__fmul_rn(__fsub_rn(B, 16.0f), 255.0f), 239.0f);
    return pack_bgra(B, G, R);
}
// What the kernel computes instead.  k_bgra is bound by instruction issue, and 23 of its ~53 instructions per pixel were this
// arithmetic (three general divisions in round 1: ~45).  The inputs are discrete -- Y is a byte, U and V are multiples of 1/4 in
// [-128, 127] (one, two or four samples of c - 128 averaged) -- so whether a cheaper expression yields the same BYTE after clamp
// and truncation is decided by enumeration: 256 x 1021 pairs for R and B, 256 x 1021 x 1021 triples for G
// (tools/probe/bgra_formula.cu tried a dozen; k_bgra_selftest below repeats the enumeration on the shipped function and
// tests/test_gpu_parity.py requires zero differences).  What survives:
//   R, B   two contracted multiply-adds with the constants premultiplied by k+ = the float above 255/239:
//          fma(V, 1.420f * k+, fma(Y, k+, -16 k+))  -- 2 instructions instead of 7, the inner one shared by R and B
//   G      the reference's own four matrix operations, then fma(c, 255, -4080) * r with r = RN(1/239) instead of subtraction,
//          multiplication and division (every contraction of the G matrix, and k+ on G, changes between 128 and 4693 of the
//          267 M bytes: tools/probe/bgra_formula_g.cu)
__device__ __forceinline__ uint32_t bgra_px_moflex(float Y2, float U, float V) {
    const float KUP = __uint_as_float(0x3f8891adu), M16K = __uint_as_float(0xc18891adu);       // k+, -16 k+
    const float RV = __uint_as_float(0x3fc1ed94u), BU = __uint_as_float(0x3ff20017u);           // RN(1.420f * k+), RN(1.772f * k+)
    const float R239 = __uint_as_float(0x3b891ac7u);                                            // RN(1 / 239)
    const float t = __fmaf_rn(Y2, KUP, M16K);
    const float R = __fmaf_rn(V, RV, t), B = __fmaf_rn(U, BU, t);
    const float G = __fmul_rn(__fmaf_rn(__fsub_rn(__fsub_rn(Y2, __fmul_rn(0.344f, U)), __fmul_rn(0.714f, V)), 255.0f, -4080.0f), R239);
    return pack_bgra(B, G, R);
}
// [0] number of (Y, U, V) whose bytes differ between bgra_px_moflex and bgra_px_reference, [1] one such triple (Y << 20 | u << 10 | v)
__global__ void k_bgra_selftest(unsigned long long* out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 256 * 1021) return;
    const int yi = idx / 1021, ui = idx - yi * 1021;
    const float Y = (float)yi, U = (float)(ui - 512) * 0.25f;
    unsigned long long bad = 0, where = 0;
    for (int vi = 0; vi < 1021; vi++) {
        const float V = (float)(vi - 512) * 0.25f;
        if (bgra_px_moflex(Y, U, V) != bgra_px_reference(Y, U, V)) { bad++; where = (unsigned long long)yi << 20 | (unsigned long long)ui << 10 | (unsigned long long)vi; }
    }
    if (bad) { atomicAdd(out, bad); atomicExch(out + 1, where); }
}
__device__ __forceinline__ uint32_t bgra_px(bool moflex, float Y2, float U, float V) {
    if (moflex) return bgra_px_moflex(Y2, U, V);
    // ModsDS (MD:309-311): integer sums of the truncated values
    const float R = (float)((int)Y2 + (int)U - (int)V), G = (float)((int)Y2 + (int)V), B = (float)((int)Y2 - (int)U - (int)V);
    return pack_bgra(B, G, R);
}
// (float)(byte k of word) minus a bias in two instructions, one of them on the FP32 pipe: PRMT drops the byte into the low
// mantissa bits of 2^23, the subtraction removes 2^23 + bias.  Exact.  (Shifts, masks and integer adds run at half the FP32
// rate on this chip and conversions at a quarter: the kernel is issue-bound, so pixel unpacking is kept off those pipes.)
template <int K>
__device__ __forceinline__ float byte_to_float(uint32_t word, float bias_plus_2p23) {
    return __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7540 + K)) - bias_plus_2p23;
}

// Eight pixels per thread: one 64-bit luma load, the five chroma columns the eight pixels touch as one aligned word + one byte
// per plane (two rows of them on odd lines), two 16-byte stores.  Width is a multiple of 16, so a picture row is W / 8 threads.
// The reference averages chroma in float -- (c - 128) summed over the 1, 2 or 4 samples a pixel uses, divided by their number
// (MD:269-297); these are sums of small integers and divisions by powers of two, exact in binary32 in any order, so the
// samples are converted once each ((c - 128) straight from the packed word) and summed on the FP32 pipe.
template <bool MOFLEX>   // the two colour matrices (MD:300-305 / 309-311) as two kernels: no per-pixel branch, half the code
__global__ void __launch_bounds__(256) k_bgra(const uint8_t* const* __restrict__ srcs, uint8_t* __restrict__ dst, int pitch, size_t per, Geom g, uint32_t wo_magic) {
    const int wo = g.W >> 3;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (uint32_t)(wo * g.H)) return;
    const int y = (int)__umulhi(i, wo_magic), x = ((int)i - y * wo) * 8;   // exact: i < 2^26 (magic = ceil(2^32 / wo))
    const uint8_t* __restrict__ src = srcs[blockIdx.y];
    const int S = g.S, h = S >> 1;
    const uint8_t* C = src + (size_t)S * g.H + (y >> 1) * S + (x >> 1);
    const uint2 yw = *reinterpret_cast<const uint2*>(src + y * S + x);
    // chroma columns x/2 .. x/2 + 4 of this line's chroma row and (odd lines) the next one.  The fifth column and the next row
    // are only read where a pixel uses them: not past the picture's last column / row (MD:269).
    const bool last8 = x + 8 >= g.W, lasty = y == g.H - 1;
    const bool below = (y & 1) && !lasty;
    const uint32_t uw = *reinterpret_cast<const uint32_t*>(C), vw = *reinterpret_cast<const uint32_t*>(C + h);
    constexpr float BC = 8388608.0f + 128.0f;
    float u[5], v[5];
    u[0] = byte_to_float<0>(uw, BC); u[1] = byte_to_float<1>(uw, BC); u[2] = byte_to_float<2>(uw, BC); u[3] = byte_to_float<3>(uw, BC);
    v[0] = byte_to_float<0>(vw, BC); v[1] = byte_to_float<1>(vw, BC); v[2] = byte_to_float<2>(vw, BC); v[3] = byte_to_float<3>(vw, BC);
    u[4] = last8 ? 0.0f : byte_to_float<0>(C[4], BC); v[4] = last8 ? 0.0f : byte_to_float<0>(C[h + 4], BC);
    const float u3top = u[3], v3top = v[3];   // (the last column's odd pixel uses its own sample only, even on an odd line)
    if (below) {
        const uint32_t ub = *reinterpret_cast<const uint32_t*>(C + S), vb = *reinterpret_cast<const uint32_t*>(C + S + h);
        u[0] += byte_to_float<0>(ub, BC); u[1] += byte_to_float<1>(ub, BC); u[2] += byte_to_float<2>(ub, BC); u[3] += byte_to_float<3>(ub, BC);
        v[0] += byte_to_float<0>(vb, BC); v[1] += byte_to_float<1>(vb, BC); v[2] += byte_to_float<2>(vb, BC); v[3] += byte_to_float<3>(vb, BC);
        if (!last8) { u[4] += byte_to_float<0>(C[S + 4], BC); v[4] += byte_to_float<0>(C[S + h + 4], BC); }
    }
    // samples per pixel: even x -> 1 (2 on odd lines), odd x -> twice that, except in the last row / column (none but its own)
    const float k1 = below ? 0.5f : 1.0f;
    const bool horiz = !lasty;                                                                     // odd x may use its right-hand neighbour
    const float k2 = horiz ? k1 * 0.5f : 1.0f;
    float Yf[8];
    Yf[0] = byte_to_float<0>(yw.x, 8388608.0f); Yf[1] = byte_to_float<1>(yw.x, 8388608.0f); Yf[2] = byte_to_float<2>(yw.x, 8388608.0f); Yf[3] = byte_to_float<3>(yw.x, 8388608.0f);
    Yf[4] = byte_to_float<0>(yw.y, 8388608.0f); Yf[5] = byte_to_float<1>(yw.y, 8388608.0f); Yf[6] = byte_to_float<2>(yw.y, 8388608.0f); Yf[7] = byte_to_float<3>(yw.y, 8388608.0f);
    uint32_t out[8];
#pragma unroll
    for (int p = 0; p < 8; p++) {
        const int c = p >> 1;
        float U, V;
        if (!(p & 1)) { U = u[c] * k1; V = v[c] * k1; }
        else if (p == 7 && (last8 || !horiz)) { U = u3top; V = v3top; }
        else if (horiz) { U = (u[c] + u[c + 1]) * k2; V = (v[c] + v[c + 1]) * k2; }
        else { U = u[c]; V = v[c]; }   // last row (never an odd line's second row: `below` is off)
        out[p] = bgra_px(MOFLEX, Yf[p], U, V);
    }
    uint4* o = reinterpret_cast<uint4*>(dst + per * blockIdx.y + (size_t)y * pitch + (size_t)x * 4);
    o[0] = make_uint4(out[0], out[1], out[2], out[3]);
    o[1] = make_uint4(out[4], out[5], out[6], out[7]);
}

// strided planes -> tight I420, 4 bytes per thread
__global__ void __launch_bounds__(256) k_pack_i420(const uint8_t* const* __restrict__ srcs, uint8_t* __restrict__ dst, Geom g) {
    const uint8_t* src = srcs[blockIdx.y];
    const int W = g.W, H = g.H, S = g.S;
    const int ywords = W * H / 4, cwords = W * H / 16;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ywords + 2 * cwords) return;
    uint8_t* out = dst + (size_t)blockIdx.y * (size_t)(W * H * 3 / 2);
    uint32_t v;
    if (i < ywords) {
        const int b = i * 4, y = b / W, x = b % W;
        v = *reinterpret_cast<const uint32_t*>(src + y * S + x);
    } else {
        const int j = i - ywords, p = j >= cwords, b = (j - p * cwords) * 4, y = b / (W / 2), x = b % (W / 2);
        v = *reinterpret_cast<const uint32_t*>(src + (size_t)S * H + y * S + x + p * (S / 2));
    }
    reinterpret_cast<uint32_t*>(out)[i] = v;
}

}  // namespace

cudaError_t init_kernel_tables() {
    uint32_t lut[64];
    for (uint32_t m = 0; m < 64; m++) {
        uint32_t list = 0; int n = 0;
        for (uint32_t k = 0; k < 6; k++) if ((m >> k) & 1u) { list |= k << (4 * n); n++; }
        lut[m] = list;
    }
    return cudaMemcpyToSymbol(c_blklist, lut, sizeof lut);
}

// The inter path.  Three formulations exist, all bit-exact (tests/test_gpu_parity.py runs each) and all within 5 % of each other
// on the bench mix -- they share their bound, the rate at which TMA delivers small boxes at a 50 % L2 hit rate (DESIGN.md 4,
// tools/probe/tma_rate.cu) -- so the default is the fastest, last round's fused kernel:
//   (default)                  k_inter_chunk: MC + residual fused, runs of 4 macroblocks, 4 box slots per run
//   MOBI_INTER_KERNEL=v3       k_inter_v3 (mobi_inter_v3.cuh): box slots in rounds, per-leaf eligibility, conflict-free coefficient
//                              pool in visiting order, DC-only blocks without a transform; MOBI_INTER_CHUNK=8 (8-macroblock
//                              chunks) and MOBI_INTER_CTAS=6 / 8 (CTAs per SM the layout is sized for; default 7)
//   MOBI_INTER_KERNEL=split    k_mc + k_res (mobi_inter_split.cuh): motion compensation and residual as two kernels at 36-40
//                              warps per SM, macroblocks of more than two leaves by boxes too; gives the MC and the IDCT kernel
//                              their own times (bench.py reports them)
static int inter_kernel_choice() {   // -1: k_mc + k_res; 0: k_inter_chunk; else k_inter_v3 with CTAs per SM * 2 + (8-macroblock chunks ? 1 : 0)
    static const int c = [] {
        const char* e = getenv("MOBI_INTER_KERNEL");
        const char* ch = getenv("MOBI_INTER_CHUNK");
        const char* ct = getenv("MOBI_INTER_CTAS");
        if (e && !strcmp(e, "split")) return -1;
        if (!e || strcmp(e, "v3")) return 0;
        const int ctas = ct && (!strcmp(ct, "6") || !strcmp(ct, "8")) ? atoi(ct) : 7;
        return ctas * 2 + ((ch && !strcmp(ch, "8")) ? 1 : 0);
    }();
    return c;
}
const char* inter_kernel_name() { const int c = inter_kernel_choice(); return c < 0 ? "k_mc+k_res" : c == 0 ? "k_inter_chunk" : "k_inter_v3"; }

cudaError_t launch_inter(const DevJob* jobs, int n_jobs, Geom g, const InterMaps& tm, int sm_count, uint32_t* tickets, uint32_t* ticket_base,
                         cudaStream_t st, cudaEvent_t* between) {
    if (n_jobs <= 0) return cudaSuccess;
    const uint32_t magic = (uint32_t)((0x100000000ull + (uint64_t)g.mbw - 1) / (uint64_t)g.mbw);
    const int choice = inter_kernel_choice();
    static const uint32_t exp_flags = [] { const char* e = getenv("MOBI_INTER_EXP"); return e ? (uint32_t)atoi(e) : 0u; }();   // timing experiments: 1 no residual (wrong pixels), 2 v3 without the multi-leaf box path, 4 k_mc without chroma boxes (wrong pixels), 8 L2 prefetch of the chunk's region
    static const bool chunk8 = [] { const char* ch = getenv("MOBI_INTER_CHUNK"); return ch && !strcmp(ch, "8"); }();
    const uint32_t chunk_mbs = (choice > 0 ? (choice & 1) != 0 : (choice == 0 && chunk8)) ? 8u : 16u;
    const uint32_t cpp = ((uint32_t)(g.mbw * g.mbh) + chunk_mbs - 1) / chunk_mbs, n_chunks = cpp * (uint32_t)n_jobs;
    const uint32_t cpp_magic = (uint32_t)((0x100000000ull + (uint64_t)cpp - 1) / (uint64_t)cpp);
    if (choice < 0) {
        uint32_t ctas = (uint32_t)sm_count * (uint32_t)MC_CTAS, rctas = (uint32_t)sm_count * (uint32_t)RES_CTAS;
        if (ctas > (n_chunks + MC_WARPS - 1) / MC_WARPS) ctas = (n_chunks + MC_WARPS - 1) / MC_WARPS;
        if (rctas > (n_chunks + RES_WARPS - 1) / RES_WARPS) rctas = (n_chunks + RES_WARPS - 1) / RES_WARPS;
#define MOBI_MC(L) k_mc<L><<<ctas, MC_WARPS * 32, 0, st>>>(jobs, n_chunks, cpp, cpp_magic, g.mbw, magic, g.H, tm.ring_rows, tickets, ticket_base[0], exp_flags, tm.l2, tm.c3)
#define MOBI_RES(L) k_res<L><<<rctas, RES_WARPS * 32, 0, st>>>(jobs, n_chunks, cpp, cpp_magic, g.mbw, magic, g.H, tickets + 16, ticket_base[1])
        if (g.log2S == 8) MOBI_MC(8); else if (g.log2S == 9) MOBI_MC(9); else MOBI_MC(10);
        ticket_base[0] += n_chunks + ctas * MC_WARPS;   // every warp draws exactly one ticket past the end
        if (between) { cudaEventRecord(between[0], st); cudaEventRecord(between[1], st); }
        if (!(exp_flags & 1u)) {
            if (g.log2S == 8) MOBI_RES(8); else if (g.log2S == 9) MOBI_RES(9); else MOBI_RES(10);
            ticket_base[1] += n_chunks + rctas * RES_WARPS;
        }
#undef MOBI_MC
#undef MOBI_RES
        return cudaGetLastError();
    }
    const uint32_t warps = choice == 0 ? CH_WARPS : V3_WARPS;
    uint32_t ctas = (uint32_t)sm_count * (choice == 0 ? 6u : (uint32_t)(choice >> 1));
    if (ctas > (n_chunks + warps - 1) / warps) ctas = (n_chunks + warps - 1) / warps;
    // k_inter_chunk: the last pictures -- about MOBI_INTER_TAIL (default 1) chunks' worth of macroblocks per resident warp --
    // are handed out run by run
    InterTail tail{n_chunks, (uint32_t)n_jobs, 1u, 0u, chunk_mbs};
    uint32_t n_tickets = n_chunks;
    if (choice == 0) {
        static const float tail_chunks = [] { const char* e = getenv("MOBI_INTER_TAIL"); return e ? (float)atof(e) : 1.0f; }();
        const uint32_t mb_per_pic = (uint32_t)(g.mbw * g.mbh);
        uint32_t tail_jobs = (uint32_t)((double)tail_chunks * ctas * warps * chunk_mbs / mb_per_pic);
        if (tail_jobs > (uint32_t)n_jobs) tail_jobs = (uint32_t)n_jobs;
        tail.first_job = (uint32_t)n_jobs - tail_jobs;
        tail.n_big = tail.first_job * cpp;
        tail.rpp = (mb_per_pic + 3u) / 4u;
        tail.rpp_magic = (uint32_t)((0x100000000ull + (uint64_t)tail.rpp - 1) / (uint64_t)tail.rpp);
        n_tickets = tail.n_big + tail_jobs * tail.rpp;
    }
#define MOBI_LAUNCH(K) K<<<ctas, warps * 32, 0, st>>>(jobs, n_tickets, cpp, cpp_magic, g.mbw, magic, g.H, tickets, ticket_base[0], (exp_flags & 8u) ? 1u : 0u, tail, tm.l3, tm.c4)
#define MOBI_LAUNCH3(K) K<<<ctas, warps * 32, 0, st>>>(jobs, n_chunks, cpp, cpp_magic, g.mbw, magic, g.H, tm.ring_rows, tickets, ticket_base[0], exp_flags, tm.l2, tm.c3)
#define MOBI_BY_STRIDE(C, T) do { if (g.log2S == 8) MOBI_LAUNCH3((k_inter_v3<8, C, T>)); else if (g.log2S == 9) MOBI_LAUNCH3((k_inter_v3<9, C, T>)); else MOBI_LAUNCH3((k_inter_v3<10, C, T>)); } while (0)
    switch (choice) {
    case 0:
        if (g.log2S == 8) MOBI_LAUNCH(k_inter_chunk<8>); else if (g.log2S == 9) MOBI_LAUNCH(k_inter_chunk<9>); else MOBI_LAUNCH(k_inter_chunk<10>);
        break;
    case 12: MOBI_BY_STRIDE(16, 6); break;
    case 13: MOBI_BY_STRIDE(8, 6); break;
    case 14: MOBI_BY_STRIDE(16, 7); break;
    case 15: MOBI_BY_STRIDE(8, 7); break;
    case 16: MOBI_BY_STRIDE(16, 8); break;
    default: MOBI_BY_STRIDE(8, 8); break;
    }
#undef MOBI_BY_STRIDE
#undef MOBI_LAUNCH3
#undef MOBI_LAUNCH
    ticket_base[0] += n_tickets + ctas * warps;   // every warp draws exactly one ticket past the end
    if (between) { cudaEventRecord(between[0], st); cudaEventRecord(between[1], st); }
    return cudaGetLastError();
}

cudaError_t launch_intra(const DevJob* jobs, const IntraWork* work, uint32_t n_work, uint32_t* ticket, uint32_t ticket_base,
                         uint32_t stamp, Geom g, uint32_t max_warps, cudaStream_t st, uint32_t* warps_launched) {
    *warps_launched = 0;
    if (n_work == 0) return cudaSuccess;
    // Tickets are handed out in dependency order (whoever is awaited holds an earlier ticket), so a warp only ever waits for tickets drawn before its own,
    // which are held by warps that are already running: any grid size makes progress.
    unsigned blocks = (unsigned)((n_work + INTRA_WARPS - 1) / INTRA_WARPS);
    unsigned cap = (max_warps + INTRA_WARPS - 1) / INTRA_WARPS;
    if (cap < 1) cap = 1;
    if (blocks > cap) blocks = cap;
    static const bool pdl = [] { const char* e = getenv("MOBI_PDL"); return !e || atoi(e) != 0; }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(INTRA_WARPS * 32); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, k_intra, jobs, work, n_work, ticket, ticket_base, stamp, g);
    *warps_launched = blocks * INTRA_WARPS;
    return e != cudaSuccess ? e : cudaGetLastError();
}

cudaError_t launch_intra_key(const DevJob* jobs, const IntraWork* work, const void* pics, int n_pics, Geom g, uint32_t* resident, cudaStream_t st) {
    if (n_pics <= 0) return cudaSuccess;
    // per device, so set on every launch (a host-side table lookup)
    const cudaError_t attr = cudaFuncSetAttribute(k_intra_key, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)key_smem_bytes(g.S));
    if (attr != cudaSuccess) return attr;
    k_intra_key<<<(unsigned)n_pics, KEY_WARPS * 32, key_smem_bytes(g.S), st>>>(jobs, work, reinterpret_cast<const KeyPic*>(pics), g, resident);
    return cudaGetLastError();
}

// Holds back the stream it is launched on until the I-picture CTAs launched on the other stream are resident (or a
// time budget runs out): those CTAs are large (16 warps) and few, and if the tens of thousands of small k_inter CTAs
// get to the SMs first they cannot find room until k_inter drains -- the two kernels then run back to back instead of
// side by side (measured: 0.86 ms instead of 0.56 ms per step, at random).
__global__ void k_gate(const uint32_t* resident, uint32_t target, long long budget_cycles) {
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        while ((int32_t)(*(volatile const uint32_t*)resident - target) < 0 && clock64() - t0 < budget_cycles) __nanosleep(200);
    }
}
cudaError_t launch_gate(const uint32_t* resident, uint32_t target, cudaStream_t st) {
    k_gate<<<1, 32, 0, st>>>(resident, target, 400000);   // ~0.2 ms at 1.965 GHz, then give up waiting
    return cudaGetLastError();
}

cudaError_t launch_bgra(const uint8_t* const* srcs, int n, uint8_t* dst, int dst_pitch, size_t dst_picture_bytes, Geom g, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const int wo = g.W >> 3, items = wo * g.H;   // eight pixels per thread
    const uint32_t magic = (uint32_t)((0x100000000ull + (uint64_t)wo - 1) / (uint64_t)wo);
    dim3 grid((unsigned)((items + 255) / 256), (unsigned)n);
    if (g.version == MOBI_MOFLEX3DS) k_bgra<true><<<grid, 256, 0, st>>>(srcs, dst, dst_pitch, dst_picture_bytes, g, magic);
    else k_bgra<false><<<grid, 256, 0, st>>>(srcs, dst, dst_pitch, dst_picture_bytes, g, magic);
    return cudaGetLastError();
}

cudaError_t selftest_bgra(unsigned long long* mismatches_host) {
    unsigned long long* d = nullptr;
    cudaError_t e = cudaMalloc(&d, 2 * sizeof *d);
    if (e != cudaSuccess) return e;
    cudaMemset(d, 0, 2 * sizeof *d);
    k_bgra_selftest<<<(256 * 1021 + 255) / 256, 256>>>(d);
    e = cudaMemcpy(mismatches_host, d, 2 * sizeof *d, cudaMemcpyDeviceToHost);
    cudaFree(d);
    return e;
}

cudaError_t launch_pack_i420(const uint8_t* const* srcs, int n, uint8_t* dst, Geom g, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const int words = g.W * g.H * 3 / 8;
    dim3 grid((unsigned)((words + 255) / 256), (unsigned)n);
    k_pack_i420<<<grid, 256, 0, st>>>(srcs, dst, g);
    return cudaGetLastError();
}

}  // namespace mobi
