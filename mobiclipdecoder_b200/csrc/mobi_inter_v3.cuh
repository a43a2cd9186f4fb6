// k_inter_v3 -- the inter kernel (motion compensation + dequantisation + inverse transforms + add / clip), included by
// mobi_kernels.cu inside its anonymous namespace.  "MD:n" = LibMobiclip/Codec/Mobiclip/MobiclipDecoder.cs:n.
//
// Same decomposition as k_inter_chunk (persistent warps, ticketed chunks of 16 macroblocks, runs of 4, lane-parallel
// set-up, TMA boxes, pooled transform passes, run-parallel 16-byte stores); what changed is where the round-1 profile
// (profiles/r01l) put the time: the ALU pipe at 62 % and the L1 data pipe at 68 % of peak, both fed by
//   * the 22 % of macroblocks that found no box slot and went down the load-per-lane path (250 data-pipe wavefronts each);
//   * box reads with 4-way bank conflicts (32-byte row pitch);
//   * per-pass address arithmetic of the coefficient pool (slot-dependent XOR swizzle) and a mixed 8x8 / 4x4 code path;
//   * compiler-made moves and branches around all of it (IMAD.MOV 107, BRA 31 of 459 instructions per macroblock).
// So:
//   * box slots are handed out in ROUNDS: what does not fit the five slots of a run in the first round is fetched when its
//     turn comes, into the slots the macroblocks before it have finished with.  Only windows that leave their pixel row
//     (flat addressing wraps where TMA zero-fills) and macroblocks of more than two leaves still load per lane;
//   * eligibility is decided per LEAF (its own columns must stay inside the row), not per 16x16 window;
//   * a box row is read as two aligned 8-byte units (2-way instead of 4-way conflicts, 2 loads instead of 3);
//   * the horizontal half-pel filter halves the aligned words BEFORE the byte shift (the halving is bytewise, so it
//     commutes with the shift): 3 halvings per row instead of 4 per output word;
//   * pooled coefficient blocks lie 288 bytes apart in the order the passes visit them (8x8-transformed first), which the
//     host makes possible by telling the kernel which coded blocks are 8x8-transformed (mobi_mb.info): every address of a
//     pass is a per-lane constant plus slot * 288, conflict-free, and passes are all-8x8 or all-4x4 except at most one;
//   * blocks whose only coefficient is the DC (MD:2938 IDCT1Px8, MD:2954 IDCT1Px4: the reference's own size dispatch) skip
//     the transform: the residual is one constant, (dc + 32) >> 6, added to the whole block.
// 4 warps per CTA, 6 CTAs per SM (85 registers instead of 72: the round-1 kernel sat on its register limit).

constexpr int V3_WARPS = 4, V3_RUN = 4, V3_PSTRIDE = 72;   // pool stride in words (288 bytes)
constexpr uint32_t MB_M8_LO = 24, MB_M8_HI = 29;   // mobi_mb.info (inter): bits 24-27 blocks 0-3, bits 29-30 blocks 4-5 are 8x8-transformed

// Per-warp shared memory.  CTAS = CTAs (of four warps) per SM the layout is sized for: 6 -> six box slots, 7 -> five, 8 -> four.
template <int CTAS>
struct V3Smem {
    static constexpr int SLOTS = CTAS >= 8 ? 4 : CTAS == 7 ? 5 : 6;
    static constexpr int POOL = SLOTS * 1280 / (V3_PSTRIDE * 4) > 24 ? 24 : SLOTS * 1280 / (V3_PSTRIDE * 4);   // pooled coefficient blocks that fit the boxes' space (24 / 22 / 17)
    union {
        uint8_t box[SLOTS * 1280];               // per slot: 32x17 luma box at +0 (544 bytes), 32x2x9 chroma box at +640 (576 bytes)
        int32_t pool[POOL * V3_PSTRIDE];         // pooled coefficient blocks, 288 bytes apart
    } u;
    static constexpr int TILE = 416;             // 384 used: keeps the four tiles 32 bytes apart modulo 128 (run-parallel final store)
    // prediction (+ residual): luma 16 rows x 16, then U 8x8, V 8x8.  The 32 spare bytes of tiles 0-2 hold slotinfo[24] (per
    // pooled block in visiting order: byte offset of its pixels in tile[][] | chroma << 11 | pool slot << 12), those of tile 3
    // the two mbarriers (one per round of boxes).
    uint8_t tile[V3_RUN][TILE];
    __device__ __forceinline__ uint32_t& slotinfo(uint32_t p) { return *reinterpret_cast<uint32_t*>(&tile[p >> 3][384 + (p & 7u) * 4u]); }
    __device__ __forceinline__ uint64_t* bar(uint32_t round) { return reinterpret_cast<uint64_t*>(&tile[3][384]) + round; }
};
static_assert(sizeof(V3Smem<6>) == 9344 && sizeof(V3Smem<7>) == 8064 && sizeof(V3Smem<8>) == 6784, "per-warp shared memory");
static_assert(sizeof(V3Smem<7>) % 128 == 0 && sizeof(V3Smem<8>) % 128 == 0, "TMA destinations must stay 128-byte aligned");
static_assert((sizeof(V3Smem<6>) * V3_WARPS + 1024) * 6 <= 233472 && (sizeof(V3Smem<7>) * V3_WARPS + 1024) * 7 <= 233472 && (sizeof(V3Smem<8>) * V3_WARPS + 1024) * 8 <= 233472, "CTAs per SM");

// Four pixels + one constant residual -> saturated bytes (the DC-only blocks).
__device__ __forceinline__ uint32_t addsat4c(uint32_t px, int r) {
    const int p0 = (int)(px & 255u) + r, p1 = (int)((px >> 8) & 255u) + r, p2 = (int)((px >> 16) & 255u) + r, p3 = (int)(px >> 24) + r;
    uint32_t hi, out;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(p3), "r"(p2), "r"(0));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(out) : "r"(p1), "r"(p0), "r"(hi));
    return out;
}
// Three aligned words holding bytes [off & 7 .. (off & 7) + 11] of a box row: two 8-byte loads, the right three of the
// four words picked by bit 2 of the offset.
__device__ __forceinline__ void v3_row3(const uint8_t* base, uint32_t off, uint32_t& x0, uint32_t& x1, uint32_t& x2) {
    const uint2* q = reinterpret_cast<const uint2*>(base + (off & ~7u));
    const uint2 p0 = q[0], p1 = q[1];
    const bool hi = (off & 4u) != 0u;
    x0 = hi ? p0.y : p0.x; x1 = hi ? p1.x : p0.y; x2 = hi ? p1.y : p1.x;
}
// CopyBlock (MD:418-456) for this lane's eight luma pixels out of a box (row pitch 32): truncating averages on packed bytes.
__device__ __forceinline__ void v3_luma(const uint8_t* base, uint32_t off, uint32_t phase, uint32_t& o0, uint32_t& o1) {
    const uint32_t sh = (off & 3u) * 8u;
    uint32_t x0, x1, x2;
    v3_row3(base, off, x0, x1, x2);
    if (!(phase & 1u)) {
        o0 = __funnelshift_r(x0, x1, sh); o1 = __funnelshift_r(x1, x2, sh);
        if (phase == 0u) return;
        uint32_t y0, y1, y2;
        v3_row3(base, off + 32u, y0, y1, y2);
        o0 = half4(o0) + half4(__funnelshift_r(y0, y1, sh)); o1 = half4(o1) + half4(__funnelshift_r(y1, y2, sh));
        return;
    }
    x0 = half4(x0); x1 = half4(x1); x2 = half4(x2);
    o0 = __funnelshift_r(x0, x1, sh) + __funnelshift_rc(x0, x1, sh + 8u);
    o1 = __funnelshift_r(x1, x2, sh) + __funnelshift_rc(x1, x2, sh + 8u);
    if (phase == 1u) return;
    uint32_t y0, y1, y2;
    v3_row3(base, off + 32u, y0, y1, y2);
    y0 = half4(y0); y1 = half4(y1); y2 = half4(y2);
    o0 = half4(o0) + half4(__funnelshift_r(y0, y1, sh) + __funnelshift_rc(y0, y1, sh + 8u));
    o1 = half4(o1) + half4(__funnelshift_r(y1, y2, sh) + __funnelshift_rc(y1, y2, sh + 8u));
}
// The same for this lane's four chroma pixels out of the merged U/V box (rows of the two planes alternate: the next
// pixel row is 64 bytes on).
__device__ __forceinline__ uint32_t v3_chroma(const uint8_t* base, uint32_t off, uint32_t phase) {
    const uint32_t sh = (off & 3u) * 8u;
    const uint32_t* q = reinterpret_cast<const uint32_t*>(base + (off & ~3u));
    uint32_t x0 = q[0], x1 = q[1];
    if (!(phase & 1u)) {
        const uint32_t a = __funnelshift_r(x0, x1, sh);
        if (phase == 0u) return a;
        return half4(a) + half4(__funnelshift_r(q[16], q[17], sh));
    }
    x0 = half4(x0); x1 = half4(x1);
    const uint32_t s = __funnelshift_r(x0, x1, sh) + __funnelshift_rc(x0, x1, sh + 8u);
    if (phase == 1u) return s;
    const uint32_t y0 = half4(q[16]), y1 = half4(q[17]);
    return half4(s) + half4(__funnelshift_r(y0, y1, sh) + __funnelshift_rc(y0, y1, sh + 8u));
}

// Per-macroblock word that travels by shuffle: bits 0-11 leaf 0 (column of its window inside the 16-byte-aligned box 4,
// luma half-pel phase 2, the same for the chroma window 4 + 2), 12-23 leaf 1, 24-26 what to do, 27-29 first box slot.
constexpr uint32_t V3_SKIP = 0, V3_LPL = 1, V3_BOX1 = 2, V3_MULTI = 3, V3_BOX2 = 4;   // BOX2 | 1: split left / right (else top / bottom); | 2: leaf 0 is the bottom / right half

// Load-per-lane prediction of one macroblock (CopyBlock MD:418-456 on the flat planes; any partition tree, any vector the
// parser accepted): this lane's eight luma pixels (x, y) and four chroma pixels (z).  Leaf records are read from memory as
// they are needed (they sit in one or two cache lines); the 2x2-granular partition map goes through `map` (64 bytes of the
// macroblock's own, not yet written, tile).  Rare (about one macroblock in eleven on the bench mix) and large: kept out of
// line so that the kernel's hot loop stays inside the instruction cache.
template <int LOG2S>
__device__ __forceinline__ uint3 v3_lpl(const DevJob& J, int n_parts, uint32_t first_part, int yo, int H, uint8_t* map, int lane) {
    constexpr int S = 1 << LOG2S;
    const int lrow = lane >> 1, lhalf = lane & 1, cpl = lane >> 4, crow = (lane >> 1) & 7;
    const size_t ysz = (size_t)S * H;
    const uint2* const sp = reinterpret_cast<const uint2*>(J.parts + first_part);
    uint32_t ml = 0, mc = 0;
    if (n_parts > 1) {
        // partition map at 2x2-pixel granularity (leaves go down to 2x2, MD:1726): lane l owns cells 2l, 2l+1 of the 8x8 grid
        const int cy2 = lane >> 2, cx2 = (lane & 3) * 2;
        uint32_t i0 = 0, i1 = 0;
#pragma unroll 1
        for (int p = 0; p < n_parts; p++) {
            const uint32_t pw = __ldg(&sp[p].x);
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
    auto leaf = [&](uint32_t idx) { const uint2 pw = __ldg(sp + idx); return part_of(pw.x, pw.y); };
    uint3 out;
    if (ml == (ml & 255u) * 0x01010101u) {
        const PartV p = leaf(ml & 255u);
        mc_row8(J.ref[p.ref - 1] + ypix + ((p.mvy >> 1) << LOG2S) + (p.mvx >> 1), S, (p.mvx & 1) | ((p.mvy & 1) << 1), out.x, out.y);
    } else {
        uint32_t o[2] = {0, 0};
#pragma unroll 1
        for (int c = 0; c < 4; c++) {
            const PartV p = leaf((ml >> (8 * c)) & 255u);
            const uint8_t* s = J.ref[p.ref - 1] + ypix + 2 * c + ((p.mvy >> 1) << LOG2S) + (p.mvx >> 1);
            const int ph = (p.mvx & 1) | ((p.mvy & 1) << 1);
            const uint32_t v = (mc_px(s, S, ph) | mc_px(s + 1, S, ph) << 8) << (16 * (c & 1));
            if (c < 2) o[0] |= v; else o[1] |= v;
        }
        out.x = o[0]; out.y = o[1];
    }
    if (mc == (mc & 255u) * 0x01010101u) {
        const PartV p = leaf(mc & 255u);
        const int cx = p.mvx >> 1, cy = p.mvy >> 1;
        out.z = mc_row4(J.ref[p.ref - 1] + ysz + cpix + ((cy >> 1) << LOG2S) + (cx >> 1), S, (cx & 1) | ((cy & 1) << 1));
    } else {
        out.z = 0;
#pragma unroll 1
        for (int c = 0; c < 4; c++) {
            const PartV p = leaf((mc >> (8 * c)) & 255u);
            const int cx = p.mvx >> 1, cy = p.mvy >> 1;
            out.z |= mc_px(J.ref[p.ref - 1] + ysz + cpix + c + ((cy >> 1) << LOG2S) + (cx >> 1), S, (cx & 1) | ((cy & 1) << 1)) << (8 * c);
        }
    }
    __syncwarp();   // everyone has read the partition map before the prediction overwrites it
    return out;
}

// Prediction of a macroblock of MORE THAN TWO leaves from boxes (after the run's other macroblocks, when all box slots are
// free): the leaves go through the slots SLOTS at a time -- lane p fetches leaf p's records, decides whether its columns stay
// inside their pixel row and issues the leaf's two boxes -- and every lane merges, leaf by leaf, the pixels of its eight luma /
// four chroma positions the leaf covers (leaves go down to 2x2: a lane's pixels may belong to four of them).  A leaf whose
// window leaves its row sends the whole macroblock down the load-per-lane path.  Out of line like v3_lpl: rare and large.
template <int LOG2S, int CTAS>
__device__ __forceinline__ uint4 v3_multi(const DevJob& J, V3Smem<CTAS>& sm, const CUtensorMap* tm_l2, const CUtensorMap* tm_c3, int n_parts, uint32_t first_part,
                                       int mbx, int mby, int H, int ring_rows, uint32_t phases, uint8_t* map, int lane) {   // .w: phases, updated
    constexpr int S = 1 << LOG2S;
    constexpr int SLOTS = V3Smem<CTAS>::SLOTS;
    const int lrow = lane >> 1, lhalf = lane & 1, cpl = lane >> 4, crow = (lane >> 1) & 7;
    const uint32_t l_off = (uint32_t)(lrow * 32 + lhalf * 8), c_off = (uint32_t)((crow * 2 + cpl) * 32 + lhalf * 4);
    const uint2* const sp = reinterpret_cast<const uint2*>(J.parts + first_part);
    // every leaf's columns inside its own pixel row?  (two leaves per lane: a macroblock has at most 64)
    bool safe = true;
#pragma unroll 1
    for (int p = lane; p < n_parts; p += 32) {
        const uint2 pw = __ldg(sp + p);
        const PartV v = part_of(pw.x, pw.y);
        const int lx = (int)(pw.x & 15u) * 2, lw = 2 << ((pw.x >> 8) & 3u);
        const int xw = mbx * 16 + (v.mvx >> 1) + lx, cxw = mbx * 8 + (v.mvx >> 2) + (lx >> 1);
        safe = safe && xw >= 0 && xw + lw + 1 <= S && cxw >= 0 && cxw + (lw >> 1) + 1 <= (S >> 1);
    }
    if (!__all_sync(0xffffffffu, safe)) {
        const uint3 px = v3_lpl<LOG2S>(J, n_parts, first_part, ((mby * 16) << LOG2S) + mbx * 16, H, map, lane);
        return make_uint4(px.x, px.y, px.z, phases);
    }
    uint4 out = make_uint4(0u, 0u, 0u, 0u);
    const uint32_t bar = smem_u32(sm.bar(0));
#pragma unroll 1
    for (int g0 = 0; g0 < n_parts; g0 += SLOTS) {
        const int cnt = min(SLOTS, n_parts - g0);
        uint32_t rect = 0, word = 0;
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"((uint32_t)cnt * (TMA_BYTES_L + TMA_BYTES_C4)) : "memory");
        // everyone is done with what the slots held; that generic-proxy traffic is ordered before the boxes
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane < cnt) {
            const uint2 pw = __ldg(sp + g0 + lane);
            const PartV v = part_of(pw.x, pw.y);
            const int x0 = mbx * 16 + (v.mvx >> 1), cx0 = mbx * 8 + (v.mvx >> 2);
            const int prow = (int)J.ref_pic[v.ref - 1] * ring_rows;
            const uint32_t b0 = smem_u32(sm.u.box + lane * 1280);
            tma_load_2d(b0, tm_l2, x0 & ~15, prow + mby * 16 + (v.mvy >> 1), bar);
            tma_load_3d(b0 + 640u, tm_c3, cx0 & ~15, 0, prow + H + mby * 8 + (v.mvy >> 2), bar);
            rect = pw.x & 0xFFFu;   // x/2 | y/2 << 4 | log2(w)-1 << 8 | log2(h)-1 << 10
            word = leaf_word(x0, cx0, v.mvx, v.mvy);
        }
        {
            const uint32_t par = phases & 1u;
            phases ^= 1u;
            uint32_t done, spins = 0;
            do {
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar), "r"(par) : "memory");
                if (!done && ++spins > (1u << 22)) __trap();
            } while (!done);
        }
#pragma unroll 1
        for (int p = 0; p < cnt; p++) {
            const uint32_t rc = __shfl_sync(0xffffffffu, rect, p), w = __shfl_sync(0xffffffffu, word, p);
            const int lx = (int)(rc & 15u) * 2, ly = (int)((rc >> 4) & 15u) * 2, lw = 2 << ((rc >> 8) & 3u), lh = 2 << ((rc >> 10) & 3u);
            // luma: this lane's pixels (8 * lhalf .. + 7, lrow); in 2-pixel cells, which of the four the leaf covers
            int lo = max(lx - 8 * lhalf, 0) >> 1, hi = min(lx + lw - 8 * lhalf, 8) >> 1;
            if ((unsigned)(lrow - ly) < (unsigned)lh && hi > lo) {
                const uint32_t cells = ((1u << hi) - 1u) & ~((1u << lo) - 1u);
                uint32_t a0, a1;
                v3_luma(sm.u.box + p * 1280, l_off + (w & 15u), (w >> 4) & 3u, a0, a1);
                // byte k of a word comes from the new value (selector k) where its cell is covered, else from the old one (4 + k)
                out.x = __byte_perm(a0, out.x, 0x7654u ^ ((cells & 1u) * 0x0044u + ((cells >> 1) & 1u) * 0x4400u));
                out.y = __byte_perm(a1, out.y, 0x7654u ^ (((cells >> 2) & 1u) * 0x0044u + ((cells >> 3) & 1u) * 0x4400u));
            }
            // chroma: pixels (4 * lhalf .. + 3, crow) of plane cpl; the leaf covers (lx/2 .. , ly/2 ..), at least one pixel
            lo = max((lx >> 1) - 4 * lhalf, 0); hi = min(((lx + lw) >> 1) - 4 * lhalf, 4);
            if ((unsigned)(crow - (ly >> 1)) < (unsigned)(lh >> 1) && hi > lo) {
                const uint32_t b = ((1u << hi) - 1u) & ~((1u << lo) - 1u);
                const uint32_t a = v3_chroma(sm.u.box + p * 1280 + 640, c_off + ((w >> 6) & 15u), (w >> 10) & 3u);
                out.z = __byte_perm(a, out.z, 0x7654u ^ (((b & 1u) | (b & 2u) << 3 | (b & 4u) << 6 | (b & 8u) << 9) * 4u));
            }
        }
    }
    out.w = phases;
    return out;
}

// tm_l2: the ring as a rank-2 u8 tensor (Stride, rows of all pictures), box 32x17; tm_c3: as a rank-3 tensor (Stride/2, 2, rows of
// all pictures) -- each row split into its U and V halves -- box 32x2x9.  A picture is `ring_rows` rows (luma, chroma, padding).
template <int LOG2S, int CHUNK, int CTAS>
__global__ void __launch_bounds__(V3_WARPS * 32, CTAS)
k_inter_v3(const DevJob* __restrict__ jobs, uint32_t n_chunks, uint32_t cpp, uint32_t cpp_magic, int mbw, uint32_t mbw_magic, int H, int ring_rows,
           uint32_t* __restrict__ ticket, uint32_t ticket_base, uint32_t exp_flags,
           const __grid_constant__ CUtensorMap tm_l2, const __grid_constant__ CUtensorMap tm_c3) {
    constexpr int RUN = V3_RUN, S = 1 << LOG2S;
    using Smem = V3Smem<CTAS>;
    constexpr uint32_t SLOTS = Smem::SLOTS, POOL = Smem::POOL;
    static_assert(CHUNK == 8 || CHUNK == 16, "a chunk is two or four runs");
    __shared__ __align__(128) Smem s_all[V3_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Smem& sm = s_all[warp];
    uint8_t* const box = sm.u.box;
    const size_t ysz = (size_t)S * H;
    if (lane < 2) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(sm.bar(lane))) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    uint32_t phases = 0;   // bit r: parity of the phase the barrier of round r completes next
    const int lrow = lane >> 1, lhalf = lane & 1;
    const int cpl = lane >> 4, crow = (lane >> 1) & 7;
    const int kc = lane & (CHUNK - 1), k4 = lane & 3;
    // per-lane constants of the box reads, the tile accesses and the transform passes
    const uint32_t l_off = (uint32_t)(lrow * 32 + lhalf * 8), c_off = (uint32_t)((crow * 2 + cpl) * 32 + lhalf * 4);
    const uint32_t tl_off = (uint32_t)(lrow * 16 + lhalf * 8), tc_off = (uint32_t)(256 + cpl * 64 + crow * 8 + lhalf * 4);
    const int g = lane >> 3, rr = lane & 7, i4 = rr & 3, s0 = (rr >> 2) * 2;
    // 8x8-transformed block: this lane's row rr, its 16-byte halves swapped for rows 4-7 (conflict-free row loads);
    // transposed stores go to column rr of rows 0-3 and column rr ^ 4 of rows 4-7
    const uint32_t p8_lo = (uint32_t)(rr * 8 + (rr & 4)), p8_hi = (uint32_t)(rr * 8 + (4 ^ (rr & 4)));
    // block transformed as four 4x4 units (elements unit * 16 + 4 * row + col): this lane's row i4 of units s0 and s0 + 1,
    // i.e. block rows 2 * s0 + (i4 >> 1) and + 2, columns 4 * (i4 & 1) ..; the same swap for rows 4-7
    const uint32_t r4 = (uint32_t)(2 * s0 + (i4 >> 1)), x4 = r4 & 4u;
    const uint32_t p4_lo = r4 * 8u + (((uint32_t)(4 * (i4 & 1))) ^ x4), p4_hi = p4_lo + 16u;
    // transposed stores of a 4x4 pass: element s0 * 16 + 4 * q + i4 -> row 2 * s0 + (q >> 1), column 4 * (q & 1) + i4
    const uint32_t t4_e = (uint32_t)(2 * s0) * 8u + ((uint32_t)i4 ^ x4), t4_o = (uint32_t)(2 * s0) * 8u + ((uint32_t)(4 + i4) ^ x4);

    uint32_t t = 0;
    if (lane == 0) asm volatile("atom.global.add.u32 %0, [%1], 1;" : "=r"(t) : "l"(ticket) : "memory");
    t = __shfl_sync(0xffffffffu, t, 0) - ticket_base;
    while (t < n_chunks) {
        uint32_t t_next = 0;   // requested now, looked at when this chunk is done
        if (lane == 0) asm volatile("atom.global.add.u32 %0, [%1], 1;" : "=r"(t_next) : "l"(ticket) : "memory");
        const uint32_t job = __umulhi(t, cpp_magic), chunk = t - job * cpp;
        const DevJob& J = jobs[job];
        const uint32_t n_mb = J.n_mb;
        if (J.n_intra != n_mb) {   // an I-picture has nothing for this kernel
            // ---- lane-parallel set-up: lane l (< CHUNK) looks after macroblock mbc + l ----
            const uint32_t mbc = chunk * CHUNK;
            if (exp_flags & 8u) {   // experiment (measured: slower, see prefetch_chunk_region): the region of picture 1 this chunk's windows can be expected in, as full lines into L2
                const int fy = (int)__umulhi(mbc, mbw_magic);
                prefetch_chunk_region<LOG2S>(J.ref[0], H, mbw, (int)mbc - fy * mbw, fy, (int)min((uint32_t)CHUNK, n_mb - mbc), lane);
            }
            const bool in = mbc + kc < n_mb;
            const uint32_t mbk = in ? mbc + kc : n_mb - 1;
            const uint4 d = __ldg(reinterpret_cast<const uint4*>(J.mbs + mbk));
            const uint32_t* const coefs = reinterpret_cast<const uint32_t*>(J.coefs);
            const uint32_t* const qtab = J.hdr->qtab;
            if (lane < 20) asm volatile("prefetch.global.L1 [%0];" :: "l"(qtab + lane * 4));   // the picture's 80 dequantisation words
            uint8_t* const dst = J.dst;
            const bool inter = in && !(d.x & 3u);
            const uint32_t n_parts_k = (d.x >> 2) & 127u;
            const uint32_t n_coef = inter ? (d.x >> 9) & 511u : 0u;
            const uint32_t bm = inter ? (d.x >> 18) & 63u : 0u;
            const uint32_t m8 = inter ? (((d.x >> MB_M8_LO) & 15u) | ((d.x >> MB_M8_HI) & 3u) << 4) & bm : 0u;
            const int mby = (int)__umulhi(mbk, mbw_magic), mbx = (int)mbk - mby * mbw;
            const int yoff = ((mby * 16) << LOG2S) + mbx * 16;
            // leaves 0 and 1: the inline copy of an unsplit macroblock's single leaf (info bit 28), else the records
            const bool inl = n_parts_k == 1u && (d.x & (1u << 28));
            uint2 pp0 = make_uint2(0u, 0u), pp1 = make_uint2(0u, 0u);
            if (inter && !inl && lane < CHUNK && n_parts_k <= 2u) {
                const uint2* pr = reinterpret_cast<const uint2*>(J.parts + d.y);
                pp0 = __ldg(pr);
                if (n_parts_k == 2u) pp1 = __ldg(pr + 1);
            }
            if (inl) pp0 = make_uint2((0xFu | (d.w >> 28) << 4) << 8 | (uint32_t)(((int)(d.w << 18)) >> 18) << 16, (uint32_t)(((int)(d.w << 4)) >> 18) & 0xFFFFu);
            const PartV v0 = part_of(pp0.x, pp0.y), v1 = part_of(pp1.x, pp1.y);
            // A leaf may come by box when every column ITS pixels need lies inside its own pixel row (flat addressing wraps
            // there, TMA zero-fills).  The box itself is the macroblock's 16x16 (+1) window displaced by the leaf's vector.
            const int x00 = mbx * 16 + (v0.mvx >> 1), cx00 = mbx * 8 + (v0.mvx >> 2);
            const int x01 = mbx * 16 + (v1.mvx >> 1), cx01 = mbx * 8 + (v1.mvx >> 2);
            auto row_ok = [&](uint32_t px, int xw, int cxw) {
                const int lx = (int)(px & 15u) * 2, lw = 2 << ((px >> 8) & 3u);
                return xw + lx >= 0 && xw + lx + lw + 1 <= S && cxw + (lx >> 1) >= 0 && cxw + (lx >> 1) + (lw >> 1) + 1 <= (S >> 1);
            };
            uint32_t need = 0;   // box slots wanted: 1 unsplit, 2 split once (each leaf fetches the whole window at its vector)
            if (inter && lane < CHUNK) {
                if (n_parts_k == 1u && row_ok(pp0.x, x00, cx00)) need = 1;
                else if (n_parts_k == 2u && row_ok(pp0.x, x00, cx00) && row_ok(pp1.x, x01, cx01)) need = 2;
            }
            // Slots in rounds: running sum over the four macroblocks of the run; the longest prefix that fits SLOTS is round 0
            // (boxes issued when the run starts), what follows is round 1, issued when its first macroblock's turn comes
            // (4 x 2 leaves <= 2 x SLOTS: two rounds always suffice).
            uint32_t incl = need;
            { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, 1); if (k4 >= 1) incl += u; }
            { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, 2); if (k4 >= 2) incl += u; }
            uint32_t r0 = incl <= SLOTS ? incl : 0u;   // slots round 0 takes: maximum over the run
            r0 = max(r0, __shfl_xor_sync(0xffffffffu, r0, 1));
            r0 = max(r0, __shfl_xor_sync(0xffffffffu, r0, 2));
            const uint32_t rnd = incl > SLOTS ? 1u : 0u;
            const uint32_t slot0 = (rnd ? incl - r0 : incl) - need;
            uint32_t what = V3_SKIP;
            if (inter) {
                what = (n_parts_k > 2u && !(exp_flags & 2u)) ? V3_MULTI : V3_LPL;
                if (need == 1u) what = V3_BOX1;
                else if (need == 2u) what = V3_BOX2 | (((pp0.x | pp1.x) & 15u) ? 1u : 0u) | ((pp0.x & 255u) ? 2u : 0u);
            }
            const uint32_t mcw = leaf_word(x00, cx00, v0.mvx, v0.mvy) | leaf_word(x01, cx01, v1.mvx, v1.mvy) << 12 | what << 24 | slot0 << 27;
            // first tensor row of the pictures the leaves point into
            int prow0 = 0, prow1 = 0;
            if (need) prow0 = (int)J.ref_pic[v0.ref - 1] * ring_rows;
            if (need == 2u) prow1 = (int)J.ref_pic[v1.ref - 1] * ring_rows;

#pragma unroll 1
            for (int r = 0; r < CHUNK / RUN; r++) {
                if (mbc + (uint32_t)(RUN * r) >= n_mb) break;
                const int l0 = RUN * r;              // lanes l0 .. l0+3 hold this run's macroblocks
                const bool mine = (kc >> 2) == r && lane < CHUNK;
                // bytes each round brings: known to the run's last lane (its running sum is the run's total)
                const uint32_t tx0 = r0 * (TMA_BYTES_L + TMA_BYTES_C4), tx1 = (incl - r0) * (TMA_BYTES_L + TMA_BYTES_C4);
                auto issue = [&](uint32_t round) {
                    if (lane == l0 + RUN - 1) {
                        const uint32_t tx = round ? tx1 : tx0;
                        if (tx) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(sm.bar(round))), "r"(tx) : "memory");
                    }
                    // everyone is done with what the slots held; that generic-proxy traffic is ordered before the boxes
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (mine && need && rnd == round) {
                        const uint32_t bar = smem_u32(sm.bar(round));
                        const uint32_t b0 = smem_u32(box + slot0 * 1280u);
                        tma_load_2d(b0, &tm_l2, x00 & ~15, prow0 + mby * 16 + (v0.mvy >> 1), bar);
                        tma_load_3d(b0 + 640u, &tm_c3, cx00 & ~15, 0, prow0 + H + mby * 8 + (v0.mvy >> 2), bar);
                        if (need == 2u) {
                            tma_load_2d(b0 + 1280u, &tm_l2, x01 & ~15, prow1 + mby * 16 + (v1.mvy >> 1), bar);
                            tma_load_3d(b0 + 1920u, &tm_c3, cx01 & ~15, 0, prow1 + H + mby * 8 + (v1.mvy >> 2), bar);
                        }
                    }
                };
                // first macroblock of round 1 (RUN: none), and whether each round has boxes at all
                const uint32_t late = __ballot_sync(0xffffffffu, mine && rnd && need) >> l0;
                const int first_late = late ? __ffs((int)late) - 1 : RUN;
                bool fresh0 = __shfl_sync(0xffffffffu, tx0, l0 + RUN - 1) != 0u;   // round 0 has boxes nobody has waited for yet
                // the run's coefficient records: one range of the picture's array; the first 64 are fetched now
                const uint32_t jmin = __reduce_min_sync(0xffffffffu, mine && n_coef ? d.z : 0xffffffffu);
                const uint32_t jmax = __reduce_max_sync(0xffffffffu, mine && n_coef ? d.z + n_coef : 0u);
                const uint32_t ntot = (jmax > jmin && !(exp_flags & 1u)) ? jmax - jmin : 0u;   // (exp_flags: timing experiments only)
                const uint32_t* cf = coefs + (ntot ? jmin : 0u);
                uint32_t ca = 0, cb = 0;
                if ((uint32_t)lane < ntot) ca = __ldg(cf + lane);
                if ((uint32_t)lane + 32u < ntot) cb = __ldg(cf + 32 + lane);
                // coded blocks of the run, bit 6 * macroblock + block; which of them are transformed as one 8x8
                const uint32_t CM = __reduce_or_sync(0xffffffffu, mine ? bm << (6 * k4) : 0u);
                const uint32_t M8 = __reduce_or_sync(0xffffffffu, mine ? m8 << (6 * k4) : 0u);

                // ---- prediction, one macroblock at a time, into its tile ----
#pragma unroll 1
                for (int i = 0; i < RUN; i++) {
                    if (i == 0 || i == first_late) issue(i ? 1u : 0u);   // (one call site: the waterfall code of four TMA instructions exists once)
                    const uint32_t w = __shfl_sync(0xffffffffu, mcw, l0 + i);
                    const uint32_t todo = (w >> 24) & 7u;
                    if (todo == V3_SKIP || todo == V3_MULTI) continue;   // intra (k_intra's job), past the picture's last macroblock, or later (below)
                    uint32_t y0, y1, c0;
                    if (todo >= V3_BOX1) {
                        if (i == first_late || fresh0) {   // the first macroblock that needs a round's boxes waits for the round
                            const uint32_t round = i == first_late ? 1u : 0u;
                            if (!round) fresh0 = false;
                            const uint32_t bar = smem_u32(sm.bar(round)), par = (phases >> round) & 1u;
                            phases ^= 1u << round;
                            uint32_t done, spins = 0;
                            do {
                                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar), "r"(par) : "memory");
                                if (!done && ++spins > (1u << 22)) __trap();   // a box that never arrives must not hang the device
                            } while (!done);
                        }
                        // which leaf covers this lane's pixels: split top/bottom -> by row, left/right -> by half (luma 8, chroma 4 pixels per lane)
                        uint32_t sl = 0, sc = 0;
                        if (todo >= V3_BOX2) {   // leaf 0 is the top / left half unless the records come in the other order
                            const uint32_t sw = (todo >> 1) & 1u;
                            sl = ((todo & 1u) ? (uint32_t)lhalf : (uint32_t)(lrow >> 3)) ^ sw;
                            sc = ((todo & 1u) ? (uint32_t)lhalf : (uint32_t)(crow >> 2)) ^ sw;
                        }
                        const uint8_t* const b0 = box + ((w >> 27) & 7u) * 1280u;
                        const uint32_t wl = sl ? w >> 12 : w, wc = sc ? w >> 12 : w;
                        v3_luma(b0 + sl * 1280u, l_off + (wl & 15u), (wl >> 4) & 3u, y0, y1);
                        c0 = v3_chroma(b0 + 640u + sc * 1280u, c_off + ((wc >> 6) & 15u), (wc >> 10) & 3u);
                    } else {
                        // load-per-lane path: windows that leave their pixel row, macroblocks of more than two leaves
                        const uint32_t dx = __shfl_sync(0xffffffffu, d.x, l0 + i), dy = __shfl_sync(0xffffffffu, d.y, l0 + i);
                        const int yo = __shfl_sync(0xffffffffu, yoff, l0 + i);
                        const uint3 px = v3_lpl<LOG2S>(J, (int)((dx >> 2) & 127u), dy, yo, H, sm.tile[i], lane);
                        y0 = px.x; y1 = px.y; c0 = px.z;
                    }
                    uint8_t* tl = sm.tile[i];
                    *reinterpret_cast<uint2*>(tl + tl_off) = make_uint2(y0, y1);
                    *reinterpret_cast<uint32_t*>(tl + tc_off) = c0;
                }
                // macroblocks of more than two leaves, now that every box slot is free
                for (uint32_t mm = __ballot_sync(0xffffffffu, mine && what == V3_MULTI) >> l0; mm; mm &= mm - 1u) {
                    const int i = __ffs((int)mm) - 1;
                    const uint32_t dx = __shfl_sync(0xffffffffu, d.x, l0 + i), dy = __shfl_sync(0xffffffffu, d.y, l0 + i);
                    const int bx = __shfl_sync(0xffffffffu, mbx, l0 + i), by = __shfl_sync(0xffffffffu, mby, l0 + i);
                    const uint4 px = v3_multi<LOG2S, CTAS>(J, sm, &tm_l2, &tm_c3, (int)((dx >> 2) & 127u), dy, bx, by, H, ring_rows, phases, sm.tile[i], lane);
                    phases = px.w;
                    uint8_t* tl = sm.tile[i];
                    *reinterpret_cast<uint2*>(tl + tl_off) = make_uint2(px.x, px.y);
                    *reinterpret_cast<uint32_t*>(tl + tc_off) = px.z;
                }
                __syncwarp();   // tiles complete; the boxes are dead, the coefficient pool may overwrite them

                if (ntot) {
                    // The pool holds POOL blocks; a run that codes more (at most 24) is worked off as two pairs of macroblocks.
                    const uint32_t nall = (uint32_t)__popc(CM);
                    const uint32_t halves = nall > POOL ? 2u : 1u;
#pragma unroll 1
                    for (uint32_t hv = 0; hv < halves; hv++) {
                        const uint32_t sel = halves == 1u ? 0xFFFFFFu : (hv ? 0xFFF000u : 0x000FFFu);
                        const uint32_t cm = CM & sel, c8 = M8 & sel, c4m = cm & ~c8;
                        const uint32_t ns = (uint32_t)__popc(cm), n8 = (uint32_t)__popc(c8);
                        if (hv) __syncwarp();   // the first pair's passes are done with the pool
                        {
                            int4* z = reinterpret_cast<int4*>(sm.u.pool);
                            for (uint32_t q = lane; q < ns * 18u; q += 32u) z[q] = make_int4(0, 0, 0, 0);
                        }
                        uint32_t dcm = 0;   // 8x8-transformed blocks whose last coefficient sits at scan position 0 (filled in by the scatter)
                        __syncwarp();
                        // ---- dequantise into the pool (MD:3424-3429): pooled block p lives at pool + 72 * p words, the blocks
                        // transformed as one 8x8 first ----
                        // whose record: the parser tags every record with its macroblock's index & 3 (mobi_coef.blk bits 3-4), and
                        // a run is four macroblocks aligned to four.  Records of intra macroblocks lying inside the range (k_intra's),
                        // or naming a block their macroblock does not code, find no bit in the mask and are passed over.
                        auto scatter = [&](uint32_t c) {
                            const uint32_t bit = ((c >> 27) & 3u) * 6u + ((c >> 24) & 7u);
                            if ((cm >> bit) & 1u) {
                                const int level = (int)(int16_t)(c & 0xFFFFu);
                                const uint32_t pos = (c >> 16) & 63u, sub = (c >> 22) & 3u;
                                const bool is8 = (c8 >> bit) & 1u;
                                const uint32_t wq = __ldg(qtab + (is8 ? pos : 64u + (pos & 15u)));
                                const uint32_t lt = (1u << bit) - 1u;
                                const uint32_t p = is8 ? (uint32_t)__popc(c8 & lt) : n8 + (uint32_t)__popc(c4m & lt);
                                const uint32_t e = is8 ? (wq & 63u) : sub * 16u + (wq & 15u);
                                sm.u.pool[p * V3_PSTRIDE + (e ^ ((e >> 3) & 4u))] = (int)(wq >> 8) * level;
                                // an 8x8-transformed block whose LAST record sits at scan position 0 holds nothing but its DC (MD:2938)
                                if (is8 && (c & (1u << 30)) && pos == 0u) dcm |= 1u << bit;
                            }
                        };
                        if ((uint32_t)lane < ntot) scatter(ca);
                        if ((uint32_t)lane + 32u < ntot) scatter(cb);
                        for (uint32_t j = 64u + (uint32_t)lane; j < ntot; j += 32u) scatter(__ldg(cf + j));
                        dcm = __reduce_or_sync(0xffffffffu, dcm);
                        // Visiting order of the passes: blocks that need the 8x8 transform, blocks transformed as four 4x4, DC-only
                        // 8x8 blocks (no transform at all) -- so that a pass is of one kind except at the two boundaries.
                        const uint32_t c8t = c8 & ~dcm;
                        const uint32_t na = (uint32_t)__popc(c8t), nab = na + (ns - n8);
                        if (lane < 6 * RUN && ((cm >> lane) & 1u)) {   // one lane per (macroblock, block) of the run: lane == bit index
                            const uint32_t kk = (uint32_t)lane / 6u, b = (uint32_t)lane - 6u * kk;
                            const uint32_t lt = (1u << lane) - 1u;
                            const bool t8 = (c8 >> lane) & 1u, dc = (dcm >> lane) & 1u;
                            const uint32_t slot = t8 ? (uint32_t)__popc(c8 & lt) : n8 + (uint32_t)__popc(c4m & lt);
                            const uint32_t vp = dc ? nab + (uint32_t)__popc(dcm & lt) : t8 ? (uint32_t)__popc(c8t & lt) : na + (uint32_t)__popc(c4m & lt);
                            const uint32_t toff = kk * (uint32_t)Smem::TILE + (b < 4u ? ((b >> 1) * 8u) * 16u + (b & 1u) * 8u : 256u + (b - 4u) * 64u);
                            sm.slotinfo(vp) = toff | (b < 4u ? 0u : 1u << 11) | slot << 12;
                        }
                        __syncwarp();

                        // ---- inverse transforms: eight lanes per pooled block (one row each), four blocks per pass ----
#pragma unroll 1
                        for (uint32_t base = 0; base < ns; base += 4u) {
                            const uint32_t idx = base + (uint32_t)g;
                            const bool has = idx < ns;
                            const uint32_t info = sm.slotinfo(has ? idx : base);
                            int32_t* const B = sm.u.pool + ((info >> 12) & 31u) * V3_PSTRIDE;
                            uint8_t* const tp = &sm.tile[0][0] + (info & 0x7FFu) + rr * ((info >> 11) & 1u ? 8 : 16);
                            int32_t in[8], v[8];
                            if (base >= nab) {
                                // DC-only blocks: the residual is the constant (dc + 32) >> 6 (IDCT1Px8 MD:3710-3725), no transform
                                if (has) {
                                    const int r = (B[0] + 32) >> 6;
                                    uint2 px = *reinterpret_cast<uint2*>(tp);
                                    px.x = addsat4c(px.x, r); px.y = addsat4c(px.y, r);
                                    *reinterpret_cast<uint2*>(tp) = px;
                                }
                                continue;
                            }
                            // One body for both transforms (the instruction cache is what this kernel lives on): passes are all-8x8 or
                            // all-4x4 except at the two boundaries of the visiting order, so the branches on is8 are uniform almost always
                            // (a DC-only block that shares a pass with 4x4 blocks simply takes the full 8x8 transform).
                            const bool is8 = idx < na || idx >= nab;
                            const uint32_t plo = is8 ? p8_lo : p4_lo, phi = is8 ? p8_hi : p4_hi;
                            {
                                const int4 lo = *reinterpret_cast<const int4*>(B + plo), hi = *reinterpret_cast<const int4*>(B + phi);
                                in[0] = lo.x; in[1] = lo.y; in[2] = lo.z; in[3] = lo.w; in[4] = hi.x; in[5] = hi.y; in[6] = hi.z; in[7] = hi.w;
                            }
                            if (is8) { if (rr == 0) in[0] += 32; bfly8(in, v); }
                            else { if (i4 == 0) { in[0] += 32; in[4] += 32; } bfly4(in, v); bfly4(in + 4, v + 4); }
                            __syncwarp();
                            if (is8) {
#pragma unroll
                                for (int q = 0; q < 4; q++) { B[q * 8 + rr] = v[q]; B[(q + 4) * 8 + (rr ^ 4)] = v[q + 4]; }
                            } else {
                                B[t4_e] = v[0]; B[t4_o] = v[1]; B[t4_e + 8] = v[2]; B[t4_o + 8] = v[3];
                                B[t4_e + 16] = v[4]; B[t4_o + 16] = v[5]; B[t4_e + 24] = v[6]; B[t4_o + 24] = v[7];
                            }
                            __syncwarp();
                            {
                                const int4 lo = *reinterpret_cast<const int4*>(B + plo), hi = *reinterpret_cast<const int4*>(B + phi);
                                in[0] = lo.x; in[1] = lo.y; in[2] = lo.z; in[3] = lo.w; in[4] = hi.x; in[5] = hi.y; in[6] = hi.z; in[7] = hi.w;
                            }
                            if (is8) bfly8(in, v); else { bfly4(in, v); bfly4(in + 4, v + 4); }
                            // either way the lane now holds the residuals of row rr, columns 0..7 of its block: add onto the prediction
                            if (has) {
                                uint2 px = *reinterpret_cast<uint2*>(tp);
                                px.x = addsat4(px.x, v[0], v[1], v[2], v[3]);
                                px.y = addsat4(px.y, v[4], v[5], v[6], v[7]);
                                *reinterpret_cast<uint2*>(tp) = px;
                            }
                        }
                    }
                    __syncwarp();
                }

                // ---- the tiles leave, lane-parallel over the run again: 16 luma bytes / 8 chroma bytes per lane and round ----
                const int yoff_k = __shfl_sync(0xffffffffu, yoff, l0 + k4);
                const bool inter_k = ((__shfl_sync(0xffffffffu, mcw, l0 + k4) >> 24) & 7u) != V3_SKIP;
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
