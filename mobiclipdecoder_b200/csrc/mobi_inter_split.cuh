// k_mc + k_res -- the inter path as TWO kernels, included by mobi_kernels.cu inside its anonymous namespace.
// "MD:n" = LibMobiclip/Codec/Mobiclip/MobiclipDecoder.cs:n.
//
// Why two.  The fused kernels (k_inter_chunk, k_inter_v3) sit at 0.25-0.27 ms per 360 k macroblocks whatever is done to
// their instruction count (415 ... 505 per macroblock, profiles/r02*): with 72-80 registers and 8-9 KB of shared memory per
// warp only 24-28 warps fit an SM, and what a warp spends its life on is WAITING -- 30 % of the stall samples of the fused
// kernel, 64 % once the residual work is taken out, are long-scoreboard waits (descriptor -> leaf records -> TMA boxes, the
// load-per-lane path's dependent trips to memory).  Taking the residual out of the kernel shortens it by a quarter, not by the
// 150 instructions per macroblock it removes: latency-bound, not issue-bound.  The cure for that is warps, and the two halves
// want different resources:
//   k_mc   motion compensation only (CopyBlock MD:418-456): TMA boxes -> half-pel filter on packed bytes -> pixels straight
//          from registers to the picture.  No coefficient pool, no tiles: 5.4 KB of shared memory and <= 48 registers per
//          warp, nine 4-warp CTAs per SM.  Boxes of macroblock i+1 are in flight while macroblock i is filtered.
//   k_res  dequantisation + inverse transforms (MD:3424-3798) added IN PLACE onto the prediction k_mc left in the picture
//          (exactly what the reference does: MC into Dst, then the IDCT adds onto Dst): coefficient pool of 16 blocks, pixels
//          read and written as the 8-byte rows of each coded 8x8 block.
// The price is one more pass over the coded blocks' pixels (1.8 blocks of 64 bytes per macroblock on the bench mix, mostly
// still in L2) -- which is also how BASELINE.json / north_star name them: "the MC and IDCT kernels".

constexpr int MC_WARPS = 4, MC_CHUNK = 16;
constexpr uint32_t MCK_SKIP = 0, MCK_LPL = 1, MCK_BOX1 = 2, MCK_MULTI = 3, MCK_BOX2 = 4;   // BOX2 | 1: split left / right; | 2: leaf 0 is the bottom / right half

struct McSmem {
    uint8_t box[2][2][1280];     // [macroblock parity][leaf]: 32x17 luma box at +0 (544 bytes), 32x2x9 chroma box at +640 (576 bytes)
    uint4 tab[MC_CHUNK][2];      // per macroblock of the chunk: what to fetch and how to read it (see k_mc)
    uint64_t bar[2];             // one mbarrier per macroblock parity
    uint8_t pad[128 - 16];
};
static_assert(sizeof(McSmem) == 5120 + 512 + 128 && sizeof(McSmem) % 128 == 0, "TMA destinations must stay 128-byte aligned");
constexpr int MC_CTAS = 9, RES_CTAS = 10;   // CTAs per SM the kernels are sized for (shared memory and registers)
static_assert((sizeof(McSmem) * MC_WARPS + 1024) * MC_CTAS <= 233472, "nine 4-warp CTAs per SM");

template <int LOG2S>
__global__ void __launch_bounds__(MC_WARPS * 32, MC_CTAS)
k_mc(const DevJob* __restrict__ jobs, uint32_t n_chunks, uint32_t cpp, uint32_t cpp_magic, int mbw, uint32_t mbw_magic, int H, int ring_rows,
     uint32_t* __restrict__ ticket, uint32_t ticket_base, uint32_t exp_flags,
     const __grid_constant__ CUtensorMap tm_l2, const __grid_constant__ CUtensorMap tm_c3) {
    constexpr int S = 1 << LOG2S;
    __shared__ __align__(128) McSmem s_all[MC_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    McSmem& sm = s_all[warp];
    const size_t ysz = (size_t)S * H;
    if (lane < 2) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&sm.bar[lane])) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    uint32_t phases = 0;   // bit s: parity of the phase barrier s completes next
    const int lrow = lane >> 1, lhalf = lane & 1;
    const int cpl = lane >> 4, crow = (lane >> 1) & 7;
    const uint32_t l_off = (uint32_t)(lrow * 32 + lhalf * 8), c_off = (uint32_t)((crow * 2 + cpl) * 32 + lhalf * 4);
    const int py_off = (lrow << LOG2S) + lhalf * 8, pc_off = (cpl ? (S >> 1) : 0) + (crow << LOG2S) + lhalf * 4;

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
            const uint32_t mbc = chunk * MC_CHUNK;
            const int n_here = (int)min((uint32_t)MC_CHUNK, n_mb - mbc);
            uint8_t* const dst = J.dst;
            if (exp_flags & 8u) {   // experiment (measured: slower, see prefetch_chunk_region): the region of picture 1 this chunk's windows can be expected in, as full lines into L2
                const int fy = (int)__umulhi(mbc, mbw_magic);
                prefetch_chunk_region<LOG2S>(J.ref[0], H, mbw, (int)mbc - fy * mbw, fy, n_here, lane);
            }
            // ---- lane-parallel set-up: lane l (< 16) works out macroblock mbc + l and files it in tab[l] ----
            //   [0].x  leaf 0 (column of its window inside the 16-byte-aligned box 4, luma half-pel phase 2, the same for chroma 4 + 2)
            //          | leaf 1 << 12 | kind << 24
            //   [0].y  box columns of leaf 0: luma (s16) | chroma << 16      [0].z / .w  its first luma / chroma tensor row
            //   [1].x .. .z  the same for leaf 1 (more than two leaves, load-per-lane path: .x = index of the first leaf record)
            //   [1].w  luma offset of the macroblock inside the picture >> 4 | leaves << 20
            if (lane < MC_CHUNK) {
                uint4 e0 = make_uint4(0u, 0u, 0u, 0u), e1 = e0;
                if (lane < n_here) {
                    const uint32_t mbk = mbc + (uint32_t)lane;
                    const uint4 d = __ldg(reinterpret_cast<const uint4*>(J.mbs + mbk));
                    if (!(d.x & 3u)) {
                        const uint32_t n_parts = (d.x >> 2) & 127u;
                        const int mby = (int)__umulhi(mbk, mbw_magic), mbx = (int)mbk - mby * mbw;
                        const bool inl = n_parts == 1u && (d.x & (1u << 28));
                        uint2 pp0 = make_uint2(0u, 0u), pp1 = make_uint2(0u, 0u);
                        if (!inl && n_parts <= 2u) {
                            const uint2* pr = reinterpret_cast<const uint2*>(J.parts + d.y);
                            pp0 = __ldg(pr);
                            if (n_parts == 2u) pp1 = __ldg(pr + 1);
                        }
                        if (inl) pp0 = make_uint2((0xFu | (d.w >> 28) << 4) << 8 | (uint32_t)(((int)(d.w << 18)) >> 18) << 16, (uint32_t)(((int)(d.w << 4)) >> 18) & 0xFFFFu);
                        const PartV v0 = part_of(pp0.x, pp0.y), v1 = part_of(pp1.x, pp1.y);
                        // A leaf may come by box when every column ITS pixels need lies inside its own pixel row (flat addressing
                        // wraps there, TMA zero-fills).  The box is the macroblock's 16x16 (+1) window displaced by the leaf's vector.
                        const int x00 = mbx * 16 + (v0.mvx >> 1), cx00 = mbx * 8 + (v0.mvx >> 2);
                        const int x01 = mbx * 16 + (v1.mvx >> 1), cx01 = mbx * 8 + (v1.mvx >> 2);
                        auto row_ok = [&](uint32_t px, int xw, int cxw) {
                            const int lx = (int)(px & 15u) * 2, lw = 2 << ((px >> 8) & 3u);
                            return xw + lx >= 0 && xw + lx + lw + 1 <= S && cxw + (lx >> 1) >= 0 && cxw + (lx >> 1) + (lw >> 1) + 1 <= (S >> 1);
                        };
                        uint32_t kind = MCK_LPL;
                        if (n_parts == 1u && row_ok(pp0.x, x00, cx00)) kind = MCK_BOX1;
                        else if (n_parts == 2u && row_ok(pp0.x, x00, cx00) && row_ok(pp1.x, x01, cx01))
                            kind = MCK_BOX2 | (((pp0.x | pp1.x) & 15u) ? 1u : 0u) | ((pp0.x & 255u) ? 2u : 0u);
                        else if (n_parts > 2u) {   // more than two leaves: by boxes, two leaves at a time, if every leaf's columns stay inside its row
                            const uint2* pr = reinterpret_cast<const uint2*>(J.parts + d.y);
                            bool safe = true;
#pragma unroll 1
                            for (uint32_t p = 0; p < n_parts; p++) {
                                const uint2 pw = __ldg(pr + p);
                                const int mvx = (int)(int16_t)(pw.x >> 16);
                                safe = safe && row_ok(pw.x, mbx * 16 + (mvx >> 1), mbx * 8 + (mvx >> 2));
                            }
                            if (safe) kind = MCK_MULTI;
                        }
                        e0.x = leaf_word(x00, cx00, v0.mvx, v0.mvy) | leaf_word(x01, cx01, v1.mvx, v1.mvy) << 12 | kind << 24;
                        if (kind >= MCK_BOX1 && kind != MCK_MULTI) {
                            const int prow0 = (int)J.ref_pic[v0.ref - 1] * ring_rows;
                            e0.y = (uint32_t)((x00 & ~15) & 0xFFFF) | (uint32_t)(cx00 & ~15) << 16;
                            e0.z = (uint32_t)(prow0 + mby * 16 + (v0.mvy >> 1)); e0.w = (uint32_t)(prow0 + H + mby * 8 + (v0.mvy >> 2));
                            if (kind >= MCK_BOX2) {
                                const int prow1 = (int)J.ref_pic[v1.ref - 1] * ring_rows;
                                e1.x = (uint32_t)((x01 & ~15) & 0xFFFF) | (uint32_t)(cx01 & ~15) << 16;
                                e1.y = (uint32_t)(prow1 + mby * 16 + (v1.mvy >> 1)); e1.z = (uint32_t)(prow1 + H + mby * 8 + (v1.mvy >> 2));
                            }
                        } else e1.x = d.y;
                        e1.w = (uint32_t)((((mby * 16) << LOG2S) + mbx * 16) >> 4) | n_parts << 20;
                    }
                }
                sm.tab[lane][0] = e0; sm.tab[lane][1] = e1;
            }
            __syncwarp();
            // One leaf's two boxes (32x17 luma, 32x2x9 U/V) into slot `slot` of box set `set`, by the calling lane.
            auto fetch_leaf = [&](uint32_t set, uint32_t slot, int xl, int rowl, int xc, int rowc) {
                const uint32_t bar = smem_u32(&sm.bar[set]), b0 = smem_u32(sm.box[set][slot]);
                tma_load_2d(b0, &tm_l2, xl, rowl, bar);
                if (!(exp_flags & 4u)) tma_load_3d(b0 + 640u, &tm_c3, xc, 0, rowc, bar);   // (exp_flags: timing experiments only)
            };
            // Leaves first .. first + 1 (as many as the macroblock has) of a macroblock of more than two leaves: lane p fetches leaf
            // first + p.  Returns, in lanes 0 and 1, what the merge needs of the leaf: rectangle (12 bits) | leaf word << 12.
            auto fetch_pair = [&](uint32_t set, const uint2* sp, int first, int n, int yo) -> uint32_t {
                uint32_t info = 0;
                if (lane < 2 && first + lane < n) {
                    const uint2 pw = __ldg(sp + first + lane);
                    const PartV v = part_of(pw.x, pw.y);
                    const int mbx16 = yo & (S - 1), mby16 = yo >> LOG2S;
                    const int x0 = mbx16 + (v.mvx >> 1), cx0 = (mbx16 >> 1) + (v.mvx >> 2);
                    const int prow = (int)J.ref_pic[v.ref - 1] * ring_rows;
                    fetch_leaf(set, (uint32_t)lane, x0 & ~15, prow + mby16 + (v.mvy >> 1), cx0 & ~15, prow + H + (mby16 >> 1) + (v.mvy >> 2));
                    info = (pw.x & 0xFFFu) | leaf_word(x0, cx0, v.mvx, v.mvy) << 12;
                }
                return info;
            };
            // Boxes of macroblock j into box set j & 1 (for a macroblock of more than two leaves: its first two), one leaf per lane.
            auto issue = [&](int j) -> uint32_t {
                // Everyone is done with what the set held (macroblock j - 2): its box reads returned their data before the pixels
                // computed from them were stored, in program order before this point -- nothing is left in flight that the boxes
                // could overtake.  (A fence.proxy.async here compiles to MEMBAR.ALL.CTA, which also waits for the pixel stores of
                // the macroblock just finished: 12 % of the kernel's stall samples.)
                __syncwarp();
                const uint4 e0 = sm.tab[j][0], e1 = sm.tab[j][1];
                const uint32_t kind = (e0.x >> 24) & 7u, set = (uint32_t)j & 1u;
                if (kind < MCK_BOX1) return 0u;
                const uint32_t n = kind == MCK_MULTI ? 2u : kind >= MCK_BOX2 ? 2u : 1u;
                if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&sm.bar[set])), "r"(n * (TMA_BYTES_L + ((exp_flags & 4u) ? 0u : TMA_BYTES_C4))) : "memory");
                if (kind == MCK_MULTI) return fetch_pair(set, reinterpret_cast<const uint2*>(J.parts + e1.x), 0, (int)(e1.w >> 20), (int)(e1.w & 0xFFFFFu) << 4);
                if (lane == 0) fetch_leaf(set, 0u, (int)(int16_t)(e0.y & 0xFFFFu), (int)e0.z, (int)e0.y >> 16, (int)e0.w);
                else if (lane == 1 && n == 2u) fetch_leaf(set, 1u, (int)(int16_t)(e1.x & 0xFFFFu), (int)e1.y, (int)e1.x >> 16, (int)e1.z);
                return 0u;
            };
            auto wait = [&](uint32_t set) {
                const uint32_t bar = smem_u32(&sm.bar[set]), par = (phases >> set) & 1u;
                phases ^= 1u << set;
                uint32_t done, spins = 0;
                do {
                    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar), "r"(par) : "memory");
                    if (!done && ++spins > (1u << 22)) __trap();   // a box that never arrives must not hang the device
                } while (!done);
            };
            uint32_t pair_next = issue(0), pair_cur;   // (what fetch_pair returned for the macroblock in hand / the next one)
#pragma unroll 1
            for (int i = 0; i < n_here; i++) {
                pair_cur = pair_next;
                if (i + 1 < n_here) pair_next = issue(i + 1);
                const uint32_t w = sm.tab[i][0].x;
                const uint32_t kind = (w >> 24) & 7u;
                if (kind == MCK_SKIP) continue;   // intra: k_intra's job
                const uint4 e1 = sm.tab[i][1];
                const int yo = (int)(e1.w & 0xFFFFFu) << 4;
                const uint32_t s = (uint32_t)i & 1u;
                uint32_t y0, y1, c0;
                if (kind == MCK_MULTI) {
                    // More than two leaves: through the set's two slots, two leaves at a time (the first two came with the prefetch).
                    // Every lane merges, leaf by leaf, the pixels of its eight luma / four chroma positions the leaf covers (leaves
                    // go down to 2x2: a lane's pixels may belong to four of them).
                    const int n = (int)(e1.w >> 20);
                    const uint2* const sp = reinterpret_cast<const uint2*>(J.parts + e1.x);
                    y0 = y1 = c0 = 0u;
#pragma unroll 1
                    for (int g0 = 0; g0 < n; g0 += 2) {
                        if (g0) {
                            __syncwarp();
                            if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&sm.bar[s])), "r"((uint32_t)min(2, n - g0) * (TMA_BYTES_L + ((exp_flags & 4u) ? 0u : TMA_BYTES_C4))) : "memory");
                            pair_cur = fetch_pair(s, sp, g0, n, yo);
                        }
                        wait(s);
#pragma unroll 1
                        for (int p = 0; p < min(2, n - g0); p++) {
                            const uint32_t inf = __shfl_sync(0xffffffffu, pair_cur, p), lw_ = inf >> 12;
                            const int lx = (int)(inf & 15u) * 2, ly = (int)((inf >> 4) & 15u) * 2, lw = 2 << ((inf >> 8) & 3u), lh = 2 << ((inf >> 10) & 3u);
                            // luma: this lane's pixels (8 * lhalf .. + 7, lrow); in 2-pixel cells, which of the four the leaf covers
                            int lo = max(lx - 8 * lhalf, 0) >> 1, hi = min(lx + lw - 8 * lhalf, 8) >> 1;
                            if ((unsigned)(lrow - ly) < (unsigned)lh && hi > lo) {
                                const uint32_t cells = ((1u << hi) - 1u) & ~((1u << lo) - 1u);
                                uint32_t a0, a1;
                                v3_luma(sm.box[s][p], l_off + (lw_ & 15u), (lw_ >> 4) & 3u, a0, a1);
                                // byte k of a word comes from the new value (selector k) where its cell is covered, else from the old one (4 + k)
                                y0 = __byte_perm(a0, y0, 0x7654u ^ ((cells & 1u) * 0x0044u + ((cells >> 1) & 1u) * 0x4400u));
                                y1 = __byte_perm(a1, y1, 0x7654u ^ (((cells >> 2) & 1u) * 0x0044u + ((cells >> 3) & 1u) * 0x4400u));
                            }
                            // chroma: pixels (4 * lhalf .. + 3, crow) of plane cpl; the leaf covers (lx/2 .., ly/2 ..), at least one pixel
                            lo = max((lx >> 1) - 4 * lhalf, 0); hi = min(((lx + lw) >> 1) - 4 * lhalf, 4);
                            if ((unsigned)(crow - (ly >> 1)) < (unsigned)(lh >> 1) && hi > lo) {
                                const uint32_t b = ((1u << hi) - 1u) & ~((1u << lo) - 1u);
                                const uint32_t a = v3_chroma(sm.box[s][p] + 640, c_off + ((lw_ >> 6) & 15u), (lw_ >> 10) & 3u);
                                c0 = __byte_perm(a, c0, 0x7654u ^ (((b & 1u) | (b & 2u) << 3 | (b & 4u) << 6 | (b & 8u) << 9) * 4u));
                            }
                        }
                    }
                } else if (kind >= MCK_BOX1) {
                    wait(s);
                    // which leaf covers this lane's pixels: split top/bottom -> by row, left/right -> by half (luma 8, chroma 4 pixels per lane)
                    uint32_t sl = 0, sc = 0;
                    if (kind >= MCK_BOX2) {   // leaf 0 is the top / left half unless the records come in the other order
                        const uint32_t sw = (kind >> 1) & 1u;
                        sl = ((kind & 1u) ? (uint32_t)lhalf : (uint32_t)(lrow >> 3)) ^ sw;
                        sc = ((kind & 1u) ? (uint32_t)lhalf : (uint32_t)(crow >> 2)) ^ sw;
                    }
                    const uint32_t wl = sl ? w >> 12 : w, wc = sc ? w >> 12 : w;
                    v3_luma(sm.box[s][sl], l_off + (wl & 15u), (wl >> 4) & 3u, y0, y1);
                    c0 = v3_chroma(sm.box[s][sc] + 640, c_off + ((wc >> 6) & 15u), (wc >> 10) & 3u);
                } else {
                    // windows that leave their pixel row: load per lane.  The 64-byte partition map goes through the box set of this
                    // macroblock's parity, which holds nothing (no boxes were issued for it).
                    const uint3 px = v3_lpl<LOG2S>(J, (int)(e1.w >> 20), e1.x, yo, H, sm.box[s][0], lane);
                    y0 = px.x; y1 = px.y; c0 = px.z;
                }
                *reinterpret_cast<uint2*>(dst + yo + py_off) = make_uint2(y0, y1);
                *reinterpret_cast<uint32_t*>(dst + ysz + (yo >> 1) + pc_off) = c0;
            }
            __syncwarp();   // everyone is done with the table before the next chunk's set-up rewrites it
        }
        t = __shfl_sync(0xffffffffu, t_next, 0) - ticket_base;
    }
}

// ------------------------------------------------------------------------------------------------
// k_res: residuals of the inter macroblocks, added in place
// ------------------------------------------------------------------------------------------------
constexpr int RES_WARPS = 4, RES_CHUNK = 16, RES_RUN = 4, RES_POOL = 16, RES_PSTRIDE = 72;   // pool stride in words (288 bytes)

struct ResSmem {
    int32_t pool[RES_POOL * RES_PSTRIDE];   // pooled coefficient blocks, 288 bytes apart
    uint32_t slotinfo[6 * RES_RUN];         // per pooled block in visiting order: byte offset of its pixels in the picture | pool slot << 24 | chroma << 31 (unused)
    uint8_t pad[32];
};
static_assert(sizeof(ResSmem) % 128 == 0 && (sizeof(ResSmem) * RES_WARPS + 1024) * RES_CTAS <= 233472, "ten 4-warp CTAs per SM");

template <int LOG2S>
__global__ void __launch_bounds__(RES_WARPS * 32, RES_CTAS)
k_res(const DevJob* __restrict__ jobs, uint32_t n_chunks, uint32_t cpp, uint32_t cpp_magic, int mbw, uint32_t mbw_magic, int H,
      uint32_t* __restrict__ ticket, uint32_t ticket_base) {
    constexpr int RUN = RES_RUN, S = 1 << LOG2S;
    __shared__ __align__(16) ResSmem s_all[RES_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    ResSmem& sm = s_all[warp];
    const uint32_t ysz = (uint32_t)S * (uint32_t)H;
    const int kc = lane & (RES_CHUNK - 1), k4 = lane & 3;
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
        uint32_t t_next = 0;
        if (lane == 0) asm volatile("atom.global.add.u32 %0, [%1], 1;" : "=r"(t_next) : "l"(ticket) : "memory");
        const uint32_t job = __umulhi(t, cpp_magic), chunk = t - job * cpp;
        const DevJob& J = jobs[job];
        const uint32_t n_mb = J.n_mb;
        if (J.n_intra != n_mb) {
            // ---- lane-parallel set-up: lane l (and l + 16) looks after macroblock mbc + l ----
            const uint32_t mbc = chunk * RES_CHUNK;
            const bool in = mbc + kc < n_mb;
            const uint32_t mbk = in ? mbc + kc : n_mb - 1;
            const uint4 d = __ldg(reinterpret_cast<const uint4*>(J.mbs + mbk));
            const uint32_t* const coefs = reinterpret_cast<const uint32_t*>(J.coefs);
            const uint32_t* const qtab = J.hdr->qtab;
            if (lane < 20) asm volatile("prefetch.global.L1 [%0];" :: "l"(qtab + lane * 4));   // the picture's 80 dequantisation words
            uint8_t* const dst = J.dst;
            const bool inter = in && !(d.x & 3u);
            const uint32_t n_coef = inter ? (d.x >> 9) & 511u : 0u;
            const uint32_t bm = n_coef ? (d.x >> 18) & 63u : 0u;
            const uint32_t m8 = (((d.x >> MB_M8_LO) & 15u) | ((d.x >> MB_M8_HI) & 3u) << 4) & bm;
            const int mby = (int)__umulhi(mbk, mbw_magic), mbx = (int)mbk - mby * mbw;
            const uint32_t yoff = (uint32_t)(((mby * 16) << LOG2S) + mbx * 16);
            if (!__any_sync(0xffffffffu, bm != 0u)) { t = __shfl_sync(0xffffffffu, t_next, 0) - ticket_base; continue; }

#pragma unroll 1
            for (int r = 0; r < RES_CHUNK / RUN; r++) {
                const int l0 = RUN * r;              // lanes l0 .. l0+3 hold this run's macroblocks
                const bool mine = (kc >> 2) == r && lane < RES_CHUNK;
                // coded blocks of the run, bit 6 * macroblock + block; which of them are transformed as one 8x8
                const uint32_t CM = __reduce_or_sync(0xffffffffu, mine ? bm << (6 * k4) : 0u);
                if (!CM) continue;
                const uint32_t M8 = __reduce_or_sync(0xffffffffu, mine ? m8 << (6 * k4) : 0u);
                // the run's coefficient records: one range of the picture's array
                const uint32_t jmin = __reduce_min_sync(0xffffffffu, mine && n_coef ? d.z : 0xffffffffu);
                const uint32_t jmax = __reduce_max_sync(0xffffffffu, mine && n_coef ? d.z + n_coef : 0u);
                const uint32_t ntot = jmax - jmin;
                const uint32_t* cf = coefs + jmin;
                uint32_t ca = 0, cb = 0;
                if ((uint32_t)lane < ntot) ca = __ldg(cf + lane);
                if ((uint32_t)lane + 32u < ntot) cb = __ldg(cf + 32 + lane);
                // where the pixels of (macroblock kk, block b) -- lane 6 * kk + b -- sit in the picture
                uint32_t pix = 0;
                {
                    const uint32_t kk = (uint32_t)lane / 6u, b = (uint32_t)lane - 6u * kk;
                    const uint32_t yo = __shfl_sync(0xffffffffu, yoff, l0 + (int)(kk & 3u));
                    pix = b < 4u ? yo + (((b >> 1) * 8u) << LOG2S) + (b & 1u) * 8u : ysz + (yo >> 1) + (b == 5u ? (uint32_t)(S >> 1) : 0u);
                }
                // The pool holds RES_POOL blocks; a run that codes more (at most 24) is worked off as two pairs of macroblocks.
                const uint32_t halves = (uint32_t)__popc(CM) > (uint32_t)RES_POOL ? 2u : 1u;
#pragma unroll 1
                for (uint32_t hv = 0; hv < halves; hv++) {
                    const uint32_t sel = halves == 1u ? 0xFFFFFFu : (hv ? 0xFFF000u : 0x000FFFu);
                    const uint32_t cm = CM & sel, c8 = M8 & sel, c4m = cm & ~c8;
                    const uint32_t ns = (uint32_t)__popc(cm), n8 = (uint32_t)__popc(c8);
                    if (!ns) continue;
                    __syncwarp();   // the passes before are done with the pool
                    {
                        int4* z = reinterpret_cast<int4*>(sm.pool);
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
                            sm.pool[p * RES_PSTRIDE + (e ^ ((e >> 3) & 4u))] = (int)(wq >> 8) * level;
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
                        const uint32_t lt = (1u << lane) - 1u;
                        const bool t8 = (c8 >> lane) & 1u, dc = (dcm >> lane) & 1u;
                        const uint32_t slot = t8 ? (uint32_t)__popc(c8 & lt) : n8 + (uint32_t)__popc(c4m & lt);
                        const uint32_t vp = dc ? nab + (uint32_t)__popc(dcm & lt) : t8 ? (uint32_t)__popc(c8t & lt) : na + (uint32_t)__popc(c4m & lt);
                        sm.slotinfo[vp] = pix | slot << 24;   // (a picture is at most 1.5 MB: 21 bits)
                    }
                    __syncwarp();

                    // ---- inverse transforms: eight lanes per pooled block (one row each), four blocks per pass ----
#pragma unroll 1
                    for (uint32_t base = 0; base < ns; base += 4u) {
                        const uint32_t idx = base + (uint32_t)g;
                        const bool has = idx < ns;
                        const uint32_t info = sm.slotinfo[has ? idx : base];
                        int32_t* const B = sm.pool + ((info >> 24) & 31u) * RES_PSTRIDE;
                        // this lane's eight pixels: row rr of the block (luma and chroma rows alike are S bytes apart); fetched now,
                        // needed after the transform
                        uint2* const tp = reinterpret_cast<uint2*>(dst + (info & 0xFFFFFFu) + ((uint32_t)rr << LOG2S));
                        uint2 px = make_uint2(0u, 0u);
                        if (has) px = *tp;
                        int32_t in[8], v[8];
                        if (base >= nab) {
                            // DC-only blocks: the residual is the constant (dc + 32) >> 6 (IDCT1Px8 MD:3710-3725), no transform
                            if (has) {
                                const int rs = (B[0] + 32) >> 6;
                                px.x = addsat4c(px.x, rs); px.y = addsat4c(px.y, rs);
                                *tp = px;
                            }
                            continue;
                        }
                        // One body for both transforms: passes are all-8x8 or all-4x4 except at the two boundaries of the visiting
                        // order, so the branches on is8 are uniform almost always (a DC-only block that shares a pass with 4x4 blocks
                        // simply takes the full 8x8 transform).
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
                            px.x = addsat4(px.x, v[0], v[1], v[2], v[3]);
                            px.y = addsat4(px.y, v[4], v[5], v[6], v[7]);
                            *tp = px;
                        }
                    }
                }
            }
        }
        t = __shfl_sync(0xffffffffu, t_next, 0) - ticket_base;
    }
}
