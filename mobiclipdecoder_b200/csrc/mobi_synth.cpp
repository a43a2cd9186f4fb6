// libmobisynth.so -- syntax-directed random Mobiclip stream writer.  See include/mobisynth.h.
//
// Grammar follows the reference decoder (MD = LibMobiclip/Codec/Mobiclip/MobiclipDecoder.cs):
//   frame header MD:113-143, 222-236; partition tree MD:469-1746; inter residual MD:1818-1833,
//   2909-2929; intra MBs MD:1759-1880, 2776-2896; residual VLC MD:3330-3432; Elias-gamma
//   MD:2970-3015 (writer mirror: BitWriter.cs:16-45).
// The model state a decoder derives from the bits (quantiser, intra-mode context grid, MV row
// cache and median predictor) is tracked here so that every emitted symbol decodes to the value
// that was intended.
#include "../../include/mobisynth.h"
#include "mobi_tables.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace {

struct Rng {
    uint64_t s;
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    bool chance(double p) { return uni() < p; }
    int range(int lo, int hi) { return lo + (int)(next() % (uint64_t)(hi - lo + 1)); }  // inclusive
    int geometric(double mean) {  // >= 0, E = mean
        if (mean <= 0) return 0;
        double p = 1.0 / (mean + 1.0), u = uni();
        int k = (int)std::floor(std::log(1.0 - u) / std::log(1.0 - p));
        return k < 0 ? 0 : k;
    }
};

struct BitW {  // 16-bit little-endian words, MSB first (BitWriter.cs:16-22, 57-64)
    std::vector<uint8_t> out;
    uint32_t acc = 0;
    int n = 0;
    void put(uint32_t v, int bits) {
        for (int i = bits - 1; i >= 0; i--) {
            acc = (acc << 1) | ((v >> i) & 1u);
            if (++n == 16) { out.push_back((uint8_t)acc); out.push_back((uint8_t)(acc >> 8)); acc = 0; n = 0; }
        }
    }
    void gamma(uint32_t x) {  // x >= 1: floor(log2 x) zeros, then x in binary
        int k = 31 - __builtin_clz(x);
        put(0, k);
        put(x, k + 1);
    }
    void uvar(uint32_t v) { gamma(v + 1); }
    void svar(int s) { gamma(s > 0 ? (uint32_t)(2 * s) : (uint32_t)(1 - 2 * s)); }
    void flush() { if (n) { acc <<= (16 - n); out.push_back((uint8_t)acc); out.push_back((uint8_t)(acc >> 8)); acc = 0; n = 0; } }
};

struct VlcBook {  // inverse of the 12-bit peek LUT
    uint8_t nbits[2][64][32];
    uint16_t code[2][64][32];
    void build(const mobi_vlc_run_t* runs, int nruns) {
        std::memset(nbits, 0, sizeof nbits);
        int idx = 0;
        for (int r = 0; r < nruns; r++) {
            uint16_t w = runs[r].word;
            int nb = w & 15, lvl = (w >> 4) & 31, run = (w >> 9) & 63, last = w >> 15;
            if (lvl != 0) { nbits[last][run][lvl] = (uint8_t)nb; code[last][run][lvl] = (uint16_t)(idx >> (13 - nb)); }
            idx += runs[r].span;
        }
    }
    bool has(int last, int run, int lvl) const { return run >= 0 && run < 64 && lvl > 0 && lvl < 32 && nbits[last][run][lvl]; }
};

struct Coef { int run, level, last; };

static inline void bfly8(const int32_t in[8], int32_t out[8]) {
    int32_t a0 = in[0] + in[4], a1 = in[0] - in[4], a2 = in[2] + (in[6] >> 1), a3 = (in[2] >> 1) - in[6];
    int32_t e0 = a0 + a2, e3 = a0 - a2, e1 = a1 + a3, e2 = a1 - a3;
    int32_t b0 = in[1] + in[7] - in[3] - (in[3] >> 1), b1 = in[7] - in[1] + in[5] + (in[5] >> 1);
    int32_t b2 = in[5] - (in[7] + (in[7] >> 1)) - in[3], b3 = in[3] + in[5] + in[1] + (in[1] >> 1);
    int32_t o1 = b2 + (b3 >> 2), o7 = b3 - (b2 >> 2), o3 = b0 + (b1 >> 2), o5 = (b0 >> 2) - b1;
    out[0] = e0 + o7; out[7] = e0 - o7; out[1] = e1 + o5; out[6] = e1 - o5;
    out[2] = e2 + o3; out[5] = e2 - o3; out[3] = e3 + o1; out[4] = e3 - o1;
}
static inline void bfly4(const int32_t in[4], int32_t out[4]) {
    int32_t s = in[0] + in[2], d = in[0] - in[2], p = (in[1] >> 1) - in[3], q = in[1] + (in[3] >> 1);
    out[0] = s + q; out[3] = s - q; out[1] = d + p; out[2] = d - p;
}
// largest |residual| the block would add to any pixel (transform of MD:3435 / MD:3728)
static int max_delta(const int32_t* c, int N) {
    int32_t t[64], in[8], out[8], cc[64];
    std::memcpy(cc, c, sizeof(int32_t) * N * N);
    cc[0] += 32;
    for (int i = 0; i < N; i++) {
        for (int k = 0; k < N; k++) in[k] = cc[N * i + k];
        if (N == 8) bfly8(in, out); else bfly4(in, out);
        for (int k = 0; k < N; k++) t[N * k + i] = out[k];
    }
    int m = 0;
    for (int r = 0; r < N; r++) {
        for (int k = 0; k < N; k++) in[k] = t[N * r + k];
        if (N == 8) bfly8(in, out); else bfly4(in, out);
        for (int k = 0; k < N; k++) m = std::max(m, std::abs(out[k] >> 6));
    }
    return m;
}

}  // namespace

struct mobi_synth {
    mobi_synth_params P;
    Rng rng;
    int S, mbw, mbh;
    int frame_idx = 0, decoded = 0;
    // decoder-visible model state
    uint32_t quant = 0;
    uint32_t scale8[64], scale4[16];
    uint8_t ctx[40];
    int vlcsel = 0;
    std::vector<int> mvc;  // MV row cache, 2 ints per entry (MD:145-154)
    int mvpx = 0, mvpy = 0;
    VlcBook book[2];
    int inv_cbp6_intra[64], inv_cbp6_inter[64], inv_cbp4_inter[16];
    std::vector<int> idx_cbp4_intra[16];
    BitW bw;
    mobi_synth_stats st;

    void setup_quant(uint32_t q) {  // MD:3884-3925
        if (P.version == 2) { if (q < 12) q = 12; if (q > 52) q = 52; }
        quant = q;
        int row = (int)(q % 6), sh = (int)(q / 6);
        for (int i = 0; i < 16; i++) scale4[i] = (((uint32_t)MOBI_SCALE4[row * 16 + i] << (sh + 8)) | MOBI_SCAN4[i]) >> 8;
        for (int i = 0; i < 64; i++) scale8[i] = (((uint32_t)MOBI_SCALE8[row * 64 + i] << (sh + 6)) | MOBI_SCAN8[i]) >> 8;
        ctx[1] = ctx[2] = ctx[3] = ctx[4] = 9; ctx[8] = ctx[0x10] = ctx[0x18] = ctx[0x20] = 9;
    }

    // ---- residual ---------------------------------------------------------------------------
    void put_coef(const Coef& c) {
        const VlcBook& B = book[vlcsel];
        const uint8_t* E = vlcsel ? MOBI_VLC1_ESC : MOBI_VLC0_ESC;
        int a = std::abs(c.level), sign = c.level < 0;
        bool direct = B.has(c.last, c.run, a);
        int lo = a - E[(c.last << 6) | c.run];                       // level-offset escape (MD:3347-3366)
        bool esc_l = B.has(c.last, c.run, lo);
        int ro = (a < 32) ? c.run - E[0x80 + a + (c.last << 6)] : -1;  // run-offset escape (MD:3371-3390)
        bool esc_r = a < 32 && B.has(c.last, ro, a);
        int form;  // 0 direct, 1 level-offset, 2 run-offset, 3 raw
        if (direct && !rng.chance(P.p_escape)) form = 0;
        else {
            int opts[4], n = 0;
            if (direct && false) opts[n++] = 0;
            if (esc_l) opts[n++] = 1;
            if (esc_r) opts[n++] = 2;
            opts[n++] = 3;
            form = opts[rng.range(0, n - 1)];
            if (!direct && esc_l && !rng.chance(0.3)) form = 1;  // what a real encoder would prefer
        }
        switch (form) {
        case 0: bw.put(B.code[c.last][c.run][a], B.nbits[c.last][c.run][a] - 1); bw.put(sign, 1); break;
        case 1: bw.put(3, 7); bw.put(0, 1); bw.put(B.code[c.last][c.run][lo], B.nbits[c.last][c.run][lo] - 1); bw.put(sign, 1); break;
        case 2: bw.put(3, 7); bw.put(2, 2); bw.put(B.code[c.last][ro][a], B.nbits[c.last][ro][a] - 1); bw.put(sign, 1); break;
        default: bw.put(3, 7); bw.put(3, 2); bw.put(c.last, 1); bw.put(c.run, 6); bw.put((uint32_t)c.level & 0xFFF, 12); break;
        }
        st.n_coefs++;
    }
    void gen_block(int N) {  // one coded transform block: coefficient list within the clip-safe bound
        const uint32_t* scale = N == 8 ? scale8 : scale4;
        const uint8_t* scan = N == 8 ? MOBI_SCAN8 : MOBI_SCAN4;
        int maxpos = N * N - 1;
        std::vector<Coef> cs;
        int want = 1 + rng.geometric(std::max(0.0f, P.mean_coefs - 1.0f));
        want = std::min(want, N * N);
        int pos = 0;
        for (int i = 0; i < want; i++) {
            int room = maxpos - pos - (want - 1 - i);
            if (room < 0) break;
            int run = std::min(rng.geometric(N == 8 ? 1.5 : 0.7), room);
            int mag = 1 + rng.geometric(2.0);
            if (rng.chance(0.02)) mag = rng.range(32, 300);
            mag = std::min(mag, 2047);
            cs.push_back(Coef{run, rng.chance(0.5) ? -mag : mag, 0});
            pos += run + 1;
        }
        if (cs.empty()) cs.push_back(Coef{0, 1, 0});
        for (int attempt = 0;; attempt++) {
            int32_t c[64] = {0};
            int p = 0;
            for (auto& k : cs) { p += k.run; c[scan[p]] = (int32_t)(scale[p] * (uint32_t)k.level); p++; }
            if (max_delta(c, N) <= 64) break;
            if (attempt >= 12) { cs.assign(1, Coef{0, rng.chance(0.5) ? -1 : 1, 0}); break; }
            for (auto& k : cs) { int h = k.level / 2; k.level = h ? h : (k.level < 0 ? -1 : 1); }
        }
        cs.back().last = 1;
        for (auto& k : cs) put_coef(k);
        if (N == 8) st.n_blk8++; else st.n_blk4++;
    }

    // ---- intra ------------------------------------------------------------------------------
    // May the predictor `m` (0..8) be used for a block whose top-left pixel is (x,y) of its plane?
    // Row 0 has nothing above it (negative index -> exception, MD:1893 etc.); pixel (0,0) has nothing
    // to its left either.
    static bool mode_ok(int m, int x, int y) {
        if (y > 0) return true;
        if (m == 3) return true;
        if (m == 1 || m == 4) return x > 0;
        return false;
    }
    int pick_mode(int x, int y, bool allow8) {
        for (;;) { int m = rng.range(0, allow8 ? 8 : 7); if (mode_ok(m, x, y)) return m; }
    }
    void put_mode_ctx(int ci, int m, bool all4) {  // MD:1840-1856 / 2841-2854
        int pred = std::min(ctx[ci - 8], ctx[ci - 1]);
        if (pred == 9) pred = 3;
        if (m == pred) bw.put(1, 1);
        else bw.put((uint32_t)(m > pred ? m - 1 : m), 4);
        ctx[ci] = (uint8_t)m;
        if (all4) ctx[ci + 1] = ctx[ci + 8] = ctx[ci + 9] = (uint8_t)m;
    }
    void put_cbp4_intra(int v) { const std::vector<int>& ix = idx_cbp4_intra[v]; bw.uvar((uint32_t)ix[rng.range(0, (int)ix.size() - 1)]); }
    void plane_delta() { bw.svar(rng.range(-12, 12)); }

    // coded block with an MB-level mode (sub_116508 MD:2869); m is 0..7 or 9
    void intra_coded_fixed(int m) {
        (void)m;
        if (rng.chance(P.p_blk8)) { bw.put(1, 1); gen_block(8); }
        else {
            int cbp4 = rng.range(0, 15);
            put_cbp4_intra(cbp4);
            for (int k = 0; k < 4; k++) if ((cbp4 >> k) & 1) gen_block(4);
        }
    }
    void intra_chroma(int cbp6, int mbx, int mby) {  // loc_116290 MD:1864
        int cx = mbx * 8, cy = mby * 8, m;
        for (;;) { m = rng.range(0, 7); if (m == 2 ? cy > 0 : mode_ok(m, cx, cy)) break; }
        bw.put((uint32_t)m, 3);
        st.mode_hist[m]++;
        if (m == 2) { plane_delta(); plane_delta(); m = 9; }
        if (cbp6 & 16) intra_coded_fixed(m);
        if (cbp6 & 32) intra_coded_fixed(m);
    }
    void intra_mb(bool sub, int mbx, int mby) {
        int cbp6 = 0;
        for (int b = 0; b < 6; b++) if (rng.chance(P.p_cbp)) cbp6 |= 1 << b;
        bw.uvar((uint32_t)inv_cbp6_intra[cbp6]);
        int px = mbx * 16, py = mby * 16;
        static const int bx[4] = {0, 8, 0, 8}, by[4] = {0, 0, 8, 8}, ci[4] = {9, 0xB, 0x19, 0x1B};
        if (!sub) {  // DecIntraFullBlockPMode MD:1759
            int m;
            for (;;) { m = rng.range(0, 7); if (m == 2 ? py > 0 : (mode_ok(m, px, py) && mode_ok(m, px + 8, py))) break; }
            bw.put((uint32_t)m, 3);
            st.mode_hist[m]++;
            if (m == 2) { plane_delta(); m = 9; }
            for (int b = 0; b < 4; b++) if ((cbp6 >> b) & 1) intra_coded_fixed(m);
        } else {  // DecIntraSubBlockPMode MD:1789
            static const int sx[4] = {0, 4, 0, 4}, sy[4] = {0, 0, 4, 4}, sc[4] = {0, 1, 8, 9};
            for (int b = 0; b < 4; b++) {
                int x = px + bx[b], y = py + by[b];
                bool coded = (cbp6 >> b) & 1;
                bool whole = !coded || rng.chance(P.p_blk8);
                if (coded) bw.put(whole ? 1 : 0, whole ? 1 : 0);  // the split path's leading 0 belongs to its uvar
                if (whole) {
                    int m = pick_mode(x, y, true);
                    put_mode_ctx(ci[b], m, true);
                    st.mode_hist[m]++;
                    if (m == 2) plane_delta();
                    if (coded) gen_block(8);
                } else {
                    int cbp4 = rng.range(0, 15);
                    put_cbp4_intra(cbp4);
                    for (int k = 0; k < 4; k++) {
                        int m = pick_mode(x + sx[k], y + sy[k], true);
                        put_mode_ctx(ci[b] + sc[k], m, false);
                        st.mode_hist[10 + m]++;
                        if (m == 2) plane_delta();
                        if ((cbp4 >> k) & 1) gen_block(4);
                    }
                }
            }
        }
        intra_chroma(cbp6, mbx, mby);
        st.n_intra_mb++;
    }

    // ---- inter ------------------------------------------------------------------------------
    bool mv_legal(int x, int y, int w, int h, int dx, int dy, bool oob) const {
        int W = (int)P.width, H = (int)P.height;
        long long off = (long long)y * S + x;
        long long first = off + (long long)(dy >> 1) * S + (dx >> 1);
        long long last = off + (long long)((dy >> 1) + h - 1 + (dy & 1)) * S + (dx >> 1) + w - 1 + (dx & 1);
        if (first < 0 || last >= (long long)S * H) return false;
        int cdx = dx >> 1, cdy = dy >> 1, cw = w >> 1, ch = h >> 1;
        long long coff = off / 2;
        long long cfirst = coff + (long long)(cdy >> 1) * S + (cdx >> 1);
        long long clast = coff + S / 2 + (long long)((cdy >> 1) + ch - 1 + (cdy & 1)) * S + (cdx >> 1) + cw - 1 + (cdx & 1);
        if (cfirst < 0 || clast >= (long long)S * H / 2) return false;
        if (oob) return true;
        int x0 = x + (dx >> 1), x1 = x0 + w - 1 + (dx & 1), y0 = y + (dy >> 1), y1 = y0 + h - 1 + (dy & 1);
        if (x0 < 0 || y0 < 0 || x1 >= W || y1 >= H) return false;
        int cx0 = x / 2 + (cdx >> 1), cx1 = cx0 + cw - 1 + (cdx & 1), cy0 = y / 2 + (cdy >> 1), cy1 = cy0 + ch - 1 + (cdy & 1);
        return cx0 >= 0 && cy0 >= 0 && cx1 < W / 2 && cy1 < H / 2;
    }
    void put_part_sym(int lw, int lh, int sym) {
        const mobi_part_code_t& pc = MOBI_PART_CODE[P.version == 2 ? 0 : 1][lw][lh];
        int len = pc.len[sym], first = 0;
        while (pc.sym[first] != sym) first++;
        bw.put((uint32_t)(first >> (pc.peek - len)), len);
    }
    void leaf(int lw, int lh, int x, int y, int cache_slot) {
        int w = 2 << lw, h = 2 << lh;
        int nref = std::min(5, decoded);
        bool oob = rng.chance(P.p_oob_mv);
        int dx, dy, ref = 1, sym;
        if (rng.chance(P.p_zero_mv) && mv_legal(x, y, w, h, mvpx, mvpy, true)) { sym = 0; dx = mvpx; dy = mvpy; }
        else {
            ref = (nref > 1 && !rng.chance(P.p_ref1)) ? rng.range(1, nref) : 1;
            int tries = 0;
            for (;; tries++) {
                int r = oob ? 48 : P.mv_range;
                dx = mvpx + rng.range(-r, r); dy = mvpy + rng.range(-r, r);
                if (tries >= 24) { dx = 0; dy = 0; }
                dx = std::max(-48, std::min(48, dx)); dy = std::max(-48, std::min(48, dy));
                if (mv_legal(x, y, w, h, dx, dy, oob || tries >= 24)) break;
            }
            sym = ref;
        }
        put_part_sym(lw, lh, sym);
        if (sym) { bw.svar(dx - mvpx); bw.svar(dy - mvpy); }
        mvc[cache_slot] = dx; mvc[cache_slot + 1] = dy;
        st.n_leaves++; st.shape_hist[lw * 4 + lh]++; st.phase_hist[(dx & 1) | ((dy & 1) << 1)]++; st.ref_hist[ref]++;
    }
    void part(int lw, int lh, int x, int y, int cache_slot) {
        const mobi_part_code_t& pc = MOBI_PART_CODE[P.version == 2 ? 0 : 1][lw][lh];
        bool can8 = lh > 0 && pc.len[8] > 0, can9 = lw > 0 && pc.len[9] > 0;
        if ((can8 || can9) && rng.chance(P.p_split)) {
            bool tb = can8 && (!can9 || rng.chance(0.5));
            put_part_sym(lw, lh, tb ? 8 : 9);
            int w = 2 << lw, h = 2 << lh;
            if (tb) { part(lw, lh - 1, x, y, cache_slot); part(lw, lh - 1, x, y + h / 2, cache_slot); }
            else { part(lw - 1, lh, x, y, cache_slot); part(lw - 1, lh, x + w / 2, y, cache_slot); }
        } else leaf(lw, lh, x, y, cache_slot);
    }
    void inter_residual() {  // loc_1161A0 MD:1818
        int cbp6 = 0;
        for (int b = 0; b < 6; b++) if (rng.chance(P.p_cbp)) cbp6 |= 1 << b;
        bw.uvar((uint32_t)inv_cbp6_inter[cbp6]);
        for (int b = 0; b < 6; b++) if ((cbp6 >> b) & 1) {
            if (rng.chance(P.p_blk8)) { bw.put(1, 1); gen_block(8); }
            else {
                int cbp4 = rng.range(1, 15);
                bw.uvar((uint32_t)inv_cbp4_inter[cbp4]);
                for (int k = 0; k < 4; k++) if ((cbp4 >> k) & 1) gen_block(4);
            }
        }
    }
    static int med3(int a, int b, int c) { return std::max(std::min(a, b), std::min(std::max(a, b), c)); }

    int next_frame(uint8_t* out, int cap, int* is_key) {
        bool key = frame_idx == 0 || (P.gop > 0 && (frame_idx + P.gop_phase) % P.gop == 0);
        std::memset(&st, 0, sizeof st);
        bw = BitW();
        if (key) {
            bw.put(1, 1);
            bw.put((uint32_t)rng.range(0, 1), 1);  // yuv_format: parsed, never used (MD:224)
            vlcsel = rng.range(0, 1);
            bw.put((uint32_t)vlcsel, 1);
            uint32_t q = frame_idx == 0 ? (uint32_t)P.quant : quant;
            if (frame_idx != 0 && rng.chance(P.p_dquant)) q = (uint32_t)std::max(12, std::min(46, (int)q + rng.range(-3, 3)));
            bw.put(q, 6);
            if (quant != q) setup_quant(q);
            for (int my = 0; my < mbh; my++) for (int mx = 0; mx < mbw; mx++) {
                bool sub = rng.chance(P.p_sub_mb);
                bw.put(sub, 1);
                intra_mb(sub, mx, my);
                st.n_mb++;
            }
        } else {
            bw.put(0, 1);
            int dq = 0;
            if (rng.chance(P.p_dquant)) { dq = rng.range(-2, 2); if ((int)quant + dq < 12 || (int)quant + dq > 46) dq = 0; }
            bw.svar(dq);
            if (P.version == 2) { if (quant == 0) setup_quant(0); else if (dq) setup_quant((uint32_t)((int)quant + dq)); }
            else if (dq) setup_quant((uint32_t)((int)quant + dq));
            vlcsel = 0;
            std::fill(mvc.begin(), mvc.end(), 0);
            for (int my = 0; my < mbh; my++) for (int mx = 0; mx < mbw; mx++) {
                const int* e = &mvc[2 * mx];
                mvpx = med3(e[0], e[2], e[4]); mvpy = med3(e[1], e[3], e[5]);
                int slot = 2 * (mx + 1);
                mvc[slot] = mvc[slot + 1] = 0;
                if (!P.inter_only && rng.chance(P.p_intra_mb)) {
                    bool sub = rng.chance(P.p_sub_mb);
                    put_part_sym(3, 3, sub ? 7 : 6);
                    intra_mb(sub, mx, my);
                } else {
                    part(3, 3, mx * 16, my * 16, slot);
                    inter_residual();
                }
                st.n_mb++;
            }
        }
        bw.flush();
        frame_idx++; decoded++;
        if (is_key) *is_key = key;
        if ((int)bw.out.size() > cap) return -(int)bw.out.size();
        std::memcpy(out, bw.out.data(), bw.out.size());
        return (int)bw.out.size();
    }
};

extern "C" {

void mobi_synth_default_params(mobi_synth_params* p, uint32_t width, uint32_t height, int version, uint64_t seed) {
    std::memset(p, 0, sizeof *p);
    p->width = width; p->height = height; p->version = version; p->seed = seed;
    p->gop = 90;  // reference encoder inserts an I-frame at least every 90 P-frames (MobiEncoder.cs:123)
    p->quant = 24;
    p->p_dquant = 0.05f; p->p_split = 0.25f; p->p_intra_mb = 0.05f; p->p_sub_mb = 0.5f;
    p->p_cbp = 0.3f; p->p_blk8 = 0.6f; p->mean_coefs = 4.0f; p->p_escape = 0.03f;
    p->mv_range = 16; p->p_ref1 = 0.8f; p->p_zero_mv = 0.35f; p->p_oob_mv = 0.03f; p->inter_only = 0;
}

mobi_synth_t* mobi_synth_create(const mobi_synth_params* p) {
    if (!p || p->width == 0 || p->height == 0 || (p->width & 15) || (p->height & 15) || p->width > 1024) return nullptr;
    if (p->version != 1 && p->version != 2) return nullptr;
    mobi_synth* s = new mobi_synth();
    s->P = *p;
    s->P.quant = std::max(12, std::min(46, p->quant));
    s->P.mv_range = std::max(0, std::min(32, p->mv_range));
    if (s->P.gop_phase < 0) s->P.gop_phase = 0;
    s->rng.s = p->seed * 0x9E3779B97F4A7C15ull + 0x1234567ull;
    s->S = p->width <= 256 ? 256 : p->width <= 512 ? 512 : 1024;
    s->mbw = (int)p->width / 16; s->mbh = (int)p->height / 16;
    std::memset(s->ctx, 0, sizeof s->ctx);
    std::memset(s->scale8, 0, sizeof s->scale8); std::memset(s->scale4, 0, sizeof s->scale4);
    s->mvc.assign(2 * (s->mbw + 3), 0);
    s->book[0].build(MOBI_VLC0_RUNS, (int)(sizeof(MOBI_VLC0_RUNS) / sizeof(MOBI_VLC0_RUNS[0])));
    s->book[1].build(MOBI_VLC1_RUNS, (int)(sizeof(MOBI_VLC1_RUNS) / sizeof(MOBI_VLC1_RUNS[0])));
    for (int i = 0; i < 64; i++) { s->inv_cbp6_intra[MOBI_CBP6_INTRA[i]] = i; s->inv_cbp6_inter[MOBI_CBP6_INTER[i]] = i; }
    for (int i = 0; i < 16; i++) s->inv_cbp4_inter[MOBI_CBP4_INTER[i]] = i;
    for (int i = 1; i < 20; i++) s->idx_cbp4_intra[MOBI_CBP4_INTRA[i]].push_back(i);  // index 0 is the bare '1' = "8x8" flag
    std::memset(&s->st, 0, sizeof s->st);
    return s;
}

void mobi_synth_destroy(mobi_synth_t* s) { delete s; }

int mobi_synth_next_frame(mobi_synth_t* s, uint8_t* out, int cap, int* is_key) { return s->next_frame(out, cap, is_key); }

void mobi_synth_last_stats(const mobi_synth_t* s, mobi_synth_stats* st) { *st = s->st; }

}  // extern "C"
