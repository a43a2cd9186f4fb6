// Host entropy parser.  "MD:n" = LibMobiclip/Codec/Mobiclip/MobiclipDecoder.cs line n of the reference.
//
// Bit-exactness notes (things a tidy re-implementation would get subtly wrong):
//  * the reader keeps a 32-bit window whose top 16 bits are always valid and refills ONE 16-bit
//    little-endian word when its counter goes negative (MD:2988-2996); several call sites consume
//    bits without the refill check (MD:2780-2781, 2873-2874, 2913-2914) -- mirrored 1:1 so that
//    Offset after the call matches the reference even on odd streams;
//  * Offset is an output of DecodeFrame (callers locate audio at Offset-2, MobiConverter/Program.cs:250);
//  * the intra-mode context grid is only re-bordered when the quantiser tables are rebuilt (MD:3913-3924);
//  * every condition under which the C# code would throw (-> null Bitmap, MD:325) is detected here,
//    before anything is sent to the GPU, and reported as an error status.
#include "mobi_parse.h"
#include "mobi_tables.h"
#include <cstring>
#include <mutex>

namespace mobi {

namespace {

struct ParseError { int code; const char* msg; };
[[noreturn]] inline void fail(int code, const char* msg) { throw ParseError{code, msg}; }

uint16_t g_vlc[2][4096];
std::once_flag g_once;
void build_tables() {
    const mobi_vlc_run_t* runs[2] = {MOBI_VLC0_RUNS, MOBI_VLC1_RUNS};
    const int n[2] = {(int)(sizeof(MOBI_VLC0_RUNS) / sizeof(mobi_vlc_run_t)), (int)(sizeof(MOBI_VLC1_RUNS) / sizeof(mobi_vlc_run_t))};
    for (int t = 0; t < 2; t++) {
        int k = 0;
        for (int r = 0; r < n[t]; r++) for (int j = 0; j < runs[t][r].span; j++) g_vlc[t][k++] = runs[t][r].word;
    }
}

}  // namespace

// ---- bit window (MD:109-112, 2970-3015) --------------------------------------------------------
struct Parser::Bits {
    const uint8_t* d;
    int len, off;
    uint32_t win;
    int nb;
    uint32_t u16(int at) const {
        if (at < 0 || at + 1 >= len) fail(MOBI_ERR_BITSTREAM, "read past end of frame data");
        return (uint32_t)d[at] | (uint32_t)d[at + 1] << 8;
    }
    void fill() {
        if (off >= len) return;
        uint32_t w = u16(off);
        off += 2;
        nb += 16;
        win |= w << ((16 - nb) & 31);
    }
    void drop(int n) { win <<= (n & 31); nb -= n; }
    // FillBits when the counter is negative (MD:2988-2996).  Whether it is negative is data-dependent and poorly
    // predicted, so while a whole word is available the refill is done arithmetically; the last word takes fill().
    void chk() {
        if (__builtin_expect(off + 1 < len, 1)) {
            const uint32_t m = (uint32_t)(nb >> 31);
            const uint32_t w = (uint32_t)d[off] | (uint32_t)d[off + 1] << 8;
            nb += (int)(16u & m); off += (int)(2u & m);
            win |= (w << ((16 - nb) & 31)) & m;
        } else if (nb < 0) fill();
    }
    uint32_t take(int n) { uint32_t v = win >> (32 - n); drop(n); chk(); return v; }
    uint32_t gamma() {
        int z = win ? __builtin_clz(win) : 32;
        win <<= (z & 31);
        win += win;
        int sh = 32 - z;
        uint32_t v = (sh == 32) ? 0 : win >> (sh & 31);
        v += (uint32_t)(1u << (z & 31));
        win <<= (z & 31);
        nb -= z << 1;
        --nb; chk();
        return v;
    }
    uint32_t uvar() { return gamma() - 1; }
    int svar() { const int v = (int)gamma(), n = -(v & 1); return ((v ^ n) - n + (v & 1)) >> 1; }   // odd v: (1 - v) >> 1, even v: v >> 1, branch-free
};

struct FrameParse {
    Parser& P;
    Parser::Bits b;
    ParsedFrame& out;
    uint32_t vlcsel = 0;
    int32_t mvpx = 0, mvpy = 0;
    int S, W, H;
    uint32_t max_ref = 0;
    uint32_t cur_deps = 0;
    int cur_mbx = 0, cur_mby = 0;
    uint32_t cur_m8 = 0;   // coded blocks of the macroblock in hand that are transformed as one 8x8
    uint8_t own_tag = 0;   // (macroblock index & 3) << 3: rides in bits 3-4 of mobi_coef.blk (the inter kernel pools the coefficients of four macroblocks)

    int log2S;
    FrameParse(Parser& p, ParsedFrame& o) : P(p), out(o), S(p.S_), W((int)p.W_), H((int)p.H_), log2S(p.S_ == 256 ? 8 : p.S_ == 512 ? 9 : 10) {}

    static uint32_t tab(const uint8_t* t, uint32_t n, uint32_t i) {
        if (i >= n) fail(MOBI_ERR_BITSTREAM, "code index outside table");
        return t[i];
    }

    // SetupQuantizationTables MD:3884-3925
    void setup_quant(uint32_t q) {
        if (P.ver_ == MOBI_MOFLEX3DS) { if (q < 12) q = 12; if (q > 52) q = 52; }
        P.st_.quant = q;
        if (q >= MOBI_QTAB_MAXQ) fail(MOBI_ERR_BITSTREAM, "quantiser outside table");
        int sh = (int)(q / 6) + 8, row = (int)(q % 6);
        for (int i = 0; i < 16; i++) P.st_.qtab[64 + i] = (uint32_t)MOBI_SCAN4[i] | (uint32_t)MOBI_SCALE4[row * 16 + i] << sh;
        sh -= 2;
        for (int i = 0; i < 64; i++) P.st_.qtab[i] = (uint32_t)MOBI_SCAN8[i] | (uint32_t)MOBI_SCALE8[row * 64 + i] << sh;
        bool clean = true;  // q < 12 on ModsDS lets the scale leak into the matrix-index byte (MD:3909-3911 vs MD:3426)
        for (int i = 0; i < 80; i++) if ((P.st_.qtab[i] & 0xFF) >= (i < 64 ? 64u : 16u)) clean = false;
        P.st_.qtab_clean = clean;
        uint8_t* c = P.st_.ctx;
        c[1] = c[2] = c[3] = c[4] = 9; c[8] = c[0x10] = c[0x18] = c[0x20] = 9;
    }

    // ReadDCTMatrix MD:3330-3432: emits quantised levels; scaling happens on the device.
    // n = 64 or 16; tag = blk | sub<<... as stored in mobi_coef.
    // The bit window lives in locals for the duration of the block (the compiler cannot keep members in registers
    // across the stores to the output array); the refill rule is the reference's, one 16-bit word per check.
    void coefs(int n, uint8_t blk, uint8_t sub, uint32_t& blkmask_any) {
        const uint16_t* A = g_vlc[vlcsel == 1];
        const uint8_t* B = vlcsel == 1 ? MOBI_VLC1_ESC : MOBI_VLC0_ESC;
        const uint32_t* qt = n == 64 ? P.st_.qtab : P.st_.qtab + 64;
        const bool check_qt = !P.st_.qtab_clean;
        const uint8_t* const data = b.d;
        const int len = b.len;
        uint32_t win = b.win;
        int nb = b.nb, off = b.off;
        mobi_coef* dst = &out.coefs.v[out.coefs.n];
        const uint8_t tag = (uint8_t)(blk | (n == 64 ? 0x80 : 0) | own_tag);
        // Refill (MD:2988-2996) without a data-dependent branch: whether the counter went negative is close to a coin flip
        // per coefficient, so the common case (a whole word is still available) does it arithmetically; the tail of the
        // buffer takes the literal path.
#define MOBI_CHK() do { \
            if (__builtin_expect(off + 1 < len, 1)) { \
                const uint32_t m_ = (uint32_t)(nb >> 31);                                   /* all ones iff nb < 0 */ \
                const uint32_t w_ = (uint32_t)data[off] | (uint32_t)data[off + 1] << 8; \
                nb += (int)(16u & m_); off += (int)(2u & m_); \
                win |= (w_ << ((16 - nb) & 31)) & m_; \
            } else if (nb < 0 && off < len) fail(MOBI_ERR_BITSTREAM, "read past end of frame data"); \
        } while (0)
        uint32_t pos = 0;
        for (;;) {
            int run, level, nbits;
            uint32_t e, last;
            if (__builtin_expect((win >> 25) == 3, 0)) {
                win <<= 7;
                uint32_t c = win >> 31; win <<= 1;
                if (!c) {
                    nb -= 8; MOBI_CHK();
                    e = A[win >> 20];
                    int add = B[e >> 9];
                    nbits = e & 15; e >>= 4; level = (int)(e & 31) + add; e >>= 5;
                    win <<= ((nbits - 1) & 31);
                    if (win >> 31) level = -level;
                    win <<= 1; nb -= nbits; MOBI_CHK();
                    run = e & 63; last = e >> 6;
                } else {
                    c = win >> 31; win <<= 1;
                    if (!c) {
                        nb -= 9; MOBI_CHK();
                        e = A[win >> 20];
                        nbits = e & 15; e >>= 4; level = e & 31; e >>= 5;
                        uint32_t r = e & 63; e >>= 6;
                        int add = B[0x80 + level + (e << 6)];
                        win <<= ((nbits - 1) & 31);
                        if (win >> 31) level = -level;
                        win <<= 1; nb -= nbits; MOBI_CHK();
                        run = (int)r + add; last = e;
                    } else {
                        nb -= 9; MOBI_CHK();
                        last = win >> 31; win <<= 1;
                        run = win >> 26; win <<= 6;
                        nb -= 7; MOBI_CHK();
                        level = (int32_t)win >> 20; win <<= 12;
                        nb -= 12; MOBI_CHK();
                    }
                }
            } else {
                e = A[win >> 20];
                nbits = e & 15; e >>= 4; level = e & 31; e >>= 5;
                win <<= ((nbits - 1) & 31);
                const int32_t neg = -(int32_t)(win >> 31);          // branch-free sign: 0 or -1
                level = (level ^ neg) - neg;
                win <<= 1; nb -= nbits; MOBI_CHK();
                run = e & 63; last = e >> 6;
            }
            pos += (uint32_t)run;
            // The reference indexes Internal[] with no check: a run past the block walks into the next table.
            if (__builtin_expect(pos >= (uint32_t)n, 0)) fail(MOBI_ERR_BITSTREAM, "coefficient run past end of block");
            if (check_qt && (qt[pos] & 0xFF) >= 64u) fail(MOBI_ERR_BITSTREAM, "quantiser < 12 corrupts the scan table (MD:3909-3911)");
            {   // one 32-bit store: level (s16) | pos << 16 | blk << 24 (little-endian layout of mobi_coef)
                const uint32_t rec = ((uint32_t)level & 0xFFFFu) | (pos | (uint32_t)sub << 6) << 16 | ((uint32_t)tag | (last & 1u) << 6) << 24;
                std::memcpy(dst, &rec, 4);
                dst++;
            }
            pos++;
            if (last & 1) break;
        }
#undef MOBI_CHK
        out.coefs.n = (size_t)(dst - out.coefs.v.data());
        b.win = win; b.nb = nb; b.off = off;
        blkmask_any |= 1u << (blk & 7);
        if (n == 64) cur_m8 |= 1u << (blk & 7);
    }

    // ---- intra -----------------------------------------------------------------------------
    // Lowest / highest plane index a predictor touches outside its block (MD:1883-2774, 3017-3327);
    // anything below 0 would be an IndexOutOfRangeException in the reference.
    void check_intra_reads(uint32_t mode, int off) {
        int lo = 0;
        switch (mode) {
        case 0: case 8: case 2: case 10: case 18: case 12: case 20: lo = off - S; break;
        case 1: case 4: case 11: case 14: lo = off - 1; break;
        case 5: case 6: case 7: case 15: case 16: case 17: lo = off - S - 1; break;
        default: return;  // 3/13 test availability themselves; 9/19 read nothing
        }
        if (mode == 2 || mode == 12 || mode == 20) { if (off - 1 < lo) lo = off - 1; }
        if (lo < 0) fail(MOBI_ERR_RANGE, "intra predictor reads above/left of the picture");
    }
    int plane_off(int plane, int mboff, int x4, int y4) const {
        if (plane == 0) return mboff + y4 * 4 * S + x4 * 4;
        return mboff / 2 + (plane == 2 ? S / 2 : 0) + y4 * 4 * S + x4 * 4;
    }
    // Which neighbouring macroblocks does this predictor read?  bit 0 left (m-1), 1 top-left (m-mbw-1), 2 top (m-mbw),
    // 3 top-right (m-mbw+1), in raster numbering -- which is also what the reference's flat addressing resolves to at
    // the picture edges: the "left" of column 0 is the last byte of the previous row, i.e. the previous MB row's last
    // macroblock when Width == Stride and zero padding otherwise (SURVEY.md 8a hazard 2).  Footprints: SURVEY.md App. C.
    void note_deps(uint32_t mode, int plane, int x4, int y4, int off) {
        int N = 8, pm = (int)mode;
        if (mode == 20) { N = 16; pm = 2; } else if (mode >= 10) { N = 4; pm = (int)mode - 10; }
        if (pm == 9) return;
        bool top = false, left = false, tl = false;
        int ext = 0;  // pixels read beyond the block's right edge in the row above
        switch (pm) {
        case 0: top = true; break;
        case 1: case 4: left = true; break;
        case 2: top = left = true; break;
        case 3: {
            const int po = plane == 2 ? off - S / 2 : off;
            left = (po & (S - 1)) != 0; top = off >= S;   // MD:1923-1924 (Stride is a power of two, MD:50-52)
            break; }
        case 5: case 6: case 7: top = left = tl = true; break;
        case 8: top = true; ext = N == 8 ? 5 : 4; break;   // T0..12 / two top words (MD:2371-2466, 2737-2746)
        }
        const int cells = plane == 0 ? 4 : 2;          // macroblock width in 4-pixel cells on this plane
        const int xe = x4 + N / 4;                     // first cell right of the block
        uint32_t d = 0;
        if (top && y4 == 0) d |= 4u;
        if (left && x4 == 0) d |= 1u;
        if (tl) d |= (x4 == 0 && y4 == 0) ? 2u : (x4 == 0 ? 1u : (y4 == 0 ? 4u : 0u));
        const bool beyond = ext && xe * 4 + ext > cells * 4;  // the row above is read past the macroblock's right edge
        if (beyond && y4 == 0) d |= 8u;
        if (cur_mbx == 0 && W != S) d &= ~3u;              // reads land in the zero padding of the previous row
        if (cur_mbx == P.mbw_ - 1 && W != S) d &= ~8u;
        if (cur_mby == 0) d &= ~14u;
        // Width == Stride, last column, block below the MB's top edge: "right of the macroblock" wraps onto the next
        // pixel row, columns 0.., which belong to the first macroblock of this MB row (index m-mbw+1), already decoded
        if (beyond && y4 > 0 && cur_mbx == P.mbw_ - 1 && W == S && P.mbw_ > 1) d |= 8u;
        cur_deps |= d;
    }
    void emit_op(uint32_t mode, bool res, int plane, int x4, int y4, int delta, int mboff) {
        if (delta < -32768 || delta > 32767) fail(MOBI_ERR_BITSTREAM, "plane-predictor delta outside 16 bits");
        check_intra_reads(mode, plane_off(plane, mboff, x4, y4));
        note_deps(mode, plane, x4, y4, plane_off(plane, mboff, x4, y4));
        if ((mode == 9 || mode == 19) && !res) return;
        out.ops.push_back((mode & 31) | (res ? 32u : 0u) | (uint32_t)plane << 6 | (uint32_t)x4 << 8 | (uint32_t)y4 << 10 | (uint32_t)(uint16_t)(int16_t)delta << 16);
    }
    uint32_t read_mode(int ci, int& nbits) {  // MD:1840-1852
        uint8_t* c = P.st_.ctx;
        uint32_t pred = c[ci - 8], l = c[ci - 1];
        if (pred > l) pred = l;
        if (pred == 9) pred = 3;
        uint32_t x = b.win >> 28;
        if (x >= pred) x++;
        if (x < 9) { nbits = 4; return x; }
        nbits = 1;
        return pred;
    }
    int delta_for(uint32_t mode) { return (mode == 2 || mode == 12) ? b.svar() : 0; }  // read inside PredictIntra (MD:1917, 2498)

    // sub_116508 MD:2869: coded block whose predictor comes from the MB header
    void coded_fixed(int plane, int x4, int y4, uint32_t m, uint8_t blk, int mboff, uint32_t& mask) {
        if (b.win >> 31) {
            b.win += b.win; b.nb--;
            emit_op(m, true, plane, x4, y4, 0, mboff);
            coefs(64, blk, 0, mask);
        } else {
            m += 10;
            uint32_t cbp4 = tab(MOBI_CBP4_INTRA, 20, b.uvar());
            for (int k = 0; k < 4; k++) {
                bool res = (cbp4 >> k) & 1;
                emit_op(m, res, plane, x4 + (k & 1), y4 + (k >> 1), 0, mboff);
                if (res) coefs(16, blk, (uint8_t)k, mask);
            }
        }
    }
    void chroma(uint32_t cbp6, int mboff, uint32_t& mask) {  // loc_116290 MD:1864
        uint32_t m = b.take(3);
        if (m == 2) {
            m = 9;
            int du = b.svar();
            emit_op(2, false, 1, 0, 0, du, mboff);
            int dv = b.svar();
            emit_op(2, false, 2, 0, 0, dv, mboff);
        }
        for (int p = 1; p <= 2; p++) {
            if ((cbp6 >> (3 + p)) & 1) coded_fixed(p, 0, 0, m, (uint8_t)(3 + p), mboff, mask);
            else emit_op(m, false, p, 0, 0, 0, mboff);
        }
    }
    void intra_full(int mboff, uint32_t& mask) {  // DecIntraFullBlockPMode MD:1759
        uint32_t cbp6 = tab(MOBI_CBP6_INTRA, 64, b.uvar());
        uint32_t m = b.take(3);
        if (m == 2) { m = 9; int d = b.svar(); emit_op(20, false, 0, 0, 0, d, mboff); }
        for (int k = 0; k < 4; k++) {
            int x4 = (k & 1) * 2, y4 = (k >> 1) * 2;
            if ((cbp6 >> k) & 1) coded_fixed(0, x4, y4, m, (uint8_t)k, mboff, mask);
            else emit_op(m, false, 0, x4, y4, 0, mboff);
        }
        chroma(cbp6, mboff, mask);
    }
    void intra_sub(int mboff, uint32_t& mask) {  // DecIntraSubBlockPMode MD:1789
        uint32_t cbp6 = tab(MOBI_CBP6_INTRA, 64, b.uvar());
        static const int ci[4] = {9, 0xB, 0x19, 0x1B}, dc[4] = {0, 1, 8, 9};
        uint8_t* c = P.st_.ctx;
        for (int k = 0; k < 4; k++) {
            int x4 = (k & 1) * 2, y4 = (k >> 1) * 2, n;
            if (!((cbp6 >> k) & 1)) {  // loc_116220 MD:1835
                uint32_t m = read_mode(ci[k], n);
                c[ci[k]] = c[ci[k] + 1] = c[ci[k] + 8] = c[ci[k] + 9] = (uint8_t)m;
                b.drop(n); b.chk();
                int d = delta_for(m);
                emit_op(m, false, 0, x4, y4, d, mboff);
            } else if ((b.win >> 31) & 1) {  // loc_116368 MD:2776, one 8x8
                b.win <<= 1; b.nb--;
                uint32_t m = read_mode(ci[k], n);
                b.drop(n); b.chk();
                c[ci[k]] = c[ci[k] + 1] = c[ci[k] + 8] = c[ci[k] + 9] = (uint8_t)m;
                int d = delta_for(m);
                emit_op(m, true, 0, x4, y4, d, mboff);
                coefs(64, (uint8_t)k, 0, mask);
            } else {  // four 4x4s, each with its own mode (sub_1163DC MD:2836)
                uint32_t cbp4 = tab(MOBI_CBP4_INTRA, 20, b.uvar());
                for (int j = 0; j < 4; j++) {
                    int cj = ci[k] + dc[j];
                    uint32_t m = read_mode(cj, n);
                    c[cj] = (uint8_t)m;
                    m += 10;
                    b.drop(n); b.chk();
                    bool res = (cbp4 >> j) & 1;
                    int d = delta_for(m);
                    emit_op(m, res, 0, x4 + (j & 1), y4 + (j >> 1), d, mboff);
                    if (res) coefs(16, (uint8_t)k, (uint8_t)j, mask);
                }
            }
        }
        chroma(cbp6, mboff, mask);
    }
    void intra_mb(bool sub, int mboff) {
        own_tag = (uint8_t)((out.mbs.size() & 3u) << 3);
        mobi_mb mb;
        uint32_t first_op = (uint32_t)out.ops.size(), first_coef = (uint32_t)out.coefs.size(), mask = 0;
        cur_deps = 0;
        cur_mbx = (mboff & (S - 1)) >> 4; cur_mby = (mboff >> log2S) >> 4;
        if (sub) intra_sub(mboff, mask); else intra_full(mboff, mask);
        uint32_t nops = (uint32_t)out.ops.size() - first_op, nco = (uint32_t)out.coefs.size() - first_coef;
        mb.info = 1u | nops << 2 | nco << 9 | mask << 18 | cur_deps << 24;
        mb.first_sub = first_op; mb.first_coef = first_coef; mb.intra_rank = (uint32_t)out.intra.size();
        out.intra.push_back((uint32_t)out.mbs.size());
        out.mbs.push_back(mb);
    }

    // ---- inter -----------------------------------------------------------------------------
    void leaf(int lw, int lh, uint32_t ref, int dx, int dy, int off, int mboff, int slot) {
        P.mvc_[slot] = dx; P.mvc_[slot + 1] = dy;  // MD:411-412, last leaf wins
        int w = 2 << lw, h = 2 << lh;
        if ((int)ref > P.st_.decoded) fail(MOBI_ERR_REFERENCE, "P-frame references a picture that is not in the ring");
        // the packed record holds 16-bit vectors; with that established, 32-bit arithmetic below cannot overflow
        if (dx < -32768 || dx > 32767 || dy < -32768 || dy > 32767) fail(MOBI_ERR_RANGE, "motion vector outside 16 bits");
        // CopyBlock reads (MD:418-456): luma, then both chroma planes at (dx>>1, dy>>1), half size
        const int first = off + (dy >> 1) * S + (dx >> 1);   // (multiplications: the vector components may be negative)
        const int last = first + (h - 1 + (dy & 1)) * S + w - 1 + (dx & 1);
        if (first < 0 || last >= H * S) fail(MOBI_ERR_RANGE, "motion vector reads outside the luma array");
        const int cdx = dx >> 1, cdy = dy >> 1;
        const int cfirst = (off >> 1) + (cdy >> 1) * S + (cdx >> 1);
        const int clast = cfirst + (S >> 1) + ((h >> 1) - 1 + (cdy & 1)) * S + (w >> 1) - 1 + (cdx & 1);
        if (cfirst < 0 || clast >= H * S / 2) fail(MOBI_ERR_RANGE, "motion vector reads outside the chroma array");
        int rel = off - mboff, x = rel & (S - 1), y = rel >> log2S;
        mobi_part p;
        p.xy = (uint8_t)((x >> 1) | (y >> 1) << 4);
        p.shape = (uint8_t)(lw | lh << 2 | ref << 4);
        p.mvx = (int16_t)dx; p.mvy = (int16_t)dy; p.pad = 0;
        out.parts.push_back(p);
        if (ref > max_ref) max_ref = ref;
    }
    // ReadPBlockWxH + SwitchPBlockWxH (MD:469-1746); returns false when the MB turned out to be intra
    bool pblock(int lw, int lh, int off, int mboff, int slot) {
        const mobi_part_code_t& pc = MOBI_PART_CODE[P.ver_ == MOBI_MOFLEX3DS ? 0 : 1][lw][lh];
        uint32_t sym = pc.sym[b.win >> (32 - pc.peek)];
        int n = pc.len[sym];
        b.drop(n); b.chk();
        int w = 2 << lw, h = 2 << lh;
        bool top = lw == 3 && lh == 3;
        if (sym <= 5) {
            int dx = mvpx, dy = mvpy;
            uint32_t ref = 1;
            if (sym) { int ax = b.svar(); int ay = b.svar(); dx += ax; dy += ay; ref = sym; }
            leaf(lw, lh, ref, dx, dy, off, mboff, slot);
        } else if (sym == 8 && lh > 0) {
            pblock(lw, lh - 1, off, mboff, slot);
            pblock(lw, lh - 1, off + S * (h / 2), mboff, slot);
        } else if (sym == 9 && lw > 0) {
            pblock(lw - 1, lh, off, mboff, slot);
            pblock(lw - 1, lh, off + w / 2, mboff, slot);
        } else if (top && (sym == 6 || sym == 7)) {
            intra_mb(sym == 7, mboff);
            return false;
        } else fail(MOBI_ERR_BITSTREAM, "illegal partition code");
        return true;
    }
    void blk8_inter(uint8_t blk, uint32_t& mask) {  // loc_11652C MD:2909
        if ((b.win >> 31) & 1) { b.win += b.win; b.nb--; coefs(64, blk, 0, mask); }
        else {
            uint32_t cbp4 = tab(MOBI_CBP4_INTER, 16, b.uvar());
            for (uint32_t m = cbp4 & 15u; m; m &= m - 1) coefs(16, blk, (uint8_t)__builtin_ctz(m), mask);
        }
    }
    void inter_mb(int mboff, int slot) {
        uint32_t first_part = (uint32_t)out.parts.size(), first_coef = (uint32_t)out.coefs.size(), mask = 0;
        own_tag = (uint8_t)((out.mbs.size() & 3u) << 3);
        cur_m8 = 0;
        if (!pblock(3, 3, mboff, mboff, slot)) return;
        uint32_t cbp6 = tab(MOBI_CBP6_INTER, 64, b.uvar());  // loc_1161A0 MD:1818
        for (uint32_t m = cbp6 & 63u; m; m &= m - 1) blk8_inter((uint8_t)__builtin_ctz(m), mask);   // set bits only, ascending: no coin-flip branch per block
        mobi_mb mb;
        uint32_t np = (uint32_t)out.parts.size() - first_part, nco = (uint32_t)out.coefs.size() - first_coef;
        mb.info = 0u | np << 2 | nco << 9 | mask << 18 | (cur_m8 & 15u) << 24 | (cur_m8 >> 4) << 29;   // bits 24-27, 29-30: which coded blocks are 8x8-transformed
        mb.first_sub = first_part; mb.first_coef = first_coef; mb.intra_rank = 0;
        if (np == 1) {  // an unsplit macroblock carries its vector inside the descriptor (saves the device a dependent load)
            const mobi_part& p = out.parts[first_part];
            if (p.mvx >= -8192 && p.mvx < 8192 && p.mvy >= -8192 && p.mvy < 8192) {
                mb.info |= 1u << 28;
                mb.intra_rank = ((uint32_t)p.mvx & 0x3FFFu) | ((uint32_t)p.mvy & 0x3FFFu) << 14 | (uint32_t)(p.shape >> 4) << 28;
            }
        }
        out.mbs.push_back(mb);
        out.hdr.n_inter_coefs += nco;
    }

    static int med3(int a, int c, int e) {  // the sorting network of MD:171-188 leaves the median in the middle; min/max form: no branches
        const int lo = a < c ? a : c, hi = a < c ? c : a;
        const int m = hi < e ? hi : e;
        return lo > m ? lo : m;
    }

    void run(const uint8_t* data, int len, int start) {
        b.d = data; b.len = len; b.off = start; b.nb = 0;
        b.win = b.u16(b.off) << 16;
        b.off += 2;
        uint32_t intra = b.win >> 31;
        b.win += b.win;
        out.hdr.flags = intra;
        if (!intra) {
            if (--b.nb < 0) b.fill();
            if (P.ver_ == MOBI_MOFLEX3DS) {
                uint32_t q = P.st_.quant;
                int dq = b.svar();
                if (q == 0) setup_quant(q);
                else if (dq != 0) setup_quant((uint32_t)(q + dq));
            } else {
                int dq = b.svar();
                if (dq != 0) setup_quant((uint32_t)(P.st_.quant + dq));
            }
            vlcsel = 0;
            std::fill(P.mvc_.begin(), P.mvc_.end(), 0);
            int off = 0, h = H;
            do {
                int w = W, mx = 0;
                do {
                    const int32_t* e = &P.mvc_[2 * mx];
                    mvpx = med3(e[0], e[2], e[4]); mvpy = med3(e[1], e[3], e[5]);
                    int slot = 2 * (mx + 1);
                    P.mvc_[slot] = P.mvc_[slot + 1] = 0;
                    out.room_for_mb();
                    inter_mb(off, slot);
                    off += 16; w -= 16; mx++;
                } while (w > 0);
                off += S * 16 - W; h -= 16;
            } while (h > 0);
        } else {
            P.st_.yuvfmt = b.win >> 31; b.win += b.win;
            vlcsel = b.win >> 31; b.win += b.win;
            b.nb -= 3; b.chk();
            uint32_t q = b.win >> 26;
            b.drop(6); b.chk();
            if (P.st_.quant != q) setup_quant(q);
            int off = 0, h = H;
            do {
                int w = W;
                do {
                    uint32_t sub = b.win >> 31;
                    b.win += b.win; b.nb--; b.chk();
                    out.room_for_mb();
                    intra_mb(sub != 0, off);
                    off += 16; w -= 16;
                } while (w > 0);
                off += S * 16 - W; h -= 16;
            } while (h > 0);
        }
    }
};

Parser::Parser(uint32_t w, uint32_t h, int version) : W_(w), H_(h), ver_(version), S_(stride_for(w)), mbw_((int)w / 16), mbh_((int)h / 16) {
    std::call_once(g_once, build_tables);
    mvc_.assign(2 * (mbw_ + 3), 0);
}

void Parser::reset() { st_ = State(); std::fill(mvc_.begin(), mvc_.end(), 0); err_.clear(); }

int Parser::parse(const uint8_t* data, int len, int* offset, ParsedFrame& out) {
    if (!data || !offset || len < 0) { err_ = "null argument"; return MOBI_ERR_ARG; }
    if (ver_ != MOBI_MODSDS && ver_ != MOBI_MOFLEX3DS) { err_ = "VxDS (VXS1) is not implemented by the reference either"; return MOBI_ERR_UNSUPPORTED; }
    State saved = st_;
    out.clear();
    std::memset(&out.hdr, 0, sizeof out.hdr);
    FrameParse fp(*this, out);
    try {
        fp.run(data, len, *offset);
    } catch (const ParseError& e) {
        st_ = saved;
        err_ = e.msg;
        return e.code;
    }
    mobi_frame_hdr& h = out.hdr;
    h.n_mb = (uint32_t)out.mbs.size(); h.n_parts = (uint32_t)out.parts.size(); h.n_ops = (uint32_t)out.ops.size();
    h.n_coefs = (uint32_t)out.coefs.size(); h.n_intra = (uint32_t)out.intra.size();
    h.quantizer = st_.quant; h.yuv_format = st_.yuvfmt;
    h.bytes_consumed = (uint32_t)(fp.b.off - *offset);
    h.max_ref = fp.max_ref;
    std::memcpy(h.qtab, st_.qtab, sizeof h.qtab);
    *offset = fp.b.off;
    if (st_.decoded < 6) st_.decoded++;
    err_.clear();
    return MOBI_OK;
}

}  // namespace mobi

// ---- C ABI: host-only parser ---------------------------------------------------------------------
struct mobi_parser {
    mobi::Parser p;
    mobi::ParsedFrame f;
    mobi_parser(uint32_t w, uint32_t h, int v) : p(w, h, v) {}
};

extern "C" {

int mobi_parser_create(uint32_t width, uint32_t height, int version, mobi_parser_t** out) {
    if (!out) return MOBI_ERR_ARG;
    *out = nullptr;
    if (width == 0 || height == 0 || (width & 15) || (height & 15) || width > 1024 || height > 1024) return MOBI_ERR_ARG;
    if (version != MOBI_MODSDS && version != MOBI_MOFLEX3DS) return version == MOBI_VXDS ? MOBI_ERR_UNSUPPORTED : MOBI_ERR_ARG;
    try { *out = new mobi_parser(width, height, version); } catch (...) { return MOBI_ERR_NOMEM; }
    return MOBI_OK;
}
void mobi_parser_destroy(mobi_parser_t* p) { delete p; }
int mobi_parser_parse(mobi_parser_t* p, const uint8_t* data, int len, int* offset_inout, mobi_packed_frame* out) {
    if (!p || !out) return MOBI_ERR_ARG;
    int rc;
    try { rc = p->p.parse(data, len, offset_inout, p->f); } catch (...) { return MOBI_ERR_NOMEM; }
    if (rc == MOBI_OK) *out = p->f.view();
    return rc;
}
const char* mobi_parser_last_error(const mobi_parser_t* p) { return p ? p->p.error().c_str() : "null parser"; }
int mobicuda_abi_version(void) { return MOBICUDA_ABI_VERSION; }

}  // extern "C"
