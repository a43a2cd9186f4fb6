/* TEST INFRASTRUCTURE -- see mobi_oracle.h.  Plain-C restatement of the reference decoder.
 * "MD:n" = LibMobiclip/Codec/Mobiclip/MobiclipDecoder.cs line n; "MC:n" = MobiConst.cs line n.
 *
 * Conventions kept from the reference because they are observable:
 *  - one u32 scratch array I[392] (MD:28) holding, by word index: bytes 0..39 intra-mode context
 *    grid, [10..73] 8x8 (scale<<8|scan) words, [74..89] 4x4 words, [90..153] coefficient block,
 *    [154..217] transform scratch, [218] VLC table select, [219..220] MV predictor, [221..] MV row
 *    cache.  Run overflows and q<12 leaks therefore land where they land in the reference.
 *  - planes are flat byte arrays addressed Offset = y*Stride + x, never clamped; any access
 *    outside an array aborts the frame (C# exception -> catch-all MD:325 -> null).
 *  - shift counts are masked to 5 bits as C# does for 32-bit operands.
 */
#include "mobi_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <setjmp.h>
#include "../mobiclipdecoder_b200/csrc/mobi_tables.h"

enum { VER_VXDS = 0, VER_MODS = 1, VER_MOFLEX = 2 };
enum { I_Q8 = 10, I_Q4 = 74, I_COEF = 90, I_TMP = 154, I_VLC = 218, I_MVP = 219, I_MVC = 221, I_LEN = 392 };

typedef struct { uint8_t* p; int n; } plane_t;

struct mobi_oracle {
    uint32_t W, H;
    int ver, S;
    plane_t Y[6], UV[6];
    uint32_t quant, yuvfmt;
    uint32_t I[I_LEN];
    uint16_t vlc[2][4096];
    const uint8_t* d;
    int dlen, off;
    uint32_t win;
    int nb;
    jmp_buf jb;
};

#define ABORT(o) longjmp((o)->jb, 1)

/* ---- managed-array accessors ------------------------------------------------------------ */
static inline uint8_t* px(mobi_oracle_t* o, plane_t pl, long long i) {
    if (!pl.p || i < 0 || i >= pl.n) ABORT(o);
    return pl.p + i;
}
#define RD(pl, i) (*px(o, (pl), (long long)(i)))
static inline uint32_t* iw(mobi_oracle_t* o, long long i) {
    if (i < 0 || i >= I_LEN) ABORT(o);
    return &o->I[i];
}
static inline uint32_t rd32(mobi_oracle_t* o, plane_t pl, long long i) { /* IOUtil.ReadU32LE, IO:75 */
    uint32_t b3 = RD(pl, i + 3), b2 = RD(pl, i + 2), b1 = RD(pl, i + 1), b0 = RD(pl, i);
    return b3 << 24 | b2 << 16 | b1 << 8 | b0;
}
static inline void wr32(mobi_oracle_t* o, plane_t pl, long long i, uint32_t v) { /* IOUtil.WriteU32LE, IO:80 */
    RD(pl, i) = (uint8_t)v; RD(pl, i + 1) = (uint8_t)(v >> 8); RD(pl, i + 2) = (uint8_t)(v >> 16); RD(pl, i + 3) = (uint8_t)(v >> 24);
}

/* ---- bit reader (MD:2970-3015, 3927-3937) ----------------------------------------------- */
static uint32_t rd16(mobi_oracle_t* o, int at) { /* IOUtil.ReadU16LE, IO:39 */
    if (at < 0 || at + 1 >= o->dlen) ABORT(o);
    return (uint32_t)o->d[at] | (uint32_t)o->d[at + 1] << 8;
}
static void fill(mobi_oracle_t* o) { /* FillBits MD:2988 */
    if (o->off >= o->dlen) return;
    uint32_t w = rd16(o, o->off);
    o->off += 2;
    o->nb += 16;
    o->win |= w << ((16 - o->nb) & 31);
}
static inline void drop(mobi_oracle_t* o, int n) { o->win <<= (n & 31); o->nb -= n; }
static inline void chk(mobi_oracle_t* o) { if (o->nb < 0) fill(o); }
static inline uint32_t take(mobi_oracle_t* o, int n) { /* n in 1..31 */
    uint32_t v = o->win >> (32 - n);
    drop(o, n); chk(o);
    return v;
}
static int clz32(uint32_t v) { int n = 0; while (v) { v >>= 1; n++; } return 32 - n; }
static uint32_t gamma_raw(mobi_oracle_t* o) { /* shared body of MD:2970 / MD:2998 */
    int z = clz32(o->win);
    o->win <<= (z & 31);
    o->win += o->win;
    int sh = 32 - z;
    uint32_t v = (sh == 32) ? 0 : o->win >> (sh & 31);
    v += (uint32_t)(1 << (z & 31));
    o->win <<= (z & 31);
    o->nb -= z << 1;
    if (--o->nb < 0) fill(o);
    return v;
}
static uint32_t uvar(mobi_oracle_t* o) { return gamma_raw(o) - 1; }
static int svar(mobi_oracle_t* o) {
    int v = (int)gamma_raw(o);
    if (v & 1) v = 1 - v;
    return v >> 1;
}

/* ---- quantiser tables (MD:3884-3925) ---------------------------------------------------- */
static void setup_quant(mobi_oracle_t* o, uint32_t q) {
    if (o->ver == VER_MOFLEX) { if (q < 12) q = 12; if (q > 52) q = 52; }
    o->quant = q;
    if (q >= MOBI_QTAB_MAXQ) ABORT(o); /* byte_119004[q] out of range */
    int sh = (int)(q / 6) + 8, row = (int)(q % 6);
    for (int i = 0; i < 16; i++) o->I[I_Q4 + i] = (uint32_t)MOBI_SCAN4[i] | (uint32_t)MOBI_SCALE4[row * 16 + i] << sh;
    sh -= 2;
    for (int i = 0; i < 64; i++) o->I[I_Q8 + i] = (uint32_t)MOBI_SCAN8[i] | (uint32_t)MOBI_SCALE8[row * 64 + i] << sh;
    uint8_t* c = (uint8_t*)o->I; /* context-grid border = "no mode" (9) */
    c[1] = c[2] = c[3] = c[4] = 9; c[8] = c[0x10] = c[0x18] = c[0x20] = 9;
}

/* ---- residual VLC (MD:3330-3432) -------------------------------------------------------- */
static void read_coefs(mobi_oracle_t* o, uint32_t* pos) {
    const uint16_t* A = o->vlc[o->I[I_VLC] == 1];
    const uint8_t* B = o->I[I_VLC] == 1 ? MOBI_VLC1_ESC : MOBI_VLC0_ESC;
    for (;;) {
        int run, level, nbits;
        uint32_t e, last;
        if ((o->win >> 25) == 3) {
            o->win <<= 7;
            uint32_t c = o->win >> 31; o->win <<= 1;
            if (!c) { /* level-offset escape */
                o->nb -= 8; chk(o);
                e = A[o->win >> 20];
                int add = B[e >> 9];
                nbits = e & 15; e >>= 4; level = (int)(e & 31) + add; e >>= 5;
                o->win <<= ((nbits - 1) & 31);
                if (o->win >> 31) level = -level;
                o->win <<= 1; o->nb -= nbits; chk(o);
                run = e & 63; last = e >> 6;
            } else {
                c = o->win >> 31; o->win <<= 1;
                if (!c) { /* run-offset escape */
                    o->nb -= 9; chk(o);
                    e = A[o->win >> 20];
                    nbits = e & 15; e >>= 4; level = e & 31; e >>= 5;
                    uint32_t r = e & 63; e >>= 6;
                    uint32_t bi = 0x80 + (uint32_t)level + (e << 6);
                    if (bi >= 256) ABORT(o);
                    int add = B[bi];
                    o->win <<= ((nbits - 1) & 31);
                    if (o->win >> 31) level = -level;
                    o->win <<= 1; o->nb -= nbits; chk(o);
                    run = (int)r + add; last = e;
                } else { /* raw escape: last(1) run(6) level(s12) */
                    o->nb -= 9; chk(o);
                    last = o->win >> 31; o->win <<= 1;
                    run = o->win >> 26; o->win <<= 6;
                    o->nb -= 7; chk(o);
                    level = (int32_t)o->win >> 20; o->win <<= 12;
                    o->nb -= 12; chk(o);
                }
            }
        } else {
            e = A[o->win >> 20];
            nbits = e & 15; e >>= 4; level = e & 31; e >>= 5;
            o->win <<= ((nbits - 1) & 31);
            if (o->win >> 31) level = -level;
            o->win <<= 1; o->nb -= nbits; chk(o);
            run = e & 63; last = e >> 6;
        }
        *pos = (uint32_t)(*pos + run);
        uint32_t w = *iw(o, (*pos)++);
        int32_t scaled = (int32_t)((uint32_t)(int32_t)(w >> 8) * (uint32_t)level);
        *iw(o, I_COEF + (long long)(w & 0xFF)) = (uint32_t)scaled;
        if (last & 1) break;
    }
}

/* ---- inverse transforms + add + clip (MD:3435-3798; clip table MC:587 = clamp(i-64,0,255)) */
static inline uint8_t clipadd(mobi_oracle_t* o, int pix, int delta) {
    int v = pix + delta;
    if (v < -64 || v > 319) ABORT(o); /* MinMaxTable index out of range */
    return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
}
static inline void bfly8(const int32_t in[8], int32_t out[8]) { /* one 8-point pass, MD:3452-3485 */
    int32_t a0 = in[0] + in[4], a1 = in[0] - in[4];
    int32_t a2 = in[2] + (in[6] >> 1), a3 = (in[2] >> 1) - in[6];
    int32_t e0 = a0 + a2, e3 = a0 - a2, e1 = a1 + a3, e2 = a1 - a3;
    int32_t b0 = in[1] + in[7] - in[3] - (in[3] >> 1);
    int32_t b1 = in[7] - in[1] + in[5] + (in[5] >> 1);
    int32_t b2 = in[5] - (in[7] + (in[7] >> 1)) - in[3];
    int32_t b3 = in[3] + in[5] + in[1] + (in[1] >> 1);
    int32_t o1 = b2 + (b3 >> 2), o7 = b3 - (b2 >> 2);
    int32_t o3 = b0 + (b1 >> 2), o5 = (b0 >> 2) - b1;
    out[0] = e0 + o7; out[7] = e0 - o7;
    out[1] = e1 + o5; out[6] = e1 - o5;
    out[2] = e2 + o3; out[5] = e2 - o3;
    out[3] = e3 + o1; out[4] = e3 - o1;
}
static void idct8_add(mobi_oracle_t* o, plane_t dst, int off) {
    int32_t* c = (int32_t*)&o->I[I_COEF];
    int32_t* t = (int32_t*)&o->I[I_TMP];
    int32_t in[8], out[8];
    c[0] += 32;
    for (int i = 0; i < 8; i++) {
        for (int k = 0; k < 8; k++) in[k] = c[8 * i + k];
        bfly8(in, out);
        for (int k = 0; k < 8; k++) t[8 * k + i] = out[k];
    }
    for (int r = 0; r < 8; r++) {
        for (int k = 0; k < 8; k++) in[k] = t[8 * r + k];
        bfly8(in, out);
        for (int k = 0; k < 8; k++) { uint8_t* p = px(o, dst, (long long)off + k); *p = clipadd(o, *p, out[k] >> 6); }
        off += o->S;
    }
}
static inline void bfly4(const int32_t in[4], int32_t out[4]) { /* MD:3740-3747 */
    int32_t s = in[0] + in[2], d = in[0] - in[2];
    int32_t p = (in[1] >> 1) - in[3], q = in[1] + (in[3] >> 1);
    out[0] = s + q; out[3] = s - q; out[1] = d + p; out[2] = d - p;
}
static void idct4_add(mobi_oracle_t* o, plane_t dst, int off) {
    int32_t* c = (int32_t*)&o->I[I_COEF];
    int32_t t[16], in[4], out[4];
    c[0] += 32;
    for (int i = 0; i < 4; i++) {
        for (int k = 0; k < 4; k++) in[k] = c[4 * i + k];
        bfly4(in, out);
        for (int k = 0; k < 4; k++) t[4 * k + i] = out[k];
    }
    for (int r = 0; r < 4; r++) {
        for (int k = 0; k < 4; k++) in[k] = t[4 * r + k];
        bfly4(in, out);
        for (int k = 0; k < 4; k++) { uint8_t* p = px(o, dst, (long long)off + k); *p = clipadd(o, *p, out[k] >> 6); }
        off += o->S;
    }
}
static void residual8(mobi_oracle_t* o, plane_t dst, int off) { /* loc_116540 MD:2931 */
    for (int i = 0; i < 64; i++) o->I[I_COEF + i] = 0;
    uint32_t pos = I_Q8;
    read_coefs(o, &pos);
    idct8_add(o, dst, off); /* the 1/3/16-coefficient variants (MD:2939-2941) are specialisations */
}
static void residual4(mobi_oracle_t* o, plane_t dst, int off) { /* sub_1166E8 MD:2958 */
    for (int i = 0; i < 16; i++) o->I[I_COEF + i] = 0;
    uint32_t pos = I_Q4;
    read_coefs(o, &pos);
    idct4_add(o, dst, off);
}

/* ---- intra prediction (MD:1883-2774, plane predictors MD:3017-3327) --------------------- */
static void pack_row(mobi_oracle_t* o, plane_t dst, long long off, const int32_t* v, int n) {
    /* the reference ORs unclipped values into a u32 per 4 pixels (MD:3064-3074, 3212-3219, 3314-3321) */
    for (int g = 0; g < n; g += 4) {
        uint32_t w = (uint32_t)v[g] | (uint32_t)v[g + 1] << 8 | (uint32_t)v[g + 2] << 16 | (uint32_t)v[g + 3] << 24;
        wr32(o, dst, off + g, w);
    }
}
static void plane16(mobi_oracle_t* o, plane_t dst, int off) { /* sub_1167BC MD:3017 */
    int S = o->S, d = svar(o);
    int32_t T[16], A[16], B[16], v[16];
    if (off - S < 0 || off - S + 16 > dst.n || !dst.p) ABORT(o);
    for (int i = 0; i < 16; i++) T[i] = dst.p[off - S + i];
    int32_t l = RD(dst, (long long)off + S * 15 - 1), t = T[15];
    int32_t m = ((l + t + 1) >> 1) + d * 2;
    int32_t gx = m - l + 1, acc = l << 3;
    for (int i = 0; i < 16; i++) { acc += gx >> 1; A[i] = T[i] * 64; B[i] = acc - T[i] * 8 + 1; }
    int32_t gy = m - t + 1, ry = t << 3;
    for (int r = 0; r < 16; r++) {
        ry += gy >> 1;
        int32_t L = RD(dst, (long long)off - 1);
        int32_t step = ry - (L << 3) + 1, run = L << 6;
        for (int i = 0; i < 16; i++) { A[i] += B[i] >> 1; run += step >> 1; v[i] = (A[i] + run + 64) >> 7; }
        pack_row(o, dst, off, v, 16);
        off += S;
    }
}
static void plane8(mobi_oracle_t* o, plane_t dst, int off) { /* sub_116CCC MD:3168 */
    int S = o->S, d = svar(o);
    int32_t T[8], A[8], B[8], v[8];
    if (off - S < 0 || off - S + 8 > dst.n || !dst.p) ABORT(o);
    for (int i = 0; i < 8; i++) T[i] = dst.p[off - S + i];
    int32_t l = RD(dst, (long long)off + S * 7 - 1), t = T[7];
    int32_t m = ((l + t + 1) >> 1) + d * 2;
    int32_t gx = m - l, acc = l * 8;
    for (int i = 0; i < 8; i++) { acc += gx; A[i] = T[i] * 64; B[i] = acc - T[i] * 8; }
    int32_t gy = m - t, ry = t << 3;
    for (int r = 0; r < 8; r++) {
        ry += gy;
        int32_t L = RD(dst, (long long)off - 1);
        int32_t step = ry - L * 8, run = L * 64;
        for (int i = 0; i < 8; i++) { A[i] += B[i]; run += step; v[i] = (A[i] + run + 64) >> 7; }
        pack_row(o, dst, off, v, 8);
        off += S;
    }
}
static void plane4(mobi_oracle_t* o, plane_t dst, int off) { /* sub_117E98 MD:3253 */
    int S = o->S, d = svar(o);
    int32_t T[4], A[4], B[4], v[4];
    uint32_t tw = rd32(o, dst, (long long)off - S);
    for (int i = 0; i < 4; i++) T[i] = (tw >> (8 * i)) & 0xFF;
    int32_t l = RD(dst, (long long)off + S * 3 - 1), t = T[3];
    int32_t m = ((l + t + 1) >> 1) + d * 2;
    int32_t gx = m - l, acc = l << 2;
    for (int i = 0; i < 4; i++) { acc += gx; A[i] = T[i] << 4; B[i] = acc - (T[i] << 2); }
    int32_t gy = m - t, ry = t << 2;
    for (int r = 0; r < 4; r++) {
        ry += gy;
        int32_t L = RD(dst, (long long)off - 1);
        int32_t step = ry - (L << 2), run = L << 4;
        for (int i = 0; i < 4; i++) { A[i] += B[i]; run += step; v[i] = (A[i] + run + 16) >> 5; }
        pack_row(o, dst, off, v, 4);
        off += S;
    }
}

/* Directional predictors, N = 8 (modes 0..8) or 4 (modes 10..18).  The reference spells each one
 * out as register shuffles (MD:1890-2472, 2475-2769); written here as the closed forms they
 * compute (H.264-style, unfiltered edges).  e[k] = TL for k = -1, top T[k] for k >= 0 ... */
static void predict_dir(mobi_oracle_t* o, int mode, int N, plane_t dst, int off, int isV) {
    int S = o->S;
    int T[13], L[8], TL = 0, P[8][8];
    int needT = 0, needL = 0, needTL = 0;
    switch (mode) {
    case 0: needT = N; break;
    case 1: needL = N; break;
    case 4: needL = N; break;
    case 5: needT = N; needL = N; needTL = 1; break;   /* 4x4 reads the whole top word (MD:2638) */
    case 6: needT = N; needL = N - 1; needTL = 1; break;
    case 7: needT = N; needL = N; needTL = 1; break;
    case 8: needT = (N == 8) ? 13 : 8; break;          /* 4x4 reads two top words (MD:2737, 2746) */
    default: break;
    }
    if (mode == 3) {
        int left = ((off - (isV ? S / 2 : 0)) % S) != 0, top = off >= S;
        needT = top ? N : 0; needL = left ? N : 0;
        for (int i = 0; i < needT; i++) T[i] = RD(dst, (long long)off - S + i);
        for (int i = 0; i < needL; i++) L[i] = RD(dst, (long long)off + (long long)i * S - 1);
        int sum = 0, dc;
        for (int i = 0; i < needT; i++) sum += T[i];
        for (int i = 0; i < needL; i++) sum += L[i];
        if (top && left) dc = (sum + N) / (2 * N);
        else if (top || left) dc = (sum + N / 2) / N;
        else dc = 0x80;
        for (int y = 0; y < N; y++) for (int x = 0; x < N; x++) RD(dst, (long long)off + (long long)y * S + x) = (uint8_t)dc;
        return;
    }
    for (int i = 0; i < needT; i++) T[i] = RD(dst, (long long)off - S + i);
    if (needTL) TL = RD(dst, (long long)off - S - 1);
    for (int i = 0; i < needL; i++) L[i] = RD(dst, (long long)off + (long long)i * S - 1);
#define TT(k) ((k) < 0 ? TL : T[k])
#define LL(k) ((k) < 0 ? TL : L[k])
    for (int y = 0; y < N; y++) for (int x = 0; x < N; x++) {
        int p = 0;
        switch (mode) {
        case 0: p = T[x]; break;
        case 1: p = L[y]; break;
        case 4: { /* horizontal-up */
            int z = x + 2 * y, k = y + (x >> 1);
            if (z > 2 * N - 3) p = L[N - 1];
            else if (z == 2 * N - 3) p = (L[N - 2] + 3 * L[N - 1] + 2) >> 2;
            else if (z & 1) p = (L[k] + 2 * L[k + 1] + L[k + 2] + 2) >> 2;
            else p = (L[k] + L[k + 1] + 1) >> 1;
            break; }
        case 5: { /* horizontal-down */
            int z = 2 * y - x, k = y - (x >> 1);
            if (z < -1) p = (TT(x - 2 * y - 1) + 2 * TT(x - 2 * y - 2) + TT(x - 2 * y - 3) + 2) >> 2;
            else if (z == -1) p = (L[0] + 2 * TL + T[0] + 2) >> 2;
            else if (z & 1) p = (LL(k - 2) + 2 * LL(k - 1) + LL(k) + 2) >> 2;
            else p = (LL(k - 1) + LL(k) + 1) >> 1;
            break; }
        case 6: { /* vertical-right */
            int z = 2 * x - y, k = x - (y >> 1);
            if (z < -1) p = (LL(y - 2 * x - 1) + 2 * LL(y - 2 * x - 2) + LL(y - 2 * x - 3) + 2) >> 2;
            else if (z == -1) p = (L[0] + 2 * TL + T[0] + 2) >> 2;
            else if (z & 1) p = (TT(k - 2) + 2 * TT(k - 1) + TT(k) + 2) >> 2;
            else p = (TT(k - 1) + TT(k) + 1) >> 1;
            break; }
        case 7: /* diagonal down-right */
            if (x > y) p = (TT(x - y - 2) + 2 * TT(x - y - 1) + TT(x - y) + 2) >> 2;
            else if (x < y) p = (LL(y - x - 2) + 2 * LL(y - x - 1) + LL(y - x) + 2) >> 2;
            else p = (T[0] + 2 * TL + L[0] + 2) >> 2;
            break;
        case 8: { /* vertical-left */
            int k = x + (y >> 1);
            if (y & 1) p = (T[k] + 2 * T[k + 1] + T[k + 2] + 2) >> 2;
            else p = (T[k] + T[k + 1] + 1) >> 1;
            break; }
        }
        P[y][x] = p;
    }
#undef TT
#undef LL
    for (int y = 0; y < N; y++) for (int x = 0; x < N; x++) RD(dst, (long long)off + (long long)y * S + x) = (uint8_t)P[y][x];
}

static void predict_intra(mobi_oracle_t* o, uint32_t mode, plane_t dst, int off) { /* PredictIntra MD:1883 */
    int isV = (dst.p == o->UV[0].p) && (off % o->S) >= o->S / 2; /* VOffsetfix MD:1885-1887 */
    if (mode == 9 || mode == 19 || mode > 19) return;
    if (mode == 2) { plane8(o, dst, off); return; }
    if (mode == 12) { plane4(o, dst, off); return; }
    if (mode < 9) predict_dir(o, (int)mode, 8, dst, off, isV);
    else predict_dir(o, (int)mode - 10, 4, dst, off, isV);
}

/* ---- intra macroblocks (MD:1759-1880, 2776-2902, 2945-2956) ----------------------------- */
static uint32_t tab(mobi_oracle_t* o, const uint8_t* t, int n, uint32_t i) { if (i >= (uint32_t)n) ABORT(o); return t[i]; }

static uint32_t read_mode(mobi_oracle_t* o, int ci, int* nbits) { /* shared by MD:1840-1852, 2785-2797, 2841-2853 */
    uint8_t* c = (uint8_t*)o->I;
    uint32_t pred = c[ci - 8], b = c[ci - 1];
    if (pred > b) pred = b;
    if (pred == 9) pred = 3;
    uint32_t x = o->win >> 28;
    if (x >= pred) x++;
    if (x < 9) { *nbits = 4; return x; }
    *nbits = 1;
    return pred;
}
static void intra_pred_res8(mobi_oracle_t* o, plane_t dst, int off, uint32_t mode) { /* loc_116518 */
    predict_intra(o, mode, dst, off);
    residual8(o, dst, off);
}
static void intra_pred_res4(mobi_oracle_t* o, plane_t dst, int off, uint32_t mode) { /* loc_116628 */
    predict_intra(o, mode, dst, off);
    residual4(o, dst, off);
}
static void intra_coded_fixed(mobi_oracle_t* o, plane_t dst, int off, uint32_t mode) { /* sub_116508 MD:2869 */
    if (o->win >> 31) {
        o->win += o->win; o->nb--;
        intra_pred_res8(o, dst, off, mode);
    } else {
        mode += 10;
        uint32_t cbp4 = tab(o, MOBI_CBP4_INTRA, 20, uvar(o));
        static const int dx[4] = {0, 4, 0, 4}, dy[4] = {0, 0, 4, 4};
        for (int k = 0; k < 4; k++) {
            int so = off + dy[k] * o->S + dx[k];
            if ((cbp4 >> k) & 1) intra_pred_res4(o, dst, so, mode);
            else predict_intra(o, mode, dst, so);
        }
    }
}
static void intra_uncoded_ctx(mobi_oracle_t* o, int ci, plane_t dst, int off) { /* loc_116220 MD:1835 */
    uint8_t* c = (uint8_t*)o->I;
    int n;
    uint32_t m = read_mode(o, ci, &n);
    c[ci] = c[ci + 1] = c[ci + 8] = c[ci + 9] = (uint8_t)m;
    drop(o, n); chk(o);
    predict_intra(o, m, dst, off);
}
static uint32_t read_mode4(mobi_oracle_t* o, int ci) { /* sub_1163DC MD:2836 */
    uint8_t* c = (uint8_t*)o->I;
    int n;
    uint32_t m = read_mode(o, ci, &n);
    c[ci] = (uint8_t)m;
    drop(o, n); chk(o);
    return m + 10;
}
static void intra_coded_ctx(mobi_oracle_t* o, int ci, plane_t dst, int off) { /* loc_116368 MD:2776 */
    uint8_t* c = (uint8_t*)o->I;
    if ((o->win >> 31) & 1) {
        o->win <<= 1; o->nb--;
        int n;
        uint32_t m = read_mode(o, ci, &n);
        drop(o, n); chk(o);
        c[ci] = c[ci + 1] = c[ci + 8] = c[ci + 9] = (uint8_t)m;
        intra_pred_res8(o, dst, off, m);
    } else {
        uint32_t cbp4 = tab(o, MOBI_CBP4_INTRA, 20, uvar(o));
        static const int dx[4] = {0, 4, 0, 4}, dy[4] = {0, 0, 4, 4}, dc[4] = {0, 1, 8, 9};
        for (int k = 0; k < 4; k++) {
            uint32_t m = read_mode4(o, ci + dc[k]);
            int so = off + dy[k] * o->S + dx[k];
            if ((cbp4 >> k) & 1) intra_pred_res4(o, dst, so, m);
            else predict_intra(o, m, dst, so);
        }
    }
}
static void intra_chroma(mobi_oracle_t* o, uint32_t cbp6, int off) { /* loc_116290 MD:1864 */
    uint32_t m = take(o, 3);
    int S = o->S;
    if (m == 2) {
        m = 9;
        plane8(o, o->UV[0], off / 2);
        plane8(o, o->UV[0], off / 2 + S / 2);
    }
    if ((cbp6 >> 4) & 1) intra_coded_fixed(o, o->UV[0], off / 2, m); else predict_intra(o, m, o->UV[0], off / 2);
    if ((cbp6 >> 5) & 1) intra_coded_fixed(o, o->UV[0], off / 2 + S / 2, m); else predict_intra(o, m, o->UV[0], off / 2 + S / 2);
}
static void intra_full_mb(mobi_oracle_t* o, int off) { /* DecIntraFullBlockPMode MD:1759 */
    uint32_t cbp6 = tab(o, MOBI_CBP6_INTRA, 64, uvar(o));
    uint32_t m = take(o, 3);
    int S = o->S;
    if (m == 2) { m = 9; plane16(o, o->Y[0], off); }
    static const int dx[4] = {0, 8, 0, 8}, dy[4] = {0, 0, 8, 8};
    for (int b = 0; b < 4; b++) {
        int bo = off + dy[b] * S + dx[b];
        if ((cbp6 >> b) & 1) intra_coded_fixed(o, o->Y[0], bo, m);
        else predict_intra(o, m, o->Y[0], bo);
    }
    intra_chroma(o, cbp6, off);
}
static void intra_sub_mb(mobi_oracle_t* o, int off) { /* DecIntraSubBlockPMode MD:1789 */
    uint32_t cbp6 = tab(o, MOBI_CBP6_INTRA, 64, uvar(o));
    int S = o->S;
    static const int dx[4] = {0, 8, 0, 8}, dy[4] = {0, 0, 8, 8}, ci[4] = {9, 0xB, 0x19, 0x1B};
    for (int b = 0; b < 4; b++) {
        int bo = off + dy[b] * S + dx[b];
        if ((cbp6 >> b) & 1) intra_coded_ctx(o, ci[b], o->Y[0], bo);
        else intra_uncoded_ctx(o, ci[b], o->Y[0], bo);
    }
    intra_chroma(o, cbp6, off);
}

/* ---- inter macroblocks (MD:400-456, 469-1746, 1818-1833, 2909-2929) --------------------- */
static void copy_block(mobi_oracle_t* o, plane_t src, int dx, int dy, int w, int h, plane_t dst, int off) { /* CopyBlock MD:418 */
    int S = o->S;
    uint8_t row[16];
    for (int i = 0; i < h; i++) {
        long long pos = (long long)off + (long long)((dy >> 1) + i) * S + (dx >> 1);
        switch ((dx & 1) | ((dy & 1) << 1)) {
        case 0:
            if (!src.p || pos < 0 || pos + w > src.n) ABORT(o);
            memcpy(row, src.p + pos, (size_t)w);
            break;
        case 1: for (int j = 0; j < w; j++) row[j] = (uint8_t)((RD(src, pos + j) >> 1) + (RD(src, pos + j + 1) >> 1)); break;
        case 2: for (int j = 0; j < w; j++) row[j] = (uint8_t)((RD(src, pos + j) >> 1) + (RD(src, pos + j + S) >> 1)); break;
        case 3: for (int j = 0; j < w; j++)
                row[j] = (uint8_t)((((RD(src, pos + j) >> 1) + (RD(src, pos + j + 1) >> 1)) >> 1) +
                                   (((RD(src, pos + j + S) >> 1) + (RD(src, pos + j + 1 + S) >> 1)) >> 1));
            break;
        }
        long long d = (long long)off + (long long)i * S;
        if (!dst.p || d < 0 || d + w > dst.n) ABORT(o);
        memcpy(dst.p + d, row, (size_t)w);
    }
}
static void inter_leaf(mobi_oracle_t* o, int io, uint32_t ref, int w, int h, int dx, int dy, int off) { /* loc_1147B0 & clones */
    *iw(o, io) = (uint32_t)dx; *iw(o, io + 1) = (uint32_t)dy;
    int S = o->S;
    copy_block(o, o->Y[ref], dx, dy, w, h, o->Y[0], off);
    copy_block(o, o->UV[ref], dx >> 1, dy >> 1, w >> 1, h >> 1, o->UV[0], off / 2);
    copy_block(o, o->UV[ref], dx >> 1, dy >> 1, w >> 1, h >> 1, o->UV[0], off / 2 + S / 2);
}
static void blk8_inter(mobi_oracle_t* o, plane_t dst, int off) { /* loc_11652C MD:2909 */
    if ((o->win >> 31) & 1) {
        o->win += o->win; o->nb--;
        residual8(o, dst, off);
    } else {
        uint32_t cbp4 = tab(o, MOBI_CBP4_INTER, 16, uvar(o));
        if (cbp4 & 1) residual4(o, dst, off);
        if (cbp4 & 2) residual4(o, dst, off + 4);
        if (cbp4 & 4) residual4(o, dst, off + o->S * 4);
        if (cbp4 & 8) residual4(o, dst, off + o->S * 4 + 4);
    }
}
static void inter_residual(mobi_oracle_t* o, int off) { /* loc_1161A0 MD:1818 */
    uint32_t cbp6 = tab(o, MOBI_CBP6_INTER, 64, uvar(o));
    int S = o->S;
    if (cbp6 & 1) blk8_inter(o, o->Y[0], off);
    if (cbp6 & 2) blk8_inter(o, o->Y[0], off + 8);
    if (cbp6 & 4) blk8_inter(o, o->Y[0], off + S * 8);
    if (cbp6 & 8) blk8_inter(o, o->Y[0], off + S * 8 + 8);
    if (cbp6 & 16) blk8_inter(o, o->UV[0], off / 2);
    if (cbp6 & 32) blk8_inter(o, o->UV[0], off / 2 + S / 2);
}
static void pblock(mobi_oracle_t* o, int lw, int lh, int io, int off) { /* ReadPBlockWxH + SwitchPBlockWxH */
    const mobi_part_code_t* pc = &MOBI_PART_CODE[o->ver == VER_MOFLEX ? 0 : 1][lw][lh];
    uint32_t sym = pc->sym[o->win >> (32 - pc->peek)];
    int n = pc->len[sym];
    drop(o, n); chk(o);
    int w = 2 << lw, h = 2 << lh, top = (lw == 3 && lh == 3);
    if (sym <= 5) {
        int dx = (int)o->I[I_MVP], dy = (int)o->I[I_MVP + 1];
        uint32_t ref = 1;
        if (sym) { int ax = svar(o), ay = svar(o); dx += ax; dy += ay; ref = sym; }
        inter_leaf(o, io, ref, w, h, dx, dy, off);
    } else if (sym == 8 && lh > 0) {
        pblock(o, lw, lh - 1, io, off);
        pblock(o, lw, lh - 1, io, off + o->S * (h / 2));
    } else if (sym == 9 && lw > 0) {
        pblock(o, lw - 1, lh, io, off);
        pblock(o, lw - 1, lh, io, off + w / 2);
    } else if (top && sym == 6) { intra_full_mb(o, off); return;
    } else if (top && sym == 7) { intra_sub_mb(o, off); return;
    } else ABORT(o); /* "error?" throws, e.g. MD:625 */
    if (top) inter_residual(o, off);
}

/* ---- YUV -> BGRA (MD:260-323) ----------------------------------------------------------- */
static void to_bgra(mobi_oracle_t* o, uint8_t* out) {
    const uint8_t* Yp = o->Y[0].p; const uint8_t* C = o->UV[0].p;
    int S = o->S, W = (int)o->W, H = (int)o->H;
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
        float Y2 = Yp[y * S + x];
        int ci = y / 2 * S + x / 2;
        float U = C[ci] - 128.0f, V = C[ci + S / 2] - 128.0f;
        if (x != W - 1 && y != H - 1) {
            switch ((x & 1) | ((y & 1) << 1)) {
            case 1: U += C[ci + 1] - 128.0f; V += C[ci + 1 + S / 2] - 128.0f; U /= 2.0f; V /= 2.0f; break;
            case 2: U += C[ci + S] - 128.0f; V += C[ci + S + S / 2] - 128.0f; U /= 2.0f; V /= 2.0f; break;
            case 3:
                U += C[ci + 1] - 128.0f; V += C[ci + 1 + S / 2] - 128.0f;
                U += C[ci + S] - 128.0f; V += C[ci + S + S / 2] - 128.0f;
                U += C[ci + 1 + S] - 128.0f; V += C[ci + 1 + S + S / 2] - 128.0f;
                U /= 4.0f; V /= 4.0f; break;
            }
        }
        float R, G, B;
        if (o->ver == VER_MOFLEX) {
            R = Y2 + 1.420f * V; G = Y2 - 0.344f * U - 0.714f * V; B = Y2 + 1.772f * U;
            R = (R - 16.0f) * 255.0f / (255.0f - 16.0f);
            G = (G - 16.0f) * 255.0f / (255.0f - 16.0f);
            B = (B - 16.0f) * 255.0f / (255.0f - 16.0f);
        } else if (o->ver == VER_MODS) {
            R = (float)((int)Y2 + (int)U - (int)V); G = (float)((int)Y2 + (int)V); B = (float)((int)Y2 - (int)U - (int)V);
        } else R = G = B = 0;
        if (R < 0) R = 0; if (R > 255) R = 255;
        if (G < 0) G = 0; if (G > 255) G = 255;
        if (B < 0) B = 0; if (B > 255) B = 255;
        uint8_t* p = out + ((size_t)y * W + x) * 4;
        p[0] = (uint8_t)(int)B; p[1] = (uint8_t)(int)G; p[2] = (uint8_t)(int)R; p[3] = 255;
    }
}

/* ---- frame driver (DecodeVXS2 MD:97-329) ------------------------------------------------ */
static int med3(int a, int b, int c) { /* sort network MD:171-188 leaves the median in the middle */
    int t;
    if (a > b) { t = a; a = b; b = t; }
    if (b > c) { t = b; b = c; c = t; }
    if (a > b) { t = a; a = b; b = t; }
    return b;
}
static void decode_body(mobi_oracle_t* o) {
    int S = o->S, W = (int)o->W, H = (int)o->H;
    o->nb = 0;
    o->win = rd16(o, o->off) << 16;
    o->off += 2;
    uint32_t intra = o->win >> 31;
    o->win += o->win;
    if (!intra) {
        if (--o->nb < 0) fill(o);
        if (o->ver == VER_MOFLEX) {
            uint32_t q = o->quant;
            int dq = svar(o);
            if (q == 0) setup_quant(o, q);
            else if (dq != 0) setup_quant(o, (uint32_t)(q + dq));
        } else if (o->ver == VER_MODS) {
            int dq = svar(o);
            if (dq != 0) setup_quant(o, (uint32_t)(o->quant + dq));
        }
        o->I[I_VLC] = 0;
        int io = I_MVC, w = W + 0x20;
        do { *iw(o, io) = 0; *iw(o, io + 1) = 0; io += 2; w -= 16; } while (w > 0);
        int off = 0, h = H;
        do {
            w = W; io = I_MVC;
            do {
                int v[6];
                for (int k = 0; k < 6; k++) v[k] = (int)*iw(o, io + k);
                io += 2;
                o->I[I_MVP] = (uint32_t)med3(v[0], v[2], v[4]);
                o->I[I_MVP + 1] = (uint32_t)med3(v[1], v[3], v[5]);
                *iw(o, io) = 0; *iw(o, io + 1) = 0;
                pblock(o, 3, 3, io, off);
                off += 16; w -= 16;
            } while (w > 0);
            off += S * 16 - W; h -= 16;
        } while (h > 0);
    } else {
        o->yuvfmt = o->win >> 31; o->win += o->win;
        o->I[I_VLC] = o->win >> 31; o->win += o->win;
        o->nb -= 3; chk(o);
        uint32_t q = o->win >> 26;
        drop(o, 6); chk(o);
        if (o->quant != q) setup_quant(o, q);
        int off = 0, h = H;
        do {
            int w = W;
            do {
                uint32_t sub = o->win >> 31;
                o->win += o->win; o->nb--; chk(o);
                if (sub) intra_sub_mb(o, off); else intra_full_mb(o, off);
                off += 16; w -= 16;
            } while (w > 0);
            off += S * 16 - W; h -= 16;
        } while (h > 0);
    }
}

int mobi_oracle_decode(mobi_oracle_t* o, const uint8_t* data, int len, int* offset_inout, uint8_t* bgra) {
    if (o->ver != VER_MODS && o->ver != VER_MOFLEX) return 0; /* VXS1 is a stub in the reference (MD:63-95) */
    free(o->Y[5].p); free(o->UV[5].p);
    for (int i = 5; i > 0; i--) { o->Y[i] = o->Y[i - 1]; o->UV[i] = o->UV[i - 1]; }
    o->Y[0].n = (int)(o->S * o->H); o->UV[0].n = (int)(o->S * o->H / 2);
    o->Y[0].p = (uint8_t*)calloc((size_t)o->Y[0].n + 1, 1);
    o->UV[0].p = (uint8_t*)calloc((size_t)o->UV[0].n + 1, 1);
    o->d = data; o->dlen = len; o->off = *offset_inout;
    volatile int ok = 0;
    if (setjmp(o->jb) == 0) { decode_body(o); ok = 1; }
    *offset_inout = o->off;
    if (ok && bgra) to_bgra(o, bgra);
    return ok;
}

mobi_oracle_t* mobi_oracle_create(uint32_t width, uint32_t height, int version) {
    mobi_oracle_t* o = (mobi_oracle_t*)calloc(1, sizeof(*o));
    o->W = width; o->H = height; o->ver = version;
    o->S = width <= 256 ? 256 : width <= 512 ? 512 : 1024; /* MD:50-52 */
    const mobi_vlc_run_t* runs[2] = {MOBI_VLC0_RUNS, MOBI_VLC1_RUNS};
    int nruns[2] = {(int)(sizeof(MOBI_VLC0_RUNS) / sizeof(MOBI_VLC0_RUNS[0])), (int)(sizeof(MOBI_VLC1_RUNS) / sizeof(MOBI_VLC1_RUNS[0]))};
    for (int t = 0; t < 2; t++) {
        int k = 0;
        for (int r = 0; r < nruns[t]; r++) for (int j = 0; j < runs[t][r].span; j++) o->vlc[t][k++] = runs[t][r].word;
    }
    return o;
}
void mobi_oracle_destroy(mobi_oracle_t* o) {
    if (!o) return;
    for (int i = 0; i < 6; i++) { free(o->Y[i].p); free(o->UV[i].p); }
    free(o);
}
const uint8_t* mobi_oracle_y(const mobi_oracle_t* o) { return o->Y[0].p; }
const uint8_t* mobi_oracle_uv(const mobi_oracle_t* o) { return o->UV[0].p; }
int mobi_oracle_stride(const mobi_oracle_t* o) { return o->S; }
uint32_t mobi_oracle_quantizer(const mobi_oracle_t* o) { return o->quant; }
uint32_t mobi_oracle_yuvformat(const mobi_oracle_t* o) { return o->yuvfmt; }

/* ---- primitive-level hooks for differential unit tests (mirror oracle/ref_capi.cpp) ------ */
void mobi_oracle_set_planes(mobi_oracle_t* o, const uint8_t* y, const uint8_t* uv) {
    free(o->Y[0].p); free(o->UV[0].p);
    o->Y[0].n = (int)(o->S * o->H); o->UV[0].n = (int)(o->S * o->H / 2);
    o->Y[0].p = (uint8_t*)malloc((size_t)o->Y[0].n + 1); o->UV[0].p = (uint8_t*)malloc((size_t)o->UV[0].n + 1);
    memcpy(o->Y[0].p, y, (size_t)o->Y[0].n); memcpy(o->UV[0].p, uv, (size_t)o->UV[0].n);
}
static void prime(mobi_oracle_t* o, uint32_t window) { o->d = (const uint8_t*)""; o->dlen = 0; o->off = 0; o->win = window; o->nb = 16; }
int mobi_oracle_predict_intra(mobi_oracle_t* o, uint32_t mode, int plane, int offset, uint32_t window) {
    prime(o, window);
    if (setjmp(o->jb)) return 0;
    predict_intra(o, mode, plane ? o->UV[0] : o->Y[0], offset);
    return 1;
}
int mobi_oracle_plane16(mobi_oracle_t* o, int offset, uint32_t window) {
    prime(o, window);
    if (setjmp(o->jb)) return 0;
    plane16(o, o->Y[0], offset);
    return 1;
}
int mobi_oracle_copy_block(mobi_oracle_t* o, int plane, const uint8_t* src, int dx, int dy, uint32_t w, uint32_t h, int offset) {
    plane_t dst = plane ? o->UV[0] : o->Y[0];
    plane_t s = {(uint8_t*)src, dst.n};
    if (setjmp(o->jb)) return 0;
    copy_block(o, s, dx, dy, (int)w, (int)h, dst, offset);
    return 1;
}
int mobi_oracle_idct(mobi_oracle_t* o, int plane, int n, const int32_t* coef, int endpos, int offset) {
    (void)endpos;
    for (int i = 0; i < n * n; i++) o->I[I_COEF + i] = (uint32_t)coef[i];
    if (setjmp(o->jb)) return 0;
    if (n == 8) idct8_add(o, plane ? o->UV[0] : o->Y[0], offset); else idct4_add(o, plane ? o->UV[0] : o->Y[0], offset);
    return 1;
}
