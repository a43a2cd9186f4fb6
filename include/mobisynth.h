/* libmobisynth.so -- seeded, syntax-directed Mobiclip bitstream synthesiser (host only, no CUDA).
 *
 * The reference ships no sample media and its encoder (MobiEncoder.cs) is a work in progress that
 * never emits half-pel vectors, sub-block intra modes or table-1 residuals, so
 * tests and bench.py draw their streams from here instead.  Every frame it writes is *in contract*
 * for MobiclipDecoder.DecodeFrame (MobiclipDecoder.cs:56):
 *   - every motion vector keeps all luma and chroma reads of CopyBlock (MD:418) inside the flat
 *     plane arrays, and references only pictures that exist in the 6-deep ring (MD:19, 102-106);
 *   - no intra block in the top pixel row uses a predictor that reads above the picture (MD:1883);
 *   - every residual keeps pixel+delta inside the 384-entry clip table (MobiConst.cs:587);
 *   - every Elias-gamma code fits the 16 valid bits the reader guarantees (MD:2970-2996).
 * It does not reconstruct pixels; it is not an encoder.
 */
#ifndef MOBISYNTH_H
#define MOBISYNTH_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct mobi_synth mobi_synth_t;

typedef struct mobi_synth_params {
    uint32_t width, height;   /* multiples of 16 */
    int32_t version;          /* 1 ModsDS, 2 Moflex3DS */
    uint64_t seed;
    int32_t gop;              /* an I-frame every `gop` frames (frame 0 is always I); <=0: only frame 0 */
    int32_t quant;            /* initial quantiser, 12..52 */
    float p_dquant;           /* P-frames: probability of a +-1..2 quantiser step */
    float p_split;            /* probability that a partition splits (per level) */
    float p_intra_mb;         /* P-frames: probability of an intra macroblock */
    float p_sub_mb;           /* intra MBs: probability of per-block (context-coded) modes */
    float p_cbp;              /* probability that an 8x8 block carries a residual */
    float p_blk8;             /* coded 8x8 block: probability of one 8x8 transform (else 4x4 split) */
    float mean_coefs;         /* mean number of coefficients per coded transform block */
    float p_escape;           /* probability of forcing an escape form for a coefficient */
    int32_t mv_range;         /* |mv delta| <= mv_range half-pels (<= 32) */
    float p_ref1;             /* probability of ref 1 when more references are available */
    float p_zero_mv;          /* probability of partition code 0 (predicted vector, ref 1) when legal */
    float p_oob_mv;           /* probability of letting a vector read outside the visible picture
                                 (still inside the flat arrays: exercises stride padding / row wrap) */
    int32_t inter_only;       /* 1: P-frames carry no intra MBs (BASELINE config 2) */
    int32_t gop_phase;        /* I-frames fall where (frame index + gop_phase) % gop == 0 (frame 0 is always I):
                                 staggers the keyframes of streams that advance in lock step */
} mobi_synth_params;

/* Fills *p with the defaults used for BASELINE configs 1/3. */
void mobi_synth_default_params(mobi_synth_params* p, uint32_t width, uint32_t height, int version, uint64_t seed);

mobi_synth_t* mobi_synth_create(const mobi_synth_params* p);
void mobi_synth_destroy(mobi_synth_t* s);

/* Writes the next frame's payload (no container framing, no trailing pad) to out[0..cap).
 * Returns the byte count (always even), or <0 if cap is too small.  *is_key receives 1 for I-frames. */
int mobi_synth_next_frame(mobi_synth_t* s, uint8_t* out, int cap, int* is_key);

/* Statistics of the frame just written (for roofline accounting and test coverage reports). */
typedef struct mobi_synth_stats {
    uint32_t n_mb, n_intra_mb, n_leaves, n_coefs, n_blk8, n_blk4;
    uint32_t shape_hist[16];  /* leaves by [log2(w)-1][log2(h)-1] */
    uint32_t phase_hist[4];   /* leaves by half-pel phase */
    uint32_t mode_hist[20];   /* intra predictor modes used */
    uint32_t ref_hist[6];
} mobi_synth_stats;
void mobi_synth_last_stats(const mobi_synth_t* s, mobi_synth_stats* st);

#ifdef __cplusplus
}
#endif
#endif
