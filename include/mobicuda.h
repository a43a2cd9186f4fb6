/* libmobicuda.so -- B200-native (sm_100a) Mobiclip frame reconstruction behind a C ABI.
 *
 * Drop-in for the per-frame Data -> YUV(+RGB) call of the reference decoder class
 *   LibMobiclip.Codec.Mobiclip.MobiclipDecoder   (LibMobiclip/Codec/Mobiclip/MobiclipDecoder.cs, "MD")
 * The serial entropy parse runs on the host inside this library and emits packed per-macroblock
 * arrays; reconstruction (motion compensation, dequant + inverse transforms, intra prediction,
 * add/clip, YUV->BGRA) runs in hand-written CUDA kernels.  All entry points are plain C: opaque
 * handles, raw pointers, sizes; they return 0 on success and a negative mobi_status otherwise; no
 * exception crosses the boundary.  A handle is not thread-safe; distinct handles are independent.
 *
 * The P/Invoke binding a LibMobiclip maintainer would add is in csharp/MobiclipDecoder.cs and described in
 * INTEGRATION.md; each export below cites the reference member it stands in for.
 */
#ifndef MOBICUDA_H
#define MOBICUDA_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MOBICUDA_ABI_VERSION 1

typedef enum mobi_status {
    MOBI_OK = 0,
    MOBI_ERR_ARG = -1,         /* bad argument / unsupported geometry (W,H must be multiples of 16, W <= 1024) */
    MOBI_ERR_UNSUPPORTED = -2, /* VxDS: DecodeVXS1 is a stub in the reference too (MD:63-95) */
    MOBI_ERR_BITSTREAM = -3,   /* the reference would throw while parsing (illegal partition code MD:625, table index, run overflow) */
    MOBI_ERR_REFERENCE = -4,   /* P-frame names a picture that is not in the ring (Y[k]==null, MD:413) */
    MOBI_ERR_RANGE = -5,       /* a motion vector or intra predictor reads outside the plane arrays (C# IndexOutOfRange) */
    MOBI_ERR_CUDA = -6,        /* CUDA runtime failure; see mobi_last_error */
    MOBI_ERR_NOMEM = -7,
    MOBI_ERR_STATE = -8        /* call sequence error (e.g. read before any frame was decoded) */
} mobi_status;

/* MobiclipDecoder.MobiclipVersion (MD:32-37) */
typedef enum mobi_version { MOBI_VXDS = 0, MOBI_MODSDS = 1, MOBI_MOFLEX3DS = 2 } mobi_version;

/* ------------------------------------------------------------------------------------------------
 * Packed per-frame arrays: the host-parse -> device-reconstruct contract.
 * Everything is little-endian, every array 16-byte aligned.  A parsed frame is self-contained apart
 * from the pictures it references.
 * ---------------------------------------------------------------------------------------------- */

/* One per frame. qtab mirrors Internal[10..89] (MD:3897-3912): (scale << 8) | matrix index, first the
 * 64 8x8 scan positions, then the 16 4x4 ones; the device computes coef[word & 0xFF] = (word >> 8) * level. */
typedef struct mobi_frame_hdr {
    uint32_t flags;        /* bit 0: I-frame */
    uint32_t n_mb, n_parts, n_ops, n_coefs, n_intra;
    uint32_t quantizer, yuv_format;
    uint32_t bytes_consumed; /* value of Offset after the call, relative to the Offset passed in */
    uint32_t max_ref;      /* highest ring index referenced (0 for I-frames) */
    uint32_t n_inter_coefs; /* coefficients that belong to inter macroblocks (accounting only) */
    uint32_t reserved[5];
    uint32_t qtab[80];
} mobi_frame_hdr;            /* 384 bytes */

/* One per macroblock, raster order. */
typedef struct mobi_mb {
    uint32_t info;         /* bits 0-1 kind: 0 inter, 1 intra; bits 2-8 n_sub (partitions <= 64, or intra ops <= 27);
                              bits 9-17 n_coefs (<= 384); bits 18-23 mask of 8x8 blocks holding >= 1 coefficient;
                              bits 24-27 (intra) neighbouring macroblocks whose pixels the predictors read:
                              1 left (m-1), 2 top-left (m-mbw-1), 4 top (m-mbw), 8 top-right (m-mbw+1);
                              bits 24-27 and 29-30 (inter) which of the coded blocks 0-3 / 4-5 are transformed as one 8x8
                              (the rest as 4x4 units): the same fact as bit 7 of their coefficient records' blk;
                              bit 28 (inter) the single partition is also stored inline in intra_rank */
    uint32_t first_sub;    /* index of the first mobi_part (inter) or mobi_op (intra) */
    uint32_t first_coef;   /* index of the first mobi_coef */
    uint32_t intra_rank;   /* intra MBs: position in the frame's intra list.  Inter MBs with info bit 28 set (exactly one
                              partition, vector within 14 bits): that partition inline, mvx & 0x3FFF | (mvy & 0x3FFF) << 14
                              | ref << 28 (the mobi_part record is emitted as well); otherwise 0 */
} mobi_mb;                   /* 16 bytes */

/* Motion partition leaf (MD:400-416 and clones): luma rect (x,y,w,h) inside the MB, ring index, half-pel vector. */
typedef struct mobi_part {
    uint8_t xy;            /* (x/2) | (y/2) << 4 */
    uint8_t shape;         /* log2(w)-1 | (log2(h)-1) << 2 | ref << 4   (ref 1..5) */
    int16_t mvx, mvy;      /* luma half-pel units; chroma uses (mv >> 1) then half-pel again (MD:414-415) */
    uint16_t pad;
} mobi_part;                 /* 8 bytes */

/* Residual coefficient: quantised level, not yet scaled (MD:3424-3429). */
typedef struct mobi_coef {
    int16_t level;
    uint8_t pos;           /* bits 0-5 scan position; bits 6-7 4x4 sub-block inside a split 8x8 */
    uint8_t blk;           /* bits 0-2 8x8 block (0-3 luma raster, 4 U, 5 V); bits 3-4 macroblock index & 3 (the inter kernel
                              pools the coefficients of four consecutive macroblocks and files each record by this tag);
                              bit 6 last record of its transform unit; bit 7: one 8x8 transform (else 4x4) */
} mobi_coef;                 /* 4 bytes */

/* Intra operation, executed in stream order by one warp (MD:1759-1880, 2776-2902):
 *   bits 0-4  predictor 0..8 (8x8) / 10..18 (4x4) / 9,19 none / 20 = 16x16 plane (MD:3017)
 *   bit  5    a residual follows for this block (its coefficients carry the same blk/sub tags)
 *   bits 6-7  plane 0 Y, 1 U, 2 V
 *   bits 8-9  x/4, bits 10-11 y/4 inside the MB's 16x16 (luma) or 8x8 (chroma) area
 *   bits 16-31 signed plane-predictor delta (modes 2, 12, 20; MD:3019, 3170, 3255)
 * A macroblock's ops come in decode order: all luma ops before the chroma ops (MD:1759-1880); at most 27 per macroblock. */
typedef uint32_t mobi_op;

typedef struct mobi_packed_frame {
    const mobi_frame_hdr* hdr;
    const mobi_mb* mbs;
    const mobi_part* parts;
    const mobi_op* ops;
    const mobi_coef* coefs;
    const uint32_t* intra_list; /* raster indices of intra MBs, ascending */
} mobi_packed_frame;

/* ------------------------------------------------------------------------------------------------
 * Host-only parse (no GPU needed): the entropy half of DecodeVXS2 (MD:97-259).
 * ---------------------------------------------------------------------------------------------- */
typedef struct mobi_parser mobi_parser_t;

int mobi_parser_create(uint32_t width, uint32_t height, int version, mobi_parser_t** out);
void mobi_parser_destroy(mobi_parser_t* p);
/* Parses one frame starting at data[*offset_inout]; on success *offset_inout advances exactly as
 * MobiclipDecoder.Offset does (MD:110-111, 2990-2992) and *out views storage owned by the parser,
 * valid until the next call.  On error the parser state is rolled back to before the call. */
int mobi_parser_parse(mobi_parser_t* p, const uint8_t* data, int len, int* offset_inout, mobi_packed_frame* out);
const char* mobi_parser_last_error(const mobi_parser_t* p);

/* ------------------------------------------------------------------------------------------------
 * Single-stream decoder: mirrors the MobiclipDecoder object (MD:13-61).
 * ---------------------------------------------------------------------------------------------- */
typedef struct mobi_decoder mobi_t;

/* new MobiclipDecoder(Width, Height, Version) (MD:41-54); allocates the 6-picture ring in HBM with
 * the reference's Stride rule (MD:50-52). device = CUDA ordinal. */
int mobi_create(uint32_t width, uint32_t height, int version, int device, mobi_t** out);
void mobi_destroy(mobi_t* d);

/* d.Data = data; d.Offset = *offset_inout; d.DecodeFrame() (MD:15-16, 56).  Host buffers.
 * Returns MOBI_OK where the reference returns a Bitmap; an error where it returns null (MD:325-328),
 * in which case the ring is left untouched (the reference leaves a partially written picture). */
int mobi_decode_frame(mobi_t* d, const uint8_t* data, int len, int* offset_inout);

/* Pre-parsed path (BASELINE config 2): reconstruct a frame from packed arrays in host memory. */
int mobi_submit_packed(mobi_t* d, const mobi_packed_frame* f);

/* The checks mobi_submit_packed applies to caller-supplied arrays before anything reaches the GPU, as a host-only call
 * (no device needed): every index the kernels would use is range-checked (descriptor fields, op fields, partition
 * geometry and tiling, vectors against the plane arrays, references against `pictures` = pictures in the ring, coefficient
 * tags and the contiguity of the per-macroblock coefficient ranges).  err (optional) receives a diagnostic. */
int mobi_packed_validate(uint32_t width, uint32_t height, int version, const mobi_packed_frame* f, int pictures, char* err, size_t err_len);

/* Y[0] / UV[0] (MD:19-20).  strided = byte-identical to the reference arrays (Stride*H, Stride*H/2);
 * tight = cropped planar I420.  Host destinations; each call synchronises the decoder's stream. */
int mobi_read_planes_strided(mobi_t* d, uint8_t* y, uint8_t* uv);
int mobi_read_yuv(mobi_t* d, uint8_t* y, uint8_t* u, uint8_t* v);
/* The returned Bitmap (MD:260-323): W*H 32bpp, memory order B,G,R,A, dst_stride bytes per row. */
int mobi_read_bgra(mobi_t* d, uint8_t* dst, int dst_stride);
/* Public fields Quantizer / YuvFormat / Stride (MD:26-30). */
int mobi_get_state(const mobi_t* d, uint32_t* quantizer, uint32_t* yuv_format, int* stride);
const char* mobi_last_error(const mobi_t* d);

/* ------------------------------------------------------------------------------------------------
 * Lock-step batch: N independent streams of equal geometry advance one frame per step on one GPU.
 * Parsing fans out over host threads; upload, reconstruction and read-back are one arena copy, one
 * set of kernel launches and one copy back.  This is the throughput path.
 * ---------------------------------------------------------------------------------------------- */
typedef struct mobi_batch mobi_batch_t;

int mobi_batch_create(uint32_t width, uint32_t height, int version, int device, int n_streams, int n_threads, mobi_batch_t** out);
void mobi_batch_destroy(mobi_batch_t* b);
/* One frame per stream: data[i], len[i], offset_inout[i] as in mobi_decode_frame; status[i] per stream.
 * Streams whose frame fails to parse keep their previous picture (status[i] < 0); the rest advance. */
int mobi_batch_decode(mobi_batch_t* b, const uint8_t* const* data, const int* len, int* offset_inout, int* status);
/* Tight I420 of every stream's newest picture into one host buffer: n_streams * W*H*3/2 bytes. */
int mobi_batch_read_yuv(mobi_batch_t* b, uint8_t* dst);
int mobi_batch_read_planes_strided(mobi_batch_t* b, int stream, uint8_t* y, uint8_t* uv);
int mobi_batch_read_bgra(mobi_batch_t* b, int stream, uint8_t* dst, int dst_stride);
/* The Bitmap of every stream's newest picture (MD:260-323), n_streams * W*H*4 bytes, tightly packed.
 * dst == NULL converts on the device only (no copy back). */
int mobi_batch_read_bgra_all(mobi_batch_t* b, uint8_t* dst);
const char* mobi_batch_last_error(const mobi_batch_t* b);

/* --- pipelined decode: the host parses step k+1 while the GPU reconstructs, converts and copies back step k.
 * mobi_batch_submit = mobi_batch_decode + conversion of every stream's new picture (format MOBI_OUT_I420: tight
 * planar Y,U,V, W*H*3/2 bytes per stream; MOBI_OUT_BGRA: the Bitmap of MD:260-323, W*H*4 bytes per stream) + an
 * asynchronous copy into pinned host memory.  It returns without waiting for the GPU.  At most two results may be
 * outstanding.  mobi_batch_fetch waits for the OLDEST outstanding result: dst (optional) receives a copy,
 * *view (optional) the library's pinned buffer itself (valid until the second mobi_batch_submit from now),
 * *bytes (optional) its size. --- */
#define MOBI_OUT_I420 1
#define MOBI_OUT_BGRA 2
int mobi_batch_submit(mobi_batch_t* b, const uint8_t* const* data, const int* len, int* offset_inout, int* status, int format);
int mobi_batch_fetch(mobi_batch_t* b, uint8_t* dst, const uint8_t** view, size_t* bytes);

/* --- pre-parsed, device-resident replay (bench "value" leg / ncu captures): frames are parsed and
 * uploaded once with mobi_batch_stage(), then mobi_batch_replay() runs reconstruction only. --- */
/* Parse + upload step `step` of every stream (data/len/offset as above) into resident staging. */
int mobi_batch_stage(mobi_batch_t* b, const uint8_t* const* data, const int* len, int* offset_inout);
/* Reconstruct staged steps [first, first+count) in order; no host<->device traffic. Asynchronous. */
int mobi_batch_replay(mobi_batch_t* b, int first, int count);
/* The same, and every step's new pictures are converted on the device as well (format MOBI_OUT_BGRA: the Bitmap of MD:260-323,
 * into the batch's device-side output buffer; 0: no conversion): the whole north_star path -- reconstruction and YUV->RGB --
 * without host<->device traffic. */
int mobi_batch_replay_convert(mobi_batch_t* b, int first, int count, int format);
int mobi_batch_staged_steps(const mobi_batch_t* b);
void mobi_batch_clear_staged(mobi_batch_t* b);
/* Restore every stream's ring to "no picture decoded" (replay from an I-frame again). */
int mobi_batch_reset(mobi_batch_t* b);
/* Same, and every stream's parser goes back to a freshly constructed decoder (Quantizer 0, no pictures);
 * staged steps must be cleared by the caller. */
int mobi_batch_reset_streams(mobi_batch_t* b);
int mobi_batch_sync(mobi_batch_t* b);
/* The CUDA stream (cudaStream_t) the batch launches on, for event timing by the caller. */
void* mobi_batch_cuda_stream(mobi_batch_t* b);
/* Accounting of the last replay/decode: kernel launches issued and algorithmic bytes per SURVEY 8(d). */
typedef struct mobi_batch_stats {
    uint64_t launches;        /* kernels launched */
    uint64_t frames, mbs, inter_mbs, intra_mbs, parts, coefs, ops;
    uint64_t h2d_bytes, d2h_bytes;
    uint64_t inter_coefs;     /* of coefs: those of inter macroblocks */
} mobi_batch_stats;
int mobi_batch_get_stats(const mobi_batch_t* b, mobi_batch_stats* st);
/* Per-kernel device time, for roofline accounting: while enabled every reconstruction kernel launch is bracketed
 * by CUDA events on the batch's stream.  mobi_batch_get_kernel_times synchronises, returns the summed durations
 * (milliseconds) and launch counts since the last call, and clears them. */
int mobi_batch_set_kernel_timing(mobi_batch_t* b, int enabled);
/* index 0: k_mc (motion compensation of the inter macroblocks); 1: k_intra over the intra macroblocks of P-pictures; 2: k_intra
 * over I-pictures (runs on a second CUDA stream, concurrently with the others); 3: k_res (dequantisation + inverse transforms
 * of the inter macroblocks, added in place; only with MOBI_INTER_KERNEL=split); 4: k_bgra (mobi_batch_replay_convert) */
int mobi_batch_get_kernel_times(mobi_batch_t* b, double ms[5], uint64_t launches[5]);
void mobi_batch_clear_stats(mobi_batch_t* b);
/* Where the calling thread's time went since mobi_batch_clear_stats, in milliseconds of wall time: [0] entropy parse (fanned out
 * over the batch's threads), [1] packing the parsed arrays into the pinned upload arena, [2] enqueueing upload + kernels,
 * [3] waiting in mobi_batch_fetch for a result's copy-back. */
int mobi_batch_get_phase_times(const mobi_batch_t* b, double ms[4]);

int mobicuda_abi_version(void);
/* Device-side exhaustive check of the YUV->RGB kernel's Moflex colour arithmetic (MD:300-305): the kernel evaluates a cheaper
 * expression than the reference's float sequence; for EVERY possible input -- Y 0..255, U and V the 1021 multiples of 1/4 in
 * [-128, 127], 267 M triples -- the resulting B, G, R, A bytes are compared with those of the reference sequence (IEEE division
 * included).  mismatches[0] receives the number of differing triples (0 expected), mismatches[1] one of them (Y << 20 | u << 10 | v). */
int mobicuda_selftest_bgra(int device, unsigned long long* mismatches);

#ifdef __cplusplus
}
#endif
#endif
