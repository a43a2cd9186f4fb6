/* TEST INFRASTRUCTURE -- CPU oracle for the Mobiclip frame path.  NOT part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this.  libmobicuda.so never links or calls it.
 *
 * Plain-C restatement of LibMobiclip/Codec/Mobiclip/MobiclipDecoder.cs (reference @ c88b67d3):
 * one call = one DecodeFrame() (MD:56): bitstream parse + reconstruction + YUV->RGB in a single
 * pass, on the reference's flat Stride-addressed planes and 6-deep ring.
 *
 * Parity status: the reference ships no golden vectors and cannot run here (C#, no .NET).  This
 * oracle is pinned instead against oracle/_ref (the reference's own source, transliterated
 * syntactically to C++ and compiled -- see oracle/build_ref.py) on seeded random streams, and the
 * resulting plane hashes are committed under tests/golden/.  See DESIGN.md "Oracle".
 */
#ifndef MOBI_ORACLE_H
#define MOBI_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct mobi_oracle mobi_oracle_t;

/* version: 0 VxDS (unsupported, as in the reference MD:63-95), 1 ModsDS, 2 Moflex3DS (MD:32-37) */
mobi_oracle_t* mobi_oracle_create(uint32_t width, uint32_t height, int version);
void mobi_oracle_destroy(mobi_oracle_t* o);

/* Data/Offset/DecodeFrame (MD:15-16, 56).  Returns 1 if the reference would return a Bitmap,
 * 0 if it would return null (any exception -> MD:325).  *offset_inout is updated like Offset.
 * bgra: optional W*H*4 destination, memory order B,G,R,A (MD:320). */
int mobi_oracle_decode(mobi_oracle_t* o, const uint8_t* data, int len, int* offset_inout, uint8_t* bgra);

/* Y[0] (Stride*H bytes) and UV[0] (Stride*H/2 bytes), MD:19-20, 107-108 */
const uint8_t* mobi_oracle_y(const mobi_oracle_t* o);
const uint8_t* mobi_oracle_uv(const mobi_oracle_t* o);
int mobi_oracle_stride(const mobi_oracle_t* o);
uint32_t mobi_oracle_quantizer(const mobi_oracle_t* o);
uint32_t mobi_oracle_yuvformat(const mobi_oracle_t* o);

#ifdef __cplusplus
}
#endif
#endif
