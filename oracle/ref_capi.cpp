// TEST INFRASTRUCTURE -- C API over the transliterated reference decoder (see build_ref.py).
// Mirrors the call sequence of the reference's callers (MobiConverter/Program.cs:64-71, 243-245):
//   d = new MobiclipDecoder(W, H, Version); d.Data = bytes; d.Offset = off; bmp = d.DecodeFrame();
#include "ref_shim.h"
#include "gen_IOUtil.h"
using namespace LibMobiclip_Utils;
#include "gen_MobiConst.h"
#include "gen_MobiclipDecoder.h"

using LibMobiclip_Codec_Mobiclip::MobiclipDecoder;

extern "C" {

void* mobiref_create(unsigned w, unsigned h, int version /*0 VxDS, 1 ModsDS, 2 Moflex3DS*/) {
    return new MobiclipDecoder(w, h, (MobiclipDecoder::MobiclipVersion)version);
}

void mobiref_destroy(void* h) { delete (MobiclipDecoder*)h; }

// Returns 1 when DecodeFrame() returned a Bitmap, 0 when it returned null (frame aborted).
// bgra (optional) receives W*H*4 bytes in memory order B,G,R,A.
int mobiref_decode(void* h, const uint8_t* data, int len, int* offset_inout, uint8_t* bgra) {
    MobiclipDecoder* d = (MobiclipDecoder*)h;
    d->Data = Arr<byte>::New(len);
    if (len) std::memcpy(d->Data.raw(), data, (size_t)len);
    d->Offset = *offset_inout;
    Bitmap b = d->DecodeFrame();
    *offset_inout = d->Offset;
    if (b.IsNull()) return 0;
    if (bgra) std::memcpy(bgra, b.px->data(), b.px->size());
    return 1;
}

// Copies the strided planes Y[0] (Stride*H) and UV[0] (Stride*H/2) exactly as the reference holds them.
int mobiref_planes(void* h, uint8_t* y, uint8_t* uv) {
    MobiclipDecoder* d = (MobiclipDecoder*)h;
    if (!d->Y[0].p || !d->UV[0].p) return 0;
    if (y) std::memcpy(y, d->Y[0].raw(), (size_t)d->Y[0].Length);
    if (uv) std::memcpy(uv, d->UV[0].raw(), (size_t)d->UV[0].Length);
    return 1;
}

int mobiref_stride(void* h) { return ((MobiclipDecoder*)h)->Stride; }
unsigned mobiref_quantizer(void* h) { return ((MobiclipDecoder*)h)->Quantizer; }
unsigned mobiref_yuvformat(void* h) { return ((MobiclipDecoder*)h)->YuvFormat; }

}  // extern "C"

// ---- primitive-level hooks for differential unit tests (tests/test_oracle_primitives.py) ----
extern "C" {
// Allocate fresh Y[0]/UV[0] and fill them from caller buffers (Stride*H and Stride*H/2 bytes).
void mobiref_set_planes(void* h, const uint8_t* y, const uint8_t* uv) {
    MobiclipDecoder* d = (MobiclipDecoder*)h;
    d->Y[0] = Arr<byte>::New(d->Stride * d->Height);
    d->UV[0] = Arr<byte>::New(d->Stride * d->Height / 2);
    std::memcpy(d->Y[0].raw(), y, (size_t)d->Y[0].Length);
    std::memcpy(d->UV[0].raw(), uv, (size_t)d->UV[0].Length);
}
// PredictIntra (MD:1883) on Y[0] (plane=0) or UV[0] (plane=1); `window` primes the bit window
// (modes 2/12 read a signed Elias-gamma delta from it).  Returns 0 if the reference threw.
int mobiref_predict_intra(void* h, unsigned mode, int plane, int offset, unsigned window) {
    MobiclipDecoder* d = (MobiclipDecoder*)h;
    d->Data = Arr<byte>::New(0);
    d->Offset = 0;
    int nbits = 16;
    uint r3 = window;
    try { d->PredictIntra(nbits, r3, mode, plane ? d->UV[0] : d->Y[0], offset); } catch (...) { return 0; }
    return 1;
}
// 16x16 plane predictor sub_1167BC (MD:3017)
int mobiref_plane16(void* h, int offset, unsigned window) {
    MobiclipDecoder* d = (MobiclipDecoder*)h;
    d->Data = Arr<byte>::New(0);
    d->Offset = 0;
    int nbits = 16;
    uint r3 = window;
    try { d->sub_1167BC(d->Y[0], offset, nbits, r3); } catch (...) { return 0; }
    return 1;
}
// CopyBlock (MD:418): Src = caller plane copy of Y[0]/UV[0] size, Dst = Y[0] / UV[0]
int mobiref_copy_block(void* h, int plane, const uint8_t* src, int dx, int dy, unsigned w, unsigned hgt, int offset) {
    MobiclipDecoder* d = (MobiclipDecoder*)h;
    Arr<byte>& dst = plane ? d->UV[0] : d->Y[0];
    Arr<byte> s = Arr<byte>::New(dst.Length);
    std::memcpy(s.raw(), src, (size_t)dst.Length);
    try { d->CopyBlock(s, dx, dy, w, hgt, dst, offset); } catch (...) { return 0; }
    return 1;
}
// The size-dispatched inverse transforms exactly as loc_116540 / sub_1166E8 pick them (MD:2939-2942,
// 2954-2955): coef = 64 or 16 s32 in Internal[90..] order, endpos = scan position after the last coefficient.
int mobiref_idct(void* h, int plane, int n, const int32_t* coef, int endpos, int offset) {
    MobiclipDecoder* d = (MobiclipDecoder*)h;
    Arr<byte>& dst = plane ? d->UV[0] : d->Y[0];
    for (int i = 0; i < n * n; i++) d->Internal[90 + i] = (uint)coef[i];
    try {
        if (n == 8) {
            uint r12 = 10 + endpos;
            if (r12 <= 11) d->IDCT1Px8(dst, offset);
            else if (r12 <= 13) d->IDCT3Px8(dst, offset);
            else if (r12 <= 20) d->IDCT16Px8(dst, offset);
            else d->IDCT64Px8(dst, offset);
        } else {
            uint r12 = 74 + endpos;
            if (r12 <= 75) d->IDCT1Px4(dst, offset);
            else d->IDCT16Px4(dst, offset);
        }
    } catch (...) { return 0; }
    return 1;
}
}  // extern "C"

// ---- the reference's SECOND copies of the primitives: FrameUtil.GetPBlock (twin of CopyBlock),
// MobiEncoder.IDCT64 / IDCT16 (+ the forward transforms), MacroBlock.GetCompvals8x8 / 4x4 and the three plane
// predictors.  Compiled from the reference's files like the decoder; tests/test_second_copies.py checks decoder copy
// == encoder copy == oracle on random inputs. ----
using namespace LibMobiclip_Codec_Mobiclip;
#include "gen_SecondCopies.h"
using LibMobiclip_Codec_Mobiclip_Encoder::EncTransforms;
using LibMobiclip_Codec_Mobiclip_Encoder::EncPredictors;

namespace {
template <class T> Arr<T> arr_from(const T* p, int n) { Arr<T> a = Arr<T>::New(n); if (n) std::memcpy(a.raw(), p, (size_t)n * sizeof(T)); return a; }
}

extern "C" {
// FrameUtil.GetPBlock(Src, Dx, Dy, Width, Height, Offset, Stride) -> Width*Height bytes.  Returns 0 if it threw.
int mobiref2_pblock(const uint8_t* src, int len, int dx, int dy, unsigned w, unsigned h, int offset, int stride, uint8_t* out) {
    try {
        Arr<byte> r = FrameUtil::GetPBlock(arr_from<byte>(src, len), dx, dy, w, h, offset, stride);
        std::memcpy(out, r.raw(), (size_t)r.Length);
        return 1;
    } catch (...) { return 0; }
}
// MobiEncoder.IDCT64 / IDCT16 (DCT[n*n], PPixels[n*n]) -> n*n reconstructed pixels.
int mobiref2_idct(int n, const int32_t* dct, const uint8_t* pred, uint8_t* out) {
    try {
        Arr<int> d = arr_from<int>(dct, n * n);
        Arr<byte> p = arr_from<byte>(pred, n * n);
        Arr<byte> r = n == 8 ? EncTransforms::IDCT64(d, p) : EncTransforms::IDCT16(d, p);
        std::memcpy(out, r.raw(), (size_t)r.Length);
        return 1;
    } catch (...) { return 0; }
}
// MobiEncoder.DCT64 / DCT16 (InPixels[n*n], residual as int) -> n*n coefficients.
int mobiref2_fdct(int n, const int32_t* px, int32_t* out) {
    try {
        Arr<int> r = n == 8 ? EncTransforms::DCT64(arr_from<int>(px, n * n)) : EncTransforms::DCT16(arr_from<int>(px, n * n));
        std::memcpy(out, r.raw(), (size_t)r.Length * sizeof(int));
        return 1;
    } catch (...) { return 0; }
}
// MacroBlock.GetCompvals8x8 / GetCompvals4x4 (BlockType, Data, X, Y, Stride, Offset) -> n*n predicted pixels.
int mobiref2_compvals(int n, int mode, const uint8_t* data, int len, int x, int y, int stride, int offset, uint8_t* out) {
    try {
        Arr<byte> r = n == 8 ? EncPredictors::GetCompvals8x8(mode, arr_from<byte>(data, len), x, y, stride, offset)
                             : EncPredictors::GetCompvals4x4(mode, arr_from<byte>(data, len), x, y, stride, offset);
        if (!r.p) return 0;
        std::memcpy(out, r.raw(), (size_t)r.Length);
        return r.Length;
    } catch (...) { return 0; }
}
// MacroBlock.PredictIntraPlane16x16 / 8x8 / 4x4 (Data, Offset, Stride, Param) -> n*n predicted pixels.
int mobiref2_plane(int n, const uint8_t* data, int len, int offset, int stride, int param, uint8_t* out) {
    try {
        Arr<byte> d = arr_from<byte>(data, len);
        Arr<byte> r = n == 16 ? EncPredictors::PredictIntraPlane16x16(d, offset, stride, param)
                    : n == 8 ? EncPredictors::PredictIntraPlane8x8(d, offset, stride, param) : EncPredictors::PredictIntraPlane4x4(d, offset, stride, param);
        std::memcpy(out, r.raw(), (size_t)r.Length);
        return r.Length;
    } catch (...) { return 0; }
}
}  // extern "C"

// ---- the reference's bit writer and coefficient entropy coder (BitWriter.cs, MobiEncoder.EncodeDCT ME:675-765):
// an independent writer for bitstreams the parsers are tested on. ----
#include "gen_EntropyWriter.h"
using LibMobiclip_Codec_Mobiclip::BitWriter;
using LibMobiclip_Codec_Mobiclip_Encoder::EncEntropy;

extern "C" {
void* mobiref2_bw_create() { return new BitWriter(); }
void mobiref2_bw_destroy(void* h) { delete (BitWriter*)h; }
void mobiref2_bw_bits(void* h, unsigned value, int nbits) { ((BitWriter*)h)->WriteBits(value, nbits); }
void mobiref2_bw_uvar(void* h, unsigned value) { ((BitWriter*)h)->WriteVarIntUnsigned(value); }
void mobiref2_bw_svar(void* h, int value) { ((BitWriter*)h)->WriteVarIntSigned(value); }
// EncodeDCT(DCT, Table, b): DCT = quantised levels in SCAN order (n = 64 or 16).  Returns 0 if the reference threw.
int mobiref2_bw_dct(void* h, const int32_t* dct, int n, int table) {
    try { EncEntropy::EncodeDCT(arr_from<int>(dct, n), table, *(BitWriter*)h); return 1; } catch (...) { return 0; }
}
// ToArray(): flushes and returns the byte count (the bytes if they fit into cap).
int mobiref2_bw_bytes(void* h, uint8_t* out, int cap) {
    Arr<byte> a = ((BitWriter*)h)->ToArray();
    if (a.Length <= cap && a.Length) std::memcpy(out, a.raw(), (size_t)a.Length);
    return a.Length;
}
}  // extern "C"


// ---- containers (tests/test_containers_vs_ref.py): the reference's ModsDemuxer, MoLiveDemux and MoflexMuxer ----------------
#include "gen_Containers.h"
using LibMobiclip_Containers_Mods::ModsDemuxer;
using namespace LibMobiclip_Containers_Moflex;

extern "C" {

struct RefMods { CsStream s; ModsDemuxer* d = nullptr; Arr<byte> last; };
// new ModsDemuxer(stream).  Returns null when the reference throws (tables outside the file ...).
void* mobiref_mods_open(const uint8_t* data, size_t len) {
    RefMods* r = new RefMods();
    r->s = CsStream(data, len);
    try { r->d = new ModsDemuxer(&r->s); } catch (...) { delete r; return nullptr; }
    return r;
}
void mobiref_mods_close(void* h) { delete (RefMods*)h; }
// the 0x30-byte header as 14 u32 (magic as its four bytes, little endian)
void mobiref_mods_header(void* h, uint32_t* out) {
    ModsDemuxer::ModsHeader* H = ((RefMods*)h)->d->Header;
    uint32_t magic = 0;
    for (int i = 0; i < 4 && i < (int)H->ModsString.size(); i++) magic |= (uint32_t)(uint8_t)H->ModsString[(size_t)i] << (8 * i);
    const uint32_t v[14] = {magic, H->TagId, H->TagIdSizeDword, H->FrameCount, H->Width, H->Height, H->Fps, H->AudioCodec, H->NbChannel,
                            H->Frequency, H->BiggestFrame, H->AudioOffset, H->KeyframeIndexOffset, H->KeyframeCount};
    std::memcpy(out, v, sizeof v);
}
int mobiref_mods_keyframe(void* h, uint32_t i, uint32_t* frame_number, uint32_t* data_offset) {
    ModsDemuxer* d = ((RefMods*)h)->d;
    if ((long long)i >= d->KeyFrames.Length) return -1;
    *frame_number = d->KeyFrames[i]->FrameNumber; *data_offset = d->KeyFrames[i]->DataOffset;
    return 0;
}
// ReadFrame(out NrAudioPackets, out IsKeyFrame): 1 and a view of the packet, 0 when it returned null, -1 when it threw
int mobiref_mods_read_frame(void* h, const uint8_t** frame, uint32_t* len, uint32_t* nr_audio, int* is_key) {
    RefMods* r = (RefMods*)h;
    try {
        uint na = 0; bool key = false;
        r->last = r->d->ReadFrame(na, key);
        if (!r->last.p) return 0;
        *frame = r->last.raw(); *len = (uint32_t)r->last.Length; *nr_audio = na; *is_key = key ? 1 : 0;
        return 1;
    } catch (...) { return -1; }
}

struct RefFrame { MoLiveChunk* chunk; Arr<byte> data; };
struct RefMoflex { CsStream s; MoLiveDemux* d = nullptr; std::vector<RefFrame> q; size_t head = 0; };
void* mobiref_moflex_open(const uint8_t* data, size_t len) {
    RefMoflex* r = new RefMoflex();
    r->s = CsStream(data, len);
    r->d = new MoLiveDemux(&r->s);
    r->d->OnCompleteFrameReceived = [r](MoLiveChunk* c, Arr<byte> d) { r->q.push_back(RefFrame{c, d}); };
    return r;
}
void mobiref_moflex_close(void* h) { delete (RefMoflex*)h; }
// ReadPacket(): its return value; 0xFFFFFFFF when the reference threw (index outside the packet buffer, duplicate stream index)
uint32_t mobiref_moflex_read_packet(void* h) {
    RefMoflex* r = (RefMoflex*)h;
    try { return r->d->ReadPacket(); } catch (...) { return 0xFFFFFFFFu; }
}
long long mobiref_moflex_position(void* h) { return ((RefMoflex*)h)->s.Position; }
// Oldest frame OnCompleteFrameReceived delivered: the chunk's fields as 14 u32 in the order of mobi_moflex_stream
int mobiref_moflex_next_frame(void* h, uint32_t* st, const uint8_t** data, uint32_t* len) {
    RefMoflex* r = (RefMoflex*)h;
    if (r->head >= r->q.size()) return 0;
    RefFrame& f = r->q[r->head++];
    uint32_t v[14] = {0};
    v[1] = f.chunk->Id;
    if (f.chunk->IsStream()) v[0] = (uint32_t)((MoLiveStream*)f.chunk)->StreamIndex;
    if (f.chunk->Id == 1 || f.chunk->Id == 3) {
        MoLiveStreamVideo* c = (MoLiveStreamVideo*)f.chunk;
        v[2] = c->CodecId; v[3] = c->FpsRate; v[4] = c->FpsScale; v[5] = c->Width; v[6] = c->Height; v[7] = c->PelRatioRate; v[8] = c->PelRatioScale;
        if (f.chunk->Id == 3) { MoLiveStreamVideoWithLayout* w = (MoLiveStreamVideoWithLayout*)f.chunk; v[9] = (uint32_t)w->ImageLayout; v[10] = w->ImageRotation; }
    } else if (f.chunk->Id == 2) {
        MoLiveStreamAudio* c = (MoLiveStreamAudio*)f.chunk;
        v[2] = c->CodecId; v[11] = c->Frequency; v[12] = c->Channel;
    } else if (f.chunk->Id == 4) v[13] = ((MoLiveStreamTimeline*)f.chunk)->AssociatedStreamIndex;
    std::memcpy(st, v, sizeof v);
    *data = f.data.raw(); *len = (uint32_t)f.data.Length;
    return 1;
}

// The reference's MoflexMuxer (MoflexMuxer.cs:11-96) as a writer
struct RefMux { CsStream s; MoflexMuxer* m = nullptr; };
void* mobiref_mux_create(void) { RefMux* r = new RefMux(); r->m = new MoflexMuxer(&r->s); return r; }
void mobiref_mux_destroy(void* h) { delete (RefMux*)h; }
void mobiref_mux_synchro_header(void* h) { ((RefMux*)h)->m->WriteSynchroHeader(); }
void mobiref_mux_video_chunk(void* h, uint32_t fps_rate, uint32_t fps_scale, uint32_t w, uint32_t hh, uint32_t par, uint32_t pas, int stream_index, uint32_t codec) {
    MoLiveStreamVideo* c = new MoLiveStreamVideo(fps_rate, fps_scale, w, hh, par, pas);
    c->StreamIndex = stream_index; c->CodecId = codec;
    ((RefMux*)h)->m->WriteSynchroChunk(c);
}
void mobiref_mux_end_chunks(void* h) { ((RefMux*)h)->m->WriteSynchroChunk(nullptr); }
void mobiref_mux_data_block(void* h) { ((RefMux*)h)->m->WriteDataBlock(); }
void mobiref_mux_ep(void* h, int ep, const uint8_t* data, int len, int end_frame) {
    RefMux* r = (RefMux*)h;
    if (!data) { r->m->WriteEp(ep, nullptr, 0, 0); return; }
    Arr<byte> a = Arr<byte>::New(len);
    if (len) std::memcpy(a.raw(), data, (size_t)len);
    r->m->WriteEp(ep, a, 0, len, end_frame != 0);
}
void mobiref_mux_pad(void* h, int n) { RefMux* r = (RefMux*)h; Arr<byte> z = Arr<byte>::New(n); r->s.Write(z, 0, n); }
size_t mobiref_mux_bytes(void* h, uint8_t* out, size_t cap) {
    RefMux* r = (RefMux*)h;
    if (out && cap >= r->s.buf.size()) std::memcpy(out, r->s.buf.data(), r->s.buf.size());
    return r->s.buf.size();
}

}  // extern "C"
