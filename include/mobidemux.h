/* Container framing in front of the frame path (host only, part of libmobicuda.so).
 * These hand the decoder the same (Data, Offset) pairs the reference front-ends do.  Zero-copy: frames are views into
 * the caller's container bytes, which must stay alive while the handle is used.
 *
 *   Mods (Nintendo DS)   LibMobiclip/Containers/Mods/ModsDemuxer.cs
 *   MOC5 (Wii)           framing parsed ad hoc in MobiclipDecoder/Form1.cs:282-320 (the CLI refuses MOC5, Program.cs:360-366)
 *   Moflex (3DS)         LibMobiclip/Containers/Moflex/MoLiveDemux.cs (packets, stream table, end-point reassembly)
 */
#ifndef MOBIDEMUX_H
#define MOBIDEMUX_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ModsDemuxer.ModsHeader (ModsDemuxer.cs:44-80): the 0x30-byte file header, field for field. */
typedef struct mobi_mods_header {
    char magic[4];             /* ModsString */
    uint16_t tag_id, tag_id_size_dword;
    uint32_t frame_count, width, height, fps; /* fps is 8.24 fixed point (MobiConverter/Program.cs:246) */
    uint16_t audio_codec, nb_channel;
    uint32_t frequency, biggest_frame, audio_offset, keyframe_index_offset, keyframe_count;
} mobi_mods_header;

typedef struct mobi_mods mobi_mods_t;

/* new ModsDemuxer(stream) (ModsDemuxer.cs:16-41): header, key-frame table, position on key frame 0.
 * Returns 0, or <0 (mobi_status) when the bytes cannot be a Mods file (truncated header / tables outside the file). */
int mobi_mods_open(const uint8_t* data, size_t len, mobi_mods_t** out);
void mobi_mods_close(mobi_mods_t* m);
int mobi_mods_get_header(const mobi_mods_t* m, mobi_mods_header* h);
/* KeyFrames[i] (ModsDemuxer.cs:82-87). Returns <0 when i is out of range. */
int mobi_mods_keyframe(const mobi_mods_t* m, uint32_t i, uint32_t* frame_number, uint32_t* data_offset);
/* ReadFrame(out NrAudioPackets, out IsKeyFrame) (ModsDemuxer.cs:97-117): *frame / *frame_len view the packet
 * (size = info >> 14, audio packets = info & 0x3FFF).  Returns 1 when a frame was produced, 0 at end of stream
 * (the reference returns null), <0 when the packet runs past the end of the file. */
int mobi_mods_read_frame(mobi_mods_t* m, const uint8_t** frame, uint32_t* frame_len, uint32_t* nr_audio_packets, int* is_key_frame);
/* JumpToKeyFrame (ModsDemuxer.cs:88-95); the reference only ever calls it with 0. */
int mobi_mods_jump_to_keyframe(mobi_mods_t* m, uint32_t keyframe);

/* MOC5: header fields at 0x4 (offset of the first block - 8), 0xC (fps * 128), 0x1C / 0x20 (width / height);
 * then per frame a u32 block size; the decoder is handed the WHOLE file and Offset = block + 8; the next block is at
 * block + 4 + (size & ~1), rounded up to a multiple of 4 (Form1.cs:285-318). */
typedef struct mobi_moc5_info { uint32_t width, height, fps_x128, first_block; } mobi_moc5_info;
int mobi_moc5_open(const uint8_t* data, size_t len, mobi_moc5_info* info);
/* Advances *cursor (initialise it with info.first_block).  *decode_offset receives the Offset to hand the decoder
 * together with the whole file as Data.  Returns 1 per frame, 0 when cursor >= len (Form1.cs:294), <0 on a block header
 * outside the file. */
int mobi_moc5_next(const uint8_t* data, size_t len, uint32_t* cursor, uint32_t* decode_offset, uint32_t* block_size);

/* ---- Moflex (3DS): MoLiveDemux (MoLiveDemux.cs:11-416) -------------------------------------------------------------
 * Packets start with an optional synchro header ("L2", checksum, 64-bit timestamp, packet size; :375-414) followed by
 * synchro chunks describing the streams (:168-215), then a data block flag byte (:217-268) and end-points: bit-packed
 * headers (stream index, end-of-frame marker, 13-bit size; :270-373) each followed by a slice of a stream's current
 * frame.  A frame is complete at an end-point flagged EndFrame; two zero bytes are appended (:353) and the
 * OnCompleteFrameReceived event fires -- here: the frame is queued for mobi_moflex_next_frame. */
typedef struct mobi_moflex mobi_moflex_t;

typedef struct mobi_moflex_stream {   /* the MoLiveStream* chunk of the frame's stream */
    int32_t stream_index;
    uint32_t chunk_id;                /* 1 MoLiveStreamVideo, 2 MoLiveStreamAudio, 3 MoLiveStreamVideoWithLayout, 4 MoLiveStreamTimeline */
    uint32_t codec_id;
    uint32_t fps_rate, fps_scale, width, height, pel_ratio_rate, pel_ratio_scale;  /* video (MoLiveStreamVideo.cs:33-49) */
    uint32_t image_layout, image_rotation;                                        /* chunk 3 (MoLiveStreamVideoWithLayout.cs) */
    uint32_t frequency, channels;                                                 /* audio (MoLiveStreamAudio.cs) */
    uint32_t associated_stream_index;                                             /* timeline */
} mobi_moflex_stream;

int mobi_moflex_open(const uint8_t* data, size_t len, mobi_moflex_t** out);
void mobi_moflex_close(mobi_moflex_t* m);
/* MoLiveDemux.ReadPacket() (:67-164).  Returns the reference's status code: 0 packet consumed (or resynchronised), 1 fewer
 * than 14 bytes left, 0x80 no synchro pattern, 73 short / inconsistent packet (the CLI's end condition, Program.cs:162-166),
 * 0x43-0x50 framing errors (the demuxer desynchronises as the reference does).  Where the reference would throw
 * (index outside the 4 KiB packet buffer, duplicate stream index) 0x43 / 0x45 is returned instead. */
uint32_t mobi_moflex_read_packet(mobi_moflex_t* m);
/* Oldest completed frame not yet handed out.  Returns 1 and fills the outputs (valid until the next call on this handle),
 * 0 when none is queued. */
int mobi_moflex_next_frame(mobi_moflex_t* m, mobi_moflex_stream* stream, const uint8_t** data, uint32_t* len);

#ifdef __cplusplus
}
#endif
#endif
