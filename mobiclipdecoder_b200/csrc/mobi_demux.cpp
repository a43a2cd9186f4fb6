// Container framing (include/mobidemux.h).  Host only.  "MODS:n" = LibMobiclip/Containers/Mods/ModsDemuxer.cs:n,
// "GUI:n" = MobiclipDecoder/Form1.cs:n of the reference.
#include "../../include/mobidemux.h"
#include "../../include/mobicuda.h"
#include <cstring>
#include <new>
#include <vector>

namespace {
inline uint16_t u16(const uint8_t* p) { return (uint16_t)(p[0] | p[1] << 8); }
inline uint32_t u32(const uint8_t* p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
}  // namespace

struct mobi_mods {
    const uint8_t* d;
    size_t len;
    mobi_mods_header h;
    std::vector<uint32_t> key_frame, key_offset;
    size_t pos = 0;          // Stream.Position
    uint32_t cur_frame = 0;  // CurFrame
    int next_key = -1;       // NextKeyFrame
};

extern "C" {

int mobi_mods_open(const uint8_t* data, size_t len, mobi_mods_t** out) {
    if (!out) return MOBI_ERR_ARG;
    *out = nullptr;
    if (!data || len < 0x30) return MOBI_ERR_BITSTREAM;
    mobi_mods* m = new (std::nothrow) mobi_mods();
    if (!m) return MOBI_ERR_NOMEM;
    m->d = data; m->len = len;
    mobi_mods_header& h = m->h;  // MODS:48-64
    std::memcpy(h.magic, data, 4);
    h.tag_id = u16(data + 4); h.tag_id_size_dword = u16(data + 6);
    h.frame_count = u32(data + 8); h.width = u32(data + 0xC); h.height = u32(data + 0x10); h.fps = u32(data + 0x14);
    h.audio_codec = u16(data + 0x18); h.nb_channel = u16(data + 0x1A);
    h.frequency = u32(data + 0x1C); h.biggest_frame = u32(data + 0x20); h.audio_offset = u32(data + 0x24);
    h.keyframe_index_offset = u32(data + 0x28); h.keyframe_count = u32(data + 0x2C);
    // MODS:30-39: the key-frame table; a table outside the file would read zeros / throw in the reference
    if ((uint64_t)h.keyframe_index_offset + 8ull * h.keyframe_count > len) { delete m; return MOBI_ERR_BITSTREAM; }
    m->key_frame.resize(h.keyframe_count); m->key_offset.resize(h.keyframe_count);
    for (uint32_t i = 0; i < h.keyframe_count; i++) {
        m->key_frame[i] = u32(data + h.keyframe_index_offset + 8 * (size_t)i);
        m->key_offset[i] = u32(data + h.keyframe_index_offset + 8 * (size_t)i + 4);
    }
    // MODS:40 JumpToKeyFrame(0); with no key frames the stream position stays behind the key-frame table (MODS:31-38)
    m->pos = (size_t)h.keyframe_index_offset + 8 * (size_t)h.keyframe_count;
    mobi_mods_jump_to_keyframe(m, 0);
    *out = m;
    return MOBI_OK;
}
void mobi_mods_close(mobi_mods_t* m) { delete m; }
int mobi_mods_get_header(const mobi_mods_t* m, mobi_mods_header* h) {
    if (!m || !h) return MOBI_ERR_ARG;
    *h = m->h;
    return MOBI_OK;
}
int mobi_mods_keyframe(const mobi_mods_t* m, uint32_t i, uint32_t* frame_number, uint32_t* data_offset) {
    if (!m || i >= m->h.keyframe_count) return MOBI_ERR_ARG;
    if (frame_number) *frame_number = m->key_frame[i];
    if (data_offset) *data_offset = m->key_offset[i];
    return MOBI_OK;
}
int mobi_mods_jump_to_keyframe(mobi_mods_t* m, uint32_t k) {  // MODS:88-95
    if (!m) return MOBI_ERR_ARG;
    if (k >= m->h.keyframe_count) return MOBI_OK;  // the reference returns silently
    m->pos = m->key_offset[k];
    m->cur_frame = m->key_frame[k];
    m->next_key = (k + 1 < m->h.keyframe_count) ? (int)k + 1 : -1;
    return MOBI_OK;
}
int mobi_mods_read_frame(mobi_mods_t* m, const uint8_t** frame, uint32_t* frame_len, uint32_t* nr_audio, int* is_key) {  // MODS:97-117
    if (!m || !frame || !frame_len) return MOBI_ERR_ARG;
    if (nr_audio) *nr_audio = 0;
    if (is_key) *is_key = 0;
    if (m->cur_frame >= m->h.frame_count) return 0;
    bool key = false;
    if (m->next_key >= 0 && (uint32_t)m->next_key < m->h.keyframe_count && m->cur_frame == m->key_frame[m->next_key]) {
        key = true;
        m->next_key = ((uint32_t)m->next_key + 1 < m->h.keyframe_count) ? m->next_key + 1 : -1;
    }
    if (m->pos + 4 > m->len) return MOBI_ERR_BITSTREAM;
    const uint32_t info = u32(m->d + m->pos), size = info >> 14;
    if (m->pos + 4 + (size_t)size > m->len) return MOBI_ERR_BITSTREAM;
    m->cur_frame++;
    *frame = m->d + m->pos + 4; *frame_len = size;
    if (nr_audio) *nr_audio = info & 0x3FFF;
    if (is_key) *is_key = key;
    m->pos += 4 + (size_t)size;
    return 1;
}

int mobi_moc5_open(const uint8_t* data, size_t len, mobi_moc5_info* info) {  // GUI:285-289
    if (!data || !info) return MOBI_ERR_ARG;
    if (len < 0x24) return MOBI_ERR_BITSTREAM;
    info->first_block = u32(data + 4) + 8;
    info->fps_x128 = u32(data + 0xC);
    info->width = u32(data + 0x1C);
    info->height = u32(data + 0x20);
    return MOBI_OK;
}
int mobi_moc5_next(const uint8_t* data, size_t len, uint32_t* cursor, uint32_t* decode_offset, uint32_t* block_size) {  // GUI:294-318
    if (!data || !cursor) return MOBI_ERR_ARG;
    uint32_t offs = *cursor;
    if (offs >= len) return 0;
    if ((size_t)offs + 4 > len) return MOBI_ERR_BITSTREAM;
    const uint32_t bs = u32(data + offs);
    if (decode_offset) *decode_offset = offs + 8;
    if (block_size) *block_size = bs;
    offs += 4 + (bs & ~1u);
    while (offs % 4) offs++;
    *cursor = offs;
    return 1;
}

}  // extern "C"
