// Container framing (include/mobidemux.h).  Host only.  "MODS:n" = LibMobiclip/Containers/Mods/ModsDemuxer.cs:n,
// "GUI:n" = MobiclipDecoder/Form1.cs:n of the reference.
#include "../../include/mobidemux.h"
#include "../../include/mobicuda.h"
#include <algorithm>
#include <cstring>
#include <deque>
#include <map>
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
    // the next cursor in 64 bits: a crafted block size must not wrap it (the reference's int arithmetic ends in an exception)
    uint64_t next = (uint64_t)offs + 4u + (uint64_t)(bs & ~1u);
    next = (next + 3u) & ~(uint64_t)3u;
    if (next <= offs || next > 0xFFFFFFFFull) return MOBI_ERR_BITSTREAM;
    *cursor = (uint32_t)next;
    return 1;
}

}  // extern "C"

// ---- Moflex ("MLD:n" = LibMobiclip/Containers/Moflex/MoLiveDemux.cs:n) ---------------------------------------------
namespace {

struct OutOfPacket {};  // an index outside the packet buffer: IndexOutOfRangeException in the reference

struct Packet {         // the reference's `byte[] packet`: PacketSize (or 0x1000) bytes, zero beyond what the stream had
    std::vector<uint8_t> b;
    uint8_t at(uint32_t i) const { if (i >= b.size()) throw OutOfPacket(); return b[i]; }
    uint32_t be16(uint32_t i) const { return (uint32_t)at(i) << 8 | at(i + 1); }
    uint32_t be24(uint32_t i) const { return (uint32_t)at(i) << 16 | (uint32_t)at(i + 1) << 8 | at(i + 2); }
    uint32_t be32(uint32_t i) const { return (uint32_t)at(i) << 24 | (uint32_t)at(i + 1) << 16 | (uint32_t)at(i + 2) << 8 | at(i + 3); }
};

// MSB-first reader over packet bytes; Pos after n bits = first byte not touched (MoLiveInBitStream.cs: bytes are pulled
// lazily, so the byte position is ceil(bits / 8) past the start)
struct EpBits {
    const Packet& p;
    uint32_t start, bits = 0;
    uint64_t pop(int n) {
        uint64_t v = 0;
        for (int i = 0; i < n; i++) {
            const uint32_t bit = bits++;
            v = v << 1 | ((p.at(start + (bit >> 3)) >> (7 - (bit & 7))) & 1u);
        }
        return v;
    }
    uint32_t pos() const { return start + ((bits + 7) >> 3); }
};

struct Endpoint { mobi_moflex_stream info; std::vector<uint8_t> data; };
struct DoneFrame { mobi_moflex_stream info; std::vector<uint8_t> data; };

}  // namespace

struct mobi_moflex {
    const uint8_t* d;
    size_t len, pos = 0;                  // Reader / Reader.Position
    uint64_t gts = 0, delta_gts = 0;
    uint32_t packet_size = 0, synchro_counter = 64, last_counter = 0;   // field defaults of MLD:22-31
    bool variable_packet_size = true, has_reference_ts = false, synchronized = false;
    std::map<int, Endpoint> streams;
    std::deque<DoneFrame> ready;
    DoneFrame current;

    void desynchronize() {  // MLD:57-65
        gts = 0; delta_gts = 0; synchro_counter = 64; last_counter = 65536; synchronized = false; streams.clear();
    }
    static bool synchro_header(const Packet& p, uint32_t o, uint64_t& ts, uint32_t& packet_size) {  // MLD:375-414
        ts = 0; packet_size = 0;
        if (!(p.at(o) == 0x4C && p.at(o + 1) == 0x32)) return false;
        const uint32_t sum = p.be16(o + 2);
        ts = (uint64_t)p.be32(o + 4) << 32 | p.be32(o + 8);
        uint32_t hi = (uint32_t)(ts >> 32);
        if ((int32_t)(hi - 1u) < 0) hi &= 0x7FFFFFFFu;
        packet_size = (p.be16(o + 12) + 1u) & 0xFFFFu;
        return sum == (uint32_t)(((ts >> 16) & 0xFFFF) ^ (hi >> 16) ^ 0xAAAA ^ (hi & 0xFFFF) ^ (ts & 0xFFFF));
    }
    static bool variable_byte(const Packet& p, uint32_t& value, uint32_t& pos, uint32_t psize) {  // MoLive.cs ReadVariableByte
        value = 0;
        for (int k = 0; k < 4; k++) {
            if (pos == psize) return false;
            const uint32_t b = p.at(pos++);
            if (k == 3) { value = value << 7 | b; return true; }   // the fourth byte contributes all eight bits
            if (!(b & 0x80)) { value = (k ? value << 7 : 0) | b; return true; }
            value = (k ? value << 7 : 0) | (b & 0x7F);
        }
        return true;
    }
    uint32_t synchro_chunk(const Packet& p, uint32_t& pos, uint32_t psize) {  // MLD:168-215
        uint32_t type, size;
        if (!variable_byte(p, type, pos, psize) || !variable_byte(p, size, pos, psize)) { desynchronize(); return 0x43; }
        uint32_t want;
        switch (type) {
        case 0: pos += size; return 0x100;
        case 1: want = 12; break;
        case 2: want = 6; break;
        case 3: want = 13; break;
        case 4: want = 2; break;
        case 0x100000: return 0x45;   // MoLiveChunkFoo.Read throws NotImplementedException
        default: return 0x44;
        }
        if (want != size) return 0x45;
        mobi_moflex_stream s;
        std::memset(&s, 0, sizeof s);
        s.chunk_id = type;
        s.stream_index = p.at(pos);
        const uint32_t have = (uint32_t)p.b.size();
        if (type == 4) { if (pos + 1 < have) s.associated_stream_index = p.at(pos + 1); }
        else if (pos + 1 < have) {
            s.codec_id = p.at(pos + 1);
            if (type == 2) {
                if (have - (pos + 2) >= 4) { s.frequency = p.be24(pos + 2) + 1; s.channels = p.at(pos + 5) + 1u; }
            } else if (have - (pos + 2) >= 0xA) {
                s.fps_rate = p.be16(pos + 2); s.fps_scale = p.be16(pos + 4); s.width = p.be16(pos + 6); s.height = p.be16(pos + 8);
                if (type == 1) { s.pel_ratio_rate = p.at(pos + 10); s.pel_ratio_scale = p.at(pos + 11); }
                else {
                    s.pel_ratio_rate = p.at(pos + 11);   // the reference assigns both bytes to PelRatioRate (MoLiveStreamVideoWithLayout.cs)
                    if (pos + 12 < have) { s.image_layout = p.at(pos + 12) & 0xF; s.image_rotation = p.at(pos + 12) >> 4; }
                }
            }
        }
        if (streams.count(s.stream_index)) return 0x45;   // Dictionary.Add on an existing key throws in the reference
        streams[s.stream_index].info = s;
        pos += size;
        if (pos <= psize) return 0;
        desynchronize();
        return 0x43;
    }
    uint32_t data_block(const Packet& p, uint32_t& pos, uint32_t psize) {  // MLD:217-268
        if (pos >= psize) { desynchronize(); return 67; }
        const uint32_t flags = p.at(pos++);
        variable_packet_size = flags & 1;
        const bool counting = (flags >> 1) & 1;
        const uint32_t sc = flags >> 2;
        if (synchro_counter == 64) synchro_counter = sc;
        else if (synchro_counter != sc) {
            if (delta_gts == 0) { desynchronize(); return 70; }
            gts += (uint64_t)(uint32_t)(sc - synchro_counter) * delta_gts;
            synchro_counter = sc;
            for (auto& kv : streams) kv.second.data.clear();
        }
        if (counting) {
            const uint32_t val = p.be16(pos);
            pos += 2;
            if (pos > psize) { desynchronize(); return 67; }
            const uint32_t expected = last_counter == 65536 ? val : last_counter + 1;
            if (expected != val) { last_counter = 65536; return 0x50; }
            last_counter = val;
        }
        return 0;
    }
    uint32_t end_point(const Packet& p, uint32_t& pos, uint32_t psize) {  // MLD:270-373
        if (pos == psize) return 0x101;
        if (pos > psize) { desynchronize(); return 0x43; }
        if (p.at(pos) == 0) {
            pos++;
            if (!variable_packet_size) pos = packet_size;
            return 0x101;
        }
        EpBits bs{p, pos};
        int idx_bits = 1;
        while (bs.pop(1) == 0) idx_bits++;
        const int stream_idx = (int)bs.pop(idx_bits);
        const bool end_frame = bs.pop(1) == 1;
        if (end_frame) {   // frame type and a signed timestamp delta: parsed, not used (MLD:298-314)
            int type_bits = 1;
            while (bs.pop(1) == 0) type_bits++;
            bs.pop(type_bits);
            int ts_bits = 28;
            bs.pop(1);
            while (bs.pop(1) == 0) ts_bits += 2;
            if (ts_bits > 64) throw OutOfPacket();   // Pop throws ArgumentException
            bs.pop(ts_bits);
        }
        const uint32_t ep_size = (uint32_t)bs.pop(13) + 1;
        pos = bs.pos();
        if (pos + ep_size > psize) { desynchronize(); return 0x43; }
        auto it = streams.find(stream_idx);
        if (it != streams.end()) {
            if (pos + ep_size > p.b.size()) throw OutOfPacket();
            it->second.data.insert(it->second.data.end(), p.b.begin() + pos, p.b.begin() + pos + ep_size);
        }
        pos += ep_size;
        if (end_frame && it != streams.end()) {
            it->second.data.push_back(0); it->second.data.push_back(0);   // MLD:353
            ready.push_back(DoneFrame{it->second.info, std::move(it->second.data)});
            it->second.data.clear();
        }
        return pos < psize ? 0 : 0x101;
    }
    uint32_t read_packet() {  // MLD:67-164
        Packet p;
        p.b.assign(packet_size == 0 ? 0x1000 : packet_size, 0);
        const uint32_t length = (uint32_t)std::min<size_t>(p.b.size(), len - std::min(pos, len));
        if (length) std::memcpy(p.b.data(), d + pos, length);
        uint64_t ts; uint32_t psz;
        if (!synchronized) {
            if (length < 0xE) return 1;
            uint32_t off = 0;
            while (!synchro_header(p, off, ts, psz)) { off++; if (off == length - 0xE) return 0x80; }
            has_reference_ts = (int64_t)(ts - 1) < 0;
            if (psz < 0x10) return 73;
            synchronized = true;
            pos += off;
            return 0;
        }
        if (packet_size != 0 && packet_size != length) return 73;
        uint32_t o2 = 0;
        if (length > 0xE && synchro_header(p, 0, ts, psz)) {
            if ((int64_t)(ts - 1) < 0) { has_reference_ts = true; ts &= 0x7FFFFFFFFFFFFFFFull; } else has_reference_ts = false;
            if (psz < 0x10) return 73;
            if (ts != 0) {
                if (gts != 0 && delta_gts == 0) delta_gts = ts - gts;
                gts = ts;
                streams.clear();
            }
            if (packet_size != psz) {
                const bool retry = (packet_size == 0 ? 0x1000u : packet_size) < psz;
                packet_size = psz;
                if (retry) return 0;
            }
            o2 = 0xE;
            const uint32_t size = packet_size > length ? length : packet_size;
            for (;;) {
                const uint32_t r = synchro_chunk(p, o2, size);
                if (r == 0x100) break;
                if (r != 0) return r;
            }
            if (o2 > length) return 0x43;
        }
        uint32_t r2 = data_block(p, o2, length);
        if (!synchronized) return 0;
        if (r2 == 0) {
            for (;;) {
                r2 = end_point(p, o2, length);
                if (r2 == 0x101) break;
                if (r2 != 0) return r2;
            }
            if (o2 > length) return 0x43;
            pos += o2;
            return 0;
        }
        return r2;
    }
};

extern "C" {

int mobi_moflex_open(const uint8_t* data, size_t len, mobi_moflex_t** out) {
    if (!out) return MOBI_ERR_ARG;
    *out = nullptr;
    if (!data) return MOBI_ERR_ARG;
    mobi_moflex* m = new (std::nothrow) mobi_moflex();
    if (!m) return MOBI_ERR_NOMEM;
    m->d = data; m->len = len;
    *out = m;
    return MOBI_OK;
}
void mobi_moflex_close(mobi_moflex_t* m) { delete m; }
uint32_t mobi_moflex_read_packet(mobi_moflex_t* m) {
    if (!m) return 0x43;
    try { return m->read_packet(); }
    catch (const OutOfPacket&) { m->desynchronize(); return 0x43; }
    catch (...) { return 0x43; }
}
int mobi_moflex_next_frame(mobi_moflex_t* m, mobi_moflex_stream* stream, const uint8_t** data, uint32_t* len) {
    if (!m || m->ready.empty()) return 0;
    m->current = std::move(m->ready.front());
    m->ready.pop_front();
    if (stream) *stream = m->current.info;
    if (data) *data = m->current.data.data();
    if (len) *len = (uint32_t)m->current.data.size();
    return 1;
}

}  // extern "C"
