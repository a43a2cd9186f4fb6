"""TEST INFRASTRUCTURE: (1) writers that wrap synthetic frames into Mods / MOC5 container bytes laid out the way the
reference readers expect, (2) a pure-Python restatement of those readers (ModsDemuxer.cs:16-117, Form1.cs:285-318) used
as the oracle for the native demuxers."""
import struct


def write_mods(frames, width, height, fps=0x18000000, tag_id=0x324E, audio_packets=None):
    """frames: list of (payload bytes, is_key).  Layout: 0x30 header | key-frame table | packets (u32 info + payload)."""
    keys = [i for i, (_, k) in enumerate(frames) if k]
    table_off = 0x30
    data_off = table_off + 8 * len(keys)
    packets, offs, at = [], [], data_off
    for i, (p, _) in enumerate(frames):
        na = 0 if audio_packets is None else audio_packets[i]
        offs.append(at)
        packets.append(struct.pack('<I', (len(p) << 14) | (na & 0x3FFF)) + p)
        at += 4 + len(p)
    hdr = b'MODS' + struct.pack('<HHIIIIHHIIIII', tag_id, 0, len(frames), width, height, fps, 0, 0, 0, max(len(p) for p, _ in frames), 0,
                                table_off, len(keys))
    assert len(hdr) == 0x30
    table = b''.join(struct.pack('<II', k, offs[k]) for k in keys)
    return hdr + table + b''.join(packets)


def read_mods_reference(data):
    """ModsDemuxer restated: returns (header dict, key-frame list, [(payload, nr_audio, is_key), ...])."""
    f = struct.unpack_from('<4sHHIIIIHHIIIII', data, 0)
    names = ['magic', 'tag_id', 'tag_id_size_dword', 'frame_count', 'width', 'height', 'fps', 'audio_codec', 'nb_channel', 'frequency',
             'biggest_frame', 'audio_offset', 'keyframe_index_offset', 'keyframe_count']
    h = dict(zip(names, f))
    keys = [struct.unpack_from('<II', data, h['keyframe_index_offset'] + 8 * i) for i in range(h['keyframe_count'])]
    out = []
    if not keys:
        return h, keys, out
    pos, cur, nxt = keys[0][1], keys[0][0], (1 if len(keys) > 1 else -1)     # JumpToKeyFrame(0)
    while cur < h['frame_count']:
        is_key = False
        if 0 <= nxt < len(keys) and cur == keys[nxt][0]:
            is_key = True
            nxt = nxt + 1 if nxt + 1 < len(keys) else -1
        cur += 1
        info = struct.unpack_from('<I', data, pos)[0]
        size = info >> 14
        out.append((data[pos + 4:pos + 4 + size], info & 0x3FFF, is_key))
        pos += 4 + size
    return h, keys, out


def write_moc5(frames, width, height, fps_x128=30 * 128, header_len=0xE0):
    """Header fields at 0x4 / 0xC / 0x1C / 0x20, then per frame: u32 block size, 4 more header bytes, payload, padded the way
    the reader steps (Form1.cs:317-318)."""
    hdr = bytearray(header_len)
    hdr[0:4] = b'MOC5'
    struct.pack_into('<I', hdr, 4, header_len - 8)
    struct.pack_into('<I', hdr, 0xC, fps_x128)
    struct.pack_into('<II', hdr, 0x1C, width, height)
    body = bytearray()
    for p, _ in frames:
        block = struct.pack('<I', 0) + p      # what lies between the size word and the next block: 4 bytes + payload
        bs = len(block)
        assert bs % 2 == 0   # the reader steps by size & ~1 (Form1.cs:317): payloads are whole 16-bit words
        body += struct.pack('<I', bs) + block
        while (len(hdr) + len(body)) % 4:
            body += b'\\0'
    return bytes(hdr) + bytes(body)


def read_moc5_reference(data):
    offs = struct.unpack_from('<I', data, 4)[0] + 8
    w, h = struct.unpack_from('<II', data, 0x1C)
    fps = struct.unpack_from('<I', data, 0xC)[0]
    out = []
    while offs < len(data):
        bs = struct.unpack_from('<I', data, offs)[0]
        out.append((offs + 8, bs))
        offs += 4 + (bs & ~1)
        while offs % 4:
            offs += 1
    return (w, h, fps), out


# ---- Moflex writer: MoflexMuxer.cs:21-95 + MoflexSimpleVideoMuxer.cs:13-70 restated (the reference's own writer) ----
def _variable_byte(v):
    assert v < (1 << 28)
    if v < 0x80:
        return bytes([v])
    if v < 0x2000:
        return bytes([(v >> 7) | 0x80, v & 0x7F])
    if v < 0x200000:
        return bytes([(v >> 14) | 0x80, ((v >> 7) & 0x7F) | 0x80, v & 0x7F])
    return bytes([((v >> 21) | 0x80) & 0xFF, ((v >> 14) | 0x80) & 0xFF, ((v >> 7) & 0x7F) | 0x80, v & 0x7F])


def _synchro_header(ts=1, packet_size_field=0x1000):
    hi = (ts >> 32) & 0xFFFFFFFF
    v19 = hi & 0x7FFFFFFF if ((hi - 1) & 0xFFFFFFFF) >= 0x80000000 else hi
    crc = ((ts >> 16) & 0xFFFF) ^ (v19 >> 16) ^ 0xAAAA ^ (v19 & 0xFFFF) ^ (ts & 0xFFFF)
    return b'L2' + struct.pack('>HQH', crc & 0xFFFF, ts, packet_size_field)


def _video_chunk(stream_index, codec_id, fps_rate, fps_scale, w, h):
    body = struct.pack('>BBHHHHBB', stream_index, codec_id, fps_rate, fps_scale, w, h, 1, 1)
    return _variable_byte(1) + _variable_byte(12) + body


def _ep(ep, data, end_frame):
    """WriteEp (MoflexMuxer.cs:52-94); like the reference's writer this is only right for Ep 0 and 1."""
    if data is None:
        return b'\0'
    nrbits = 1 if ep == 0 else ep.bit_length()
    bits = '0' * (nrbits - 1) + '1' + format(ep, '0%db' % nrbits) + ('1' if end_frame else '0')
    if end_frame:
        bits += '1' + '0' + '0' + '1' + '0' * 28     # frame type 1 bit = 1; sign 0; ts length marker; 28-bit timestamp 0
    bits += format(len(data) - 1, '013b')
    nbytes = (len(bits) + 4) // 8
    bits = bits.ljust(nbytes * 8, '0')
    return int(bits, 2).to_bytes(nbytes, 'big') + data


def write_moflex(frames, width, height, fps_rate=24, fps_scale=1, max_ep=0x1000 - 0x80):
    """frames: list of payload bytes (without the two pad bytes the demuxer appends)."""
    out = bytearray(_synchro_header() + _video_chunk(0, 0, fps_rate, fps_scale, width, height) + _variable_byte(0) + _variable_byte(0))
    for data in frames:
        pos, left = 0, len(data)
        if left <= max_ep:
            out += b'\x01' + _ep(0, data, True) + _ep(0, None, False)
            continue
        while left >= max_ep:
            out += b'\x01' + _ep(0, data[pos:pos + max_ep], left == max_ep) + _ep(0, None, False)
            pos += max_ep
            left -= max_ep
        if left > 0:
            out += b'\x01' + _ep(0, data[pos:pos + left], True) + _ep(0, None, False)
    out += bytes(0x1000)    # FinalizeMoflex
    return bytes(out)
