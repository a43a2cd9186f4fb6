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
