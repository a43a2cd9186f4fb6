"""ctypes bindings of include/mobicuda.h and include/mobisynth.h (one declaration per export)."""
import ctypes as C
import os

_LIBDIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'lib')

u8p = C.POINTER(C.c_uint8)


class FrameHdr(C.Structure):
    _fields_ = [('flags', C.c_uint32), ('n_mb', C.c_uint32), ('n_parts', C.c_uint32), ('n_ops', C.c_uint32),
                ('n_coefs', C.c_uint32), ('n_intra', C.c_uint32), ('quantizer', C.c_uint32), ('yuv_format', C.c_uint32),
                ('bytes_consumed', C.c_uint32), ('max_ref', C.c_uint32), ('n_inter_coefs', C.c_uint32), ('reserved', C.c_uint32 * 5), ('qtab', C.c_uint32 * 80)]


class Mb(C.Structure):
    _fields_ = [('info', C.c_uint32), ('first_sub', C.c_uint32), ('first_coef', C.c_uint32), ('intra_rank', C.c_uint32)]


class Part(C.Structure):
    _fields_ = [('xy', C.c_uint8), ('shape', C.c_uint8), ('mvx', C.c_int16), ('mvy', C.c_int16), ('pad', C.c_uint16)]


class Coef(C.Structure):
    _fields_ = [('level', C.c_int16), ('pos', C.c_uint8), ('blk', C.c_uint8)]


class PackedFrame(C.Structure):
    _fields_ = [('hdr', C.POINTER(FrameHdr)), ('mbs', C.POINTER(Mb)), ('parts', C.POINTER(Part)), ('ops', C.POINTER(C.c_uint32)),
                ('coefs', C.POINTER(Coef)), ('intra_list', C.POINTER(C.c_uint32))]


class BatchStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ('launches', 'frames', 'mbs', 'inter_mbs', 'intra_mbs', 'parts', 'coefs', 'ops', 'h2d_bytes', 'd2h_bytes', 'inter_coefs')]


# every symbol include/mobicuda.h declares: name -> (restype, argtypes)
MOBICUDA_EXPORTS = {
    'mobicuda_abi_version': (C.c_int, []),
    'mobicuda_selftest_bgra': (C.c_int, [C.c_int, C.POINTER(C.c_ulonglong)]),
    'mobi_parser_create': (C.c_int, [C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_void_p)]),
    'mobi_parser_destroy': (None, [C.c_void_p]),
    'mobi_parser_parse': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(PackedFrame)]),
    'mobi_parser_last_error': (C.c_char_p, [C.c_void_p]),
    'mobi_create': (C.c_int, [C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    'mobi_destroy': (None, [C.c_void_p]),
    'mobi_decode_frame': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    'mobi_submit_packed': (C.c_int, [C.c_void_p, C.POINTER(PackedFrame)]),
    'mobi_packed_validate': (C.c_int, [C.c_uint32, C.c_uint32, C.c_int, C.POINTER(PackedFrame), C.c_int, C.c_char_p, C.c_size_t]),
    'mobi_read_planes_strided': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    'mobi_read_yuv': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'mobi_read_bgra': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    'mobi_get_state': (C.c_int, [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_int)]),
    'mobi_last_error': (C.c_char_p, [C.c_void_p]),
    'mobi_batch_create': (C.c_int, [C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    'mobi_batch_destroy': (None, [C.c_void_p]),
    'mobi_batch_decode': (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    'mobi_batch_submit': (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int]),
    'mobi_batch_fetch': (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    'mobi_batch_read_yuv': (C.c_int, [C.c_void_p, C.c_void_p]),
    'mobi_batch_read_planes_strided': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    'mobi_batch_read_bgra': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    'mobi_batch_read_bgra_all': (C.c_int, [C.c_void_p, C.c_void_p]),
    'mobi_batch_last_error': (C.c_char_p, [C.c_void_p]),
    'mobi_batch_stage': (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    'mobi_batch_replay': (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    'mobi_batch_replay_convert': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    'mobi_batch_staged_steps': (C.c_int, [C.c_void_p]),
    'mobi_batch_clear_staged': (None, [C.c_void_p]),
    'mobi_batch_reset': (C.c_int, [C.c_void_p]),
    'mobi_batch_reset_streams': (C.c_int, [C.c_void_p]),
    'mobi_batch_sync': (C.c_int, [C.c_void_p]),
    'mobi_batch_cuda_stream': (C.c_void_p, [C.c_void_p]),
    'mobi_batch_get_stats': (C.c_int, [C.c_void_p, C.POINTER(BatchStats)]),
    'mobi_batch_clear_stats': (None, [C.c_void_p]),
    'mobi_batch_get_phase_times': (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    'mobi_batch_set_kernel_timing': (C.c_int, [C.c_void_p, C.c_int]),
    'mobi_batch_get_kernel_times': (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
}


class ModsHeader(C.Structure):
    _fields_ = [('magic', C.c_char * 4), ('tag_id', C.c_uint16), ('tag_id_size_dword', C.c_uint16), ('frame_count', C.c_uint32),
                ('width', C.c_uint32), ('height', C.c_uint32), ('fps', C.c_uint32), ('audio_codec', C.c_uint16), ('nb_channel', C.c_uint16),
                ('frequency', C.c_uint32), ('biggest_frame', C.c_uint32), ('audio_offset', C.c_uint32), ('keyframe_index_offset', C.c_uint32),
                ('keyframe_count', C.c_uint32)]


class Moc5Info(C.Structure):
    _fields_ = [('width', C.c_uint32), ('height', C.c_uint32), ('fps_x128', C.c_uint32), ('first_block', C.c_uint32)]


class MoflexStream(C.Structure):
    _fields_ = [('stream_index', C.c_int32)] + [(n, C.c_uint32) for n in (
        'chunk_id', 'codec_id', 'fps_rate', 'fps_scale', 'width', 'height', 'pel_ratio_rate', 'pel_ratio_scale', 'image_layout',
        'image_rotation', 'frequency', 'channels', 'associated_stream_index')]


# every symbol include/mobidemux.h declares (same library)
MOBIDEMUX_EXPORTS = {
    'mobi_moflex_open': (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    'mobi_moflex_close': (None, [C.c_void_p]),
    'mobi_moflex_read_packet': (C.c_uint32, [C.c_void_p]),
    'mobi_moflex_next_frame': (C.c_int, [C.c_void_p, C.POINTER(MoflexStream), C.POINTER(C.c_void_p), C.POINTER(C.c_uint32)]),
    'mobi_mods_open': (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    'mobi_mods_close': (None, [C.c_void_p]),
    'mobi_mods_get_header': (C.c_int, [C.c_void_p, C.POINTER(ModsHeader)]),
    'mobi_mods_keyframe': (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    'mobi_mods_read_frame': (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_int)]),
    'mobi_mods_jump_to_keyframe': (C.c_int, [C.c_void_p, C.c_uint32]),
    'mobi_moc5_open': (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(Moc5Info)]),
    'mobi_moc5_next': (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
}


class SynthParamsC(C.Structure):
    _fields_ = [('width', C.c_uint32), ('height', C.c_uint32), ('version', C.c_int32), ('seed', C.c_uint64), ('gop', C.c_int32),
                ('quant', C.c_int32), ('p_dquant', C.c_float), ('p_split', C.c_float), ('p_intra_mb', C.c_float), ('p_sub_mb', C.c_float),
                ('p_cbp', C.c_float), ('p_blk8', C.c_float), ('mean_coefs', C.c_float), ('p_escape', C.c_float), ('mv_range', C.c_int32),
                ('p_ref1', C.c_float), ('p_zero_mv', C.c_float), ('p_oob_mv', C.c_float), ('inter_only', C.c_int32), ('gop_phase', C.c_int32)]


class SynthStats(C.Structure):
    _fields_ = [('n_mb', C.c_uint32), ('n_intra_mb', C.c_uint32), ('n_leaves', C.c_uint32), ('n_coefs', C.c_uint32), ('n_blk8', C.c_uint32),
                ('n_blk4', C.c_uint32), ('shape_hist', C.c_uint32 * 16), ('phase_hist', C.c_uint32 * 4), ('mode_hist', C.c_uint32 * 20),
                ('ref_hist', C.c_uint32 * 6)]


MOBISYNTH_EXPORTS = {
    'mobi_synth_default_params': (None, [C.POINTER(SynthParamsC), C.c_uint32, C.c_uint32, C.c_int, C.c_uint64]),
    'mobi_synth_create': (C.c_void_p, [C.POINTER(SynthParamsC)]),
    'mobi_synth_destroy': (None, [C.c_void_p]),
    'mobi_synth_next_frame': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    'mobi_synth_last_stats': (None, [C.c_void_p, C.POINTER(SynthStats)]),
}

_libs = {}


def _load(name, exports):
    if name in _libs:
        return _libs[name]
    path = os.path.join(_LIBDIR, name)
    if not os.path.exists(path):
        raise ImportError('%s is not built: run `python -c "import __graft_entry__ as g; g.build()"` (or python -m mobiclipdecoder_b200._build) '
                          'from the repository root. There is no CPU fallback.' % path)
    lib = C.CDLL(path)
    for sym, (res, args) in exports.items():
        fn = getattr(lib, sym)  # AttributeError if the library does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    _libs[name] = lib
    return lib


def mobicuda():
    return _load('libmobicuda.so', dict(MOBICUDA_EXPORTS, **MOBIDEMUX_EXPORTS))


def mobisynth():
    return _load('libmobisynth.so', MOBISYNTH_EXPORTS)
