"""concentus_b200 — B200-native batched CELT (Opus) codec engine.

The product is libconcentus_b200.so (hand-written sm_100a CUDA + a libopus-compatible C ABI, see
include/opus_b200.h).  This package is only the thin Python binding used by the tests and bench.py: it
loads the shared library with ctypes and mirrors the C entry points one-to-one.  It never falls back
to a CPU implementation: if the library is missing it raises, and on a box without a usable CUDA device
every codec call returns OPUS_INTERNAL_ERROR.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CB200_LIB", os.path.join(HERE, "libconcentus_b200.so"))   # override: A/B builds only

OPUS_OK = 0
OPUS_BAD_ARG = -1
OPUS_BUFFER_TOO_SMALL = -2
OPUS_INTERNAL_ERROR = -3
OPUS_INVALID_PACKET = -4
OPUS_UNIMPLEMENTED = -5
OPUS_INVALID_STATE = -6
OPUS_ALLOC_FAIL = -7
OPUS_RESET_STATE = 4028
OPUS_GET_FINAL_RANGE_REQUEST = 4031
OPUS_GET_BANDWIDTH_REQUEST = 4009
OPUS_GET_SAMPLE_RATE_REQUEST = 4029
OPUS_GET_PITCH_REQUEST = 4033
OPUS_SET_GAIN_REQUEST = 4034
OPUS_GET_GAIN_REQUEST = 4045
OPUS_GET_LAST_PACKET_DURATION_REQUEST = 4039

_lib = None

EXPORTS = [
    "opus_decoder_get_size", "opus_decoder_create", "opus_decoder_init", "opus_decode", "opus_decoder_ctl",
    "opus_decoder_destroy", "opus_packet_parse", "opus_packet_get_bandwidth", "opus_packet_get_samples_per_frame",
    "opus_packet_get_nb_channels", "opus_packet_get_nb_frames", "opus_packet_get_nb_samples",
    "opus_decoder_get_nb_samples", "opus_strerror", "opus_get_version_string", "opus_decode_batch", "opus_decode_span",
    "opus_decode_span_device", "opus_decoder_sync", "opus_b200_init", "opus_b200_synchronize", "opus_b200_stream",
    "opus_b200_kernel_launches", "opus_b200_last_kernel_ms", "opus_b200_stage_times",
]


def lib():
    """Load the CUDA library (building is __graft_entry__.build()'s job).  Raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libconcentus_b200.so is not built (run `python -m concentus_b200.build`); "
                               "there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.opus_decoder_create.restype = C.c_void_p
        L.opus_decoder_create.argtypes = [C.c_int32, C.c_int, C.POINTER(C.c_int)]
        L.opus_decoder_init.argtypes = [C.c_void_p, C.c_int32, C.c_int]
        L.opus_decoder_destroy.argtypes = [C.c_void_p]
        L.opus_decoder_destroy.restype = None
        L.opus_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int, C.c_int]
        L.opus_decode_span.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_int, C.c_void_p]
        L.opus_decode_span_device.argtypes = L.opus_decode_span.argtypes
        L.opus_decode_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.opus_decoder_sync.argtypes = [C.c_void_p, C.c_int]
        L.opus_strerror.restype = C.c_char_p
        L.opus_get_version_string.restype = C.c_char_p
        L.opus_b200_stream.restype = C.c_void_p
        L.opus_b200_kernel_launches.restype = C.c_longlong
        L.opus_b200_last_kernel_ms.restype = C.c_float
        L.opus_b200_stage_times.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class DecoderBatch:
    """n independent decoders (same Fs / channels) driven through the batch C ABI."""

    def __init__(self, n, Fs=48000, channels=2):
        L = lib()
        self.n, self.Fs, self.channels = n, Fs, channels
        err = C.c_int(0)
        self.handles = (C.c_void_p * n)()
        for i in range(n):
            h = L.opus_decoder_create(Fs, channels, C.byref(err))
            if not h:
                raise RuntimeError("opus_decoder_create failed: %d" % err.value)
            self.handles[i] = h

    def decode_span(self, data, offs, lens, F, frame_size):
        """Host buffers in, host buffers out: returns (pcm [n*F*frame_size, channels], rets [n*F])."""
        L = lib()
        pcm = np.zeros((self.n * F * frame_size, self.channels), dtype=np.int16)
        rets = np.zeros(self.n * F, dtype=np.int32)
        rc = L.opus_decode_span(self.handles, self.n, F, _p(data), _p(np.ascontiguousarray(offs, dtype=np.int64)),
                                _p(np.ascontiguousarray(lens, dtype=np.int32)), _p(pcm), frame_size, _p(rets))
        if rc != OPUS_OK:
            raise RuntimeError("opus_decode_span: %s" % L.opus_strerror(rc).decode())
        return pcm, rets

    def final_ranges(self):
        L = lib()
        out = np.zeros(self.n, dtype=np.uint32)
        v = C.c_uint32(0)
        for i in range(self.n):
            L.opus_decoder_ctl(C.c_void_p(self.handles[i]), OPUS_GET_FINAL_RANGE_REQUEST, C.byref(v))
            out[i] = v.value
        return out

    def close(self):
        L = lib()
        for i in range(self.n):
            if self.handles[i]:
                L.opus_decoder_destroy(C.c_void_p(self.handles[i]))
                self.handles[i] = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
