"""GPU parity of the array-of-streams entry points opus_decode_batch / opus_encode_batch (include/opus_b200.h): one packet /
frame per stream and call, per-stream pointers, ret[i] = what the scalar call on stream i would return.  Streams of one call
differ in channel count, bitrate and (decode) packet length; NULL packets are lost packets."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu


def _cb():
    import concentus_b200 as cb
    assert cb.lib().opus_b200_init(0) == 0, "CUDA device required: no CPU fallback exists"
    return cb


def test_decode_batch_mixed_channels_and_loss():
    cb = _cb()
    L = cb.lib()
    fs, F = 960, 30
    cfgs = [(2, 64000, "music"), (1, 48000, "tone"), (2, 128000, "clicks"), (1, 32000, "noise"), (2, 96000, "tone"), (2, 510000, "music"),
            (1, 64000, "clicks")]
    n = len(cfgs)
    streams = []
    for i, (ch, br, kind) in enumerate(cfgs):
        x = O.test_signal(fs * F, ch, 200 + i, kind)
        d, o, l, _ = O.encode_stream(x, fs, br, ch, vbr=1, cvbr=0)
        d, o = O.pack(d, o, l)
        l = l.copy()
        l[(5 + i)::9] = 0                                     # lost packets: a NULL data pointer in the batch call
        streams.append((d, o, l, ch, O.decode_stream(d, o, l, fs, ch)))
    L.opus_decoder_create.restype = C.c_void_p
    err = C.c_int(0)
    hs = (C.c_void_p * n)(*[L.opus_decoder_create(48000, s[3], C.byref(err)) for s in streams])
    outs = [np.zeros((fs, s[3]), dtype=np.int16) for s in streams]
    pcm_ptrs = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
    rets = np.zeros(n, dtype=np.int32)
    v = C.c_uint32(0)
    for f in range(F):
        bufs = [np.ascontiguousarray(s[0][s[1][f]:s[1][f] + max(int(s[2][f]), 1)]) for s in streams]
        data_ptrs = (C.c_void_p * n)(*[(b.ctypes.data if int(s[2][f]) > 0 else None) for b, s in zip(bufs, streams)])
        lens = np.array([int(s[2][f]) for s in streams], dtype=np.int32)
        assert L.opus_decode_batch(hs, data_ptrs, O.ptr(lens), pcm_ptrs, fs, 0, O.ptr(rets), n) == 0
        for i, s in enumerate(streams):
            rp, rr, rret = s[4]
            assert rets[i] == rret[f], (i, f, int(rets[i]), int(rret[f]))
            assert np.array_equal(outs[i], rp[f * fs:(f + 1) * fs]), (i, f)
        if f % 6 == 2:                                        # a ctl in between pulls single states back from HBM
            for i in (0, 3, 6):
                L.opus_decoder_ctl(C.c_void_p(hs[i]), cb.OPUS_GET_FINAL_RANGE_REQUEST, C.byref(v))
                assert v.value == int(streams[i][4][1][f]), (i, f)
    for h in hs:
        L.opus_decoder_destroy(C.c_void_p(h))


def test_encode_batch_mixed_channels_and_settings():
    cb = _cb()
    L = cb.lib()
    fs, F = 960, 24
    cfgs = [(2, 96000, 1, 0, 10, "music"), (1, 32000, 0, 0, 5, "tone"), (2, 64000, 1, 1, 0, "clicks"), (1, 128000, 1, 0, 10, "noise"),
            (2, 256000, 0, 0, 8, "music"), (2, 48000, 1, 0, 3, "tone")]
    n = len(cfgs)
    L.opus_encoder_create.restype = C.c_void_p
    err = C.c_int(0)
    hs, pcms, refs = [], [], []
    for i, (ch, br, vbr, cvbr, cx, kind) in enumerate(cfgs):
        x = O.test_signal(fs * F, ch, 400 + i, kind)
        pcms.append(x)
        refs.append(O.encode_stream(x, fs, br, ch, vbr=vbr, cvbr=cvbr, complexity=cx, max_bytes=1276))
        h = L.opus_encoder_create(48000, ch, O.OPUS_APPLICATION_RESTRICTED_LOWDELAY, C.byref(err))
        for req, val in ((cb.OPUS_SET_BITRATE_REQUEST, br), (cb.OPUS_SET_VBR_REQUEST, vbr), (cb.OPUS_SET_VBR_CONSTRAINT_REQUEST, cvbr),
                         (cb.OPUS_SET_COMPLEXITY_REQUEST, cx)):
            assert L.opus_encoder_ctl(C.c_void_p(h), req, C.c_int32(val)) == 0
        hs.append(h)
    hs = (C.c_void_p * n)(*hs)
    outs = [np.zeros(1276, dtype=np.uint8) for _ in range(n)]
    out_ptrs = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
    rets = np.zeros(n, dtype=np.int32)
    v = C.c_uint32(0)
    for f in range(F):
        frames = [np.ascontiguousarray(p[f * fs:(f + 1) * fs]) for p in pcms]
        pcm_ptrs = (C.c_void_p * n)(*[fr.ctypes.data for fr in frames])
        assert L.opus_encode_batch(hs, pcm_ptrs, fs, out_ptrs, 1276, O.ptr(rets), n) == 0
        for i in range(n):
            rd, ro, rl, rr = refs[i]
            assert rets[i] == rl[f], (i, f, int(rets[i]), int(rl[f]))
            assert np.array_equal(outs[i][:rl[f]], rd[ro[f]:ro[f] + rl[f]]), (i, f)
        if f % 5 == 1:
            for i in (1, 4):
                L.opus_encoder_ctl(C.c_void_p(hs[i]), cb.OPUS_GET_FINAL_RANGE_REQUEST, C.byref(v))
                assert v.value == int(refs[i][3][f]), (i, f)
    for h in hs:
        L.opus_encoder_destroy(C.c_void_p(h))
