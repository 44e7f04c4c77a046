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


def test_decode_batch_mixed_sample_rates():
    """Decoders of different API rates (and channel counts) in ONE opus_decode_batch call: the call groups them by (Fs, channels);
    a span call given mixed rates is an argument error (its staging geometry is per call)."""
    cb = _cb()
    L = cb.lib()
    F = 20
    x = O.test_signal(960 * F, 2, 77, "music")
    d, o, l, _ = O.encode_stream(x, 960, 64000, 2, vbr=1, cvbr=0)
    d, o = O.pack(d, o, l)
    cfgs = [(48000, 2), (8000, 2), (16000, 1), (24000, 2), (48000, 1), (12000, 1)]
    n = len(cfgs)
    refs = [O.decode_stream(d, o, l, 960 * Fs // 48000, ch, Fs=Fs) for Fs, ch in cfgs]
    L.opus_decoder_create.restype = C.c_void_p
    err = C.c_int(0)
    hs = (C.c_void_p * n)(*[L.opus_decoder_create(Fs, ch, C.byref(err)) for Fs, ch in cfgs])
    cap = 960                                                 # one capacity for all: 20 ms at 48 kHz, more than enough below
    outs = [np.zeros((cap, ch), dtype=np.int16) for _, ch in cfgs]
    pcm_ptrs = (C.c_void_p * n)(*[b.ctypes.data for b in outs])
    rets = np.zeros(n, dtype=np.int32)
    for f in range(F):
        pk = np.ascontiguousarray(d[o[f]:o[f] + l[f]])
        data_ptrs = (C.c_void_p * n)(*[pk.ctypes.data] * n)
        lens = np.full(n, int(l[f]), dtype=np.int32)
        assert L.opus_decode_batch(hs, data_ptrs, O.ptr(lens), pcm_ptrs, cap, 0, O.ptr(rets), n) == 0
        for i, (Fs, ch) in enumerate(cfgs):
            fs_i = 960 * Fs // 48000
            rp, rr, rret = refs[i]
            assert rets[i] == fs_i == rret[f], (i, f, int(rets[i]))
            assert np.array_equal(outs[i][:fs_i], rp[f * fs_i:(f + 1) * fs_i]), (i, f)
    # span entry points: mixed rates -> OPUS_BAD_ARG, nothing launched
    two = (C.c_void_p * 2)(hs[0], hs[1])
    offs = np.zeros(2, dtype=np.int64)
    lens = np.full(2, int(l[0]), dtype=np.int32)
    blob = np.ascontiguousarray(d[:int(l[0])])
    pcm = np.zeros((2 * 960, 2), dtype=np.int16)
    r2 = np.zeros(2, dtype=np.int32)
    assert L.opus_decode_span(two, 2, 1, O.ptr(blob), O.ptr(offs), O.ptr(lens), O.ptr(pcm), 960, O.ptr(r2)) == cb.OPUS_BAD_ARG
    offs[1] = -4
    two = (C.c_void_p * 2)(hs[0], hs[4])
    assert L.opus_decode_span(two, 1, 2, O.ptr(blob), O.ptr(offs), O.ptr(lens), O.ptr(pcm), 960, O.ptr(r2)) == cb.OPUS_BAD_ARG   # negative offset
    for h in hs:
        L.opus_decoder_destroy(C.c_void_p(h))


@pytest.mark.parametrize("cap", [961, 1001, 964])
def test_decode_odd_capacity_stereo(cap):
    """PCM capacities that are not a multiple of 8 (channel 1's staging row is then not 16-byte aligned): both lanes of a
    stereo pair must take the same de-emphasis path."""
    cb = _cb()
    n, F = 6, 12
    datas, offs, lens, refs = [], [], [], []
    base = 0
    for s in range(n):
        x = O.test_signal(960 * F, 2, 900 + s, ["music", "tone", "clicks"][s % 3])
        d, o, l, _ = O.encode_stream(x, 960, 96000, 2, vbr=1, cvbr=0)
        d, o = O.pack(d, o, l)
        refs.append(O.decode_stream(d, o, l, 960, 2))
        datas.append(d); offs.append(o + base); lens.append(l); base += len(d)
    data, offs, lens = np.concatenate(datas), np.concatenate(offs), np.concatenate(lens)
    dec = cb.DecoderBatch(n, 48000, 2)
    pcm, rets = dec.decode_span(data, offs, lens, F, cap)
    dec.close()
    assert (rets == 960).all()
    pcm = pcm.reshape(n, F, cap, 2)
    for s in range(n):
        rp = refs[s][0].reshape(F, 960, 2)
        assert np.array_equal(pcm[s, :, :960], rp), s
        assert not pcm[s, :, 960:].any()


def test_duplicate_state_in_large_batch_is_rejected():
    cb = _cb()
    L = cb.lib()
    n = 200
    dec = cb.DecoderBatch(n, 48000, 2)
    hs = (C.c_void_p * n)(*dec.handles)
    hs[150] = hs[3]
    offs = np.zeros(n, dtype=np.int64)
    lens = np.zeros(n, dtype=np.int32)
    pcm = np.zeros((n * 120, 2), dtype=np.int16)
    r = np.zeros(n, dtype=np.int32)
    assert L.opus_decode_span(hs, n, 1, None, O.ptr(offs), O.ptr(lens), O.ptr(pcm), 120, O.ptr(r)) == cb.OPUS_BAD_ARG
    dec.close()
    enc = cb.EncoderBatch(n, 48000, 2, bitrate=64000)
    hs = (C.c_void_p * n)(*enc.handles)
    hs[199] = hs[0]
    x = np.zeros((n * 120, 2), dtype=np.int16)
    out = np.zeros((n, 100), dtype=np.uint8)
    assert L.opus_encode_span(hs, n, 1, O.ptr(x), 120, O.ptr(out), 100, O.ptr(r)) == cb.OPUS_BAD_ARG
    enc.close()


def test_encode_span_rejects_sizes_the_stream_would_not_code():
    """opus_encode runs frame_size_select first (opus_encoder.c:807-826); the span calls take the size actually coded."""
    cb = _cb()
    L = cb.lib()
    n = 3
    enc = cb.EncoderBatch(n, 48000, 2, bitrate=64000)
    x = np.zeros((n * 5000, 2), dtype=np.int16)
    out = np.zeros((n, 1276), dtype=np.uint8)
    r = np.zeros(n, dtype=np.int32)
    for bad in (900, 1000, 1500, 5000, 100):
        assert L.opus_encode_span(enc.handles, n, 1, O.ptr(x), bad, O.ptr(out), 1276, O.ptr(r)) == cb.OPUS_BAD_ARG, bad
    # OPUS_SET_EXPERT_FRAME_DURATION(10 ms) on one stream: a 20 ms span is not what that stream codes
    assert L.opus_encoder_ctl(C.c_void_p(enc.handles[1]), 4040, C.c_int32(5003)) == 0
    assert L.opus_encode_span(enc.handles, n, 1, O.ptr(x), 960, O.ptr(out), 1276, O.ptr(r)) == cb.OPUS_BAD_ARG
    assert L.opus_encode_span(enc.handles, n, 1, O.ptr(x), 480, O.ptr(out), 1276, O.ptr(r)) == 0   # 10 ms is what every stream codes when given 480
    assert (r > 0).all()
    assert L.opus_encoder_ctl(C.c_void_p(enc.handles[1]), 4040, C.c_int32(5000)) == 0
    assert L.opus_encode_span(enc.handles, n, 1, O.ptr(x), 960, O.ptr(out), 1276, O.ptr(r)) == 0
    assert (r > 0).all()
    enc.close()
