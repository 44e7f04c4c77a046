"""GPU parity: CUDA encode (through the C ABI of libconcentus_b200.so) vs the oracle (unmodified opus-fix build): packets
byte-for-byte, packet length per frame, final range after the last frame — the same comparison CSharp/ParityTest makes
(reference CSharp/ParityTest/TestDriver.cs) and tests/test_opus_encode.c:306 (encoder/decoder final-range agreement)."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu


def _cb():
    import concentus_b200 as cb
    assert cb.lib().opus_b200_init(0) == 0, "CUDA device required: no CPU fallback exists"
    return cb


def _ref_encode(pcm, fs, br, ch, vbr, cvbr, cx, application=O.OPUS_APPLICATION_RESTRICTED_LOWDELAY):
    d, o, l, r = O.encode_stream(pcm, fs, br, ch, vbr=vbr, cvbr=cvbr, complexity=cx, application=application, max_bytes=1276)
    return d.reshape(-1, 1276), l, r


def _check_group(signals, ch, fs, br, vbr, cvbr, cx, nsec=1, spans=2):
    """One batch = streams that share every encoder setting; signals = list of (kind, seed)."""
    cb = _cb()
    n = len(signals)
    pcms = [O.test_signal(48000 * nsec, ch, seed, kind) for (kind, seed) in signals]
    F = pcms[0].shape[0] // fs
    enc = cb.EncoderBatch(n, 48000, ch, bitrate=br, vbr=vbr, cvbr=cvbr, complexity=cx)
    allp = np.stack([p[:F * fs].reshape(F, fs * ch) for p in pcms])   # [n, F, fs*ch]
    data = np.zeros((n, F, 1276), dtype=np.uint8)
    lens = np.zeros((n, F), dtype=np.int32)
    cuts = np.linspace(0, F, spans + 1).astype(int)
    for a, b in zip(cuts[:-1], cuts[1:]):
        if b == a:
            continue
        d, l = enc.encode_span(allp[:, a:b].reshape(-1, ch), b - a, fs)
        data[:, a:b] = d.reshape(n, b - a, 1276)
        lens[:, a:b] = l.reshape(n, b - a)
    fr = enc.final_ranges()
    enc.close()
    for s in range(n):
        rd, rl, rr = _ref_encode(pcms[s], fs, br, ch, vbr, cvbr, cx)
        tag = (signals[s], ch, fs, br, vbr, cvbr, cx)
        assert np.array_equal(rl, lens[s]), ("packet lengths", tag, int(np.nonzero(rl != lens[s])[0][0]))
        for f in range(F):
            assert np.array_equal(rd[f, :rl[f]], data[s, f, :rl[f]]), ("packet bytes", tag, "frame", f)
        assert int(rr[-1]) == int(fr[s]), ("final range", tag)


SIGS = [("music", 11), ("tone", 12), ("clicks", 13), ("noise", 14)]


@pytest.fixture(params=["pipeline", "one_kernel"])
def enc_path(request):
    """The encoder's two device paths (include/opus_b200.h): the frame-synchronous kernel pipeline and the one-kernel path."""
    cb = _cb()
    L = cb.lib()
    prev = L.opus_b200_enc_set_pipeline(1 if request.param == "pipeline" else 0)
    assert prev >= 0
    p0, l0 = C.c_longlong(0), C.c_longlong(0)
    L.opus_b200_enc_path_counts(C.byref(p0), C.byref(l0))
    yield request.param
    p1, l1 = C.c_longlong(0), C.c_longlong(0)
    L.opus_b200_enc_path_counts(C.byref(p1), C.byref(l1))
    L.opus_b200_enc_set_pipeline(1)
    if request.param == "pipeline":
        assert p1.value > p0.value, "the pipeline did not take the streams of this test"
    else:
        assert p1.value == p0.value and l1.value > l0.value


@pytest.mark.parametrize("fs", [120, 240, 480, 960])
@pytest.mark.parametrize("ch", [1, 2])
def test_encode_frame_sizes_vbr(fs, ch, enc_path):
    _check_group(SIGS, ch, fs, 96000, 1, 0, 10)


@pytest.mark.parametrize("br", [32000, 48000, 64000, 128000, 192000, 256000, 510000])
def test_encode_bitrates_cbr_stereo(br, enc_path):
    _check_group(SIGS, 2, 960, br, 0, 0, 10)


@pytest.mark.parametrize("br", [32000, 64000, 256000])
@pytest.mark.parametrize("fs", [120, 480])
def test_encode_cvbr_mono(br, fs, enc_path):
    _check_group(SIGS, 1, fs, br, 1, 1, 10)


@pytest.mark.parametrize("cx", [0, 2, 4, 5, 8])
def test_encode_complexities(cx, enc_path):
    _check_group(SIGS, 2, 960, 64000, 1, 1, cx)
    _check_group(SIGS[:2], 2, 240, 128000, 1, 0, cx)


def test_encode_config2_shape_long():
    """BASELINE configs[2] shape (48 kHz stereo 96 kbps complexity 10, VBR and CBR), long enough for the VBR controller to settle
    (vbr_count saturates at 970 frames) and for the pre-filter / transient paths to be exercised."""
    _check_group([("music", 21), ("tone", 22), ("clicks", 23)], 2, 960, 96000, 1, 0, 10, nsec=22, spans=3)
    _check_group([("music", 24), ("clicks", 25)], 2, 960, 96000, 0, 0, 10, nsec=4, spans=1)


def test_encode_many_streams_one_launch():
    """More streams than one wave of warps: 300 streams x 10 frames, every stream a different signal."""
    sigs = [(("music", "tone", "clicks", "noise")[i % 4], 1000 + i) for i in range(300)]
    cb = _cb()
    ch, fs, F = 2, 960, 10
    pcms = [O.test_signal(fs * F, ch, seed, kind) for (kind, seed) in sigs]
    enc = cb.EncoderBatch(len(sigs), 48000, ch, bitrate=96000, vbr=1, cvbr=0, complexity=10)
    d, l = enc.encode_span(np.concatenate(pcms), F, fs)
    enc.close()
    d = d.reshape(len(sigs), F, 1276)
    l = l.reshape(len(sigs), F)
    for s in range(len(sigs)):
        rd, rl, _ = _ref_encode(pcms[s], fs, 96000, ch, 1, 0, 10)
        assert np.array_equal(rl, l[s]), ("len", sigs[s])
        for f in range(F):
            assert np.array_equal(rd[f, :rl[f]], d[s, f, :rl[f]]), ("bytes", sigs[s], f)


def test_encode_pipelined_host_span():
    """A host-buffer span big enough (>= 32 MB of PCM) to be cut into sub-spans whose PCM upload / packet download overlap the
    coding of their neighbours: 320 streams x 30 frames in ONE call, 40 distinct signals; every stream against the oracle."""
    cb = _cb()
    ch, fs, F, n = 2, 960, 30, 320
    kinds = ("music", "tone", "clicks", "noise")
    base = [O.test_signal(fs * F, ch, 4000 + i, kinds[i % 4]) for i in range(40)]
    refs = [_ref_encode(b, fs, 96000, ch, 1, 0, 10) for b in base]
    enc = cb.EncoderBatch(n, 48000, ch, bitrate=96000, vbr=1, cvbr=0, complexity=10)
    d, l = enc.encode_span(np.concatenate([base[s % 40] for s in range(n)]), F, fs)
    fr = enc.final_ranges()
    enc.close()
    d = d.reshape(n, F, 1276)
    l = l.reshape(n, F)
    for s in range(n):
        rd, rl, rr = refs[s % 40]
        assert np.array_equal(rl, l[s]), ("len", s)
        for f in range(F):
            assert np.array_equal(rd[f, :rl[f]], d[s, f, :rl[f]]), ("bytes", s, f)
        assert int(rr[-1]) == int(fr[s]), ("final range", s)


def test_scalar_api_and_state_copy():
    """opus_encode one frame at a time, a memcpy'd state block continues identically (tests/test_opus_encode.c:198,214),
    OPUS_RESET_STATE restarts the stream, ctl argument checks (tests/test_opus_api.c)."""
    cb = _cb()
    L = cb.lib()
    ch, fs = 2, 960
    pcm = O.test_signal(48000, ch, 77, "tone")
    F = pcm.shape[0] // fs
    rd, rl, rr = _ref_encode(pcm, fs, 64000, ch, 1, 1, 10)
    err = C.c_int(0)
    h = L.opus_encoder_create(48000, ch, cb.OPUS_APPLICATION_RESTRICTED_LOWDELAY, C.byref(err))
    assert h and err.value == 0
    hp = C.c_void_p(h)
    assert L.opus_encoder_ctl(hp, cb.OPUS_SET_BITRATE_REQUEST, C.c_int32(64000)) == 0
    assert L.opus_encoder_ctl(hp, cb.OPUS_SET_COMPLEXITY_REQUEST, C.c_int32(10)) == 0
    assert L.opus_encoder_ctl(hp, cb.OPUS_SET_COMPLEXITY_REQUEST, C.c_int32(11)) == cb.OPUS_BAD_ARG
    assert L.opus_encoder_ctl(hp, 999999, C.c_int32(0)) == cb.OPUS_UNIMPLEMENTED
    # the multistream-only switches of the CELT layer (opus_encoder.c:2455-2469): their defaults are accepted, nothing else
    assert L.opus_encoder_ctl(hp, 10024, C.c_int32(0)) == 0                       # OPUS_SET_LFE(0)
    assert L.opus_encoder_ctl(hp, 10024, C.c_int32(1)) == cb.OPUS_UNIMPLEMENTED
    assert L.opus_encoder_ctl(hp, 10026, C.c_void_p(None)) == 0                   # OPUS_SET_ENERGY_MASK(NULL)
    out = np.zeros(1276, dtype=np.uint8)
    half = F // 2
    for f in range(half):
        n = L.opus_encode(hp, O.ptr(pcm[f * fs:(f + 1) * fs]), fs, O.ptr(out), 1276)
        assert n == rl[f] and np.array_equal(out[:n], rd[f, :n]), f
        v = C.c_uint32(0)
        L.opus_encoder_ctl(hp, cb.OPUS_GET_FINAL_RANGE_REQUEST, C.byref(v))
        assert v.value == int(rr[f])
    # clone the state block with memcpy, destroy the original, continue on the clone
    size = L.opus_encoder_get_size(ch)
    clone = C.create_string_buffer(size)
    C.memmove(clone, h, size)
    L.opus_encoder_destroy(hp)
    cp = C.cast(clone, C.c_void_p)
    for f in range(half, F):
        n = L.opus_encode(cp, O.ptr(pcm[f * fs:(f + 1) * fs]), fs, O.ptr(out), 1276)
        assert n == rl[f] and np.array_equal(out[:n], rd[f, :n]), f
    # reset: the stream starts over
    assert L.opus_encoder_ctl(cp, cb.OPUS_RESET_STATE) == 0
    n = L.opus_encode(cp, O.ptr(pcm[:fs]), fs, O.ptr(out), 1276)
    assert n == rl[0] and np.array_equal(out[:n], rd[0, :n])
    # argument errors
    assert L.opus_encode(cp, O.ptr(pcm[:fs]), 100, O.ptr(out), 1276) == cb.OPUS_BAD_ARG
    assert L.opus_encode(cp, O.ptr(pcm[:fs]), fs, O.ptr(out), 0) == cb.OPUS_BAD_ARG
    assert L.opus_encoder_create(44100, 2, cb.OPUS_APPLICATION_AUDIO, C.byref(err)) is None and err.value == cb.OPUS_BAD_ARG


def test_encode_decode_roundtrip_on_device():
    """Our encoder's packets through our decoder: final ranges agree frame by frame (tests/test_opus_encode.c:306) and the PCM
    equals what the oracle decodes from the oracle's packets."""
    cb = _cb()
    ch, fs, F = 2, 960, 50
    pcm = O.test_signal(fs * F, ch, 5, "music")
    enc = cb.EncoderBatch(1, 48000, ch, bitrate=128000, vbr=1, cvbr=0, complexity=10)
    d, l = enc.encode_span(pcm, F, fs)
    enc.close()
    offs = (np.arange(F, dtype=np.int64) * 1276)
    dec = cb.DecoderBatch(1, 48000, ch)
    out, rets = dec.decode_span(d.reshape(-1), offs, l, F, fs)
    dec.close()
    assert (rets == fs).all()
    rd, rl, _ = _ref_encode(pcm, fs, 128000, ch, 1, 0, 10)
    rp, _, _ = O.decode_stream(rd.reshape(-1), offs, rl, fs, ch)
    assert np.array_equal(rp, out)


def test_audio_application_forced_celt_and_unimplemented():
    """OPUS_APPLICATION_AUDIO with OPUS_SET_FORCE_MODE(MODE_CELT_ONLY): 4 ms of look-ahead compensation through the delay buffer.
    Without forcing, a low-rate frame the reference would hand to SILK returns OPUS_UNIMPLEMENTED and leaves the state alone."""
    cb = _cb()
    ch, fs, F = 2, 960, 25
    pcm = O.test_signal(fs * F, ch, 9, "music")
    enc = cb.EncoderBatch(1, 48000, ch, application=cb.OPUS_APPLICATION_AUDIO, bitrate=96000, vbr=1, cvbr=1, complexity=10)
    d, l = enc.encode_span(pcm, F, fs)
    enc.close()
    rd, rl, _ = _ref_encode(pcm, fs, 96000, ch, 1, 1, 10, application=O.OPUS_APPLICATION_AUDIO)
    # the reference picks CELT-only for music-like input at this rate (TOC bit 7)
    if (rd[:, 0] & 0x80).all():
        assert np.array_equal(rl, l)
        for f in range(F):
            assert np.array_equal(rd[f, :rl[f]], d[f, :rl[f]]), f
    enc = cb.EncoderBatch(1, 48000, 1, application=cb.OPUS_APPLICATION_AUDIO, bitrate=12000, vbr=1, cvbr=1, complexity=10)
    d, l = enc.encode_span(pcm[:, :1].copy(), F, fs)
    enc.close()
    assert (l == cb.OPUS_UNIMPLEMENTED).all()


def test_mixed_settings_in_one_launch():
    """BASELINE configs[4] in miniature: one span launch over streams that differ in bitrate, CBR/VBR/CVBR and complexity
    (per-stream settings live in each stream's own state block; only Fs, channel count and frame size are shared)."""
    cb = _cb()
    L = cb.lib()
    ch, fs, F = 2, 960, 25
    rs = np.random.RandomState(3)
    rates = [32000, 48000, 64000, 96000, 128000, 192000, 256000, 510000]
    cfgs = [(rates[rs.randint(8)], [(0, 0), (1, 0), (1, 1)][rs.randint(3)], [0, 5, 10][rs.randint(3)]) for _ in range(48)]
    pcms = [O.test_signal(fs * F, ch, 9000 + i, ("music", "tone", "clicks", "noise")[i % 4]) for i in range(len(cfgs))]
    enc = cb.EncoderBatch(len(cfgs), 48000, ch)
    for i, (br, (vbr, cvbr), cx) in enumerate(cfgs):
        hp = C.c_void_p(enc.handles[i])
        for req, v in ((cb.OPUS_SET_BITRATE_REQUEST, br), (cb.OPUS_SET_VBR_REQUEST, vbr), (cb.OPUS_SET_VBR_CONSTRAINT_REQUEST, cvbr),
                       (cb.OPUS_SET_COMPLEXITY_REQUEST, cx)):
            assert L.opus_encoder_ctl(hp, req, C.c_int32(v)) == 0
    d, l = enc.encode_span(np.concatenate(pcms), F, fs)
    enc.close()
    d = d.reshape(len(cfgs), F, 1276)
    l = l.reshape(len(cfgs), F)
    # decode everything in one launch too and compare with the oracle's decode of the oracle's packets
    offs = np.arange(len(cfgs) * F, dtype=np.int64) * 1276
    dec = cb.DecoderBatch(len(cfgs), 48000, ch)
    out, rets = dec.decode_span(d.reshape(-1), offs, l.reshape(-1), F, fs)
    dec.close()
    assert (rets == fs).all()
    out = out.reshape(len(cfgs), F * fs, ch)
    for i, (br, (vbr, cvbr), cx) in enumerate(cfgs):
        rd, rl, _ = _ref_encode(pcms[i], fs, br, ch, vbr, cvbr, cx)
        assert np.array_equal(rl, l[i]), ("len", i, cfgs[i])
        for f in range(F):
            assert np.array_equal(rd[f, :rl[f]], d[i, f, :rl[f]]), ("bytes", i, cfgs[i], f)
        rp, _, _ = O.decode_stream(rd.reshape(-1), np.arange(F, dtype=np.int64) * 1276, rl, fs, ch)
        assert np.array_equal(rp, out[i]), ("pcm", i, cfgs[i])


@pytest.mark.parametrize("Fs", [8000, 12000, 16000, 24000])
def test_encode_other_api_rates(Fs, enc_path):
    """API rates below 48 kHz: zero-stuffing pre-emphasis (celt_encoder.c:490-533), MDCT bound
    (:451-460), bandwidth capped at the input's Nyquist rate; both applications, every frame size."""
    cb = _cb()
    L = cb.lib()
    k = 0
    for ch in (1, 2):
        for ms, br, vbr, cvbr, app in ((20, 64000, 1, 0, cb.OPUS_APPLICATION_RESTRICTED_LOWDELAY), (10, 32000, 0, 0, cb.OPUS_APPLICATION_RESTRICTED_LOWDELAY),
                                       (5, 96000, 1, 1, cb.OPUS_APPLICATION_AUDIO), (2.5, 128000, 1, 0, cb.OPUS_APPLICATION_RESTRICTED_LOWDELAY)):
            fs = int(Fs * ms / 1000)
            pcms = [O.test_signal(Fs, ch, 600 + k + i, kind) for i, kind in enumerate(("music", "tone", "clicks"))]
            k += 3
            F = pcms[0].shape[0] // fs
            enc = cb.EncoderBatch(len(pcms), Fs, ch, application=app, bitrate=br, vbr=vbr, cvbr=cvbr, complexity=10)
            d, l = enc.encode_span(np.concatenate([p[:F * fs] for p in pcms]), F, fs)
            enc.close()
            d = d.reshape(len(pcms), F, 1276)
            l = l.reshape(len(pcms), F)
            for s_, x in enumerate(pcms):
                rd, ro, rl, _ = O.encode_stream(x, fs, br, ch, Fs=Fs, vbr=vbr, cvbr=cvbr, complexity=10, application=app, max_bytes=1276)
                rd = rd.reshape(-1, 1276)
                if not (rd[:, 0] & 0x80).all():
                    continue                      # the reference chose SILK/hybrid for this one: outside the engine
                assert np.array_equal(rl, l[s_]), ("len", Fs, ch, ms, br, s_)
                for f in range(F):
                    assert np.array_equal(rd[f, :rl[f]], d[s_, f, :rl[f]]), ("bytes", Fs, ch, ms, br, s_, f)


@pytest.mark.parametrize("ms", [40, 60])
def test_encode_long_frames_repacketized(ms):
    """40 / 60 ms frames: coded as 2 / 3 forced-configuration 20 ms frames and merged by the
    repacketizer into code 1 / 2 / 3 packets, CBR ones padded (opus_encoder.c:1362-1438, repacketizer.c:102-227)."""
    cb = _cb()
    k = 0
    for Fs, ch, br, vbr, cvbr, maxb in ((48000, 2, 96000, 1, 0, 1276), (48000, 2, 64000, 0, 0, 1276), (48000, 1, 32000, 1, 1, 1276),
                                        (48000, 2, 510000, 1, 0, 1276), (48000, 2, 128000, 1, 0, 400), (16000, 2, 48000, 0, 0, 1276)):
        fs = Fs * ms // 1000
        pcms = [O.test_signal(Fs * 2, ch, 800 + k + i, kind) for i, kind in enumerate(("music", "tone", "clicks"))]
        k += 3
        F = pcms[0].shape[0] // fs
        enc = cb.EncoderBatch(len(pcms), Fs, ch, bitrate=br, vbr=vbr, cvbr=cvbr, complexity=10)
        d, l = enc.encode_span(np.concatenate([p[:F * fs] for p in pcms]), F, fs, max_data_bytes=maxb)
        fr = enc.final_ranges()
        enc.close()
        d = d.reshape(len(pcms), F, maxb)
        l = l.reshape(len(pcms), F)
        for s_, x in enumerate(pcms):
            rd, ro, rl, rr = O.encode_stream(x, fs, br, ch, Fs=Fs, vbr=vbr, cvbr=cvbr, complexity=10, max_bytes=maxb)
            assert np.array_equal(rl, l[s_]), ("len", ms, Fs, ch, br, s_)
            for f in range(F):
                assert np.array_equal(rd[ro[f]:ro[f] + rl[f]], d[s_, f, :rl[f]]), ("bytes", ms, Fs, ch, br, s_, f)
            assert int(rr[-1]) == int(fr[s_])
        # and our decoder takes the multi-frame packets back (codes 1-3)
        dec = cb.DecoderBatch(1, Fs, ch)
        offs = np.arange(F, dtype=np.int64) * maxb
        out, rets = dec.decode_span(d[0].reshape(-1), offs, l[0], F, fs)
        dec.close()
        assert (rets == fs).all()
        rp, _, _ = O.decode_stream(d[0].reshape(-1), offs, l[0], fs, ch, Fs=Fs)
        assert np.array_equal(rp, out)


@pytest.mark.parametrize("seed", [1, 2, 6, 11])
def test_scalar_api_ctl_fuzz(seed):
    """opus_encoder_ctl between opus_encode calls (state round-trips host <-> HBM every frame): random setting changes, the
    reference's own fuzz shape (tests/test_opus_encode.c:236-330)."""
    cb = _cb()
    L = cb.lib()
    ch = 1 + seed % 2
    fs = (120, 240, 480, 960, 1920, 2880)[seed % 6]
    F = 50 if fs <= 960 else 20
    x = O.test_signal(fs * F, ch, seed, ("music", "tone", "clicks", "noise")[seed % 4])
    script = O.ctl_script(seed, F, ch)
    rd, rl, rr = O.encode_stream_script(x, fs, ch, script)
    err = C.c_int(0)
    h = C.c_void_p(L.opus_encoder_create(48000, ch, cb.OPUS_APPLICATION_RESTRICTED_LOWDELAY, C.byref(err)))
    for req, v in ((4002, 64000), (4006, 1), (4020, 1), (4010, 10)):
        assert L.opus_encoder_ctl(h, req, C.c_int32(v)) == 0
    out = np.zeros(1276, dtype=np.uint8)
    v = C.c_uint32(0)
    for f in range(F):
        for k in range(script.shape[1]):
            req, val = int(script[f, k, 0]), int(script[f, k, 1])
            if req == 4028:
                L.opus_encoder_ctl(h, req)
            elif req:
                L.opus_encoder_ctl(h, req, C.c_int32(val))
        n = L.opus_encode(h, O.ptr(x[f * fs:(f + 1) * fs]), fs, O.ptr(out), 1276)
        assert n == rl[f], (seed, f, n, int(rl[f]))
        assert np.array_equal(out[:n], rd[f, :n]), (seed, f)
        L.opus_encoder_ctl(h, cb.OPUS_GET_FINAL_RANGE_REQUEST, C.byref(v))
        assert v.value == int(rr[f]), (seed, f)
    L.opus_encoder_destroy(h)


def test_pipeline_groups_chunks_and_mixed_paths():
    """The frame-synchronous pipeline at a size that uses several stream groups (>= 512 streams) and several chunks of frames
    (front end of chunk k+1 overlapping the frame steps of chunk k, buffers reused from chunk k+2 on), in the same launch as
    streams only the one-kernel path takes (OPUS_APPLICATION_AUDIO forced to CELT); two calls, state carried across."""
    cb = _cb()
    L = cb.lib()
    ch, fs, F, n = 2, 960, 44, 640
    kinds = ("music", "tone", "clicks", "noise")
    nb = 16
    base = [O.test_signal(fs * F, ch, 7000 + i, kinds[i % 4]) for i in range(nb)]
    apps = [cb.OPUS_APPLICATION_AUDIO if (s % 7) == 3 else cb.OPUS_APPLICATION_RESTRICTED_LOWDELAY for s in range(n)]
    refs = {}
    for b in range(nb):
        for app in (cb.OPUS_APPLICATION_AUDIO, cb.OPUS_APPLICATION_RESTRICTED_LOWDELAY):
            d, o, l, r = O.encode_stream(base[b], fs, 96000, ch, vbr=1, cvbr=(b % 2), complexity=10, application=app, max_bytes=1276)
            refs[(b, app)] = (d.reshape(-1, 1276), l, r)
    err = C.c_int(0)
    hs = []
    for s in range(n):
        h = L.opus_encoder_create(48000, ch, apps[s], C.byref(err))
        for req, val in ((cb.OPUS_SET_BITRATE_REQUEST, 96000), (cb.OPUS_SET_VBR_REQUEST, 1), (cb.OPUS_SET_VBR_CONSTRAINT_REQUEST, (s % nb) % 2),
                         (cb.OPUS_SET_COMPLEXITY_REQUEST, 10)):
            assert L.opus_encoder_ctl(C.c_void_p(h), req, C.c_int32(val)) == 0
        if apps[s] == cb.OPUS_APPLICATION_AUDIO:
            assert L.opus_encoder_ctl(C.c_void_p(h), cb.OPUS_SET_FORCE_MODE_REQUEST, C.c_int32(1002)) == 0
        hs.append(h)
    hs = (C.c_void_p * n)(*hs)
    p0, l0 = C.c_longlong(0), C.c_longlong(0)
    L.opus_b200_enc_path_counts(C.byref(p0), C.byref(l0))
    allp = np.stack([base[s % nb].reshape(F, fs * ch) for s in range(n)])
    data = np.zeros((n, F, 1276), dtype=np.uint8)
    lens = np.zeros((n, F), dtype=np.int32)
    for a, b in ((0, 37), (37, 44)):
        x = np.ascontiguousarray(allp[:, a:b]).reshape(-1)
        d = np.zeros((n * (b - a), 1276), dtype=np.uint8)
        l = np.zeros(n * (b - a), dtype=np.int32)
        assert L.opus_encode_span(hs, n, b - a, O.ptr(x), fs, O.ptr(d), 1276, O.ptr(l)) == 0
        data[:, a:b] = d.reshape(n, b - a, 1276)
        lens[:, a:b] = l.reshape(n, b - a)
    p1, l1 = C.c_longlong(0), C.c_longlong(0)
    L.opus_b200_enc_path_counts(C.byref(p1), C.byref(l1))
    n_audio = sum(1 for a in apps if a == cb.OPUS_APPLICATION_AUDIO)
    assert p1.value - p0.value == 2 * (n - n_audio) and l1.value - l0.value == 2 * n_audio
    v = C.c_uint32(0)
    for s in range(n):
        rd, rl, rr = refs[(s % nb, apps[s])]
        assert np.array_equal(rl, lens[s]), ("len", s, int(np.nonzero(rl != lens[s])[0][0]))
        for f in range(F):
            assert np.array_equal(rd[f, :rl[f]], data[s, f, :rl[f]]), ("bytes", s, f)
        if s % 41 == 0:
            L.opus_encoder_ctl(C.c_void_p(hs[s]), cb.OPUS_GET_FINAL_RANGE_REQUEST, C.byref(v))
            assert v.value == int(rr[-1]), s
    for h in hs:
        L.opus_encoder_destroy(C.c_void_p(h))


def _encode_mixed_batch(cb, pcms, ch, fs, Fs, settings, application=None):
    """One span launch over streams with per-stream (bitrate, vbr, cvbr, complexity); returns (data [n, F, 1276], lens [n, F], final ranges)."""
    L = cb.lib()
    n = len(pcms)
    F = pcms[0].shape[0] // fs
    enc = cb.EncoderBatch(n, Fs, ch, application=application or cb.OPUS_APPLICATION_RESTRICTED_LOWDELAY)
    for i, (br, vbr, cvbr, cx) in enumerate(settings):
        hp = C.c_void_p(enc.handles[i])
        for req, v in ((cb.OPUS_SET_BITRATE_REQUEST, br), (cb.OPUS_SET_VBR_REQUEST, vbr), (cb.OPUS_SET_VBR_CONSTRAINT_REQUEST, cvbr),
                       (cb.OPUS_SET_COMPLEXITY_REQUEST, cx)):
            assert L.opus_encoder_ctl(hp, req, C.c_int32(v)) == 0
    d, l = enc.encode_span(np.concatenate([p[:F * fs] for p in pcms]), F, fs)
    fr = enc.final_ranges()
    enc.close()
    return d.reshape(n, F, 1276), l.reshape(n, F), fr


@pytest.mark.parametrize("fs", [120, 240, 480, 960])
@pytest.mark.parametrize("ch", [1, 2])
def test_encode_sweep_matrix(fs, ch, enc_path):
    """BASELINE configs[3], encoder side: frame size x {mono, stereo} x {32, 64, 128, 256, 510 kbps} x {CBR, VBR, CVBR}: the 15
    rate / mode combinations of a (frame size, channels) cell are one span launch (per-stream settings), every stream a different
    signal, packets byte-for-byte against the oracle."""
    cb = _cb()
    settings = [(br, vbr, cvbr, 10) for br in (32000, 64000, 128000, 256000, 510000) for (vbr, cvbr) in ((0, 0), (1, 0), (1, 1))]
    kinds = ("music", "tone", "clicks", "noise")
    pcms = [O.test_signal(48000, ch, 5000 + 17 * fs + i, kinds[i % 4]) for i in range(len(settings))]
    d, l, fr = _encode_mixed_batch(cb, pcms, ch, fs, 48000, settings)
    F = l.shape[1]
    for i, (br, vbr, cvbr, cx) in enumerate(settings):
        rd, rl, rr = _ref_encode(pcms[i], fs, br, ch, vbr, cvbr, cx)
        assert np.array_equal(rl, l[i]), ("len", fs, ch, settings[i], int(np.nonzero(rl != l[i])[0][0]))
        for f in range(F):
            assert np.array_equal(rd[f, :rl[f]], d[i, f, :rl[f]]), ("bytes", fs, ch, settings[i], f)
        assert int(rr[-1]) == int(fr[i]), ("final range", fs, ch, settings[i])


def _raw_fixture(Fs, ch, seconds):
    import os
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    name = "%dKhz_%s.raw" % (Fs // 1000, "Stereo" if ch == 2 else "Mono")
    p = os.path.join(root, name) if (Fs == 48000 and ch == 2) else os.path.join(root, "raw", name)
    x = np.fromfile(p, dtype="<i2").reshape(-1, ch)
    n = Fs * seconds
    return np.ascontiguousarray(np.resize(x, (n, ch)) if len(x) < n else x[:n])


@pytest.mark.parametrize("fs", [120, 240, 480, 960])
def test_encode_long_streams_from_the_reference_fixture(fs, enc_path):
    """>= 20 s per stream for every frame size (the VBR controller's vbr_count saturates at 970 frames, the energy histories and
    the CVBR reservoir settle well after one second), on the reference's own parity input 48Khz Stereo.raw (CSharp/ParityTest)."""
    cb = _cb()
    x = _raw_fixture(48000, 2, 20)
    pcms = [x, np.ascontiguousarray(np.roll(x, 48000 * 3, axis=0))]
    settings = [(96000, 1, 0, 10), (64000, 1, 1, 10)]
    d, l, fr = _encode_mixed_batch(cb, pcms, 2, fs, 48000, settings)
    for i, (br, vbr, cvbr, cx) in enumerate(settings):
        rd, rl, rr = _ref_encode(pcms[i], fs, br, 2, vbr, cvbr, cx)
        assert np.array_equal(rl, l[i]), ("len", fs, settings[i], int(np.nonzero(rl != l[i])[0][0]))
        bad = [f for f in range(len(rl)) if not np.array_equal(rd[f, :rl[f]], d[i, f, :rl[f]])]
        assert not bad, ("bytes", fs, settings[i], bad[0])
        assert int(rr[-1]) == int(fr[i])


@pytest.mark.parametrize("Fs", [8000, 12000, 16000, 24000, 48000])
@pytest.mark.parametrize("ch", [1, 2])
def test_raw_fixtures_every_api_rate(Fs, ch):
    """The reference's parity inputs at every API rate (Java/ConcentusTestConsole/.../AudioData/<rate>Khz {Mono,Stereo}.raw, the files
    CSharp/ParityTest feeds both codecs): encode 5 s (20 ms frames, VBR and CBR), packets against the oracle, then decode the
    packets at the same rate, PCM against the oracle."""
    cb = _cb()
    x = _raw_fixture(Fs, ch, 5)
    fs = Fs // 50
    for br, vbr in ((64000, 1), (32000, 0)):
        d, l, fr = _encode_mixed_batch(cb, [x], ch, fs, Fs, [(br, vbr, 0, 10)])
        rd, ro, rl, rr = O.encode_stream(x, fs, br, ch, Fs=Fs, vbr=vbr, cvbr=0, complexity=10, max_bytes=1276)
        rd = rd.reshape(-1, 1276)
        assert np.array_equal(rl, l[0]), ("len", Fs, ch, br)
        for f in range(len(rl)):
            assert np.array_equal(rd[f, :rl[f]], d[0, f, :rl[f]]), ("bytes", Fs, ch, br, f)
        F = len(rl)
        offs = np.arange(F, dtype=np.int64) * 1276
        dec = cb.DecoderBatch(1, Fs, ch)
        pcm, rets = dec.decode_span(rd.reshape(-1), offs, rl, F, fs)
        dec.close()
        rp, _, rret = O.decode_stream(rd.reshape(-1), offs, rl, fs, ch, Fs=Fs)
        assert np.array_equal(rets, rret) and np.array_equal(pcm, rp), ("decode", Fs, ch, br)
