"""ctypes bindings to the oracle (oracle/_ref/libopus_ref.so = UNMODIFIED opus-fix build + oracle/ref_harness.c).

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs; nothing under concentus_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libopus_ref.so")

OPUS_APPLICATION_VOIP = 2048
OPUS_APPLICATION_AUDIO = 2049
OPUS_APPLICATION_RESTRICTED_LOWDELAY = 2051


class RefEncCfg(C.Structure):
    _fields_ = [(n, C.c_int) for n in
                ("application", "bitrate", "vbr", "cvbr", "complexity", "max_bytes", "force_channels", "bandwidth")]


def build_ref():
    """(Re)build oracle/_ref from /root/reference when the sources are present (dev container only)."""
    if os.path.isdir("/root/reference/opus-fix"):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "-j8"], check=True,
                       stdout=subprocess.DEVNULL)
    return os.path.exists(REF_SO)


_lib = None


def ref():
    global _lib
    if _lib is None:
        if not os.path.exists(REF_SO):
            build_ref()
        lib = C.CDLL(REF_SO)
        lib.ref_generate_music.argtypes = [C.c_void_p, C.c_int32, C.c_uint32]
        lib.ref_encode_stream.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(RefEncCfg),
                                          C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        lib.ref_decode_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ref_decode_stream_gain.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 7 + [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ref_decode_stream_info.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 5 + [C.c_void_p] * 4
        lib.ref_decode_streams_mt.restype = C.c_double
        lib.ref_decode_streams_mt.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ref_encode_stream_script.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(RefEncCfg), C.c_void_p, C.c_int,
                                                 C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        lib.ref_encode_streams_mt.restype = C.c_double
        lib.ref_encode_streams_mt.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                              C.POINTER(RefEncCfg), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        _lib = lib
    return _lib


def ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def generate_music(n_samples, seed=13371337):
    """Stereo int16 [n_samples, 2] — the signal of opus-fix/tests/test_opus_encode.c:59-90."""
    buf = np.zeros((n_samples, 2), dtype=np.int16)
    ref().ref_generate_music(ptr(buf), n_samples, seed & 0xFFFFFFFF)
    return buf


def test_signal(n_samples, channels, seed=13371337, kind="music"):
    """Deterministic test signals: 'music' (generate_music), 'noise' (full-scale), 'tone' (pitched, drives the
    post-filter), 'clicks' (music + impulses, drives transients)."""
    rng = np.random.RandomState(seed & 0x7FFFFFFF)
    if kind == "music":
        x = generate_music(n_samples, seed)
    elif kind == "noise":
        x = rng.randint(-32768, 32768, size=(n_samples, 2)).astype(np.int16)
    elif kind == "tone":
        t = np.arange(n_samples)
        f0 = 180.0 + (seed % 7) * 37.0
        s = sum(np.sin(2 * np.pi * f0 * k * t / 48000.0) / k for k in range(1, 9))
        env = 0.6 + 0.4 * np.sin(2 * np.pi * t / 48000.0 * 0.7)
        y = (s * env * 6000).astype(np.int32)
        x = np.stack([y, (y * 0.8).astype(np.int32)], axis=1)
        x = (x + rng.randint(-40, 40, size=x.shape)).clip(-32768, 32767).astype(np.int16)
    elif kind == "clicks":
        x = generate_music(n_samples, seed).astype(np.int32) // 2
        pos = rng.randint(0, n_samples, size=max(1, n_samples // 9000))
        for p in pos:
            seg = x[p:p + 64]
            seg += (rng.randint(-20000, 20000, size=seg.shape))
        x = x.clip(-32768, 32767).astype(np.int16)
    else:
        raise ValueError(kind)
    if channels == 1:
        x = x[:, :1].copy()
    return np.ascontiguousarray(x)


def encode_stream(pcm, frame_size, bitrate, channels=None, Fs=48000, vbr=1, cvbr=0, complexity=10,
                  application=OPUS_APPLICATION_RESTRICTED_LOWDELAY, max_bytes=1275, stride=1276):
    """Encode with the reference.  Returns (data uint8 [F*stride], offs int64 [F], lens int32 [F], ranges uint32 [F])."""
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    if channels is None:
        channels = pcm.shape[1]
    F = pcm.shape[0] // frame_size
    out = np.zeros(F * stride, dtype=np.uint8)
    lens = np.zeros(F, dtype=np.int32)
    ranges = np.zeros(F, dtype=np.uint32)
    cfg = RefEncCfg(application, bitrate, vbr, cvbr, complexity, max_bytes, 0, 0)
    rc = ref().ref_encode_stream(ptr(pcm), F, frame_size, channels, Fs, C.byref(cfg), ptr(out), stride, ptr(lens), ptr(ranges))
    assert rc == 0, rc
    offs = (np.arange(F, dtype=np.int64) * stride)
    return out, offs, lens, ranges


def pack(data, offs, lens):
    """Re-pack strided packets contiguously -> (data, offs[F])."""
    F = len(lens)
    noffs = np.zeros(F, dtype=np.int64)
    noffs[1:] = np.cumsum(lens[:-1].clip(min=0))
    total = int(noffs[-1] + max(int(lens[-1]), 0)) if F else 0
    out = np.zeros(max(total, 1), dtype=np.uint8)
    for f in range(F):
        n = int(lens[f])
        if n > 0:
            out[noffs[f]:noffs[f] + n] = data[offs[f]:offs[f] + n]
    return out, noffs


def decode_stream(data, offs, lens, frame_size, channels, Fs=48000):
    """Decode with the reference.  Returns (pcm int16 [F*frame_size, channels], ranges, rets)."""
    F = len(lens)
    pcm = np.zeros((F * frame_size, channels), dtype=np.int16)
    ranges = np.zeros(F, dtype=np.uint32)
    rets = np.zeros(F, dtype=np.int32)
    ref().ref_decode_stream(ptr(data), ptr(np.ascontiguousarray(offs, dtype=np.int64)),
                            ptr(np.ascontiguousarray(lens, dtype=np.int32)), F, frame_size, channels, Fs,
                            ptr(pcm), ptr(ranges), ptr(rets))
    return pcm, ranges, rets


def ctl_script(seed, F, channels, K=2, allow_mode=False):
    """Random encoder ctl changes, K slots before every frame: int32 [F, K, 2] of (request, value); request 0 = nothing.
    The shape of the reference's own encoder fuzz loop (opus-fix/tests/test_opus_encode.c:236-330)."""
    rs = np.random.RandomState(seed)
    script = np.zeros((F, K, 2), dtype=np.int32)
    for f in range(F):
        for k in range(K):
            if rs.rand() >= 0.25:
                continue
            c = rs.randint(12)
            if c == 0:
                script[f, k] = (4002, int(rs.choice([6000, 12000, 24000, 32000, 48000, 64000, 96000, 128000, 256000, 510000, -1000, -1])))
            elif c == 1:
                script[f, k] = (4006, rs.randint(2))
            elif c == 2:
                script[f, k] = (4020, rs.randint(2))
            elif c == 3:
                script[f, k] = (4010, rs.randint(11))
            elif c == 4:
                script[f, k] = (4022, int(rs.choice([1, 2, -1000])) if channels == 2 else -1000)
            elif c == 5:
                script[f, k] = (4008, int(rs.choice([-1000, 1101, 1102, 1103, 1104, 1105])))
            elif c == 6:
                script[f, k] = (4004, int(rs.choice([1101, 1102, 1103, 1104, 1105])))
            elif c == 7:
                script[f, k] = (4014, rs.randint(0, 30))
            elif c == 8:
                script[f, k] = (4036, rs.randint(8, 25))
            elif c == 9:
                script[f, k] = (4042, rs.randint(2))
            elif c == 10:
                script[f, k] = (4024, int(rs.choice([-1000, 3001, 3002])))
            elif rs.rand() < 0.2:
                script[f, k] = (4028, 0)
    return script


def encode_stream_script(pcm, frame_size, channels, script, Fs=48000, application=OPUS_APPLICATION_RESTRICTED_LOWDELAY, stride=1276):
    """Reference encode with a ctl script.  Returns (packets [F, stride], lens, ranges)."""
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    F, K = script.shape[0], script.shape[1]
    out = np.zeros((F, stride), dtype=np.uint8)
    lens = np.zeros(F, dtype=np.int32)
    ranges = np.zeros(F, dtype=np.uint32)
    cfg = RefEncCfg(application, 64000, 1, 1, 10, stride, 0, 0)
    ref().ref_encode_stream_script(ptr(pcm), F, frame_size, channels, Fs, C.byref(cfg), ptr(np.ascontiguousarray(script)), K, ptr(out), stride,
                                   ptr(lens), ptr(ranges))
    return out, lens, ranges
