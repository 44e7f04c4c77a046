"""The `.bit` packet container of the reference's command-line codec, and batch file coding on top of the C ABI.

Format (opus-fix/src/opus_demo.c:68-80 int_to_char / char_to_int, :651-670 read, :748-760 write), per packet:
    4 bytes  payload length, big-endian
    4 bytes  encoder final range (OPUS_GET_FINAL_RANGE after the frame), big-endian; 0 = "do not check"
    len bytes payload (one Opus packet, TOC first); length 0 = a lost packet (decoded as concealment, :763)
Host-side data plumbing only: the coding itself goes through opus_decode_span / opus_encode_span_ranges (CUDA); there is no
CPU codec here.  SURVEY.md §8f rank 4.
"""
import ctypes as C
import struct

import numpy as np

MAX_PAYLOAD = 1500   # opus_demo.c:58 MAX_PACKET


def read_bit(path):
    """-> (data uint8 [sum(lens)], offs int64 [F], lens int32 [F], ranges uint32 [F]).  A truncated tail ends the stream the way
    opus_demo does (it stops at the first short read)."""
    raw = np.fromfile(path, dtype=np.uint8)
    buf = raw.tobytes()
    pos, n = 0, len(buf)
    offs, lens, ranges = [], [], []
    while pos + 8 <= n:
        ln, rng = struct.unpack_from(">II", buf, pos)
        if ln > MAX_PAYLOAD or pos + 8 + ln > n:
            break
        offs.append(pos + 8)
        lens.append(ln)
        ranges.append(rng)
        pos += 8 + ln
    return raw, np.asarray(offs, dtype=np.int64), np.asarray(lens, dtype=np.int32), np.asarray(ranges, dtype=np.uint32)


def write_bit(path, data, offs, lens, ranges=None):
    """Packets data[offs[f] : offs[f]+lens[f]] (+ their encoder final ranges, default 0 = unchecked) -> `.bit` file."""
    data = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1)
    with open(path, "wb") as f:
        for k in range(len(lens)):
            ln = int(lens[k])
            f.write(struct.pack(">II", ln, int(ranges[k]) if ranges is not None else 0))
            f.write(data[int(offs[k]):int(offs[k]) + ln].tobytes())


def decode_files(paths, Fs=48000, channels=2, frame_size=960):
    """Decode a batch of `.bit` files in one span call (one stream per file; ragged lengths are padded with lost packets whose
    output is dropped).  frame_size = samples per packet at Fs (the files' packet duration).  -> list of int16 [samples, channels]."""
    from . import DecoderBatch
    streams = [read_bit(p) for p in paths]
    n = len(streams)
    F = max((len(s[2]) for s in streams), default=0)
    if n == 0 or F == 0:
        return [np.zeros((0, channels), dtype=np.int16) for _ in streams]
    blobs, offs, lens, base = [], np.zeros((n, F), dtype=np.int64), np.zeros((n, F), dtype=np.int32), 0
    for i, (d, o, l, _) in enumerate(streams):
        blobs.append(d)
        offs[i, :len(l)] = o + base
        lens[i, :len(l)] = l
        base += len(d)
    data = np.concatenate(blobs) if base else np.zeros(1, dtype=np.uint8)
    dec = DecoderBatch(n, Fs, channels)
    try:
        pcm, rets = dec.decode_span(data, offs.reshape(-1), lens.reshape(-1), F, frame_size)
    finally:
        dec.close()
    pcm = pcm.reshape(n, F, frame_size, channels)
    rets = rets.reshape(n, F)
    out = []
    for i, s in enumerate(streams):
        Fi = len(s[2])
        if (rets[i, :Fi] < 0).any():
            raise RuntimeError("%s: packet %d: error %d" % (paths[i], int(np.nonzero(rets[i, :Fi] < 0)[0][0]), int(rets[i, :Fi].min())))
        out.append(np.concatenate([pcm[i, f, :rets[i, f]] for f in range(Fi)]) if Fi else np.zeros((0, channels), dtype=np.int16))
    return out


def encode_files(pcms, paths, Fs=48000, channels=2, frame_size=960, **enc_settings):
    """Encode a batch of PCM arrays (int16 [samples, channels]) into `.bit` files in one span call.  Like opus_demo -e
    (src/opus_demo.c:672-690) the last partial frame is zero-padded, and an input that ends on a frame boundary gets one extra
    all-zero frame.  enc_settings go to EncoderBatch (bitrate, vbr, cvbr, complexity, application ...)."""
    from . import EncoderBatch, lib, _p
    n = len(pcms)
    Fs_each = [len(p) // frame_size + 1 for p in pcms]
    F = max(Fs_each)
    x = np.zeros((n, F * frame_size, channels), dtype=np.int16)
    for i, p in enumerate(pcms):
        x[i, :len(p)] = np.asarray(p, dtype=np.int16).reshape(-1, channels)
    enc = EncoderBatch(n, Fs, channels, **enc_settings)
    stride = 1276
    data = np.zeros((n * F, stride), dtype=np.uint8)
    lens = np.zeros(n * F, dtype=np.int32)
    ranges = np.zeros(n * F, dtype=np.uint32)
    try:
        rc = lib().opus_encode_span_ranges(enc.handles, n, F, _p(x), frame_size, _p(data), stride, _p(lens), _p(ranges))
        if rc != 0:
            raise RuntimeError("opus_encode_span_ranges: %d" % rc)
    finally:
        enc.close()
    lens = lens.reshape(n, F)
    ranges = ranges.reshape(n, F)
    for i in range(n):
        Fi = Fs_each[i]
        if (lens[i, :Fi] < 0).any():
            raise RuntimeError("stream %d: encoder error %d" % (i, int(lens[i, :Fi].min())))
        write_bit(paths[i], data, (i * F + np.arange(Fi)) * stride, lens[i, :Fi], ranges[i, :Fi])
    return [Fs_each[i] for i in range(n)]
