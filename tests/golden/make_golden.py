"""Generate tests/golden/*.npz from the oracle (oracle/_ref = unmodified opus-fix build).  Dev container only.

Each fixture: a short CELT stream produced by the reference ENCODER (packets, lens, encoder final ranges) plus the
reference DECODER's output (decoder final ranges, per-packet return codes, CRC32 of each decoded frame and the first
two frames of PCM verbatim).  Small on purpose (a few KB each)."""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O

CASES = [
    # name, kind, channels, frame_size, bitrate, vbr, cvbr, seconds
    ("music_st_20ms_64k_cbr", "music", 2, 960, 64000, 0, 0, 1.0),
    ("music_st_20ms_128k_vbr", "music", 2, 960, 128000, 1, 0, 1.0),
    ("tone_st_10ms_96k_cvbr", "tone", 2, 480, 96000, 1, 1, 0.5),
    ("clicks_mono_5ms_48k_vbr", "clicks", 1, 240, 48000, 1, 0, 0.5),
    ("noise_st_2p5ms_510k_cbr", "noise", 2, 120, 510000, 0, 0, 0.25),
    ("music_mono_20ms_32k_vbr", "music", 1, 960, 32000, 1, 0, 1.0),
    ("tone_st_2p5ms_32k_vbr_monocoded", "tone", 2, 120, 32000, 1, 0, 0.25),
]


def main():
    for (name, kind, ch, fs, br, vbr, cvbr, sec) in CASES:
        pcm = O.test_signal(int(48000 * sec), ch, 4242, kind)
        d, o, l, er = O.encode_stream(pcm, fs, br, vbr=vbr, cvbr=cvbr)
        d, o = O.pack(d, o, l)
        rp, rr, rret = O.decode_stream(d, o, l, fs, ch)
        F = len(l)
        crc = np.array([zlib.crc32(rp[f * fs:(f + 1) * fs].tobytes()) for f in range(F)], dtype=np.uint32)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), data=d, offs=o, lens=l, enc_ranges=er, dec_ranges=rr, rets=rret,
                            pcm_crc=crc, pcm_head=rp[:2 * fs], channels=ch, frame_size=fs, bitrate=br, vbr=vbr, cvbr=cvbr)
        print(name, F, "frames", len(d), "bytes")


if __name__ == "__main__":
    main()
