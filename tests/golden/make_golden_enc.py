"""Generate tests/golden/enc_*.npz from the oracle (oracle/_ref = unmodified opus-fix build).  Dev container only.

Each fixture: a short PCM input (verbatim), the encoder settings, and what the reference ENCODER produced from it —
packets, packet lengths, final range per frame.  The encoder tests replay the PCM through our encoder (host simulation on
CPU, CUDA on the GPU box) and require byte-identical packets."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O

CASES = [
    # name, kind, channels, frame_size, bitrate, vbr, cvbr, complexity, seconds
    ("enc_music_st_20ms_96k_vbr_cx10", "music", 2, 960, 96000, 1, 0, 10, 0.4),
    ("enc_clicks_st_20ms_96k_cbr_cx10", "clicks", 2, 960, 96000, 0, 0, 10, 0.4),
    ("enc_tone_st_10ms_64k_cvbr_cx10", "tone", 2, 480, 64000, 1, 1, 10, 0.3),
    ("enc_music_mono_5ms_48k_vbr_cx5", "music", 1, 240, 48000, 1, 0, 5, 0.2),
    ("enc_noise_st_2p5ms_256k_cbr_cx0", "noise", 2, 120, 256000, 0, 0, 0, 0.1),
]


def main():
    for (name, kind, ch, fs, br, vbr, cvbr, cx, sec) in CASES:
        pcm = O.test_signal(int(48000 * sec), ch, 777, kind)
        pcm = pcm[:(pcm.shape[0] // fs) * fs]
        d, o, l, er = O.encode_stream(pcm, fs, br, ch, vbr=vbr, cvbr=cvbr, complexity=cx)
        d, o = O.pack(d, o, l)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), pcm=pcm, data=d, offs=o, lens=l, enc_ranges=er, channels=ch, frame_size=fs,
                            bitrate=br, vbr=vbr, cvbr=cvbr, complexity=cx)
        print(name, len(l), "frames", len(d), "bytes")


if __name__ == "__main__":
    main()
