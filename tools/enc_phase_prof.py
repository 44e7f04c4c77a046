"""Dev tool: per-phase latency of stream 0 in the encoder kernel (needs a -DCB_PHASE_PROF build, CB200_LIB=...)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import concentus_b200 as cb
import oracle_lib as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2072
F = int(sys.argv[2]) if len(sys.argv) > 2 else 20
fs, ch = 960, 2
L = cb.lib()
assert L.opus_b200_init(0) == 0
base = [O.test_signal(fs * F, ch, 500 + i, ("music", "tone", "clicks", "music")[i % 4]) for i in range(64)]
pcm = np.concatenate([base[i % 64] for i in range(n)])
enc = cb.EncoderBatch(n, 48000, ch, bitrate=96000, vbr=1, cvbr=0, complexity=10)
cyc = (C.c_longlong * 64)()
wait = (C.c_longlong * 64)()
enc.encode_span(pcm, F, fs)
L.opus_b200_enc_phase_prof(cyc, wait, 1)
enc.encode_span(pcm, F, fs)
L.opus_b200_enc_phase_prof(cyc, wait, 1)
tot = sum(cyc[:30]) + sum(wait[:30])
print("n=%d F=%d: stream 0, cycles per frame by phase slot (work | barrier wait), total %.0f cycles/frame" % (n, F, tot / F))
for k in range(30):
    print("slot %2d: work %8.0f (%4.1f %%)  wait %8.0f (%4.1f %%)" % (k, cyc[k] / F, 100.0 * cyc[k] / tot, wait[k] / F, 100.0 * wait[k] / tot))
enc.close()
