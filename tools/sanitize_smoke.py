"""Dev tool: a tiny decode + encode workload for compute-sanitizer (memcheck / racecheck / initcheck / synccheck).
Covers transient + tonal + noisy signals, packet loss, two launches per direction; compares with the oracle.

usage (GPU box): compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import concentus_b200 as cb
import oracle_lib as O

L = cb.lib()
assert L.opus_b200_init(0) == 0
n, fs, ch = 6, 960, 2
F = int(sys.argv[1]) if len(sys.argv) > 1 else 8
kinds = ("music", "tone", "clicks", "noise", "music", "clicks")
pcms = [O.test_signal(fs * F, ch, 70 + i, kinds[i]) for i in range(n)]
enc = cb.EncoderBatch(n, 48000, ch, bitrate=96000, vbr=1, cvbr=0, complexity=10)
x = np.stack([p.reshape(F, fs * ch) for p in pcms])
d1, l1 = enc.encode_span(x[:, :F // 2].reshape(-1, ch), F // 2, fs)
d2, l2 = enc.encode_span(x[:, F // 2:].reshape(-1, ch), F - F // 2, fs)
enc.close()
d = np.concatenate([d1.reshape(n, F // 2, 1276), d2.reshape(n, F - F // 2, 1276)], axis=1)
l = np.concatenate([l1.reshape(n, F // 2), l2.reshape(n, F - F // 2)], axis=1)
bad = 0
refs = []
for i in range(n):
    rd, ro, rl, rr = O.encode_stream(pcms[i], fs, 96000, ch, vbr=1, cvbr=0, complexity=10, max_bytes=1276)
    refs.append((rd, ro, rl))
    rd2 = rd.reshape(F, 1276)
    bad += not (np.array_equal(rl, l[i]) and all(np.array_equal(rd2[f, :rl[f]], d[i, f, :rl[f]]) for f in range(F)))
lens = np.stack([r[2] for r in refs]).copy()
lens[1, 2] = 0
lens[2, 3:6] = 0
lens[3, 1] = 1
blob = np.concatenate([r[0] for r in refs])
offs = (np.arange(n * F, dtype=np.int64) * 1276).reshape(n, F)
dec = cb.DecoderBatch(n, 48000, ch)
p1, r1 = dec.decode_span(blob, offs[:, :F // 2].reshape(-1), lens[:, :F // 2].reshape(-1), F // 2, fs)
p2, r2 = dec.decode_span(blob, offs[:, F // 2:].reshape(-1), lens[:, F // 2:].reshape(-1), F - F // 2, fs)
dec.close()
pcm = np.concatenate([p1.reshape(n, F // 2, -1), p2.reshape(n, F - F // 2, -1)], axis=1)
for i in range(n):
    rp, rr, rret = O.decode_stream(blob, offs[i], lens[i], fs, ch)
    bad += not np.array_equal(rp.reshape(F, -1), pcm[i])
print("sanitize_smoke: %d mismatching streams of %d" % (bad, 2 * n))
sys.exit(1 if bad else 0)
