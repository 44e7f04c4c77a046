#!/usr/bin/env python
"""SASS listing per kernel of the built objects (cuobjdump -sass), instruction text only (encodings stripped), with an opcode
histogram in front.  usage: python tools/sass_listing.py outdir tag"""
import collections
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
outdir, tag = sys.argv[1], sys.argv[2]
KEEP = ["parse_kernel", "synth_kernel", "deemph_kernel", "pipe_walk_kernelILi14ELb1", "pipe_transform_kernel", "pipe_prep_kernel",
        "pipe_decide_kernelILi2", "pipe_transient2_kernel", "pipe_comb_kernel", "pipe_head_kernelILi2", "pipe_prepass2_kernel",
        "pipe_fe1_kernel", "pipe_fe2_kernel"]
for obj in ("opus_capi.o", "opus_enc_pipe.o"):
    txt = subprocess.run(["cuobjdump", "-sass", os.path.join(HERE, "concentus_b200", "build", obj)], capture_output=True, text=True).stdout
    parts = re.split(r"\n\s*Function : ", txt)
    for part in parts[1:]:
        name = part.split("\n", 1)[0].strip()
        key = [k for k in KEEP if k in name]
        if not key:
            continue
        ins = re.findall(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);\s*/\*", part)
        hist = collections.Counter(re.sub(r"^@!?U?P\d\s+", "", i[1]).split()[0].split(".")[0] for i in ins)
        short = key[0].replace("ILi14ELb1", "_14warps").replace("ILi2", "_2perwarp")
        with open(os.path.join(outdir, "%s_sass_%s.txt" % (tag, short)), "w") as f:
            f.write("# %s\n# %d SASS instructions = %d bytes (sm_100a, cuobjdump -sass, encodings stripped)\n" % (name, len(ins), 16 * len(ins)))
            f.write("# opcode histogram: " + ", ".join("%s %d" % kv for kv in hist.most_common(40)) + "\n")
            f.write("# tensor-core / TMA opcodes (UTCMMA, UTMALDG, UBLKCP ...): %d — none expected: the path has no contraction (DESIGN.md 3)\n\n"
                    % sum(v for k, v in hist.items() if k.startswith(("UTC", "UTMA", "UBLKCP", "HMMA", "IMMA"))))
            for a, t in ins:
                f.write("%s  %s\n" % (a, t))
        print(short, len(ins))
