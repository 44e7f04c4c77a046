# oracle/celt_only_decode.sed — turns opus-fix/tests/test_opus_decode.c into the CELT-only test this engine is in scope for
# (SURVEY.md 8b): TOC configurations 16..31 only.  Applied at build time to the reference's source where it lies; the result goes to
# oracle/_ref/gen/ (git-ignored) and is compiled twice: against the reference library (the filtered test must still pass there,
# including the cres[] known-answer sums) and against libconcentus_b200.so.
# 1. loops over all 64 (configuration, stereo) prefixes -> the 32 CELT ones (packet[0] = i << 2, CELT <=> i >= 32)
s/for(i=0;i<64;i++)/for(i=32;i<64;i++)/
# 2. the SILK known-answer block (lmodes / lres): removed
/mode=fast_rand()%3;/,/lmodes\[mode\]);/d
# 3. de Bruijn sequence over all mode pairs -> over all CELT mode pairs
s/packet\[0\]=modes\[i\]<<2;/packet[0]=(32|(modes[i]\&31))<<2;/
# 4. the pre-selected packet (tmodes = 25 << 2 = TOC 0x64) is a HYBRID packet: removed
/int tmodes\[1\]={25<<2};/,/pre-selected random packets OK/d
