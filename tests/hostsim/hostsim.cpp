// tests/hostsim/hostsim.cpp — TEST TOOL, never part of the product library.
//
// Compiles the codec headers of concentus_b200/csrc with a plain C++ compiler (team width 1, see
// celt_simt.cuh) so the integer semantics of the kernels can be diffed against the oracle on a
// machine with no GPU.  The CUDA build of the very same headers is what ships; this file exists
// only so `pytest -m "not gpu"` can localise a bit mismatch before GPU time is spent.
#include <cstdlib>
#include <cstring>
#include "../../concentus_b200/csrc/opus_decoder_dev.cuh"

extern "C" {

int hostsim_dec_state_size(void) { return (int)sizeof(CbDecState); }

// Decode F packets of one stream (packed layout).  cap = pcm capacity per packet (samples per channel).
int hostsim_decode_stream(const uint8_t *data, const int64_t *offs, const int32_t *lens, int F, int cap, int channels,
                          int Fs, int16_t *pcm, uint32_t *ranges, int32_t *rets) {
    CbDecState *st = (CbDecState *)calloc(1, sizeof(CbDecState));
    cb::DecScratch *S = (cb::DecScratch *)calloc(1, sizeof(cb::DecScratch));
    if (cb::dec_state_init(st, Fs, channels) != 0) return -1;
    cb::Team tm{0};
    for (int f = 0; f < F; f++) {
        const uint8_t *p = lens[f] > 0 ? data + offs[f] : nullptr;
        int r = cb::opus_decode_packet(tm, st, *S, p, lens[f], pcm + (size_t)f * cap * channels, cap, 0);
        if (rets) rets[f] = r;
        if (ranges) ranges[f] = st->rangeFinal;
    }
    free(S);
    free(st);
    return 0;
}
}
