// tests/hostsim/hostsim.cpp — TEST TOOL, never part of the product library.
//
// Compiles the codec headers of concentus_b200/csrc with a plain C++ compiler (1-lane teams, see
// celt_simt.cuh) so the integer semantics of both pipeline stages can be diffed against the oracle on a
// machine with no GPU.  The CUDA build of the very same headers is what ships; this file exists
// only so `pytest -m "not gpu"` can localise a bit mismatch before GPU time is spent.
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../concentus_b200/csrc/opus_decoder_dev.cuh"

extern "C" {

int hostsim_dec_state_size(void) { return (int)sizeof(CbDecState); }

// Decode F packets of one stream (packed layout).  cap = pcm capacity per packet (samples per channel).
// Stage A (parse -> IR), stage B (synth) and stage C (de-emphasis) per packet, exactly the hand-offs the kernels use.
int hostsim_decode_stream(const uint8_t *data, const int64_t *offs, const int32_t *lens, int F, int cap, int channels,
                          int Fs, int16_t *pcm, uint32_t *ranges, int32_t *rets) {
    CbDecState *st = (CbDecState *)calloc(1, sizeof(CbDecState));
    cb::SynthScratch *S = (cb::SynthScratch *)calloc(1, sizeof(cb::SynthScratch));
    cb::ParseScratch *ps = (cb::ParseScratch *)calloc(1, sizeof(cb::ParseScratch));
    if (cb::dec_state_init(st, Fs, channels) != 0) return -1;
    const int kmax = cap / (Fs / 400) < 48 ? cap / (Fs / 400) : 48;
    std::vector<CbFrameIR> fr(kmax > 0 ? kmax : 1);
    std::vector<int16_t> X((size_t)cap * (48000 / Fs) * 2 + 16);
    std::vector<int> sig((size_t)cap * (48000 / Fs) * 2 + 16);
    cb::SoloTeam tm;
    for (int f = 0; f < F; f++) {
        const uint8_t *p = lens[f] > 0 ? data + offs[f] : nullptr;
        CbPacketIR pk;
        unsigned seed = st->rng;   // chained by stage A in the kernels; equal to st->rng here because B(f-1) has run
        // dry pass first (what a run's first thread does to recover its seed): must leave the same final range behind
        unsigned dry_seed = 12345u;
        CbPacketIR dry_pk;
        cb::opus_parse_packet(p, lens[f], cap, Fs, 0, kmax, &dry_seed, dry_pk, fr.data(), X.data(), *ps, true);
        cb::opus_parse_packet(p, lens[f], cap, Fs, 0, kmax, &seed, pk, fr.data(), X.data(), *ps, false);
        if (pk.ret >= 0 && !pk.lost && pk.count > 0 && !(fr[pk.count - 1].flags & CB_IR_LOST) && dry_seed != seed) {
            if (rets) rets[f] = -99;   // dry/full divergence: flag loudly
            continue;
        }
        cb::CbSigRange rg;
        int r = cb::opus_synth_packet(tm, st, *S, pk, fr.data(), X.data(), pcm + (size_t)f * cap * channels, cap, sig.data(), &rg);
        for (int c = 0; c < channels; c++) cb::opus_deemph_packet(st, c, sig.data(), rg, pcm + (size_t)f * cap * channels, cap);   // stage C
        if (rets) rets[f] = r;
        if (ranges) ranges[f] = st->rangeFinal;
    }
    free(ps);
    free(S);
    free(st);
    return 0;
}
}

// ---- encoder ----------------------------------------------------------------------------------------------------
#include "../../concentus_b200/csrc/opus_encoder_dev.cuh"

extern "C" {

int hostsim_enc_state_size(void) { return (int)sizeof(CbEncState); }

// Encode F frames of one stream with the 1-lane team.  cfg = {application, bitrate, vbr, cvbr, complexity, max_bytes,
// force_channels, bandwidth} (the oracle harness's ref_enc_cfg).  out: F slots of `stride` bytes.
int hostsim_encode_stream(const int16_t *pcm, int F, int frame_size, int channels, int Fs, const int *cfg, uint8_t *out, int stride,
                          int32_t *lens, uint32_t *ranges) {
    CbEncState *st = (CbEncState *)calloc(1, sizeof(CbEncState));
    cb::EncShared *S = (cb::EncShared *)calloc(1, sizeof(cb::EncShared));
    cb::EncGlobal *G = (cb::EncGlobal *)calloc(1, sizeof(cb::EncGlobal));
    if (cb::enc_state_init(st, Fs, channels, cfg[0]) != 0) return -1;
    int dummy = 0;
    cb::enc_ctl(st, 4002, cfg[1], &dummy);
    cb::enc_ctl(st, 4006, cfg[2], &dummy);
    cb::enc_ctl(st, 4020, cfg[3], &dummy);
    cb::enc_ctl(st, 4010, cfg[4], &dummy);
    if (cfg[6]) cb::enc_ctl(st, 4022, cfg[6], &dummy);
    if (cfg[7]) cb::enc_ctl(st, 4008, cfg[7], &dummy);
    cb::SoloTeam tm;
    int rc = 0;
    for (int f = 0; f < F; f++) {
        int n = cb::opus_encode_frame(tm, st, st, *S, *G, pcm + (size_t)f * frame_size * channels, frame_size, out + (size_t)f * stride,
                                      cfg[5] < stride ? cfg[5] : stride);
        lens[f] = n;
        if (n < 0) { rc = n; break; }
        if (ranges) ranges[f] = st->rangeFinal;
    }
    free(G);
    free(S);
    free(st);
    return rc;
}
}
