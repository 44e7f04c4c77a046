// opus_decoder_dev.cuh — the Opus decoder layer that sits between the C API and the CELT frame decoder,
// executed by the team that owns the stream (so a whole span of packets is decoded without touching the host).
//
// Restates the MODE_CELT_ONLY subset of opus-fix/src/opus_decoder.c:200-596 (opus_decode_frame) and
// :598-709 (opus_decode_native), plus state init / reset (:96-133, celt/celt_decoder.c:134-172,1188-1202).
// SILK-only and hybrid packets are outside this engine (SURVEY.md §8b "scope edge"): they return
// OPUS_UNIMPLEMENTED and leave the state untouched — there is deliberately no CPU fallback.
#pragma once
#include "celt_decoder.cuh"
#include "celt_plc.cuh"
#include "opus_packet.h"

namespace cb {

// OPUS_RESET_STATE of the CELT decoder (celt_decoder.c:1188-1202) — host or device, scalar.
CB_HD void dec_state_reset_celt(CbDecState *st) {
    st->rng = 0; st->error = 0; st->last_pitch_index = 0; st->loss_count = 0;
    st->postfilter_period = st->postfilter_period_old = 0;
    st->postfilter_gain = st->postfilter_gain_old = 0;
    st->postfilter_tapset = st->postfilter_tapset_old = 0;
    st->preemph_memD[0] = st->preemph_memD[1] = 0;
    for (int i = 0; i < 2 * CB_NB_EBANDS; i++) {
        st->oldEBands[i] = 0; st->backgroundLogE[i] = 0;
        st->oldLogE[i] = st->oldLogE2[i] = -28672;
    }
    for (int i = 0; i < 2 * CB_LPC_ORDER; i++) st->lpc[i] = 0;
    for (int i = 0; i < 2 * CB_DEC_MEM; i++) st->decode_mem[i] = 0;
}
// OPUS_RESET_STATE of the Opus decoder (opus_decoder.c:873-887)
CB_HD void dec_state_reset(CbDecState *st) {
    st->bandwidth = 0; st->mode = 0; st->prev_mode = 0; st->prev_redundancy = 0;
    st->last_packet_duration = 0; st->rangeFinal = 0;
    dec_state_reset_celt(st);
    st->stream_channels = st->channels;
    st->frame_size = st->Fs / 400;
}
// opus_decoder_init (opus_decoder.c:96-133)
CB_HD int dec_state_init(CbDecState *st, int Fs, int channels) {
    if ((Fs != 48000 && Fs != 24000 && Fs != 16000 && Fs != 12000 && Fs != 8000) || (channels != 1 && channels != 2)) return -1;
    st->channels = channels;
    st->Fs = Fs;
    st->downsample = 48000 / Fs;
    st->decode_gain = 0;
    dec_state_reset(st);
    return 0;
}

// opus_decode_frame, CELT-only (opus_decoder.c:200-596).  data == nullptr / len <= 1 => concealment.
CB_DEV int opus_decode_frame(Team tm, CbDecState *st, DecScratch &S, const uint8_t *data, int len, int16_t *pcm,
                             int frame_size, int decode_fec) {
    const int F20 = st->Fs / 50, F10 = F20 >> 1, F5 = F10 >> 1, F2_5 = F5 >> 1;
    if (frame_size < F2_5) return OPUS_BUFFER_TOO_SMALL_;
    frame_size = imin(frame_size, st->Fs / 25 * 3);
    if (len <= 1) {
        data = nullptr;
        frame_size = imin(frame_size, st->frame_size);
    }
    int audiosize, mode;
    EcDec dec;
    if (data != nullptr) {
        audiosize = st->frame_size;
        mode = st->mode;
        dec.init(data, (unsigned)len);
    } else {
        audiosize = frame_size;
        mode = st->prev_mode;
        if (mode == 0) {
            CB_TEAM_FOR(i, audiosize * st->channels, tm) pcm[i] = 0;
            CB_SYNC();
            return audiosize;
        }
        // audiosize > 20 ms cannot happen here: it is clamped to st->frame_size, and CELT frames are <= 20 ms
        // (the reference's chunking loop at opus_decoder.c:275-288 is only reachable after SILK packets).
        if (audiosize < F20) {
            if (audiosize > F10) audiosize = F10;
            else if (mode != CB_MODE_SILK_ONLY && audiosize > F5 && audiosize < F10) audiosize = F5;
        }
    }
    // transitions to/from SILK cannot occur: this engine only ever latches MODE_CELT_ONLY.
    if (audiosize > frame_size) return OPUS_BAD_ARG_;
    frame_size = audiosize;
    int endband = 21;
    switch (st->bandwidth) {
    case kBwNarrow: endband = 13; break;
    case kBwMedium:
    case kBwWide: endband = 17; break;
    case kBwSuperWide: endband = 19; break;
    case kBwFull: endband = 21; break;
    }
    const int C = st->stream_channels;
    const int celt_frame_size = imin(F20, frame_size);
    int celt_ret;
    if (data == nullptr || decode_fec) {
        celt_ret = celt_decode_lost_frame(tm, st, S, pcm, celt_frame_size, C, 0, endband);
    } else {
        celt_ret = celt_decode_frame(tm, st, S, data, len, pcm, celt_frame_size, C, 0, endband, dec);
    }
    if (st->decode_gain) {
        int gain = celt_exp2(s16(mul16_16_p15(21771, st->decode_gain)));   // QCONST16(6.48814081e-4f, 25)
        CB_TEAM_FOR(i, frame_size * st->channels, tm) {
            int x = mul16_32_p16(pcm[i], gain);
            pcm[i] = (int16_t)(x > 32767 ? 32767 : (x < -32767 ? -32767 : x));
        }
        CB_SYNC();
    }
    if (tm.lane == 0) {
        st->rangeFinal = len <= 1 ? 0 : dec.rng;
        st->prev_mode = mode;
        st->prev_redundancy = 0;
    }
    CB_SYNC();
    return celt_ret < 0 ? celt_ret : audiosize;
}

// Lost-packet branch of opus_decode_native (opus_decoder.c:613-627): conceal frame_size samples.
CB_DEV int opus_conceal(Team tm, CbDecState *st, DecScratch &S, int16_t *pcm, int frame_size) {
    int pcm_count = 0;
    do {
        int ret = opus_decode_frame(tm, st, S, nullptr, 0, pcm + pcm_count * st->channels, frame_size - pcm_count, 0);
        if (ret < 0) return ret;
        pcm_count += ret;
    } while (pcm_count < frame_size);
    if (tm.lane == 0) st->last_packet_duration = pcm_count;
    CB_SYNC();
    return pcm_count;
}

// opus_decode_native (opus_decoder.c:598-709).  frame_size = capacity of pcm in samples per channel.
CB_DEV int opus_decode_packet(Team tm, CbDecState *st, DecScratch &S, const uint8_t *data, int len, int16_t *pcm,
                              int frame_size, int decode_fec) {
    if (decode_fec < 0 || decode_fec > 1) return OPUS_BAD_ARG_;
    if ((decode_fec || len == 0 || data == nullptr) && frame_size % (st->Fs / 400) != 0) return OPUS_BAD_ARG_;
    if (len == 0 || data == nullptr) {
        return opus_conceal(tm, st, S, pcm, frame_size);
    } else if (len < 0) {
        return OPUS_BAD_ARG_;
    }
    const int packet_mode = pkt_mode(data);
    const int packet_bandwidth = pkt_bandwidth(data);
    const int packet_frame_size = pkt_samples_per_frame(data, st->Fs);
    const int packet_stream_channels = pkt_nb_channels(data);
    int16_t size[48];
    int offset;
    uint8_t toc;
    const int count = pkt_parse(data, len, 0, &toc, size, &offset, nullptr);
    if (count < 0) return count;
    if (packet_mode != CB_MODE_CELT_ONLY) return OPUS_UNIMPLEMENTED_;   // scope edge: no SILK / hybrid
    data += offset;
    if (decode_fec) {
        // CELT carries no in-band FEC: run concealment (opus_decoder.c:655-657)
        if (frame_size % (st->Fs / 400) != 0) return OPUS_BAD_ARG_;
        return opus_conceal(tm, st, S, pcm, frame_size);
    }
    if (count * packet_frame_size > frame_size) return OPUS_BUFFER_TOO_SMALL_;
    if (tm.lane == 0) {
        st->mode = packet_mode;
        st->bandwidth = packet_bandwidth;
        st->frame_size = packet_frame_size;
        st->stream_channels = packet_stream_channels;
    }
    CB_SYNC();
    int nb_samples = 0;
    for (int i = 0; i < count; i++) {
        int ret = opus_decode_frame(tm, st, S, data, size[i], pcm + nb_samples * st->channels, frame_size - nb_samples, 0);
        if (ret < 0) return ret;
        data += size[i];
        nb_samples += ret;
    }
    if (tm.lane == 0) st->last_packet_duration = nb_samples;
    CB_SYNC();
    return nb_samples;
}

}  // namespace cb
