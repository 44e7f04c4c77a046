// opus_decoder_dev.cuh — the Opus decoder layer between the C API and the CELT frame stages.
//
// Restates the MODE_CELT_ONLY subset of opus-fix/src/opus_decoder.c:200-596 (opus_decode_frame) and
// :598-709 (opus_decode_native), plus state init / reset (:96-133, celt/celt_decoder.c:134-172,1188-1202),
// split along the pipeline: opus_parse_packet (stage A: TOC, framing, per-frame CELT parse -> IR) and
// opus_synth_packet (stage B: state latching, per-frame synthesis, gain, final range).
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
    CB_NOUNROLL for (int i = 0; i < 2 * CB_NB_EBANDS; i++) {
        st->oldEBands[i] = 0; st->backgroundLogE[i] = 0;
        st->oldLogE[i] = st->oldLogE2[i] = -28672;
    }
    CB_NOUNROLL for (int i = 0; i < 2 * CB_LPC_ORDER; i++) st->lpc[i] = 0;
    CB_NOUNROLL for (int i = 0; i < 2 * CB_DEC_MEM; i++) st->decode_mem[i] = 0;
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

CB_DEV int bandwidth_to_endband(int bw) {   // opus_decoder.c:431-450
    switch (bw) {
    case kBwNarrow: return 13;
    case kBwMedium:
    case kBwWide: return 17;
    case kBwSuperWide: return 19;
    default: return 21;
    }
}

// ---------------------------------------------------------------------------------------------------
// Stage A: one packet -> IR.  `cap` = PCM capacity for this packet (samples per channel at Fs), kmax = frame IR
// slots per packet.  *seed: fold/noise seed chained from frame to frame (previous frame's final rng).
// Nothing here reads or writes decoder state.
// ---------------------------------------------------------------------------------------------------
CB_DEV void opus_parse_packet(const uint8_t *data, int len, int cap, int Fs, int decode_fec, int kmax, unsigned *seed,
                              CbPacketIR &pk, CbFrameIR *fr, int16_t *Xarea, ParseScratch &ps, bool dry) {
    pk.count = 0; pk.lost = 0; pk.frame_size = 0; pk.mode = 0; pk.bandwidth = 0; pk.stream_channels = 0; pk.reserved = 0;
    if (decode_fec < 0 || decode_fec > 1) { pk.ret = OPUS_BAD_ARG_; return; }
    if ((decode_fec || len == 0 || data == nullptr) && cap % (Fs / 400) != 0) { pk.ret = OPUS_BAD_ARG_; return; }
    if (len == 0 || data == nullptr) { pk.ret = 0; pk.lost = 1; return; }
    if (len < 0) { pk.ret = OPUS_BAD_ARG_; return; }
    const int packet_mode = pkt_mode(data);
    const int packet_frame_size = pkt_samples_per_frame(data, Fs);
    int16_t size[48];
    int offset;
    uint8_t toc;
    const int count = pkt_parse(data, len, 0, &toc, size, &offset, nullptr);
    if (count < 0) { pk.ret = count; return; }
    if (packet_mode != CB_MODE_CELT_ONLY) { pk.ret = OPUS_UNIMPLEMENTED_; return; }   // scope edge: no SILK / hybrid
    if (decode_fec) { pk.ret = 0; pk.lost = 1; return; }   // CELT carries no in-band FEC (opus_decoder.c:655-657)
    if (count * packet_frame_size > cap) { pk.ret = OPUS_BUFFER_TOO_SMALL_; return; }
    if (count > kmax) { pk.ret = OPUS_BUFFER_TOO_SMALL_; return; }   // cannot happen when kmax = cap / (Fs/400)
    pk.ret = 0;
    pk.count = (int16_t)count;
    pk.frame_size = packet_frame_size;
    pk.mode = (int16_t)packet_mode;
    pk.bandwidth = (int16_t)pkt_bandwidth(data);
    pk.stream_channels = (int16_t)pkt_nb_channels(data);
    const int C = pk.stream_channels;
    const int end = bandwidth_to_endband(pk.bandwidth);
    const int N = packet_frame_size * (48000 / Fs);   // CELT frame length at 48 kHz
    int LM = 0;
    while ((kShortMdct << LM) != N && LM < kMaxLM) LM++;
    const uint8_t *p = data + offset;
    int xoff = 0;
    CB_NOUNROLL for (int i = 0; i < count; i++) {
        CbFrameIR &ir = fr[i];
        ir.x_off = xoff;
        if (size[i] <= 1) {
            // concealment frame: no symbols; the seed passes through (pitch PLC leaves st->rng alone)
            ir.flags = CB_IR_LOST;
            ir.len = size[i];
            ir.LM = (uint8_t)LM; ir.C = (uint8_t)C; ir.end = (uint8_t)end;
            ir.rng_final = 0; ir.seed_bands = *seed;
        } else {
            celt_parse_frame(p, size[i], LM, C, end, seed, ir, Xarea + xoff, ps, dry);
        }
        xoff += N * C;
        p += size[i];
    }
}

// ---------------------------------------------------------------------------------------------------
// Stage B: consume one packet's IR in stream order.
// ---------------------------------------------------------------------------------------------------

// What stage B tells stage C about one packet: which 48 kHz sample range of the packet's staging area holds signal
// that still has to be de-emphasised (PCM outside it was written by stage B itself: leading zero frames).
struct CbSigRange {
    int32_t begin, end;
};

// opus_decode_frame remainder for one frame (opus_decoder.c:246-252,265-272,452-596).  sig[c] points at this frame's
// position in the packet's staging area; *staged is set when the frame left signal there.
template <class TM>
CB_DEV int opus_synth_frame(TM tm, CbDecState *st, SynthScratch &S, const CbFrameIR &ir, int16_t *X, int16_t *pcm, int room,
                            int *const *sig, bool *staged) {
    const int F20 = st->Fs / 50, F10 = F20 >> 1, F5 = F10 >> 1, F2_5 = F5 >> 1;
    *staged = false;
    if (room < F2_5) return OPUS_BUFFER_TOO_SMALL_;
    int frame_size = imin(room, st->Fs / 25 * 3);
    const bool lost = (ir.flags & CB_IR_LOST) != 0;
    int audiosize, mode;
    if (!lost) {
        audiosize = st->frame_size;
        mode = st->mode;
    } else {
        frame_size = imin(frame_size, st->frame_size);
        audiosize = frame_size;
        mode = st->prev_mode;
        if (mode == 0) {
            CB_TEAM_FOR(i, audiosize * st->channels, tm) pcm[i] = 0;
            tm.sync();
            return audiosize;
        }
        // audiosize > 20 ms cannot happen: it is clamped to st->frame_size and CELT frames are <= 20 ms
        if (audiosize < F20) {
            if (audiosize > F10) audiosize = F10;
            else if (mode != CB_MODE_SILK_ONLY && audiosize > F5 && audiosize < F10) audiosize = F5;
        }
    }
    if (audiosize > frame_size) return OPUS_BAD_ARG_;
    frame_size = audiosize;
    int celt_ret;
    if (lost) celt_ret = celt_decode_lost_frame(tm, st, S, sig, imin(F20, frame_size));
    else celt_ret = celt_synth_frame(tm, st, S, ir, X, sig);
    *staged = !(lost && celt_ret < 0);
    // decode gain (opus_decoder.c:567-577) is applied by stage C when it writes the PCM
    if (tm.lane() == 0) {
        st->rangeFinal = (lost || ir.len <= 1) ? 0 : ir.rng_final;
        st->prev_mode = mode;
        st->prev_redundancy = 0;
    }
    tm.sync();
    return celt_ret < 0 ? celt_ret : audiosize;
}

// opus_decode_native remainder for one packet (opus_decoder.c:613-627,682-708).  Returns what opus_decode returns.
// sigbase: the packet's staging area, channel c at sigbase + c*cap48 (cap48 = cap * downsample samples).
template <class TM>
CB_DEV int opus_synth_packet(TM tm, CbDecState *st, SynthScratch &S, const CbPacketIR &pk, const CbFrameIR *fr, int16_t *Xarea,
                             int16_t *pcm, int cap, int *sigbase, CbSigRange *range) {
    const int ds = st->downsample;
    const int cap48 = cap * ds;
    int sb = 0, se = 0;   // staged range in 48 kHz samples
    bool any = false;
    int result;
    if (pk.ret < 0) {
        result = pk.ret;
    } else if (pk.lost) {
        // conceal `cap` samples, frame by frame
        CbFrameIR lostir;
        lostir.flags = CB_IR_LOST; lostir.len = 0; lostir.rng_final = 0;
        int pcm_count = 0;
        result = 0;
        do {
            int *sig[2] = {sigbase + pcm_count * ds, sigbase + cap48 + pcm_count * ds};
            bool staged;
            int ret = opus_synth_frame(tm, st, S, lostir, nullptr, pcm + pcm_count * st->channels, cap - pcm_count, sig, &staged);
            if (ret < 0) { result = ret; break; }
            if (staged) { if (!any) sb = pcm_count * ds; any = true; se = (pcm_count + ret) * ds; }
            pcm_count += ret;
        } while (pcm_count < cap);
        if (result >= 0) {
            if (tm.lane() == 0) st->last_packet_duration = pcm_count;
            tm.sync();
            result = pcm_count;
        }
    } else {
        if (tm.lane() == 0) {
            st->mode = pk.mode;
            st->bandwidth = pk.bandwidth;
            st->frame_size = pk.frame_size;
            st->stream_channels = pk.stream_channels;
        }
        tm.sync();
        int nb_samples = 0;
        result = 0;
        CB_NOUNROLL for (int i = 0; i < pk.count; i++) {
            int *sig[2] = {sigbase + nb_samples * ds, sigbase + cap48 + nb_samples * ds};
            bool staged;
            int ret = opus_synth_frame(tm, st, S, fr[i], Xarea + fr[i].x_off, pcm + nb_samples * st->channels, cap - nb_samples, sig,
                                       &staged);
            if (staged) { if (!any) sb = nb_samples * ds; any = true; se = (nb_samples + (ret < 0 ? pk.frame_size : ret)) * ds; }
            if (ret < 0) { result = ret; break; }
            nb_samples += ret;
        }
        if (result >= 0) {
            if (tm.lane() == 0) st->last_packet_duration = nb_samples;
            tm.sync();
            result = nb_samples;
        }
    }
    if (tm.lane() == 0) { range->begin = any ? sb : 0; range->end = any ? se : 0; }
    return result;
}

// Stage C for one packet and one channel: de-emphasis + decode gain + PCM store (celt_decoder.c:185-275).
CB_DEV void opus_deemph_packet(CbDecState *st, int c, const int *sigbase, const CbSigRange &rg, int16_t *pcm, int cap) {
    if (rg.end <= rg.begin) return;
    const int ds = st->downsample, CC = st->channels;
    const int gain = st->decode_gain ? celt_exp2(s16(mul16_16_p15(21771, st->decode_gain))) : 0;   // QCONST16(6.48814081e-4f, 25)
    const int *x = sigbase + c * (cap * ds) + rg.begin;
    int16_t *y = pcm + (rg.begin / ds) * CC + c;
    st->preemph_memD[c] = deemphasis_channel(x, rg.end - rg.begin, y, CC, ds, st->preemph_memD[c], gain);
}

}  // namespace cb
