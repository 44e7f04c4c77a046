// opus_encoder_dev.cuh — the Opus layer around the CELT encoder: state init/reset, per-frame rate / channel / bandwidth
// decisions, DC rejection, low-rate stereo narrowing, TOC and CBR padding.
//
// Restates the MODE_CELT_ONLY path of opus-fix/src/opus_encoder.c: opus_encoder_init :164-252, gen_toc :254-284, dc_reject
// :362-385, stereo_fade :411-441, user_bitrate_to_bitrate :512-521, compute_stereo_width :861-936, opus_encode_native
// :938-2005, the OPUS_RESET_STATE ctl :2437-2458, and opus_packet_pad for a single frame (src/repacketizer.c:102-258).
// Scope: frames the reference would code with SILK or hybrid, VOIP's variable high-pass, API rates other
// than 48 kHz and frames longer than 20 ms return OPUS_UNIMPLEMENTED and leave the state untouched — there is no CPU path.
#pragma once
#include "celt_encoder.cuh"
#include "opus_packet.h"

#ifndef CB_HD
#if defined(__CUDACC__)
#define CB_HD __host__ __device__ inline
#else
#define CB_HD static inline
#endif
#endif

namespace cb {

enum { kAppVoip = 2048, kAppAudio = 2049, kAppLowdelay = 2051, kSignalVoice = 3001, kSignalMusic = 3002 };

// CELT OPUS_RESET_STATE (celt_encoder.c:2438-2457)
CB_HD void enc_state_reset_celt(CbEncState *st) {
    st->rng = 0;
    st->spread_decision = 2;   // SPREAD_NORMAL
    st->delayedIntra = 1;
    st->tonal_average = 256;
    st->lastCodedBands = 0; st->hf_average = 0; st->tapset_decision = 0;
    st->prefilter_period = 0; st->prefilter_gain = 0; st->prefilter_tapset = 0; st->consec_transient = 0;
    st->preemph_memE[0] = st->preemph_memE[1] = 0;
    st->vbr_reservoir = 0; st->vbr_drift = 0; st->vbr_offset = 0; st->vbr_count = 0;
    st->overlap_max = 0; st->stereo_saving = 0; st->intensity = 0; st->spec_avg = 0;
    CB_NOUNROLL for (int i = 0; i < 2 * CB_OVERLAP; i++) st->in_mem[i] = 0;
    CB_NOUNROLL for (int i = 0; i < 2 * CB_COMB_MAXPERIOD; i++) st->prefilter_mem[i] = 0;
    CB_NOUNROLL for (int i = 0; i < 2 * CB_NB_EBANDS; i++) { st->oldBandE[i] = 0; st->oldLogE[i] = st->oldLogE2[i] = -28672; }
}
// Opus OPUS_RESET_STATE (opus_encoder.c:2437-2458)
CB_HD void enc_state_reset(CbEncState *st) {
    st->stream_channels = st->channels;
    st->hybrid_stereo_width_Q14 = 1 << 14;
    CB_NOUNROLL for (int i = 0; i < 4; i++) st->hp_mem[i] = 0;
    st->mode = CB_MODE_HYBRID;
    st->prev_mode = 0; st->prev_channels = 0; st->prev_framesize = 0;
    st->bandwidth = 1105;
    st->first = 1;
    st->width_XX = st->width_XY = st->width_YY = 0; st->width_smoothed = 0; st->width_max_follower = 0;
    CB_NOUNROLL for (int i = 0; i < CB_ENC_DELAY_BUF; i++) st->delay_buffer[i] = 0;
    st->rangeFinal = 0;
    enc_state_reset_celt(st);
}
// opus_encoder_init (opus_encoder.c:164-252)
CB_HD int enc_state_init(CbEncState *st, int Fs, int channels, int application) {
    if ((Fs != 48000 && Fs != 24000 && Fs != 16000 && Fs != 12000 && Fs != 8000) || (channels != 1 && channels != 2) ||
        (application != kAppVoip && application != kAppAudio && application != kAppLowdelay))
        return -1;
    st->application = application; st->channels = channels; st->Fs = Fs;
    st->force_channels = -1000; st->signal_type = -1000; st->user_bandwidth = -1000; st->max_bandwidth = 1105;
    st->user_forced_mode = -1000;
    st->use_vbr = 1; st->vbr_constraint = 1; st->variable_duration = 5000; st->user_bitrate_bps = -1000; st->lsb_depth = 24;
    st->complexity = 9; st->packet_loss_perc = 0; st->prediction_disabled = 0; st->inband_fec = 0; st->dtx = 0;
    st->delay_compensation = Fs / 250; st->encoder_buffer = Fs / 100;
    st->voice_ratio = -1; st->bitrate_bps = 3000 + Fs * channels;
    st->upsample = 48000 / Fs; st->celt_force_intra = 0; st->celt_disable_pf = 0;
    enc_state_reset(st);
    return 0;
}


// opus_encoder_ctl (opus_encoder.c:2031-2507), integer-valued requests.  SET: `v` is the argument.  GET: *out receives the
// value.  Host-side code (the C ABI and the host simulation) — the state block must be host-current.
CB_HD int enc_ctl(CbEncState *st, int request, int v, int *out) {
    switch (request) {
    case 4000:   // OPUS_SET_APPLICATION
        if ((v != kAppVoip && v != kAppAudio && v != kAppLowdelay) || (!st->first && st->application != v)) return -1;
        st->application = v;
        return 0;
    case 4001: *out = st->application; return 0;
    case 4002:   // OPUS_SET_BITRATE
        if (v != -1000 && v != -1) {
            if (v <= 0) return -1;
            else if (v <= 500) v = 500;
            else if (v > 300000 * st->channels) v = 300000 * st->channels;
        }
        st->user_bitrate_bps = v;
        return 0;
    case 4003: {   // OPUS_GET_BITRATE: user_bitrate_to_bitrate(st, st->prev_framesize, 1276)
        int fs = st->prev_framesize ? st->prev_framesize : st->Fs / 400;
        if (st->user_bitrate_bps == -1000) *out = 60 * st->Fs / fs + st->Fs * st->channels;
        else if (st->user_bitrate_bps == -1) *out = 1276 * 8 * st->Fs / fs;
        else *out = st->user_bitrate_bps;
        return 0;
    }
    case 4022:   // OPUS_SET_FORCE_CHANNELS
        if ((v < 1 || v > st->channels) && v != -1000) return -1;
        st->force_channels = v;
        return 0;
    case 4023: *out = st->force_channels; return 0;
    case 4004:   // OPUS_SET_MAX_BANDWIDTH
        if (v < 1101 || v > 1105) return -1;
        st->max_bandwidth = v;
        return 0;
    case 4005: *out = st->max_bandwidth; return 0;
    case 4008:   // OPUS_SET_BANDWIDTH
        if ((v < 1101 || v > 1105) && v != -1000) return -1;
        st->user_bandwidth = v;
        return 0;
    case 4009: *out = st->bandwidth; return 0;
    case 4016: if (v < 0 || v > 1) return -1; st->dtx = v; return 0;
    case 4017: *out = st->dtx; return 0;
    case 4010: if (v < 0 || v > 10) return -1; st->complexity = v; return 0;
    case 4011: *out = st->complexity; return 0;
    case 4012: if (v < 0 || v > 1) return -1; st->inband_fec = v; return 0;
    case 4013: *out = st->inband_fec; return 0;
    case 4014: if (v < 0 || v > 100) return -1; st->packet_loss_perc = v; return 0;
    case 4015: *out = st->packet_loss_perc; return 0;
    case 4006: if (v < 0 || v > 1) return -1; st->use_vbr = v; return 0;
    case 4007: *out = st->use_vbr; return 0;
    case 11018: if (v < -1 || v > 100) return -1; st->voice_ratio = v; return 0;
    case 11019: *out = st->voice_ratio; return 0;
    case 4020: if (v < 0 || v > 1) return -1; st->vbr_constraint = v; return 0;
    case 4021: *out = st->vbr_constraint; return 0;
    case 4024: if (v != -1000 && v != kSignalVoice && v != kSignalMusic) return -1; st->signal_type = v; return 0;
    case 4025: *out = st->signal_type; return 0;
    case 4027:   // OPUS_GET_LOOKAHEAD
        *out = st->Fs / 400;
        if (st->application != kAppLowdelay) *out += st->delay_compensation;
        return 0;
    case 4029: *out = st->Fs; return 0;
    case 4031: *out = (int)st->rangeFinal; return 0;
    case 4036: if (v < 8 || v > 24) return -1; st->lsb_depth = v; return 0;
    case 4037: *out = st->lsb_depth; return 0;
    case 4040:   // OPUS_SET_EXPERT_FRAME_DURATION
        if (v != 5000 && v != 5001 && v != 5002 && v != 5003 && v != 5004 && v != 5005 && v != 5006 && v != 5010) return -1;
        st->variable_duration = v;
        return 0;
    case 4041: *out = st->variable_duration; return 0;
    case 4042: if (v > 1 || v < 0) return -1; st->prediction_disabled = v; return 0;
    case 4043: *out = st->prediction_disabled; return 0;
    case 4028: enc_state_reset(st); return 0;   // OPUS_RESET_STATE
    case 11002:   // OPUS_SET_FORCE_MODE
        if ((v < CB_MODE_SILK_ONLY || v > CB_MODE_CELT_ONLY) && v != -1000) return -1;
        st->user_forced_mode = v;
        return 0;
    default: return -5;
    }
}
// 1 if `request` takes a pointer (GET), 0 if it takes a value (SET / RESET)
CB_HD int enc_ctl_is_get(int request) {
    switch (request) {
    case 4001: case 4003: case 4005: case 4007: case 4009: case 4011: case 4013: case 4015: case 4017: case 4021: case 4023: case 4025:
    case 4027: case 4029: case 4031: case 4037: case 4041: case 4043: case 11019: return 1;
    default: return 0;
    }
}

// frame_size_select (opus_encoder.c:807-826): the frame size opus_encode actually codes for OPUS_SET_EXPERT_FRAME_DURATION
CB_HD int frame_size_select(int frame_size, int variable_duration, int Fs) {
    int new_size;
    if (frame_size < Fs / 400) return -1;
    if (variable_duration == 5000) new_size = frame_size;
    else if (variable_duration == 5010) new_size = Fs / 50;
    else if (variable_duration >= 5001 && variable_duration <= 5006) {
        new_size = (Fs / 400) << (variable_duration - 5001);
        if (new_size > 3 * Fs / 50) new_size = 3 * Fs / 50;
    } else return -1;
    if (new_size > frame_size) return -1;
    if (400 * new_size != Fs && 200 * new_size != Fs && 100 * new_size != Fs && 50 * new_size != Fs && 25 * new_size != Fs && 50 * new_size != 3 * Fs)
        return -1;
    return new_size;
}

// gen_toc (opus_encoder.c:254-284)
CB_DEV int gen_toc(int mode, int framerate, int bandwidth, int channels) {
    int period = 0;
    while (framerate < 400) { framerate <<= 1; period++; }
    int toc;
    if (mode == CB_MODE_SILK_ONLY) {
        toc = (bandwidth - 1101) << 5;
        toc |= (period - 2) << 3;
    } else if (mode == CB_MODE_CELT_ONLY) {
        int tmp = bandwidth - 1102;
        if (tmp < 0) tmp = 0;
        toc = 0x80;
        toc |= tmp << 5;
        toc |= period << 3;
    } else {
        toc = 0x60;
        toc |= (bandwidth - 1104) << 4;
        toc |= (period - 2) << 3;
    }
    toc |= (channels == 2) << 2;
    return toc & 0xff;
}

// opus_packet_pad of a one-frame code-0 packet (repacketizer.c:239-258 -> :102-227): the payload moves up by one byte
// (plus the padding-length bytes), the packet becomes code 3 with count 1, zero padding follows.
CB_HD int packet_pad_single(uint8_t *data, int len, int new_len) {
    if (len < 1) return OPUS_BAD_ARG_;
    if (len == new_len) return OPUS_OK_;
    if (len > new_len) return OPUS_BAD_ARG_;
    const int payload = len - 1;
    const int tot_size = payload + 2;
    const int pad_amount = new_len - tot_size;
    int hdr = 2;
    int nb_255s = 0;
    if (pad_amount != 0) {
        nb_255s = (pad_amount - 1) / 255;
        hdr += nb_255s + 1;
    }
    CB_NOUNROLL for (int i = payload - 1; i >= 0; i--) data[hdr + i] = data[1 + i];   // memmove upwards
    data[0] = (uint8_t)((data[0] & 0xFC) | 0x3);
    data[1] = (uint8_t)(1 | (pad_amount != 0 ? 0x40 : 0));
    if (pad_amount != 0) {
        CB_NOUNROLL for (int i = 0; i < nb_255s; i++) data[2 + i] = 255;
        data[2 + nb_255s] = (uint8_t)(pad_amount - 255 * nb_255s - 1);
    }
    CB_NOUNROLL for (int i = hdr + payload; i < new_len; i++) data[i] = 0;
    return OPUS_OK_;
}

// dc_reject (opus_encoder.c:362-385) for one channel: two cascaded one-pole sections, order dependent
CB_DEV_NOINLINE void dc_reject_channel(const int16_t *in, int16_t *out, int32_t *hp_mem, int len, int channels, int c, int shift) {
    int m0 = hp_mem[2 * c], m1 = hp_mem[2 * c + 1];
    CB_NOUNROLL for (int i = 0; i < len; i++) {
        const int x = shl32(in[channels * i + c], 15);
        const int tmp = wsub(x, m0);
        m0 = wadd(m0, pshr32(wsub(x, m0), shift));
        const int y = wsub(tmp, m1);
        m1 = wadd(m1, pshr32(wsub(tmp, m1), shift));
        int v = pshr32(y, 15);
        v = v > 32767 ? 32767 : (v < -32767 ? -32767 : v);
        out[channels * i + c] = (int16_t)v;
    }
    hp_mem[2 * c] = m0;
    hp_mem[2 * c + 1] = m1;
}

struct StereoWidth {
    int XX, XY, YY, smoothed, max_follower, width;
};
// compute_stereo_width (opus_encoder.c:861-936), second half: the smoothed memories and the width from the frame's three sums.
// The new memory is returned in `w` (committed by the caller).
CB_DEV_NOINLINE void stereo_width_finish(int xx, int xy, int yy, int frame_size, int Fs, const CbEncState *st, StereoWidth &w) {
    const int frame_rate = Fs / frame_size;
    const int short_alpha = s16(32767 - 25 * 32767 / imax(50, frame_rate));
    w.XX = wadd(st->width_XX, mul16_32_q15(short_alpha, wsub(xx, st->width_XX)));
    w.XY = wadd(st->width_XY, mul16_32_q15(short_alpha, wsub(xy, st->width_XY)));
    w.YY = wadd(st->width_YY, mul16_32_q15(short_alpha, wsub(yy, st->width_YY)));
    w.XX = imax(0, w.XX); w.XY = imax(0, w.XY); w.YY = imax(0, w.YY);
    w.smoothed = st->width_smoothed;
    w.max_follower = st->width_max_follower;
    if (imax(w.XX, w.YY) > 210) {
        const int sqrt_xx = s16(celt_sqrt(w.XX)), sqrt_yy = s16(celt_sqrt(w.YY));
        const int qrrt_xx = s16(celt_sqrt(sqrt_xx)), qrrt_yy = s16(celt_sqrt(sqrt_yy));
        w.XY = imin(w.XY, sqrt_xx * sqrt_yy);
        const int corr = s16(frac_div32(w.XY, wadd(1, mul16_16(sqrt_xx, sqrt_yy))) >> 16);
        const int ldiff = s16(32767 * iabs(qrrt_xx - qrrt_yy) / (1 + qrrt_xx + qrrt_yy));
        const int width = s16(mul16_16_q15(celt_sqrt(1073741824 - mul16_16(corr, corr)), ldiff));
        w.smoothed = s16(w.smoothed + (width - w.smoothed) / frame_rate);
        w.max_follower = s16(imax(w.max_follower - 655 / frame_rate, w.smoothed));
    }
    w.width = s16(imin(32767, 20 * w.max_follower));
}
// first half: the three correlation sums over the frame (groups of four samples, opus_encoder.c:880-903)
template <class TM>
CB_DEV void compute_stereo_width_team(TM tm, const int16_t *pcm, int frame_size, int Fs, const CbEncState *st, StereoWidth &w) {
    int xx = 0, xy = 0, yy = 0;
    CB_TEAM_FOR(g, (frame_size - 3 + 3) / 4, tm) {
        const int i = 4 * g;
        if (i < frame_size - 3) {
            int pxx = 0, pxy = 0, pyy = 0;
            CB_NOUNROLL for (int k = 0; k < 4; k++) {
                const int x = pcm[2 * (i + k)], y = pcm[2 * (i + k) + 1];
                pxx += mul16_16(x, x) >> 2;
                pxy += mul16_16(x, y) >> 2;
                pyy += mul16_16(y, y) >> 2;
            }
            xx = wadd(xx, pxx >> 10);
            xy = wadd(xy, pxy >> 10);
            yy = wadd(yy, pyy >> 10);
        }
    }
    xx = tm.sum(xx); xy = tm.sum(xy); yy = tm.sum(yy);
    stereo_width_finish(xx, xy, yy, frame_size, Fs, st, w);
}

// stereo_fade (opus_encoder.c:411-441), in place
template <class TM>
CB_DEV void stereo_fade_team(TM tm, int16_t *buf, int g1, int g2, int frame_size, int Fs) {
    const int inc = 48000 / Fs;
    const int overlap = kOverlap / inc;
    g1 = s16(32767 - g1);
    g2 = s16(32767 - g2);
    CB_TEAM_FOR(i, frame_size, tm) {
        int g = g2;
        if (i < overlap) {
            const int w = s16(mul16_16_q15(kWindow120[i * inc], kWindow120[i * inc]));
            g = s16(mac16_16(mul16_16(w, g2), 32767 - w, g1) >> 15);
        }
        int diff = s16(((int)buf[i * 2] - (int)buf[i * 2 + 1]) >> 1);
        diff = mul16_16_q15(g, diff);
        buf[i * 2] = (int16_t)(buf[i * 2] - diff);
        buf[i * 2 + 1] = (int16_t)(buf[i * 2 + 1] + diff);
    }
    tm.sync();
}

CB_TABLE int32_t kBwThreshMonoVoice[8] = {11000, 1000, 14000, 1000, 17000, 1000, 21000, 2000};
CB_TABLE int32_t kBwThreshMonoMusic[8] = {12000, 1000, 15000, 1000, 18000, 2000, 22000, 2000};
CB_TABLE int32_t kBwThreshStereoVoice[8] = {11000, 1000, 14000, 1000, 21000, 2000, 28000, 2000};
CB_TABLE int32_t kBwThreshStereoMusic[8] = {12000, 1000, 18000, 2000, 21000, 2000, 30000, 2000};

template <class TM>
CB_DEV void skip_phases(TM tm, int nsub = 1) {
    for (int i = 0; i < nsub * kEncPhases; i++) tm.phase();
}

// What a 40 / 60 ms frame leaves for its 20 ms sub-frames (opus_encoder.c:1362-1438): the settings the outer call forces on the
// inner calls, and what it restores afterwards.
struct LongFrameCtx {
    int nb_frames, bytes_per_frame;
    int bak_mode, bak_bandwidth, bak_channels;
};
enum { kLongFrame = -1000 };   // internal: "outer decisions committed, now code the sub-frames"

// opus_repacketizer_cat + opus_repacketizer_out_range_impl (repacketizer.c:49-227) for nb (2 or 3) single-frame packets of
// one configuration: src[i] / len[i] are whole sub-packets (TOC + payload, possibly CBR-padded code 3).  Writes the merged
// packet to `data` (<= maxlen), padded to maxlen when `pad`.  Returns its length or a negative error.
CB_DEV int repacketize_frames(uint8_t *data, int maxlen, const uint8_t *const *src, const int *plen, int nb, int pad) {
    const uint8_t *frames[3];
    int len[3];
    uint8_t toc = 0;
    for (int i = 0; i < nb; i++) {
        if (plen[i] < 1) return OPUS_INVALID_PACKET_;
        uint8_t t;
        int16_t size[48];
        int offset;
        const int cnt = pkt_parse(src[i], plen[i], 0, &t, size, &offset, nullptr);
        if (cnt != 1) return cnt < 0 ? cnt : OPUS_INVALID_PACKET_;
        if (i == 0) toc = t;
        else if ((toc & 0xFC) != (t & 0xFC)) return OPUS_INVALID_PACKET_;
        frames[i] = src[i] + offset;
        len[i] = size[0];
    }
    int tot_size = 0;
    uint8_t *ptr = data;
    if (nb == 2) {
        if (len[1] == len[0]) {
            tot_size = 2 * len[0] + 1;
            if (tot_size > maxlen) return OPUS_BUFFER_TOO_SMALL_;
            *ptr++ = (uint8_t)((toc & 0xFC) | 0x1);
        } else {
            tot_size = len[0] + len[1] + 2 + (len[0] >= 252);
            if (tot_size > maxlen) return OPUS_BUFFER_TOO_SMALL_;
            *ptr++ = (uint8_t)((toc & 0xFC) | 0x2);
            if (len[0] < 252) *ptr++ = (uint8_t)len[0];
            else { ptr[0] = (uint8_t)(252 + (len[0] & 3)); ptr[1] = (uint8_t)((len[0] - ptr[0]) >> 2); ptr += 2; }
        }
    }
    if (nb > 2 || (pad && tot_size < maxlen)) {
        ptr = data;
        tot_size = 0;
        int vbr = 0;
        for (int i = 1; i < nb; i++)
            if (len[i] != len[0]) { vbr = 1; break; }
        if (vbr) {
            tot_size += 2;
            for (int i = 0; i < nb - 1; i++) tot_size += 1 + (len[i] >= 252) + len[i];
            tot_size += len[nb - 1];
            if (tot_size > maxlen) return OPUS_BUFFER_TOO_SMALL_;
            *ptr++ = (uint8_t)((toc & 0xFC) | 0x3);
            *ptr++ = (uint8_t)(nb | 0x80);
        } else {
            tot_size += nb * len[0] + 2;
            if (tot_size > maxlen) return OPUS_BUFFER_TOO_SMALL_;
            *ptr++ = (uint8_t)((toc & 0xFC) | 0x3);
            *ptr++ = (uint8_t)nb;
        }
        const int pad_amount = pad ? maxlen - tot_size : 0;
        if (pad_amount != 0) {
            data[1] |= 0x40;
            const int nb_255s = (pad_amount - 1) / 255;
            for (int i = 0; i < nb_255s; i++) *ptr++ = 255;
            *ptr++ = (uint8_t)(pad_amount - 255 * nb_255s - 1);
            tot_size += pad_amount;
        }
        if (vbr)
            for (int i = 0; i < nb - 1; i++) {
                if (len[i] < 252) *ptr++ = (uint8_t)len[i];
                else { ptr[0] = (uint8_t)(252 + (len[i] & 3)); ptr[1] = (uint8_t)((len[i] - ptr[0]) >> 2); ptr += 2; }
            }
    }
    for (int i = 0; i < nb; i++) {
        for (int k = 0; k < len[i]; k++) ptr[k] = frames[i][k];
        ptr += len[i];
    }
    if (pad)
        while (ptr < data + maxlen) *ptr++ = 0;
    return tot_size;
}

// opus_encode_native (opus_encoder.c:938-2005), MODE_CELT_ONLY path.  pcm: frame_size x channels int16; out: >= out_data_bytes.
// st: head of the state (maybe a shared-memory copy), gst: the full block in HBM (delay buffer and sample histories).
// Returns (uniformly on all lanes) the packet length or a negative error code.
template <class TM>
CB_DEV int opus_encode_one(TM tm, CbEncState *st, CbEncState *gst, EncShared &S, EncGlobal &G, const int16_t *pcm, int frame_size, uint8_t *out,
                           int out_data_bytes, LongFrameCtx &lc) {
    const bool L0 = tm.lane() == 0;
    const int Fs = st->Fs, channels = st->channels;
    const int nsub = frame_size > Fs / 50 ? (frame_size > Fs / 25 ? 3 : 2) : 1;   // phase barriers this frame owes its block
    int max_data_bytes = imin(1276, out_data_bytes);
    // frame_size has been through frame_size_select() on the host (opus_encode, opus_encoder.c:2007-2025)
    if ((400 * frame_size != Fs && 200 * frame_size != Fs && 100 * frame_size != Fs && 50 * frame_size != Fs && 25 * frame_size != Fs &&
         50 * frame_size != 3 * Fs) || 400 * frame_size < Fs || max_data_bytes <= 0)
        { skip_phases(tm, nsub); return OPUS_BAD_ARG_; }
    const int delay_compensation = st->application == kAppLowdelay ? 0 : st->delay_compensation;
    const int lsb_depth = imin(16, st->lsb_depth);
    const int total_buffer = delay_compensation;
    int bitrate_bps;
    if (st->user_bitrate_bps == kOpusAuto) bitrate_bps = 60 * Fs / frame_size + Fs * channels;
    else if (st->user_bitrate_bps == kBitrateMax) bitrate_bps = max_data_bytes * 8 * Fs / frame_size;
    else bitrate_bps = st->user_bitrate_bps;
    const int frame_rate = Fs / frame_size;
    if (!st->use_vbr) {
        const int frame_rate3 = 3 * Fs / frame_size;
        const int cbrBytes = imin((3 * bitrate_bps / 8 + frame_rate3 / 2) / frame_rate3, max_data_bytes);
        bitrate_bps = cbrBytes * frame_rate3 * 8 / 3;
        max_data_bytes = cbrBytes;
    }
    // mode / channel / bandwidth decisions into locals first: a frame we cannot code must leave the state untouched
    int voice_est;
    if (st->signal_type == kSignalVoice) voice_est = 127;
    else if (st->signal_type == kSignalMusic) voice_est = 0;
    else if (st->application == kAppVoip) voice_est = 115;
    else voice_est = 48;
    int equiv_rate = bitrate_bps - (40 * channels + 20) * (Fs / frame_size - 50);
    int stream_channels;
    if (st->force_channels != kOpusAuto && channels == 2) {
        stream_channels = st->force_channels;
    } else if (channels == 2) {
        int stereo_threshold = 30000 + ((voice_est * voice_est * (30000 - 30000)) >> 14);
        if (st->stream_channels == 2) stereo_threshold -= 1000;
        else stereo_threshold += 1000;
        stream_channels = equiv_rate > stereo_threshold ? 2 : 1;
    } else {
        stream_channels = channels;
    }
    const bool tiny = max_data_bytes < 3 || bitrate_bps < 3 * frame_rate * 8 || (frame_rate < 50 && (max_data_bytes * frame_rate < 300 || bitrate_bps < 2400));
    StereoWidth sw;
    sw.width = 0;
    const bool want_width = channels == 2 && st->force_channels != 1;
    if (want_width) compute_stereo_width_team(tm, pcm, frame_size, Fs, st, sw);
    if (tiny) {
        // "PLC frame": a bare TOC (opus_encoder.c:1062-1090)
        int tocmode = st->mode;
        int bw = st->bandwidth == 0 ? 1101 : st->bandwidth;
        if (tocmode == 0) tocmode = CB_MODE_SILK_ONLY;
        if (frame_rate > 100) tocmode = CB_MODE_CELT_ONLY;
        if (frame_rate < 50) tocmode = CB_MODE_SILK_ONLY;
        if (tocmode == CB_MODE_SILK_ONLY && bw > 1103) bw = 1103;
        else if (tocmode == CB_MODE_CELT_ONLY && bw == 1102) bw = 1101;
        else if (tocmode == CB_MODE_HYBRID && bw <= 1104) bw = 1104;
        int ret = 1;
        tm.sync();
        if (L0) {
            st->rangeFinal = 0;
            st->voice_ratio = -1;
            st->bitrate_bps = bitrate_bps;
            if (want_width) { st->width_XX = sw.XX; st->width_XY = sw.XY; st->width_YY = sw.YY; st->width_smoothed = sw.smoothed; st->width_max_follower = sw.max_follower; }
            out[0] = (uint8_t)gen_toc(tocmode, frame_rate, bw, st->stream_channels);
            if (!st->use_vbr) packet_pad_single(out, 1, max_data_bytes);
        }
        if (!st->use_vbr) ret = max_data_bytes;
        tm.sync();
        skip_phases(tm, nsub);
        return ret;
    }
    equiv_rate = bitrate_bps - (40 * stream_channels + 20) * (Fs / frame_size - 50);
    int mode;
    if (st->application == kAppLowdelay) {
        mode = CB_MODE_CELT_ONLY;
    } else if (st->user_forced_mode == kOpusAuto) {
        const int stereo_width = sw.width;
        const int mode_voice = mul16_32_q15(32767 - stereo_width, 64000) + mul16_32_q15(stereo_width, 36000);
        const int mode_music = mul16_32_q15(32767 - stereo_width, 16000) + mul16_32_q15(stereo_width, 16000);
        int threshold = mode_music + ((voice_est * voice_est * (mode_voice - mode_music)) >> 14);
        if (st->application == kAppVoip) threshold += 8000;
        if (st->prev_mode == CB_MODE_CELT_ONLY) threshold -= 4000;
        else if (st->prev_mode > 0) threshold += 4000;
        mode = equiv_rate >= threshold ? CB_MODE_CELT_ONLY : CB_MODE_SILK_ONLY;
        if (st->inband_fec && st->packet_loss_perc > (128 - voice_est) >> 4) mode = CB_MODE_SILK_ONLY;
        if (st->dtx && voice_est > 100) mode = CB_MODE_SILK_ONLY;
    } else {
        mode = st->user_forced_mode;
    }
    if (mode != CB_MODE_CELT_ONLY && frame_size < Fs / 100) mode = CB_MODE_CELT_ONLY;
    if (max_data_bytes < (frame_rate > 50 ? 12000 : 8000) * frame_size / (Fs * 8)) mode = CB_MODE_CELT_ONLY;
    if (mode != CB_MODE_CELT_ONLY) { skip_phases(tm, nsub); return OPUS_UNIMPLEMENTED_; }                 // SILK / hybrid: not in this engine
    if (st->prev_mode > 0 && st->prev_mode != CB_MODE_CELT_ONLY) { skip_phases(tm, nsub); return OPUS_UNIMPLEMENTED_; }
    if (st->application == kAppVoip) { skip_phases(tm, nsub); return OPUS_UNIMPLEMENTED_; }               // hp_cutoff (SILK biquad) path
    // bandwidth (opus_encoder.c:1229-1292)
    int bandwidth;
    {
        const int32_t *voice_t, *music_t;
        if (channels == 2 && st->force_channels != 1) { voice_t = kBwThreshStereoVoice; music_t = kBwThreshStereoMusic; }
        else { voice_t = kBwThreshMonoVoice; music_t = kBwThreshMonoMusic; }
        int thr[8];
        CB_NOUNROLL for (int i = 0; i < 8; i++) thr[i] = music_t[i] + ((voice_est * voice_est * (voice_t[i] - music_t[i])) >> 14);
        bandwidth = 1105;
        do {
            int threshold = thr[2 * (bandwidth - 1102)];
            const int hysteresis = thr[2 * (bandwidth - 1102) + 1];
            if (!st->first) {
                if (st->bandwidth >= bandwidth) threshold -= hysteresis;
                else threshold += hysteresis;
            }
            if (equiv_rate >= threshold) break;
        } while (--bandwidth > 1101);
    }
    if (bandwidth > st->max_bandwidth) bandwidth = st->max_bandwidth;
    if (st->user_bandwidth != kOpusAuto) bandwidth = st->user_bandwidth;
    // nothing above the Nyquist rate of the input (opus_encoder.c:1308-1315)
    if (Fs <= 24000 && bandwidth > 1104) bandwidth = 1104;
    if (Fs <= 16000 && bandwidth > 1103) bandwidth = 1103;
    if (Fs <= 12000 && bandwidth > 1102) bandwidth = 1102;
    if (Fs <= 8000 && bandwidth > 1101) bandwidth = 1101;
    if (bandwidth == 1102) bandwidth = 1103;   // CELT has no mediumband
    const int curr_bandwidth = bandwidth;
    const int bytes_target = imin(max_data_bytes, bitrate_bps * frame_size / (Fs * 8)) - 1;

    if (frame_size > Fs / 50) {
        // 40 / 60 ms in CELT-only mode: the decisions above stand for the whole frame, the audio is coded as 2 or 3 20 ms frames
        // with mode / bandwidth / channels forced, then merged by the repacketizer (opus_encoder.c:1362-1438)
        tm.sync();
        if (L0) {
            st->rangeFinal = 0;
            st->voice_ratio = -1;
            st->bitrate_bps = bitrate_bps;
            st->stream_channels = stream_channels;
            st->mode = mode;
            st->bandwidth = bandwidth;
            if (want_width) { st->width_XX = sw.XX; st->width_XY = sw.XY; st->width_YY = sw.YY; st->width_smoothed = sw.smoothed; st->width_max_follower = sw.max_follower; }
        }
        lc.nb_frames = nsub;
        lc.bytes_per_frame = imin(1276, (out_data_bytes - 3) / nsub);
        lc.bak_mode = st->user_forced_mode;
        lc.bak_bandwidth = st->user_bandwidth;
        lc.bak_channels = st->force_channels;
        tm.sync();
        if (L0) {
            st->user_forced_mode = mode;
            st->user_bandwidth = bandwidth;
            st->force_channels = stream_channels;
            st->prev_channels = stream_channels;
        }
        tm.sync();
        return kLongFrame;
    }

    // ---- commit point: from here on the frame is coded ----
    int16_t *pcm_buf = S.u.pcm_buf;
    CB_TEAM_FOR(i, total_buffer * channels, tm) pcm_buf[i] = gst->delay_buffer[(st->encoder_buffer - total_buffer) * channels + i];
    {
        // stage the frame in shared memory (coalesced), then filter it in place: one lane per channel, order dependent
        int16_t *dst = pcm_buf + total_buffer * channels;
        CB_TEAM_FOR(i, frame_size * channels, tm) dst[i] = pcm[i];
        tm.sync();
        const int shift = celt_ilog2(Fs / (3 * 3));
        CB_NOUNROLL for (int c = tm.lane(); c < channels; c += TM::W) dc_reject_channel(dst, dst, st->hp_mem, frame_size, channels, c, shift);
    }
    tm.sync();
    // delay buffer (opus_encoder.c:1773-1781)
    {
        const int eb = st->encoder_buffer;
        if (channels * (eb - (frame_size + total_buffer)) > 0) {
            const int keep = channels * (eb - frame_size - total_buffer);
            // the move overlaps itself: stage through registers in two passes of the whole team
            CB_NOUNROLL for (int base = 0; base < keep; base += TM::W) {
                const int i = base + tm.lane();
                const int v = i < keep ? gst->delay_buffer[channels * frame_size + i] : 0;
                tm.sync();
                if (i < keep) gst->delay_buffer[i] = (int16_t)v;
                tm.sync();
            }
            CB_TEAM_FOR(i, (frame_size + total_buffer) * channels, tm) gst->delay_buffer[keep + i] = pcm_buf[i];
        } else {
            CB_TEAM_FOR(i, eb * channels, tm) gst->delay_buffer[i] = pcm_buf[(frame_size + total_buffer - eb) * channels + i];
        }
        tm.sync();
    }
    // low-rate stereo narrowing (opus_encoder.c:1790-1810)
    const int stereoWidth_Q14 = imin(1 << 14, 2 * imax(0, equiv_rate - 30000));
    if (channels == 2 && (st->hybrid_stereo_width_Q14 < (1 << 14) || stereoWidth_Q14 < (1 << 14))) {
        int g1 = st->hybrid_stereo_width_Q14, g2 = stereoWidth_Q14;
        g1 = g1 == 16384 ? 32767 : shl16(g1, 1);
        g2 = g2 == 16384 ? 32767 : shl16(g2, 1);
        stereo_fade_team(tm, pcm_buf, g1, g2, frame_size, Fs);
        if (L0) st->hybrid_stereo_width_Q14 = stereoWidth_Q14;
    }
    CeltEncCfg cfg;
    cfg.C = stream_channels;
    cfg.end = curr_bandwidth == 1101 ? 13 : curr_bandwidth <= 1103 ? 17 : curr_bandwidth == 1104 ? 19 : 21;
    cfg.complexity = st->complexity;
    cfg.lsb_depth = lsb_depth;
    cfg.loss_rate = st->packet_loss_perc;
    cfg.variable_duration = st->variable_duration;
    const int celt_pred = st->prediction_disabled ? 0 : 2;
    cfg.disable_pf = celt_pred <= 1;
    cfg.force_intra = celt_pred == 0;
    int nb_compr_bytes;
    if (st->use_vbr) {
        cfg.vbr = 1;
        cfg.constrained_vbr = st->vbr_constraint;
        cfg.bitrate = imin(bitrate_bps, 260000 * channels);
        nb_compr_bytes = max_data_bytes - 1;
    } else {
        cfg.vbr = 0;
        cfg.constrained_vbr = st->vbr_constraint;
        cfg.bitrate = kBitrateMax;
        nb_compr_bytes = bytes_target;
    }
    nb_compr_bytes = imin(max_data_bytes - 1, nb_compr_bytes);
    if (L0) {
        st->rangeFinal = 0;
        st->voice_ratio = -1;
        st->bitrate_bps = bitrate_bps;
        st->stream_channels = stream_channels;
        st->mode = mode;
        st->bandwidth = bandwidth;
        if (want_width) { st->width_XX = sw.XX; st->width_XY = sw.XY; st->width_YY = sw.YY; st->width_smoothed = sw.smoothed; st->width_max_follower = sw.max_follower; }
        EcEnc ec;
        ec.init(out + 1, (unsigned)(max_data_bytes - 1));
        ec.shrink((unsigned)nb_compr_bytes);
        S.v.ec = ec;
    }
    tm.sync();
    int ret = 0;
    // "If false, we already busted the budget" cannot happen here: nothing has been coded before the CELT frame
    ret = celt_encode_frame(tm, st, gst, S, G, cfg, pcm_buf, frame_size, nb_compr_bytes);
    if (ret < 0) return OPUS_INTERNAL_ERROR_;
    if (L0) {
        out[0] = (uint8_t)gen_toc(mode, Fs / frame_size, curr_bandwidth, stream_channels);
        st->rangeFinal = S.v.ec.rng;
        st->prev_mode = mode;
        st->prev_channels = stream_channels;
        st->prev_framesize = frame_size;
        st->first = 0;
    }
    ret += 1;
    if (!st->use_vbr) {
        if (L0) packet_pad_single(out, ret, max_data_bytes);
        ret = max_data_bytes;
    }
    tm.sync();
    return ret;
}


// opus_encode_native for any frame size (2.5-60 ms).  pcm: frame_size x channels int16 at the API rate; out: >= out_data_bytes.
// One call site of the frame coder: the outer pass of a 40/60 ms frame only commits its decisions, the sub-frames follow and the
// repacketizer merges them.  (A separate out-of-line copy for long frames measured no faster and 150 KB bigger.)
template <class TM>
CB_DEV int opus_encode_frame(TM tm, CbEncState *st, CbEncState *gst, EncShared &S, EncGlobal &G, const int16_t *pcm, int frame_size, uint8_t *out,
                             int out_data_bytes) {
    LongFrameCtx lc;
    lc.nb_frames = 0;
    int sub_len[3] = {0, 0, 0};
    int i = -1;
    int failed = 0;
    const int Fs = st->Fs, channels = st->channels;
    for (;;) {
        const int16_t *p = i < 0 ? pcm : pcm + (size_t)i * channels * (Fs / 50);
        const int fs = i < 0 ? frame_size : Fs / 50;
        // sub-packets are coded into the TAIL of the caller's slot (nb * bytes_per_frame <= out_data_bytes - 3 by construction)
        uint8_t *o = i < 0 ? out : out + out_data_bytes - (lc.nb_frames - i) * lc.bytes_per_frame;
        const int ob = i < 0 ? out_data_bytes : lc.bytes_per_frame;
        int r;
        if (i >= 0 && (failed || ob < 1)) { skip_phases(tm); r = OPUS_INTERNAL_ERROR_; }
        else r = opus_encode_one(tm, st, gst, S, G, p, fs, o, ob, lc);
        if (i < 0) {
            if (r != kLongFrame) return r;
            i = 0;
            continue;
        }
        if (r < 0) failed = 1;
        sub_len[i] = r;
        if (++i == lc.nb_frames) break;
    }
    int ret = OPUS_INTERNAL_ERROR_;
    tm.sync();
    // the merged packet is assembled at the front of the same slot: stage the sub-packets in shared memory first
    uint8_t *stg = S.u.subpackets;
    {
        const uint8_t *tail = out + out_data_bytes - lc.nb_frames * lc.bytes_per_frame;
        CB_TEAM_FOR(k, lc.nb_frames * lc.bytes_per_frame, tm) stg[k] = tail[k];
    }
    tm.sync();
    if (tm.lane() == 0) {
        if (!failed) {
            const int repacketize_len = st->use_vbr ? out_data_bytes : imin(3 * st->bitrate_bps / (3 * 8 * 50 / lc.nb_frames), out_data_bytes);
            const uint8_t *src[3] = {stg, stg + lc.bytes_per_frame, stg + 2 * lc.bytes_per_frame};
            ret = repacketize_frames(out, repacketize_len, src, sub_len, lc.nb_frames, !st->use_vbr);
            if (ret < 0) ret = OPUS_INTERNAL_ERROR_;
        }
        st->user_forced_mode = lc.bak_mode;
        st->user_bandwidth = lc.bak_bandwidth;
        st->force_channels = lc.bak_channels;
        S.v.ret = ret;
    }
    tm.sync();
    return S.v.ret;
}

}  // namespace cb
