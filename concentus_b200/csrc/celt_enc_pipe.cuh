// celt_enc_pipe.cuh — the frame-synchronous encoder pipeline: one frame of EVERY stream of a launch advances through a chain of
// small kernels, each with the thread mapping its work wants, instead of one warp per stream walking a 250 KB kernel
// (csrc/opus_enc_capi.cu `encode_span_kernel`, kept for the streams this pipeline does not take).
//
// The arithmetic is that of celt_encoder.cuh / opus_encoder_dev.cuh (same functions, same order inside a stream); what changes is
// who executes which slice (opus-fix/src/opus_encoder.c:938-2005, celt/celt_encoder.c:1379-2273):
//
//   per chunk of Fc frames, ahead of the frame steps (nothing here depends on the codec's quantisation state):
//     P0  prepass      thread / stream          Opus-layer decisions for every frame of the chunk (rate -> bytes, channels, bandwidth:
//                                               a recurrence over a handful of scalars), dc_reject (:362-385, a per-channel IIR over
//                                               the whole chunk), compute_stereo_width, stereo_fade -> D (int16) + one EncPlan per frame
//     FE1 pre-emphasis warp / (stream, frame)   celt_preemphasis (:464-535) -> P (int32, with the 1024-sample history in front), the
//                                               frame's maxabs values (silence detection)
//     FE2 pitch        warp / (stream, frame)   pitch_downsample, pitch_search and every candidate remove_doubling can select
//                                               (pitch.c:147-505): `pre` is built from the UN-filtered pre-emphasised input
//                                               (celt_encoder.c:1090-1105,1176-1185), so the analysis is frame-parallel; only the
//                                               selection among the candidates needs the previous frame's period / gain
//   per frame step f (all streams):
//     K1  head         thread / stream          range coder init, byte budget, silence flag, pitch selection + gain decisions,
//                                               post-filter header symbols
//     K2  comb         warp / (stream, channel) comb pre-filter P -> in (FIR), in_mem, transient_analysis of the channel
//     K3  transform    warp / stream            transient decision, MDCT(s), band energies, transient patch, normalisation, tf metrics,
//                                               and the X-only statistics of spreading / stereo / trim analysis
//     K4  decide       thread / stream          tf Viterbi, coarse energy (two-pass), tf / spread / dynalloc / trim symbols, VBR,
//                                               compute_allocation, fine energy
//     K5  bands        the band loop, cut into a data-parallel part and a scalar chain (celt_enc_bandpipe.cuh):
//         K5a prep     warp / stream            stereo angles, mid / side vectors, reordering, split-angle trees
//         K5b chain-S  thread / stream          the loop on a budget-only range coder -> list of leaves (position, N, K)
//         K5c leaves   warp / stream            rotation + PVQ search + codeword index of every listed leaf
//         K5d chain-X  thread / stream          the loop with the real coder and indices, finalise, packet tail
//         (pipe_bands, the loop as one warp-per-stream stage, is kept: hostsim / A-B)
//
// Every stage is a function template over the team type so tests/hostsim can run the very same slices with 1-lane teams.
#pragma once
#include "celt_enc_bandpipe.cuh"
#include "opus_encoder_dev.cuh"

namespace cb {

enum { kPipeHist = kCombMaxPeriod };

// Geometry of one pipeline launch: uniform over its streams.
struct PipeGeom {
    int n;           // streams in the pipeline
    int CC;          // input channels
    int Fs, upsample;
    int fsz;         // frame size at the API rate (samples per channel)
    int N;           // frame size at 48 kHz
    int LM;
    int F;           // frames per stream in the call (row pitch of pcm / data / rets)
    int Fc;          // frames per chunk
    int max_bytes;   // min(1276, max_data_bytes)
    int stride;      // packet slot size
    int pstride;     // ints per (stream, channel) row of P: kPipeHist + Fc * N
};

// Opus-layer plan of one frame (P0 -> K1, K5)
struct EncPlan {
    int code;        // 1: a CELT frame is coded this step; 0: `ret` is final and nothing else happens
    int ret;
    int mode, curr_bandwidth, stream_channels, max_data_bytes, nb_compr_bytes, use_vbr;
    CeltEncCfg cfg;
};

// Everything remove_doubling (pitch.c:372-505) can select from, computed without the previous frame's period / gain.
// Entry 0 is the initial period itself, entry k-1 the sub-harmonic T0/k (k = 2..15, as far as the reference's loop runs).
struct PitchCand {
    int ncand;
    int T0h;                 // the halved, clamped initial period
    int T[15];               // halved candidate periods
    int g[15], xy[15], yy[15];
    int xc[15][3];           // x . x[-(T-1)], x . x[-T], x . x[-(T+1)] for the final +-1 refinement
};

// Front-end results of one frame
struct FeFrame {
    int maxabs_a, maxabs_b;  // celt_maxabs16 of the body / the overlap tail of the frame (celt_encoder.c:1567-1571)
    int pitch_index;         // pitch_search result (undoubled); valid when `pitch_done`
    int pitch_done;
    PitchCand pc;
};

// Per-stream working set of the frame in flight (global memory; the L2 holds it between the kernels of a step).
struct EncPipeCtx {
    EncVars v;
    int code;                // copy of the plan's flag
    int LM;
    CeltEncCfg cfg;
    int pf_T0, pf_g0, pf_tap0;                        // previous post-filter parameters as K2 needs them
    int spread_sum, spread_nb, spread_hf, spread_skip;   // spreading_decision statistics of X
    int st_sumLR, st_sumMS;                           // stereo_analysis sums
    int16_t trim_xc[kNbEBands];                       // alloc_trim_analysis: per-band L.R correlation >> 18
    int dual_stereo;
    int metric[kNbEBands];
    int tf_sum_team;
    int bandE[2 * kNbEBands];
    int16_t bandLogE[2 * kNbEBands], bandLogE2[2 * kNbEBands], error[2 * kNbEBands];
    int16_t band_g[2 * kNbEBands];
    int8_t band_shift[2 * kNbEBands];
    int tf_res[kNbEBands], offsets[kNbEBands], cap[kNbEBands], fine_quant[kNbEBands], pulses[kNbEBands], fine_priority[kNbEBands];
    AllocScratch alloc;
    CoarseScratch coarse;
};

// Per-stream sample buffers of the frame in flight
struct EncPipeBuf {
    int in[2 * (kMaxFrame + kOverlap)];     // comb-filtered input + overlap history, per channel
    int freq[2 * kMaxFrame];                // MDCT output
    int16_t X[2 * kMaxFrame];               // normalised spectrum
};

// ---------------------------------------------------------------------------------------------------------------------------
// P0 — Opus layer of one frame for a stream the pipeline takes (opus_encode_one's decisions, committed as the legacy path does).
// Eligibility (host side, enc_pipe_eligible): OPUS_APPLICATION_RESTRICTED_LOWDELAY, frame <= 20 ms, not a "PLC frame" budget.
// Returns the plan; when plan.code the frame's PCM has been DC-rejected (and narrowed) into `D`.
// ---------------------------------------------------------------------------------------------------------------------------
// What the decisions of a frame leave for its sample pass and its commit
struct PlanPre {
    int code, ret;
    int want_width, fade, g1, g2, dc_shift;
    int bitrate_bps, equiv_rate, stream_channels, mode, bandwidth, curr_bandwidth, max_data_bytes, bytes_target, lsb_depth, stereoWidth_Q14;
};

// (1) the decisions: scalars only, nothing is written
CB_DEV_NOINLINE void pipe_plan_pre(const CbEncState *st, int frame_size, int out_data_bytes, PlanPre &pp) {
    const int Fs = st->Fs, channels = st->channels;
    pp.code = 0;
    pp.want_width = 0; pp.fade = 0;
    int max_data_bytes = imin(1276, out_data_bytes);
    if ((400 * frame_size != Fs && 200 * frame_size != Fs && 100 * frame_size != Fs && 50 * frame_size != Fs) || max_data_bytes <= 0 ||
        st->application != kAppLowdelay) {
        pp.ret = OPUS_INTERNAL_ERROR_;   // not a frame for this pipeline: the host's eligibility test keeps these out
        return;
    }
    pp.lsb_depth = imin(16, st->lsb_depth);
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
    int voice_est;
    if (st->signal_type == kSignalVoice) voice_est = 127;
    else if (st->signal_type == kSignalMusic) voice_est = 0;
    else voice_est = 48;
    int equiv_rate = bitrate_bps - (40 * channels + 20) * (Fs / frame_size - 50);
    int stream_channels;
    if (st->force_channels != kOpusAuto && channels == 2) {
        stream_channels = st->force_channels;
    } else if (channels == 2) {
        int stereo_threshold = 30000;
        if (st->stream_channels == 2) stereo_threshold -= 1000;
        else stereo_threshold += 1000;
        stream_channels = equiv_rate > stereo_threshold ? 2 : 1;
    } else {
        stream_channels = channels;
    }
    const bool tiny = max_data_bytes < 3 || bitrate_bps < 3 * frame_rate * 8 || (frame_rate < 50 && (max_data_bytes * frame_rate < 300 || bitrate_bps < 2400));
    if (tiny) { pp.ret = OPUS_INTERNAL_ERROR_; return; }
    pp.want_width = channels == 2 && st->force_channels != 1;
    equiv_rate = bitrate_bps - (40 * stream_channels + 20) * (Fs / frame_size - 50);
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
    if (Fs <= 24000 && bandwidth > 1104) bandwidth = 1104;
    if (Fs <= 16000 && bandwidth > 1103) bandwidth = 1103;
    if (Fs <= 12000 && bandwidth > 1102) bandwidth = 1102;
    if (Fs <= 8000 && bandwidth > 1101) bandwidth = 1101;
    if (bandwidth == 1102) bandwidth = 1103;
    pp.bytes_target = imin(max_data_bytes, bitrate_bps * frame_size / (Fs * 8)) - 1;
    pp.dc_shift = celt_ilog2(Fs / (3 * 3));
    pp.stereoWidth_Q14 = imin(1 << 14, 2 * imax(0, equiv_rate - 30000));
    if (channels == 2 && (st->hybrid_stereo_width_Q14 < (1 << 14) || pp.stereoWidth_Q14 < (1 << 14))) {
        int g1 = st->hybrid_stereo_width_Q14, g2 = pp.stereoWidth_Q14;
        pp.g1 = g1 == 16384 ? 32767 : shl16(g1, 1);
        pp.g2 = g2 == 16384 ? 32767 : shl16(g2, 1);
        pp.fade = 1;
    }
    pp.bitrate_bps = bitrate_bps; pp.equiv_rate = equiv_rate; pp.stream_channels = stream_channels; pp.mode = CB_MODE_CELT_ONLY;
    pp.bandwidth = bandwidth; pp.curr_bandwidth = bandwidth; pp.max_data_bytes = max_data_bytes;
    pp.code = 1;
    pp.ret = 0;
}

// (3) the commit, after the frame's samples have been processed (xx / xy / yy: compute_stereo_width's sums when pp.want_width)
CB_DEV_NOINLINE void pipe_plan_post(CbEncState *st, const PlanPre &pp, int frame_size, int xx, int xy, int yy, EncPlan &pl) {
    pl.code = pp.code;
    pl.ret = pp.ret;
    if (!pp.code) return;
    const int channels = st->channels;
    if (pp.want_width) {
        StereoWidth sw;
        stereo_width_finish(xx, xy, yy, frame_size, st->Fs, st, sw);
        st->width_XX = sw.XX; st->width_XY = sw.XY; st->width_YY = sw.YY; st->width_smoothed = sw.smoothed; st->width_max_follower = sw.max_follower;
    }
    if (pp.fade) st->hybrid_stereo_width_Q14 = pp.stereoWidth_Q14;
    pl.cfg.C = pp.stream_channels;
    pl.cfg.end = pp.curr_bandwidth == 1101 ? 13 : pp.curr_bandwidth <= 1103 ? 17 : pp.curr_bandwidth == 1104 ? 19 : 21;
    pl.cfg.complexity = st->complexity;
    pl.cfg.lsb_depth = pp.lsb_depth;
    pl.cfg.loss_rate = st->packet_loss_perc;
    pl.cfg.variable_duration = st->variable_duration;
    const int celt_pred = st->prediction_disabled ? 0 : 2;
    pl.cfg.disable_pf = celt_pred <= 1;
    pl.cfg.force_intra = celt_pred == 0;
    int nb_compr_bytes;
    if (st->use_vbr) {
        pl.cfg.vbr = 1;
        pl.cfg.constrained_vbr = st->vbr_constraint;
        pl.cfg.bitrate = imin(pp.bitrate_bps, 260000 * channels);
        nb_compr_bytes = pp.max_data_bytes - 1;
    } else {
        pl.cfg.vbr = 0;
        pl.cfg.constrained_vbr = st->vbr_constraint;
        pl.cfg.bitrate = kBitrateMax;
        nb_compr_bytes = pp.bytes_target;
    }
    nb_compr_bytes = imin(pp.max_data_bytes - 1, nb_compr_bytes);
    st->voice_ratio = -1;
    st->bitrate_bps = pp.bitrate_bps;
    st->stream_channels = pp.stream_channels;
    st->mode = pp.mode;
    st->bandwidth = pp.bandwidth;
    st->prev_mode = pp.mode;
    st->prev_channels = pp.stream_channels;
    st->prev_framesize = frame_size;
    st->first = 0;
    pl.mode = pp.mode; pl.curr_bandwidth = pp.curr_bandwidth; pl.stream_channels = pp.stream_channels;
    pl.max_data_bytes = pp.max_data_bytes; pl.nb_compr_bytes = nb_compr_bytes; pl.use_vbr = st->use_vbr;
}

// (2) one sample of dc_reject (opus_encoder.c:362-385) for one channel; m0 / m1: the channel's two filter memories
CB_DEV int dc_reject_step(int x16, int &m0, int &m1, int shift) {
    const int x = shl32(x16, 15);
    const int tmp = wsub(x, m0);
    m0 = wadd(m0, pshr32(tmp, shift));
    const int y = wsub(tmp, m1);
    m1 = wadd(m1, pshr32(y, shift));
    int v = pshr32(y, 15);
    return v > 32767 ? 32767 : (v < -32767 ? -32767 : v);
}
// the stereo_fade gain at sample i (opus_encoder.c:411-441)
CB_DEV int stereo_fade_gain(int i, int g1, int g2, int Fs) {
    const int inc = 48000 / Fs;
    const int overlap = kOverlap / inc;
    g1 = s16(32767 - g1);
    g2 = s16(32767 - g2);
    if (i >= overlap) return g2;
    const int w = s16(mul16_16_q15(kWindow120[i * inc], kWindow120[i * inc]));
    return s16(mac16_16(mul16_16(w, g2), 32767 - w, g1) >> 15);
}

// The three steps for one frame by one thread (tests/hostsim; the prepass kernel spreads step 2 over a lane per channel).
CB_DEV_NOINLINE void pipe_plan_frame(CbEncState *st, const int16_t *pcm, int frame_size, int out_data_bytes, int16_t *D, EncPlan &pl) {
    PlanPre pp;
    pipe_plan_pre(st, frame_size, out_data_bytes, pp);
    int xx = 0, xy = 0, yy = 0;
    if (pp.code) {
        const int channels = st->channels, Fs = st->Fs;
        if (pp.want_width) {
            CB_NOUNROLL for (int i = 0; i < frame_size - 3; i += 4) {
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
        CB_NOUNROLL for (int c = 0; c < channels; c++) {
            int m0 = st->hp_mem[2 * c], m1 = st->hp_mem[2 * c + 1];
            CB_NOUNROLL for (int i = 0; i < frame_size; i++) D[channels * i + c] = (int16_t)dc_reject_step(pcm[channels * i + c], m0, m1, pp.dc_shift);
            st->hp_mem[2 * c] = m0; st->hp_mem[2 * c + 1] = m1;
        }
        if (pp.fade)
            CB_NOUNROLL for (int i = 0; i < frame_size; i++) {
                const int gq = stereo_fade_gain(i, pp.g1, pp.g2, Fs);
                int diff = s16(((int)D[i * 2] - (int)D[i * 2 + 1]) >> 1);
                diff = mul16_16_q15(gq, diff);
                D[i * 2] = (int16_t)(D[i * 2] - diff);
                D[i * 2 + 1] = (int16_t)(D[i * 2 + 1] + diff);
            }
    }
    pipe_plan_post(st, pp, frame_size, xx, xy, yy, pl);
}

// Host-side test: may this stream go through the pipeline for a span of `frame_size` frames with `out_data_bytes` per packet?
// Only ctl-visible configuration is read (it never changes on the device).
CB_HD int enc_pipe_eligible(const CbEncState *st, int frame_size, int out_data_bytes) {
    const int Fs = st->Fs, channels = st->channels;
    if (st->application != kAppLowdelay) return 0;
    if (400 * frame_size != Fs && 200 * frame_size != Fs && 100 * frame_size != Fs && 50 * frame_size != Fs) return 0;
    int max_data_bytes = out_data_bytes < 1276 ? out_data_bytes : 1276;
    if (max_data_bytes <= 0) return 0;
    int bitrate_bps;
    if (st->user_bitrate_bps == -1000) bitrate_bps = 60 * Fs / frame_size + Fs * channels;
    else if (st->user_bitrate_bps == -1) bitrate_bps = max_data_bytes * 8 * Fs / frame_size;
    else bitrate_bps = st->user_bitrate_bps;
    const int frame_rate = Fs / frame_size;
    if (!st->use_vbr) {
        const int frame_rate3 = 3 * Fs / frame_size;
        int cbrBytes = (3 * bitrate_bps / 8 + frame_rate3 / 2) / frame_rate3;
        if (cbrBytes > max_data_bytes) cbrBytes = max_data_bytes;
        bitrate_bps = cbrBytes * frame_rate3 * 8 / 3;
        max_data_bytes = cbrBytes;
    }
    if (max_data_bytes < 3 || bitrate_bps < 3 * frame_rate * 8 || (frame_rate < 50 && (max_data_bytes * frame_rate < 300 || bitrate_bps < 2400))) return 0;
    return 1;
}

// ---------------------------------------------------------------------------------------------------------------------------
// FE1 — pre-emphasis of one frame into the chunk's P rows, and the silence-detection maxima.
// D: the frame's DC-rejected PCM (CC-interleaved, fsz samples per channel); Prow: P row of channel 0 of this stream (channel c at
// + c * pstride); the frame's samples go to Prow[kPipeHist + fi * N + i].  m0[c]: the pre-emphasis memory entering the frame.
// ---------------------------------------------------------------------------------------------------------------------------
template <class TM>
CB_DEV void pipe_preemph_frame(TM tm, const PipeGeom &g, const EncPlan &pl, const int16_t *D, int *Prow, int fi, const int *m0, FeFrame &fe) {
    const int CC = g.CC, N = g.N, up = g.upsample, ov = kOverlap, C = pl.cfg.C;
    const int a = team_maxabs16(tm, D, C * (N - ov) / up);
    const int b = team_maxabs16(tm, D + C * (N - ov) / up, C * ov / up);
    if (tm.lane() == 0) { fe.maxabs_a = a; fe.maxabs_b = b; }
    CB_NOUNROLL for (int c = 0; c < CC; c++) {
        int *inp = Prow + c * g.pstride + kPipeHist + fi * N;
        const int mc = m0[c];
        if (up == 1) {
            CB_TEAM_FOR(i, N, tm) {
                const int x = D[CC * i + c];
                const int m = i == 0 ? mc : mul16_16(kPreemphCoef0, D[CC * (i - 1) + c]) >> 3;
                inp[i] = wsub(shl32(x, 12), m);
            }
        } else {
            CB_TEAM_FOR(i, N, tm) {
                const int x = i % up == 0 ? (int)D[CC * (i / up) + c] : 0;
                const int xp = i == 0 ? 0 : ((i - 1) % up == 0 ? (int)D[CC * ((i - 1) / up) + c] : 0);
                const int m = i == 0 ? mc : mul16_16(kPreemphCoef0, xp) >> 3;
                inp[i] = wsub(shl32(x, 12), m);
            }
        }
    }
}
// the pre-emphasis memory a frame leaves behind (celt_encoder.c:476-488 fast path, :490-533 zero-stuffed path)
CB_DEV int pipe_preemph_mem_after(const PipeGeom &g, const int16_t *D, int c) {
    return g.upsample == 1 ? mul16_16(kPreemphCoef0, D[g.CC * (g.fsz - 1) + c]) >> 3 : 0;
}

// Scratch of FE2 (per team): the pitch buffers of celt_encoder.cuh's EncShared::Phase::pf
struct PitchScratch {
    int16_t pitch_buf[(kCombMaxPeriod + kMaxFrame) / 2];
    int16_t x_lp4[kMaxFrame / 4], y_lp4[(kMaxFrame + kCombMaxPeriod) / 4];
    union {
        int16_t pitch_raw[(kCombMaxPeriod + kMaxFrame) / 2];
        struct { int xcorr[kCombMaxPeriod / 2]; int yy_lookup[kCombMaxPeriod / 2 + 1]; } c;
    } a;
};

// remove_doubling (pitch.c:372-505), the part that does not depend on the previous frame: every candidate period with its gain,
// correlations and refinement taps.  x: pitch_buf; T0: the pitch_search result (already maxperiod - index).
template <class TM>
CB_DEV_NOINLINE void remove_doubling_candidates(TM tm, const int16_t *x, int maxperiod, int minperiod, int N, int T0_in, int *yy_lookup, PitchCand &pc) {
    maxperiod /= 2; minperiod /= 2; N /= 2;
    int T0 = T0_in / 2;
    x += maxperiod;
    if (T0 >= maxperiod) T0 = maxperiod - 1;
    int xx, xy;
    {
        int a = 0, b = 0;
        CB_TEAM_FOR(i, N, tm) { a = mac16_16(a, x[i], x[i]); b = mac16_16(b, x[i], x[i - T0]); }
        xx = tm.sum(a);
        xy = tm.sum(b);
    }
    {
        const int per = (maxperiod + TM::W - 1) / TM::W;
        const int first = 1 + tm.lane() * per;
        int local = 0;
        CB_NOUNROLL for (int i = first; i < first + per && i <= maxperiod; i++)
            local = wsub(wadd(local, mul16_16(x[-i], x[-i])), mul16_16(x[N - i], x[N - i]));
        int yy = wadd(xx, tm.exscan(local));
        CB_NOUNROLL for (int i = first; i < first + per && i <= maxperiod; i++) {
            yy = wsub(wadd(yy, mul16_16(x[-i], x[-i])), mul16_16(x[N - i], x[N - i]));
            yy_lookup[i] = imax(0, yy);
        }
        if (tm.lane() == 0) yy_lookup[0] = xx;
        tm.sync();
    }
    int nc = 0;
    int Tl[15], gl[15], xyl[15], yyl[15];
    {
        const int yy = yy_lookup[T0];
        int x2y2 = wadd(1, mul32_32_q31(xx, yy) >> 1);
        int sh = celt_ilog2(x2y2) >> 1;
        int t = vshr32(x2y2, 2 * (sh - 7));
        Tl[0] = T0; gl[0] = vshr32(mul16_32_q15(celt_rsqrt_norm(t), xy), sh + 1); xyl[0] = xy; yyl[0] = yy;
        nc = 1;
    }
    CB_NOUNROLL for (int k = 2; k <= 15; k++) {
        int T1 = (int)udiv((unsigned)(2 * T0 + k), (unsigned)(2 * k));
        if (T1 < minperiod) break;
        int T1b;
        if (k == 2) {
            if (T1 + T0 > maxperiod) T1b = T0;
            else T1b = T0 + T1;
        } else {
            T1b = (int)udiv((unsigned)(2 * kSecondCheck[k] * T0 + k), (unsigned)(2 * k));
        }
        int a = 0, b = 0;
        CB_TEAM_FOR(i, N, tm) { a = mac16_16(a, x[i], x[i - T1]); b = mac16_16(b, x[i], x[i - T1b]); }
        const int xyk = wadd(tm.sum(a), tm.sum(b));
        const int yyk = wadd(yy_lookup[T1], yy_lookup[T1b]);
        int x2y2 = wadd(1, mul32_32_q31(xx, yyk));
        int sh = celt_ilog2(x2y2) >> 1;
        int t = vshr32(x2y2, 2 * (sh - 7));
        Tl[k - 1] = T1; gl[k - 1] = vshr32(mul16_32_q15(celt_rsqrt_norm(t), xyk), sh + 1); xyl[k - 1] = xyk; yyl[k - 1] = yyk;
        nc = k;
    }
    CB_NOUNROLL for (int j = 0; j < nc; j++) {
        const int T = Tl[j];
        int a = 0, b = 0, c = 0;
        CB_TEAM_FOR(i, N, tm) {
            a = mac16_16(a, x[i], x[i - (T - 1)]);
            b = mac16_16(b, x[i], x[i - T]);
            c = mac16_16(c, x[i], x[i - (T + 1)]);
        }
        a = tm.sum(a); b = tm.sum(b); c = tm.sum(c);
        if (tm.lane() == 0) {
            pc.T[j] = T; pc.g[j] = gl[j]; pc.xy[j] = xyl[j]; pc.yy[j] = yyl[j];
            pc.xc[j][0] = a; pc.xc[j][1] = b; pc.xc[j][2] = c;
        }
    }
    if (tm.lane() == 0) { pc.ncand = nc; pc.T0h = T0; }
    tm.sync();
}

// ... and the part that does: the walk over the candidates with the previous period / gain.  Returns the gain, *T0_ = the period.
CB_DEV_NOINLINE int remove_doubling_select(const PitchCand &pc, int minperiod, int *T0_, int prev_period, int prev_gain) {
    const int minperiod0 = minperiod;
    minperiod /= 2; prev_period /= 2;
    const int T0 = pc.T0h;
    const int g0 = pc.g[0];
    int best = 0;
    CB_NOUNROLL for (int k = 2; k <= pc.ncand; k++) {
        const int T1 = pc.T[k - 1];
        const int g1 = pc.g[k - 1];
        int cont;
        if (iabs(T1 - prev_period) <= 1) cont = prev_gain;
        else if (iabs(T1 - prev_period) <= 2 && 5 * k * k < T0) cont = s16(prev_gain >> 1);
        else cont = 0;
        int thresh = imax(9830, wsub(mul16_32_q15(22938, g0), cont));
        if (T1 < 3 * minperiod) thresh = imax(13107, wsub(mul16_32_q15(27853, g0), cont));
        else if (T1 < 2 * minperiod) thresh = imax(16384, wsub(mul16_32_q15(29491, g0), cont));
        if (g1 > thresh) best = k - 1;
    }
    const int best_xy = imax(0, pc.xy[best]);
    const int best_yy = pc.yy[best];
    const int T = pc.T[best], g = pc.g[best];
    int pg;
    if (best_yy <= best_xy) pg = 32767;
    else pg = s16(frac_div32(best_xy, wadd(best_yy, 1)) >> 16);
    const int *xc = pc.xc[best];
    int offset;
    if (wsub(xc[2], xc[0]) > mul16_32_q15(22938, wsub(xc[1], xc[0]))) offset = 1;
    else if (wsub(xc[0], xc[2]) > mul16_32_q15(22938, wsub(xc[1], xc[2]))) offset = -1;
    else offset = 0;
    if (pg > g) pg = s16(g);
    *T0_ = 2 * T + offset;
    if (*T0_ < minperiod0) *T0_ = minperiod0;
    return pg;
}

// FE2 — pitch analysis of one frame.  pre0 / pre1: the frame's window of P (kCombMaxPeriod history + N new), per channel.
template <class TM>
CB_DEV void pipe_pitch_frame(TM tm, const PipeGeom &g, const EncPlan &pl, const int *pre0, const int *pre1, PitchScratch &ps, FeFrame &fe) {
    const int N = g.N;
    if (!pl.code || pl.cfg.disable_pf || pl.cfg.complexity < 5) {
        if (tm.lane() == 0) fe.pitch_done = 0;
        return;
    }
    pitch_downsample_team(tm, pre0, pre1, kCombMaxPeriod + N, g.CC, ps.a.pitch_raw, ps.pitch_buf);
    int pitch_index = pitch_search_team(tm, ps.pitch_buf + (kCombMaxPeriod >> 1), ps.pitch_buf, N, kCombMaxPeriod - 3 * kCombMinPeriod, ps.x_lp4, ps.y_lp4,
                                        ps.a.c.xcorr, ps.a.c.yy_lookup);
    pitch_index = kCombMaxPeriod - pitch_index;
    tm.sync();
    remove_doubling_candidates(tm, ps.pitch_buf, kCombMaxPeriod, kCombMinPeriod, N, pitch_index, ps.a.c.yy_lookup, fe.pc);
    if (tm.lane() == 0) { fe.pitch_index = pitch_index; fe.pitch_done = 1; }
}

// ---------------------------------------------------------------------------------------------------------------------------
// K1 — head of the frame (celt_encoder.c:1480-1674 minus the signal processing): scalar, one thread per stream.
// ---------------------------------------------------------------------------------------------------------------------------
CB_DEV_NOINLINE void pipe_head(CbEncState *st, const PipeGeom &g, const EncPlan &pl, const FeFrame &fe, EncPipeCtx &X, uint8_t *out) {
    EncVars &V = X.v;
    X.code = pl.code;
    if (!pl.code) { V.ret = pl.ret; return; }
    const CeltEncCfg cfg = pl.cfg;
    X.cfg = cfg;
    X.LM = g.LM;
    const int C = cfg.C, LM = g.LM, N = g.N, start = 0;
    st->rangeFinal = 0;
    EcEnc ec;
    ec.init(out + 1, (unsigned)(pl.max_data_bytes - 1));
    ec.shrink((unsigned)pl.nb_compr_bytes);
    // ---- rate bookkeeping (:1480-1560) ----
    int nbCompressedBytes = pl.nb_compr_bytes;
    int tell = ec.tell();
    const int nbFilledBytes = (tell + 4) >> 3;
    nbCompressedBytes = imin(nbCompressedBytes, 1275);
    int nbAvailableBytes = nbCompressedBytes - nbFilledBytes;
    int vbr_rate, effectiveBytes;
    if (cfg.vbr && cfg.bitrate != kBitrateMax) {
        const int den = 48000 >> kBitRes;
        vbr_rate = (cfg.bitrate * N + (den >> 1)) / den;
        effectiveBytes = vbr_rate >> (3 + kBitRes);
    } else {
        vbr_rate = 0;
        int tmp = wmul(cfg.bitrate, N);
        if (tell > 1) tmp += tell;
        if (cfg.bitrate != kBitrateMax) nbCompressedBytes = imax(2, imin(nbCompressedBytes, (tmp + 4 * 48000) / (8 * 48000)));
        effectiveBytes = nbCompressedBytes;
    }
    int equiv_rate = 510000;
    if (cfg.bitrate != kBitrateMax) equiv_rate = cfg.bitrate - (40 * C + 20) * ((400 >> LM) - 50);
    if (vbr_rate > 0 && cfg.constrained_vbr) {
        const int vbr_bound = vbr_rate;
        const int max_allowed = imin(imax(tell == 1 ? 2 : 0, (vbr_rate + vbr_bound - st->vbr_reservoir) >> (kBitRes + 3)), nbAvailableBytes);
        if (max_allowed < nbAvailableBytes) {
            nbCompressedBytes = nbFilledBytes + max_allowed;
            nbAvailableBytes = max_allowed;
            ec.shrink((unsigned)nbCompressedBytes);
        }
    }
    int total_bits = nbCompressedBytes * 8;
    // ---- silence (:1567-1571, :1605-1635) ----
    const int sample_max = imax(imax(st->overlap_max, fe.maxabs_a), fe.maxabs_b);
    st->overlap_max = fe.maxabs_b;
    // pre-emphasis memory: FE1 keeps it (per chunk); nothing to do here
    int silence = sample_max == 0;
    if (tell == 1) ec.bit_logp(silence, 15);
    else silence = 0;
    if (silence) {
        if (vbr_rate > 0) {
            effectiveBytes = nbCompressedBytes = imin(nbCompressedBytes, nbFilledBytes + 2);
            total_bits = nbCompressedBytes * 8;
            nbAvailableBytes = 2;
            ec.shrink((unsigned)nbCompressedBytes);
        }
        tell = nbCompressedBytes * 8;
        ec.nbits_total += tell - ec.tell();
    }
    const int enabled = nbAvailableBytes > 12 * C && start == 0 && !silence && !cfg.disable_pf && cfg.complexity >= 5 &&
                        !(st->consec_transient && LM != 3 && cfg.variable_duration == kFramesizeVariable);
    const int prefilter_tapset = st->tapset_decision;
    const int prefilter_period0 = imax(st->prefilter_period, kCombMinPeriod);
    // ---- pitch decision (run_prefilter, :1107-1160) ----
    int pitch_index, gain1;
    const int prev_period = st->prefilter_period, prev_gain = st->prefilter_gain;
    if (enabled) {
        pitch_index = fe.pitch_index;
        gain1 = remove_doubling_select(fe.pc, kCombMinPeriod, &pitch_index, prev_period, prev_gain);
        if (pitch_index > kCombMaxPeriod - 2) pitch_index = kCombMaxPeriod - 2;
        gain1 = s16(mul16_16_q15(22938, gain1));
        if (cfg.loss_rate > 2) gain1 = gain1 >> 1;
        if (cfg.loss_rate > 4) gain1 = gain1 >> 1;
        if (cfg.loss_rate > 8) gain1 = 0;
    } else {
        gain1 = 0;
        pitch_index = kCombMinPeriod;
    }
    int pf_threshold = 6554;
    if (iabs(pitch_index - prev_period) * 10 > pitch_index) pf_threshold += 6554;
    if (nbAvailableBytes < 25) pf_threshold += 3277;
    if (nbAvailableBytes < 35) pf_threshold += 3277;
    if (prev_gain > 13107) pf_threshold -= 3277;
    if (prev_gain > 18022) pf_threshold -= 3277;
    pf_threshold = imax(pf_threshold, 6554);
    int pf_on, qg;
    if (gain1 < pf_threshold) {
        gain1 = 0; pf_on = 0; qg = 0;
    } else {
        if (iabs(gain1 - prev_gain) < 3277) gain1 = prev_gain;
        qg = ((gain1 + 1536) >> 10) / 3 - 1;
        qg = imax(0, imin(7, qg));
        gain1 = 3072 * (qg + 1);
        pf_on = 1;
    }
    if (pf_on == 0) {
        if (start == 0 && tell + 16 <= total_bits) ec.bit_logp(0, 1);
    } else {
        ec.bit_logp(1, 1);
        const int pi1 = pitch_index + 1;
        const int octave = ec_ilog((unsigned)pi1) - 5;
        ec.uint_((unsigned)octave, 6);
        ec.bits((unsigned)(pi1 - (16 << octave)), (unsigned)(4 + octave));
        ec.bits((unsigned)qg, 3);
        ec.icdf(prefilter_tapset, kTapsetIcdf, 2);
    }
    V.ec = ec;
    V.nbCompressedBytes = nbCompressedBytes; V.nbAvailableBytes = nbAvailableBytes; V.nbFilledBytes = nbFilledBytes;
    V.vbr_rate = vbr_rate; V.effectiveBytes = effectiveBytes; V.equiv_rate = equiv_rate; V.total_bits = total_bits;
    V.silence = silence; V.tell = tell; V.enabled = enabled;
    V.pf_on = pf_on; V.pitch_index = pitch_index; V.gain1 = gain1; V.qg = qg;
    V.prefilter_tapset = prefilter_tapset; V.prefilter_period0 = prefilter_period0;
    X.pf_T0 = prefilter_period0; X.pf_g0 = prev_gain; X.pf_tap0 = st->prefilter_tapset;
    V.ret = 0;
}

// ---------------------------------------------------------------------------------------------------------------------------
// K2 — one channel: comb pre-filter (celt_encoder.c:1164-1185) and transient_analysis (:227-378).  tin: N + overlap ints of team
// scratch, sc: 4 ints.  pre: the channel's window of P (history first).  inc: the channel's row of `in` (overlap + N).
// ---------------------------------------------------------------------------------------------------------------------------
// (tin == nullptr: the comb filter only — the pipeline runs transient_analysis in its own kernel, one thread per channel)
template <class TM>
CB_DEV void pipe_comb_channel(TM tm, CbEncState *st, const PipeGeom &g, EncPipeCtx &X, const int *pre, int *inc, int c, int *tin, int *sc) {
    if (!X.code) return;
    const EncVars &V = X.v;
    const int N = g.N, ov = kOverlap;
    CB_TEAM_FOR(i, ov, tm) inc[i] = st->in_mem[c * ov + i];
    comb_filter_fir_team(tm, inc + ov, pre + kCombMaxPeriod, X.pf_T0, V.pitch_index, N, -X.pf_g0, -V.gain1, X.pf_tap0, V.prefilter_tapset, ov);
    tm.sync();
    CB_TEAM_FOR(i, ov, tm) st->in_mem[c * ov + i] = inc[N + i];
    if (X.cfg.complexity >= 1 && tin != nullptr) {
        CB_TEAM_FOR(i, N + ov, tm) tin[i] = inc[i] >> 12;
        tm.sync();
        transient_analysis_team(tm, tin, N + ov, 1, sc, &X.v.mask_metric[c]);
    }
}

// transient_analysis (celt_encoder.c:227-378) of ONE channel by ONE thread, in two parts: the high-pass recurrence is fed sample by
// sample (the caller streams the channel through tiles), the rest works on the channel's row of high-passed int16 samples.
struct TransientHp {
    int mem0, mem1, mx, mn;
    CB_MEM void reset() { mem0 = mem1 = mx = mn = 0; }
    // x: the input sample >> SIG_SHIFT, i: its index.  Returns the high-passed sample (opus_val16).
    CB_MEM int step(int x, int i) {
        const int y = wadd(mem0, x);
        mem0 = wsub(wadd(mem1, y), shl32(x, 1));
        mem1 = wsub(x, y >> 1);
        const int t = i < 12 ? 0 : s16(y >> 2);
        mx = imax(mx, t);
        mn = imin(mn, t);
        return t;
    }
};
// row: len high-passed samples; clobbered (pairwise energies and the two followers are built in place).  Returns mask_metric.
CB_DEV_NOINLINE int transient_finish_row(int16_t *row, int len, int mx, int mn) {
    const int len2 = len / 2;
    const int shift = 14 - celt_ilog2(1 + imax(mx, -mn));
    int mean = 0;
    CB_NOUNROLL for (int i = 0; i < len2; i++) {
        int a = row[2 * i], b = row[2 * i + 1];
        if (shift != 0) {   // SHL16 with a NEGATIVE count (maxabs == 32768): what the reference's C expression does on x86
            a = (int16_t)((unsigned)(uint16_t)a << (shift & 31));
            b = (int16_t)((unsigned)(uint16_t)b << (shift & 31));
        }
        const int x2 = s16(pshr32(wadd(mul16_16(a, a), mul16_16(b, b)), 16));
        row[i] = (int16_t)x2;
        mean = wadd(mean, x2);
    }
    int mem0 = 0;
    CB_NOUNROLL for (int i = 0; i < len2; i++) {
        mem0 = s16(mem0 + pshr32(row[i] - mem0, 4));
        row[i] = (int16_t)mem0;
    }
    mem0 = 0;
    int maxE = 0;
    CB_NOUNROLL for (int i = len2 - 1; i >= 0; i--) {
        mem0 = s16(mem0 + pshr32(row[i] - mem0, 3));
        row[i] = (int16_t)mem0;
        maxE = imax(maxE, mem0);
    }
    const int m = mul16_16(celt_sqrt(mean), celt_sqrt(mul16_16(maxE, len2 >> 1)));
    const int norm = shl32(len2, 6 + 14) / wadd(1, m >> 1);
    int unmask = 0;
    CB_NOUNROLL for (int i = 12; i < len2 - 5; i += 4) {
        const int id = imax(0, imin(127, mul16_32_q15(row[i] + 1, norm)));
        unmask += kInvTable[id];
    }
    return 64 * unmask * 4 / (6 * (len2 - 17));
}

// spreading_decision (bands.c:428-519), the statistics of X (everything before the recursive averages)
template <class TM>
CB_DEV_NOINLINE void spreading_stats_team(TM tm, const int16_t *X, int end, int C, int M, int *sum_, int *nb_, int *hf_) {
    int sum = 0, nbBands = 0, hf_sum = 0;
    const int N0 = M * kShortMdct;
    CB_NOUNROLL for (int c = 0; c < C; c++) {
        CB_NOUNROLL for (int i = 0; i < end; i++) {
            const int16_t *x = X + M * kEBands[i] + c * N0;
            const int N = M * (kEBands[i + 1] - kEBands[i]);
            if (N <= 8) continue;
            int packed = 0;
            CB_TEAM_FOR(j, N, tm) {
                int x2N = mul16_16(mul16_16_q15(x[j], x[j]), N);
                packed += (x2N < 2048) + ((x2N < 512) << 10) + ((x2N < 128) << 20);
            }
            packed = tm.sum(packed);
            const int t0 = packed & 1023, t1 = (packed >> 10) & 1023, t2 = (packed >> 20) & 1023;
            if (i > kNbEBands - 4) hf_sum += (int)udiv((unsigned)(32 * (t1 + t0)), (unsigned)N);
            int tmp = (2 * t2 >= N) + (2 * t1 >= N) + (2 * t0 >= N);
            sum += tmp * 256;
            nbBands++;
        }
    }
    *sum_ = sum; *nb_ = nbBands; *hf_ = hf_sum;
}
// ... and the rest of it
CB_DEV int spreading_finish(int sum, int nbBands, int hf_sum, int *average, int last_decision, int *hf_average, int *tapset_decision, int update_hf, int end, int C) {
    if (update_hf) {
        if (hf_sum) hf_sum = (int)udiv((unsigned)hf_sum, (unsigned)(C * (4 - kNbEBands + end)));
        *hf_average = (*hf_average + hf_sum) >> 1;
        hf_sum = *hf_average;
        if (*tapset_decision == 2) hf_sum += 4;
        else if (*tapset_decision == 0) hf_sum -= 4;
        if (hf_sum > 22) *tapset_decision = 2;
        else if (hf_sum > 18) *tapset_decision = 1;
        else *tapset_decision = 0;
    }
    sum = (int)udiv((unsigned)sum, (unsigned)nbBands);
    sum = (sum + *average) >> 1;
    *average = sum;
    sum = (3 * sum + (((3 - last_decision) << 7) + 64) + 2) >> 2;
    if (sum < 80) return kSpreadAggressive;
    if (sum < 256) return kSpreadNormal;
    if (sum < 384) return kSpreadLight;
    return kSpreadNone;
}

// Team scratch of K3: one channel's FFT buffer, then the spectrum and the tf work arrays (phase overlay)
struct TransformScratch {
    union {
        int fft[kMaxFrame];
        struct { int16_t tf_tmp[kMaxFrame], tf_tmp1[kMaxFrame]; } tf;
    } u;
    int16_t X[2 * kMaxFrame];
};

// tf_analysis metrics (celt_encoder.c:580-640) of every band at once.  tf_band_metric (celt_encoder.cuh) walks one band per lane:
// the bands are 8 to 176 values wide, so a few lanes did nearly all of it (4.5 active lanes, 30 % of K3's instructions).  Here the
// Haar steps run over the whole spectrum as one position-parallel pass (a band starts at a multiple of 2^LM, so the butterflies of
// levels < LM never straddle bands; the extra level of the long-block case pairs blocks of 2^(LM+1) values, which only the
// width-1 bands — excluded from that level — could break), and a band's L1 norm is a team sum.  Band b's running best lives in
// lane b % W.  tmp / tmp1: team scratch for the analysed channel.  Returns the lane's share of tf_sum.
template <class TM>
CB_DEV void haar_pass_team(TM tm, int16_t *X, int p0, int p1, int stride) {
    const int npairs = (p1 - p0) >> 1;
    CB_TEAM_FOR(w, npairs, tm) {
        const int blk = w / stride, i = w - blk * stride;
        const int a = p0 + blk * 2 * stride + i, b = a + stride;
        const int t1 = mul16_16(23170, X[a]);
        const int t2 = mul16_16(23170, X[b]);
        X[a] = (int16_t)pshr32(wadd(t1, t2), 15);
        X[b] = (int16_t)pshr32(wsub(t1, t2), 15);
    }
    tm.sync();
}
template <class TM>
CB_DEV int tf_metrics_team(TM tm, const int16_t *Xc, int effEnd, int LM, int isTransient, int bias, int16_t *tmp, int16_t *tmp1, int *metric) {
    constexpr int BPL = (kNbEBands + TM::W - 1) / TM::W;   // bands per lane
    int bestL1[BPL], bestLvl[BPL];
    const int total = kEBands[effEnd] << LM;
    int nnarrow = 0;                                       // the leading width-1 bands
    while (nnarrow < effEnd && band_width(nnarrow) == 1) nnarrow++;
    const int pw = kEBands[nnarrow] << LM;                 // first position of the wider bands
    CB_TEAM_FOR(j, total, tm) tmp[j] = Xc[j];
    tm.sync();
    // L1 of every band of t at weight LMw; what(b, L1) is applied by the band's lane
    auto all_bands = [&](const int16_t *t, int first, int LMw, auto what) {
        CB_NOUNROLL for (int b = first; b < effEnd; b++) {
            const int lo = kEBands[b] << LM, n = band_width(b) << LM;
            int part = 0;
            CB_TEAM_FOR(j, n, tm) part += iabs((int)t[lo + j]);
            const int L1 = tm.sum(part);
            if (tm.lane() == b % TM::W) what(BPL == 1 ? 0 : b / TM::W, mac16_32_q15(L1, LMw * bias, L1));
        }
    };
    all_bands(tmp, 0, isTransient ? LM : 0, [&](int k, int L1) { bestL1[k] = L1; bestLvl[k] = 0; });
    if (isTransient && nnarrow < effEnd) {
        CB_TEAM_FOR(j, total - pw, tm) tmp1[pw + j] = tmp[pw + j];
        tm.sync();
        haar_pass_team(tm, tmp1, pw, total, 1 << LM);
        all_bands(tmp1, nnarrow, LM + 1, [&](int k, int L1) { if (L1 < bestL1[k]) { bestL1[k] = L1; bestLvl[k] = -1; } });
    }
    CB_NOUNROLL for (int k = 0; k < LM + !isTransient; k++) {
        const int Bw = isTransient ? LM - k - 1 : k + 1;
        const int first = k < LM ? 0 : nnarrow;            // level LM: the long-block extra, not for the width-1 bands
        if (first >= effEnd) break;
        haar_pass_team(tm, tmp, k < LM ? 0 : pw, total, 1 << k);
        all_bands(tmp, first, Bw, [&](int q, int L1) { if (L1 < bestL1[q]) { bestL1[q] = L1; bestLvl[q] = k + 1; } });
    }
    int tf_sum = 0;
    CB_TEAM_FOR(i, effEnd, tm) {
        const int q = BPL == 1 ? 0 : i / TM::W;
        const int narrow = band_width(i) == 1;
        int m = isTransient ? 2 * bestLvl[q] : -2 * bestLvl[q];
        tf_sum += (isTransient ? LM : 0) - m / 2;
        if (narrow && (m == 0 || m == -2 * LM)) m -= 1;
        metric[i] = m;
    }
    return tf_sum;
}

// ---------------------------------------------------------------------------------------------------------------------------
// K3 — transform and analysis of one stream's frame (celt_encoder.c:1642-1880 plus the X-only halves of :1900-1990).
// ---------------------------------------------------------------------------------------------------------------------------
template <class TM>
CB_DEV void pipe_transform(TM tm, CbEncState *st, const PipeGeom &g, EncPipeCtx &X, EncPipeBuf &B, TransformScratch &S) {
    if (!X.code) return;
    EncVars &V = X.v;
    const CeltEncCfg &cfg = X.cfg;
    const bool L0 = tm.lane() == 0;
    const int CC = g.CC, C = cfg.C, LM = g.LM, M = 1 << LM, N = g.N, upsample = g.upsample;
    const int start = 0, end = cfg.end, effEnd = end;
    // ---- transient decision (:1642-1657) ----
    if (L0) {
        int isTransient = 0, shortBlocks = 0, tf_estimate = 0, tf_chan = 0, transient_got_disabled = 0;
        if (cfg.complexity >= 1) {
            int mask_metric = 0;
            CB_NOUNROLL for (int c = 0; c < CC; c++)
                if (V.mask_metric[c] > mask_metric) { tf_chan = c; mask_metric = V.mask_metric[c]; }
            isTransient = mask_metric > 200;
            const int tf_max = imax(0, s16(celt_sqrt(27 * mask_metric)) - 42);
            tf_estimate = s16(celt_sqrt(imax(0, wsub(shl32(mul16_16(113, imin(163, tf_max)), 14), 37312528))));
        }
        if (LM > 0 && V.ec.tell() + 3 <= V.total_bits) {
            if (isTransient) shortBlocks = M;
        } else {
            isTransient = 0;
            transient_got_disabled = 1;
        }
        V.isTransient = isTransient; V.shortBlocks = shortBlocks; V.tf_estimate = tf_estimate; V.tf_chan = tf_chan;
        V.transient_got_disabled = transient_got_disabled;
        V.secondMdct = shortBlocks && cfg.complexity >= 8;
    }
    tm.sync();
    // ---- MDCT, band energies (:1660-1690) ----
    if (V.secondMdct) {
        compute_mdcts_team(tm, 0, B.in, B.freq, C, CC, LM, upsample, S.u.fft);
        band_energies_team(tm, B.freq, X.bandE, X.bandLogE2, effEnd, end, C, LM);
        CB_TEAM_FOR(i, C * kNbEBands, tm) X.bandLogE2[i] = (int16_t)(X.bandLogE2[i] + (shl16(LM, 10) >> 1));
        tm.sync();
    }
    compute_mdcts_team(tm, V.shortBlocks, B.in, B.freq, C, CC, LM, upsample, S.u.fft);
    band_energies_team(tm, B.freq, X.bandE, X.bandLogE, effEnd, end, C, LM);
    // ---- temporal VBR, bandLogE2, transient patch (:1803-1848) ----
    if (L0) {
        if (CC == 2 && C == 1) V.tf_chan = 0;
        {
            int follow = -10240;
            int frame_avg = 0;
            const int offset = V.shortBlocks ? (shl16(LM, 10) >> 1) : 0;
            CB_NOUNROLL for (int i = start; i < end; i++) {
                follow = s16(imax(follow - 1024, X.bandLogE[i] - offset));
                if (C == 2) follow = s16(imax(follow, X.bandLogE[i + kNbEBands] - offset));
                frame_avg += follow;
            }
            frame_avg /= (end - start);
            int temporal_vbr = s16(s16(frame_avg) - s16(st->spec_avg));
            temporal_vbr = imin(3072, imax(-1536, temporal_vbr));
            st->spec_avg = s16(st->spec_avg + mul16_16_q15(655, temporal_vbr));
            V.temporal_vbr = temporal_vbr;
        }
        if (!V.secondMdct)
            CB_NOUNROLL for (int i = 0; i < C * kNbEBands; i++) X.bandLogE2[i] = X.bandLogE[i];
        V.patch = 0;
        if (LM > 0 && V.ec.tell() + 3 <= V.total_bits && !V.isTransient && cfg.complexity >= 5) {
            if (patch_transient_decision(X.bandLogE, st->oldBandE, start, end, C)) {
                V.patch = 1;
                V.isTransient = 1;
                V.shortBlocks = M;
            }
        }
    }
    tm.sync();
    if (V.patch) {
        compute_mdcts_team(tm, V.shortBlocks, B.in, B.freq, C, CC, LM, upsample, S.u.fft);
        band_energies_team(tm, B.freq, X.bandE, X.bandLogE, effEnd, end, C, LM);
        CB_TEAM_FOR(i, C * kNbEBands, tm) X.bandLogE2[i] = (int16_t)(X.bandLogE2[i] + (shl16(LM, 10) >> 1));
        if (L0) V.tf_estimate = 3277;
        tm.sync();
    }
    if (L0) {
        if (LM > 0 && V.ec.tell() + 3 <= V.total_bits) {
            EcEnc ec = V.ec;
            ec.bit_logp(V.isTransient, 3);
            V.ec = ec;
        }
        V.do_tf = V.effectiveBytes >= 15 * C && start == 0 && cfg.complexity >= 2;
    }
    tm.sync();
    // ---- band normalisation (:1856) ----
    normalise_bands_team(tm, B.freq, S.X, X.bandE, effEnd, C, M, LM, X.band_g, X.band_shift);
    // ---- tf_analysis metrics (:1858-1880): one band per lane ----
    const int isTransient = V.isTransient;
    const int shortBlocks = V.shortBlocks;
    if (V.do_tf) {
        const int bias = mul16_16_q14(1311, imax(-4096, 8192 - V.tf_estimate));
        const int16_t *Xc = S.X + V.tf_chan * N;
        int tf_sum = tf_metrics_team(tm, Xc, effEnd, LM, isTransient, bias, S.u.tf.tf_tmp, S.u.tf.tf_tmp1, X.metric);
        tf_sum = tm.sum(tf_sum);
        if (L0) X.tf_sum_team = tf_sum;
    }
    // ---- statistics of X for the decisions K4 takes: spreading (bands.c:428-519), stereo_analysis (:840-870), trim (:756-838) ----
    {
        int sum = 0, nb = 0, hf = 0;
        const int skip = M * (kEBands[effEnd] - kEBands[effEnd - 1]) <= 8;
        const int may_spread = !(shortBlocks || cfg.complexity < 3 || V.nbAvailableBytes < 10 * C);
        if (!skip && may_spread) spreading_stats_team(tm, S.X, effEnd, C, M, &sum, &nb, &hf);
        if (L0) { X.spread_sum = sum; X.spread_nb = nb; X.spread_hf = hf; X.spread_skip = skip; }
    }
    if (C == 2) {
        if (LM != 0) {
            int lr = 0, ms = 0;
            CB_TEAM_FOR(j, kEBands[13] << LM, tm) {
                int Lv = S.X[j], Rv = S.X[N + j];
                int Mv = Lv + Rv, Sv = Lv - Rv;
                lr = wadd(lr, iabs(Lv) + iabs(Rv));
                ms = wadd(ms, iabs(Mv) + iabs(Sv));
            }
            lr = tm.sum(lr); ms = tm.sum(ms);
            if (L0) { X.st_sumLR = lr; X.st_sumMS = ms; }
        }
        CB_NOUNROLL for (int i = 0; i < end; i++) {
            const int partial = team_inner16(tm, &S.X[kEBands[i] << LM], &S.X[N + (kEBands[i] << LM)], band_width(i) << LM);
            if (L0) X.trim_xc[i] = (int16_t)s16(partial >> 18);
        }
    }
    tm.sync();
    CB_TEAM_FOR(i, C * N, tm) B.X[i] = S.X[i];
}

// alloc_trim_analysis (celt_encoder.c:756-838) from the per-band correlations K3 left
CB_DEV_NOINLINE int alloc_trim_finish(const int16_t *trim_xc, const int16_t *bandLogE, int end, int C, int *stereo_saving, int tf_estimate, int intensity) {
    int diff = 0;
    int trim = 1280;
    if (C == 2) {
        int sum = 0;
        CB_NOUNROLL for (int i = 0; i < 8; i++) sum = s16(sum + trim_xc[i]);
        sum = mul16_16_q15(4096, sum);
        sum = imin(1024, iabs(sum));
        int minXC = sum;
        CB_NOUNROLL for (int i = 8; i < intensity; i++) minXC = imin(minXC, iabs((int)trim_xc[i]));
        minXC = imin(1024, iabs(minXC));
        int logXC = celt_log2(1049625 - mul16_16(sum, sum));
        int logXC2 = imax(logXC >> 1, celt_log2(1049625 - mul16_16(minXC, minXC)));
        logXC = s16(pshr32(logXC - 6144, 2));
        logXC2 = s16(pshr32(logXC2 - 6144, 2));
        trim = s16(trim + imax(-1024, mul16_16_q15(24576, logXC)));
        *stereo_saving = s16(imin(*stereo_saving + 64, -(logXC2 >> 1)));
    }
    CB_NOUNROLL for (int c = 0; c < C; c++)
        CB_NOUNROLL for (int i = 0; i < end - 1; i++) diff += bandLogE[i + c * kNbEBands] * (2 + 2 * i - end);
    diff /= C * (end - 1);
    trim = s16(trim - imax(-512, imin(512, ((diff + 1024) >> 2) / 6)));
    trim = s16(trim - 2 * (tf_estimate >> 6));
    int trim_index = pshr32(trim, 8);
    return imax(0, imin(10, trim_index));
}

// ---------------------------------------------------------------------------------------------------------------------------
// K4 — the scalar decisions between the analysis and the band loop (celt_encoder.c:1858-2106): one thread per stream.
// ---------------------------------------------------------------------------------------------------------------------------
CB_DEV_NOINLINE void pipe_decide(CbEncState *st, const PipeGeom &g, EncPipeCtx &X) {
    if (!X.code) return;
    EncVars &V = X.v;
    const CeltEncCfg cfg = X.cfg;
    const int C = cfg.C, LM = g.LM, M = 1 << LM, start = 0, end = cfg.end, effEnd = end;
    const int isTransient = V.isTransient, shortBlocks = V.shortBlocks;
    (void)M;
    // ---- tf_analysis: the Viterbi search over the band metrics ----
    if (V.do_tf) {
        int lambda;
        if (V.effectiveBytes < 40) lambda = 12;
        else if (V.effectiveBytes < 60) lambda = 6;
        else if (V.effectiveBytes < 100) lambda = 4;
        else lambda = 3;
        lambda *= 2;
        V.tf_select = tf_viterbi(X.metric, effEnd, isTransient, X.tf_res, lambda, LM);
        CB_NOUNROLL for (int i = effEnd; i < end; i++) X.tf_res[i] = X.tf_res[effEnd - 1];
        V.tf_sum = X.tf_sum_team;
    } else {
        V.tf_sum = 0;
        CB_NOUNROLL for (int i = 0; i < end; i++) X.tf_res[i] = isTransient;
        V.tf_select = 0;
    }
    // ---- coarse energy, tf flags, spread (:1882-1925) ----
    EcEnc ec = V.ec;
    quant_coarse_energy(start, end, effEnd, X.bandLogE, st->oldBandE, (unsigned)V.total_bits, X.error, ec, C, LM, V.nbAvailableBytes,
                        cfg.force_intra, &st->delayedIntra, cfg.complexity >= 4, cfg.loss_rate, X.coarse);
    tf_encode(start, end, isTransient, X.tf_res, LM, V.tf_select, ec);
    if (ec.tell() + 4 <= V.total_bits) {
        if (shortBlocks || cfg.complexity < 3 || V.nbAvailableBytes < 10 * C || start != 0) {
            st->spread_decision = cfg.complexity == 0 ? kSpreadNone : kSpreadNormal;
        } else if (X.spread_skip) {
            st->spread_decision = kSpreadNone;
        } else {
            int average = st->tonal_average, hf_average = st->hf_average, tapset_decision = st->tapset_decision;
            const int dec = spreading_finish(X.spread_sum, X.spread_nb, X.spread_hf, &average, st->spread_decision, &hf_average, &tapset_decision,
                                             V.pf_on && !shortBlocks, effEnd, C);
            st->tonal_average = average; st->hf_average = hf_average; st->tapset_decision = tapset_decision;
            st->spread_decision = dec;
        }
        ec.icdf(st->spread_decision, kSpreadIcdf, 5);
    }
    // ---- dynalloc (:1927-1972) ----
    V.maxDepth = dynalloc_analysis(X.bandLogE, X.bandLogE2, start, end, C, X.offsets, cfg.lsb_depth, isTransient, cfg.vbr, cfg.constrained_vbr, LM,
                                   V.effectiveBytes, &V.tot_boost);
    init_caps(X.cap, LM, C);
    int dynalloc_logp = 6;
    int total_bits = V.total_bits << kBitRes;
    int total_boost = 0;
    int tell = (int)ec.tell_frac();
    CB_NOUNROLL for (int i = start; i < end; i++) {
        const int width = C * band_width(i) << LM;
        const int quanta = imin(width << kBitRes, imax(6 << kBitRes, width));
        int loop_logp = dynalloc_logp;
        int boost = 0;
        int j;
        CB_NOUNROLL for (j = 0; tell + (loop_logp << kBitRes) < total_bits - total_boost && boost < X.cap[i]; j++) {
            const int flag = j < X.offsets[i];
            ec.bit_logp(flag, (unsigned)loop_logp);
            tell = (int)ec.tell_frac();
            if (!flag) break;
            boost += quanta;
            total_boost += quanta;
            loop_logp = 1;
        }
        if (j) dynalloc_logp = imax(2, dynalloc_logp - 1);
        X.offsets[i] = boost;
    }
    // ---- stereo decisions (:1974-1990), allocation trim (:1992-2000) ----
    int dual_stereo = 0;
    if (C == 2) {
        if (LM != 0) {
            int sumLR = wadd(1, X.st_sumLR);
            int sumMS = wadd(1, X.st_sumMS);
            sumMS = mul16_32_q15(23170, sumMS);
            int thetas = 13;
            if (LM <= 1) thetas -= 8;
            dual_stereo = mul16_32_q15((kEBands[13] << (LM + 1)) + thetas, sumMS) > mul16_32_q15(kEBands[13] << (LM + 1), sumLR);
        }
        st->intensity = hysteresis_decision(s16(V.equiv_rate / 1000), kIntensityThresholds, kIntensityHisteresis, 21, st->intensity);
        st->intensity = imin(end, imax(start, st->intensity));
    }
    int alloc_trim = 5;
    if (tell + (6 << kBitRes) <= total_bits - total_boost) {
        int stereo_saving = st->stereo_saving;
        alloc_trim = alloc_trim_finish(X.trim_xc, X.bandLogE, end, C, &stereo_saving, V.tf_estimate, st->intensity);
        st->stereo_saving = stereo_saving;
        ec.icdf(alloc_trim, kTrimIcdf, 7);
        tell = (int)ec.tell_frac();
    }
    // ---- rate control (:2002-2087) ----
    int nbCompressedBytes = V.nbCompressedBytes;
    int nbAvailableBytes = V.nbAvailableBytes;
    const int nbFilledBytes = V.nbFilledBytes;
    const int vbr_rate = V.vbr_rate;
    if (vbr_rate > 0) {
        const int lm_diff = kMaxLM - LM;
        nbCompressedBytes = imin(nbCompressedBytes, 1275 >> (3 - LM));
        int base_target = vbr_rate - ((40 * C + 20) << kBitRes);
        if (cfg.constrained_vbr) base_target += (st->vbr_offset >> lm_diff);
        int target = compute_vbr(base_target, LM, V.equiv_rate, st->lastCodedBands, C, st->intensity, cfg.constrained_vbr, st->stereo_saving, V.tot_boost,
                                 V.tf_estimate, V.maxDepth, cfg.variable_duration, V.temporal_vbr);
        target = target + tell;
        const int min_allowed = ((tell + total_boost + (1 << (kBitRes + 3)) - 1) >> (kBitRes + 3)) + 2 - nbFilledBytes;
        nbAvailableBytes = (target + (1 << (kBitRes + 2))) >> (kBitRes + 3);
        nbAvailableBytes = imax(min_allowed, nbAvailableBytes);
        nbAvailableBytes = imin(nbCompressedBytes, nbAvailableBytes + nbFilledBytes) - nbFilledBytes;
        int delta = target - vbr_rate;
        target = nbAvailableBytes << (kBitRes + 3);
        if (V.silence) {
            nbAvailableBytes = 2;
            target = 2 * 8 << kBitRes;
            delta = 0;
        }
        int alpha;
        if (st->vbr_count < 970) {
            st->vbr_count++;
            alpha = s16(celt_rcp(shl32(st->vbr_count + 20, 16)));
        } else {
            alpha = 33;
        }
        if (cfg.constrained_vbr) st->vbr_reservoir += target - vbr_rate;
        if (cfg.constrained_vbr) {
            st->vbr_drift += mul16_32_q15(alpha, (delta * (1 << lm_diff)) - st->vbr_offset - st->vbr_drift);
            st->vbr_offset = -st->vbr_drift;
        }
        if (cfg.constrained_vbr && st->vbr_reservoir < 0) {
            const int adjust = (-st->vbr_reservoir) / (8 << kBitRes);
            nbAvailableBytes += V.silence ? 0 : adjust;
            st->vbr_reservoir = 0;
        }
        nbCompressedBytes = imin(nbCompressedBytes, nbAvailableBytes + nbFilledBytes);
        ec.shrink((unsigned)nbCompressedBytes);
    }
    // ---- allocation, fine energy (:2089-2133) ----
    int bits = ((nbCompressedBytes * 8) << kBitRes) - (int)ec.tell_frac() - 1;
    const int anti_collapse_rsv = isTransient && LM >= 2 && bits >= ((LM + 2) << kBitRes) ? (1 << kBitRes) : 0;
    bits -= anti_collapse_rsv;
    const int signalBandwidth = end - 1;
    int balance = 0;
    int intensity = st->intensity;
    AllocEncIo io{ec, start, st->lastCodedBands, signalBandwidth, LM};
    const int codedBands = compute_allocation(io, X.alloc, start, end, X.offsets, X.cap, alloc_trim, &intensity, &dual_stereo, bits, &balance, X.pulses,
                                              X.fine_quant, X.fine_priority, C, LM);
    st->intensity = intensity;
    if (st->lastCodedBands) st->lastCodedBands = imin(st->lastCodedBands + 1, imax(st->lastCodedBands - 1, codedBands));
    else st->lastCodedBands = codedBands;
    quant_fine_energy(start, end, st->oldBandE, X.error, X.fine_quant, ec, C);
    V.ec = ec;
    V.total_bits = total_bits;
    V.nbCompressedBytes = nbCompressedBytes;
    V.anti_collapse_rsv = anti_collapse_rsv;
    V.balance = balance;
    V.codedBands = codedBands;
    V.dual_stereo = dual_stereo;
    V.alloc_trim = alloc_trim;
}

// Team scratch of K5
struct BandScratch {
    int16_t X[2 * kMaxFrame];
    PvqScratch pvq;
    int16_t had_tmp[176];
};

// the frame's tail: anti-collapse bit, energy finalise, state update, ec_enc_done (celt_encoder.c:2135-2262), then the Opus layer's
// (opus_encoder.c:1927-1971): TOC, final range, CBR padding.  Scalar.  Returns the packet length.
CB_DEV_NOINLINE int pipe_finish(CbEncState *st, const PipeGeom &g, const EncPlan &pl, EncPipeCtx &X, uint8_t *out) {
    EncVars &V = X.v;
    const CeltEncCfg &cfg = X.cfg;
    const int CC = g.CC, C = cfg.C, start = 0, end = cfg.end;
    const int isTransient = V.isTransient;
    EcEnc ec = V.ec;
    const int nbCompressedBytes = V.nbCompressedBytes;
    if (V.anti_collapse_rsv > 0) {
        const int anti_collapse_on = st->consec_transient < 2;
        ec.bits((unsigned)anti_collapse_on, 1);
    }
    quant_energy_finalise(start, end, st->oldBandE, X.error, X.fine_quant, X.fine_priority, nbCompressedBytes * 8 - ec.tell(), ec, C);
    if (V.silence)
        CB_NOUNROLL for (int i = 0; i < C * kNbEBands; i++) st->oldBandE[i] = -28672;
    st->prefilter_period = V.pitch_index;
    st->prefilter_gain = V.gain1;
    st->prefilter_tapset = V.prefilter_tapset;
    if (CC == 2 && C == 1)
        CB_NOUNROLL for (int i = 0; i < kNbEBands; i++) st->oldBandE[kNbEBands + i] = st->oldBandE[i];
    if (!isTransient) {
        CB_NOUNROLL for (int i = 0; i < CC * kNbEBands; i++) { st->oldLogE2[i] = st->oldLogE[i]; st->oldLogE[i] = st->oldBandE[i]; }
    } else {
        CB_NOUNROLL for (int i = 0; i < CC * kNbEBands; i++) st->oldLogE[i] = (int16_t)imin((int)st->oldLogE[i], (int)st->oldBandE[i]);
    }
    CB_NOUNROLL for (int c = 0; c < CC; c++) {
        CB_NOUNROLL for (int i = end; i < kNbEBands; i++) {
            st->oldBandE[c * kNbEBands + i] = 0;
            st->oldLogE[c * kNbEBands + i] = st->oldLogE2[c * kNbEBands + i] = -28672;
        }
    }
    if (isTransient || V.transient_got_disabled) st->consec_transient++;
    else st->consec_transient = 0;
    st->rng = ec.rng;
    ec.done();
    int ret = ec.error ? OPUS_INTERNAL_ERROR_ : nbCompressedBytes;
    if (ret < 0) return OPUS_INTERNAL_ERROR_;
    out[0] = (uint8_t)gen_toc(pl.mode, g.Fs / g.fsz, pl.curr_bandwidth, pl.stream_channels);
    st->rangeFinal = ec.rng;
    ret += 1;
    if (!pl.use_vbr) {
        packet_pad_single(out, ret, pl.max_data_bytes);
        ret = pl.max_data_bytes;
    }
    return ret;
}

// ---------------------------------------------------------------------------------------------------------------------------
// K5 — residual quantisation (quant_all_bands, celt_encoder.c:2208) with the team splitting the vector work, then the tail.
// ---------------------------------------------------------------------------------------------------------------------------
template <class TM>
CB_DEV int pipe_bands(TM tm, CbEncState *st, const PipeGeom &g, const EncPlan &pl, EncPipeCtx &X, EncPipeBuf &B, BandScratch &S, uint8_t *out) {
    if (!X.code) return X.v.ret;
    EncVars &V = X.v;
    const CeltEncCfg &cfg = X.cfg;
    const int C = cfg.C, N = g.N, LM = g.LM;
    CB_TEAM_FOR(i, C * N, tm) S.X[i] = B.X[i];
    tm.sync();
    {
        EcEnc ec = V.ec;
        quant_all_bands_enc(tm, 0, cfg.end, S.X, C == 2 ? S.X + N : nullptr, X.bandE, X.pulses, V.shortBlocks, st->spread_decision, V.dual_stereo,
                            st->intensity, X.tf_res, V.nbCompressedBytes * (8 << kBitRes) - V.anti_collapse_rsv, V.balance, ec, LM, V.codedBands, &S.pvq,
                            S.had_tmp);
        tm.sync();
        if (tm.lane() == 0) V.ec = ec;
    }
    tm.sync();
    if (tm.lane() == 0) V.ret = pipe_finish(st, g, pl, X, out);
    tm.sync();
    return V.ret;
}

// ---- K5a..K5d: the band loop as prep / chain-S / leaves / chain-X (celt_enc_bandpipe.cuh) ------------------------------------------
struct PrepScratch {
    int16_t Xall[kXallStride];
    int16_t tmp[176];
    int seg[16];
};
template <class TM>
CB_DEV void pipe_band_prep(TM tm, const CbEncState *st, const PipeGeom &g, const EncPipeCtx &X, const EncPipeBuf &B, BandPrep &P, int16_t *XallG,
                           PrepScratch &S) {
    if (!X.code) return;
    const int C = X.cfg.C, N = g.N;
    CB_TEAM_FOR(i, N, tm) S.Xall[i] = B.X[i];
    if (C == 2) CB_TEAM_FOR(i, N, tm) S.Xall[kMaxFrame + i] = B.X[N + i];
    tm.sync();
    band_prep_team(tm, S.Xall, X.bandE, X.tf_res, X.cfg.end, C, g.LM, X.v.shortBlocks, X.v.dual_stereo, st->intensity, P, S.tmp, S.seg);
    tm.sync();
    CB_NOUNROLL for (int v = 0; v < 3; v++) {
        if (v > 0 && C == 1) break;
        CB_TEAM_FOR(i, N, tm) XallG[v * kMaxFrame + i] = S.Xall[v * kMaxFrame + i];
    }
}
CB_DEV_NOINLINE void pipe_band_spec(const CbEncState *st, const PipeGeom &g, const EncPipeCtx &X, const BandPrep &P, LeafList &L) {
    if (!X.code) { L.count = 0; return; }
    const EncVars &V = X.v;
    SpecPolicy p;
    p.ec.from(V.ec);
    p.list = &L;
    p.cur_band = -1;
    L.count = 0;
    L.overflow = 0;
    band_walk(p, P, X.cfg.end, X.cfg.C, X.pulses, V.shortBlocks, st->spread_decision, V.dual_stereo, st->intensity, X.tf_res,
              V.nbCompressedBytes * (8 << kBitRes) - V.anti_collapse_rsv, V.balance, g.LM, V.codedBands);
}
struct LeafScratch {
    int16_t V[176];
    PvqScratch pvq;
};
template <class TM>
CB_DEV void pipe_leaves(TM tm, const CbEncState *st, const EncPipeCtx &X, LeafList &L, const int16_t *XallG, LeafScratch &S) {
    if (!X.code) return;
    const int spread = st->spread_decision;
    const int cnt = L.count;
    CB_NOUNROLL for (int k = 0; k < cnt; k++) {
        const LeafTask t = L.task[k];
        CB_TEAM_FOR(j, t.N, tm) S.V[j] = XallG[t.off + j];
        tm.sync();
        const unsigned idx = leaf_quant_team(tm, S.V, t.N, t.K, spread, t.B, S.pvq);
        if (tm.lane() == 0) L.index[k] = idx;
        tm.sync();
    }
}
CB_DEV_NOINLINE int pipe_band_exact_finish(CbEncState *st, const PipeGeom &g, const EncPlan &pl, EncPipeCtx &X, const BandPrep &P, const LeafList &L,
                                           int16_t *XallG, uint8_t *out, int *misses) {
    if (!X.code) return X.v.ret;
    EncVars &V = X.v;
    PvqScratch ps;
    ExactPolicy p;
    p.ec = V.ec;
    p.list = &L;
    p.Xall = XallG;
    p.prep = &P;
    p.ps = &ps;
    p.cursor = 0;
    p.misses = 0;
    band_walk(p, P, X.cfg.end, X.cfg.C, X.pulses, V.shortBlocks, st->spread_decision, V.dual_stereo, st->intensity, X.tf_res,
              V.nbCompressedBytes * (8 << kBitRes) - V.anti_collapse_rsv, V.balance, g.LM, V.codedBands);
    V.ec = p.ec;
    if (misses) *misses += p.misses;
    V.ret = pipe_finish(st, g, pl, X, out);
    return V.ret;
}

// ---- the band loop as prep + ONE walk with the real coder, leaves searched inline by the team (no speculation) -----------------------
struct WalkScratch {
    int16_t leaf[176];
    PvqScratch pvq;
    WalkShared sh;
};
template <class TM>
CB_DEV int pipe_band_inline_finish(TM tm, CbEncState *st, const PipeGeom &g, const EncPlan &pl, EncPipeCtx &X, const BandPrep &P, const int16_t *XallG,
                                   WalkScratch &S, uint8_t *out, unsigned sync_mask = 0x1fffffu) {
    if (!X.code) {
        CB_NOUNROLL for (int i = 0; i < kNbEBands; i++)
            if ((sync_mask >> i) & 1u) tm.phase();
        return X.v.ret;
    }
    EncVars &V = X.v;
    const int C = X.cfg.C;
#if defined(CB_WALK_LOCAL)
    WalkShared lsh;            // A/B: the scalar state as 32 private copies on the stack
    WalkShared &SH = lsh;
#else
    WalkShared &SH = S.sh;
#endif
    SH.ec = V.ec;              // every lane stores the same values
    tm.sync();
    InlinePolicy<TM> p{tm, &SH, SH.ec, XallG, S.leaf, &P, &S.pvq, sync_mask};
    band_walk(p, P, X.cfg.end, C, X.pulses, V.shortBlocks, st->spread_decision, V.dual_stereo, st->intensity, X.tf_res,
              V.nbCompressedBytes * (8 << kBitRes) - V.anti_collapse_rsv, V.balance, g.LM, V.codedBands);
    tm.sync();
    if (tm.lane() == 0) {
        V.ec = SH.ec;
        V.ret = pipe_finish(st, g, pl, X, out);
    }
    tm.sync();
    return V.ret;
}

}  // namespace cb
