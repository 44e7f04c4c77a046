// celt_plc.cuh — packet-loss concealment for the CELT decoder (stage B, one team per stream).
//
// Restates opus-fix/celt/celt_decoder.c:399-711: celt_plc_pitch_search (:399-413) and celt_decode_lost (:415-711) with both
// branches — noise-based concealment / comfort noise after 5 consecutive losses (:447-498) and pitch-based concealment in the
// excitation domain (:499-707: LPC analysis of the last 1024 samples, celt_fir to the excitation, decay estimate, periodic
// extrapolation with per-period attenuation, celt_iir back to the signal domain, energy guard, pre-filtered TDAC tail) —
// plus celt_fir / celt_iir (celt/celt_lpc.c:93-231) and the windowed _celt_autocorr / _celt_lpc of order 24 it relies on.
// the part of opus_decode's contract the reference tests exercise with NULL packets
// (tests/test_opus_decode.c:125-132, opus_demo -loss).
//
// A lost frame is rare, so only what is cheap to spread is spread over the team (sample-parallel loops, sums); the two
// recursions (Levinson, the synthesis IIR) run on one lane.
#pragma once
#include "celt_decoder.cuh"
#include "celt_pitch.cuh"

namespace cb {

enum { kPlcPitchLagMax = 720, kPlcPitchLagMin = 100, kMaxPeriod = 1024 };

// Per-stream concealment scratch (HBM; touched only by lost frames).
struct PlcScratch {
    int16_t lp_raw[kDecBuf / 2], lp_buf[kDecBuf / 2];   // pitch_downsample staging / output; lp_raw doubles as autocorr scratch
    int16_t x_lp4[336], y_lp4[512];
    int xcorr[312], syy[312];
    int16_t exc[kMaxPeriod];                             // excitation of the channel being concealed
    int16_t fir_out[kMaxPeriod];
    int16_t iir_hist[kMaxFrame + kOverlap + kLpcOrder];  // negated, rounded synthesis-filter outputs (celt_lpc.c:181-215)
    int etmp[kOverlap];
    int16_t X[2 * kMaxFrame];                            // noise branch: normalised spectrum
};

// celt_decode_lost for one frame of `frame_size` samples per channel at the API rate; `end` = the end band the Opus layer set
// from the last good packet's bandwidth (opus_decoder.c:431-450).  The concealed (pre-de-emphasis) signal of channel c is
// staged to sig[c] for stage C.  Returns samples per channel at the API rate.
template <class TM>
CB_DEV_NOINLINE int celt_decode_lost_frame(TM tm, CbDecState *st, SynthScratch &S, PlcScratch &P, int *const *sig, int frame_size, int end) {
    const int C = st->channels;
    const int N = frame_size * st->downsample;
    int LM;
    for (LM = 0; LM <= kMaxLM; LM++)
        if (kShortMdct << LM == N) break;
    if (LM > kMaxLM) return OPUS_BAD_ARG_;
    int *decode_mem[2], *out_syn[2];
    for (int c = 0; c < C; c++) {
        decode_mem[c] = st->decode_mem + c * CB_DEC_MEM;
        out_syn[c] = decode_mem[c] + kDecBuf - N;
    }
    int16_t *oldBandE = st->oldEBands, *backgroundLogE = st->backgroundLogE;
    const int loss_count = st->loss_count;
    const int start = 0;
    const bool noise_based = loss_count >= 5 || start != 0;
    const bool L0 = tm.lane() == 0;
    tm.sync();
    if (noise_based) {
        const int effEnd = imax(start, imin(end, kNbEBands));
        const int decay = loss_count == 0 ? 1536 : 512;   // QCONST16(1.5f / .5f, DB_SHIFT)
        CB_TEAM_FOR(w, C * kNbEBands, tm) {
            const int i = w % kNbEBands;
            if (i >= start && i < end) oldBandE[w] = (int16_t)imax((int)backgroundLogE[w], oldBandE[w] - decay);
        }
        unsigned seed = st->rng;
        int16_t *X = P.X;
        for (int c = 0; c < C; c++) {
            for (int i = start; i < effEnd; i++) {
                const int boffs = N * c + (kEBands[i] << LM);
                const int blen = band_width(i) << LM;
                // every lane steps the generator so the seed stays team-uniform; lane j%W stores slot j
                CB_NOUNROLL for (int j = 0; j < blen; j++) {
                    seed = lcg_rand(seed);
                    if ((j % TM::W) == tm.lane()) X[boffs + j] = (int16_t)((int)seed >> 20);
                }
                tm.sync();
                renormalise_vector(tm, X + boffs, blen, 32767);
            }
        }
        if (L0) st->rng = seed;
        for (int c = 0; c < C; c++) history_shift(tm, decode_mem[c], N);
        celt_synthesis_team(tm, S, X, out_syn, oldBandE, start, effEnd, C, C, 0, LM, st->downsample, 0);
    } else {
        int fade = 32767;
        int pitch_index;
        if (loss_count == 0) {
            // celt_plc_pitch_search (:399-413)
            pitch_downsample_team(tm, decode_mem[0], decode_mem[C - 1], kDecBuf, C, P.lp_raw, P.lp_buf);
            pitch_index = pitch_search_team(tm, P.lp_buf + (kPlcPitchLagMax >> 1), P.lp_buf, kDecBuf - kPlcPitchLagMax,
                                            kPlcPitchLagMax - kPlcPitchLagMin, P.x_lp4, P.y_lp4, P.xcorr, P.syy);
            pitch_index = kPlcPitchLagMax - pitch_index;
            if (L0) st->last_pitch_index = pitch_index;
        } else {
            pitch_index = st->last_pitch_index;
            fade = 26214;   // QCONST16(.8f,15)
        }
        const int pf_period = st->postfilter_period, pf_gain = st->postfilter_gain, pf_tapset = st->postfilter_tapset;
        for (int c = 0; c < C; c++) {
            int *buf = decode_mem[c];
            int16_t *exc = P.exc;
            int16_t *lpc = st->lpc + c * kLpcOrder;
            CB_TEAM_FOR(i, kMaxPeriod, tm) exc[i] = (int16_t)round16(buf[kDecBuf - kMaxPeriod + i], 12);
            tm.sync();
            if (loss_count == 0) {
                if (L0) {
                    int ac[kLpcOrder + 1];
                    celt_autocorr(exc, ac, kWindow120, kOverlap, kLpcOrder, kMaxPeriod, P.lp_raw);
                    ac[0] = wadd(ac[0], ac[0] >> 13);
                    for (int i = 1; i <= kLpcOrder; i++) ac[i] = wsub(ac[i], mul16_32_q15(2 * i * i, ac[i]));
                    celt_lpc(lpc, ac, kLpcOrder);
                }
                tm.sync();
            }
            const int exc_length = imin(2 * pitch_index, kMaxPeriod);
            // celt_fir (celt_lpc.c:93-147): y[i] = sat16(x[i] + PSHR32(sum_k lpc[k] * x[i-1-k], SIG_SHIFT)); the history before
            // the region is ROUND16 of the same buffer, i.e. x[i] = ROUND16(buf[kDecBuf - exc_length + i]) for every i >= -24
            CB_TEAM_FOR(i, exc_length, tm) {
                int sum = 0;
                CB_NOUNROLL for (int k = 0; k < kLpcOrder; k++)
                    sum = mac16_16(sum, lpc[k], round16(buf[kDecBuf - exc_length + i - 1 - k], 12));
                P.fir_out[i] = (int16_t)sat16(wadd(exc[kMaxPeriod - exc_length + i], pshr32(sum, 12)));
            }
            tm.sync();
            CB_TEAM_FOR(i, exc_length, tm) exc[kMaxPeriod - exc_length + i] = P.fir_out[i];
            tm.sync();
            // decay of the excitation energy over the last two half-windows (:569-588)
            int decay;
            {
                const int shift = imax(0, 2 * celt_zlog2(team_maxabs16(tm, &exc[kMaxPeriod - exc_length], exc_length)) - 20);
                const int decay_length = exc_length >> 1;
                int e1 = 0, e2 = 0;
                CB_TEAM_FOR(i, decay_length, tm) {
                    int e = exc[kMaxPeriod - decay_length + i];
                    e1 = wadd(e1, mul16_16(e, e) >> shift);
                    e = exc[kMaxPeriod - 2 * decay_length + i];
                    e2 = wadd(e2, mul16_16(e, e) >> shift);
                }
                int E1 = wadd(1, tm.sum(e1));
                const int E2 = wadd(1, tm.sum(e2));
                E1 = imin(E1, E2);
                decay = s16(celt_sqrt(frac_div32(E1 >> 1, E2)));
            }
            history_shift(tm, buf, N, kDecBuf - N);
            // periodic extrapolation (:596-624)
            const int extrapolation_offset = kMaxPeriod - pitch_index;
            const int extrapolation_len = N + kOverlap;
            int s1 = 0;
            {
                const int att0 = s16(mul16_16_q15(fade, decay));
                CB_TEAM_FOR(i, extrapolation_len, tm) {
                    const int w = i / pitch_index, j = i - w * pitch_index;
                    int att = att0;
                    CB_NOUNROLL for (int k = 0; k < w; k++) att = s16(mul16_16_q15(att, decay));
                    buf[kDecBuf - N + i] = shl32(s16(mul16_16_q15(att, exc[extrapolation_offset + j])), 12);
                    const int tmp = round16(buf[kDecBuf - kMaxPeriod - N + extrapolation_offset + j], 12);
                    s1 = wadd(s1, mul16_16(tmp, tmp) >> 8);
                }
            }
            const int S1 = tm.sum(s1);
            tm.sync();
            // celt_iir (celt_lpc.c:151-231): recursive, one lane.  The reference keeps the NEGATED rounded outputs as its history
            // (y[i+ord] = -ROUND16(sum)) and adds den * history; extrapolation_len is a multiple of 4, so only that form runs.
            if (L0) {
                int16_t *h = P.iir_hist;   // h[ord + i] = -ROUND16(out[i]); h[0..ord) from the samples before the region
                for (int i = 0; i < kLpcOrder; i++) h[i] = (int16_t)(-round16(buf[kDecBuf - N - 1 - (kLpcOrder - 1 - i)], 12));
                CB_NOUNROLL for (int i = 0; i < extrapolation_len; i++) {
                    int sum = buf[kDecBuf - N + i];
                    CB_NOUNROLL for (int k = 0; k < kLpcOrder; k++) sum = mac16_16(sum, lpc[k], h[kLpcOrder + i - 1 - k]);
                    h[kLpcOrder + i] = (int16_t)(-round16(sum, 12));
                    buf[kDecBuf - N + i] = sum;
                }
            }
            tm.sync();
            // energy guard (:641-671)
            {
                int s2 = 0;
                CB_TEAM_FOR(i, extrapolation_len, tm) {
                    const int tmp = round16(buf[kDecBuf - N + i], 12);
                    s2 = wadd(s2, mul16_16(tmp, tmp) >> 8);
                }
                const int S2 = tm.sum(s2);
                tm.sync();
                if (!(S1 > (S2 >> 2))) {
                    CB_TEAM_FOR(i, extrapolation_len, tm) buf[kDecBuf - N + i] = 0;
                } else if (S1 < S2) {
                    const int ratio = s16(celt_sqrt(frac_div32(wadd(S1 >> 1, 1), wadd(S2, 1))));
                    CB_TEAM_FOR(i, extrapolation_len, tm) {
                        const int g = i < kOverlap ? s16(32767 - mul16_16_q15(kWindow120[i], 32767 - ratio)) : ratio;
                        buf[kDecBuf - N + i] = mul16_32_q15(g, buf[kDecBuf - N + i]);
                    }
                }
                tm.sync();
            }
            // pre-filter the overlap so the post-filter of the next frame undoes it, then simulate the TDAC (:673-687)
            if (L0) comb_filter_fir(P.etmp, buf + kDecBuf, pf_period, pf_period, kOverlap, -pf_gain, -pf_gain, pf_tapset, pf_tapset, 0);
            tm.sync();
            CB_TEAM_FOR(i, kOverlap / 2, tm)
                buf[kDecBuf + i] = wadd(mul16_32_q15(kWindow120[i], P.etmp[kOverlap - 1 - i]), mul16_32_q15(kWindow120[kOverlap - i - 1], P.etmp[i]));
            tm.sync();
        }
    }
    if (L0) st->loss_count = loss_count + 1;
    tm.sync();
    for (int c = 0; c < C; c++) {
        const int *src = out_syn[c];
        int *dst = sig[c];
        CB_TEAM_FOR(j, N, tm) dst[j] = src[j];
    }
    tm.sync();
    return N / st->downsample;
}

}  // namespace cb
