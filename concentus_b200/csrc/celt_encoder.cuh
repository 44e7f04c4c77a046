// celt_encoder.cuh — one CELT frame, encoder side: analysis, transform and the bitstream writer.
//
// Restates opus-fix/celt/celt_encoder.c:227-378 (transient_analysis), :380-414 (patch_transient_decision), :418-461
// (compute_mdcts), :464-535 (celt_preemphasis), :539-712 (l1_metric, tf_analysis), :756-838 (alloc_trim_analysis), :840-870
// (stereo_analysis), :873-930 (medians), :932-1064 (dynalloc_analysis), :1067-1192 (run_prefilter), :1194-1312 (compute_vbr)
// and :1379-2273 (celt_encode_with_ec) for the standard 48 kHz mode, FIXED_POINT, no CUSTOM_MODES, no RESYNTH, no LFE,
// no surround energy mask (the Opus layer of this engine never sets them).
//
// Execution model (one TEAM per stream, see celt_simt.cuh):
//   * every data-parallel loop — pre-emphasis, pitch down-sampling and correlation, comb pre-filter, MDCT, band energies,
//     normalisation, tf / spreading / stereo metrics — is strided over the lanes, reductions are wrapping 32-bit sums
//     (order-free, so the lane split cannot change a bit);
//   * everything order-dependent — the range coder, energy quantisation, allocation, the band loop with its PVQ search —
//     runs on lane 0, between team barriers.  Scalars cross the lane-0 / team boundary through `EncVars` in the stream's
//     scratch block; persistent state is written by lane 0 only.
#pragma once
#include "celt_enc_bands.cuh"
#include "celt_enc_energy.cuh"
#include "celt_mdct.cuh"
#include "opus_state.h"

namespace cb {

#if defined(CB_NO_BAND_PHASE)
enum { kEncPhases = 7 + 2 };
#else
enum { kEncPhases = 7 + 2 + kNbEBands };
#endif   // tm.phase() calls per coded frame (celt_encode_frame)
enum { kBitrateMax = -1, kOpusAuto = -1000, kFramesizeArg = 5000, kFramesizeVariable = 5010 };

// What the Opus layer sets on the CELT encoder before a frame (celt_encoder_ctl calls, src/opus_encoder.c:1715-1770)
struct CeltEncCfg {
    int C;              // stream channels (CELT_SET_CHANNELS)
    int end;            // end band (CELT_SET_END_BAND)
    int bitrate;        // OPUS_SET_BITRATE at the CELT level (kBitrateMax for CBR)
    int vbr, constrained_vbr;
    int complexity, lsb_depth, loss_rate, variable_duration;
    int disable_pf, force_intra;   // CELT_SET_PREDICTION
};

// Frame-level scalars shared between lane 0 and the team.
struct EncVars {
    EcEnc ec;
    int nbCompressedBytes, nbAvailableBytes, nbFilledBytes, vbr_rate, effectiveBytes, total_bits, equiv_rate;
    int silence, tell, enabled;
    int pf_on, pitch_index, gain1, qg, prefilter_tapset, prefilter_period0;
    int isTransient, shortBlocks, transient_got_disabled, tf_estimate, tf_chan, secondMdct, patch;
    int mask_metric[2], tr_scratch[4];
    int tf_select, tf_sum, do_tf, do_spread, do_trim, dual_stereo, alloc_trim;
    int temporal_vbr, maxDepth, tot_boost, total_boost;
    int anti_collapse_rsv, balance, codedBands;
    int ret;
};

// Per-stream working set of one frame, split by temperature.
//   EncShared — per-warp SHARED memory (~11 KB): everything lane 0 walks serially or the team hits repeatedly, overlaid by
//               phase (the phases of a frame are strictly sequential), plus the band-sized arrays and the frame scalars.
//   EncGlobal — HBM/L2: the three big sample buffers, touched only by coalesced team-wide passes.
struct EncShared {
    union Phase {
        int16_t pcm_buf[2 * (kMaxFrame + 192)];       // Opus layer: delay-compensation samples + DC-rejected input of this frame
        struct {                                       // pitch pre-filter analysis
            int16_t pitch_buf[(kCombMaxPeriod + kMaxFrame) / 2];
            int16_t x_lp4[kMaxFrame / 4], y_lp4[(kMaxFrame + kCombMaxPeriod) / 4];
            union {
                int16_t pitch_raw[(kCombMaxPeriod + kMaxFrame) / 2];
                struct { int xcorr[kCombMaxPeriod / 2]; int yy_lookup[kCombMaxPeriod / 2 + 1]; } c;
            } a;
        } pf;
        int tin[2 * (kMaxFrame + kOverlap)];          // transient_analysis: staged input (int32), rewritten in place as int16
        int fft[kMaxFrame];                            // MDCT: one channel's FFT buffer (all short blocks at once)
        uint8_t subpackets[3 * 1276];                  // 40/60 ms frames: the coded 20 ms sub-packets, staged for the repacketizer
        struct {                                       // normalised spectrum and what works on it
            int16_t X[2 * kMaxFrame];
            union {
                struct { int16_t tf_tmp[kMaxFrame], tf_tmp1[kMaxFrame]; } tf;
                struct { PvqScratch pvq; int16_t had_tmp[176]; } q;
            } w;
        } x;
    } u;
    int metric[kNbEBands];
    int bandE[2 * kNbEBands];
    int16_t bandLogE[2 * kNbEBands], bandLogE2[2 * kNbEBands], error[2 * kNbEBands];
    int16_t band_g[2 * kNbEBands];                    // normalise_bands gains
    int8_t band_shift[2 * kNbEBands];
    int tf_res[kNbEBands], offsets[kNbEBands], cap[kNbEBands], fine_quant[kNbEBands], pulses[kNbEBands], fine_priority[kNbEBands];
    AllocScratch alloc;
    CoarseScratch coarse;
    EncVars v;
};
struct EncGlobal {
    int in[2 * (kMaxFrame + kOverlap)];              // pre-emphasised input + overlap history, per channel
    int pre[2 * (kCombMaxPeriod + kMaxFrame)];       // pre-filter history + new samples, per channel
    int freq[2 * kMaxFrame];                          // MDCT output
};

// Encoder-side compute_allocation hooks (rate.c:346-364,391-411)
struct AllocEncIo {
    EcEnc &ec;
    int start, prev, signalBandwidth, LM;
    CB_MEM int skip_flag(int band_bits, int j, int codedBands, int bw) {
        if (codedBands <= start + 2 || (band_bits > ((j < prev ? 7 : 9) * bw << LM << kBitRes) >> 4 && j <= signalBandwidth)) {
            ec.bit_logp(1, 1);
            return 1;
        }
        ec.bit_logp(0, 1);
        return 0;
    }
    CB_MEM int intensity(int want, int start_, int codedBands) {
        want = imin(want, codedBands);
        ec.uint_((unsigned)(want - start_), (unsigned)(codedBands + 1 - start_));
        return want;
    }
    CB_MEM int dual_stereo(int want) {
        ec.bit_logp(want, 1);
        return want;
    }
};

// remove_doubling (pitch.c:372-505).  opus-fix keeps g, g0 32-bit (pitch.c:376,416-420).
template <class TM>
CB_DEV_NOINLINE int remove_doubling_team(TM tm, const int16_t *x, int maxperiod, int minperiod, int N, int *T0_, int prev_period, int prev_gain,
                                int *yy_lookup) {
    const int minperiod0 = minperiod;
    maxperiod /= 2; minperiod /= 2; *T0_ /= 2; prev_period /= 2; N /= 2;
    x += maxperiod;
    if (*T0_ >= maxperiod) *T0_ = maxperiod - 1;
    int T, T0;
    T = T0 = *T0_;
    int xx, xy;
    {
        int a = 0, b = 0;
        CB_TEAM_FOR(i, N, tm) { a = mac16_16(a, x[i], x[i]); b = mac16_16(b, x[i], x[i - T0]); }
        xx = tm.sum(a);
        xy = tm.sum(b);
    }
    // yy_lookup[i] = max(0, xx + sum_{k=1..i} (x[-k]^2 - x[N-k]^2)): a prefix sum (wrapping adds, order-free)
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
    int yy = yy_lookup[T0];
    int best_xy = xy, best_yy = yy;
    int g, g0;
    {
        int x2y2 = wadd(1, mul32_32_q31(xx, yy) >> 1);
        int sh = celt_ilog2(x2y2) >> 1;
        int t = vshr32(x2y2, 2 * (sh - 7));
        g = g0 = vshr32(mul16_32_q15(celt_rsqrt_norm(t), xy), sh + 1);
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
        xy = wadd(tm.sum(a), tm.sum(b));
        yy = wadd(yy_lookup[T1], yy_lookup[T1b]);
        int g1;
        {
            int x2y2 = wadd(1, mul32_32_q31(xx, yy));
            int sh = celt_ilog2(x2y2) >> 1;
            int t = vshr32(x2y2, 2 * (sh - 7));
            g1 = vshr32(mul16_32_q15(celt_rsqrt_norm(t), xy), sh + 1);
        }
        int cont;
        if (iabs(T1 - prev_period) <= 1) cont = prev_gain;
        else if (iabs(T1 - prev_period) <= 2 && 5 * k * k < T0) cont = s16(prev_gain >> 1);
        else cont = 0;
        int thresh = imax(9830, wsub(mul16_32_q15(22938, g0), cont));
        if (T1 < 3 * minperiod) thresh = imax(13107, wsub(mul16_32_q15(27853, g0), cont));
        else if (T1 < 2 * minperiod) thresh = imax(16384, wsub(mul16_32_q15(29491, g0), cont));
        if (g1 > thresh) {
            best_xy = xy; best_yy = yy; T = T1; g = g1;
        }
    }
    best_xy = imax(0, best_xy);
    int pg;
    if (best_yy <= best_xy) pg = 32767;
    else pg = s16(frac_div32(best_xy, wadd(best_yy, 1)) >> 16);
    int xc[3];
    {
        int a = 0, b = 0, c = 0;
        CB_TEAM_FOR(i, N, tm) {
            a = mac16_16(a, x[i], x[i - (T - 1)]);
            b = mac16_16(b, x[i], x[i - T]);
            c = mac16_16(c, x[i], x[i - (T + 1)]);
        }
        xc[0] = tm.sum(a); xc[1] = tm.sum(b); xc[2] = tm.sum(c);
    }
    int offset;
    if (wsub(xc[2], xc[0]) > mul16_32_q15(22938, wsub(xc[1], xc[0]))) offset = 1;
    else if (wsub(xc[0], xc[2]) > mul16_32_q15(22938, wsub(xc[1], xc[2]))) offset = -1;
    else offset = 0;
    if (pg > g) pg = s16(g);
    *T0_ = 2 * T + offset;
    if (*T0_ < minperiod0) *T0_ = minperiod0;
    tm.sync();
    return pg;
}

// comb_filter with y != x (celt.c:183-244), every output sample independent
template <class TM>
CB_DEV_NOINLINE void comb_filter_fir_team(TM tm, int *y, const int *x, int T0, int T1, int N, int g0, int g1, int tapset0, int tapset1, int overlap) {
    if (g0 == 0 && g1 == 0) {
        CB_TEAM_FOR(i, N, tm) y[i] = x[i];
        return;
    }
    const int g00 = s16(mul16_16_p15(g0, kCombGains[tapset0][0]));
    const int g01 = s16(mul16_16_p15(g0, kCombGains[tapset0][1]));
    const int g02 = s16(mul16_16_p15(g0, kCombGains[tapset0][2]));
    const int g10 = s16(mul16_16_p15(g1, kCombGains[tapset1][0]));
    const int g11 = s16(mul16_16_p15(g1, kCombGains[tapset1][1]));
    const int g12 = s16(mul16_16_p15(g1, kCombGains[tapset1][2]));
    if (g0 == g1 && T0 == T1 && tapset0 == tapset1) overlap = 0;
    CB_TEAM_FOR(i, N, tm) {
        int v = x[i];
        if (i < overlap) {
            const int f = s16(mul16_16_q15(kWindow120[i], kWindow120[i]));
            const int nf = 32767 - f;
            v = wadd(v, mul16_32_q15(mul16_16_q15(nf, g00), x[i - T0]));
            v = wadd(v, mul16_32_q15(mul16_16_q15(nf, g01), wadd(x[i - T0 + 1], x[i - T0 - 1])));
            v = wadd(v, mul16_32_q15(mul16_16_q15(nf, g02), wadd(x[i - T0 + 2], x[i - T0 - 2])));
            v = wadd(v, mul16_32_q15(mul16_16_q15(f, g10), x[i - T1]));
            v = wadd(v, mul16_32_q15(mul16_16_q15(f, g11), wadd(x[i - T1 + 1], x[i - T1 - 1])));
            v = wadd(v, mul16_32_q15(mul16_16_q15(f, g12), wadd(x[i - T1 + 2], x[i - T1 - 2])));
        } else if (g1 != 0) {
            v = wadd(v, mul16_32_q15(g10, x[i - T1]));
            v = wadd(v, mul16_32_q15(g11, wadd(x[i - T1 + 1], x[i - T1 - 1])));
            v = wadd(v, mul16_32_q15(g12, wadd(x[i - T1 + 2], x[i - T1 - 2])));
        }
        y[i] = v;
    }
}

// ---- transient analysis (celt_encoder.c:227-378): one channel, order dependent -> one lane per channel ------------------
CB_TABLE uint8_t kInvTable[128] = {
    255, 255, 156, 110, 86, 70, 59, 51, 45, 40, 37, 33, 31, 28, 26, 25, 23, 22, 21, 20, 19, 18, 17, 16, 16, 15, 15, 14, 13, 13, 12, 12,
    12, 12, 11, 11, 11, 10, 10, 10, 9, 9, 9, 9, 9, 9, 8, 8, 8, 8, 8, 7, 7, 7, 7, 7, 7, 6, 6, 6, 6, 6, 6, 6,
    6, 6, 6, 6, 6, 6, 6, 6, 6, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
    4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 2};

// transient_analysis (celt_encoder.c:227-378), split by what is order dependent:
//   A  lane per channel : the high-pass recurrence (floor() inside the feedback: not a linear scan) + max/min of its output
//   B  whole team       : normalisation shift and the pairwise energies x2[i] (independent), their sum `mean`
//   C  lane per channel : forward (post-echo) and backward (pre-echo) one-pole followers over x2, maxE
//   D  whole team       : the inverse-table sum over every fourth follower value
// tin: per channel `len` int32 (input >> SIG_SHIFT, staged in shared memory).  Stage A rewrites the first half of a channel's
// block in place as int16 (element i of the int16 view trails element i of the int32 view); x2 and the followers live in the
// second half.  sc: 4 ints of team-shared scratch.  mask_metric[c] receives the channel's metric.
template <class TM>
CB_DEV_NOINLINE void transient_analysis_team(TM tm, int *tin, int len, int CC, int *sc, int *mask_metric) {
    const int len2 = len / 2;
    CB_NOUNROLL for (int c = tm.lane(); c < CC; c += TM::W) {
        const int *xin = tin + c * len;
        int16_t *tmp = reinterpret_cast<int16_t *>(tin + c * len);
        int mem0 = 0, mem1 = 0, mxv = 0, mnv = 0;
        CB_NOUNROLL for (int i = 0; i < len; i++) {
            const int x = xin[i];
            const int y = wadd(mem0, x);
            mem0 = wsub(wadd(mem1, y), shl32(x, 1));
            mem1 = wsub(x, y >> 1);
            const int t = i < 12 ? 0 : s16(y >> 2);
            tmp[i] = (int16_t)t;
            mxv = imax(mxv, t);
            mnv = imin(mnv, t);
        }
        sc[c] = 14 - celt_ilog2(1 + imax(mxv, -mnv));
    }
    tm.sync();
    CB_NOUNROLL for (int c = 0; c < CC; c++) {
        const int16_t *tmp = reinterpret_cast<const int16_t *>(tin + c * len);
        int16_t *x2buf = reinterpret_cast<int16_t *>(tin + c * len) + len;
        // SHL16 with a NEGATIVE count (maxabs == 32768) is what the reference's C expression does on x86: a 32-bit shift by
        // (count & 31) of the zero-extended value, truncated to 16 bits
        const int shift = sc[c];
        int part = 0;
        CB_TEAM_FOR(i, len2, tm) {
            int a = tmp[2 * i], b = tmp[2 * i + 1];
            if (shift != 0) {
                a = (int16_t)((unsigned)(uint16_t)a << (shift & 31));
                b = (int16_t)((unsigned)(uint16_t)b << (shift & 31));
            }
            const int x2 = s16(pshr32(wadd(mul16_16(a, a), mul16_16(b, b)), 16));
            x2buf[i] = (int16_t)x2;
            part = wadd(part, x2);
        }
        const int mean = tm.sum(part);
        if (tm.lane() == 0) sc[2 + c] = mean;
    }
    tm.sync();
    CB_NOUNROLL for (int c = tm.lane(); c < CC; c += TM::W) {
        int16_t *f = reinterpret_cast<int16_t *>(tin + c * len) + len;
        int mem0 = 0;
        CB_NOUNROLL for (int i = 0; i < len2; i++) {
            mem0 = s16(mem0 + pshr32(f[i] - mem0, 4));
            f[i] = (int16_t)mem0;
        }
        mem0 = 0;
        int maxE = 0;
        CB_NOUNROLL for (int i = len2 - 1; i >= 0; i--) {
            mem0 = s16(mem0 + pshr32(f[i] - mem0, 3));
            f[i] = (int16_t)mem0;
            maxE = imax(maxE, mem0);
        }
        sc[c] = maxE;
    }
    tm.sync();
    CB_NOUNROLL for (int c = 0; c < CC; c++) {
        const int16_t *f = reinterpret_cast<const int16_t *>(tin + c * len) + len;
        const int mean = mul16_16(celt_sqrt(sc[2 + c]), celt_sqrt(mul16_16(sc[c], len2 >> 1)));
        const int norm = shl32(len2, 6 + 14) / wadd(1, mean >> 1);
        const int cnt = (len2 - 5 - 12 + 3) / 4;   // i = 12, 16, ... < len2 - 5
        int part = 0;
        CB_TEAM_FOR(k, cnt, tm) {
            const int id = imax(0, imin(127, mul16_32_q15(f[12 + 4 * k] + 1, norm)));
            part += kInvTable[id];
        }
        const int unmask = tm.sum(part);
        if (tm.lane() == 0) mask_metric[c] = 64 * unmask * 4 / (6 * (len2 - 17));
    }
    tm.sync();
}

// patch_transient_decision (celt_encoder.c:380-414)
CB_DEV int patch_transient_decision(const int16_t *newE, const int16_t *oldE, int start, int end, int C) {
    int mean_diff = 0;
    int spread_old[26];
    if (C == 1) {
        spread_old[start] = oldE[start];
        CB_NOUNROLL for (int i = start + 1; i < end; i++) spread_old[i] = s16(imax(spread_old[i - 1] - 1024, (int)oldE[i]));
    } else {
        spread_old[start] = imax((int)oldE[start], (int)oldE[start + kNbEBands]);
        CB_NOUNROLL for (int i = start + 1; i < end; i++)
            spread_old[i] = s16(imax(spread_old[i - 1] - 1024, imax((int)oldE[i], (int)oldE[i + kNbEBands])));
    }
    CB_NOUNROLL for (int i = end - 2; i >= start; i--) spread_old[i] = s16(imax(spread_old[i], spread_old[i + 1] - 1024));
    CB_NOUNROLL for (int c = 0; c < C; c++)
        CB_NOUNROLL for (int i = imax(2, start); i < end - 1; i++) {
            int x1 = imax(0, (int)newE[i + c * kNbEBands]);
            int x2 = imax(0, spread_old[i]);
            mean_diff = wadd(mean_diff, imax(0, x1 - x2));
        }
    mean_diff = mean_diff / (C * (end - 1 - imax(2, start)));
    return mean_diff > 1024;
}

// ---- tf_analysis (celt_encoder.c:539-712) -------------------------------------------------------------------------
CB_DEV int l1_metric(const int16_t *tmp, int N, int LM, int bias) {
    int L1 = 0;
    CB_NOUNROLL for (int i = 0; i < N; i++) L1 += iabs((int)tmp[i]);
    return mac16_32_q15(L1, LM * bias, L1);
}

// metric of one band (the per-band body of tf_analysis, :580-640); tmp / tmp_1: this band's N int16 of scratch
CB_DEV_NOINLINE int tf_band_metric(const int16_t *Xb, int N, int narrow, int isTransient, int LM, int bias, int16_t *tmp, int16_t *tmp_1,
                                   int *tf_sum_term) {
    CB_NOUNROLL for (int j = 0; j < N; j++) tmp[j] = Xb[j];
    int L1 = l1_metric(tmp, N, isTransient ? LM : 0, bias);
    int best_L1 = L1;
    int best_level = 0;
    if (isTransient && !narrow) {
        CB_NOUNROLL for (int j = 0; j < N; j++) tmp_1[j] = tmp[j];
        haar1(tmp_1, N >> LM, 1 << LM);
        L1 = l1_metric(tmp_1, N, LM + 1, bias);
        if (L1 < best_L1) { best_L1 = L1; best_level = -1; }
    }
    CB_NOUNROLL for (int k = 0; k < LM + !(isTransient || narrow); k++) {
        int B = isTransient ? LM - k - 1 : k + 1;
        haar1(tmp, N >> k, 1 << k);
        L1 = l1_metric(tmp, N, B, bias);
        if (L1 < best_L1) { best_L1 = L1; best_level = k + 1; }
    }
    int metric = isTransient ? 2 * best_level : -2 * best_level;
    *tf_sum_term = (isTransient ? LM : 0) - metric / 2;
    if (narrow && (metric == 0 || metric == -2 * LM)) metric -= 1;
    return metric;
}

// the Viterbi search over the band metrics (:641-710).  Returns tf_select.
CB_DEV_NOINLINE int tf_viterbi(const int *metric, int len, int isTransient, int *tf_res, int lambda, int LM) {
    int path0[kNbEBands], path1[kNbEBands];
    int selcost[2];
    int tf_select = 0;
    CB_NOUNROLL for (int sel = 0; sel < 2; sel++) {
        int cost0 = 0;
        int cost1 = isTransient ? 0 : lambda;
        CB_NOUNROLL for (int i = 1; i < len; i++) {
            int curr0 = imin(cost0, cost1 + lambda);
            int curr1 = imin(cost0 + lambda, cost1);
            cost0 = curr0 + iabs(metric[i] - 2 * kTfSelect[LM][4 * isTransient + 2 * sel + 0]);
            cost1 = curr1 + iabs(metric[i] - 2 * kTfSelect[LM][4 * isTransient + 2 * sel + 1]);
        }
        selcost[sel] = imin(cost0, cost1);
    }
    if (selcost[1] < selcost[0] && isTransient) tf_select = 1;
    int cost0 = 0;
    int cost1 = isTransient ? 0 : lambda;
    CB_NOUNROLL for (int i = 1; i < len; i++) {
        int curr0, curr1;
        int from0 = cost0, from1 = cost1 + lambda;
        if (from0 < from1) { curr0 = from0; path0[i] = 0; }
        else { curr0 = from1; path0[i] = 1; }
        from0 = cost0 + lambda;
        from1 = cost1;
        if (from0 < from1) { curr1 = from0; path1[i] = 0; }
        else { curr1 = from1; path1[i] = 1; }
        cost0 = curr0 + iabs(metric[i] - 2 * kTfSelect[LM][4 * isTransient + 2 * tf_select + 0]);
        cost1 = curr1 + iabs(metric[i] - 2 * kTfSelect[LM][4 * isTransient + 2 * tf_select + 1]);
    }
    tf_res[len - 1] = cost0 < cost1 ? 0 : 1;
    CB_NOUNROLL for (int i = len - 2; i >= 0; i--) tf_res[i] = tf_res[i + 1] == 1 ? path1[i + 1] : path0[i + 1];
    return tf_select;
}

// ---- dynalloc_analysis (celt_encoder.c:873-1064) ------------------------------------------------------------------
CB_DEV int median_of_5(const int16_t *x) {
    int t0, t1, t2 = x[2], t3, t4;
    if (x[0] > x[1]) { t0 = x[1]; t1 = x[0]; } else { t0 = x[0]; t1 = x[1]; }
    if (x[3] > x[4]) { t3 = x[4]; t4 = x[3]; } else { t3 = x[3]; t4 = x[4]; }
    if (t0 > t3) { int t = t0; t0 = t3; t3 = t; t = t1; t1 = t4; t4 = t; }
    if (t2 > t1) return t1 < t3 ? imin(t2, t3) : imin(t4, t1);
    return t2 < t3 ? imin(t1, t3) : imin(t2, t4);
}
CB_DEV int median_of_3(const int16_t *x) {
    int t0, t1, t2 = x[2];
    if (x[0] > x[1]) { t0 = x[1]; t1 = x[0]; } else { t0 = x[0]; t1 = x[1]; }
    if (t1 < t2) return t1;
    if (t0 < t2) return t2;
    return t0;
}

CB_DEV_NOINLINE int dynalloc_analysis(const int16_t *bandLogE, const int16_t *bandLogE2, int start, int end, int C, int *offsets,
                                      int lsb_depth, int isTransient, int vbr, int constrained_vbr, int LM, int effectiveBytes,
                                      int *tot_boost_) {
    const int nb = kNbEBands;
    int tot_boost = 0;
    int16_t follower[2 * kNbEBands], noise_floor[kNbEBands];
    CB_NOUNROLL for (int i = 0; i < nb; i++) offsets[i] = 0;
    int maxDepth = -32666;
    CB_NOUNROLL for (int i = 0; i < end; i++)
        noise_floor[i] = (int16_t)(mul16_16(64, kLogN[i]) + 512 + shl16(9 - lsb_depth, 10) - shl16(kEMeans[i], 6) + mul16_16(6, (i + 5) * (i + 5)));
    CB_NOUNROLL for (int c = 0; c < C; c++)
        CB_NOUNROLL for (int i = 0; i < end; i++) maxDepth = imax(maxDepth, (int)bandLogE[c * nb + i] - noise_floor[i]);
    maxDepth = s16(maxDepth);
    if (effectiveBytes > 50 && LM >= 1) {
        int last = 0;
        CB_NOUNROLL for (int c = 0; c < C; c++) {
            int16_t *f = &follower[c * nb];
            const int16_t *e2 = &bandLogE2[c * nb];
            f[0] = e2[0];
            CB_NOUNROLL for (int i = 1; i < end; i++) {
                if (e2[i] > e2[i - 1] + 512) last = i;
                f[i] = (int16_t)imin(f[i - 1] + 1536, (int)e2[i]);
            }
            CB_NOUNROLL for (int i = last - 1; i >= 0; i--) f[i] = (int16_t)imin((int)f[i], imin(f[i + 1] + 2048, (int)e2[i]));
            const int offset = 1024;
            CB_NOUNROLL for (int i = 2; i < end - 2; i++) f[i] = (int16_t)imax((int)f[i], median_of_5(&e2[i - 2]) - offset);
            int tmp = median_of_3(&e2[0]) - offset;
            f[0] = (int16_t)imax((int)f[0], tmp);
            f[1] = (int16_t)imax((int)f[1], tmp);
            tmp = median_of_3(&e2[end - 3]) - offset;
            f[end - 2] = (int16_t)imax((int)f[end - 2], tmp);
            f[end - 1] = (int16_t)imax((int)f[end - 1], tmp);
            CB_NOUNROLL for (int i = 0; i < end; i++) f[i] = (int16_t)imax((int)f[i], (int)noise_floor[i]);
        }
        if (C == 2) {
            CB_NOUNROLL for (int i = start; i < end; i++) {
                follower[nb + i] = (int16_t)imax((int)follower[nb + i], follower[i] - 4096);
                follower[i] = (int16_t)imax((int)follower[i], follower[nb + i] - 4096);
                follower[i] = (int16_t)((imax(0, bandLogE[i] - follower[i]) + imax(0, bandLogE[nb + i] - follower[nb + i])) >> 1);
            }
        } else {
            CB_NOUNROLL for (int i = start; i < end; i++) follower[i] = (int16_t)imax(0, bandLogE[i] - follower[i]);
        }
        // surround_dynalloc is all zero here: follower[i] = MAX16(follower[i], 0) is the identity for these non-negative values
        if ((!vbr || constrained_vbr) && !isTransient)
            CB_NOUNROLL for (int i = start; i < end; i++) follower[i] = (int16_t)(follower[i] >> 1);
        CB_NOUNROLL for (int i = start; i < end; i++) {
            if (i < 8) follower[i] = (int16_t)(follower[i] * 2);
            if (i >= 12) follower[i] = (int16_t)(follower[i] >> 1);
            follower[i] = (int16_t)imin((int)follower[i], 4096);
            const int width = C * band_width(i) << LM;
            int boost, boost_bits;
            if (width < 6) {
                boost = (int)follower[i] >> 10;
                boost_bits = boost * width << kBitRes;
            } else if (width > 48) {
                boost = ((int)follower[i] * 8) >> 10;
                boost_bits = (boost * width << kBitRes) / 8;
            } else {
                boost = ((int)follower[i] * width / 6) >> 10;
                boost_bits = boost * 6 << kBitRes;
            }
            if ((!vbr || (constrained_vbr && !isTransient)) && (tot_boost + boost_bits) >> kBitRes >> 3 > effectiveBytes / 4) {
                int cap = ((effectiveBytes / 4) << kBitRes << 3);
                offsets[i] = cap - tot_boost;
                tot_boost = cap;
                break;
            } else {
                offsets[i] = boost;
                tot_boost += boost_bits;
            }
        }
    }
    *tot_boost_ = tot_boost;
    return maxDepth;
}

// ---- compute_vbr (celt_encoder.c:1194-1312), no analysis, no surround mask, no LFE ---------------------------------
CB_DEV_NOINLINE int compute_vbr(int base_target, int LM, int bitrate, int lastCodedBands, int C, int intensity, int constrained_vbr,
                                int stereo_saving, int tot_boost, int tf_estimate, int maxDepth, int variable_duration, int temporal_vbr) {
    const int coded_bands = lastCodedBands ? lastCodedBands : kNbEBands;
    int coded_bins = kEBands[coded_bands] << LM;
    if (C == 2) coded_bins += kEBands[imin(intensity, coded_bands)] << LM;
    int target = base_target;
    if (C == 2) {
        const int coded_stereo_bands = imin(intensity, coded_bands);
        const int coded_stereo_dof = (kEBands[coded_stereo_bands] << LM) - coded_stereo_bands;
        const int max_frac = s16(mul16_16(26214, coded_stereo_dof) / s16(coded_bins));
        stereo_saving = imin(stereo_saving, 256);
        target -= imin(mul16_32_q15(max_frac, target), mul16_16(stereo_saving - 26, coded_stereo_dof << kBitRes) >> 8);
    }
    target += tot_boost - (16 << LM);
    const int tf_calibration = variable_duration == kFramesizeVariable ? 328 : 655;
    target += shl32(mul16_32_q15(tf_estimate - tf_calibration, target), 1);
    {
        const int bins = kEBands[kNbEBands - 2] << LM;
        int floor_depth = mul16_16(C * bins << kBitRes, maxDepth) >> 10;
        floor_depth = imax(floor_depth, target >> 2);
        target = imin(target, floor_depth);
    }
    if (constrained_vbr || bitrate < 64000) {
        int rate_factor = imax(0, bitrate - 32000);
        if (constrained_vbr) rate_factor = imin(rate_factor, 21955);
        target = base_target + mul16_32_q15(rate_factor, target - base_target);
    }
    if (tf_estimate < 3277) {
        const int amount = s16(mul16_16_q15(3329, imax(0, imin(32000, 96000 - bitrate))));
        const int tvbr_factor = s16(mul16_16(temporal_vbr, amount) >> 10);
        target += mul16_32_q15(tvbr_factor, target);
    }
    return imin(2 * base_target, target);
}

// ---- alloc_trim_analysis (celt_encoder.c:756-838) & stereo_analysis (:840-870), team reductions over X -------------
template <class TM>
CB_DEV_NOINLINE int alloc_trim_analysis_team(TM tm, const int16_t *X, const int16_t *bandLogE, int end, int LM, int C, int N0, int *stereo_saving,
                                    int tf_estimate, int intensity) {
    int diff = 0;
    int trim = 1280;
    if (C == 2) {
        int sum = 0;
        CB_NOUNROLL for (int i = 0; i < 8; i++) {
            int partial = team_inner16(tm, &X[kEBands[i] << LM], &X[N0 + (kEBands[i] << LM)], band_width(i) << LM);
            sum = s16(sum + s16(partial >> 18));
        }
        sum = mul16_16_q15(4096, sum);
        sum = imin(1024, iabs(sum));
        int minXC = sum;
        CB_NOUNROLL for (int i = 8; i < intensity; i++) {
            int partial = team_inner16(tm, &X[kEBands[i] << LM], &X[N0 + (kEBands[i] << LM)], band_width(i) << LM);
            minXC = imin(minXC, iabs(s16(partial >> 18)));
        }
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

template <class TM>
CB_DEV int stereo_analysis_team(TM tm, const int16_t *X, int LM, int N0) {
    int lr = 0, ms = 0;
    CB_TEAM_FOR(j, kEBands[13] << LM, tm) {
        int L = X[j], R = X[N0 + j];
        int M = L + R, S = L - R;
        lr = wadd(lr, iabs(L) + iabs(R));
        ms = wadd(ms, iabs(M) + iabs(S));
    }
    int sumLR = wadd(1, tm.sum(lr));
    int sumMS = wadd(1, tm.sum(ms));
    sumMS = mul16_32_q15(23170, sumMS);
    int thetas = 13;
    if (LM <= 1) thetas -= 8;
    return mul16_32_q15((kEBands[13] << (LM + 1)) + thetas, sumMS) > mul16_32_q15(kEBands[13] << (LM + 1), sumLR);
}

// spreading_decision (bands.c:428-519): threshold counts per band as one packed team sum
template <class TM>
CB_DEV_NOINLINE int spreading_decision_team(TM tm, const int16_t *X, int *average, int last_decision, int *hf_average, int *tapset_decision,
                                   int update_hf, int end, int C, int M) {
    int sum = 0, nbBands = 0, hf_sum = 0;
    const int N0 = M * kShortMdct;
    if (M * (kEBands[end] - kEBands[end - 1]) <= 8) return kSpreadNone;
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

// ---- transform + band energies (team) --------------------------------------------------------------------------------

// clt_mdct_forward (mdct.c:121-259) for the B blocks of one channel at once: window/fold and pre-rotation are fused and write
// straight into the shared FFT buffer in bit-reversed order, one batched FFT, post-rotation writes the interleaved output.
// in: the channel's B*N2 + overlap samples (HBM); out: freq of this channel, coefficient k of block b at out[b + k*B].
template <class TM>
CB_DEV void mdct_forward_blocks(TM tm, const int *in, int *out, int shift, int B, int *fftbuf) {
    const int N2 = (kMaxFrame * 2 >> shift) >> 1;
    const int N4 = N2 >> 1;
    int trig_off = 0;
    CB_NOUNROLL for (int i = 0, n = kMaxFrame * 2; i < shift; i++) { n >>= 1; trig_off += n; }
    const int16_t *t = kMdctTwiddles + trig_off;
    const int16_t *bitrev = fft_bitrev(shift);
    const int scale_shift = kFftPlan[shift].scale_shift - 1;
    const int ov = kOverlap, q = (ov + 3) >> 2;
    CB_TEAM_FOR(w, B * N4, tm) {
        const int b = w / N4, i = w - b * N4;
        const int *xp1 = in + b * N2 + (ov >> 1) + 2 * i;
        const int *xp2 = in + b * N2 + N2 - 1 + (ov >> 1) - 2 * i;
        int re, im;
        if (i < q) {
            const int w1 = kWindow120[(ov >> 1) + 2 * i], w2 = kWindow120[(ov >> 1) - 1 - 2 * i];
            re = wadd(smul(xp1[N2], w2), smul(*xp2, w1));
            im = wsub(smul(*xp1, w1), smul(xp2[-N2], w2));
        } else if (i < N4 - q) {
            re = *xp2;
            im = *xp1;
        } else {
            const int k = i - (N4 - q);
            const int w1 = kWindow120[2 * k], w2 = kWindow120[ov - 1 - 2 * k];
            re = wadd(wneg(smul(xp1[-N2], w1)), smul(*xp2, w2));
            im = wadd(smul(*xp1, w2), smul(xp2[N2], w1));
        }
        const int t0 = t[i], t1 = t[N4 + i];
        int yr = wsub(smul(re, t0), smul(im, t1));
        int yi = wadd(smul(im, t0), smul(re, t1));
        yr = pshr32(mul16_32_q16(kFftScale, yr), scale_shift);
        yi = pshr32(mul16_32_q16(kFftScale, yi), scale_shift);
        const int rev = bitrev[i];
        fftbuf[b * N2 + 2 * rev] = yr;
        fftbuf[b * N2 + 2 * rev + 1] = yi;
    }
    tm.sync();
    fft_inplace(tm, (Cpx *)fftbuf, shift, B);
    CB_TEAM_FOR(w, B * N4, tm) {
        const int b = w / N4, i = w - b * N4;
        const int fr = fftbuf[b * N2 + 2 * i], fi = fftbuf[b * N2 + 2 * i + 1];
        const int yr = wsub(smul(fi, t[N4 + i]), smul(fr, t[i]));
        const int yi = wadd(smul(fr, t[N4 + i]), smul(fi, t[i]));
        out[b + B * (2 * i)] = yr;
        out[b + B * (N2 - 1 - 2 * i)] = yi;
    }
    tm.sync();
}

// compute_mdcts (celt_encoder.c:418-461)
template <class TM>
CB_DEV_NOINLINE void compute_mdcts_team(TM tm, int shortBlocks, const int *in, int *freq, int C, int CC, int LM, int upsample, int *fftbuf) {
    int B, N, shift;
    if (shortBlocks) { B = shortBlocks; N = kShortMdct; shift = kMaxLM; }
    else { B = 1; N = kShortMdct << LM; shift = kMaxLM - LM; }
    CB_NOUNROLL for (int c = 0; c < CC; c++) mdct_forward_blocks(tm, in + c * (B * N + kOverlap), freq + c * N * B, shift, B, fftbuf);
    if (CC == 2 && C == 1) {
        CB_TEAM_FOR(i, B * N, tm) freq[i] = wadd(freq[i] >> 1, freq[B * N + i] >> 1);
        tm.sync();
    }
    if (upsample != 1) {   // API rate below 48 kHz: the zero-stuffed input's images above the original Nyquist are dropped
        const int bound = B * N / upsample;
        CB_TEAM_FOR(w, C * B * N, tm) {
            const int i = w % (B * N);
            freq[w] = i < bound ? wmul(freq[w], upsample) : 0;
        }
        tm.sync();
    }
}

// compute_band_energies (bands.c:97-143) + amp2Log2 (quant_bands.c:551-572)
template <class TM>
CB_DEV_NOINLINE void band_energies_team(TM tm, const int *freq, int *bandE, int16_t *bandLogE, int effEnd, int end, int C, int LM) {
    const int N = kShortMdct << LM;
    CB_NOUNROLL for (int c = 0; c < C; c++) {
        CB_NOUNROLL for (int i = 0; i < effEnd; i++) {
            const int lo = kEBands[i] << LM, hi = kEBands[i + 1] << LM;
            const int *x = freq + c * N;
            const int maxval = team_maxabs32(tm, x + lo, hi - lo);
            int e = 1;
            if (maxval > 0) {
                const int shift = celt_ilog2(maxval) - 14 + (((kLogN[i] >> kBitRes) + LM + 1) >> 1);
                int sum = 0;
                if (shift > 0) {
                    CB_TEAM_FOR(j, hi - lo, tm) { int v = s16(x[lo + j] >> shift); sum = mac16_16(sum, v, v); }
                } else {
                    CB_TEAM_FOR(j, hi - lo, tm) { int v = s16(shl32(x[lo + j], -shift)); sum = mac16_16(sum, v, v); }
                }
                sum = tm.sum(sum);
                e = wadd(1, vshr32(celt_sqrt(sum), -shift));
            }
            if (tm.lane() == 0) {
                bandE[i + c * kNbEBands] = e;
                bandLogE[i + c * kNbEBands] = (int16_t)(celt_log2(shl32(e, 2)) - shl16(kEMeans[i], 6));
            }
        }
        if (tm.lane() == 0)
            CB_NOUNROLL for (int i = effEnd; i < end; i++) bandLogE[c * kNbEBands + i] = -14336;
    }
    tm.sync();
}

// normalise_bands (bands.c:146-164)
template <class TM>
CB_DEV_NOINLINE void normalise_bands_team(TM tm, const int *freq, int16_t *X, const int *bandE, int end, int C, int M, int LM, int16_t *band_g,
                                 int8_t *band_shift) {
    const int N = M * kShortMdct;
    CB_TEAM_FOR(k, C * kNbEBands, tm) {
        const int c = k / kNbEBands, i = k - c * kNbEBands;
        if (i < end) {
            int shift = celt_zlog2(bandE[i + c * kNbEBands]) - 13;
            int E = s16(vshr32(bandE[i + c * kNbEBands], shift));
            band_g[k] = (int16_t)celt_rcp(shl32(E, 3));
            band_shift[k] = (int8_t)shift;
        }
    }
    tm.sync();
    const int top = M * kEBands[end];
    CB_TEAM_FOR(w, C * top, tm) {
        const int c = w / top, j = w - c * top;
        const int k = c * kNbEBands + kBinToBand[j >> LM];
        X[j + c * N] = (int16_t)mul16_16_q15(s16(vshr32(freq[j + c * N], band_shift[k] - 1)), band_g[k]);
    }
    tm.sync();
}

// ---- the frame -------------------------------------------------------------------------------------------------------

// celt_encode_with_ec (celt_encoder.c:1379-2273).  `pcm`: CC-interleaved int16, frame_size_api samples per channel at the API rate
// (st->upsample = 48000 / Fs: below 48 kHz the input is zero-stuffed in the pre-emphasis, :490-533).
// `st` holds the head of the state (CB_ENC_HEAD_BYTES, possibly a shared-memory copy), `gst` the full block in HBM (only its
// sample histories are touched).  `pcm` may live in S.u.pcm_buf: it is dead before the first overlay is written.
// S.v.ec must hold the range coder the Opus layer initialised (and shrank to nbCompressedBytes).  Returns (on every lane)
// the number of payload bytes, or a negative error.
template <class TM>
CB_DEV int celt_encode_frame(TM tm, CbEncState *st, CbEncState *gst, EncShared &S, EncGlobal &G, const CeltEncCfg &cfg, const int16_t *pcm,
                             int frame_size_api, int nbCompressedBytes_in) {
    EncVars &V = S.v;
    const int upsample = st->upsample;
    const int frame_size = frame_size_api * upsample;
    const bool L0 = tm.lane() == 0;
    const int CC = st->channels;
    const int C = cfg.C;
    const int start = 0;
    const int end = cfg.end;
    const int effEnd = end;   // effEBands == 21 in the 48 kHz mode
    int LM;
    CB_NOUNROLL for (LM = 0; LM <= kMaxLM; LM++)
        if (kShortMdct << LM == frame_size) break;
    if (LM > kMaxLM || nbCompressedBytes_in < 2) {
        for (int i = 0; i < kEncPhases; i++) tm.phase();
        return OPUS_BAD_ARG_;
    }
    const int M = 1 << LM;
    const int N = M * kShortMdct;
    const int ov = kOverlap;

    // ---- header: rate bookkeeping (:1480-1560) ----
    if (L0) {
        EcEnc ec = V.ec;
        int nbCompressedBytes = nbCompressedBytes_in;
        const int tell = ec.tell();
        const int nbFilledBytes = (tell + 4) >> 3;
        nbCompressedBytes = imin(nbCompressedBytes, 1275);
        int nbAvailableBytes = nbCompressedBytes - nbFilledBytes;
        int vbr_rate, effectiveBytes;
        if (cfg.vbr && cfg.bitrate != kBitrateMax) {
            const int den = 48000 >> kBitRes;
            vbr_rate = (cfg.bitrate * frame_size + (den >> 1)) / den;
            effectiveBytes = vbr_rate >> (3 + kBitRes);
        } else {
            vbr_rate = 0;
            int tmp = wmul(cfg.bitrate, frame_size);
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
        V.ec = ec;
        V.nbCompressedBytes = nbCompressedBytes; V.nbAvailableBytes = nbAvailableBytes; V.nbFilledBytes = nbFilledBytes;
        V.vbr_rate = vbr_rate; V.effectiveBytes = effectiveBytes; V.equiv_rate = equiv_rate;
        V.total_bits = nbCompressedBytes * 8;
        V.tell = tell;
    }
    // ---- silence detection (:1567-1571) and pre-emphasis (:1598-1606) ----
    int sample_max;
    {
        const int old_overlap_max = st->overlap_max;
        const int a = team_maxabs16(tm, pcm, C * (N - ov) / upsample);
        const int b = team_maxabs16(tm, pcm + C * (N - ov) / upsample, C * ov / upsample);
        sample_max = imax(imax(old_overlap_max, a), b);
        tm.sync();
        if (L0) st->overlap_max = b;
    }
    CB_NOUNROLL for (int c = 0; c < CC; c++) {
        int *inp = G.in + c * (N + ov) + ov;
        const int m0 = st->preemph_memE[c];
        if (upsample == 1) {
            CB_TEAM_FOR(i, N, tm) {
                const int x = pcm[CC * i + c];
                const int m = i == 0 ? m0 : mul16_16(kPreemphCoef0, pcm[CC * (i - 1) + c]) >> 3;
                inp[i] = wsub(shl32(x, 12), m);
            }
        } else {
            CB_TEAM_FOR(i, N, tm) {   // x[k] = pcm[k / upsample] when k is a multiple of upsample, else 0
                const int x = i % upsample == 0 ? (int)pcm[CC * (i / upsample) + c] : 0;
                const int xp = i == 0 ? 0 : ((i - 1) % upsample == 0 ? (int)pcm[CC * ((i - 1) / upsample) + c] : 0);
                const int m = i == 0 ? m0 : mul16_16(kPreemphCoef0, xp) >> 3;
                inp[i] = wsub(shl32(x, 12), m);
            }
        }
        tm.sync();
        if (L0) st->preemph_memE[c] = upsample == 1 ? mul16_16(kPreemphCoef0, pcm[CC * (N - 1) + c]) >> 3 : 0;
    }
    if (L0) {
        EcEnc ec = V.ec;
        int silence = sample_max == 0;
        int tell = V.tell;
        if (tell == 1) ec.bit_logp(silence, 15);
        else silence = 0;
        if (silence) {
            if (V.vbr_rate > 0) {
                V.effectiveBytes = V.nbCompressedBytes = imin(V.nbCompressedBytes, V.nbFilledBytes + 2);
                V.total_bits = V.nbCompressedBytes * 8;
                V.nbAvailableBytes = 2;
                ec.shrink((unsigned)V.nbCompressedBytes);
            }
            tell = V.nbCompressedBytes * 8;
            ec.nbits_total += tell - ec.tell();
        }
        V.silence = silence;
        V.tell = tell;
        V.ec = ec;
        V.enabled = V.nbAvailableBytes > 12 * C && start == 0 && !silence && !cfg.disable_pf && cfg.complexity >= 5 &&
                    !(st->consec_transient && LM != 3 && cfg.variable_duration == kFramesizeVariable);
        V.prefilter_tapset = st->tapset_decision;
        V.prefilter_period0 = imax(st->prefilter_period, kCombMinPeriod);
    }
    tm.sync();

    tm.phase();
    // ---- pitch pre-filter (run_prefilter, :1067-1192) ----
    {
        int *pre0 = G.pre, *pre1 = G.pre + (N + kCombMaxPeriod);
        CB_NOUNROLL for (int c = 0; c < CC; c++) {
            int *pre = c ? pre1 : pre0;
            CB_TEAM_FOR(i, kCombMaxPeriod, tm) pre[i] = gst->prefilter_mem[c * kCombMaxPeriod + i];
            CB_TEAM_FOR(i, N, tm) pre[kCombMaxPeriod + i] = G.in[c * (N + ov) + ov + i];
        }
        tm.sync();
        int pitch_index, gain1;
        const int prev_period = st->prefilter_period, prev_gain = st->prefilter_gain, prev_tapset = st->prefilter_tapset;
        if (V.enabled) {
            pitch_downsample_team(tm, pre0, pre1, kCombMaxPeriod + N, CC, S.u.pf.a.pitch_raw, S.u.pf.pitch_buf);
            tm.phase();
            pitch_index = pitch_search_team(tm, S.u.pf.pitch_buf + (kCombMaxPeriod >> 1), S.u.pf.pitch_buf, N, kCombMaxPeriod - 3 * kCombMinPeriod,
                                            S.u.pf.x_lp4, S.u.pf.y_lp4, S.u.pf.a.c.xcorr, S.u.pf.a.c.yy_lookup);
            pitch_index = kCombMaxPeriod - pitch_index;
            tm.phase();
            gain1 = remove_doubling_team(tm, S.u.pf.pitch_buf, kCombMaxPeriod, kCombMinPeriod, N, &pitch_index, prev_period, prev_gain, S.u.pf.a.c.yy_lookup);
            if (pitch_index > kCombMaxPeriod - 2) pitch_index = kCombMaxPeriod - 2;
            gain1 = s16(mul16_16_q15(22938, gain1));
            if (cfg.loss_rate > 2) gain1 = gain1 >> 1;
            if (cfg.loss_rate > 4) gain1 = gain1 >> 1;
            if (cfg.loss_rate > 8) gain1 = 0;
        } else {
            tm.phase();
            tm.phase();
            gain1 = 0;
            pitch_index = kCombMinPeriod;
        }
        int pf_threshold = 6554;
        if (iabs(pitch_index - prev_period) * 10 > pitch_index) pf_threshold += 6554;
        if (V.nbAvailableBytes < 25) pf_threshold += 3277;
        if (V.nbAvailableBytes < 35) pf_threshold += 3277;
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
        const int pp = V.prefilter_period0;
        const int tapset1 = V.prefilter_tapset;
        CB_NOUNROLL for (int c = 0; c < CC; c++) {
            int *pre = c ? pre1 : pre0;
            int *inc = G.in + c * (N + ov);
            CB_TEAM_FOR(i, ov, tm) inc[i] = gst->in_mem[c * ov + i];
            comb_filter_fir_team(tm, inc + ov, pre + kCombMaxPeriod, pp, pitch_index, N, -prev_gain, -gain1, prev_tapset, tapset1, ov);
            tm.sync();
            CB_TEAM_FOR(i, ov, tm) gst->in_mem[c * ov + i] = inc[N + i];
            CB_TEAM_FOR(i, kCombMaxPeriod, tm) gst->prefilter_mem[c * kCombMaxPeriod + i] = pre[N + i];
        }
        tm.sync();
        if (L0) {
            V.pf_on = pf_on; V.pitch_index = pitch_index; V.gain1 = gain1; V.qg = qg;
            EcEnc ec = V.ec;
            if (pf_on == 0) {
                if (start == 0 && V.tell + 16 <= V.total_bits) ec.bit_logp(0, 1);
            } else {
                ec.bit_logp(1, 1);
                const int pi1 = pitch_index + 1;
                const int octave = ec_ilog((unsigned)pi1) - 5;
                ec.uint_((unsigned)octave, 6);
                ec.bits((unsigned)(pi1 - (16 << octave)), (unsigned)(4 + octave));
                ec.bits((unsigned)qg, 3);
                ec.icdf(tapset1, kTapsetIcdf, 2);
            }
            V.ec = ec;
        }
    }

    tm.phase();
    // ---- transient analysis (:1642-1657): one lane per channel ----
    if (cfg.complexity >= 1) {
        CB_TEAM_FOR(i, CC * (N + ov), tm) S.u.tin[i] = G.in[i] >> 12;
        tm.sync();
        transient_analysis_team(tm, S.u.tin, N + ov, CC, V.tr_scratch, V.mask_metric);
    }
    tm.sync();
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

    tm.phase();
    // ---- MDCT, band energies (:1660-1690) ----
    if (V.secondMdct) {
        compute_mdcts_team(tm, 0, G.in, G.freq, C, CC, LM, upsample, S.u.fft);
        band_energies_team(tm, G.freq, S.bandE, S.bandLogE2, effEnd, end, C, LM);
        CB_TEAM_FOR(i, C * kNbEBands, tm) S.bandLogE2[i] = (int16_t)(S.bandLogE2[i] + (shl16(LM, 10) >> 1));
        tm.sync();
    }
    compute_mdcts_team(tm, V.shortBlocks, G.in, G.freq, C, CC, LM, upsample, S.u.fft);
    band_energies_team(tm, G.freq, S.bandE, S.bandLogE, effEnd, end, C, LM);

    // ---- temporal VBR, bandLogE2, transient patch (:1803-1848) ----
    if (L0) {
        if (CC == 2 && C == 1) V.tf_chan = 0;
        {
            int follow = -10240;
            int frame_avg = 0;
            const int offset = V.shortBlocks ? (shl16(LM, 10) >> 1) : 0;
            CB_NOUNROLL for (int i = start; i < end; i++) {
                follow = s16(imax(follow - 1024, S.bandLogE[i] - offset));
                if (C == 2) follow = s16(imax(follow, S.bandLogE[i + kNbEBands] - offset));
                frame_avg += follow;
            }
            frame_avg /= (end - start);
            int temporal_vbr = s16(s16(frame_avg) - s16(st->spec_avg));
            temporal_vbr = imin(3072, imax(-1536, temporal_vbr));
            st->spec_avg = s16(st->spec_avg + mul16_16_q15(655, temporal_vbr));
            V.temporal_vbr = temporal_vbr;
        }
        if (!V.secondMdct)
            CB_NOUNROLL for (int i = 0; i < C * kNbEBands; i++) S.bandLogE2[i] = S.bandLogE[i];
        V.patch = 0;
        if (LM > 0 && V.ec.tell() + 3 <= V.total_bits && !V.isTransient && cfg.complexity >= 5) {
            if (patch_transient_decision(S.bandLogE, st->oldBandE, start, end, C)) {
                V.patch = 1;
                V.isTransient = 1;
                V.shortBlocks = M;
            }
        }
    }
    tm.sync();
    if (V.patch) {
        compute_mdcts_team(tm, V.shortBlocks, G.in, G.freq, C, CC, LM, upsample, S.u.fft);
        band_energies_team(tm, G.freq, S.bandE, S.bandLogE, effEnd, end, C, LM);
        CB_TEAM_FOR(i, C * kNbEBands, tm) S.bandLogE2[i] = (int16_t)(S.bandLogE2[i] + (shl16(LM, 10) >> 1));
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
    tm.phase();
    // ---- band normalisation (:1856) ----
    normalise_bands_team(tm, G.freq, S.u.x.X, S.bandE, effEnd, C, M, LM, S.band_g, S.band_shift);

    // ---- tf_analysis (:1858-1880): one band per lane ----
    const int isTransient = V.isTransient;
    const int shortBlocks = V.shortBlocks;
    if (V.do_tf) {
        const int bias = mul16_16_q14(1311, imax(-4096, 8192 - V.tf_estimate));
        const int16_t *Xc = S.u.x.X + V.tf_chan * N;
        int tf_sum = 0;
        CB_TEAM_FOR(i, effEnd, tm) {
            const int lo = kEBands[i] << LM;
            const int Nb = band_width(i) << LM;
            int term;
            S.metric[i] = tf_band_metric(Xc + lo, Nb, band_width(i) == 1, isTransient, LM, bias, S.u.x.w.tf.tf_tmp + lo, S.u.x.w.tf.tf_tmp1 + lo, &term);
            tf_sum += term;
        }
        tf_sum = tm.sum(tf_sum);
        tm.sync();
        if (L0) {
            int lambda;
            if (V.effectiveBytes < 40) lambda = 12;
            else if (V.effectiveBytes < 60) lambda = 6;
            else if (V.effectiveBytes < 100) lambda = 4;
            else lambda = 3;
            lambda *= 2;
            V.tf_select = tf_viterbi(S.metric, effEnd, isTransient, S.tf_res, lambda, LM);
            CB_NOUNROLL for (int i = effEnd; i < end; i++) S.tf_res[i] = S.tf_res[effEnd - 1];
            V.tf_sum = tf_sum;
        }
    } else if (L0) {
        V.tf_sum = 0;
        CB_NOUNROLL for (int i = 0; i < end; i++) S.tf_res[i] = isTransient;
        V.tf_select = 0;
    }
    tm.phase();
    // ---- coarse energy, tf flags (:1882-1889) ----
    if (L0) {
        EcEnc ec = V.ec;
        quant_coarse_energy(start, end, effEnd, S.bandLogE, st->oldBandE, (unsigned)V.total_bits, S.error, ec, C, LM, V.nbAvailableBytes,
                            cfg.force_intra, &st->delayedIntra, cfg.complexity >= 4, cfg.loss_rate, S.coarse);
        tf_encode(start, end, isTransient, S.tf_res, LM, V.tf_select, ec);
        V.do_spread = 0;
        if (ec.tell() + 4 <= V.total_bits) {
            if (shortBlocks || cfg.complexity < 3 || V.nbAvailableBytes < 10 * C || start != 0) {
                st->spread_decision = cfg.complexity == 0 ? kSpreadNone : kSpreadNormal;
                ec.icdf(st->spread_decision, kSpreadIcdf, 5);
            } else {
                V.do_spread = 1;   // needs the team: decided below, coded right after
            }
        }
        V.ec = ec;
    }
    tm.sync();
    if (V.do_spread) {
        int average = st->tonal_average, hf_average = st->hf_average, tapset_decision = st->tapset_decision;
        const int last = st->spread_decision;
        const int dec = spreading_decision_team(tm, S.u.x.X, &average, last, &hf_average, &tapset_decision, V.pf_on && !shortBlocks, effEnd, C, M);
        tm.sync();
        if (L0) {
            st->tonal_average = average; st->hf_average = hf_average; st->tapset_decision = tapset_decision;
            st->spread_decision = dec;
            EcEnc ec = V.ec;
            ec.icdf(dec, kSpreadIcdf, 5);
            V.ec = ec;
        }
    }
    // ---- dynalloc (:1927-1972), stereo decisions (:1974-1990) ----
    if (L0) {
        EcEnc ec = V.ec;
        V.maxDepth = dynalloc_analysis(S.bandLogE, S.bandLogE2, start, end, C, S.offsets, cfg.lsb_depth, isTransient, cfg.vbr,
                                       cfg.constrained_vbr, LM, V.effectiveBytes, &V.tot_boost);
        init_caps(S.cap, LM, C);
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
            CB_NOUNROLL for (j = 0; tell + (loop_logp << kBitRes) < total_bits - total_boost && boost < S.cap[i]; j++) {
                const int flag = j < S.offsets[i];
                ec.bit_logp(flag, (unsigned)loop_logp);
                tell = (int)ec.tell_frac();
                if (!flag) break;
                boost += quanta;
                total_boost += quanta;
                loop_logp = 1;
            }
            if (j) dynalloc_logp = imax(2, dynalloc_logp - 1);
            S.offsets[i] = boost;
        }
        V.total_bits = total_bits;       // now in 1/8 bits
        V.tell = tell;
        V.total_boost = total_boost;
        V.ec = ec;
        V.do_trim = tell + (6 << kBitRes) <= total_bits - total_boost;
    }
    tm.sync();
    int dual_stereo = 0;
    if (C == 2) {
        if (LM != 0) dual_stereo = stereo_analysis_team(tm, S.u.x.X, LM, N);
        if (L0) {
            st->intensity = hysteresis_decision(s16(V.equiv_rate / 1000), kIntensityThresholds, kIntensityHisteresis, 21, st->intensity);
            st->intensity = imin(end, imax(start, st->intensity));
        }
        tm.sync();
    }
    int alloc_trim = 5;
    if (V.do_trim) {
        int stereo_saving = st->stereo_saving;
        alloc_trim = alloc_trim_analysis_team(tm, S.u.x.X, S.bandLogE, end, LM, C, N, &stereo_saving, V.tf_estimate, st->intensity);
        tm.sync();
        if (L0) st->stereo_saving = stereo_saving;
    }

    tm.phase();
    // ---- rate control, allocation, quantisation, packing: lane 0 (:1992-2262) ----
    if (L0) {
        EcEnc ec = V.ec;
        int nbCompressedBytes = V.nbCompressedBytes;
        int nbAvailableBytes = V.nbAvailableBytes;
        const int nbFilledBytes = V.nbFilledBytes;
        const int total_bits = V.total_bits;
        const int total_boost = V.total_boost;
        int tell = V.tell;
        if (V.do_trim) {
            ec.icdf(alloc_trim, kTrimIcdf, 7);
            tell = (int)ec.tell_frac();
        }
        const int vbr_rate = V.vbr_rate;
        if (vbr_rate > 0) {
            const int lm_diff = kMaxLM - LM;
            nbCompressedBytes = imin(nbCompressedBytes, 1275 >> (3 - LM));
            int base_target = vbr_rate - ((40 * C + 20) << kBitRes);
            if (cfg.constrained_vbr) base_target += (st->vbr_offset >> lm_diff);
            int target = compute_vbr(base_target, LM, V.equiv_rate, st->lastCodedBands, C, st->intensity, cfg.constrained_vbr,
                                     st->stereo_saving, V.tot_boost, V.tf_estimate, V.maxDepth, cfg.variable_duration, V.temporal_vbr);
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
        int bits = ((nbCompressedBytes * 8) << kBitRes) - (int)ec.tell_frac() - 1;
        const int anti_collapse_rsv = isTransient && LM >= 2 && bits >= ((LM + 2) << kBitRes) ? (1 << kBitRes) : 0;
        bits -= anti_collapse_rsv;
        const int signalBandwidth = end - 1;
        int balance = 0;
        int intensity = st->intensity;
        AllocEncIo io{ec, start, st->lastCodedBands, signalBandwidth, LM};
        const int codedBands = compute_allocation(io, S.alloc, start, end, S.offsets, S.cap, alloc_trim, &intensity, &dual_stereo, bits,
                                                  &balance, S.pulses, S.fine_quant, S.fine_priority, C, LM);
        st->intensity = intensity;
        if (st->lastCodedBands) st->lastCodedBands = imin(st->lastCodedBands + 1, imax(st->lastCodedBands - 1, codedBands));
        else st->lastCodedBands = codedBands;
        quant_fine_energy(start, end, st->oldBandE, S.error, S.fine_quant, ec, C);
        V.ec = ec;
        V.nbCompressedBytes = nbCompressedBytes;
        V.anti_collapse_rsv = anti_collapse_rsv;
        V.balance = balance;
        V.codedBands = codedBands;
        V.dual_stereo = dual_stereo;
    }
    tm.sync();
    tm.phase();
    // ---- residual quantisation (:2208): every lane walks the band loop with identical scalars, vector work is split ----
    {
        EcEnc ec = V.ec;
        quant_all_bands_enc(tm, start, end, S.u.x.X, C == 2 ? S.u.x.X + N : nullptr, S.bandE, S.pulses, shortBlocks, st->spread_decision,
                            V.dual_stereo, st->intensity, S.tf_res, V.nbCompressedBytes * (8 << kBitRes) - V.anti_collapse_rsv, V.balance, ec, LM,
                            V.codedBands, &S.u.x.w.q.pvq, S.u.x.w.q.had_tmp);
        tm.sync();
        if (L0) V.ec = ec;
    }
    if (L0) {
        EcEnc ec = V.ec;
        const int nbCompressedBytes = V.nbCompressedBytes;
        const int anti_collapse_rsv = V.anti_collapse_rsv;
        if (anti_collapse_rsv > 0) {
            const int anti_collapse_on = st->consec_transient < 2;
            ec.bits((unsigned)anti_collapse_on, 1);
        }
        quant_energy_finalise(start, end, st->oldBandE, S.error, S.fine_quant, S.fine_priority, nbCompressedBytes * 8 - ec.tell(), ec, C);
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
            CB_NOUNROLL for (int i = 0; i < start; i++) {
                st->oldBandE[c * kNbEBands + i] = 0;
                st->oldLogE[c * kNbEBands + i] = st->oldLogE2[c * kNbEBands + i] = -28672;
            }
            CB_NOUNROLL for (int i = end; i < kNbEBands; i++) {
                st->oldBandE[c * kNbEBands + i] = 0;
                st->oldLogE[c * kNbEBands + i] = st->oldLogE2[c * kNbEBands + i] = -28672;
            }
        }
        if (isTransient || V.transient_got_disabled) st->consec_transient++;
        else st->consec_transient = 0;
        st->rng = ec.rng;
        ec.done();
        V.ec = ec;
        V.ret = ec.error ? OPUS_INTERNAL_ERROR_ : nbCompressedBytes;
    }
    tm.sync();
    return V.ret;
}

}  // namespace cb
