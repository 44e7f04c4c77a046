// celt_decoder.cuh — one CELT frame: symbols -> allocation -> PVQ -> IMDCT -> post-filter -> PCM,
// plus the Opus packet layer above it (TOC, framing codes 0-3, CELT-only dispatch).
//
// Restates opus-fix/celt/celt_decoder.c:185-275 (deemphasis), :280-350 (celt_synthesis), :713-1072
// (celt_decode_with_ec), celt/celt.c:156-244 (comb_filter), celt/bands.c:169-238 (denormalise_bands,
// fused into the IMDCT pre-rotation here), src/opus_decoder.c:200-596 (opus_decode_frame, CELT-only
// subset) and :598-709 (opus_decode_native), src/opus.c:169-343 (packet parsing).
//
// One team (= one warp on the GPU) owns one stream.  Frame-level structure:
//   lane-0 section : header flags, coarse energy, tf, spread, dynalloc, trim, bit allocation, fine energy
//                    (scalar, range-coder bound; the coder state is then broadcast to the team)
//   team section   : band loop (uniform scalar control + lane-strided vectors), anti-collapse,
//                    fused denormalise+IMDCT, comb filter in runs of T-2, history shift
//   lane-per-channel: de-emphasis (1-pole IIR, order dependent)
#pragma once
#include "celt_bands.cuh"
#include "celt_energy.cuh"
#include "celt_mdct.cuh"
#include "opus_state.h"

namespace cb {

enum {
    OPUS_OK_ = 0, OPUS_BAD_ARG_ = -1, OPUS_BUFFER_TOO_SMALL_ = -2, OPUS_INTERNAL_ERROR_ = -3,
    OPUS_INVALID_PACKET_ = -4, OPUS_UNIMPLEMENTED_ = -5, OPUS_INVALID_STATE_ = -6, OPUS_ALLOC_FAIL_ = -7,
};
enum { kBwNarrow = 1101, kBwMedium = 1102, kBwWide = 1103, kBwSuperWide = 1104, kBwFull = 1105 };

// bin (at LM=0 resolution) -> band, from eband5ms
CB_TABLE uint8_t kBinToBand[100] = {
    0, 1, 2, 3, 4, 5, 6, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 12, 12, 13, 13, 13, 13, 14, 14, 14, 14,
    15, 15, 15, 15, 15, 15, 16, 16, 16, 16, 16, 16, 17, 17, 17, 17, 17, 17, 17, 17,
    18, 18, 18, 18, 18, 18, 18, 18, 18, 18, 18, 18,
    19, 19, 19, 19, 19, 19, 19, 19, 19, 19, 19, 19, 19, 19, 19, 19, 19, 19,
    20, 20, 20, 20, 20, 20, 20, 20, 20, 20, 20, 20, 20, 20, 20, 20, 20, 20, 20, 20, 20, 20};

// Team-shared working set of one frame decode (~9.5 KB).
struct DecScratch {
    int16_t X[2 * kMaxFrame];                         // normalised spectrum, channel-major
    union {
        int16_t norm[2 * 8 * 78];                     // folding source (band loop only)
        int fft[kMaxFrame];                           // IMDCT buffer (synthesis only)
    } u;
    int16_t tmp[176];                                 // pulse vector / hadamard staging
    int tf_res[kNbEBands], cap[kNbEBands], offsets[kNbEBands], fine_quant[kNbEBands], pulses[kNbEBands],
        fine_priority[kNbEBands];
    AllocScratch alloc;
    int16_t den_gain[2][kNbEBands];                   // denormalisation gain / shift per band
    int8_t den_shift[2][kNbEBands];
    uint8_t collapse[2 * kNbEBands];
    int hdr[16];                                      // lane-0 -> team hand-off of frame scalars
};

#if defined(__CUDACC__)
CB_DEV void ec_broadcast(EcDec &d, int src) {
    d.storage = team_bcast(d.storage, src); d.end_offs = team_bcast(d.end_offs, src);
    d.end_window = team_bcast(d.end_window, src); d.nend_bits = team_bcast(d.nend_bits, src);
    d.nbits_total = team_bcast(d.nbits_total, src); d.offs = team_bcast(d.offs, src);
    d.rng = team_bcast(d.rng, src); d.val = team_bcast(d.val, src); d.ext = team_bcast(d.ext, src);
    d.rem = team_bcast(d.rem, src); d.error = team_bcast(d.error, src);
}
#else
CB_DEV void ec_broadcast(EcDec &, int) {}
#endif

// comb_filter with y == x (celt/celt.c:183-244, as called at celt_decoder.c:1002-1013).  In place the
// filter is recursive: output i reads positions <= i-T+2, already final.  Outputs within a run of
// min(T0,T1)-2 (>= 13) samples are mutually independent, so runs are spread over the lanes.
CB_DEV void comb_filter_inplace(Team tm, int *x, int T0, int T1, int N, int g0, int g1, int tapset0, int tapset1, int overlap) {
    if (g0 == 0 && g1 == 0) return;
    const int g00 = s16(mul16_16_p15(g0, kCombGains[tapset0][0]));
    const int g01 = s16(mul16_16_p15(g0, kCombGains[tapset0][1]));
    const int g02 = s16(mul16_16_p15(g0, kCombGains[tapset0][2]));
    const int g10 = s16(mul16_16_p15(g1, kCombGains[tapset1][0]));
    const int g11 = s16(mul16_16_p15(g1, kCombGains[tapset1][1]));
    const int g12 = s16(mul16_16_p15(g1, kCombGains[tapset1][2]));
    if (g0 == g1 && T0 == T1 && tapset0 == tapset1) overlap = 0;
    // A tap set whose gain is zero contributes exactly 0 (MULT16_32_Q15(0,.) == 0) whatever its period — the
    // "off" period is 0 in the bitstream — so it is skipped and does not bound the run length.
    const int Tmin = imin(g0 != 0 ? T0 : kCombMaxPeriod, g1 != 0 ? T1 : kCombMaxPeriod);
    const int G = imin(CB_LANES, Tmin - 2);
    for (int base = 0; base < overlap; base += G) {
        int i = base + tm.lane;
        if (tm.lane < G && i < overlap) {
            int f = s16(mul16_16_q15(kWindow120[i], kWindow120[i]));
            int nf = 32767 - f;
            int v = x[i];
            if (g0 != 0) {
                v = wadd(v, mul16_32_q15(mul16_16_q15(nf, g00), x[i - T0]));
                v = wadd(v, mul16_32_q15(mul16_16_q15(nf, g01), wadd(x[i - T0 + 1], x[i - T0 - 1])));
                v = wadd(v, mul16_32_q15(mul16_16_q15(nf, g02), wadd(x[i - T0 + 2], x[i - T0 - 2])));
            }
            if (g1 != 0) {
                v = wadd(v, mul16_32_q15(mul16_16_q15(f, g10), x[i - T1]));
                v = wadd(v, mul16_32_q15(mul16_16_q15(f, g11), wadd(x[i - T1 + 1], x[i - T1 - 1])));
                v = wadd(v, mul16_32_q15(mul16_16_q15(f, g12), wadd(x[i - T1 + 2], x[i - T1 - 2])));
            }
            x[i] = v;
        }
        CB_SYNC();
    }
    if (g1 == 0) return;
    for (int base = overlap; base < N; base += G) {
        int i = base + tm.lane;
        if (tm.lane < G && i < N) {
            int v = x[i];
            v = wadd(v, mul16_32_q15(g10, x[i - T1]));
            v = wadd(v, mul16_32_q15(g11, wadd(x[i - T1 + 1], x[i - T1 - 1])));
            v = wadd(v, mul16_32_q15(g12, wadd(x[i - T1 + 2], x[i - T1 - 2])));
            x[i] = v;
        }
        CB_SYNC();
    }
}

// deemphasis (celt_decoder.c:185-275, accum = 0): lane c filters channel c.
CB_DEV void deemphasis(Team tm, const int *const *in, int16_t *pcm, int N, int CC, int downsample, int *mem) {
    if (tm.lane < CC || CB_LANES == 1) {
        const int c0 = CB_LANES == 1 ? 0 : tm.lane;
        const int c1 = CB_LANES == 1 ? CC : tm.lane + 1;
        for (int c = c0; c < c1; c++) {
            const int *x = in[c];
            int16_t *y = pcm + c;
            int m = mem[c];
            if (downsample > 1) {
                int k = 0;
                for (int j = 0; j < N; j++) {
                    int t = wadd(x[j], m);
                    m = mul16_32_q15(kPreemphCoef0, t);
                    if (j == k * downsample) { y[k * CC] = (int16_t)sig2word16(t); k++; }
                }
            } else {
                for (int j = 0; j < N; j++) {
                    int t = wadd(x[j], m);
                    m = mul16_32_q15(kPreemphCoef0, t);
                    y[j * CC] = (int16_t)sig2word16(t);
                }
            }
            mem[c] = m;
        }
    }
    CB_SYNC();
}

// Shift the synthesis history down by N samples (celt_decoder.c:962-964), team-parallel and overlap-safe:
// each chunk is read by all lanes before any lane writes it, and destinations trail sources by N >= 120.
CB_DEV void history_shift(Team tm, int *mem, int N) {
    const int count = kDecBuf - N + kOverlap / 2;
    for (int base = 0; base < count; base += CB_LANES) {
        int i = base + tm.lane;
        int v = 0;
        if (i < count) v = mem[i + N];
        CB_SYNC();
        if (i < count) mem[i] = v;
    }
    CB_SYNC();
}

// celt_decode_with_ec for a received frame (len >= 2).  `dec` must be initialised on the payload.
// pcm: interleaved int16, CC channels.  Returns samples per channel or a negative Opus error.
CB_DEV int celt_decode_frame(Team tm, CbDecState *st, DecScratch &S, const uint8_t *data, int len, int16_t *pcm,
                             int frame_size, int C, int start, int end, EcDec &dec) {
    (void)data;
    const int CC = st->channels;
    frame_size *= st->downsample;
    int LM;
    for (LM = 0; LM <= kMaxLM; LM++)
        if (kShortMdct << LM == frame_size) break;
    if (LM > kMaxLM) return OPUS_BAD_ARG_;
    const int M = 1 << LM;
    if (len < 0 || len > 1275 || pcm == nullptr) return OPUS_BAD_ARG_;
    const int N = M * kShortMdct;
    int *decode_mem[2], *out_syn[2];
    for (int c = 0; c < CC; c++) {
        decode_mem[c] = st->decode_mem + c * CB_DEC_MEM;
        out_syn[c] = decode_mem[c] + kDecBuf - N;
    }
    const int effEnd = imin(end, kNbEBands);
    int16_t *oldBandE = st->oldEBands, *oldLogE = st->oldLogE, *oldLogE2 = st->oldLogE2, *backgroundLogE = st->backgroundLogE;

    // ---------------- lane-0 section: everything up to the band loop ----------------
    enum { H_SILENCE, H_PF_PITCH, H_PF_GAIN, H_PF_TAPSET, H_TRANSIENT, H_SPREAD, H_INTENSITY, H_DUAL, H_BALANCE,
           H_CODED, H_ACRSV, H_TOTALBITS };
    if (tm.lane == 0) {
        if (C == 1)
            for (int i = 0; i < kNbEBands; i++) oldBandE[i] = (int16_t)imax(oldBandE[i], oldBandE[kNbEBands + i]);
        int total_bits = len * 8;
        int tell = dec.tell();
        int silence;
        if (tell >= total_bits) silence = 1;
        else if (tell == 1) silence = dec.bit_logp(15);
        else silence = 0;
        if (silence) {
            tell = len * 8;
            dec.nbits_total += tell - dec.tell();
        }
        int postfilter_gain = 0, postfilter_pitch = 0, postfilter_tapset = 0;
        if (start == 0 && tell + 16 <= total_bits) {
            if (dec.bit_logp(1)) {
                int octave = (int)dec.uint_(6);
                postfilter_pitch = (16 << octave) + (int)dec.bits(4 + octave) - 1;
                int qg = (int)dec.bits(3);
                if (dec.tell() + 2 <= total_bits) postfilter_tapset = dec.icdf(kTapsetIcdf, 2);
                postfilter_gain = 3072 * (qg + 1);   // QCONST16(.09375f,15)
            }
            tell = dec.tell();
        }
        int isTransient = 0;
        if (LM > 0 && tell + 3 <= total_bits) {
            isTransient = dec.bit_logp(3);
            tell = dec.tell();
        }
        int intra_ener = tell + 3 <= total_bits ? dec.bit_logp(3) : 0;
        unquant_coarse_energy(start, end, oldBandE, intra_ener, dec, C, LM);
        tf_decode(start, end, isTransient, S.tf_res, LM, dec);
        tell = dec.tell();
        int spread_decision = kSpreadNormal;
        if (tell + 4 <= total_bits) spread_decision = dec.icdf(kSpreadIcdf, 5);
        init_caps(S.cap, LM, C);
        int dynalloc_logp = 6;
        total_bits <<= kBitRes;
        tell = (int)dec.tell_frac();
        for (int i = start; i < end; i++) {
            int width = C * band_width(i) << LM;
            int quanta = imin(width << kBitRes, imax(6 << kBitRes, width));
            int loop_logp = dynalloc_logp;
            int boost = 0;
            while (tell + (loop_logp << kBitRes) < total_bits && boost < S.cap[i]) {
                int flag = dec.bit_logp(loop_logp);
                tell = (int)dec.tell_frac();
                if (!flag) break;
                boost += quanta;
                total_bits -= quanta;
                loop_logp = 1;
            }
            S.offsets[i] = boost;
            if (boost > 0) dynalloc_logp = imax(2, dynalloc_logp - 1);
        }
        int alloc_trim = tell + (6 << kBitRes) <= total_bits ? dec.icdf(kTrimIcdf, 7) : 5;
        int bits = ((len * 8) << kBitRes) - (int)dec.tell_frac() - 1;
        int anti_collapse_rsv = isTransient && LM >= 2 && bits >= ((LM + 2) << kBitRes) ? (1 << kBitRes) : 0;
        bits -= anti_collapse_rsv;
        int intensity = 0, dual_stereo = 0, balance = 0;
        AllocDecIo io{dec};
        int codedBands = compute_allocation(io, S.alloc, start, end, S.offsets, S.cap, alloc_trim, &intensity, &dual_stereo,
                                            bits, &balance, S.pulses, S.fine_quant, S.fine_priority, C, LM);
        unquant_fine_energy(start, end, oldBandE, S.fine_quant, dec, C);
        S.hdr[H_SILENCE] = silence; S.hdr[H_PF_PITCH] = postfilter_pitch; S.hdr[H_PF_GAIN] = postfilter_gain;
        S.hdr[H_PF_TAPSET] = postfilter_tapset; S.hdr[H_TRANSIENT] = isTransient; S.hdr[H_SPREAD] = spread_decision;
        S.hdr[H_INTENSITY] = intensity; S.hdr[H_DUAL] = dual_stereo; S.hdr[H_BALANCE] = balance;
        S.hdr[H_CODED] = codedBands; S.hdr[H_ACRSV] = anti_collapse_rsv;
    }
    CB_SYNC();
    ec_broadcast(dec, 0);
    const int silence = S.hdr[H_SILENCE], postfilter_pitch = S.hdr[H_PF_PITCH], postfilter_gain = S.hdr[H_PF_GAIN],
              postfilter_tapset = S.hdr[H_PF_TAPSET], isTransient = S.hdr[H_TRANSIENT], spread_decision = S.hdr[H_SPREAD],
              intensity = S.hdr[H_INTENSITY], dual_stereo = S.hdr[H_DUAL], balance = S.hdr[H_BALANCE],
              codedBands = S.hdr[H_CODED], anti_collapse_rsv = S.hdr[H_ACRSV];
    const int shortBlocks = isTransient ? M : 0;
#ifdef CB_HOSTSIM_TRACE
    fprintf(stderr, "[trace] LM=%d C=%d sil=%d pf=(%d,%d,%d) trans=%d spread=%d int=%d dual=%d coded=%d acrsv=%d\n", LM, C, silence,
            postfilter_pitch, postfilter_gain, postfilter_tapset, isTransient, spread_decision, intensity, dual_stereo, codedBands, anti_collapse_rsv);
#endif

    // ---------------- team section ----------------
    for (int c = 0; c < CC; c++) history_shift(tm, decode_mem[c], N);

    unsigned seed = st->rng;
    quant_all_bands_dec(tm, start, end, S.X, C == 2 ? S.X + N : nullptr, S.collapse, S.pulses, shortBlocks, spread_decision,
                        dual_stereo, intensity, S.tf_res, len * (8 << kBitRes) - anti_collapse_rsv, balance, dec, LM,
                        codedBands, &seed, S.u.norm, S.tmp);
    CB_SYNC();
    int anti_collapse_on = 0;
    if (anti_collapse_rsv > 0) anti_collapse_on = (int)dec.bits(1);
    if (tm.lane == 0)
        unquant_energy_finalise(start, end, oldBandE, S.fine_quant, S.fine_priority, len * 8 - dec.tell(), dec, C);
    CB_SYNC();
    ec_broadcast(dec, 0);
    if (anti_collapse_on)
        anti_collapse(tm, S.X, S.collapse, LM, C, N, start, end, oldBandE, oldLogE, oldLogE2, S.pulses, seed);
    if (silence) {
        CB_TEAM_FOR(i, C * kNbEBands, tm) oldBandE[i] = -28672;   // -QCONST16(28.f,DB_SHIFT)
        CB_SYNC();
    }

    // ---- celt_synthesis (celt_decoder.c:280-350) with denormalise_bands (bands.c:169-238) fused in ----
    {
        int bstart = start, bend = effEnd;
        int bound = M * kEBands[bend];
        if (st->downsample != 1) bound = imin(bound, N / st->downsample);
        if (silence) { bound = 0; bstart = bend = 0; }
        const int lo = M * kEBands[bstart];
        // per-band gain/shift (bands.c:195-227)
        CB_TEAM_FOR(w, C * kNbEBands, tm) {
            int c = w / kNbEBands, i = w - c * kNbEBands;
            int lg = s16(oldBandE[c * kNbEBands + i] + shl16(kEMeans[i], 6));
            int shift = 16 - (lg >> 10);
            int g;
            if (shift > 31) { shift = 0; g = 0; }
            else g = celt_exp2_frac(lg & 1023);
            if (shift < 0 && shift < -2) { g = 32767; shift = -2; }
            S.den_gain[c][i] = (int16_t)g;
            S.den_shift[c][i] = (int8_t)shift;
        }
        CB_SYNC();
        const int B = isTransient ? M : 1;
        const int shift = isTransient ? kMaxLM : kMaxLM - LM;
        const int16_t *X = S.X;
        const int hi = (bend > bstart) ? imin(bound, M * kEBands[bend]) : 0;
        auto freq_ch = [&](int c, int j) -> int {
            if (j < lo || j >= hi) return 0;
            int b = kBinToBand[j >> LM];
            int v = mul16_16(X[c * N + j], S.den_gain[c][b]);
            int sh = S.den_shift[c][b];
            return sh < 0 ? shl32(v, -sh) : (v >> sh);
        };
        if (CC == 2 && C == 1) {
            imdct_compute(tm, [&](int j) { return freq_ch(0, j); }, B, shift, S.u.fft);
            imdct_assemble(tm, out_syn[0], B, shift, S.u.fft);
            imdct_assemble(tm, out_syn[1], B, shift, S.u.fft);
        } else if (CC == 1 && C == 2) {
            imdct_compute(tm, [&](int j) { return wadd(freq_ch(0, j), freq_ch(1, j)) >> 1; }, B, shift, S.u.fft);
            imdct_assemble(tm, out_syn[0], B, shift, S.u.fft);
        } else {
            for (int c = 0; c < CC; c++) {
                imdct_compute(tm, [&](int j) { return freq_ch(c, j); }, B, shift, S.u.fft);
                imdct_assemble(tm, out_syn[c], B, shift, S.u.fft);
            }
        }
    }

    // ---- post-filter (celt_decoder.c:1001-1025) ----
    int pf_period = imax(st->postfilter_period, kCombMinPeriod);
    int pf_period_old = imax(st->postfilter_period_old, kCombMinPeriod);
    for (int c = 0; c < CC; c++) {
        comb_filter_inplace(tm, out_syn[c], pf_period_old, pf_period, kShortMdct, st->postfilter_gain_old, st->postfilter_gain,
                            st->postfilter_tapset_old, st->postfilter_tapset, kOverlap);
        if (LM != 0)
            comb_filter_inplace(tm, out_syn[c] + kShortMdct, pf_period, postfilter_pitch, N - kShortMdct, st->postfilter_gain,
                                postfilter_gain, st->postfilter_tapset, postfilter_tapset, kOverlap);
    }
    CB_SYNC();
    if (tm.lane == 0) {
        st->postfilter_period = pf_period;
        st->postfilter_period_old = pf_period_old;
        st->postfilter_period_old = st->postfilter_period;
        st->postfilter_gain_old = st->postfilter_gain;
        st->postfilter_tapset_old = st->postfilter_tapset;
        st->postfilter_period = postfilter_pitch;
        st->postfilter_gain = postfilter_gain;
        st->postfilter_tapset = postfilter_tapset;
        if (LM != 0) {
            st->postfilter_period_old = st->postfilter_period;
            st->postfilter_gain_old = st->postfilter_gain;
            st->postfilter_tapset_old = st->postfilter_tapset;
        }
        // ---- energy history (celt_decoder.c:1027-1062) ----
        if (C == 1)
            for (int i = 0; i < kNbEBands; i++) oldBandE[kNbEBands + i] = oldBandE[i];
        if (!isTransient) {
            int max_inc = st->loss_count < 10 ? M : 1024;   // M*QCONST16(0.001f,DB_SHIFT) = M*1 ; QCONST16(1.f,DB_SHIFT)
            for (int i = 0; i < 2 * kNbEBands; i++) {
                oldLogE2[i] = oldLogE[i];
                oldLogE[i] = oldBandE[i];
                backgroundLogE[i] = (int16_t)imin(backgroundLogE[i] + max_inc, (int)oldBandE[i]);
            }
        } else {
            for (int i = 0; i < 2 * kNbEBands; i++) oldLogE[i] = (int16_t)imin(oldLogE[i], oldBandE[i]);
        }
        for (int c = 0; c < 2; c++) {
            for (int i = 0; i < start; i++) {
                oldBandE[c * kNbEBands + i] = 0;
                oldLogE[c * kNbEBands + i] = oldLogE2[c * kNbEBands + i] = -28672;
            }
            for (int i = end; i < kNbEBands; i++) {
                oldBandE[c * kNbEBands + i] = 0;
                oldLogE[c * kNbEBands + i] = oldLogE2[c * kNbEBands + i] = -28672;
            }
        }
        st->rng = dec.rng;
        st->loss_count = 0;
        if (dec.error) st->error = 1;
    }
    CB_SYNC();
    deemphasis(tm, out_syn, pcm, N, CC, st->downsample, st->preemph_memD);
    if (dec.tell() > 8 * len) return OPUS_INTERNAL_ERROR_;
    return frame_size / st->downsample;
}

}  // namespace cb
