// celt_pitch.cuh — pitch analysis, LPC and the out-of-place comb filter (encoder pre-filter, PLC).
//
// Restates opus-fix/celt/pitch.c:45-505 (find_best_pitch, celt_fir5, pitch_downsample, celt_pitch_xcorr, pitch_search,
// remove_doubling), celt/pitch.h:66-169 (xcorr_kernel / inner products: plain wrapping 32-bit sums of 16x16 products, so the
// reference's 4-way unrolling is an implementation detail, not a numerical one), celt/celt_lpc.c:37-328 (_celt_lpc, celt_fir,
// celt_iir, _celt_autocorr) and celt/celt.c:156-244 (comb_filter with y != x).  Scalar code: one thread per stream.
#pragma once
#include "celt_arith.cuh"
#include "celt_tables.cuh"

namespace cb {

CB_DEV int inner_prod16(const int16_t *x, const int16_t *y, int n) {
    int s = 0;
    CB_NOUNROLL for (int i = 0; i < n; i++) s = mac16_16(s, x[i], y[i]);
    return s;
}
CB_DEV int maxabs16(const int16_t *x, int n) {
    int mx = 0, mn = 0;
    CB_NOUNROLL for (int i = 0; i < n; i++) { mx = imax(mx, x[i]); mn = imin(mn, x[i]); }
    return imax(mx, -mn);
}
CB_DEV int maxabs32(const int *x, int n) {
    int mx = 0, mn = 0;
    CB_NOUNROLL for (int i = 0; i < n; i++) { mx = imax(mx, x[i]); mn = imin(mn, x[i]); }
    return imax(mx, wneg(mn));
}

// celt_pitch_xcorr (pitch.c:225-258): xcorr[i] = sum_j x[j]*y[i+j]; returns max(1, max_i xcorr[i])
CB_DEV_NOINLINE int pitch_xcorr(const int16_t *x, const int16_t *y, int *xcorr, int len, int max_pitch) {
    int maxcorr = 1;
    CB_NOUNROLL for (int i = 0; i < max_pitch; i++) {
        int s = inner_prod16(x, y + i, len);
        xcorr[i] = s;
        maxcorr = imax(maxcorr, s);
    }
    return maxcorr;
}

// _celt_lpc (celt_lpc.c:37-93), p <= 24
CB_DEV_NOINLINE void celt_lpc(int16_t *out, const int *ac, int p) {
    int lpc[kLpcOrder];
    int error = ac[0];
    CB_NOUNROLL for (int i = 0; i < p; i++) lpc[i] = 0;
    if (ac[0] != 0) {
        CB_NOUNROLL for (int i = 0; i < p; i++) {
            int rr = 0;
            CB_NOUNROLL for (int j = 0; j < i; j++) rr = wadd(rr, mul32_32_q31(lpc[j], ac[i - j]));
            rr = wadd(rr, ac[i + 1] >> 3);
            int r = wneg(frac_div32(shl32(rr, 3), error));
            lpc[i] = r >> 3;
            CB_NOUNROLL for (int j = 0; j < (i + 1) >> 1; j++) {
                int t1 = lpc[j], t2 = lpc[i - 1 - j];
                lpc[j] = wadd(t1, mul32_32_q31(r, t2));
                lpc[i - 1 - j] = wadd(t2, mul32_32_q31(r, t1));
            }
            error = wsub(error, mul32_32_q31(mul32_32_q31(r, r), error));
            if (error < (ac[0] >> 10)) break;
        }
    }
    CB_NOUNROLL for (int i = 0; i < p; i++) out[i] = (int16_t)round16(lpc[i], 16);
}

// _celt_autocorr (celt_lpc.c:232-328).  xx: scratch for n int16.  Returns the shift.
CB_DEV_NOINLINE int celt_autocorr(const int16_t *x, int *ac, const int16_t *window, int overlap, int lag, int n, int16_t *xx) {
    const int fastN = n - lag;
    const int16_t *xptr;
    if (overlap == 0) {
        xptr = x;
    } else {
        CB_NOUNROLL for (int i = 0; i < n; i++) xx[i] = x[i];
        CB_NOUNROLL for (int i = 0; i < overlap; i++) {
            xx[i] = (int16_t)mul16_16_q15(x[i], window[i]);
            xx[n - i - 1] = (int16_t)mul16_16_q15(x[n - i - 1], window[i]);
        }
        xptr = xx;
    }
    int shift;
    {
        int ac0 = 1 + (n << 7);
        if (n & 1) ac0 = wadd(ac0, mul16_16(xptr[0], xptr[0]) >> 9);
        CB_NOUNROLL for (int i = (n & 1); i < n; i += 2) {
            ac0 = wadd(ac0, mul16_16(xptr[i], xptr[i]) >> 9);
            ac0 = wadd(ac0, mul16_16(xptr[i + 1], xptr[i + 1]) >> 9);
        }
        shift = celt_ilog2(ac0) - 30 + 10;
        shift = shift / 2;
        if (shift > 0) {
            CB_NOUNROLL for (int i = 0; i < n; i++) xx[i] = (int16_t)pshr32(xptr[i], shift);
            xptr = xx;
        } else {
            shift = 0;
        }
    }
    pitch_xcorr(xptr, xptr, ac, fastN, lag + 1);
    CB_NOUNROLL for (int k = 0; k <= lag; k++) {
        int d = 0;
        CB_NOUNROLL for (int i = k + fastN; i < n; i++) d = mac16_16(d, xptr[i], xptr[i - k]);
        ac[k] = wadd(ac[k], d);
    }
    shift = 2 * shift;
    if (shift <= 0) ac[0] = wadd(ac[0], shl32(1, -shift));
    if (ac[0] < 268435456) {
        int shift2 = 29 - ec_ilog((unsigned)ac[0]);
        CB_NOUNROLL for (int i = 0; i <= lag; i++) ac[i] = shl32(ac[i], shift2);
        shift -= shift2;
    } else if (ac[0] >= 536870912) {
        int shift2 = 1;
        if (ac[0] >= 1073741824) shift2++;
        CB_NOUNROLL for (int i = 0; i <= lag; i++) ac[i] = ac[i] >> shift2;
        shift += shift2;
    }
    return shift;
}

// pitch_downsample (pitch.c:147-217) incl. celt_fir5 (pitch.c:105-144).  x[c]: len samples; x_lp: len/2; xx: scratch len/2.
CB_DEV_NOINLINE void pitch_downsample(const int *x0, const int *x1, int16_t *x_lp, int len, int C, int16_t *xx) {
    int maxabs = maxabs32(x0, len);
    if (C == 2) maxabs = imax(maxabs, maxabs32(x1, len));
    if (maxabs < 1) maxabs = 1;
    int shift = celt_ilog2(maxabs) - 10;
    if (shift < 0) shift = 0;
    if (C == 2) shift++;
    const int half = len >> 1;
    CB_NOUNROLL for (int i = 1; i < half; i++) x_lp[i] = (int16_t)((wadd(wadd(x0[2 * i - 1], x0[2 * i + 1]) >> 1, x0[2 * i]) >> 1) >> shift);
    x_lp[0] = (int16_t)((wadd(x0[1] >> 1, x0[0]) >> 1) >> shift);
    if (C == 2) {
        CB_NOUNROLL for (int i = 1; i < half; i++)
            x_lp[i] = (int16_t)(x_lp[i] + (int16_t)((wadd(wadd(x1[2 * i - 1], x1[2 * i + 1]) >> 1, x1[2 * i]) >> 1) >> shift));
        x_lp[0] = (int16_t)(x_lp[0] + (int16_t)((wadd(x1[1] >> 1, x1[0]) >> 1) >> shift));
    }
    int ac[5];
    celt_autocorr(x_lp, ac, nullptr, 0, 4, half, xx);
    ac[0] = wadd(ac[0], ac[0] >> 13);
    CB_NOUNROLL for (int i = 1; i <= 4; i++) ac[i] = wsub(ac[i], mul16_32_q15(2 * i * i, ac[i]));
    int16_t lpc[4];
    celt_lpc(lpc, ac, 4);
    int tmp = 32767;
    CB_NOUNROLL for (int i = 0; i < 4; i++) {
        tmp = s16(mul16_16_q15(29491, tmp));   // QCONST16(.9f,15)
        lpc[i] = (int16_t)mul16_16_q15(lpc[i], tmp);
    }
    const int c1 = 26214;   // QCONST16(.8f,15)
    int16_t lpc2[5];
    lpc2[0] = (int16_t)(lpc[0] + 3277);   // QCONST16(.8f,SIG_SHIFT)
    lpc2[1] = (int16_t)(lpc[1] + mul16_16_q15(c1, lpc[0]));
    lpc2[2] = (int16_t)(lpc[2] + mul16_16_q15(c1, lpc[1]));
    lpc2[3] = (int16_t)(lpc[3] + mul16_16_q15(c1, lpc[2]));
    lpc2[4] = (int16_t)mul16_16_q15(c1, lpc[3]);
    // celt_fir5, in place (mem = previous INPUT samples)
    int m0 = 0, m1 = 0, m2 = 0, m3 = 0, m4 = 0;
    CB_NOUNROLL for (int i = 0; i < half; i++) {
        const int xi = x_lp[i];
        int sum = shl32(xi, 12);
        sum = mac16_16(sum, lpc2[0], m0);
        sum = mac16_16(sum, lpc2[1], m1);
        sum = mac16_16(sum, lpc2[2], m2);
        sum = mac16_16(sum, lpc2[3], m3);
        sum = mac16_16(sum, lpc2[4], m4);
        m4 = m3; m3 = m2; m2 = m1; m1 = m0; m0 = xi;
        x_lp[i] = (int16_t)round16(sum, 12);
    }
}

// find_best_pitch (pitch.c:45-103)
CB_DEV_NOINLINE void find_best_pitch(const int *xcorr, const int16_t *y, int len, int max_pitch, int *best_pitch, int yshift, int maxcorr) {
    int Syy = 1;
    int best_num[2] = {-1, -1};
    int best_den[2] = {0, 0};
    const int xshift = celt_ilog2(maxcorr) - 14;
    best_pitch[0] = 0;
    best_pitch[1] = 1;
    CB_NOUNROLL for (int j = 0; j < len; j++) Syy = wadd(Syy, mul16_16(y[j], y[j]) >> yshift);
    CB_NOUNROLL for (int i = 0; i < max_pitch; i++) {
        if (xcorr[i] > 0) {
            int xcorr16 = s16(vshr32(xcorr[i], xshift));
            int num = s16(mul16_16_q15(xcorr16, xcorr16));
            if (mul16_32_q15(num, best_den[1]) > mul16_32_q15(best_num[1], Syy)) {
                if (mul16_32_q15(num, best_den[0]) > mul16_32_q15(best_num[0], Syy)) {
                    best_num[1] = best_num[0]; best_den[1] = best_den[0]; best_pitch[1] = best_pitch[0];
                    best_num[0] = num; best_den[0] = Syy; best_pitch[0] = i;
                } else {
                    best_num[1] = num; best_den[1] = Syy; best_pitch[1] = i;
                }
            }
        }
        Syy = wadd(Syy, wsub(mul16_16(y[i + len], y[i + len]) >> yshift, mul16_16(y[i], y[i]) >> yshift));
        Syy = imax(1, Syy);
    }
}

// pitch_search (pitch.c:260-369).  scratch: x_lp4 (len/4), y_lp4 ((len+max_pitch)/4) int16, xcorr (max_pitch/2) int.
CB_DEV_NOINLINE void pitch_search(const int16_t *x_lp, const int16_t *y, int len, int max_pitch, int *pitch, int16_t *x_lp4,
                                  int16_t *y_lp4, int *xcorr) {
    const int lag = len + max_pitch;
    int best_pitch[2] = {0, 0};
    CB_NOUNROLL for (int j = 0; j < len >> 2; j++) x_lp4[j] = x_lp[2 * j];
    CB_NOUNROLL for (int j = 0; j < lag >> 2; j++) y_lp4[j] = y[2 * j];
    const int xmax = maxabs16(x_lp4, len >> 2);
    const int ymax = maxabs16(y_lp4, lag >> 2);
    int shift = celt_ilog2(imax(1, imax(xmax, ymax))) - 11;
    if (shift > 0) {
        CB_NOUNROLL for (int j = 0; j < len >> 2; j++) x_lp4[j] = (int16_t)(x_lp4[j] >> shift);
        CB_NOUNROLL for (int j = 0; j < lag >> 2; j++) y_lp4[j] = (int16_t)(y_lp4[j] >> shift);
        shift *= 2;
    } else {
        shift = 0;
    }
    int maxcorr = pitch_xcorr(x_lp4, y_lp4, xcorr, len >> 2, max_pitch >> 2);
    find_best_pitch(xcorr, y_lp4, len >> 2, max_pitch >> 2, best_pitch, 0, maxcorr);
    maxcorr = 1;
    CB_NOUNROLL for (int i = 0; i < max_pitch >> 1; i++) {
        xcorr[i] = 0;
        if (iabs(i - 2 * best_pitch[0]) > 2 && iabs(i - 2 * best_pitch[1]) > 2) continue;
        int sum = 0;
        CB_NOUNROLL for (int j = 0; j < len >> 1; j++) sum = wadd(sum, mul16_16(x_lp[j], y[i + j]) >> shift);
        xcorr[i] = imax(-1, sum);
        maxcorr = imax(maxcorr, sum);
    }
    find_best_pitch(xcorr, y, len >> 1, max_pitch >> 1, best_pitch, shift + 1, maxcorr);
    int offset;
    if (best_pitch[0] > 0 && best_pitch[0] < (max_pitch >> 1) - 1) {
        int a = xcorr[best_pitch[0] - 1], b = xcorr[best_pitch[0]], c = xcorr[best_pitch[0] + 1];
        if (wsub(c, a) > mul16_32_q15(22938, wsub(b, a))) offset = 1;        // QCONST16(.7f,15)
        else if (wsub(a, c) > mul16_32_q15(22938, wsub(b, c))) offset = -1;
        else offset = 0;
    } else {
        offset = 0;
    }
    *pitch = 2 * best_pitch[0] - offset;
}

CB_TABLE uint8_t kSecondCheck[16] = {0, 0, 3, 2, 3, 2, 5, 2, 3, 2, 3, 2, 5, 2, 3, 2};   // pitch.c:371

// remove_doubling (pitch.c:372-505).  NOTE opus-fix keeps g, g0 32-bit (pitch.c:376,416-420), unlike upstream 1.1.2.
// yy_lookup: scratch for maxperiod/2+1 ints.
CB_DEV_NOINLINE int remove_doubling(const int16_t *x, int maxperiod, int minperiod, int N, int *T0_, int prev_period, int prev_gain,
                                    int *yy_lookup) {
    const int minperiod0 = minperiod;
    maxperiod /= 2; minperiod /= 2; *T0_ /= 2; prev_period /= 2; N /= 2;
    x += maxperiod;
    if (*T0_ >= maxperiod) *T0_ = maxperiod - 1;
    int T, T0;
    T = T0 = *T0_;
    int xx = 0, xy = 0;
    CB_NOUNROLL for (int i = 0; i < N; i++) { xx = mac16_16(xx, x[i], x[i]); xy = mac16_16(xy, x[i], x[i - T0]); }
    yy_lookup[0] = xx;
    int yy = xx;
    CB_NOUNROLL for (int i = 1; i <= maxperiod; i++) {
        yy = wsub(wadd(yy, mul16_16(x[-i], x[-i])), mul16_16(x[N - i], x[N - i]));
        yy_lookup[i] = imax(0, yy);
    }
    yy = yy_lookup[T0];
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
        int xy2 = 0;
        xy = 0;
        CB_NOUNROLL for (int i = 0; i < N; i++) { xy = mac16_16(xy, x[i], x[i - T1]); xy2 = mac16_16(xy2, x[i], x[i - T1b]); }
        xy = wadd(xy, xy2);
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
        int thresh = imax(9830, wsub(mul16_32_q15(22938, g0), cont));                       // .3, .7
        if (T1 < 3 * minperiod) thresh = imax(13107, wsub(mul16_32_q15(27853, g0), cont));   // .4, .85
        else if (T1 < 2 * minperiod) thresh = imax(16384, wsub(mul16_32_q15(29491, g0), cont));   // .5, .9
        if (g1 > thresh) {
            best_xy = xy; best_yy = yy; T = T1; g = g1;
        }
    }
    best_xy = imax(0, best_xy);
    int pg;
    if (best_yy <= best_xy) pg = 32767;
    else pg = s16(frac_div32(best_xy, wadd(best_yy, 1)) >> 16);
    int xc[3];
    CB_NOUNROLL for (int k = 0; k < 3; k++) xc[k] = inner_prod16(x, x - (T + k - 1), N);
    int offset;
    if (wsub(xc[2], xc[0]) > mul16_32_q15(22938, wsub(xc[1], xc[0]))) offset = 1;
    else if (wsub(xc[0], xc[2]) > mul16_32_q15(22938, wsub(xc[1], xc[2]))) offset = -1;
    else offset = 0;
    if (pg > g) pg = s16(g);
    *T0_ = 2 * T + offset;
    if (*T0_ < minperiod0) *T0_ = minperiod0;
    return pg;
}

// comb_filter, y != x (celt.c:183-244): a plain FIR over x's history.  window == nullptr <=> overlap == 0.
CB_DEV_NOINLINE void comb_filter_fir(int *y, const int *x, int T0, int T1, int N, int g0, int g1, int tapset0, int tapset1, int overlap) {
    if (g0 == 0 && g1 == 0) {
        CB_NOUNROLL for (int i = 0; i < N; i++) y[i] = x[i];
        return;
    }
    const int g00 = s16(mul16_16_p15(g0, kCombGains[tapset0][0]));
    const int g01 = s16(mul16_16_p15(g0, kCombGains[tapset0][1]));
    const int g02 = s16(mul16_16_p15(g0, kCombGains[tapset0][2]));
    const int g10 = s16(mul16_16_p15(g1, kCombGains[tapset1][0]));
    const int g11 = s16(mul16_16_p15(g1, kCombGains[tapset1][1]));
    const int g12 = s16(mul16_16_p15(g1, kCombGains[tapset1][2]));
    if (g0 == g1 && T0 == T1 && tapset0 == tapset1) overlap = 0;
    int i;
    CB_NOUNROLL for (i = 0; i < overlap; i++) {
        int f = s16(mul16_16_q15(kWindow120[i], kWindow120[i]));
        int nf = 32767 - f;
        int v = x[i];
        v = wadd(v, mul16_32_q15(mul16_16_q15(nf, g00), x[i - T0]));
        v = wadd(v, mul16_32_q15(mul16_16_q15(nf, g01), wadd(x[i - T0 + 1], x[i - T0 - 1])));
        v = wadd(v, mul16_32_q15(mul16_16_q15(nf, g02), wadd(x[i - T0 + 2], x[i - T0 - 2])));
        v = wadd(v, mul16_32_q15(mul16_16_q15(f, g10), x[i - T1]));
        v = wadd(v, mul16_32_q15(mul16_16_q15(f, g11), wadd(x[i - T1 + 1], x[i - T1 - 1])));
        v = wadd(v, mul16_32_q15(mul16_16_q15(f, g12), wadd(x[i - T1 + 2], x[i - T1 - 2])));
        y[i] = v;
    }
    if (g1 == 0) {
        CB_NOUNROLL for (; i < N; i++) y[i] = x[i];
        return;
    }
    CB_NOUNROLL for (; i < N; i++) {
        int v = x[i];
        v = wadd(v, mul16_32_q15(g10, x[i - T1]));
        v = wadd(v, mul16_32_q15(g11, wadd(x[i - T1 + 1], x[i - T1 - 1])));
        v = wadd(v, mul16_32_q15(g12, wadd(x[i - T1 + 2], x[i - T1 - 2])));
        y[i] = v;
    }
}

// ---- team reductions ---------------------------------------------------------------------------------------------
template <class TM>
CB_DEV int team_maxabs16(TM tm, const int16_t *x, int n) {   // celt_maxabs16 (mathops.h:49-60)
    int mx = 0, mn = 0;
    CB_TEAM_FOR(i, n, tm) { int v = x[i]; mx = imax(mx, v); mn = imin(mn, v); }
    mx = tm.max(mx);
    mn = ~tm.max(~mn);
    return imax(mx, -mn);
}
template <class TM>
CB_DEV int team_maxabs32(TM tm, const int *x, int n) {       // celt_maxabs32 (mathops.h:67-78)
    int mx = 0, mn = 0;
    CB_TEAM_FOR(i, n, tm) { int v = x[i]; mx = imax(mx, v); mn = imin(mn, v); }
    mx = tm.max(mx);
    mn = ~tm.max(~mn);
    return imax(mx, wneg(mn));
}
template <class TM>
CB_DEV int team_inner16(TM tm, const int16_t *x, const int16_t *y, int n) {
    int s = 0;
    CB_TEAM_FOR(i, n, tm) s = mac16_16(s, x[i], y[i]);
    return tm.sum(s);
}

// ---- pitch analysis, team versions (pitch.c) --------------------------------------------------------------------

// pitch_downsample (pitch.c:147-217): x0/x1 -> x_out[len/2] (x_raw: staging of the same size).
template <class TM>
CB_DEV_NOINLINE void pitch_downsample_team(TM tm, const int *x0, const int *x1, int len, int C, int16_t *x_raw, int16_t *x_out) {
    int maxabs = team_maxabs32(tm, x0, len);
    if (C == 2) maxabs = imax(maxabs, team_maxabs32(tm, x1, len));
    if (maxabs < 1) maxabs = 1;
    int shift = celt_ilog2(maxabs) - 10;
    if (shift < 0) shift = 0;
    if (C == 2) shift++;
    const int half = len >> 1;
    CB_TEAM_FOR(i, half, tm) {
        int v = i == 0 ? (wadd(x0[1] >> 1, x0[0]) >> 1) >> shift : (wadd(wadd(x0[2 * i - 1], x0[2 * i + 1]) >> 1, x0[2 * i]) >> 1) >> shift;
        v = s16(v);
        if (C == 2) {
            int w = i == 0 ? (wadd(x1[1] >> 1, x1[0]) >> 1) >> shift : (wadd(wadd(x1[2 * i - 1], x1[2 * i + 1]) >> 1, x1[2 * i]) >> 1) >> shift;
            v = s16(v + w);
        }
        x_raw[i] = (int16_t)v;
    }
    tm.sync();
    // _celt_autocorr (celt_lpc.c:232-328) with overlap = 0, lag = 4, n = half; x_out doubles as its scaled copy
    const int n = half;
    const int16_t *xptr = x_raw;
    int sh;
    {
        int acc = 0;
        CB_TEAM_FOR(i, n, tm) acc = wadd(acc, mul16_16(x_raw[i], x_raw[i]) >> 9);
        int ac0 = wadd(wadd(1, n << 7), tm.sum(acc));
        sh = celt_ilog2(ac0) - 30 + 10;
        sh = sh / 2;
        if (sh > 0) {
            CB_TEAM_FOR(i, n, tm) x_out[i] = (int16_t)pshr32(x_raw[i], sh);
            tm.sync();
            xptr = x_out;
        } else {
            sh = 0;
        }
    }
    int ac[5];
    {
        int a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0;
        CB_TEAM_FOR(j, n, tm) {
            const int xj = xptr[j];
            a0 = mac16_16(a0, xj, xj);
            if (j + 1 < n) a1 = mac16_16(a1, xj, xptr[j + 1]);
            if (j + 2 < n) a2 = mac16_16(a2, xj, xptr[j + 2]);
            if (j + 3 < n) a3 = mac16_16(a3, xj, xptr[j + 3]);
            if (j + 4 < n) a4 = mac16_16(a4, xj, xptr[j + 4]);
        }
        ac[0] = tm.sum(a0); ac[1] = tm.sum(a1); ac[2] = tm.sum(a2); ac[3] = tm.sum(a3); ac[4] = tm.sum(a4);
    }
    tm.sync();   // x_out is rewritten below
    sh = 2 * sh;
    if (sh <= 0) ac[0] = wadd(ac[0], shl32(1, -sh));
    if (ac[0] < 268435456) {
        int shift2 = 29 - ec_ilog((unsigned)ac[0]);
        CB_NOUNROLL for (int i = 0; i <= 4; i++) ac[i] = shl32(ac[i], shift2);
    } else if (ac[0] >= 536870912) {
        int shift2 = 1;
        if (ac[0] >= 1073741824) shift2++;
        CB_NOUNROLL for (int i = 0; i <= 4; i++) ac[i] = ac[i] >> shift2;
    }
    ac[0] = wadd(ac[0], ac[0] >> 13);
    CB_NOUNROLL for (int i = 1; i <= 4; i++) ac[i] = wsub(ac[i], mul16_32_q15(2 * i * i, ac[i]));
    int16_t lpc[4];
    celt_lpc(lpc, ac, 4);
    int tmp = 32767;
    CB_NOUNROLL for (int i = 0; i < 4; i++) {
        tmp = s16(mul16_16_q15(29491, tmp));
        lpc[i] = (int16_t)mul16_16_q15(lpc[i], tmp);
    }
    const int c1 = 26214;
    const int l0 = s16(lpc[0] + 3277);
    const int l1 = s16(lpc[1] + mul16_16_q15(c1, lpc[0]));
    const int l2 = s16(lpc[2] + mul16_16_q15(c1, lpc[1]));
    const int l3 = s16(lpc[3] + mul16_16_q15(c1, lpc[2]));
    const int l4 = s16(mul16_16_q15(c1, lpc[3]));
    // celt_fir5 (pitch.c:105-144) with zero initial memory: a plain FIR over the INPUT samples, so out of place it is parallel
    CB_TEAM_FOR(i, half, tm) {
        int sum = shl32(x_raw[i], 12);
        if (i >= 1) sum = mac16_16(sum, l0, x_raw[i - 1]);
        if (i >= 2) sum = mac16_16(sum, l1, x_raw[i - 2]);
        if (i >= 3) sum = mac16_16(sum, l2, x_raw[i - 3]);
        if (i >= 4) sum = mac16_16(sum, l3, x_raw[i - 4]);
        if (i >= 5) sum = mac16_16(sum, l4, x_raw[i - 5]);
        x_out[i] = (int16_t)round16(sum, 12);
    }
    tm.sync();
}

// find_best_pitch (pitch.c:45-103), team version.  The reference walks every lag i with a running window energy
//   Syy_0 = 1 + sum_{j<len} (y[j]^2 >> yshift),  Syy_{i+1} = max(1, Syy_i + (y[i+len]^2 >> yshift) - (y[i]^2 >> yshift))
// and tests a candidate only where xcorr[i] > 0.  The window energy is a prefix sum (wrapping adds, order-free); the team builds
// it once, verifies that the max(1, .) clamp can never engage (min over all lags >= 1 — it is 1 + a sum of non-negative terms
// unless 32-bit wrap-around interferes) and then only the candidate lags are visited, in increasing order, with the reference's
// comparisons.  If the clamp could engage the reference's sequential walk is used instead (never observed).
// `cand`/`ncand`: sorted lags to visit, or ncand < 0 for "every lag".  syy: scratch for max_pitch + 1 ints.
template <class TM>
CB_DEV_NOINLINE void find_best_pitch_team(TM tm, const int *xcorr, const int16_t *y, int len, int max_pitch, int *best_pitch, int yshift,
                                          int maxcorr, const int *cand, int ncand, int *syy) {
    int part = 0;
    CB_TEAM_FOR(j, len, tm) part = wadd(part, mul16_16(y[j], y[j]) >> yshift);
    const int syy0 = wadd(1, tm.sum(part));
    // exclusive prefix of d[i] = (y[i+len]^2 >> ys) - (y[i]^2 >> ys), lane-contiguous chunks
    const int per = (max_pitch + TM::W - 1) / TM::W;
    const int first = tm.lane() * per;
    int local = 0;
    CB_NOUNROLL for (int i = first; i < first + per && i < max_pitch; i++)
        local = wadd(local, wsub(mul16_16(y[i + len], y[i + len]) >> yshift, mul16_16(y[i], y[i]) >> yshift));
    int run = wadd(syy0, tm.exscan(local));
    int mn = 0x7fffffff;
    CB_NOUNROLL for (int i = first; i < first + per && i < max_pitch; i++) {
        syy[i] = run;
        mn = imin(mn, run);
        run = wadd(run, wsub(mul16_16(y[i + len], y[i + len]) >> yshift, mul16_16(y[i], y[i]) >> yshift));
    }
    mn = ~tm.max(~mn);
    tm.sync();
    if (mn < 1) {   // the clamp would change the walk: do exactly what the reference does
        find_best_pitch(xcorr, y, len, max_pitch, best_pitch, yshift, maxcorr);
        return;
    }
    int best_num[2] = {-1, -1};
    int best_den[2] = {0, 0};
    const int xshift = celt_ilog2(maxcorr) - 14;
    best_pitch[0] = 0;
    best_pitch[1] = 1;
    const int nvisit = ncand < 0 ? max_pitch : ncand;
    CB_NOUNROLL for (int k = 0; k < nvisit; k++) {
        const int i = ncand < 0 ? k : cand[k];
        const int xc = xcorr[i];
        if (xc > 0) {
            const int Syy = syy[i];
            int xcorr16 = s16(vshr32(xc, xshift));
            int num = s16(mul16_16_q15(xcorr16, xcorr16));
            if (mul16_32_q15(num, best_den[1]) > mul16_32_q15(best_num[1], Syy)) {
                if (mul16_32_q15(num, best_den[0]) > mul16_32_q15(best_num[0], Syy)) {
                    best_num[1] = best_num[0]; best_den[1] = best_den[0]; best_pitch[1] = best_pitch[0];
                    best_num[0] = num; best_den[0] = Syy; best_pitch[0] = i;
                } else {
                    best_num[1] = num; best_den[1] = Syy; best_pitch[1] = i;
                }
            }
        }
    }
}

// pitch_search (pitch.c:260-369)
template <class TM>
CB_DEV_NOINLINE int pitch_search_team(TM tm, const int16_t *x_lp, const int16_t *y, int len, int max_pitch, int16_t *x_lp4, int16_t *y_lp4,
                             int *xcorr, int *syy) {
    const int lag = len + max_pitch;
    int best_pitch[2] = {0, 0};
    CB_TEAM_FOR(j, len >> 2, tm) x_lp4[j] = x_lp[2 * j];
    CB_TEAM_FOR(j, lag >> 2, tm) y_lp4[j] = y[2 * j];
    tm.sync();
    const int xmax = team_maxabs16(tm, x_lp4, len >> 2);
    const int ymax = team_maxabs16(tm, y_lp4, lag >> 2);
    int shift = celt_ilog2(imax(1, imax(xmax, ymax))) - 11;
    if (shift > 0) {
        CB_TEAM_FOR(j, len >> 2, tm) x_lp4[j] = (int16_t)(x_lp4[j] >> shift);
        CB_TEAM_FOR(j, lag >> 2, tm) y_lp4[j] = (int16_t)(y_lp4[j] >> shift);
        tm.sync();
        shift *= 2;
    } else {
        shift = 0;
    }
    // coarse search, 4x decimated (celt_pitch_xcorr): each lane owns 8 CONSECUTIVE lags and slides an 8-sample register window
    // over y, so a tap costs one broadcast load of x, one load of y and 8 multiply-adds (244 lags = 31 lanes x 8).
    // y_lp4 is read up to 7 + 3 entries past its valid part (inside its array); those sums belong to lags >= np and are dropped.
    int maxcorr = 1;
    {
        const int n = len >> 2, np = max_pitch >> 2;
        CB_NOUNROLL for (int blk = tm.lane(); blk * 8 < np; blk += TM::W) {
            const int16_t *yb = y_lp4 + blk * 8;
            int a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0, a7 = 0;
            int w0 = yb[0], w1 = yb[1], w2 = yb[2], w3 = yb[3], w4 = yb[4], w5 = yb[5], w6 = yb[6], w7;
            int j = 0;
#define CB_XC_TAP(W0, W1, W2, W3, W4, W5, W6, W7)                                                          \
    {                                                                                                      \
        const int xj = x_lp4[j];                                                                           \
        W7 = yb[j + 7];                                                                                    \
        a0 += xj * W0; a1 += xj * W1; a2 += xj * W2; a3 += xj * W3; a4 += xj * W4; a5 += xj * W5; a6 += xj * W6; a7 += xj * W7; \
        j++;                                                                                               \
    }
            CB_NOUNROLL for (; j + 8 <= n;) {
                CB_XC_TAP(w0, w1, w2, w3, w4, w5, w6, w7)
                CB_XC_TAP(w1, w2, w3, w4, w5, w6, w7, w0)
                CB_XC_TAP(w2, w3, w4, w5, w6, w7, w0, w1)
                CB_XC_TAP(w3, w4, w5, w6, w7, w0, w1, w2)
                CB_XC_TAP(w4, w5, w6, w7, w0, w1, w2, w3)
                CB_XC_TAP(w5, w6, w7, w0, w1, w2, w3, w4)
                CB_XC_TAP(w6, w7, w0, w1, w2, w3, w4, w5)
                CB_XC_TAP(w7, w0, w1, w2, w3, w4, w5, w6)
            }
            CB_NOUNROLL for (; j < n;) {   // tail (frame sizes whose len/4 is not a multiple of 8): slide by hand
                CB_XC_TAP(w0, w1, w2, w3, w4, w5, w6, w7)
                w0 = w1; w1 = w2; w2 = w3; w3 = w4; w4 = w5; w5 = w6; w6 = w7;
            }
#undef CB_XC_TAP
            const int acc[8] = {a0, a1, a2, a3, a4, a5, a6, a7};
            for (int k = 0; k < 8; k++)
                if (blk * 8 + k < np) {
                    xcorr[blk * 8 + k] = acc[k];
                    maxcorr = imax(maxcorr, acc[k]);
                }
        }
        maxcorr = tm.max(maxcorr);
        tm.sync();
    }
    find_best_pitch_team(tm, xcorr, y_lp4, len >> 2, max_pitch >> 2, best_pitch, 0, maxcorr, nullptr, -1, syy);
    tm.sync();   // xcorr is rewritten below
    // finer search, 2x decimated, around the two candidates (at most 10 lags, visited in increasing order): lanes split the taps
    int cand[10];
    int ncand = 0;
    {
        const int c0 = 2 * best_pitch[0], c1 = 2 * best_pitch[1];
        const int lo = imin(c0, c1), hi = imax(c0, c1);
        for (int d = -2; d <= 2; d++)
            if (lo + d >= 0 && lo + d < (max_pitch >> 1)) cand[ncand++] = lo + d;
        for (int d = -2; d <= 2; d++)
            if (hi + d >= 0 && hi + d < (max_pitch >> 1) && hi + d > lo + 2) cand[ncand++] = hi + d;
    }
    maxcorr = 1;
    CB_TEAM_FOR(i, max_pitch >> 1, tm) xcorr[i] = 0;
    tm.sync();
    CB_NOUNROLL for (int k = 0; k < ncand; k++) {
        const int i = cand[k];
        int s = 0;
        CB_TEAM_FOR(j, len >> 1, tm) s = wadd(s, mul16_16(x_lp[j], y[i + j]) >> shift);
        s = tm.sum(s);
        if (tm.lane() == 0) xcorr[i] = imax(-1, s);
        maxcorr = imax(maxcorr, s);
    }
    tm.sync();
    find_best_pitch_team(tm, xcorr, y, len >> 1, max_pitch >> 1, best_pitch, shift + 1, maxcorr, cand, ncand, syy);
    int offset = 0;
    if (best_pitch[0] > 0 && best_pitch[0] < (max_pitch >> 1) - 1) {
        int a = xcorr[best_pitch[0] - 1], b = xcorr[best_pitch[0]], c = xcorr[best_pitch[0] + 1];
        if (wsub(c, a) > mul16_32_q15(22938, wsub(b, a))) offset = 1;
        else if (wsub(a, c) > mul16_32_q15(22938, wsub(b, c))) offset = -1;
    }
    tm.sync();
    return 2 * best_pitch[0] - offset;
}

}  // namespace cb
