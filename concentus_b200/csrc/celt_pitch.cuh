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

}  // namespace cb
