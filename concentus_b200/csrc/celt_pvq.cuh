// celt_pvq.cuh — PVQ codeword (de)indexing and the decoder side of the spherical vector quantiser.
//
// Restates opus-fix/celt/cwrs.c:196-199,463-541 (CELT_PVQ_U/V, cwrsi, decode_pulses) and
// opus-fix/celt/vq.c:43-157,329-372 (exp_rotation1, exp_rotation, normalise_residual,
// extract_collapse_mask, alg_unquant, renormalise_vector).
//
// Execution model (see celt_simt.cuh): every lane of the team runs the scalar control flow
// redundantly (same registers, no divergence); vectors live in team-shared memory as int16 and
// elementwise work is strided over the lanes.  Order-dependent sweeps (the spreading rotation) run on
// lane 0.  Every helper that writes a vector ends with CB_SYNC() so the next reader sees it.
#pragma once
#include "celt_ec.cuh"
#include "celt_tables.cuh"

namespace cb {

// U(n,k) with n<=k folded onto the triangular table (cwrs.c:196)
CB_DEV unsigned pvq_u(int n, int k) {
    int lo = n < k ? n : k, hi = n < k ? k : n;
    return kPvqU[kPvqURow[lo] + hi];
}
CB_DEV unsigned pvq_v(int n, int k) { return pvq_u(n, k) + pvq_u(n, k + 1); }
CB_DEV unsigned pvq_row(int r, int c) { return kPvqU[kPvqURow[r] + c]; }

// cwrsi (cwrs.c:463-537): index -> pulse vector y[0..n), returns sum y^2.  Strictly sequential; all
// lanes walk it redundantly, lane 0 stores.
CB_DEV int pvq_decode_index(Team tm, int n, int k, unsigned i, int16_t *y) {
    unsigned p;
    int s, k0, val;
    int yy = 0;
    int pos = 0;
    while (n > 2) {
        unsigned q;
        if (k >= n) {
            p = pvq_row(n, k + 1);
            s = -(int)(i >= p);
            i -= p & (unsigned)s;
            k0 = k;
            q = pvq_row(n, n);
            if (q > i) {
                k = n;
                do p = pvq_row(--k, n);
                while (p > i);
            } else {
                for (p = pvq_row(n, k); p > i; p = pvq_row(n, k)) k--;
            }
            i -= p;
            val = s16((k0 - k + s) ^ s);
            if (tm.lane == 0) y[pos] = (int16_t)val;
            pos++;
            yy = mac16_16(yy, val, val);
        } else {
            p = pvq_row(k, n);
            q = pvq_row(k + 1, n);
            if (p <= i && i < q) {
                i -= p;
                if (tm.lane == 0) y[pos] = 0;
                pos++;
            } else {
                s = -(int)(i >= q);
                i -= q & (unsigned)s;
                k0 = k;
                do p = pvq_row(--k, n);
                while (p > i);
                i -= p;
                val = s16((k0 - k + s) ^ s);
                if (tm.lane == 0) y[pos] = (int16_t)val;
                pos++;
                yy = mac16_16(yy, val, val);
            }
        }
        n--;
    }
    // n == 2
    p = 2 * (unsigned)k + 1;
    s = -(int)(i >= p);
    i -= p & (unsigned)s;
    k0 = k;
    k = (int)((i + 1) >> 1);
    if (k) i -= 2 * (unsigned)k - 1;
    val = s16((k0 - k + s) ^ s);
    if (tm.lane == 0) y[pos] = (int16_t)val;
    pos++;
    yy = mac16_16(yy, val, val);
    // n == 1
    s = -(int)i;
    val = s16((k + s) ^ s);
    if (tm.lane == 0) y[pos] = (int16_t)val;
    yy = mac16_16(yy, val, val);
    CB_SYNC();
    return yy;
}

// exp_rotation1 (vq.c:43-67): forward then backward in-place Givens sweep — order dependent.
CB_DEV void exp_rotation1(int16_t *X, int len, int stride, int c, int s) {
    int ms = s16(-s);
    for (int i = 0; i < len - stride; i++) {
        int x1 = X[i], x2 = X[i + stride];
        X[i + stride] = (int16_t)pshr32(mac16_16(mul16_16(c, x2), s, x1), 15);
        X[i] = (int16_t)pshr32(mac16_16(mul16_16(c, x1), ms, x2), 15);
    }
    for (int i = len - 2 * stride - 1; i >= 0; i--) {
        int x1 = X[i], x2 = X[i + stride];
        X[i + stride] = (int16_t)pshr32(mac16_16(mul16_16(c, x2), s, x1), 15);
        X[i] = (int16_t)pshr32(mac16_16(mul16_16(c, x1), ms, x2), 15);
    }
}

// exp_rotation (vq.c:70-113).  The `stride` sub-vectors are independent: one lane each.
CB_DEV void exp_rotation(Team tm, int16_t *X, int len, int dir, int stride, int K, int spread) {
    if (2 * K >= len || spread == kSpreadNone) return;
    int factor = spread == 1 ? 15 : spread == 2 ? 10 : 5;
    int gain = s16(celt_div(mul16_16(32767, len), len + factor * K));
    int theta = mul16_16_q15(gain, gain) >> 1;
    int c = celt_cos_norm(theta);
    int s = celt_cos_norm(s16(32767 - theta));
    int stride2 = 0;
    if (len >= 8 * stride) {
        stride2 = 1;
        while ((stride2 * stride2 + stride2) * stride + (stride >> 2) < len) stride2++;
    }
    len = (int)udiv((unsigned)len, (unsigned)stride);
    CB_TEAM_FOR(i, stride, tm) {
        int16_t *x = X + i * len;
        if (dir < 0) {
            if (stride2) exp_rotation1(x, len, stride2, s, c);
            exp_rotation1(x, len, 1, c, s);
        } else {
            exp_rotation1(x, len, 1, c, s16(-s));
            if (stride2) exp_rotation1(x, len, stride2, s, s16(-c));
        }
    }
    CB_SYNC();
}

// gain that maps a vector of energy E (Ryy) to unit norm times `gain` (vq.c:117-136 / :349-372 share it)
CB_DEV void unit_gain(int E, int gain, int &g, int &k) {
    k = celt_ilog2(E) >> 1;
    int t = vshr32(E, 2 * (k - 7));
    g = s16(mul16_16_p15(celt_rsqrt_norm(t), gain));
}

// normalise_residual (vq.c:117-136)
CB_DEV void normalise_residual(Team tm, const int16_t *iy, int16_t *X, int N, int Ryy, int gain) {
    int g, k;
    unit_gain(Ryy, gain, g, k);
    CB_TEAM_FOR(i, N, tm) X[i] = (int16_t)pshr32(mul16_16(g, iy[i]), k + 1);
    CB_SYNC();
}

// extract_collapse_mask (vq.c:139-157)
CB_DEV unsigned extract_collapse_mask(Team tm, const int16_t *iy, int N, int B) {
    if (B <= 1) return 1;
    int N0 = (int)udiv((unsigned)N, (unsigned)B);
    unsigned mask = 0;
    CB_TEAM_FOR(j, N0 * B, tm) {
        if (iy[j] != 0) mask |= 1u << (j / N0);
    }
    return team_or(mask);
}

// alg_unquant (vq.c:329-346).  `iy` is team scratch for >= N int16.
CB_DEV unsigned alg_unquant(Team tm, int16_t *X, int N, int K, int spread, int B, EcDec &dec, int gain, int16_t *iy) {
    unsigned idx = dec.uint_(pvq_v(N, K));
    int Ryy = pvq_decode_index(tm, N, K, idx, iy);
    normalise_residual(tm, iy, X, N, Ryy, gain);
    exp_rotation(tm, X, N, -1, B, K, spread);
    return extract_collapse_mask(tm, iy, N, B);
}

// celt_inner_prod of a vector with itself (celt/pitch.h:160-169), wrapping 32-bit sum (order free).
CB_DEV int inner_prod_self(Team tm, const int16_t *x, int N) {
    int acc = 0;
    CB_TEAM_FOR(i, N, tm) acc = mac16_16(acc, x[i], x[i]);
    return team_sum(acc);
}

// renormalise_vector (vq.c:349-372)
CB_DEV void renormalise_vector(Team tm, int16_t *X, int N, int gain) {
    int E = wadd(1, inner_prod_self(tm, X, N));
    int g, k;
    unit_gain(E, gain, g, k);
    CB_TEAM_FOR(i, N, tm) X[i] = (int16_t)pshr32(mul16_16(g, X[i]), k + 1);
    CB_SYNC();
}

}  // namespace cb
