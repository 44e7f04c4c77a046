// celt_pvq.cuh — PVQ codeword (de)indexing and the decoder side of the spherical vector quantiser.
//
// Restates opus-fix/celt/cwrs.c:196-199,463-541 (CELT_PVQ_U/V, cwrsi, decode_pulses) and
// opus-fix/celt/vq.c:43-157,329-372 (exp_rotation1, exp_rotation, normalise_residual,
// extract_collapse_mask, alg_unquant, renormalise_vector).
//
// Stage A code: scalar, executed by the one thread that owns the frame; vectors are int16 arrays in
// that thread's local memory (ParseScratch).
// renormalise_vector is also used by stage B (anti-collapse) and is therefore team-templated.
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

// cwrsi (cwrs.c:463-537): index -> pulse vector y[0..n), returns sum y^2.
CB_DEV_NOINLINE int pvq_decode_index(int n, int k, unsigned i, int16_t *y) {
    unsigned p;
    int s, k0, val;
    int yy = 0;
    while (n > 2) {
        unsigned q;
        if (k >= n) {
            const int rown = kPvqURow[n];
            p = kPvqU[rown + k + 1];
            s = -(int)(i >= p);
            i -= p & (unsigned)s;
            k0 = k;
            q = kPvqU[rown + n];
            if (q > i) {
                k = n;
                do p = pvq_row(--k, n);
                while (p > i);
            } else {
                CB_NOUNROLL for (p = kPvqU[rown + k]; p > i; p = kPvqU[rown + k]) k--;
            }
            i -= p;
            val = s16((k0 - k + s) ^ s);
            *y++ = (int16_t)val;
            yy = mac16_16(yy, val, val);
        } else {
            p = pvq_row(k, n);
            q = pvq_row(k + 1, n);
            if (p <= i && i < q) {
                i -= p;
                *y++ = 0;
            } else {
                s = -(int)(i >= q);
                i -= q & (unsigned)s;
                k0 = k;
                do p = pvq_row(--k, n);
                while (p > i);
                i -= p;
                val = s16((k0 - k + s) ^ s);
                *y++ = (int16_t)val;
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
    *y++ = (int16_t)val;
    yy = mac16_16(yy, val, val);
    // n == 1
    s = -(int)i;
    val = s16((k + s) ^ s);
    *y = (int16_t)val;
    yy = mac16_16(yy, val, val);
    return yy;
}

// exp_rotation1 (vq.c:43-67): forward then backward in-place Givens sweep — order dependent.
CB_DEV_NOINLINE void exp_rotation1(int16_t *X, int len, int stride, int c, int s) {
    int ms = s16(-s);
    CB_NOUNROLL for (int i = 0; i < len - stride; i++) {
        int x1 = X[i], x2 = X[i + stride];
        X[i + stride] = (int16_t)pshr32(mac16_16(mul16_16(c, x2), s, x1), 15);
        X[i] = (int16_t)pshr32(mac16_16(mul16_16(c, x1), ms, x2), 15);
    }
    CB_NOUNROLL for (int i = len - 2 * stride - 1; i >= 0; i--) {
        int x1 = X[i], x2 = X[i + stride];
        X[i + stride] = (int16_t)pshr32(mac16_16(mul16_16(c, x2), s, x1), 15);
        X[i] = (int16_t)pshr32(mac16_16(mul16_16(c, x1), ms, x2), 15);
    }
}

// exp_rotation (vq.c:70-113), decoder direction (dir = -1).
CB_DEV void exp_rotation_dec(int16_t *X, int len, int stride, int K, int spread) {
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
    CB_NOUNROLL for (int i = 0; i < stride; i++) {
        int16_t *x = X + i * len;
        if (stride2) exp_rotation1(x, len, stride2, s, c);
        exp_rotation1(x, len, 1, c, s);
    }
}

// gain that maps a vector of energy E (Ryy) to unit norm times `gain` (vq.c:117-136 / :349-372 share it)
CB_DEV void unit_gain(int E, int gain, int &g, int &k) {
    k = celt_ilog2(E) >> 1;
    int t = vshr32(E, 2 * (k - 7));
    g = s16(mul16_16_p15(celt_rsqrt_norm(t), gain));
}

// alg_unquant (vq.c:329-346) once the codeword is known = cwrsi + normalise_residual + exp_rotation + extract_collapse_mask.
// The pulse vector is decoded straight into X and scaled in place.
CB_DEV unsigned alg_unquant_idx(int16_t *X, int N, int K, int spread, int B, unsigned idx, int gain) {
    int Ryy = pvq_decode_index(N, K, idx, X);
    int g, k;
    unit_gain(Ryy, gain, g, k);
    // normalise_residual (vq.c:117-136) fused with extract_collapse_mask (vq.c:139-157)
    unsigned mask = 0;
    if (B <= 1) {
        CB_NOUNROLL for (int i = 0; i < N; i++) X[i] = (int16_t)pshr32(mul16_16(g, X[i]), k + 1);
        mask = 1;
    } else {
        const int N0 = (int)udiv((unsigned)N, (unsigned)B);
        int i = 0;
        CB_NOUNROLL for (int blk = 0; blk < B; blk++) {
            int any = 0;
            CB_NOUNROLL for (int j = 0; j < N0; j++, i++) {
                int v = X[i];
                any |= v;
                X[i] = (int16_t)pshr32(mul16_16(g, v), k + 1);
            }
            mask |= (unsigned)(any != 0) << blk;
        }
        CB_NOUNROLL for (; i < N; i++) X[i] = (int16_t)pshr32(mul16_16(g, X[i]), k + 1);
    }
    exp_rotation_dec(X, N, B, K, spread);
    return mask;
}

// renormalise_vector (vq.c:349-372), team-parallel: Σx² is a wrapping 32-bit sum, so its order is free.
template <class TM>
CB_DEV void renormalise_vector(TM tm, int16_t *X, int N, int gain) {
    int acc = 0;
    CB_TEAM_FOR(i, N, tm) acc = mac16_16(acc, X[i], X[i]);
    int E = wadd(1, tm.sum(acc));
    int g, k;
    unit_gain(E, gain, g, k);
    CB_TEAM_FOR(i, N, tm) X[i] = (int16_t)pshr32(mul16_16(g, X[i]), k + 1);
    tm.sync();
}

}  // namespace cb
