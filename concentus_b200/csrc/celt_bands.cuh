// celt_bands.cuh — the band loop: theta/split coding, recursive partitioning, folding, stereo
// recombination, anti-collapse and denormalisation (decoder side).
//
// Restates opus-fix/celt/bands.c:169-238 (denormalise_bands), :241-335 (anti_collapse), :375-424
// (stereo_merge), :532-592 (hadamard (de)interleave, haar1), :596-616 (compute_qn), :645-817
// (compute_theta), :819-859 (quant_band_n1), :864-1040 (quant_partition), :1044-1170 (quant_band),
// :1176-1335 (quant_band_stereo) and :1337-1502 (quant_all_bands), all with encode=0 / resynth=1.
//
// Execution model: the whole band loop is executed by every lane of the team with identical scalar
// state (range decoder, bit budgets, fill masks), so there is no divergence and no broadcast; the
// normalised spectrum X, the folding source `norm` and a 176-entry scratch live in team-shared
// memory and every vector operation is strided over the lanes.  The reference's recursion
// (quant_partition, depth <= maxLM+1 = 4 splits) is unrolled into five non-inlined template
// instances so the device needs no dynamic call stack.
#pragma once
#include "celt_pvq.cuh"
#include "celt_rate.cuh"

namespace cb {

struct BandCtx {
    Team tm;
    EcDec *ec;
    int16_t *tmp;        // >= 176 int16 of team scratch (pulse vector / hadamard staging)
    int i;               // band
    int intensity, spread, tf_change;
    int remaining_bits;
    unsigned seed;
};

struct SplitCtx {
    int inv, imid, iside, delta, itheta, qalloc;
};

// ---- small vector kernels -------------------------------------------------------------------------

// haar1 (bands.c:581-594): independent 2-point butterflies.
CB_DEV void haar1(Team tm, int16_t *X, int N0, int stride) {
    N0 >>= 1;
    CB_TEAM_FOR(p, N0 * stride, tm) {
        int i = p % stride, j = p / stride;
        int a = stride * 2 * j + i, b = stride * (2 * j + 1) + i;
        int t1 = mul16_16(23170, X[a]);
        int t2 = mul16_16(23170, X[b]);
        X[a] = (int16_t)pshr32(wadd(t1, t2), 15);
        X[b] = (int16_t)pshr32(wsub(t1, t2), 15);
    }
    CB_SYNC();
}

// deinterleave_hadamard / interleave_hadamard (bands.c:532-579) through the team scratch.
CB_DEV void deinterleave_hadamard(Team tm, int16_t *X, int16_t *tmp, int N0, int stride, int hadamard) {
    int N = N0 * stride;
    const uint8_t *ordery = kOrdery + stride - 2;
    CB_TEAM_FOR(p, N, tm) {
        int i = p % stride, j = p / stride;   // source index p = j*stride+i
        int row = hadamard ? ordery[i] : i;
        tmp[row * N0 + j] = X[p];
    }
    CB_SYNC();
    CB_TEAM_FOR(p, N, tm) X[p] = tmp[p];
    CB_SYNC();
}
CB_DEV void interleave_hadamard(Team tm, int16_t *X, int16_t *tmp, int N0, int stride, int hadamard) {
    int N = N0 * stride;
    const uint8_t *ordery = kOrdery + stride - 2;
    CB_TEAM_FOR(p, N, tm) {
        int i = p % stride, j = p / stride;   // destination index p = j*stride+i
        int row = hadamard ? ordery[i] : i;
        tmp[p] = X[row * N0 + j];
    }
    CB_SYNC();
    CB_TEAM_FOR(p, N, tm) X[p] = tmp[p];
    CB_SYNC();
}

// stereo_merge (bands.c:375-424)
CB_DEV void stereo_merge(Team tm, int16_t *X, int16_t *Y, int mid, int N) {
    int xp = 0, side = 0;
    CB_TEAM_FOR(j, N, tm) {
        xp = mac16_16(xp, Y[j], X[j]);
        side = mac16_16(side, Y[j], Y[j]);
    }
    xp = team_sum(xp);
    side = team_sum(side);
    xp = mul16_32_q15(mid, xp);
    int mid2 = s16(mid >> 1);
    int El = wsub(wadd(mul16_16(mid2, mid2), side), wmul(2, xp));
    int Er = wadd(wadd(mul16_16(mid2, mid2), side), wmul(2, xp));
    if (Er < 161061 || El < 161061) {   // QCONST32(6e-4f, 28)
        CB_TEAM_FOR(j, N, tm) Y[j] = X[j];
        CB_SYNC();
        return;
    }
    int kl = celt_ilog2(El) >> 1;
    int kr = celt_ilog2(Er) >> 1;
    int t = vshr32(El, (kl - 7) << 1);
    int lgain = celt_rsqrt_norm(t);
    t = vshr32(Er, (kr - 7) << 1);
    int rgain = celt_rsqrt_norm(t);
    if (kl < 7) kl = 7;
    if (kr < 7) kr = 7;
    CB_TEAM_FOR(j, N, tm) {
        int l = s16(mul16_16_p15(mid, X[j]));
        int r = Y[j];
        X[j] = (int16_t)pshr32(mul16_16(lgain, s16(l - r)), kl + 1);
        Y[j] = (int16_t)pshr32(mul16_16(rgain, s16(l + r)), kr + 1);
    }
    CB_SYNC();
}

// compute_qn (bands.c:596-620)
CB_DEV int compute_qn(int N, int b, int offset, int pulse_cap, int stereo) {
    int N2 = 2 * N - 1;
    if (stereo && N == 2) N2--;
    int qb = sudiv(b + N2 * offset, N2);
    qb = imin(b - pulse_cap - (4 << kBitRes), qb);
    qb = imin(8 << kBitRes, qb);
    int qn;
    if (qb < (1 << kBitRes >> 1)) {
        qn = 1;
    } else {
        qn = kExp2Table8[qb & 0x7] >> (14 - (qb >> kBitRes));
        qn = (qn + 1) >> 1 << 1;
    }
    return qn;
}

// compute_theta, decoder half (bands.c:645-817): reads itheta with the pdf the split type calls for.
CB_DEV void compute_theta(BandCtx &ctx, SplitCtx &sctx, int N, int *b, int B, int B0, int LM, int stereo, int *fill) {
    EcDec &ec = *ctx.ec;
    int itheta = 0, inv = 0;
    int pulse_cap = kLogN[ctx.i] + LM * (1 << kBitRes);
    int offset = (pulse_cap >> 1) - (stereo && N == 2 ? kQThetaOffsetTwoPhase : kQThetaOffset);
    int qn = compute_qn(N, *b, offset, pulse_cap, stereo);
    if (stereo && ctx.i >= ctx.intensity) qn = 1;
    int tell = (int)ec.tell_frac();
    if (qn != 1) {
        if (stereo && N > 2) {
            const int p0 = 3;
            int x0 = qn / 2;
            unsigned ft = (unsigned)(p0 * (x0 + 1) + x0);
            int fs = (int)ec.decode(ft);
            int x;
            if (fs < (x0 + 1) * p0) x = fs / p0;
            else x = x0 + 1 + (fs - (x0 + 1) * p0);
            ec.update((unsigned)(x <= x0 ? p0 * x : (x - 1 - x0) + (x0 + 1) * p0),
                      (unsigned)(x <= x0 ? p0 * (x + 1) : (x - x0) + (x0 + 1) * p0), ft);
            itheta = x;
        } else if (B0 > 1 || stereo) {
            itheta = (int)ec.uint_((unsigned)qn + 1);
        } else {
            int fs = 1, fl = 0;
            int ft = ((qn >> 1) + 1) * ((qn >> 1) + 1);
            int fm = (int)ec.decode((unsigned)ft);
            if (fm < ((qn >> 1) * ((qn >> 1) + 1) >> 1)) {
                itheta = (int)((isqrt32(8 * (unsigned)fm + 1) - 1) >> 1);
                fs = itheta + 1;
                fl = itheta * (itheta + 1) >> 1;
            } else {
                itheta = (int)((2 * (unsigned)(qn + 1) - isqrt32(8 * (unsigned)(ft - fm - 1) + 1)) >> 1);
                fs = qn + 1 - itheta;
                fl = ft - ((qn + 1 - itheta) * (qn + 2 - itheta) >> 1);
            }
            ec.update((unsigned)fl, (unsigned)(fl + fs), (unsigned)ft);
        }
        itheta = (int)udiv((unsigned)(itheta * 16384), (unsigned)qn);
    } else if (stereo) {
        if (*b > 2 << kBitRes && ctx.remaining_bits > 2 << kBitRes) inv = ec.bit_logp(2);
        else inv = 0;
        itheta = 0;
    }
    int qalloc = (int)ec.tell_frac() - tell;
    *b -= qalloc;

    int imid, iside, delta;
    if (itheta == 0) {
        imid = 32767; iside = 0;
        *fill &= (1 << B) - 1;
        delta = -16384;
    } else if (itheta == 16384) {
        imid = 0; iside = 32767;
        *fill &= ((1 << B) - 1) << B;
        delta = 16384;
    } else {
        imid = bitexact_cos(s16(itheta));
        iside = bitexact_cos(s16(16384 - itheta));
        delta = frac_mul16((N - 1) << 7, bitexact_log2tan(iside, imid));
    }
    sctx.inv = inv; sctx.imid = imid; sctx.iside = iside;
    sctx.delta = delta; sctx.itheta = itheta; sctx.qalloc = qalloc;
}

// quant_band_n1 (bands.c:819-859)
CB_DEV unsigned quant_band_n1(BandCtx &ctx, int16_t *X, int16_t *Y, int16_t *lowband_out) {
    EcDec &ec = *ctx.ec;
    int16_t *x = X;
    int nch = Y != nullptr ? 2 : 1;
    for (int c = 0; c < nch; c++) {
        int sign = 0;
        if (ctx.remaining_bits >= 1 << kBitRes) {
            sign = (int)ec.bits(1);
            ctx.remaining_bits -= 1 << kBitRes;
        }
        if (ctx.tm.lane == 0) x[0] = sign ? -16384 : 16384;
        x = Y;
    }
    CB_SYNC();
    if (lowband_out && ctx.tm.lane == 0) lowband_out[0] = (int16_t)(X[0] >> 4);
    CB_SYNC();
    return 1;
}

// quant_partition (bands.c:864-1040).  D = remaining split depth.
template <int D>
CB_DEV_NOINLINE unsigned quant_partition(BandCtx &ctx, int16_t *X, int N, int b, int B, int16_t *lowband, int LM, int gain, int fill) {
    const Team tm = ctx.tm;
    const uint8_t *cache = pulse_cache(ctx.i, LM);
    unsigned cm = 0;
    bool split = false;
    if constexpr (D > 0) split = (LM != -1 && b > cache[cache[0]] + 12 && N > 2);
    if (split) {
        if constexpr (D > 0) {
            int B0 = B;
            SplitCtx s;
            N >>= 1;
            int16_t *Y = X + N;
            LM -= 1;
            if (B == 1) fill = (fill & 1) | (fill << 1);
            B = (B + 1) >> 1;
            compute_theta(ctx, s, N, &b, B, B0, LM, 0, &fill);
            int mid = s.imid, side = s.iside, delta = s.delta, itheta = s.itheta;
            if (B0 > 1 && (itheta & 0x3fff)) {
                if (itheta > 8192) delta -= delta >> (4 - LM);
                else delta = imin(0, delta + (N << kBitRes >> (5 - LM)));
            }
            int mbits = imax(0, imin(b, (b - delta) / 2));
            int sbits = b - mbits;
            ctx.remaining_bits -= s.qalloc;
            int16_t *next_lowband2 = lowband ? lowband + N : nullptr;
            int rebalance = ctx.remaining_bits;
            const int gmid = s16(mul16_16_p15(gain, mid));
            const int gside = s16(mul16_16_p15(gain, side));
            if (mbits >= sbits) {
                cm = quant_partition<D - 1>(ctx, X, N, mbits, B, lowband, LM, gmid, fill);
                rebalance = mbits - (rebalance - ctx.remaining_bits);
                if (rebalance > 3 << kBitRes && itheta != 0) sbits += rebalance - (3 << kBitRes);
                cm |= quant_partition<D - 1>(ctx, Y, N, sbits, B, next_lowband2, LM, gside, fill >> B) << (B0 >> 1);
            } else {
                cm = quant_partition<D - 1>(ctx, Y, N, sbits, B, next_lowband2, LM, gside, fill >> B) << (B0 >> 1);
                rebalance = sbits - (rebalance - ctx.remaining_bits);
                if (rebalance > 3 << kBitRes && itheta != 16384) mbits += rebalance - (3 << kBitRes);
                cm |= quant_partition<D - 1>(ctx, X, N, mbits, B, lowband, LM, gmid, fill);
            }
        }
    } else {
        int q = bits2pulses(ctx.i, LM, b);
        int curr_bits = pulses2bits(ctx.i, LM, q);
        ctx.remaining_bits -= curr_bits;
        while (ctx.remaining_bits < 0 && q > 0) {
            ctx.remaining_bits += curr_bits;
            q--;
            curr_bits = pulses2bits(ctx.i, LM, q);
            ctx.remaining_bits -= curr_bits;
        }
        if (q != 0) {
            int K = get_pulses(q);
            cm = alg_unquant(tm, X, N, K, ctx.spread, B, *ctx.ec, gain, ctx.tmp);
        } else {
            unsigned cm_mask = (1u << B) - 1;
            fill &= (int)cm_mask;
            if (!fill) {
                CB_TEAM_FOR(j, N, tm) X[j] = 0;
                CB_SYNC();
            } else {
                // The LCG advances once per coefficient: every lane steps the (cheap) generator through
                // the whole band so the seed stays identical team-wide, and stores only its own slots.
                if (lowband == nullptr) {
                    unsigned sd = ctx.seed;
                    for (int j = 0; j < N; j++) {
                        sd = lcg_rand(sd);
                        if ((j % CB_LANES) == tm.lane) X[j] = (int16_t)((int)sd >> 20);
                    }
                    ctx.seed = sd;
                    cm = cm_mask;
                } else {
                    unsigned sd = ctx.seed;
                    for (int j = 0; j < N; j++) {
                        sd = lcg_rand(sd);
                        if ((j % CB_LANES) == tm.lane) {
                            int t = (sd & 0x8000) ? 4 : -4;   // QCONST16(1.0f/256, 10)
                            X[j] = (int16_t)(lowband[j] + t);
                        }
                    }
                    ctx.seed = sd;
                    cm = (unsigned)fill;
                }
                CB_SYNC();
                renormalise_vector(tm, X, N, gain);
            }
        }
    }
    return cm;
}

// quant_band (bands.c:1044-1170)
CB_DEV_NOINLINE unsigned quant_band(BandCtx &ctx, int16_t *X, int N, int b, int B, int16_t *lowband, int LM,
                                    int16_t *lowband_out, int gain, int16_t *lowband_scratch, int fill) {
    const Team tm = ctx.tm;
    int N0 = N, N_B = N, N_B0, B0 = B;
    int time_divide = 0, recombine = 0;
    int tf_change = ctx.tf_change;
    int longBlocks = B0 == 1;
    unsigned cm = 0;
    N_B = (int)udiv((unsigned)N_B, (unsigned)B);
    if (N == 1) return quant_band_n1(ctx, X, nullptr, lowband_out);
    if (tf_change > 0) recombine = tf_change;
    if (lowband_scratch && lowband && (recombine || ((N_B & 1) == 0 && tf_change < 0) || B0 > 1)) {
        CB_TEAM_FOR(j, N, tm) lowband_scratch[j] = lowband[j];
        CB_SYNC();
        lowband = lowband_scratch;
    }
    for (int k = 0; k < recombine; k++) {
        if (lowband) haar1(tm, lowband, N >> k, 1 << k);
        fill = kBitInterleave[fill & 0xF] | kBitInterleave[fill >> 4] << 2;
    }
    B >>= recombine;
    N_B <<= recombine;
    while ((N_B & 1) == 0 && tf_change < 0) {
        if (lowband) haar1(tm, lowband, N_B, B);
        fill |= fill << B;
        B <<= 1;
        N_B >>= 1;
        time_divide++;
        tf_change++;
    }
    B0 = B;
    N_B0 = N_B;
    if (B0 > 1 && lowband) deinterleave_hadamard(tm, lowband, ctx.tmp, N_B >> recombine, B0 << recombine, longBlocks);

    cm = quant_partition<4>(ctx, X, N, b, B, lowband, LM, gain, fill);

    // resynthesis (decoder): undo the reorganisation
    if (B0 > 1) interleave_hadamard(tm, X, ctx.tmp, N_B >> recombine, B0 << recombine, longBlocks);
    N_B = N_B0;
    B = B0;
    for (int k = 0; k < time_divide; k++) {
        B >>= 1;
        N_B <<= 1;
        cm |= cm >> B;
        haar1(tm, X, N_B, B);
    }
    for (int k = 0; k < recombine; k++) {
        cm = kBitDeinterleave[cm];
        haar1(tm, X, N0 >> k, 1 << k);
    }
    B <<= recombine;
    if (lowband_out) {
        int n = s16(celt_sqrt(shl32(N0, 22)));
        CB_TEAM_FOR(j, N0, tm) lowband_out[j] = (int16_t)mul16_16_q15(n, X[j]);
        CB_SYNC();
    }
    cm &= (1u << B) - 1;
    return cm;
}

// quant_band_stereo (bands.c:1176-1335)
CB_DEV_NOINLINE unsigned quant_band_stereo(BandCtx &ctx, int16_t *X, int16_t *Y, int N, int b, int B, int16_t *lowband, int LM,
                                           int16_t *lowband_out, int16_t *lowband_scratch, int fill) {
    const Team tm = ctx.tm;
    EcDec &ec = *ctx.ec;
    unsigned cm = 0;
    if (N == 1) return quant_band_n1(ctx, X, Y, lowband_out);
    int orig_fill = fill;
    SplitCtx s;
    compute_theta(ctx, s, N, &b, B, B, LM, 1, &fill);
    int inv = s.inv, mid = s.imid, side = s.iside, delta = s.delta, itheta = s.itheta, qalloc = s.qalloc;
    if (N == 2) {
        int mbits = b, sbits = 0;
        if (itheta != 0 && itheta != 16384) sbits = 1 << kBitRes;
        mbits -= sbits;
        int c = itheta > 8192;
        ctx.remaining_bits -= qalloc + sbits;
        int16_t *x2 = c ? Y : X;
        int16_t *y2 = c ? X : Y;
        int sign = 0;
        if (sbits) sign = (int)ec.bits(1);
        sign = 1 - 2 * sign;
        cm = quant_band(ctx, x2, N, mbits, B, lowband, LM, lowband_out, 32767, lowband_scratch, orig_fill);
        if (tm.lane == 0) {
            y2[0] = (int16_t)(-sign * x2[1]);
            y2[1] = (int16_t)(sign * x2[0]);
            X[0] = (int16_t)mul16_16_q15(mid, X[0]);
            X[1] = (int16_t)mul16_16_q15(mid, X[1]);
            Y[0] = (int16_t)mul16_16_q15(side, Y[0]);
            Y[1] = (int16_t)mul16_16_q15(side, Y[1]);
            int t = X[0];
            X[0] = (int16_t)(t - Y[0]);
            Y[0] = (int16_t)(t + Y[0]);
            t = X[1];
            X[1] = (int16_t)(t - Y[1]);
            Y[1] = (int16_t)(t + Y[1]);
        }
        CB_SYNC();
    } else {
        int mbits = imax(0, imin(b, (b - delta) / 2));
        int sbits = b - mbits;
        ctx.remaining_bits -= qalloc;
        int rebalance = ctx.remaining_bits;
        if (mbits >= sbits) {
            cm = quant_band(ctx, X, N, mbits, B, lowband, LM, lowband_out, 32767, lowband_scratch, fill);
            rebalance = mbits - (rebalance - ctx.remaining_bits);
            if (rebalance > 3 << kBitRes && itheta != 0) sbits += rebalance - (3 << kBitRes);
            cm |= quant_band(ctx, Y, N, sbits, B, nullptr, LM, nullptr, side, nullptr, fill >> B);
        } else {
            cm = quant_band(ctx, Y, N, sbits, B, nullptr, LM, nullptr, side, nullptr, fill >> B);
            rebalance = sbits - (rebalance - ctx.remaining_bits);
            if (rebalance > 3 << kBitRes && itheta != 16384) mbits += rebalance - (3 << kBitRes);
            cm |= quant_band(ctx, X, N, mbits, B, lowband, LM, lowband_out, 32767, lowband_scratch, fill);
        }
    }
    if (N != 2) stereo_merge(tm, X, Y, mid, N);
    if (inv) {
        CB_TEAM_FOR(j, N, tm) Y[j] = (int16_t)(-Y[j]);
        CB_SYNC();
    }
    return cm;
}

// quant_all_bands, decoder (bands.c:1337-1502).  X_: C*N int16 (channel-major), norm: C*(M*eBands[20]) int16.
CB_DEV void quant_all_bands_dec(Team tm, int start, int end, int16_t *X_, int16_t *Y_, uint8_t *collapse_masks,
                                const int *pulses, int shortBlocks, int spread, int dual_stereo, int intensity,
                                const int *tf_res, int total_bits, int balance, EcDec &ec, int LM, int codedBands,
                                unsigned *seed, int16_t *norm, int16_t *tmp) {
    const int M = 1 << LM;
    const int B = shortBlocks ? M : 1;
    const int C = Y_ != nullptr ? 2 : 1;
    const int norm_offset = M * kEBands[start];
    int16_t *norm2 = norm + M * kEBands[kNbEBands - 1] - norm_offset;
    int16_t *lowband_scratch = X_ + M * kEBands[kNbEBands - 1];
    int lowband_offset = 0;
    int update_lowband = 1;
    BandCtx ctx;
    ctx.tm = tm; ctx.ec = &ec; ctx.tmp = tmp;
    ctx.intensity = intensity; ctx.spread = spread; ctx.seed = *seed;
    for (int i = start; i < end; i++) {
        ctx.i = i;
        int last = (i == end - 1);
        int16_t *X = X_ + M * kEBands[i];
        int16_t *Y = Y_ != nullptr ? Y_ + M * kEBands[i] : nullptr;
        int N = M * kEBands[i + 1] - M * kEBands[i];
        int tell = (int)ec.tell_frac();
        if (i != start) balance -= tell;
        int remaining_bits = total_bits - tell - 1;
        ctx.remaining_bits = remaining_bits;
        int b;
        if (i <= codedBands - 1) {
            int curr_balance = sudiv(balance, imin(3, codedBands - i));
            b = imax(0, imin(16383, imin(remaining_bits + 1, pulses[i] + curr_balance)));
        } else {
            b = 0;
        }
        if (M * kEBands[i] - N >= M * kEBands[start] && (update_lowband || lowband_offset == 0)) lowband_offset = i;
        int tf_change = tf_res[i];
        ctx.tf_change = tf_change;
        if (i == end - 1) lowband_scratch = nullptr;

        int effective_lowband = -1;
        unsigned x_cm, y_cm;
        if (lowband_offset != 0 && (spread != kSpreadAggressive || B > 1 || tf_change < 0)) {
            effective_lowband = imax(0, M * kEBands[lowband_offset] - norm_offset - N);
            int fold_start = lowband_offset;
            while (M * kEBands[--fold_start] > effective_lowband + norm_offset) {}
            int fold_end = lowband_offset - 1;
            while (M * kEBands[++fold_end] < effective_lowband + norm_offset + N) {}
            x_cm = y_cm = 0;
            int fold_i = fold_start;
            do {
                x_cm |= collapse_masks[fold_i * C + 0];
                y_cm |= collapse_masks[fold_i * C + C - 1];
            } while (++fold_i < fold_end);
        } else {
            x_cm = y_cm = (1u << B) - 1;
        }
        if (dual_stereo && i == intensity) {
            dual_stereo = 0;
            CB_TEAM_FOR(j, M * kEBands[i] - norm_offset, tm) norm[j] = (int16_t)((norm[j] + norm2[j]) >> 1);
            CB_SYNC();
        }
        int16_t *lb = effective_lowband != -1 ? norm + effective_lowband : nullptr;
        int16_t *lb_out = last ? nullptr : norm + M * kEBands[i] - norm_offset;
        if (dual_stereo) {
            int16_t *lb2 = effective_lowband != -1 ? norm2 + effective_lowband : nullptr;
            int16_t *lb2_out = last ? nullptr : norm2 + M * kEBands[i] - norm_offset;
            x_cm = quant_band(ctx, X, N, b / 2, B, lb, LM, lb_out, 32767, lowband_scratch, (int)x_cm);
            y_cm = quant_band(ctx, Y, N, b / 2, B, lb2, LM, lb2_out, 32767, lowband_scratch, (int)y_cm);
        } else {
            if (Y != nullptr) x_cm = quant_band_stereo(ctx, X, Y, N, b, B, lb, LM, lb_out, lowband_scratch, (int)(x_cm | y_cm));
            else x_cm = quant_band(ctx, X, N, b, B, lb, LM, lb_out, 32767, lowband_scratch, (int)(x_cm | y_cm));
            y_cm = x_cm;
        }
        if (tm.lane == 0) {
            collapse_masks[i * C + 0] = (uint8_t)x_cm;
            collapse_masks[i * C + C - 1] = (uint8_t)y_cm;
        }
        CB_SYNC();
        balance += pulses[i] + tell;
        update_lowband = b > (N << kBitRes);
    }
    *seed = ctx.seed;
}

// anti_collapse (bands.c:241-335)
CB_DEV void anti_collapse(Team tm, int16_t *X_, const uint8_t *collapse_masks, int LM, int C, int size, int start, int end,
                          const int16_t *logE, const int16_t *prev1logE, const int16_t *prev2logE, const int *pulses, unsigned seed) {
    for (int i = start; i < end; i++) {
        int N0 = band_width(i);
        int depth = (int)udiv((unsigned)(1 + pulses[i]), (unsigned)N0) >> LM;
        int thresh32 = celt_exp2(s16(-shl16(depth, 10 - kBitRes))) >> 1;
        int thresh = s16(mul16_32_q15(16384, imin(32767, thresh32)));
        int t = N0 << LM;
        int shift = celt_ilog2(t) >> 1;
        t = shl32(t, (7 - shift) << 1);
        int sqrt_1 = celt_rsqrt_norm(t);
        for (int c = 0; c < C; c++) {
            int prev1 = prev1logE[c * kNbEBands + i];
            int prev2 = prev2logE[c * kNbEBands + i];
            if (C == 1) {
                prev1 = imax(prev1, (int)prev1logE[kNbEBands + i]);
                prev2 = imax(prev2, (int)prev2logE[kNbEBands + i]);
            }
            int Ediff = (int)logE[c * kNbEBands + i] - imin(prev1, prev2);
            Ediff = imax(0, Ediff);
            int r;
            if (Ediff < 16384) {
                int r32 = celt_exp2(s16(-s16(Ediff))) >> 1;
                r = s16(2 * imin(16383, r32));
            } else {
                r = 0;
            }
            if (LM == 3) r = s16(mul16_16_q14(23170, imin(23169, r)));
            r = s16(imin(thresh, r) >> 1);
            r = s16(mul16_16_q15(sqrt_1, r) >> shift);
            int16_t *X = X_ + c * size + (kEBands[i] << LM);
            int renormalize = 0;
            unsigned mask = collapse_masks[i * C + c];
            for (int k = 0; k < 1 << LM; k++) {
                if (!(mask & (1u << k))) {
                    for (int j = 0; j < N0; j++) {
                        seed = lcg_rand(seed);
                        if ((j % CB_LANES) == tm.lane) X[(j << LM) + k] = (int16_t)((seed & 0x8000) ? r : -r);
                    }
                    renormalize = 1;
                }
            }
            CB_SYNC();
            if (renormalize) renormalise_vector(tm, X, N0 << LM, 32767);
        }
    }
}

}  // namespace cb
