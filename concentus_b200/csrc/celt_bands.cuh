// celt_bands.cuh — the band loop: theta/split coding, recursive partitioning, folding, stereo
// recombination (decoder side, stage A) and anti-collapse (stage B).
//
// Restates opus-fix/celt/bands.c:241-335 (anti_collapse), :375-424 (stereo_merge), :532-592 (hadamard
// (de)interleave, haar1), :596-616 (compute_qn), :645-817 (compute_theta), :819-859 (quant_band_n1),
// :864-1040 (quant_partition), :1044-1170 (quant_band), :1176-1335 (quant_band_stereo) and :1337-1502
// (quant_all_bands), all with encode=0 / resynth=1.
//
// Stage A is scalar: one thread walks the whole band loop of its frame with the range decoder in
// registers.  Two structural changes keep the instruction footprint small (the first CUDA version was
// instruction-cache bound, profiles/r1_v1_decode_ncu_summary.md):
//   * quant_partition's recursion (depth <= maxLM+1 splits) is an explicit walker over a 5-entry frame
//     stack, so the leaf code (PVQ decode / fold / noise) exists exactly once;
//   * a band issues its (up to two) quant_band calls from one loop, so quant_band is instantiated once
//     instead of at the reference's six call sites.
#pragma once
#include "celt_pvq.cuh"
#include "celt_rate.cuh"

namespace cb {

struct BandCtx {
    EcDec ec;            // by value: lives in registers for the whole band loop
    int16_t *tmp;        // >= 176 int16 of thread scratch (pulse vector / hadamard staging)
    int i;               // band
    int intensity, spread, tf_change;
    int remaining_bits;
    unsigned seed;
    bool dry;            // seed-recovery pass: read every symbol (identical range-coder walk), skip all vector work
};

struct SplitCtx {
    int inv, imid, iside, delta, itheta, qalloc;
};

// ---- small vector kernels (scalar) ------------------------------------------------------------------

// haar1 (bands.c:581-594)
CB_DEV_NOINLINE void haar1(int16_t *X, int N0, int stride) {
    N0 >>= 1;
    CB_NOUNROLL for (int i = 0; i < stride; i++)
        CB_NOUNROLL for (int j = 0; j < N0; j++) {
            int a = stride * 2 * j + i, b = stride * (2 * j + 1) + i;
            int t1 = mul16_16(23170, X[a]);
            int t2 = mul16_16(23170, X[b]);
            X[a] = (int16_t)pshr32(wadd(t1, t2), 15);
            X[b] = (int16_t)pshr32(wsub(t1, t2), 15);
        }
}

// deinterleave_hadamard / interleave_hadamard (bands.c:532-579) through the thread scratch (never aliases X).
CB_DEV_NOINLINE void deinterleave_hadamard(int16_t *__restrict__ X, int16_t *__restrict__ tmp, int N0, int stride, int hadamard) {
    const int N = N0 * stride;
    const uint8_t *ordery = kOrdery + stride - 2;
    CB_NOUNROLL for (int i = 0; i < stride; i++) {
        const int row = hadamard ? ordery[i] : i;
        int16_t *d = tmp + row * N0;
        const int16_t *x = X + i;
        CB_NOUNROLL for (int j = 0; j < N0; j++) d[j] = x[j * stride];
    }
    CB_NOUNROLL for (int p = 0; p < N; p++) X[p] = tmp[p];
}
CB_DEV_NOINLINE void interleave_hadamard(int16_t *__restrict__ X, int16_t *__restrict__ tmp, int N0, int stride, int hadamard) {
    const int N = N0 * stride;
    const uint8_t *ordery = kOrdery + stride - 2;
    CB_NOUNROLL for (int i = 0; i < stride; i++) {
        const int row = hadamard ? ordery[i] : i;
        int16_t *d = tmp + i;
        const int16_t *x = X + row * N0;
        CB_NOUNROLL for (int j = 0; j < N0; j++) d[j * stride] = x[j];
    }
    CB_NOUNROLL for (int p = 0; p < N; p++) X[p] = tmp[p];
}

// stereo_merge (bands.c:375-424)
CB_DEV_NOINLINE void stereo_merge(int16_t *__restrict__ X, int16_t *__restrict__ Y, int mid, int N) {
    int xp = 0, side = 0;
    CB_NOUNROLL for (int j = 0; j < N; j++) {
        xp = mac16_16(xp, Y[j], X[j]);
        side = mac16_16(side, Y[j], Y[j]);
    }
    xp = mul16_32_q15(mid, xp);
    int mid2 = s16(mid >> 1);
    int El = wsub(wadd(mul16_16(mid2, mid2), side), wmul(2, xp));
    int Er = wadd(wadd(mul16_16(mid2, mid2), side), wmul(2, xp));
    if (Er < 161061 || El < 161061) {   // QCONST32(6e-4f, 28)
        CB_NOUNROLL for (int j = 0; j < N; j++) Y[j] = X[j];
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
    CB_NOUNROLL for (int j = 0; j < N; j++) {
        int l = s16(mul16_16_p15(mid, X[j]));
        int r = Y[j];
        X[j] = (int16_t)pshr32(mul16_16(lgain, s16(l - r)), kl + 1);
        Y[j] = (int16_t)pshr32(mul16_16(rgain, s16(l + r)), kr + 1);
    }
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
    EcDec &ec = ctx.ec;
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
    int16_t *x = X;
    const int nch = Y != nullptr ? 2 : 1;
    CB_NOUNROLL for (int c = 0; c < nch; c++) {
        int sign = 0;
        if (ctx.remaining_bits >= 1 << kBitRes) {
            sign = (int)ctx.ec.bits(1);
            ctx.remaining_bits -= 1 << kBitRes;
        }
        if (!ctx.dry) x[0] = sign ? -16384 : 16384;
        x = Y;
    }
    if (lowband_out && !ctx.dry) lowband_out[0] = (int16_t)(X[0] >> 4);
    return 1;
}

// A leaf of the partition tree, recorded by the walker and executed afterwards.  The symbol walk (range decoder, bit
// accounting) diverges between the lanes of a warp — every lane parses another frame — but none of it depends on spectrum
// values: a leaf's PVQ index is read during the walk, while its pulse vector (cwrsi), normalisation, rotation, or its fold /
// noise fill run in quant_band's leaf loop, which all lanes of the warp enter together (profiles/r1_dec_v5: cwrsi and the
// rotation ran with 5 of 32 lanes when executed from inside the walk).
struct LeafRec {
    int16_t *X;
    const int16_t *lowband;   // fold source (kind kLeafFold)
    unsigned idx;             // PVQ codeword (kind kLeafPvq)
    int16_t N, K, B, gain, fill;
    uint8_t kind, shift;      // shift: position of this leaf's collapse bits in the band's mask
};
enum { kLeafPvq, kLeafZero, kLeafNoise, kLeafFold, kMaxLeaves = 16 };   // <= 2^(maxLM+1) leaves per quant_band

// One leaf of the partition tree (bands.c:989-1036), walk half: pulse count, bit accounting, the PVQ index.
CB_DEV void partition_leaf(BandCtx &ctx, LeafRec *leaves, int &nleaves, int16_t *X, int N, int b, int B, const int16_t *lowband, int LM,
                           int gain, int fill, int shift) {
    int q = bits2pulses(ctx.i, LM, b);
    int curr_bits = pulses2bits(ctx.i, LM, q);
    ctx.remaining_bits -= curr_bits;
    while (ctx.remaining_bits < 0 && q > 0) {
        ctx.remaining_bits += curr_bits;
        q--;
        curr_bits = pulses2bits(ctx.i, LM, q);
        ctx.remaining_bits -= curr_bits;
    }
    unsigned idx = 0;
    int K = 0, kind;
    if (q != 0) {
        K = get_pulses(q);
        idx = ctx.ec.uint_(pvq_v(N, K));
        kind = kLeafPvq;
    } else {
        fill &= (1 << B) - 1;
        kind = !fill ? kLeafZero : lowband == nullptr ? kLeafNoise : kLeafFold;
    }
    if (ctx.dry) return;
    LeafRec &r = leaves[nleaves++];
    r.X = X; r.lowband = lowband; r.idx = idx;
    r.N = (int16_t)N; r.K = (int16_t)K; r.B = (int16_t)B; r.gain = (int16_t)gain; r.fill = (int16_t)fill;
    r.kind = (uint8_t)kind; r.shift = (uint8_t)shift;
}

// Execute half (vq.c:329-346 alg_unquant after the index is known; bands.c:1005-1036 for the fills): returns the leaf's
// collapse mask at its position in the band's mask.
CB_DEV unsigned run_leaf(const LeafRec &r, int spread, unsigned &seed) {
    int16_t *X = r.X;
    const int N = r.N, B = r.B, gain = r.gain;
    unsigned cm;
    if (r.kind == kLeafPvq) {
        cm = alg_unquant_idx(X, N, r.K, spread, B, r.idx, gain);
    } else if (r.kind == kLeafZero) {
        CB_NOUNROLL for (int j = 0; j < N; j++) X[j] = 0;
        cm = 0;
    } else {
        unsigned sd = seed;
        if (r.kind == kLeafNoise) {
            CB_NOUNROLL for (int j = 0; j < N; j++) {
                sd = lcg_rand(sd);
                X[j] = (int16_t)((int)sd >> 20);
            }
            cm = (1u << B) - 1;
        } else {
            int16_t *__restrict__ d = X;
            const int16_t *__restrict__ src = r.lowband;
            CB_NOUNROLL for (int j = 0; j < N; j++) {
                sd = lcg_rand(sd);
                int t = (sd & 0x8000) ? 4 : -4;   // QCONST16(1.0f/256, 10)
                d[j] = (int16_t)(src[j] + t);
            }
            cm = (unsigned)r.fill;
        }
        seed = sd;
        renormalise_vector(SoloTeam{}, X, N, gain);
    }
    return cm << r.shift;
}

// quant_partition (bands.c:864-1040) as an explicit walker.  A frame holds the arguments of one call and,
// once it has split, what the reference keeps in locals across its two recursive calls.  The collapse mask of a split is
// cm(mid) | cm(side) << (B0 >> 1) (bands.c:975-985), so a leaf's bits land at the sum of its side-branch shifts.
struct PartFrame {
    int16_t *X, *lowband;
    int N, b, B, LM, gain, fill, shift;
    // after a split:
    int16_t *Y, *lowband2;
    int B0, mbits, sbits, itheta, rebalance0, gmid, gside, mid_first;
    int stage;   // 0 = entered, 1 = first child returned, 2 = second child returned
};

CB_DEV void quant_partition(BandCtx &ctx, LeafRec *leaves, int &nleaves, int16_t *X, int N, int b, int B, int16_t *lowband, int LM, int gain,
                            int fill) {
    PartFrame st[5];
    int sp = 0;
    st[0].X = X; st[0].lowband = lowband; st[0].N = N; st[0].b = b; st[0].B = B; st[0].LM = LM; st[0].gain = gain;
    st[0].fill = fill; st[0].shift = 0; st[0].stage = 0;
    while (sp >= 0) {
        PartFrame &f = st[sp];
        if (f.stage == 0) {
            const uint8_t *cache = pulse_cache(ctx.i, f.LM);
            if (f.LM != -1 && f.b > cache[cache[0]] + 12 && f.N > 2) {
                SplitCtx s;
                const int n = f.N >> 1;
                const int lm = f.LM - 1;
                int fl = f.fill;
                int bb = f.b;
                const int B0 = f.B;
                if (B0 == 1) fl = (fl & 1) | (fl << 1);
                const int Bn = (B0 + 1) >> 1;
                compute_theta(ctx, s, n, &bb, Bn, B0, lm, 0, &fl);
                int delta = s.delta;
                const int itheta = s.itheta;
                if (B0 > 1 && (itheta & 0x3fff)) {
                    if (itheta > 8192) delta -= delta >> (4 - lm);
                    else delta = imin(0, delta + (n << kBitRes >> (5 - lm)));
                }
                const int mbits = imax(0, imin(bb, (bb - delta) / 2));
                const int sbits = bb - mbits;
                ctx.remaining_bits -= s.qalloc;
                f.Y = f.X + n;
                f.lowband2 = f.lowband ? f.lowband + n : nullptr;
                f.B0 = B0; f.mbits = mbits; f.sbits = sbits; f.itheta = itheta;
                f.rebalance0 = ctx.remaining_bits;
                f.gmid = s16(mul16_16_p15(f.gain, s.imid));
                f.gside = s16(mul16_16_p15(f.gain, s.iside));
                // the frame now describes the halves
                f.N = n; f.LM = lm; f.B = Bn; f.fill = fl;
                f.stage = 1;
                f.mid_first = mbits >= sbits;
                PartFrame &c = st[sp + 1];
                c.N = n; c.B = Bn; c.LM = lm; c.stage = 0;
                if (f.mid_first) { c.X = f.X; c.lowband = f.lowband; c.b = mbits; c.gain = f.gmid; c.fill = fl; c.shift = f.shift; }
                else { c.X = f.Y; c.lowband = f.lowband2; c.b = sbits; c.gain = f.gside; c.fill = fl >> Bn; c.shift = f.shift + (B0 >> 1); }
                sp++;
            } else {
                partition_leaf(ctx, leaves, nleaves, f.X, f.N, f.b, f.B, f.lowband, f.LM, f.gain, f.fill, f.shift);
                sp--;
            }
        } else if (f.stage == 1) {
            PartFrame &c = st[sp + 1];
            c.N = f.N; c.B = f.B; c.LM = f.LM; c.stage = 0;
            if (f.mid_first) {
                int rebalance = f.mbits - (f.rebalance0 - ctx.remaining_bits);
                if (rebalance > 3 << kBitRes && f.itheta != 0) f.sbits += rebalance - (3 << kBitRes);
                c.X = f.Y; c.lowband = f.lowband2; c.b = f.sbits; c.gain = f.gside; c.fill = f.fill >> f.B; c.shift = f.shift + (f.B0 >> 1);
            } else {
                int rebalance = f.sbits - (f.rebalance0 - ctx.remaining_bits);
                if (rebalance > 3 << kBitRes && f.itheta != 16384) f.mbits += rebalance - (3 << kBitRes);
                c.X = f.X; c.lowband = f.lowband; c.b = f.mbits; c.gain = f.gmid; c.fill = f.fill; c.shift = f.shift;
            }
            f.stage = 2;
            sp++;
        } else {
            sp--;
        }
    }
}

// quant_band (bands.c:1044-1170)
CB_DEV unsigned quant_band(BandCtx &ctx, int16_t *X, int N, int b, int B, int16_t *lowband, int LM, int16_t *lowband_out, int gain,
                           int16_t *lowband_scratch, int fill) {
    int N0 = N, N_B = N, N_B0, B0 = B;
    int time_divide = 0, recombine = 0;
    int tf_change = ctx.tf_change;
    const int longBlocks = B0 == 1;
    unsigned cm = 0;
    N_B = (int)udiv((unsigned)N_B, (unsigned)B);
    if (N == 1) return quant_band_n1(ctx, X, nullptr, lowband_out);
    if (tf_change > 0) recombine = tf_change;
    if (ctx.dry) lowband = nullptr;
    if (lowband_scratch && lowband && (recombine || ((N_B & 1) == 0 && tf_change < 0) || B0 > 1)) {
        {
            int16_t *__restrict__ d = lowband_scratch;
            const int16_t *__restrict__ src = lowband;
            CB_NOUNROLL for (int j = 0; j < N; j++) d[j] = src[j];
        }
        lowband = lowband_scratch;
    }
    CB_NOUNROLL for (int k = 0; k < recombine; k++) {
        if (lowband) haar1(lowband, N >> k, 1 << k);
        fill = kBitInterleave[fill & 0xF] | kBitInterleave[fill >> 4] << 2;
    }
    B >>= recombine;
    N_B <<= recombine;
    while ((N_B & 1) == 0 && tf_change < 0) {
        if (lowband) haar1(lowband, N_B, B);
        fill |= fill << B;
        B <<= 1;
        N_B >>= 1;
        time_divide++;
        tf_change++;
    }
    B0 = B;
    N_B0 = N_B;
    if (B0 > 1 && lowband) deinterleave_hadamard(lowband, ctx.tmp, N_B >> recombine, B0 << recombine, longBlocks);

    LeafRec leaves[kMaxLeaves];
    int nleaves = 0;
    quant_partition(ctx, leaves, nleaves, X, N, b, B, lowband, LM, gain, fill);
    if (ctx.dry) return 0;
    // the leaf loop: every lane of the warp that decodes this band arrives here together
    {
        unsigned sd = ctx.seed;
        CB_NOUNROLL for (int l = 0; l < nleaves; l++) cm |= run_leaf(leaves[l], ctx.spread, sd);
        ctx.seed = sd;
    }

    // resynthesis (decoder): undo the reorganisation
    if (B0 > 1) interleave_hadamard(X, ctx.tmp, N_B >> recombine, B0 << recombine, longBlocks);
    N_B = N_B0;
    B = B0;
    CB_NOUNROLL for (int k = 0; k < time_divide; k++) {
        B >>= 1;
        N_B <<= 1;
        cm |= cm >> B;
        haar1(X, N_B, B);
    }
    CB_NOUNROLL for (int k = 0; k < recombine; k++) {
        cm = kBitDeinterleave[cm];
        haar1(X, N0 >> k, 1 << k);
    }
    B <<= recombine;
    if (lowband_out) {
        const int n = s16(celt_sqrt(shl32(N0, 22)));
        int16_t *__restrict__ d = lowband_out;
        const int16_t *__restrict__ src = X;
        CB_NOUNROLL for (int j = 0; j < N0; j++) d[j] = (int16_t)mul16_16_q15(n, src[j]);
    }
    cm &= (1u << B) - 1;
    return cm;
}

// Copy a finished band from the working copy (16-byte aligned) to the IR spectrum: band offsets are multiples of M = 1 << LM
// int16 and widths multiples of M, the frame's spectrum is 16-byte aligned.
CB_DEV void store_band(int16_t *g, const int16_t *x, int N, int LM) {
    if (LM == 3) {
        CB_NOUNROLL for (int j = 0; j < N; j += 8) *reinterpret_cast<int4 *>(g + j) = *reinterpret_cast<const int4 *>(x + j);
    } else if (LM == 2) {
        CB_NOUNROLL for (int j = 0; j < N; j += 4) *reinterpret_cast<int2 *>(g + j) = *reinterpret_cast<const int2 *>(x + j);
    } else if (LM == 1) {
        CB_NOUNROLL for (int j = 0; j < N; j += 2) *reinterpret_cast<int *>(g + j) = *reinterpret_cast<const int *>(x + j);
    } else {
        CB_NOUNROLL for (int j = 0; j < N; j++) g[j] = x[j];
    }
}

// quant_all_bands, decoder (bands.c:1337-1502) with quant_band_stereo (bands.c:1176-1335) folded into the
// per-band pass loop.  X_: C*N int16 (channel-major), norm: C*(M*eBands[20]) int16.
CB_DEV void quant_all_bands_dec(int start, int end, int16_t *X_, int16_t *Y_, uint8_t *collapse_masks, const int *pulses,
                                int shortBlocks, int spread, int dual_stereo, int intensity, const int *tf_res, int total_bits,
                                int balance, EcDec &ec_io, int LM, int codedBands, unsigned *seed, int16_t *norm, int16_t *tmp,
                                int16_t *xw, int16_t *lbs, bool dry) {
    const int M = 1 << LM;
    const int B = shortBlocks ? M : 1;
    const int C = Y_ != nullptr ? 2 : 1;
    const int norm_offset = M * kEBands[start];
    int16_t *norm2 = norm + M * kEBands[kNbEBands - 1] - norm_offset;
    int16_t *lowband_scratch = lbs;
    int lowband_offset = 0;
    int update_lowband = 1;
    BandCtx ctx;
    ctx.ec = ec_io;
    ctx.tmp = tmp;
    ctx.intensity = intensity; ctx.spread = spread; ctx.seed = *seed;
    ctx.dry = dry;
    CB_NOUNROLL for (int i = start; i < end; i++) {
        ctx.i = i;
        const int last = (i == end - 1);
        // The band is decoded in the thread's working copy and stored to the IR once, with the widest stores its alignment
        // allows: byte-granular read-modify-write of the IR made stage A's L1->L2 write path its busiest unit (profiles/).
        int16_t *X = xw;
        int16_t *Y = Y_ != nullptr ? xw + 176 : nullptr;
        const int N = M * kEBands[i + 1] - M * kEBands[i];
        const int tell = (int)ctx.ec.tell_frac();
        if (i != start) balance -= tell;
        const int remaining_bits = total_bits - tell - 1;
        ctx.remaining_bits = remaining_bits;
        int b;
        if (i <= codedBands - 1) {
            int curr_balance = sudiv(balance, imin(3, codedBands - i));
            b = imax(0, imin(16383, imin(remaining_bits + 1, pulses[i] + curr_balance)));
        } else {
            b = 0;
        }
        if (M * kEBands[i] - N >= M * kEBands[start] && (update_lowband || lowband_offset == 0)) lowband_offset = i;
        const int tf_change = tf_res[i];
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
            const int n = dry ? 0 : M * kEBands[i] - norm_offset;
            CB_NOUNROLL for (int j = 0; j < n; j++) norm[j] = (int16_t)((norm[j] + norm2[j]) >> 1);
        }
        int16_t *lb = effective_lowband != -1 ? norm + effective_lowband : nullptr;
        int16_t *lb_out = last ? nullptr : norm + M * kEBands[i] - norm_offset;

        // ---- plan the quant_band passes of this band ----
        enum { kMono, kDual, kStereo, kStereoN2 };
        int mode, npass;
        SplitCtx s;
        int mbits = 0, sbits = 0, rebalance0 = 0, sfill = 0, orig_fill = 0, sign = 1;
        int16_t *x2 = nullptr, *y2 = nullptr;
        if (Y != nullptr && !dual_stereo && N == 1) {
            x_cm = quant_band_n1(ctx, X, Y, lb_out);
            y_cm = x_cm;
            npass = 0;
            mode = kMono;
        } else if (dual_stereo) {
            mode = kDual; npass = 2;
        } else if (Y != nullptr) {
            // quant_band_stereo, first half (bands.c:1208-1246 / 1278-1284)
            orig_fill = sfill = (int)(x_cm | y_cm);
            int bs = b;   // quant_band_stereo works on its own copy of b (update_lowband below needs the original)
            compute_theta(ctx, s, N, &bs, B, B, LM, 1, &sfill);
            if (N == 2) {
                mode = kStereoN2; npass = 1;
                mbits = bs; sbits = 0;
                if (s.itheta != 0 && s.itheta != 16384) sbits = 1 << kBitRes;
                mbits -= sbits;
                const int c = s.itheta > 8192;
                ctx.remaining_bits -= s.qalloc + sbits;
                x2 = c ? Y : X;
                y2 = c ? X : Y;
                int sg = 0;
                if (sbits) sg = (int)ctx.ec.bits(1);
                sign = 1 - 2 * sg;
            } else {
                mode = kStereo; npass = 2;
                mbits = imax(0, imin(bs, (bs - s.delta) / 2));
                sbits = bs - mbits;
                ctx.remaining_bits -= s.qalloc;
                rebalance0 = ctx.remaining_bits;
            }
        } else {
            mode = kMono; npass = 1;
        }
        unsigned cm_acc = 0;
        CB_NOUNROLL for (int pass = 0; pass < npass; pass++) {
            int16_t *px, *plb, *plb_out, *pscratch;
            int pb, pgain, pfill;
            if (mode == kMono) {
                px = X; pb = b; plb = lb; plb_out = lb_out; pgain = 32767; pscratch = lowband_scratch; pfill = (int)(x_cm | y_cm);
            } else if (mode == kDual) {
                if (pass == 0) { px = X; plb = lb; plb_out = lb_out; pfill = (int)x_cm; }
                else {
                    px = Y;
                    plb = effective_lowband != -1 ? norm2 + effective_lowband : nullptr;
                    plb_out = last ? nullptr : norm2 + M * kEBands[i] - norm_offset;
                    pfill = (int)y_cm;
                }
                pb = b / 2; pgain = 32767; pscratch = lowband_scratch;
            } else if (mode == kStereoN2) {
                px = x2; pb = mbits; plb = lb; plb_out = lb_out; pgain = 32767; pscratch = lowband_scratch; pfill = orig_fill;
            } else {
                // mid first when mbits >= sbits, else side first (bands.c:1286-1317); rebalance before the second
                const bool mid_first = mbits >= sbits;
                if (pass == 1) {
                    if (mid_first) {
                        int rebalance = mbits - (rebalance0 - ctx.remaining_bits);
                        if (rebalance > 3 << kBitRes && s.itheta != 0) sbits += rebalance - (3 << kBitRes);
                    } else {
                        int rebalance = sbits - (rebalance0 - ctx.remaining_bits);
                        if (rebalance > 3 << kBitRes && s.itheta != 16384) mbits += rebalance - (3 << kBitRes);
                    }
                }
                const bool do_mid = (pass == 0) == mid_first;
                if (do_mid) { px = X; pb = mbits; plb = lb; plb_out = lb_out; pgain = 32767; pscratch = lowband_scratch; pfill = sfill; }
                else { px = Y; pb = sbits; plb = nullptr; plb_out = nullptr; pgain = s.iside; pscratch = nullptr; pfill = sfill >> B; }
            }
            unsigned cm = quant_band(ctx, px, N, pb, B, plb, LM, plb_out, pgain, pscratch, pfill);
            if (mode == kDual) { if (pass == 0) x_cm = cm; else y_cm = cm; }
            else cm_acc |= cm;
        }
        if (dry) {
            // nothing to recombine
        } else if (mode == kStereoN2) {
            // bands.c:1250-1268
            y2[0] = (int16_t)(-sign * x2[1]);
            y2[1] = (int16_t)(sign * x2[0]);
            X[0] = (int16_t)mul16_16_q15(s.imid, X[0]);
            X[1] = (int16_t)mul16_16_q15(s.imid, X[1]);
            Y[0] = (int16_t)mul16_16_q15(s.iside, Y[0]);
            Y[1] = (int16_t)mul16_16_q15(s.iside, Y[1]);
            int t = X[0];
            X[0] = (int16_t)(t - Y[0]);
            Y[0] = (int16_t)(t + Y[0]);
            t = X[1];
            X[1] = (int16_t)(t - Y[1]);
            Y[1] = (int16_t)(t + Y[1]);
        } else if (mode == kStereo) {
            stereo_merge(X, Y, s.imid, N);
        }
        if (!dry && (mode == kStereo || mode == kStereoN2) && s.inv)
            CB_NOUNROLL for (int j = 0; j < N; j++) Y[j] = (int16_t)(-Y[j]);
        if (!dry) {
            store_band(X_ + M * kEBands[i], X, N, LM);
            if (Y != nullptr) store_band(Y_ + M * kEBands[i], Y, N, LM);
        }
        if (npass > 0 && mode != kDual) x_cm = y_cm = cm_acc;
        collapse_masks[i * C + 0] = (uint8_t)x_cm;
        collapse_masks[i * C + C - 1] = (uint8_t)y_cm;
        balance += pulses[i] + tell;
        update_lowband = b > (N << kBitRes);
    }
    *seed = ctx.seed;
    ec_io = ctx.ec;
}

// anti_collapse (bands.c:241-335), stage B: X_ is the frame's spectrum in the IR (global memory).
template <class TM>
CB_DEV void anti_collapse(TM tm, int16_t *X_, const uint8_t *collapse_masks, int LM, int C, int size, int start, int end,
                          const int16_t *logE, const int16_t *prev1logE, const int16_t *prev2logE, const int16_t *pulses, unsigned seed) {
    CB_NOUNROLL for (int i = start; i < end; i++) {
        int N0 = band_width(i);
        int depth = (int)udiv((unsigned)(1 + pulses[i]), (unsigned)N0) >> LM;
        int thresh32 = celt_exp2(s16(-shl16(depth, 10 - kBitRes))) >> 1;
        int thresh = s16(mul16_32_q15(16384, imin(32767, thresh32)));
        int t = N0 << LM;
        int shift = celt_ilog2(t) >> 1;
        t = shl32(t, (7 - shift) << 1);
        int sqrt_1 = celt_rsqrt_norm(t);
        CB_NOUNROLL for (int c = 0; c < C; c++) {
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
            const unsigned mask = collapse_masks[i * C + c];
            CB_NOUNROLL for (int k = 0; k < 1 << LM; k++) {
                if (!(mask & (1u << k))) {
                    // sample j takes the generator's (j+1)-th step: an affine jump per lane instead of N0 steps on every lane
                    CB_TEAM_FOR(j, N0, tm) {
                        const unsigned sj = lcg_jump(seed, j + 1);
                        X[(j << LM) + k] = (int16_t)((sj & 0x8000) ? r : -r);
                    }
                    seed = lcg_jump(seed, N0);
                    renormalize = 1;
                }
            }
            tm.sync();
            if (renormalize) renormalise_vector(tm, X, N0 << LM, 32767);
        }
    }
}

}  // namespace cb
