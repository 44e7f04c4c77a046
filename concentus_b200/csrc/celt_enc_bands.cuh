// celt_enc_bands.cuh — encoder side of the band loop: band energies, normalisation, spreading decision, stereo angle,
// PVQ search and indexing, and quant_all_bands with encode = 1 (no resynthesis: this build has no RESYNTH).
//
// Restates opus-fix/celt/bands.c:48-61 (hysteresis_decision), :97-164 (compute_band_energies, normalise_bands), :337-373
// (intensity_stereo, stereo_split), :428-519 (spreading_decision), the encode branches of :645-1502 (compute_theta,
// quant_band_n1, quant_partition, quant_band, quant_band_stereo, quant_all_bands), celt/vq.c:70-113 (exp_rotation, dir=+1),
// :161-325 (alg_quant), :376-408 (stereo_itheta) and celt/cwrs.c:440-461 (icwrs, encode_pulses).
// Without resynthesis the encoder never folds: lowband, norm, fill and the collapse masks do not influence the bitstream
// (bands.c:1413-1414 only advances lowband_offset under `resynth`), so they are not carried here.  Scalar code.
#pragma once
#include "celt_bands.cuh"
#include "celt_pitch.cuh"

namespace cb {

// bands.c:48-61
CB_DEV int hysteresis_decision(int val, const int16_t *thresholds, const int16_t *hysteresis, int N, int prev) {
    int i;
    CB_NOUNROLL for (i = 0; i < N; i++)
        if (val < thresholds[i]) break;
    if (i > prev && val < thresholds[prev] + hysteresis[prev]) i = prev;
    if (i < prev && val > thresholds[prev - 1] - hysteresis[prev - 1]) i = prev;
    return i;
}

// bands.c:97-143.  freq: C*N int32; bandE: [c*21+i]
CB_DEV_NOINLINE void compute_band_energies(const int *freq, int *bandE, int end, int C, int LM) {
    const int N = kShortMdct << LM;
    CB_NOUNROLL for (int c = 0; c < C; c++) {
        CB_NOUNROLL for (int i = 0; i < end; i++) {
            const int lo = kEBands[i] << LM, hi = kEBands[i + 1] << LM;
            const int *x = freq + c * N;
            int maxval = maxabs32(x + lo, hi - lo);
            if (maxval > 0) {
                int shift = celt_ilog2(maxval) - 14 + (((kLogN[i] >> kBitRes) + LM + 1) >> 1);
                int sum = 0;
                if (shift > 0) {
                    CB_NOUNROLL for (int j = lo; j < hi; j++) { int v = s16(x[j] >> shift); sum = mac16_16(sum, v, v); }
                } else {
                    CB_NOUNROLL for (int j = lo; j < hi; j++) { int v = s16(shl32(x[j], -shift)); sum = mac16_16(sum, v, v); }
                }
                bandE[i + c * kNbEBands] = wadd(1, vshr32(celt_sqrt(sum), -shift));
            } else {
                bandE[i + c * kNbEBands] = 1;
            }
        }
    }
}

// bands.c:146-164
CB_DEV_NOINLINE void normalise_bands(const int *freq, int16_t *X, const int *bandE, int end, int C, int M) {
    const int N = M * kShortMdct;
    CB_NOUNROLL for (int c = 0; c < C; c++) {
        CB_NOUNROLL for (int i = 0; i < end; i++) {
            int shift = celt_zlog2(bandE[i + c * kNbEBands]) - 13;
            int E = s16(vshr32(bandE[i + c * kNbEBands], shift));
            int g = s16(celt_rcp(shl32(E, 3)));
            CB_NOUNROLL for (int j = M * kEBands[i]; j < M * kEBands[i + 1]; j++)
                X[j + c * N] = (int16_t)mul16_16_q15(s16(vshr32(freq[j + c * N], shift - 1)), g);
        }
    }
}

// bands.c:428-519
CB_DEV_NOINLINE int spreading_decision(const int16_t *X, int *average, int last_decision, int *hf_average, int *tapset_decision,
                                       int update_hf, int end, int C, int M) {
    int sum = 0, nbBands = 0, hf_sum = 0;
    const int N0 = M * kShortMdct;
    if (M * (kEBands[end] - kEBands[end - 1]) <= 8) return kSpreadNone;
    CB_NOUNROLL for (int c = 0; c < C; c++) {
        CB_NOUNROLL for (int i = 0; i < end; i++) {
            const int16_t *x = X + M * kEBands[i] + c * N0;
            const int N = M * (kEBands[i + 1] - kEBands[i]);
            if (N <= 8) continue;
            int t0 = 0, t1 = 0, t2 = 0;
            CB_NOUNROLL for (int j = 0; j < N; j++) {
                int x2N = mul16_16(mul16_16_q15(x[j], x[j]), N);
                if (x2N < 2048) t0++;
                if (x2N < 512) t1++;
                if (x2N < 128) t2++;
            }
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

// ---- team-parallel vector kernels of the band loop ----------------------------------------------------------------------
// The band loop (quant_all_bands_enc below) is executed by EVERY lane of the team with identical (uniform) scalars — range
// coder included — so no broadcast is needed; only the vector work on X is split over the lanes.  All sums are wrapping
// 32-bit (order-free), so the split cannot change a bit.  Every function leaves the team synchronised.

// vq.c:376-408
template <class TM>
CB_DEV_NOINLINE int stereo_itheta(TM tm, const int16_t *X, const int16_t *Y, int stereo, int N) {
    int em = 0, es = 0;
    if (stereo) {
        CB_TEAM_FOR(i, N, tm) {
            int m = s16((X[i] >> 1) + (Y[i] >> 1));
            int s = s16((X[i] >> 1) - (Y[i] >> 1));
            em = mac16_16(em, m, m);
            es = mac16_16(es, s, s);
        }
    } else {
        CB_TEAM_FOR(i, N, tm) {
            em = mac16_16(em, X[i], X[i]);
            es = mac16_16(es, Y[i], Y[i]);
        }
    }
    const int Emid = wadd(1, tm.sum(em)), Eside = wadd(1, tm.sum(es));
    int mid = s16(celt_sqrt(Emid));
    int side = s16(celt_sqrt(Eside));
    return mul16_16_q15(20861, celt_atan2p(side, mid));   // QCONST16(0.63662f,15)
}

// bands.c:337-360
template <class TM>
CB_DEV_NOINLINE void intensity_stereo(TM tm, int16_t *X, const int16_t *Y, const int *bandE, int i, int N) {
    int shift = celt_zlog2(imax(bandE[i], bandE[i + kNbEBands])) - 13;
    int left = s16(vshr32(bandE[i], shift));
    int right = s16(vshr32(bandE[i + kNbEBands], shift));
    int norm = s16(1 + celt_sqrt(wadd(1, wadd(mul16_16(left, left), mul16_16(right, right)))));
    int a1 = s16(shl32(left, 14) / norm);
    int a2 = s16(shl32(right, 14) / norm);
    CB_TEAM_FOR(j, N, tm) X[j] = (int16_t)(mac16_16(mul16_16(a1, X[j]), a2, Y[j]) >> 14);
    tm.sync();
}
// bands.c:362-373
template <class TM>
CB_DEV_NOINLINE void stereo_split(TM tm, int16_t *X, int16_t *Y, int N) {
    CB_TEAM_FOR(j, N, tm) {
        int l = mul16_16(23170, X[j]);
        int r = mul16_16(23170, Y[j]);
        X[j] = (int16_t)(wadd(l, r) >> 15);
        Y[j] = (int16_t)(wsub(r, l) >> 15);
    }
    tm.sync();
}
template <class TM>
CB_DEV_NOINLINE void negate_vector(TM tm, int16_t *Y, int N) {
    CB_TEAM_FOR(j, N, tm) Y[j] = (int16_t)(-Y[j]);
    tm.sync();
}

// haar1 (bands.c:581-594): N0/2 * stride independent butterflies
template <class TM>
CB_DEV_NOINLINE void haar1_team(TM tm, int16_t *X, int N0, int stride) {
    N0 >>= 1;
    CB_TEAM_FOR(w, N0 * stride, tm) {
        const int j = w / stride, i = w - j * stride;
        const int a = stride * 2 * j + i, b = stride * (2 * j + 1) + i;
        const int t1 = mul16_16(23170, X[a]);
        const int t2 = mul16_16(23170, X[b]);
        X[a] = (int16_t)pshr32(wadd(t1, t2), 15);
        X[b] = (int16_t)pshr32(wsub(t1, t2), 15);
    }
    tm.sync();
}
// deinterleave_hadamard (bands.c:532-556) through the team scratch
template <class TM>
CB_DEV_NOINLINE void deinterleave_hadamard_team(TM tm, int16_t *X, int16_t *tmp, int N0, int stride, int hadamard) {
    const int N = N0 * stride;
    const uint8_t *ordery = kOrdery + stride - 2;
    CB_TEAM_FOR(w, N, tm) {
        const int j = w / stride, i = w - j * stride;
        const int row = hadamard ? ordery[i] : i;
        tmp[row * N0 + j] = X[w];
    }
    tm.sync();
    CB_TEAM_FOR(p, N, tm) X[p] = tmp[p];
    tm.sync();
}

#if defined(CB_WALK_DEBUG)
static int g_dbg_tell[3][32]; static unsigned g_dbg_rng[3][32]; static int g_dbg_which = 1; static int g_dbg_log = 0;
#endif
// One residue class of exp_rotation1 (vq.c:43-67): the pairs (i, i+stride) with i = r (mod stride) form an independent chain.
// Each sweep carries the element it shares with the next pair in a register: one load and one store per step.
CB_DEV_TINY void exp_rotation1_chain(int16_t *X, int len, int stride, int c, int s, int r) {
    const int ms = s16(-s);
    if (r < len - stride) {
        int i = r;
        int x1 = X[i];
        CB_NOUNROLL for (; i < len - stride; i += stride) {
            const int x2 = X[i + stride];
            X[i] = (int16_t)pshr32(mac16_16(mul16_16(c, x1), ms, x2), 15);
            x1 = s16(pshr32(mac16_16(mul16_16(c, x2), s, x1), 15));
        }
        X[i] = (int16_t)x1;
    }
    const int top = len - 2 * stride - 1;
    if (top >= r) {
        int i = top - ((top - r) % stride);
        int x2 = X[i + stride];
        CB_NOUNROLL for (; i >= 0; i -= stride) {
            const int x1 = X[i];
            X[i + stride] = (int16_t)pshr32(mac16_16(mul16_16(c, x2), s, x1), 15);
            x2 = s16(pshr32(mac16_16(mul16_16(c, x1), ms, x2), 15));
        }
        X[i + stride] = (int16_t)x2;
    }
}

// The rotation's parameters (vq.c:80-93): c, s from (len, K, spread) through celt_div and two celt_cos_norm, stride2 from (len, stride).
CB_DEV void rotation_params(int len, int K, int spread, int &c, int &s) {
    const int factor = spread == 1 ? 15 : spread == 2 ? 10 : 5;
    const int gain = s16(celt_div(mul16_16(32767, len), len + factor * K));
    const int theta = mul16_16_q15(gain, gain) >> 1;
    c = celt_cos_norm(theta);
    s = celt_cos_norm(s16(32767 - theta));
}
CB_DEV int rotation_stride2(int len, int stride) {
    int stride2 = 0;
    if (len >= 8 * stride) {
        stride2 = 1;
        while ((stride2 * stride2 + stride2) * stride + (stride >> 2) < len) stride2++;
    }
    return stride2;
}
// CB_ROT_LUT (the encoder pipeline): ~170 uniform scalar instructions per leaf for those parameters — a ninth of the band walk —
// become two loads from tables the pipeline fills at start-up WITH the functions above (opus_enc_pipe.cu, rot_lut_kernel).
enum { kRotLen = 177, kRotK = 88 };   // leaves are <= 176 values wide; the rotation runs only when 2 K < len
#if defined(CB_ROT_LUT) && defined(__CUDACC__)
static __device__ uint32_t g_rot_cs[3 * kRotLen * kRotK];   // (c & 0xffff) | s << 16, by (spread - 1, len, K)
static __device__ uint8_t g_rot_s2[kRotLen * 8];            // stride2 by (len, log2 stride): the time-divided short blocks reach stride 16+
static __device__ uint32_t g_inv16[32];                     // ceil(65536 / d): x / d == x * inv >> 16 for x < 256, d < 32
#endif

// exp_rotation, encoder direction (vq.c:70-113 with dir = +1): per block the stride-1 sweep (one chain), then the stride2 sweep
// (stride2 chains); blocks are independent, so the team runs `stride` resp. `stride*stride2` chains at a time.
template <class TM>
CB_DEV_NOINLINE void exp_rotation_enc(TM tm, int16_t *X, int len, int stride, int K, int spread) {
    if (2 * K >= len || spread == kSpreadNone) return;
    int c, s, stride2;
    const int lstride = celt_ilog2(stride);                  // stride (= B) is 1, 2, 4 or 8
#if defined(CB_ROT_LUT) && defined(__CUDACC__)
    {
        const uint32_t w = g_rot_cs[((spread - 1) * kRotLen + len) * kRotK + K];
        c = (int)(w & 0xffffu);
        s = (int)(w >> 16);
        stride2 = g_rot_s2[len * 8 + lstride];
    }
#else
    rotation_params(len, K, spread, c, s);
    stride2 = rotation_stride2(len, stride);
#endif
    len >>= lstride;
    CB_TEAM_FOR(b, stride, tm) exp_rotation1_chain(X + b * len, len, 1, c, s16(-s), 0);
    tm.sync();
    if (stride2) {
#if defined(CB_ROT_LUT) && defined(__CUDACC__)
        const uint32_t inv = g_inv16[stride2];
#endif
        CB_TEAM_FOR(w, stride * stride2, tm) {
#if defined(CB_ROT_LUT) && defined(__CUDACC__)
            const int b = (int)(((uint32_t)w * inv) >> 16), r = w - b * stride2;
#else
            const int b = w / stride2, r = w - b * stride2;
#endif
            exp_rotation1_chain(X + b * len, len, stride2, s, s16(-c), r);
        }
        tm.sync();
    }
}

// icwrs (cwrs.c:440-456) as a sum over positions: with S_j = sum_{m>=j} |y_m| (so S_0 = K),
//   index = [y_{n-1} < 0] + sum_{j=0}^{n-2} ( U(n-j, S_{j+1}) + [y_j < 0] * U(n-j, S_j + 1) )      (mod 2^32)
template <class TM>
CB_DEV_NOINLINE unsigned pvq_encode_index(TM tm, int n, int K, const int16_t *y) {
    const int per = (n + TM::W - 1) / TM::W;
    const int first = tm.lane() * per;
    int local = 0;
    CB_NOUNROLL for (int j = first; j < first + per && j < n; j++) local += iabs((int)y[j]);
    int S = K - tm.exscan(local);   // S_first
    unsigned acc = 0;
    CB_NOUNROLL for (int j = first; j < first + per && j < n; j++) {
        const int a = iabs((int)y[j]);
        const int Snext = S - a;   // S_{j+1}
        if (j < n - 1) {
            acc += pvq_u(n - j, Snext);
            if (y[j] < 0) acc += pvq_u(n - j, S + 1);
        } else {
            acc += y[j] < 0;
        }
        S = Snext;
    }
    return (unsigned)tm.sum((int)acc);
}

// The same sum when position j lives in lane j (n <= team width): iy = the lane's signed pulse count (0 on lanes >= n)
template <class TM>
CB_DEV unsigned pvq_index_lane(TM tm, int n, int K, int iy) {
    const int j = tm.lane();
    const int a = iabs(iy);
    const int S = K - tm.exscan(a);   // S_j
    unsigned acc = 0;
    if (j < n - 1) {
        acc = pvq_u(n - j, S - a);
        if (iy < 0) acc += pvq_u(n - j, S + 1);
    } else if (j == n - 1) {
        acc = iy < 0;
    }
    return (unsigned)tm.sum((int)acc);
}

// Scratch of the PVQ search: y (doubled pulses), iy (pulses), sign, per band (N <= 176)
struct PvqScratch {
    int16_t y[176], iy[176];
    int8_t sign[176];
};

// best candidate of the greedy search: ratio num/den, ties to the lower index
struct PvqBest { int num, den, id; };
CB_DEV bool pvq_better(const PvqBest &a, const PvqBest &b) {
    const int l = mul16_16(b.den, a.num), r = mul16_16(a.den, b.num);
    return l > r || (l == r && a.id < b.id);
}

// Position the sequential scan of vq.c:265-294 selects, given every lane's own best (sentinel num < 0, den = 0 on lanes without
// a position).  SoloTeam: the lane's best is the answer.
CB_DEV int pvq_pick(SoloTeam, const PvqBest &best, int) { return best.id; }
#if defined(__CUDACC__)
// WarpTeam: a float estimate of num/den picks a provisional winner with one redux.max; its (num, den) are then compared
// EXACTLY (the reference's cross-multiplication) on every lane.  No lane strictly better: the winner is the lowest id among the
// exact ties (one redux.min).  Otherwise (the estimate mis-ordered two near-equal ratios) the exact shuffle tree decides.
// the exact shuffle tree (rare: the float estimate mis-ordered two near-equal ratios); a real call, out of the hot loop's code
static __device__ __noinline__ int pvq_pick_tree(WarpTeam tm, PvqBest best, int levels) {
    CB_NOUNROLL for (int lv = 0; lv < levels; lv++) {
        PvqBest o;
        o.num = tm.shfl_xor(best.num, 1 << lv);
        o.den = tm.shfl_xor(best.den, 1 << lv);
        o.id = tm.shfl_xor(best.id, 1 << lv);
        if (pvq_better(o, best)) best = o;
    }
    return tm.bcast(best.id, 0);
}
CB_DEV int pvq_pick(WarpTeam tm, PvqBest best, int levels) {
    const unsigned full = 0xffffffffu;
    const unsigned key = best.den > 0 ? __float_as_uint(__fdividef((float)best.num, (float)best.den)) + 1u : 0u;
    const unsigned kmax = __reduce_max_sync(full, key);
    const int m = __ffs((int)__ballot_sync(full, key == kmax)) - 1;
    const int num_m = __shfl_sync(full, best.num, m), den_m = __shfl_sync(full, best.den, m);
    const int l = mul16_16(den_m, best.num), r = mul16_16(best.den, num_m);
    if (__ballot_sync(full, l > r) == 0u)
        return (int)__reduce_min_sync(full, l == r && best.den > 0 ? (unsigned)best.id : 0x7fffffffu);
    return pvq_pick_tree(tm, best, levels);
}
CB_DEV int pvq_pick(FreeWarpTeam tm, PvqBest best, int levels) { return pvq_pick(static_cast<WarpTeam>(tm), best, levels); }
CB_DEV int pvq_pick(SyncWarpTeam tm, PvqBest best, int levels) { return pvq_pick(static_cast<WarpTeam>(tm), best, levels); }
#endif

// alg_quant (vq.c:161-325), no resynthesis.  The greedy search is the encoder's hottest loop (profiles/): every lane scans its
// share of the N positions in increasing order with the reference's strict '>' test (first maximum wins), then a shuffle tree
// takes the best of the lanes with ties going to the lower index — the same element the sequential scan selects, because with
// Ryy > 0 and Rxy >= 0 the cross-multiplied comparison is a strict weak order on the ratios.
// alg_quant for N <= team width: position j lives in lane j's registers (|X|, y, iy, sign), nothing touches shared memory
// until the pulse vector is handed to the indexer.  Same arithmetic and tie-breaking as the general version below.
// (the search only: the pulse vector is left in ps.iy)
// Returns the lane's signed pulse count.
template <class TM>
CB_DEV_NOINLINE int alg_quant_small_core(TM tm, int16_t *X, int N, int K, PvqScratch &ps) {
    const int j = tm.lane();
    const bool active = j < N;
    int xj = active ? (int)X[j] : 0;
    const int sgn = xj > 0 ? 1 : -1;
    if (xj <= 0) xj = s16(-xj);
    int yj = 0, iyj = 0;
    int xy = 0, yy = 0;
    int pulsesLeft = K;
    if (K > (N >> 1)) {
        int sum = tm.sum(xj);
        if (sum <= K) {
            xj = j == 0 ? 16384 : 0;
            sum = 16384;
        }
        const int rcp = s16(mul16_32_q16(K - 1, celt_rcp(sum)));
        const int v = mul16_16_q15(xj, rcp);
        iyj = v;
        const int yv = s16(v);
        yy = s16(tm.sum(mul16_16(yv, yv)));
        xy = tm.sum(mul16_16(xj, yv));
        yj = s16(yv * 2);
        pulsesLeft -= tm.sum(v);
    }
    if (pulsesLeft > N + 3) {
        const int tmp = s16(pulsesLeft);
        const int y0 = tm.bcast(yj, 0);
        yy = s16(mac16_16(yy, tmp, tmp));
        yy = s16(mac16_16(yy, tmp, y0));
        if (j == 0) iyj += pulsesLeft;
        pulsesLeft = 0;
    }
    int levels = 0;
    while ((1 << levels) < N) levels++;
    CB_NOUNROLL for (int i = 0; i < pulsesLeft; i++) {
        const int rshift = 1 + celt_ilog2(K - pulsesLeft + i + 1);
        yy = s16(wadd(yy, 1));
        PvqBest best{-32767, 0, j};
        if (active) {
            int Rxy = s16(wadd(xy, xj) >> rshift);
            best.den = s16(yy + yj);
            best.num = s16(mul16_16_q15(Rxy, Rxy));
        }
        const int best_id = pvq_pick(tm, best, levels);
        xy = wadd(xy, tm.bcast(xj, best_id));
        yy = s16(yy + tm.bcast(yj, best_id));
        if (j == best_id) {
            yj = s16(yj + 2);
            iyj++;
        }
    }
    const int iys = active ? (sgn < 0 ? -iyj : iyj) : 0;
    if (active) {
        X[j] = (int16_t)mul16_16(sgn, xj);
        ps.iy[j] = (int16_t)iys;
    }
    tm.sync();
    return iys;
}

// the general search (any N), pulse vector left in ps.iy
template <class TM>
CB_DEV_NOINLINE void alg_quant_core(TM tm, int16_t *X, int N, int K, PvqScratch &ps) {
    int16_t *y = ps.y, *iy = ps.iy;
    int8_t *signx = ps.sign;
    CB_TEAM_FOR(j, N, tm) {
        const int x = X[j];
        if (x > 0) signx[j] = 1;
        else { signx[j] = -1; X[j] = (int16_t)(-x); }
        iy[j] = 0;
        y[j] = 0;
    }
    tm.sync();
    int xy = 0;
    int yy = 0;   // opus_val16 in the reference: truncated after every update
    int pulsesLeft = K;
    if (K > (N >> 1)) {
        int part = 0;
        CB_TEAM_FOR(j, N, tm) part = wadd(part, X[j]);
        int sum = tm.sum(part);
        if (sum <= K) {
            tm.sync();
            CB_TEAM_FOR(j, N, tm) X[j] = j == 0 ? 16384 : 0;
            tm.sync();
            sum = 16384;
        }
        const int rcp = s16(mul16_32_q16(K - 1, celt_rcp(sum)));
        int pyy = 0, pxy = 0, pk = 0;
        CB_TEAM_FOR(j, N, tm) {
            const int v = mul16_16_q15(X[j], rcp);
            iy[j] = (int16_t)v;
            const int yv = s16(v);
            pyy = mac16_16(pyy, yv, yv);
            pxy = mac16_16(pxy, X[j], yv);
            y[j] = (int16_t)(yv * 2);
            pk += v;
        }
        yy = s16(tm.sum(pyy));
        xy = tm.sum(pxy);
        pulsesLeft -= tm.sum(pk);
        tm.sync();
    }
    if (pulsesLeft > N + 3) {
        int tmp = s16(pulsesLeft);
        yy = s16(mac16_16(yy, tmp, tmp));
        yy = s16(mac16_16(yy, tmp, y[0]));
        tm.sync();
        if (tm.lane() == 0) iy[0] = (int16_t)(iy[0] + pulsesLeft);
        tm.sync();
        pulsesLeft = 0;
    }
    // shuffle-tree width: lanes >= N hold the sentinel, so only ceil(log2(min(N,W))) levels are needed
    int levels = 0;
    while ((1 << levels) < TM::W && (1 << levels) < N) levels++;
    CB_NOUNROLL for (int i = 0; i < pulsesLeft; i++) {
        const int rshift = 1 + celt_ilog2(K - pulsesLeft + i + 1);
        yy = s16(wadd(yy, 1));
        PvqBest best{-32767, 0, 0};
        CB_TEAM_FOR(j, N, tm) {
            int Rxy = s16(wadd(xy, X[j]) >> rshift);
            const int Ryy = s16(yy + y[j]);
            Rxy = s16(mul16_16_q15(Rxy, Rxy));
            if (mul16_16(best.den, Rxy) > mul16_16(Ryy, best.num)) {
                best.den = Ryy;
                best.num = Rxy;
                best.id = j;
            }
        }
        const int best_id = pvq_pick(tm, best, levels);
        xy = wadd(xy, X[best_id]);
        yy = s16(yy + y[best_id]);
        tm.sync();
        if (tm.lane() == 0) {
            y[best_id] = (int16_t)(y[best_id] + 2);
            iy[best_id]++;
        }
        tm.sync();
    }
    CB_TEAM_FOR(j, N, tm) {
        X[j] = (int16_t)mul16_16(signx[j], X[j]);
        if (signx[j] < 0) iy[j] = (int16_t)(-iy[j]);
    }
    tm.sync();
}

// alg_quant (vq.c:161-325): spreading rotation, search, codeword
template <class TM>
CB_DEV_NOINLINE void alg_quant(TM tm, int16_t *X, int N, int K, int spread, int B, EcEnc &enc, PvqScratch &ps) {
    exp_rotation_enc(tm, X, N, B, K, spread);
    if (TM::W > 1 && N <= TM::W) {
        const int iy = alg_quant_small_core(tm, X, N, K, ps);
        enc.uint_(pvq_index_lane(tm, N, K, iy), pvq_v(N, K));
        return;
    }
    alg_quant_core(tm, X, N, K, ps);
#if defined(CB_WALK_DEBUG)
    { unsigned idx = pvq_encode_index(tm, N, K, ps.iy); if (g_dbg_log) printf("   [%s] leaf N=%d K=%d B=%d idx=%u\n", g_dbg_which == 0 ? "ref" : "slow", N, K, B, idx); }
#endif
    enc.uint_(pvq_encode_index(tm, N, K, ps.iy), pvq_v(N, K));
}

template <class TM>
struct EncBandCtx {
    TM tm;
    EcEnc ec;
    PvqScratch *ps;
    int16_t *tmp;        // hadamard staging, >= 176 int16
    const int *bandE;
    int i, intensity, spread, tf_change;
    int remaining_bits;
};

// compute_theta, encoder half (bands.c:645-817)
template <class TM>
CB_DEV_NOINLINE void compute_theta_enc(EncBandCtx<TM> &ctx, SplitCtx &sctx, int16_t *X, int16_t *Y, int N, int *b, int B, int B0, int LM, int stereo) {
    EcEnc &ec = ctx.ec;
    TM tm = ctx.tm;
    int inv = 0;
    int pulse_cap = kLogN[ctx.i] + LM * (1 << kBitRes);
    int offset = (pulse_cap >> 1) - (stereo && N == 2 ? kQThetaOffsetTwoPhase : kQThetaOffset);
    int qn = compute_qn(N, *b, offset, pulse_cap, stereo);
    if (stereo && ctx.i >= ctx.intensity) qn = 1;
    int itheta = stereo_itheta(tm, X, Y, stereo, N);
    int tell = (int)ec.tell_frac();
    if (qn != 1) {
        itheta = (itheta * qn + 8192) >> 14;
        if (stereo && N > 2) {
            const int p0 = 3;
            int x = itheta, x0 = qn / 2;
            unsigned ft = (unsigned)(p0 * (x0 + 1) + x0);
            ec.encode((unsigned)(x <= x0 ? p0 * x : (x - 1 - x0) + (x0 + 1) * p0),
                      (unsigned)(x <= x0 ? p0 * (x + 1) : (x - x0) + (x0 + 1) * p0), ft);
        } else if (B0 > 1 || stereo) {
            ec.uint_((unsigned)itheta, (unsigned)qn + 1);
        } else {
            int ft = ((qn >> 1) + 1) * ((qn >> 1) + 1);
            int fs = itheta <= (qn >> 1) ? itheta + 1 : qn + 1 - itheta;
            int fl = itheta <= (qn >> 1) ? itheta * (itheta + 1) >> 1 : ft - ((qn + 1 - itheta) * (qn + 2 - itheta) >> 1);
            ec.encode((unsigned)fl, (unsigned)(fl + fs), (unsigned)ft);
        }
        itheta = (int)udiv((unsigned)(itheta * 16384), (unsigned)qn);
        if (stereo) {
            tm.sync();
            if (itheta == 0) intensity_stereo(tm, X, Y, ctx.bandE, ctx.i, N);
            else stereo_split(tm, X, Y, N);
        }
    } else if (stereo) {
        inv = itheta > 8192;
        tm.sync();
        if (inv) negate_vector(tm, Y, N);
        intensity_stereo(tm, X, Y, ctx.bandE, ctx.i, N);
        if (*b > 2 << kBitRes && ctx.remaining_bits > 2 << kBitRes) ec.bit_logp(inv, 2);
        else inv = 0;
        itheta = 0;
    }
    int qalloc = (int)ec.tell_frac() - tell;
    *b -= qalloc;
    int imid, iside, delta;
    if (itheta == 0) { imid = 32767; iside = 0; delta = -16384; }
    else if (itheta == 16384) { imid = 0; iside = 32767; delta = 16384; }
    else {
        imid = bitexact_cos(s16(itheta));
        iside = bitexact_cos(s16(16384 - itheta));
        delta = frac_mul16((N - 1) << 7, bitexact_log2tan(iside, imid));
    }
    sctx.inv = inv; sctx.imid = imid; sctx.iside = iside; sctx.delta = delta; sctx.itheta = itheta; sctx.qalloc = qalloc;
}

template <class TM>
CB_DEV void quant_band_n1_enc(EncBandCtx<TM> &ctx, int16_t *X, int16_t *Y) {
    int16_t *x = X;
    const int nch = Y != nullptr ? 2 : 1;
    CB_NOUNROLL for (int c = 0; c < nch; c++) {
        if (ctx.remaining_bits >= 1 << kBitRes) {
            ctx.ec.bits((unsigned)(x[0] < 0), 1);
            ctx.remaining_bits -= 1 << kBitRes;
        }
        x = Y;
    }
}

// quant_partition, encode (bands.c:864-1040) as an explicit walker (see celt_bands.cuh)
struct EncPartFrame {
    int16_t *X, *Y;
    int N, b, B, LM;
    int mbits, sbits, itheta, rebalance0, mid_first, stage;
};

template <class TM>
CB_DEV void quant_partition_enc(EncBandCtx<TM> &ctx, int16_t *X, int N, int b, int B, int LM) {
    EncPartFrame st[5];
    int sp = 0;
    st[0].X = X; st[0].N = N; st[0].b = b; st[0].B = B; st[0].LM = LM; st[0].stage = 0;
    while (sp >= 0) {
        EncPartFrame &f = st[sp];
        if (f.stage == 0) {
            const uint8_t *cache = pulse_cache(ctx.i, f.LM);
            if (f.LM != -1 && f.b > cache[cache[0]] + 12 && f.N > 2) {
                SplitCtx s;
                const int n = f.N >> 1, lm = f.LM - 1, B0 = f.B;
                const int Bn = (B0 + 1) >> 1;
                int bb = f.b;
                f.Y = f.X + n;
                compute_theta_enc(ctx, s, f.X, f.Y, n, &bb, Bn, B0, lm, 0);
                int delta = s.delta;
                const int itheta = s.itheta;
                if (B0 > 1 && (itheta & 0x3fff)) {
                    if (itheta > 8192) delta -= delta >> (4 - lm);
                    else delta = imin(0, delta + (n << kBitRes >> (5 - lm)));
                }
                const int mbits = imax(0, imin(bb, (bb - delta) / 2));
                const int sbits = bb - mbits;
                ctx.remaining_bits -= s.qalloc;
                f.mbits = mbits; f.sbits = sbits; f.itheta = itheta; f.rebalance0 = ctx.remaining_bits;
                f.N = n; f.LM = lm; f.B = Bn;
                f.mid_first = mbits >= sbits;
                f.stage = 1;
                EncPartFrame &c = st[sp + 1];
                c.N = n; c.B = Bn; c.LM = lm; c.stage = 0;
                if (f.mid_first) { c.X = f.X; c.b = mbits; }
                else { c.X = f.Y; c.b = sbits; }
                sp++;
            } else {
                int q = bits2pulses(ctx.i, f.LM, f.b);
                int curr_bits = pulses2bits(ctx.i, f.LM, q);
                ctx.remaining_bits -= curr_bits;
                while (ctx.remaining_bits < 0 && q > 0) {
                    ctx.remaining_bits += curr_bits;
                    q--;
                    curr_bits = pulses2bits(ctx.i, f.LM, q);
                    ctx.remaining_bits -= curr_bits;
                }
                if (q != 0) alg_quant(ctx.tm, f.X, f.N, get_pulses(q), ctx.spread, f.B, ctx.ec, *ctx.ps);
                sp--;
            }
        } else if (f.stage == 1) {
            EncPartFrame &c = st[sp + 1];
            c.N = f.N; c.B = f.B; c.LM = f.LM; c.stage = 0;
            if (f.mid_first) {
                int rebalance = f.mbits - (f.rebalance0 - ctx.remaining_bits);
                if (rebalance > 3 << kBitRes && f.itheta != 0) f.sbits += rebalance - (3 << kBitRes);
                c.X = f.Y; c.b = f.sbits;
            } else {
                int rebalance = f.sbits - (f.rebalance0 - ctx.remaining_bits);
                if (rebalance > 3 << kBitRes && f.itheta != 16384) f.mbits += rebalance - (3 << kBitRes);
                c.X = f.X; c.b = f.mbits;
            }
            f.stage = 2;
            sp++;
        } else {
            sp--;
        }
    }
}

// quant_band, encode (bands.c:1044-1170 without the resynthesis tail)
template <class TM>
CB_DEV_NOINLINE void quant_band_enc(EncBandCtx<TM> &ctx, int16_t *X, int N, int b, int B, int LM) {
    int N_B = N, B0 = B;
    int recombine = 0;
    int tf_change = ctx.tf_change;
    const int longBlocks = B0 == 1;
    N_B = (int)udiv((unsigned)N_B, (unsigned)B);
    if (N == 1) { quant_band_n1_enc(ctx, X, nullptr); return; }
    if (tf_change > 0) recombine = tf_change;
    CB_NOUNROLL for (int k = 0; k < recombine; k++) haar1_team(ctx.tm, X, N >> k, 1 << k);
    B >>= recombine;
    N_B <<= recombine;
    while ((N_B & 1) == 0 && tf_change < 0) {
        haar1_team(ctx.tm, X, N_B, B);
        B <<= 1;
        N_B >>= 1;
        tf_change++;
    }
    B0 = B;
    if (B0 > 1) deinterleave_hadamard_team(ctx.tm, X, ctx.tmp, N_B >> recombine, B0 << recombine, longBlocks);
    quant_partition_enc(ctx, X, N, b, B, LM);
}

// quant_all_bands with encode = 1 (bands.c:1337-1502) and quant_band_stereo (bands.c:1176-1335) folded in.
// Called by ALL lanes with uniform arguments; ec_io is updated identically on every lane.
template <class TM>
CB_DEV void quant_all_bands_enc(TM tm, int start, int end, int16_t *X_, int16_t *Y_, const int *bandE, const int *pulses, int shortBlocks,
                                int spread, int dual_stereo, int intensity, const int *tf_res, int total_bits, int balance,
                                EcEnc &ec_io, int LM, int codedBands, PvqScratch *ps, int16_t *tmp) {
    const int M = 1 << LM;
    const int B = shortBlocks ? M : 1;
    EncBandCtx<TM> ctx;
    ctx.tm = tm;
    ctx.ec = ec_io; ctx.ps = ps; ctx.tmp = tmp; ctx.bandE = bandE;
    ctx.intensity = intensity; ctx.spread = spread;
    CB_NOUNROLL for (int i = start; i < end; i++) {
#if !defined(CB_NO_BAND_PHASE)
        tm.phase();   // one per band, kNbEBands per frame (balanced below): co-resident streams walk the band loop together
#endif
        ctx.i = i;
#if defined(CB_WALK_DEBUG)
        g_dbg_tell[0][i] = (int)ctx.ec.tell_frac(); g_dbg_rng[0][i] = ctx.ec.rng;
#endif
        int16_t *X = X_ + M * kEBands[i];
        int16_t *Y = Y_ != nullptr ? Y_ + M * kEBands[i] : nullptr;
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
        ctx.tf_change = tf_res[i];
        if (dual_stereo && i == intensity) dual_stereo = 0;
        if (dual_stereo) {
            quant_band_enc(ctx, X, N, b / 2, B, LM);
            quant_band_enc(ctx, Y, N, b / 2, B, LM);
        } else if (Y != nullptr) {
            if (N == 1) {
                quant_band_n1_enc(ctx, X, Y);
            } else {
                SplitCtx s;
                int bs = b;
                compute_theta_enc(ctx, s, X, Y, N, &bs, B, B, LM, 1);
                if (N == 2) {
                    int mbits = bs, sbits = 0;
                    if (s.itheta != 0 && s.itheta != 16384) sbits = 1 << kBitRes;
                    mbits -= sbits;
                    const int c = s.itheta > 8192;
                    ctx.remaining_bits -= s.qalloc + sbits;
                    int16_t *x2 = c ? Y : X;
                    int16_t *y2 = c ? X : Y;
                    if (sbits) {
                        int sign = wsub(wmul(x2[0], y2[1]), wmul(x2[1], y2[0])) < 0;
                        ctx.ec.bits((unsigned)sign, 1);
                    }
                    quant_band_enc(ctx, x2, N, mbits, B, LM);
                } else {
                    int mbits = imax(0, imin(bs, (bs - s.delta) / 2));
                    int sbits = bs - mbits;
                    ctx.remaining_bits -= s.qalloc;
                    int rebalance = ctx.remaining_bits;
                    if (mbits >= sbits) {
                        quant_band_enc(ctx, X, N, mbits, B, LM);
                        rebalance = mbits - (rebalance - ctx.remaining_bits);
                        if (rebalance > 3 << kBitRes && s.itheta != 0) sbits += rebalance - (3 << kBitRes);
                        quant_band_enc(ctx, Y, N, sbits, B, LM);
                    } else {
                        quant_band_enc(ctx, Y, N, sbits, B, LM);
                        rebalance = sbits - (rebalance - ctx.remaining_bits);
                        if (rebalance > 3 << kBitRes && s.itheta != 16384) mbits += rebalance - (3 << kBitRes);
                        quant_band_enc(ctx, X, N, mbits, B, LM);
                    }
                }
            }
        } else {
            quant_band_enc(ctx, X, N, b, B, LM);
        }
        balance += pulses[i] + tell;
    }
#if !defined(CB_NO_BAND_PHASE)
    CB_NOUNROLL for (int i = end - start; i < kNbEBands; i++) tm.phase();
#endif
    ec_io = ctx.ec;
}

}  // namespace cb
