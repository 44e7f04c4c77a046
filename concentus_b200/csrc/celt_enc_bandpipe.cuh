// celt_enc_bandpipe.cuh — the encoder's band loop (quant_all_bands with encode = 1, opus-fix/celt/bands.c:1337-1502) cut into a
// data-parallel part and a scalar chain, for the frame-synchronous pipeline (celt_enc_pipe.cuh).
//
// Without resynthesis the encoder never folds, so a band's vector work does not depend on other bands' results.  What serialises
// the reference's loop is the bit budget: a band's allocation b comes from ec_tell_frac() after the previous band, a split's
// angle resolution qn from b, a leaf's pulse count K from what is left.  But:
//   * every ANGLE the loop can ask for is a function of the spectrum alone — the stereo angle of a band (vq.c:376-408) and, for
//     every node of the binary split tree of a band vector (after quant_band's haar / hadamard reordering, which depends only on
//     tf_change and B), the angle between its halves; the budget only decides how finely an angle is quantised;
//   * the search for K pulses in a leaf (vq.c:161-325) needs only K; its result feeds back into the chain through ONE thing: when
//     the top symbol of the codeword index is 0 the range coder's new range is r + rng % ft instead of r (entenc.c:187-216),
//     which can move a later ec_tell_frac() by 1/8 bit.
// So the band loop runs as
//   prep    warp / stream      stereo angles, mid / side / intensity vectors, quant_band's reordering, the angle trees
//   chain-S thread / stream    the whole loop on a range coder that only tracks (rng, nbits_total), codeword top symbols assumed
//                              non-zero: emits the list of leaves (position, N, K, B)
//   leaves  sub-warp / leaf    exp_rotation + PVQ search + icwrs for every listed leaf -> codeword indices
//   chain-X thread / stream    the loop again with the real range coder and the real indices.  A leaf whose (N, K) differs from
//                              what chain-S listed — its budget moved by that 1/8 bit — is searched in the thread (rare).
// Both chains are one function template (band_walk) over a policy; bit-exactness does not rest on the speculation being right.
#pragma once
#include "celt_enc_bands.cuh"

namespace cb {

enum { kTreeNodes = 15, kMaxLeafTasks = 224, kXallStride = 3 * kMaxFrame };

// What prep leaves for the chains.  Vector ids: 0 = the X region (mono band / left / mid after stereo_split), 1 = the Y region
// (right / side), 2 = the XB region (mid of intensity_stereo, bands.c:337-360).
struct BandPrep {
    int16_t theta_st[kNbEBands];               // raw stereo angle of the band (before quantisation)
    int8_t hasB[kNbEBands];                    // XB holds the intensity mid of this band
    int n2_d[kNbEBands];                       // N == 2 stereo bands: x[0]*y[1] - x[1]*y[0] after stereo_split (bands.c:1241)
    int16_t tree[kNbEBands][3][kTreeNodes];    // raw split angle of node h (heap order) of vector v of band i
};

struct LeafTask {
    uint16_t off;      // position in the stream's vector space [X | Y | XB]
    uint8_t N, K, B, band;
    uint16_t pad;
};
struct LeafList {
    int count;
    int overflow;
    uint8_t band_first[kNbEBands + 1];         // first task of each band (tasks are in walk order)
    LeafTask task[kMaxLeafTasks];
    uint32_t index[kMaxLeafTasks];             // icwrs of the searched pulse vector
};

// bits2pulses (rate.h:53-78) as a table: the binary search over the pulse cache only compares bits - 1 with cache entries <= 255,
// so every bits >= 257 gives what 257 gives (checked exhaustively on the host for all bands and LM).  Filled at start-up by the
// function itself (opus_enc_pipe.cu).
enum { kB2pBits = 258 };
#if defined(__CUDACC__)
static __device__ uint8_t g_b2p_lut[(kMaxLM + 2) * kNbEBands * kB2pBits];
CB_DEV int bits2pulses_fast(int band, int LM, int bits) { return g_b2p_lut[((LM + 1) * kNbEBands + band) * kB2pBits + imin(bits, kB2pBits - 1)]; }
#else
CB_DEV int bits2pulses_fast(int band, int LM, int bits) { return bits2pulses(band, LM, bits); }
#endif

// ---- prep ---------------------------------------------------------------------------------------------------------------------

// quant_band's reordering of one vector (bands.c:1062-1100): recombine / time-divide haar steps, then the hadamard de-interleave.
// Returns the B the partition starts with.
template <class TM>
CB_DEV int band_reorder_team(TM tm, int16_t *X, int N, int B, int tf_change, int16_t *tmp) {
    int N_B = (int)udiv((unsigned)N, (unsigned)B);
    int recombine = 0;
    const int longBlocks = B == 1;
    if (tf_change > 0) recombine = tf_change;
    CB_NOUNROLL for (int k = 0; k < recombine; k++) haar1_team(tm, X, N >> k, 1 << k);
    B >>= recombine;
    N_B <<= recombine;
    while ((N_B & 1) == 0 && tf_change < 0) {
        haar1_team(tm, X, N_B, B);
        B <<= 1;
        N_B >>= 1;
        tf_change++;
    }
    if (B > 1) deinterleave_hadamard_team(tm, X, tmp, N_B >> recombine, B << recombine, longBlocks);
    return B;
}
// the same bookkeeping without the vector work (the chains need B)
CB_DEV int band_reorder_B(int N, int B, int tf_change) {
    int N_B = N >> celt_ilog2(B);                 // B is 1, 2, 4 or 8
    int recombine = 0;
    if (tf_change > 0) recombine = tf_change;
    B >>= recombine;
    N_B <<= recombine;
    while ((N_B & 1) == 0 && tf_change < 0) {
        B <<= 1;
        N_B >>= 1;
        tf_change++;
    }
    return B;
}

// levels of the split tree a vector of N values can have at frame resolution LM (quant_partition splits while LM != -1 && N > 2)
CB_DEV int tree_levels(int N, int LM) {
    int d = 0;
    while (LM != -1 && N > 2) { N >>= 1; LM--; d++; }
    return d;
}

// raw angle between two halves from their energies (stereo_itheta with stereo = 0, vq.c:376-408)
CB_DEV int theta_from_energies(int Eleft, int Eright) {
    const int mid = s16(celt_sqrt(wadd(1, Eleft)));
    const int side = s16(celt_sqrt(wadd(1, Eright)));
    return mul16_16_q15(20861, celt_atan2p(side, mid));
}

// The angle tree of one reordered vector: seg[] = team scratch for 16 ints.
template <class TM>
CB_DEV void band_tree_team(TM tm, const int16_t *X, int N, int LM, int16_t *tree, int *seg) {
    const int D = tree_levels(N, LM);
    if (D == 0) return;
    const int nseg = 1 << D, len = N >> D;
    CB_TEAM_FOR(k, nseg, tm) {
        int e = 0;
        CB_NOUNROLL for (int j = 0; j < len; j++) e = mac16_16(e, X[k * len + j], X[k * len + j]);
        seg[k] = e;
    }
    tm.sync();
    const int nnodes = nseg - 1;
    CB_TEAM_FOR(h, nnodes, tm) {
        int l = 0;
        while ((2 << l) - 1 <= h) l++;            // level of node h: (1<<l)-1 <= h < (2<<l)-1
        const int j = h - ((1 << l) - 1);
        const int span = nseg >> l;               // finest segments under this node
        int el = 0, er = 0;
        CB_NOUNROLL for (int k = 0; k < span / 2; k++) { el = wadd(el, seg[j * span + k]); er = wadd(er, seg[j * span + span / 2 + k]); }
        tree[h] = (int16_t)theta_from_energies(el, er);
    }
    tm.sync();
}

// prep of one stream's frame.  Xall: [X | Y | XB] in team-shared memory (X / Y = the normalised spectrum on entry).
template <class TM>
CB_DEV void band_prep_team(TM tm, int16_t *Xall, const int *bandE, const int *tf_res, int end, int C, int LM, int shortBlocks, int dual_stereo,
                           int intensity, BandPrep &P, int16_t *tmp, int *seg) {
    const int M = 1 << LM;
    const int B0 = shortBlocks ? M : 1;
    int16_t *Xb = Xall, *Yb = Xall + kMaxFrame, *XBb = Xall + 2 * kMaxFrame;
    CB_NOUNROLL for (int i = 0; i < end; i++) {
        int16_t *X = Xb + M * kEBands[i], *Y = Yb + M * kEBands[i], *XB = XBb + M * kEBands[i];
        const int N = M * kEBands[i + 1] - M * kEBands[i];
        if (dual_stereo && i == intensity) dual_stereo = 0;
        int hasB = 0;
        if (N == 1) {
            if (tm.lane() == 0) P.hasB[i] = 0;
            continue;
        }
        if (C == 2 && !dual_stereo) {
            const int itheta = stereo_itheta(tm, X, Y, 1, N);
            {
                // The intensity mid (bands.c:337-360).  It is what gets coded when the angle quantiser has a single step (qn == 1:
                // bands from `intensity` up, or any band short of bits — then the side is first negated when itheta > 8192,
                // bands.c:771-790) and when a finer quantiser rounds the angle to 0 (only possible below 4096, where inv = 0).
                const int inv = itheta > 8192;
                int shift = celt_zlog2(imax(bandE[i], bandE[i + kNbEBands])) - 13;
                int left = s16(vshr32(bandE[i], shift));
                int right = s16(vshr32(bandE[i + kNbEBands], shift));
                int norm = s16(1 + celt_sqrt(wadd(1, wadd(mul16_16(left, left), mul16_16(right, right)))));
                int a1 = s16(shl32(left, 14) / norm);
                int a2 = s16(shl32(right, 14) / norm);
                CB_TEAM_FOR(j, N, tm) XB[j] = (int16_t)(mac16_16(mul16_16(a1, X[j]), a2, inv ? (int)(int16_t)(-Y[j]) : (int)Y[j]) >> 14);
                hasB = 1;
            }
            if (i < intensity) {
                tm.sync();
                stereo_split(tm, X, Y, N);
                if (N == 2 && tm.lane() == 0) P.n2_d[i] = wsub(wmul(X[0], Y[1]), wmul(X[1], Y[0]));
            }
            if (tm.lane() == 0) P.theta_st[i] = (int16_t)itheta;
            tm.sync();
        }
        if (tm.lane() == 0) P.hasB[i] = (int8_t)hasB;
        // reorder + angle trees of every vector the chains may walk
        const int tf_change = tf_res[i];
        const int nv = C == 2 ? 2 : 1;
        const bool stereo_hi = C == 2 && !dual_stereo && i >= intensity;   // only the intensity mid is ever coded
        CB_NOUNROLL for (int v = 0; v < 3; v++) {
            if (v < 2 && (v >= nv || stereo_hi)) continue;
            if (v == 2 && !hasB) continue;
            int16_t *V = v == 0 ? X : v == 1 ? Y : XB;
            band_reorder_team(tm, V, N, B0, tf_change, tmp);
            band_tree_team(tm, V, N, LM, P.tree[i][v], seg);
        }
    }
}

// ---- the walk's scalar state --------------------------------------------------------------------------------------------------
struct WalkCtx {
    const BandPrep *prep;
    int i, intensity, spread, remaining_bits;
};
struct WalkFrame {
    int off, N, b, B, LM, h;
    int mbits, sbits, itheta, rebalance0, mid_first, stage;
};
// The scalar state of a team's walk.  In the warp-per-stream walk every lane runs the scalar code with identical values; kept in
// registers / on the stack that is 32 private copies of the coder, the partition stack and the band context per warp — half a KB
// of local memory per LANE, 28 warps per SM: far beyond L1 (profiles/r2b_pipe_walk: 29 % of the stalls were local loads that
// missed).  So the team keeps ONE copy in shared memory: every lane loads the same word (a broadcast) and stores the same value.
struct WalkShared {
    EcEnc ec;
    WalkCtx w;
    SplitCtx split;
    WalkFrame st[5];
};

// ---- the scalar chain ---------------------------------------------------------------------------------------------------------

// A range coder that only knows how many bits have been spent: (rng, nbits_total) of entenc.c, no bytes.
struct TellCoder {
    unsigned rng;
    int nbits_total;
    CB_MEM void from(const EcEnc &e) { rng = e.rng; nbits_total = e.nbits_total; }
    CB_MEM void normalize() {
        while (rng <= CB_EC_CODE_BOT) { rng <<= kEcSymBits; nbits_total += kEcSymBits; }
    }
    CB_MEM unsigned tell_frac() const {
        unsigned nbits = (unsigned)nbits_total << 3;
        int l = ec_ilog(rng);
        unsigned r = rng >> (l - 16);
        unsigned b = (r >> 12) - 8;
        const unsigned corr = b == 0 ? 35733u : b == 1 ? 38967u : b == 2 ? 42495u : b == 3 ? 46340u :
                              b == 4 ? 50535u : b == 5 ? 55109u : b == 6 ? 60097u : 65535u;
        b += r > corr;
        l = (l << 3) + (int)b;
        return nbits - (unsigned)l;
    }
    CB_MEM void encode(unsigned fl, unsigned fh, unsigned ft) {
        unsigned r = rng / ft;
        if (fl > 0) rng = r * (fh - fl);
        else rng -= r * (ft - fh);
        normalize();
    }
    CB_MEM void bit_logp(int v, unsigned logp) {
        unsigned s = rng >> logp;
        rng = v ? s : rng - s;
        normalize();
    }
    CB_MEM void bits(unsigned nb) { nbits_total += (int)nb; }
    CB_MEM void uint_(unsigned fl, unsigned ft_in) {
        unsigned ft = ft_in - 1;
        int ftb = ec_ilog(ft);
        if (ftb > kEcUintBits) {
            ftb -= kEcUintBits;
            unsigned f = (ft >> ftb) + 1;
            unsigned l = fl >> ftb;
            encode(l, l + 1, f);
            bits((unsigned)ftb);
        } else {
            encode(fl, fl + 1, ft + 1);
        }
    }
    // a codeword whose value is not known yet: its top symbol is taken to be non-zero
    CB_MEM void uint_unknown(unsigned ft_in) {
        unsigned ft = ft_in - 1;
        int ftb = ec_ilog(ft);
        if (ftb > kEcUintBits) {
            ftb -= kEcUintBits;
            rng = rng / ((ft >> ftb) + 1);
            normalize();
            bits((unsigned)ftb);
        } else {
            rng = rng / (ft + 1);
            normalize();
        }
    }
};

// chain-S policy: budget-only coder, leaves are listed
struct SpecPolicy {
    TellCoder ec;
    LeafList *list;
    int cur_band;
    WalkCtx wc;
    SplitCtx sc;
    WalkFrame fr[5];
    CB_MEM WalkCtx &wctx() { return wc; }
    CB_MEM SplitCtx &split() { return sc; }
    CB_MEM WalkFrame *frames() { return fr; }
    CB_MEM unsigned tell_frac() const { return ec.tell_frac(); }
    CB_MEM void encode(unsigned fl, unsigned fh, unsigned ft) { ec.encode(fl, fh, ft); }
    CB_MEM void uint_(unsigned fl, unsigned ft) { ec.uint_(fl, ft); }
    CB_MEM void bit_logp(int v, unsigned logp) { ec.bit_logp(v, logp); }
    CB_MEM void sign_bit(int /*off*/) { ec.bits(1); }
    CB_MEM void n2_sign(int /*band*/, int /*c*/) { ec.bits(1); }
    CB_MEM void begin_band(int i) {
        while (cur_band < i) list->band_first[++cur_band] = (uint8_t)list->count;
    }
    CB_MEM void leaf(int band, int off, int N, int K, int B, int /*spread*/) {
        const int k = list->count;
        if (k < kMaxLeafTasks) {
            LeafTask t;
            t.off = (uint16_t)off; t.N = (uint8_t)N; t.K = (uint8_t)K; t.B = (uint8_t)B; t.band = (uint8_t)band; t.pad = 0;
            list->task[k] = t;
            list->count = k + 1;
        } else {
            list->overflow = 1;
        }
        ec.uint_unknown(pvq_v(N, K));
    }
    CB_MEM void finish() {
        while (cur_band < kNbEBands) list->band_first[++cur_band] = (uint8_t)list->count;
    }
    CB_MEM void skip_bands(int) {}
};

// chain-X policy: the real coder, the real indices; a leaf chain-S did not list as such is searched here
struct ExactPolicy {
    EcEnc ec;
    const LeafList *list;
    int16_t *Xall;             // the stream's prepared vectors (global memory)
    const BandPrep *prep;
    PvqScratch *ps;            // thread scratch of the in-thread search
    int cursor;
    int misses;
    WalkCtx wc;
    SplitCtx sc;
    WalkFrame fr[5];
    CB_MEM WalkCtx &wctx() { return wc; }
    CB_MEM SplitCtx &split() { return sc; }
    CB_MEM WalkFrame *frames() { return fr; }
    CB_MEM unsigned tell_frac() const { return ec.tell_frac(); }
    CB_MEM void encode(unsigned fl, unsigned fh, unsigned ft) { ec.encode(fl, fh, ft); }
    CB_MEM void uint_(unsigned fl, unsigned ft) { ec.uint_(fl, ft); }
    CB_MEM void bit_logp(int v, unsigned logp) { ec.bit_logp(v, logp); }
    CB_MEM void sign_bit(int off) { ec.bits((unsigned)(Xall[off] < 0), 1); }
    CB_MEM void n2_sign(int band, int c) {
        const int d = prep->n2_d[band];
#if defined(CB_WALK_DEBUG)
        if (g_dbg_log) printf("   [exact] band %d n2 sign c=%d d=%d\n", band, c, d);
#endif
        ec.bits((unsigned)(c ? wneg(d) < 0 : d < 0), 1);
    }
    CB_MEM void begin_band(int) {}
    CB_MEM void skip_bands(int) {}
    CB_MEM void leaf(int band, int off, int N, int K, int B, int spread) {
        int hit = -1;
        const int cnt = list->count;
        if (cursor < cnt) {
            const LeafTask &t = list->task[cursor];
            if (t.off == off && t.N == N && t.K == K && t.B == B) hit = cursor;
        }
        if (hit < 0) {
            const int a = list->band_first[band], b = list->band_first[band + 1];
            CB_NOUNROLL for (int k = a; k < b && k < cnt; k++) {
                const LeafTask &t = list->task[k];
                if (t.off == off && t.N == N && t.K == K && t.B == B) { hit = k; break; }
            }
        }
        if (hit >= 0) {
#if defined(CB_WALK_DEBUG)
            if (g_dbg_log) printf("   [exact] band %d leaf off=%d N=%d K=%d B=%d idx=%u\n", band, off, N, K, B, list->index[hit]);
#endif
            cursor = hit + 1;
            ec.uint_(list->index[hit], pvq_v(N, K));
        } else {
            misses++;
            alg_quant(SoloTeam{}, Xall + off, N, K, spread, B, ec, *ps);
        }
    }
    CB_MEM void finish() {}
};

// Inline policy: the real coder, and every leaf searched by the whole team the moment the walk reaches it.  ALL lanes of the team
// run the walk with identical scalars (range coder included), so nothing is broadcast; only the leaf's vector work is split.
template <class TM>
struct InlinePolicy {
    TM tm;
    WalkShared *sh;            // the team's one copy of the scalar state (team-shared memory)
    EcEnc &ec;                 // = sh->ec
    const int16_t *Xall;       // the stream's prepared vectors [X | Y | intensity mid] where prep left them (global memory, L2)
    int16_t *leafbuf;          // team-shared copy of the leaf being quantised (176 values)
    const BandPrep *prep;
    PvqScratch *ps;
    CB_MEM WalkCtx &wctx() { return sh->w; }
    CB_MEM SplitCtx &split() { return sh->split; }
    CB_MEM WalkFrame *frames() { return sh->st; }
    CB_MEM unsigned tell_frac() const { return ec.tell_frac(); }
    CB_MEM void encode(unsigned fl, unsigned fh, unsigned ft) { ec.encode(fl, fh, ft); }
    CB_MEM void uint_(unsigned fl, unsigned ft) { ec.uint_(fl, ft); }
    CB_MEM void bit_logp(int v, unsigned logp) { ec.bit_logp(v, logp); }
    CB_MEM void sign_bit(int off) { ec.bits((unsigned)(Xall[off] < 0), 1); }
    CB_MEM void n2_sign(int band, int c) {
        const int d = prep->n2_d[band];
        ec.bits((unsigned)(c ? wneg(d) < 0 : d < 0), 1);
    }
    unsigned sync_mask;        // bands at whose start the block's warps meet (SyncWarpTeam); every warp passes the same barriers
    CB_MEM void begin_band(int i) { if ((sync_mask >> i) & 1u) tm.phase(); }
    CB_MEM void leaf(int, int off, int N, int K, int B, int spread) {
        // Only the leaf being searched is staged: a whole-frame copy of the vectors (5.6 KB per warp, 160 KB per SM) left the
        // kernel ~50 KB of L1 for its tables, the angle trees and the spills.
        CB_TEAM_FOR(j, N, tm) leafbuf[j] = Xall[off + j];
        tm.sync();
        alg_quant(tm, leafbuf, N, K, spread, B, ec, *ps);
        tm.sync();             // the lanes go on through the scalar state together
    }
    CB_MEM void finish() {}
    CB_MEM void skip_bands(int n) {   // the bands this frame does not code: the barriers of the mask are still passed
        CB_NOUNROLL for (int i = kNbEBands - n; i < kNbEBands; i++)
            if ((sync_mask >> i) & 1u) tm.phase();
    }
};

// compute_theta, encoder half (bands.c:645-817), given the raw angle
template <class P>
CB_DEV_NOINLINE void theta_code(P &p, WalkCtx &w, SplitCtx &sctx, int itheta, int N, int *b, int B0, int LM, int stereo) {
    int inv = 0;
    int pulse_cap = kLogN[w.i] + LM * (1 << kBitRes);
    int offset = (pulse_cap >> 1) - (stereo && N == 2 ? kQThetaOffsetTwoPhase : kQThetaOffset);
    int qn = compute_qn(N, *b, offset, pulse_cap, stereo);
    if (stereo && w.i >= w.intensity) qn = 1;
    const int tell = (int)p.tell_frac();
    if (qn != 1) {
        itheta = (itheta * qn + 8192) >> 14;
        if (stereo && N > 2) {
            const int p0 = 3;
            int x = itheta, x0 = qn / 2;
            unsigned ft = (unsigned)(p0 * (x0 + 1) + x0);
            p.encode((unsigned)(x <= x0 ? p0 * x : (x - 1 - x0) + (x0 + 1) * p0),
                     (unsigned)(x <= x0 ? p0 * (x + 1) : (x - x0) + (x0 + 1) * p0), ft);
        } else if (B0 > 1 || stereo) {
            p.uint_((unsigned)itheta, (unsigned)qn + 1);
        } else {
            int ft = ((qn >> 1) + 1) * ((qn >> 1) + 1);
            int fs = itheta <= (qn >> 1) ? itheta + 1 : qn + 1 - itheta;
            int fl = itheta <= (qn >> 1) ? itheta * (itheta + 1) >> 1 : ft - ((qn + 1 - itheta) * (qn + 2 - itheta) >> 1);
            p.encode((unsigned)fl, (unsigned)(fl + fs), (unsigned)ft);
        }
        itheta = (int)udiv((unsigned)(itheta * 16384), (unsigned)qn);
    } else if (stereo) {
        inv = itheta > 8192;
        if (*b > 2 << kBitRes && w.remaining_bits > 2 << kBitRes) p.bit_logp(inv, 2);
        else inv = 0;
        itheta = 0;
    }
    const int qalloc = (int)p.tell_frac() - tell;
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

// quant_band (bands.c:1044-1170) on vector v of the band, positioned at `off`: the reordering was done by prep, the partition
// (bands.c:864-1040) is walked with prep's angle tree
template <class P>
CB_DEV_NOINLINE void walk_band_vector(P &p, WalkCtx &w, int v, int off, int N, int b, int B, int LM, int tf_change) {
    if (N == 1) {
        if (w.remaining_bits >= 1 << kBitRes) {
            p.sign_bit(off);
            w.remaining_bits -= 1 << kBitRes;
        }
        return;
    }
    B = band_reorder_B(N, B, tf_change);
    const int16_t *tree = w.prep->tree[w.i][v];
    WalkFrame *st = p.frames();
    int sp = 0;
    st[0].off = off; st[0].N = N; st[0].b = b; st[0].B = B; st[0].LM = LM; st[0].h = 0; st[0].stage = 0;
    while (sp >= 0) {
        WalkFrame &f = st[sp];
        if (f.stage == 0) {
            const uint8_t *cache = pulse_cache(w.i, f.LM);
            if (f.LM != -1 && f.b > cache[cache[0]] + 12 && f.N > 2) {
                SplitCtx &s = p.split();
                const int n = f.N >> 1, lm = f.LM - 1, B0 = f.B;
                const int Bn = (B0 + 1) >> 1;
                int bb = f.b;
                theta_code(p, w, s, tree[f.h], n, &bb, B0, lm, 0);
                int delta = s.delta;
                const int itheta = s.itheta;
                if (B0 > 1 && (itheta & 0x3fff)) {
                    if (itheta > 8192) delta -= delta >> (4 - lm);
                    else delta = imin(0, delta + (n << kBitRes >> (5 - lm)));
                }
                const int mbits = imax(0, imin(bb, (bb - delta) / 2));
                const int sbits = bb - mbits;
                w.remaining_bits -= s.qalloc;
                f.mbits = mbits; f.sbits = sbits; f.itheta = itheta; f.rebalance0 = w.remaining_bits;
                f.N = n; f.LM = lm; f.B = Bn;
                f.mid_first = mbits >= sbits;
                f.stage = 1;
                WalkFrame &c = st[sp + 1];
                c.N = n; c.B = Bn; c.LM = lm; c.stage = 0;
                if (f.mid_first) { c.off = f.off; c.b = mbits; c.h = 2 * f.h + 1; }
                else { c.off = f.off + n; c.b = sbits; c.h = 2 * f.h + 2; }
                sp++;
            } else {
                int q = bits2pulses_fast(w.i, f.LM, f.b);
                int curr_bits = pulses2bits(w.i, f.LM, q);
                w.remaining_bits -= curr_bits;
                while (w.remaining_bits < 0 && q > 0) {
                    w.remaining_bits += curr_bits;
                    q--;
                    curr_bits = pulses2bits(w.i, f.LM, q);
                    w.remaining_bits -= curr_bits;
                }
                if (q != 0) p.leaf(w.i, f.off, f.N, get_pulses(q), f.B, w.spread);
                sp--;
            }
        } else if (f.stage == 1) {
            WalkFrame &c = st[sp + 1];
            c.N = f.N; c.B = f.B; c.LM = f.LM; c.stage = 0;
            if (f.mid_first) {
                int rebalance = f.mbits - (f.rebalance0 - w.remaining_bits);
                if (rebalance > 3 << kBitRes && f.itheta != 0) f.sbits += rebalance - (3 << kBitRes);
                c.off = f.off + f.N; c.b = f.sbits; c.h = 2 * f.h + 2;
            } else {
                int rebalance = f.sbits - (f.rebalance0 - w.remaining_bits);
                if (rebalance > 3 << kBitRes && f.itheta != 16384) f.mbits += rebalance - (3 << kBitRes);
                c.off = f.off; c.b = f.mbits; c.h = 2 * f.h + 1;
            }
            f.stage = 2;
            sp++;
        } else {
            sp--;
        }
    }
}

// quant_all_bands with encode = 1 (bands.c:1337-1502), scalar, over prep's results
template <class P>
CB_DEV void band_walk(P &p, const BandPrep &prep, int end, int C, const int *pulses, int shortBlocks, int spread, int dual_stereo, int intensity,
                      const int *tf_res, int total_bits, int balance, int LM, int codedBands) {
    const int M = 1 << LM;
    const int B = shortBlocks ? M : 1;
    WalkCtx &w = p.wctx();
    w.prep = &prep; w.intensity = intensity; w.spread = spread;
    CB_NOUNROLL for (int i = 0; i < end; i++) {
        w.i = i;
        p.begin_band(i);
#if defined(CB_WALK_DEBUG)
        g_dbg_tell[g_dbg_which][i] = (int)p.tell_frac(); g_dbg_rng[g_dbg_which][i] = p.ec.rng;
#endif
        const int offX = M * kEBands[i], offY = kMaxFrame + offX, offB = 2 * kMaxFrame + offX;
        const int N = M * kEBands[i + 1] - M * kEBands[i];
        const int tell = (int)p.tell_frac();
        if (i != 0) balance -= tell;
        const int remaining_bits = total_bits - tell - 1;
        w.remaining_bits = remaining_bits;
        int b;
        if (i <= codedBands - 1) {
            int curr_balance = sudiv(balance, imin(3, codedBands - i));
            b = imax(0, imin(16383, imin(remaining_bits + 1, pulses[i] + curr_balance)));
        } else {
            b = 0;
        }
        const int tf_change = tf_res[i];
        if (dual_stereo && i == intensity) dual_stereo = 0;
        if (C == 2 && dual_stereo) {
            walk_band_vector(p, w, 0, offX, N, b / 2, B, LM, tf_change);
            walk_band_vector(p, w, 1, offY, N, b / 2, B, LM, tf_change);
        } else if (C == 2) {
            if (N == 1) {
                CB_NOUNROLL for (int c = 0; c < 2; c++)
                    if (w.remaining_bits >= 1 << kBitRes) {
                        p.sign_bit(c ? offY : offX);
                        w.remaining_bits -= 1 << kBitRes;
                    }
            } else {
                SplitCtx s;
                int bs = b;
                {
                    SplitCtx &sr = p.split();   // the nested walks reuse the policy's split context: keep this band's own copy
                    theta_code(p, w, sr, prep.theta_st[i], N, &bs, B, LM, 1);
                    s = sr;
                }
                // itheta == 0: the band was turned into its intensity mid (XB), the side is empty
                const int vm = s.itheta == 0 ? 2 : 0;
                const int offM = s.itheta == 0 ? offB : offX;
                if (N == 2) {
                    int mbits = bs, sbits = 0;
                    if (s.itheta != 0 && s.itheta != 16384) sbits = 1 << kBitRes;
                    mbits -= sbits;
                    const int c = s.itheta > 8192;
                    w.remaining_bits -= s.qalloc + sbits;
                    if (sbits) p.n2_sign(i, c);
                    if (c) walk_band_vector(p, w, 1, offY, N, mbits, B, LM, tf_change);
                    else walk_band_vector(p, w, vm, offM, N, mbits, B, LM, tf_change);
                } else {
                    int mbits = imax(0, imin(bs, (bs - s.delta) / 2));
                    int sbits = bs - mbits;
                    w.remaining_bits -= s.qalloc;
                    int rebalance = w.remaining_bits;
                    if (mbits >= sbits) {
                        walk_band_vector(p, w, vm, offM, N, mbits, B, LM, tf_change);
                        rebalance = mbits - (rebalance - w.remaining_bits);
                        if (rebalance > 3 << kBitRes && s.itheta != 0) sbits += rebalance - (3 << kBitRes);
                        walk_band_vector(p, w, 1, offY, N, sbits, B, LM, tf_change);
                    } else {
                        walk_band_vector(p, w, 1, offY, N, sbits, B, LM, tf_change);
                        rebalance = sbits - (rebalance - w.remaining_bits);
                        if (rebalance > 3 << kBitRes && s.itheta != 16384) mbits += rebalance - (3 << kBitRes);
                        walk_band_vector(p, w, vm, offM, N, mbits, B, LM, tf_change);
                    }
                }
            }
        } else {
            walk_band_vector(p, w, 0, offX, N, b, B, LM, tf_change);
        }
        balance += pulses[i] + tell;
    }
    p.skip_bands(kNbEBands - end);
    p.finish();
}

// ---- leaves ---------------------------------------------------------------------------------------------------------------------

// One leaf with the whole team: exp_rotation + search + icwrs.  V: the leaf's N values in team-shared memory (clobbered).
template <class TM>
CB_DEV unsigned leaf_quant_team(TM tm, int16_t *V, int N, int K, int spread, int B, PvqScratch &ps) {
    exp_rotation_enc(tm, V, N, B, K, spread);
    if (TM::W > 1 && N <= TM::W) alg_quant_small_core(tm, V, N, K, ps);
    else alg_quant_core(tm, V, N, K, ps);
    return pvq_encode_index(tm, N, K, ps.iy);
}

}  // namespace cb
