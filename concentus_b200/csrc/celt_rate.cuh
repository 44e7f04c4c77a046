// celt_rate.cuh — bit allocation: pulse cache lookups, init_caps and compute_allocation.
//
// Restates opus-fix/celt/rate.h:48-85 (get_pulses, bits2pulses, pulses2bits), celt/celt.c:255-264
// (init_caps) and celt/rate.c:248-638 (interp_bits2pulses, compute_allocation) for the standard
// 48 kHz mode (21 bands, 11 allocation vectors).  Purely scalar, 21-element work: it runs inside a
// lane-0 section of the frame driver with its arrays in team-shared memory.  The range coder is a
// template parameter so the same allocation code serves the decoder (reads skip/intensity/dual-stereo
// symbols) and the encoder (writes them).
#pragma once
#include "celt_ec.cuh"
#include "celt_tables.cuh"

namespace cb {

CB_DEV int band_width(int j) { return kEBands[j + 1] - kEBands[j]; }

CB_DEV int get_pulses(int i) { return i < 8 ? i : (8 + (i & 7)) << ((i >> 3) - 1); }

CB_DEV const uint8_t *pulse_cache(int band, int LM) { return kCacheBits + kCacheIndex[(LM + 1) * kNbEBands + band]; }

CB_MATH int bits2pulses(int band, int LM, int bits) {
    const uint8_t *cache = pulse_cache(band, LM);
    int lo = 0, hi = cache[0];
    bits--;
    CB_NOUNROLL for (int i = 0; i < kLogMaxPseudo; i++) {
        int mid = (lo + hi + 1) >> 1;
        if ((int)cache[mid] >= bits) hi = mid;
        else lo = mid;
    }
    if (bits - (lo == 0 ? -1 : (int)cache[lo]) <= (int)cache[hi] - bits) return lo;
    return hi;
}
CB_DEV int pulses2bits(int band, int LM, int pulses) {
    return pulses == 0 ? 0 : pulse_cache(band, LM)[pulses] + 1;
}

CB_DEV void init_caps(int *cap, int LM, int C) {
    CB_NOUNROLL for (int i = 0; i < kNbEBands; i++) {
        int N = band_width(i) << LM;
        cap[i] = (kCacheCaps[kNbEBands * (2 * LM + C - 1) + i] + 64) * C * N >> 2;
    }
}

// Symbol hooks: the decoder reads the decision, the encoder codes the one it was given.
struct AllocDecIo {
    EcDec &ec;
    CB_MEM int skip_flag(int /*band_bits*/, int /*j*/, int /*codedBands*/, int /*band_width*/) { return ec.bit_logp(1); }
    CB_MEM int intensity(int /*want*/, int start, int codedBands) { return start + (int)ec.uint_((unsigned)(codedBands + 1 - start)); }
    CB_MEM int dual_stereo(int /*want*/) { return ec.bit_logp(1); }
};

// Scratch for one allocation: 4 x 21 ints.
struct AllocScratch {
    int bits1[kNbEBands], bits2[kNbEBands], thresh[kNbEBands], trim_offset[kNbEBands];
};

// compute_allocation (rate.c:527-638) with interp_bits2pulses (rate.c:248-525) folded in.
// Returns codedBands.  `pulses` = PVQ bits per band (1/8 bit), `ebits` = fine energy bits per channel.
template <class Io>
CB_DEV int compute_allocation(Io io, AllocScratch &sc, int start, int end, const int *offsets, const int *cap, int alloc_trim,
                              int *intensity, int *dual_stereo, int total, int *balance_out, int *pulses, int *ebits,
                              int *fine_priority, int C, int LM) {
    const int len = kNbEBands;
    int *bits1 = sc.bits1, *bits2 = sc.bits2, *thresh = sc.thresh, *trim_offset = sc.trim_offset;
    total = imax(total, 0);
    int skip_start = start;
    int skip_rsv = total >= 1 << kBitRes ? 1 << kBitRes : 0;
    total -= skip_rsv;
    int intensity_rsv = 0, dual_stereo_rsv = 0;
    if (C == 2) {
        intensity_rsv = kLog2Frac[end - start];
        if (intensity_rsv > total) intensity_rsv = 0;
        else {
            total -= intensity_rsv;
            dual_stereo_rsv = total >= 1 << kBitRes ? 1 << kBitRes : 0;
            total -= dual_stereo_rsv;
        }
    }
    CB_NOUNROLL for (int j = start; j < end; j++) {
        int w = band_width(j);
        thresh[j] = imax(C << kBitRes, (3 * w << LM << kBitRes) >> 4);
        trim_offset[j] = C * w * (alloc_trim - 5 - LM) * (end - j - 1) * (1 << (LM + kBitRes)) >> 6;
        if (w << LM == 1) trim_offset[j] -= C << kBitRes;
    }
    int lo = 1, hi = kNbAllocVectors - 1;
    do {
        int done = 0, psum = 0;
        int mid = (lo + hi) >> 1;
        CB_NOUNROLL for (int j = end; j-- > start;) {
            int bitsj = C * band_width(j) * kAllocVectors[mid * len + j] << LM >> 2;
            if (bitsj > 0) bitsj = imax(0, bitsj + trim_offset[j]);
            bitsj += offsets[j];
            if (bitsj >= thresh[j] || done) {
                done = 1;
                psum += imin(bitsj, cap[j]);
            } else if (bitsj >= C << kBitRes) {
                psum += C << kBitRes;
            }
        }
        if (psum > total) hi = mid - 1;
        else lo = mid + 1;
    } while (lo <= hi);
    hi = lo--;
    CB_NOUNROLL for (int j = start; j < end; j++) {
        int w = band_width(j);
        int b1 = C * w * kAllocVectors[lo * len + j] << LM >> 2;
        int b2 = hi >= kNbAllocVectors ? cap[j] : C * w * kAllocVectors[hi * len + j] << LM >> 2;
        if (b1 > 0) b1 = imax(0, b1 + trim_offset[j]);
        if (b2 > 0) b2 = imax(0, b2 + trim_offset[j]);
        if (lo > 0) b1 += offsets[j];
        b2 += offsets[j];
        if (offsets[j] > 0) skip_start = j;
        b2 = imax(0, b2 - b1);
        bits1[j] = b1;
        bits2[j] = b2;
    }

    // ---- interp_bits2pulses ------------------------------------------------------------------
    int *bits = pulses;
    const int alloc_floor = C << kBitRes;
    const int stereo = C > 1;
    const int logM = LM << kBitRes;
    int psum;
    lo = 0;
    hi = 1 << kAllocSteps;
    CB_NOUNROLL for (int i = 0; i < kAllocSteps; i++) {
        int mid = (lo + hi) >> 1;
        int done = 0;
        psum = 0;
        CB_NOUNROLL for (int j = end; j-- > start;) {
            int tmp = bits1[j] + (mid * bits2[j] >> kAllocSteps);
            if (tmp >= thresh[j] || done) {
                done = 1;
                psum += imin(tmp, cap[j]);
            } else if (tmp >= alloc_floor) {
                psum += alloc_floor;
            }
        }
        if (psum > total) hi = mid;
        else lo = mid;
    }
    psum = 0;
    {
        int done = 0;
        CB_NOUNROLL for (int j = end; j-- > start;) {
            int tmp = bits1[j] + (lo * bits2[j] >> kAllocSteps);
            if (tmp < thresh[j] && !done) {
                tmp = tmp >= alloc_floor ? alloc_floor : 0;
            } else {
                done = 1;
            }
            tmp = imin(tmp, cap[j]);
            bits[j] = tmp;
            psum += tmp;
        }
    }
    // band skipping, from the top
    int codedBands;
    CB_NOUNROLL for (codedBands = end;; codedBands--) {
        int j = codedBands - 1;
        if (j <= skip_start) {
            total += skip_rsv;
            break;
        }
        int left = total - psum;
        int span = kEBands[codedBands] - kEBands[start];
        int percoeff = (int)udiv((unsigned)left, (unsigned)span);
        left -= span * percoeff;
        int rem = imax(left - (kEBands[j] - kEBands[start]), 0);
        int bw = kEBands[codedBands] - kEBands[j];
        int band_bits = bits[j] + percoeff * bw + rem;
        if (band_bits >= imax(thresh[j], alloc_floor + (1 << kBitRes))) {
            if (io.skip_flag(band_bits, j, codedBands, bw)) break;
            psum += 1 << kBitRes;
            band_bits -= 1 << kBitRes;
        }
        psum -= bits[j] + intensity_rsv;
        if (intensity_rsv > 0) intensity_rsv = kLog2Frac[j - start];
        psum += intensity_rsv;
        if (band_bits >= alloc_floor) {
            psum += alloc_floor;
            bits[j] = alloc_floor;
        } else {
            bits[j] = 0;
        }
    }
    if (intensity_rsv > 0) *intensity = io.intensity(*intensity, start, codedBands);
    else *intensity = 0;
    if (*intensity <= start) {
        total += dual_stereo_rsv;
        dual_stereo_rsv = 0;
    }
    if (dual_stereo_rsv > 0) *dual_stereo = io.dual_stereo(*dual_stereo);
    else *dual_stereo = 0;

    // distribute what is left
    {
        int left = total - psum;
        int span = kEBands[codedBands] - kEBands[start];
        int percoeff = (int)udiv((unsigned)left, (unsigned)span);
        left -= span * percoeff;
        CB_NOUNROLL for (int j = start; j < codedBands; j++) bits[j] += percoeff * band_width(j);
        CB_NOUNROLL for (int j = start; j < codedBands; j++) {
            int tmp = imin(left, band_width(j));
            bits[j] += tmp;
            left -= tmp;
        }
    }
    int balance = 0;
    int j;
    CB_NOUNROLL for (j = start; j < codedBands; j++) {
        int N0 = band_width(j);
        int N = N0 << LM;
        int bit = bits[j] + balance;
        int excess;
        if (N > 1) {
            excess = imax(bit - cap[j], 0);
            bits[j] = bit - excess;
            int den = C * N + ((C == 2 && N > 2 && !*dual_stereo && j < *intensity) ? 1 : 0);
            int NClogN = den * (kLogN[j] + logM);
            int offset = (NClogN >> 1) - den * kFineOffset;
            if (N == 2) offset += den << kBitRes >> 2;
            if (bits[j] + offset < den * 2 << kBitRes) offset += NClogN >> 2;
            else if (bits[j] + offset < den * 3 << kBitRes) offset += NClogN >> 3;
            ebits[j] = imax(0, bits[j] + offset + (den << (kBitRes - 1)));
            ebits[j] = (int)udiv((unsigned)ebits[j], (unsigned)den) >> kBitRes;
            if (C * ebits[j] > (bits[j] >> kBitRes)) ebits[j] = bits[j] >> stereo >> kBitRes;
            ebits[j] = imin(ebits[j], kMaxFineBits);
            fine_priority[j] = ebits[j] * (den << kBitRes) >= bits[j] + offset;
            bits[j] -= C * ebits[j] << kBitRes;
        } else {
            excess = imax(0, bit - (C << kBitRes));
            bits[j] = bit - excess;
            ebits[j] = 0;
            fine_priority[j] = 1;
        }
        if (excess > 0) {
            int extra_fine = imin(excess >> (stereo + kBitRes), kMaxFineBits - ebits[j]);
            ebits[j] += extra_fine;
            int extra_bits = extra_fine * C << kBitRes;
            fine_priority[j] = extra_bits >= excess - balance;
            excess -= extra_bits;
        }
        balance = excess;
    }
    *balance_out = balance;
    CB_NOUNROLL for (; j < end; j++) {
        ebits[j] = bits[j] >> stereo >> kBitRes;
        bits[j] = 0;
        fine_priority[j] = ebits[j] < 1;
    }
    return codedBands;
}

}  // namespace cb
