// celt_enc_energy.cuh — band-energy quantisation, encoder side, plus tf_encode.
//
// Restates opus-fix/celt/quant_bands.c:144-433 (loss_distortion, quant_coarse_energy_impl, quant_coarse_energy with its
// intra/inter two-pass trial and range-coder snapshot, quant_fine_energy, quant_energy_finalise), :551-572 (amp2Log2) and
// celt/celt_encoder.c:715-753 (tf_encode).  Scalar code.
#pragma once
#include "celt_ec.cuh"
#include "celt_tables.cuh"

namespace cb {

// quant_bands.c:551-572
CB_DEV void amp2Log2(int effEnd, int end, const int *bandE, int16_t *bandLogE, int C) {
    CB_NOUNROLL for (int c = 0; c < C; c++) {
        CB_NOUNROLL for (int i = 0; i < effEnd; i++)
            bandLogE[i + c * kNbEBands] = (int16_t)(celt_log2(shl32(bandE[i + c * kNbEBands], 2)) - shl16(kEMeans[i], 6));
        CB_NOUNROLL for (int i = effEnd; i < end; i++) bandLogE[c * kNbEBands + i] = -14336;   // -QCONST16(14.f,DB_SHIFT)
    }
}

// quant_bands.c:144-156
CB_DEV int loss_distortion(const int16_t *eBands, const int16_t *oldEBands, int start, int end, int C) {
    int dist = 0;
    CB_NOUNROLL for (int c = 0; c < C; c++)
        CB_NOUNROLL for (int i = start; i < end; i++) {
            int d = (eBands[i + c * kNbEBands] >> 3) - (oldEBands[i + c * kNbEBands] >> 3);
            dist = mac16_16(dist, d, d);
        }
    return imin(200, dist >> 14);
}

// quant_bands.c:158-267
CB_DEV_NOINLINE int quant_coarse_energy_impl(int start, int end, const int16_t *eBands, int16_t *oldEBands, int budget, int tell,
                                             const uint8_t *prob_model, int16_t *error, EcEnc &enc, int C, int LM, int intra,
                                             int max_decay) {
    int badness = 0;
    int prev[2] = {0, 0};
    int coef, beta;
    if (tell + 3 <= budget) enc.bit_logp(intra, 3);
    if (intra) { coef = 0; beta = kBetaIntra; }
    else { beta = kBetaCoef[LM]; coef = kPredCoef[LM]; }
    CB_NOUNROLL for (int i = start; i < end; i++) {
        CB_NOUNROLL for (int c = 0; c < C; c++) {
            const int x = eBands[i + c * kNbEBands];
            const int oldE = imax(-9216, (int)oldEBands[i + c * kNbEBands]);
            const int f = wsub(wsub(shl32(x, 7), pshr32(mul16_16(coef, oldE), 8)), prev[c]);
            int qi = wadd(f, 65536) >> 17;   // QCONST32(.5f,DB_SHIFT+7)
            const int decay_bound = s16(imax(-28672, (int)oldEBands[i + c * kNbEBands] - max_decay));
            if (qi < 0 && x < decay_bound) {
                qi += (decay_bound - x) >> 10;
                if (qi > 0) qi = 0;
            }
            const int qi0 = qi;
            tell = enc.tell();
            const int bits_left = budget - tell - 3 * C * (end - i);
            if (i != start && bits_left < 30) {
                if (bits_left < 24) qi = imin(1, qi);
                if (bits_left < 16) qi = imax(-1, qi);
            }
            if (budget - tell >= 15) {
                int pi = 2 * imin(i, 20);
                enc.laplace(&qi, (unsigned)prob_model[pi] << 7, (int)prob_model[pi + 1] << 6);
            } else if (budget - tell >= 2) {
                qi = imax(-1, imin(qi, 1));
                enc.icdf(2 * qi ^ -(qi < 0), kSmallEnergyIcdf, 2);
            } else if (budget - tell >= 1) {
                qi = imin(0, qi);
                enc.bit_logp(-qi, 1);
            } else {
                qi = -1;
            }
            error[i + c * kNbEBands] = (int16_t)(pshr32(f, 7) - shl16(qi, 10));
            badness += iabs(qi0 - qi);
            const int q = shl32(qi, 10);
            int tmp = wadd(wadd(pshr32(mul16_16(coef, oldE), 8), prev[c]), shl32(q, 7));
            tmp = imax(-3670016, tmp);
            oldEBands[i + c * kNbEBands] = (int16_t)pshr32(tmp, 7);
            prev[c] = wsub(wadd(prev[c], shl32(q, 7)), mul16_16(beta, pshr32(q, 8)));
        }
    }
    return badness;
}

// Scratch of the two-pass trial: intra copies of the energies / errors and the bytes the intra pass produced.
struct CoarseScratch {
    int16_t oldE_intra[2 * kNbEBands], error_intra[2 * kNbEBands];
    uint8_t intra_bits[160];   // 42 Laplace symbols of <= 15+ bits each cannot exceed this
};

// quant_bands.c:269-367
CB_DEV_NOINLINE void quant_coarse_energy(int start, int end, int effEnd, const int16_t *eBands, int16_t *oldEBands, unsigned budget,
                                         int16_t *error, EcEnc &enc, int C, int LM, int nbAvailableBytes, int force_intra,
                                         int *delayedIntra, int two_pass, int loss_rate, CoarseScratch &cs) {
    int intra = force_intra || (!two_pass && *delayedIntra > 2 * C * (end - start) && nbAvailableBytes > (end - start) * C);
    const int intra_bias = (int)((budget * (unsigned)*delayedIntra * (unsigned)loss_rate) / (unsigned)(C * 512));
    const int new_distortion = loss_distortion(eBands, oldEBands, start, effEnd, C);
    const unsigned tell = (unsigned)enc.tell();
    if (tell + 3 > budget) two_pass = intra = 0;
    int max_decay = 16384;   // QCONST16(16.f,DB_SHIFT)
    if (end - start > 10) max_decay = s16(imin(max_decay, shl32(nbAvailableBytes, 7)));
    const EcEnc enc_start = enc;
    CB_NOUNROLL for (int i = 0; i < C * kNbEBands; i++) cs.oldE_intra[i] = oldEBands[i];
    int badness1 = 0;
    if (two_pass || intra)
        badness1 = quant_coarse_energy_impl(start, end, eBands, cs.oldE_intra, (int)budget, (int)tell, kEProbModel[LM][1],
                                            cs.error_intra, enc, C, LM, 1, max_decay);
    if (!intra) {
        const int tell_intra = (int)enc.tell_frac();
        const EcEnc enc_intra = enc;
        const unsigned nstart = enc_start.offs, nintra = enc_intra.offs;
        uint8_t *intra_buf = enc_intra.buf + nstart;
        unsigned save = nintra - nstart;
        if (save > sizeof(cs.intra_bits)) save = sizeof(cs.intra_bits);
        CB_NOUNROLL for (unsigned k = 0; k < save; k++) cs.intra_bits[k] = intra_buf[k];
        enc = enc_start;
        const int badness2 = quant_coarse_energy_impl(start, end, eBands, oldEBands, (int)budget, (int)tell, kEProbModel[LM][intra],
                                                      error, enc, C, LM, 0, max_decay);
        if (two_pass && (badness1 < badness2 || (badness1 == badness2 && (int)enc.tell_frac() + intra_bias > tell_intra))) {
            enc = enc_intra;
            CB_NOUNROLL for (unsigned k = 0; k < save; k++) intra_buf[k] = cs.intra_bits[k];
            CB_NOUNROLL for (int i = 0; i < C * kNbEBands; i++) { oldEBands[i] = cs.oldE_intra[i]; error[i] = cs.error_intra[i]; }
            intra = 1;
        }
    } else {
        CB_NOUNROLL for (int i = 0; i < C * kNbEBands; i++) { oldEBands[i] = cs.oldE_intra[i]; error[i] = cs.error_intra[i]; }
    }
    if (intra) *delayedIntra = new_distortion;
    else *delayedIntra = wadd(mul16_32_q15(mul16_16_q15(kPredCoef[LM], kPredCoef[LM]), *delayedIntra), new_distortion);
}

// quant_bands.c:369-404
CB_DEV void quant_fine_energy(int start, int end, int16_t *oldEBands, int16_t *error, const int *fine_quant, EcEnc &enc, int C) {
    CB_NOUNROLL for (int i = start; i < end; i++) {
        const int frac = s16(1 << fine_quant[i]);
        if (fine_quant[i] <= 0) continue;
        CB_NOUNROLL for (int c = 0; c < C; c++) {
            int q2 = (error[i + c * kNbEBands] + 512) >> (10 - fine_quant[i]);
            if (q2 > frac - 1) q2 = frac - 1;
            if (q2 < 0) q2 = 0;
            enc.bits((unsigned)q2, (unsigned)fine_quant[i]);
            int offset = s16(s16((shl32(q2, 10) + 512) >> fine_quant[i]) - 512);
            oldEBands[i + c * kNbEBands] = (int16_t)(oldEBands[i + c * kNbEBands] + offset);
            error[i + c * kNbEBands] = (int16_t)(error[i + c * kNbEBands] - offset);
        }
    }
}

// quant_bands.c:406-433
CB_DEV void quant_energy_finalise(int start, int end, int16_t *oldEBands, const int16_t *error, const int *fine_quant,
                                  const int *fine_priority, int bits_left, EcEnc &enc, int C) {
    CB_NOUNROLL for (int prio = 0; prio < 2; prio++) {
        CB_NOUNROLL for (int i = start; i < end && bits_left >= C; i++) {
            if (fine_quant[i] >= kMaxFineBits || fine_priority[i] != prio) continue;
            CB_NOUNROLL for (int c = 0; c < C; c++) {
                int q2 = error[i + c * kNbEBands] < 0 ? 0 : 1;
                enc.bits((unsigned)q2, 1);
                int offset = s16((shl16(q2, 10) - 512) >> (fine_quant[i] + 1));
                oldEBands[i + c * kNbEBands] = (int16_t)(oldEBands[i + c * kNbEBands] + offset);
                bits_left--;
            }
        }
    }
}

// celt_encoder.c:715-753
CB_DEV void tf_encode(int start, int end, int isTransient, int *tf_res, int LM, int tf_select, EcEnc &enc) {
    unsigned budget = enc.storage * 8;
    unsigned tell = (unsigned)enc.tell();
    int logp = isTransient ? 2 : 4;
    int tf_select_rsv = LM > 0 && tell + logp + 1 <= budget;
    budget -= tf_select_rsv;
    int curr = 0, tf_changed = 0;
    CB_NOUNROLL for (int i = start; i < end; i++) {
        if (tell + logp <= budget) {
            enc.bit_logp(tf_res[i] ^ curr, logp);
            tell = (unsigned)enc.tell();
            curr = tf_res[i];
            tf_changed |= curr;
        } else {
            tf_res[i] = curr;
        }
        logp = isTransient ? 4 : 5;
    }
    if (tf_select_rsv && kTfSelect[LM][4 * isTransient + 0 + tf_changed] != kTfSelect[LM][4 * isTransient + 2 + tf_changed])
        enc.bit_logp(tf_select, 1);
    else
        tf_select = 0;
    CB_NOUNROLL for (int i = start; i < end; i++) tf_res[i] = kTfSelect[LM][4 * isTransient + 2 * tf_select + tf_res[i]];
}

}  // namespace cb
