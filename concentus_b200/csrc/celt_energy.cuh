// celt_energy.cuh — band-energy (de)quantisation, decoder side, plus tf_decode.
//
// Restates opus-fix/celt/quant_bands.c:435-549 (unquant_coarse_energy, unquant_fine_energy,
// unquant_energy_finalise) and celt/celt_decoder.c:352-389 (tf_decode), split along the pipeline:
//   stage A (parse)  reads the SYMBOLS: coarse qi per band/channel, and the fine + finalise offsets summed
//                    per band/channel (they are wrapping int16 additions onto the same cell with no read in
//                    between, so their order is free and one sum carries them);
//   stage B (synth)  owns the state: applies the inter/intra prediction recurrence with qi, then the offsets.
// Energies are Q10 int16 (DB_SHIFT = 10).
#pragma once
#include "celt_ec.cuh"
#include "celt_tables.cuh"

namespace cb {

// symbol half of unquant_coarse_energy (quant_bands.c:457-482)
CB_DEV void decode_coarse_symbols(int start, int end, int intra, EcDec &dec, int C, int LM, int16_t *qi_out) {
    const uint8_t *prob = kEProbModel[LM][intra];
    const int budget = (int)dec.storage * 8;
    CB_NOUNROLL for (int i = start; i < end; i++) {
        CB_NOUNROLL for (int c = 0; c < C; c++) {
            int qi;
            int tell = dec.tell();
            if (budget - tell >= 15) {
                int pi = 2 * imin(i, 20);
                qi = dec.laplace((unsigned)prob[pi] << 7, (int)prob[pi + 1] << 6);
            } else if (budget - tell >= 2) {
                qi = dec.icdf(kSmallEnergyIcdf, 2);
                qi = (qi >> 1) ^ -(qi & 1);
            } else if (budget - tell >= 1) {
                qi = -dec.bit_logp(1);
            } else {
                qi = -1;
            }
            qi_out[i + c * kNbEBands] = (int16_t)qi;
        }
    }
}

// state half of unquant_coarse_energy (quant_bands.c:437-455,483-497).  qi is an `int qi` in the reference:
// Laplace values are bounded well inside int16 (|qi| <= 32768 >> 1 steps of the pdf), so int16 transport is exact.
CB_DEV void apply_coarse_energy(int start, int end, int16_t *oldE, const int16_t *qi_in, int intra, int C, int LM) {
    int prev[2] = {0, 0};
    int coef, beta;
    if (intra) { coef = 0; beta = kBetaIntra; }
    else { beta = kBetaCoef[LM]; coef = kPredCoef[LM]; }
    CB_NOUNROLL for (int i = start; i < end; i++) {
        CB_NOUNROLL for (int c = 0; c < C; c++) {
            int q = shl32((int)qi_in[i + c * kNbEBands], 10);
            int16_t *e = &oldE[i + c * kNbEBands];
            int old = imax(-9216, (int)*e);   // -QCONST16(9.f, DB_SHIFT)
            int tmp = wadd(wadd(pshr32(mul16_16(coef, old), 8), prev[c]), shl32(q, 7));
            tmp = imax(-3670016, tmp);        // -QCONST32(28.f, DB_SHIFT+7)
            *e = (int16_t)pshr32(tmp, 7);
            prev[c] = wsub(wadd(prev[c], shl32(q, 7)), mul16_16(beta, pshr32(q, 8)));
        }
    }
}

// unquant_fine_energy (quant_bands.c:500-521): offsets accumulated into eoff
CB_DEV void decode_fine_energy(int start, int end, const int *fine_quant, EcDec &dec, int C, int16_t *eoff) {
    CB_NOUNROLL for (int i = start; i < end; i++) {
        if (fine_quant[i] <= 0) continue;
        CB_NOUNROLL for (int c = 0; c < C; c++) {
            int q2 = (int)dec.bits((unsigned)fine_quant[i]);
            int offset = s16(s16((shl32(q2, 10) + 512) >> fine_quant[i]) - 512);
            eoff[i + c * kNbEBands] = (int16_t)(eoff[i + c * kNbEBands] + offset);
        }
    }
}

// unquant_energy_finalise (quant_bands.c:523-549)
CB_DEV void decode_energy_finalise(int start, int end, const int *fine_quant, const int *fine_priority, int bits_left, EcDec &dec,
                                   int C, int16_t *eoff) {
    CB_NOUNROLL for (int prio = 0; prio < 2; prio++) {
        CB_NOUNROLL for (int i = start; i < end && bits_left >= C; i++) {
            if (fine_quant[i] >= kMaxFineBits || fine_priority[i] != prio) continue;
            CB_NOUNROLL for (int c = 0; c < C; c++) {
                int q2 = (int)dec.bits(1);
                int offset = s16((shl16(q2, 10) - 512) >> (fine_quant[i] + 1));
                eoff[i + c * kNbEBands] = (int16_t)(eoff[i + c * kNbEBands] + offset);
                bits_left--;
            }
        }
    }
}

// celt_decoder.c:352-389
CB_DEV void tf_decode(int start, int end, int isTransient, int *tf_res, int LM, EcDec &dec) {
    unsigned budget = dec.storage * 8;
    unsigned tell = (unsigned)dec.tell();
    int logp = isTransient ? 2 : 4;
    int tf_select_rsv = LM > 0 && tell + logp + 1 <= budget;
    budget -= tf_select_rsv;
    int tf_changed = 0, curr = 0;
    CB_NOUNROLL for (int i = start; i < end; i++) {
        if (tell + logp <= budget) {
            curr ^= dec.bit_logp(logp);
            tell = (unsigned)dec.tell();
            tf_changed |= curr;
        }
        tf_res[i] = curr;
        logp = isTransient ? 4 : 5;
    }
    int tf_select = 0;
    if (tf_select_rsv && kTfSelect[LM][4 * isTransient + 0 + tf_changed] != kTfSelect[LM][4 * isTransient + 2 + tf_changed])
        tf_select = dec.bit_logp(1);
    CB_NOUNROLL for (int i = start; i < end; i++) tf_res[i] = kTfSelect[LM][4 * isTransient + 2 * tf_select + tf_res[i]];
}

}  // namespace cb
