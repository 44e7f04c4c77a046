// celt_tables.cuh — read-only constants of the 48 kHz standard CELT mode (mode48000_960_120,
// opus-fix/celt/static_modes_fixed.h:870-887) and of the entropy models.
//
// The large numeric arrays live in celt_tables_data.inc (generated, see tools/extract_tables.c); the small
// normative tables are written out here with their reference location.  On the GPU everything is a
// `__device__ const` array: ~13 KB in total, L1/L2-resident, indexed with lane-varying subscripts (so
// __constant__ memory, which serialises divergent indices, is the wrong place for most of them).
#pragma once
#include "celt_simt.cuh"

namespace cb {

#include "celt_tables_data.inc"

enum {
    kNbEBands = 21,
    kOverlap = 120,
    kShortMdct = 120,
    kMaxLM = 3,
    kMaxFrame = 960,
    kDecBuf = 2048,          // DECODE_BUFFER_SIZE (celt/celt_decoder.c:62)
    kCombMinPeriod = 15,     // celt/celt.h:192-193
    kCombMaxPeriod = 1024,
    kBitRes = 3,             // BITRES (celt/entcode.h:58)
    kMaxFineBits = 8,        // celt/rate.h:36-39
    kFineOffset = 21,
    kQThetaOffset = 4,
    kQThetaOffsetTwoPhase = 16,
    kLogMaxPseudo = 6,
    kAllocSteps = 6,
    kNbAllocVectors = 11,
    kLpcOrder = 24,
    kPreemphCoef0 = 27853,   // mode->preemph[0]
};
enum { kSpreadNone = 0, kSpreadLight = 1, kSpreadNormal = 2, kSpreadAggressive = 3 };

// eMeans (celt/quant_bands.c:46-52), Q4
CB_TABLE int8_t kEMeans[25] = {103, 100, 92, 85, 81, 77, 72, 70, 78, 75, 73, 71, 78, 74, 69, 72, 70, 74, 76, 71, 60, 60, 60, 60, 60};
// pred_coef / beta_coef / beta_intra (celt/quant_bands.c:64-66)
CB_TABLE int16_t kPredCoef[4] = {29440, 26112, 21248, 16384};
CB_TABLE int16_t kBetaCoef[4] = {30147, 22282, 12124, 6554};
enum { kBetaIntra = 4915 };
// e_prob_model[LM][intra][2*band] (celt/quant_bands.c:77-140): {P(0), decay} in Q8
CB_TABLE uint8_t kEProbModel[4][2][42] = {
    {{72, 127, 65, 129, 66, 128, 65, 128, 64, 128, 62, 128, 64, 128, 64, 128, 92, 78, 92, 79, 92, 78, 90, 79, 116, 41, 115, 40, 114, 40, 132, 26, 132, 26, 145, 17, 161, 12, 176, 10, 177, 11},
     {24, 179, 48, 138, 54, 135, 54, 132, 53, 134, 56, 133, 55, 132, 55, 132, 61, 114, 70, 96, 74, 88, 75, 88, 87, 74, 89, 66, 91, 67, 100, 59, 108, 50, 120, 40, 122, 37, 97, 43, 78, 50}},
    {{83, 78, 84, 81, 88, 75, 86, 74, 87, 71, 90, 73, 93, 74, 93, 74, 109, 40, 114, 36, 117, 34, 117, 34, 143, 17, 145, 18, 146, 19, 162, 12, 165, 10, 178, 7, 189, 6, 190, 8, 177, 9},
     {23, 178, 54, 115, 63, 102, 66, 98, 69, 99, 74, 89, 71, 91, 73, 91, 78, 89, 86, 80, 92, 66, 93, 64, 102, 59, 103, 60, 104, 60, 117, 52, 123, 44, 138, 35, 133, 31, 97, 38, 77, 45}},
    {{61, 90, 93, 60, 105, 42, 107, 41, 110, 45, 116, 38, 113, 38, 112, 38, 124, 26, 132, 27, 136, 19, 140, 20, 155, 14, 159, 16, 158, 18, 170, 13, 177, 10, 187, 8, 192, 6, 175, 9, 159, 10},
     {21, 178, 59, 110, 71, 86, 75, 85, 84, 83, 91, 66, 88, 73, 87, 72, 92, 75, 98, 72, 105, 58, 107, 54, 115, 52, 114, 55, 112, 56, 129, 51, 132, 40, 150, 33, 140, 29, 98, 35, 77, 42}},
    {{42, 121, 96, 66, 108, 43, 111, 40, 117, 44, 123, 32, 120, 36, 119, 33, 127, 33, 134, 34, 139, 21, 147, 23, 152, 20, 158, 25, 154, 26, 166, 21, 173, 16, 184, 13, 184, 10, 150, 13, 139, 15},
     {22, 178, 63, 114, 74, 82, 84, 83, 92, 82, 103, 62, 96, 72, 96, 67, 101, 73, 107, 72, 113, 55, 118, 52, 125, 52, 118, 52, 117, 55, 135, 49, 137, 39, 157, 32, 145, 29, 97, 33, 77, 40}},
};
CB_TABLE uint8_t kSmallEnergyIcdf[3] = {2, 1, 0};                                   // quant_bands.c:142
CB_TABLE uint8_t kTrimIcdf[11] = {126, 124, 119, 109, 87, 41, 19, 9, 4, 2, 0};        // celt/celt.h:150
CB_TABLE uint8_t kSpreadIcdf[4] = {25, 23, 2, 0};                                     // celt/celt.h:152
CB_TABLE uint8_t kTapsetIcdf[3] = {2, 1, 0};                                          // celt/celt.h:154
CB_TABLE int8_t kTfSelect[4][8] = {                                                   // celt/celt.c:247-252
    {0, -1, 0, -1, 0, -1, 0, -1}, {0, -1, 0, -2, 1, 0, 1, -1}, {0, -2, 0, -3, 2, 0, 1, -1}, {0, -2, 0, -3, 3, 0, 1, -1}};
CB_TABLE uint8_t kLog2Frac[24] = {0, 8, 13, 16, 19, 21, 23, 24, 26, 27, 28, 29, 30, 31, 32, 32, 33, 34, 34, 35, 36, 36, 37, 37};  // celt/rate.c:42-48
// comb-filter tap gains, Q15 (celt/celt.c:191-194): QCONST16 of {.3066406250,.2170410156,.1296386719},{.4638671875,.2680664062,0},{.7998046875,.1000976562,0}
CB_TABLE int16_t kCombGains[3][3] = {{10048, 7112, 4248}, {15200, 8784, 0}, {26208, 3280, 0}};
// Hadamard ordering (celt/bands.c:525-530)
CB_TABLE uint8_t kOrdery[30] = {1, 0, 3, 0, 2, 1, 7, 0, 4, 3, 6, 1, 5, 2, 15, 0, 8, 7, 12, 3, 11, 4, 14, 1, 9, 6, 13, 2, 10, 5};
CB_TABLE uint8_t kBitInterleave[16] = {0, 1, 1, 1, 2, 3, 3, 3, 2, 3, 3, 3, 2, 3, 3, 3};                       // bands.c:1091
CB_TABLE uint8_t kBitDeinterleave[16] = {0x00, 0x03, 0x0C, 0x0F, 0x30, 0x33, 0x3C, 0x3F, 0xC0, 0xC3, 0xCC, 0xCF, 0xF0, 0xF3, 0xFC, 0xFF};  // bands.c:1151
CB_TABLE int16_t kExp2Table8[8] = {16384, 17866, 19483, 21247, 23170, 25267, 27554, 30048};                    // bands.c:598

// FFT plans (celt/static_modes_fixed.h:432-499): nfft = 480>>s, twiddle stride shift = max(shift,0).
// Stage list is in EXECUTION order (last factor first), each {radix, m}; see celt_mdct.cuh.
struct FftPlan {
    int16_t nfft;
    int8_t scale_shift;   // 8,7,6,5
    int8_t tw_shift;      // 0,1,2,3
    int8_t nstages;
    int8_t radix[5];
    int16_t m[5];
};
CB_TABLE FftPlan kFftPlan[4] = {
    {480, 8, 0, 5, {4, 2, 4, 3, 5}, {1, 4, 8, 32, 96}},
    {240, 7, 1, 4, {4, 4, 3, 5, 0}, {1, 4, 16, 48, 0}},
    {120, 6, 2, 4, {4, 2, 3, 5, 0}, {1, 4, 8, 24, 0}},
    {60, 5, 3, 3, {4, 3, 5, 0, 0}, {1, 4, 12, 0, 0}},
};
enum { kFftScale = 17476 };

enum {
    OPUS_OK_ = 0, OPUS_BAD_ARG_ = -1, OPUS_BUFFER_TOO_SMALL_ = -2, OPUS_INTERNAL_ERROR_ = -3,
    OPUS_INVALID_PACKET_ = -4, OPUS_UNIMPLEMENTED_ = -5, OPUS_INVALID_STATE_ = -6, OPUS_ALLOC_FAIL_ = -7,
};
enum { kBwNarrow = 1101, kBwMedium = 1102, kBwWide = 1103, kBwSuperWide = 1104, kBwFull = 1105 };

// bin (at LM=0 resolution) -> band, from eband5ms
CB_TABLE uint8_t kBinToBand[100] = {
    0, 1, 2, 3, 4, 5, 6, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 12, 12, 13, 13, 13, 13, 14, 14, 14, 14,
    15, 15, 15, 15, 15, 15, 16, 16, 16, 16, 16, 16, 17, 17, 17, 17, 17, 17, 17, 17,
    18, 18, 18, 18, 18, 18, 18, 18, 18, 18, 18, 18,
    19, 19, 19, 19, 19, 19, 19, 19, 19, 19, 19, 19, 19, 19, 19, 19, 19, 19,
    20, 20, 20, 20, 20, 20, 20, 20, 20, 20, 20, 20, 20, 20, 20, 20, 20, 20, 20, 20, 20, 20};

// intensity_thresholds / intensity_histeresis (celt/celt_encoder.c:1973-1977)
CB_TABLE int16_t kIntensityThresholds[21] = {1, 2, 3, 4, 5, 6, 7, 8, 16, 24, 36, 44, 50, 56, 62, 67, 72, 79, 88, 106, 134};
CB_TABLE int16_t kIntensityHisteresis[21] = {1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 3, 3, 4, 5, 6, 8, 8};

CB_DEV const int16_t *fft_bitrev(int s) {
    return s == 0 ? kFftBitrev480 : s == 1 ? kFftBitrev240 : s == 2 ? kFftBitrev120 : kFftBitrev60;
}

}  // namespace cb
