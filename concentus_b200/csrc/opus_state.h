// opus_state.h — the per-stream persistent codec state as it lives in HBM (and, serialised, inside the
// caller-visible OpusDecoder / OpusEncoder blocks).  Plain C layout, no pointers: the block can be
// memcpy'd, relocated and snapshotted exactly like the reference's states
// (opus-fix/tests/test_opus_decode.c:84-95, src/opus_decoder.c:55-79, celt/celt_decoder.c:67-100).
#pragma once
#include <stddef.h>
#include <stdint.h>

#define CB_NB_EBANDS 21
#define CB_OVERLAP 120
#define CB_DEC_BUF 2048
#define CB_DEC_MEM (CB_DEC_BUF + CB_OVERLAP)   /* per channel */
#define CB_LPC_ORDER 24

#define CB_MODE_SILK_ONLY 1000
#define CB_MODE_HYBRID 1001
#define CB_MODE_CELT_ONLY 1002

typedef struct CbDecState {
    /* ---- configuration: survives OPUS_RESET_STATE (src/opus_decoder.c:55-63) ---- */
    int32_t channels;       /* output channels CC */
    int32_t Fs;             /* API sampling rate */
    int32_t downsample;     /* resampling_factor(Fs), celt/celt.c:62-91 */
    int32_t decode_gain;    /* Q8 dB */
    /* ---- OPUS_DECODER_RESET_START (src/opus_decoder.c:66-78) ---- */
    int32_t stream_channels;
    int32_t bandwidth;
    int32_t mode;
    int32_t prev_mode;
    int32_t frame_size;
    int32_t prev_redundancy;
    int32_t last_packet_duration;
    uint32_t rangeFinal;
    /* ---- CELT DECODER_RESET_START (celt/celt_decoder.c:80-99) ---- */
    uint32_t rng;
    int32_t error;
    int32_t last_pitch_index;
    int32_t loss_count;
    int32_t postfilter_period, postfilter_period_old;
    int32_t postfilter_gain, postfilter_gain_old;   /* Q15, value range of opus_val16 */
    int32_t postfilter_tapset, postfilter_tapset_old;
    int32_t preemph_memD[2];
    int16_t oldEBands[2 * CB_NB_EBANDS];
    int16_t oldLogE[2 * CB_NB_EBANDS];
    int16_t oldLogE2[2 * CB_NB_EBANDS];
    int16_t backgroundLogE[2 * CB_NB_EBANDS];
    int16_t lpc[2 * CB_LPC_ORDER];
    int32_t decode_mem[2 * CB_DEC_MEM];
} CbDecState;

#define CB_COMB_MAXPERIOD 1024
#define CB_ENC_DELAY_BUF 960   /* MAX_ENCODER_BUFFER (480) x 2 channels, src/opus_encoder.c:58 */

/* Encoder state: Opus layer (src/opus_encoder.c:62-110, CELT-only subset) + CELT layer (celt/celt_encoder.c:60-128).
 * Layout: everything small comes first (the "head": scalars and the three band-energy histories, CB_ENC_HEAD_BYTES) so a
 * kernel can keep it in shared memory for a whole span; the sample histories (the "tail") stay in HBM. */
typedef struct CbEncState {
    /* ---- Opus-layer configuration (survives OPUS_RESET_STATE) ---- */
    int32_t application, channels, Fs;
    int32_t force_channels, signal_type, user_bandwidth, max_bandwidth, user_forced_mode;
    int32_t use_vbr, vbr_constraint, variable_duration, user_bitrate_bps, lsb_depth;
    int32_t complexity, packet_loss_perc, prediction_disabled, inband_fec, dtx, delay_compensation, encoder_buffer;
    int32_t voice_ratio, bitrate_bps;
    /* ---- OPUS_ENCODER_RESET_START ---- */
    int32_t stream_channels;
    int32_t hybrid_stereo_width_Q14;
    int32_t hp_mem[4];
    int32_t mode, prev_mode, prev_channels, prev_framesize, bandwidth, first;
    int32_t width_XX, width_XY, width_YY, width_smoothed, width_max_follower;   /* StereoWidthState */
    uint32_t rangeFinal;
    /* ---- CELT configuration (celt_encoder.c:62-80) ---- */
    int32_t upsample, celt_force_intra, celt_disable_pf;
    /* ---- CELT ENCODER_RESET_START (celt_encoder.c:84-113) ---- */
    uint32_t rng;
    int32_t spread_decision, delayedIntra, tonal_average, lastCodedBands, hf_average, tapset_decision;
    int32_t prefilter_period, prefilter_gain, prefilter_tapset, consec_transient;
    int32_t preemph_memE[2];
    int32_t vbr_reservoir, vbr_drift, vbr_offset, vbr_count, overlap_max, stereo_saving, intensity, spec_avg;
    int16_t oldBandE[2 * CB_NB_EBANDS], oldLogE[2 * CB_NB_EBANDS], oldLogE2[2 * CB_NB_EBANDS];
    int32_t head_pad;                                    /* keeps the tail 8-byte aligned */
    /* ---- tail: sample histories (HBM) ---- */
    int32_t in_mem[2 * CB_OVERLAP];
    int32_t prefilter_mem[2 * CB_COMB_MAXPERIOD];
    int16_t delay_buffer[CB_ENC_DELAY_BUF];
} CbEncState;
#define CB_ENC_HEAD_BYTES ((int)offsetof(CbEncState, in_mem))
