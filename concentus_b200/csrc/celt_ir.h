// celt_ir.h — the intermediate representation between the two decoder stages.
//
// Stage A (parse, one thread per run of packets) turns a packet into: one CbPacketIR, up to K CbFrameIR, and the
// normalised spectra X of its frames (int16, channel-major per frame, consecutive inside the packet's X area).
// Stage B (synth, one warp per stream) consumes them in stream order and owns all persistent state.
// What crosses the boundary is exactly what celt_decode_with_ec (opus-fix/celt/celt_decoder.c:713-1072) computes from
// the bitstream alone — the parse of a CELT frame never reads decoder state except the fold/noise seed (= the range
// coder's final rng of the previous frame), which stage A chains itself.
#pragma once
#include <stdint.h>

#define CB_IR_SILENCE 0x01
#define CB_IR_TRANSIENT 0x02
#define CB_IR_INTRA 0x04
#define CB_IR_ANTICOLLAPSE 0x08
#define CB_IR_EC_ERROR 0x10     /* ec_dec error latch (celt_decoder.c:1069) */
#define CB_IR_OVERRUN 0x20      /* ec_tell > 8*len: OPUS_INTERNAL_ERROR after the frame is synthesised (celt_decoder.c:1067) */
#define CB_IR_LOST 0x40         /* payload <= 1 byte: concealment frame (opus_decoder.c:246-252) */

typedef struct CbFrameIR {
    uint32_t rng_final;          /* -> st->rng and OPUS_GET_FINAL_RANGE */
    uint32_t seed_bands;         /* LCG seed after the band loop (anti-collapse noise, celt_decoder.c:989) */
    int32_t x_off;               /* int16 offset of this frame's X inside the packet's X area */
    int16_t len;                 /* payload bytes */
    int16_t pf_pitch, pf_gain;   /* post-filter of this frame (gain Q15) */
    uint8_t pf_tapset;
    uint8_t LM, C, end;          /* frame size code, coded channels, end band */
    uint8_t flags;
    uint8_t pad[3];
    int16_t qi[42];              /* coarse energy symbols, [c*21+band] */
    int16_t eoff[42];            /* fine + finalise energy offsets, summed (Q10, wrapping) */
    int16_t pulses[21];          /* PVQ bit allocation per band (1/8 bit): anti-collapse depth */
    uint8_t collapse[42];        /* collapse masks [band*C+c] */
} CbFrameIR;

typedef struct CbPacketIR {
    int32_t ret;                 /* < 0: error for this packet (state must stay untouched) */
    int16_t count;               /* CELT frames that follow (0 when the whole packet is lost) */
    int16_t lost;                /* 1: run concealment for `cap` samples (opus_decoder.c:613-627) */
    int32_t frame_size;          /* packet frame size in samples at the API rate */
    int16_t mode, bandwidth;
    int16_t stream_channels, reserved;
} CbPacketIR;
