// opus_decoder_dev.cuh — the Opus decoder layer between the C API and the CELT frame stages.
//
// Restates the MODE_CELT_ONLY subset of opus-fix/src/opus_decoder.c:200-596 (opus_decode_frame) and
// :598-709 (opus_decode_native), plus state init / reset (:96-133, celt/celt_decoder.c:134-172,1188-1202),
// split along the pipeline: opus_parse_packet (stage A: TOC, framing, per-frame CELT parse -> IR) and
// opus_synth_packet (stage B: state latching, per-frame synthesis, gain, final range).
// SILK-only and hybrid packets are outside this engine (SURVEY.md §8b "scope edge"): they return
// OPUS_UNIMPLEMENTED and leave the state untouched — there is deliberately no CPU fallback.
#pragma once
#include "celt_decoder.cuh"
#include "celt_plc.cuh"
#include "opus_packet.h"

namespace cb {

// OPUS_RESET_STATE of the CELT decoder (celt_decoder.c:1188-1202) — host or device, scalar.
CB_HD void dec_state_reset_celt(CbDecState *st) {
    st->rng = 0; st->error = 0; st->last_pitch_index = 0; st->loss_count = 0;
    st->postfilter_period = st->postfilter_period_old = 0;
    st->postfilter_gain = st->postfilter_gain_old = 0;
    st->postfilter_tapset = st->postfilter_tapset_old = 0;
    st->preemph_memD[0] = st->preemph_memD[1] = 0;
    CB_NOUNROLL for (int i = 0; i < 2 * CB_NB_EBANDS; i++) {
        st->oldEBands[i] = 0; st->backgroundLogE[i] = 0;
        st->oldLogE[i] = st->oldLogE2[i] = -28672;
    }
    CB_NOUNROLL for (int i = 0; i < 2 * CB_LPC_ORDER; i++) st->lpc[i] = 0;
    CB_NOUNROLL for (int i = 0; i < 2 * CB_DEC_MEM; i++) st->decode_mem[i] = 0;
}
// OPUS_RESET_STATE of the Opus decoder (opus_decoder.c:873-887)
CB_HD void dec_state_reset(CbDecState *st) {
    st->bandwidth = 0; st->mode = 0; st->prev_mode = 0; st->prev_redundancy = 0;
    st->last_packet_duration = 0; st->rangeFinal = 0;
    dec_state_reset_celt(st);
    st->stream_channels = st->channels;
    st->frame_size = st->Fs / 400;
}
// opus_decoder_init (opus_decoder.c:96-133)
CB_HD int dec_state_init(CbDecState *st, int Fs, int channels) {
    if ((Fs != 48000 && Fs != 24000 && Fs != 16000 && Fs != 12000 && Fs != 8000) || (channels != 1 && channels != 2)) return -1;
    st->channels = channels;
    st->Fs = Fs;
    st->downsample = 48000 / Fs;
    st->decode_gain = 0;
    dec_state_reset(st);
    return 0;
}

CB_DEV int bandwidth_to_endband(int bw) {   // opus_decoder.c:431-450
    switch (bw) {
    case kBwNarrow: return 13;
    case kBwMedium:
    case kBwWide: return 17;
    case kBwSuperWide: return 19;
    default: return 21;
    }
}

// ---------------------------------------------------------------------------------------------------
// Stage A: one packet -> IR.  `cap` = PCM capacity for this packet (samples per channel at Fs), kmax = frame IR
// slots per packet.  Nothing here reads or writes decoder state.
//
// What a frame's parse needs from the past is the fold/noise seed st->rng.  A received frame sets it to its final range;
// pitch-based concealment leaves it alone; NOISE-based concealment (the 6th and later consecutive lost frame,
// celt_decoder.c:446,470-489) steps the LCG once per synthesised coefficient.  SeedTrack carries exactly the decoder state
// that decides this — the seed, the loss streak (st->loss_count), the frame size / bandwidth the last packet set and whether
// any packet was decoded yet (st->prev_mode != 0) — so stage A can follow it through lost packets without stage B.
// ---------------------------------------------------------------------------------------------------
struct SeedTrack {
    unsigned seed;
    int streak;       // st->loss_count
    int fs_last;      // st->frame_size (API rate)
    int bw_last;      // st->bandwidth
    int have_mode;    // st->prev_mode != 0
    int channels;     // decoder output channels (constant)
};

// One concealment frame of `audiosize` samples (API rate): what celt_decode_lost does to (rng, loss_count).  More than 20 ms
// (the frame size a SILK / hybrid TOC can set) is concealed in 20 ms pieces (opus_decoder.c:257-268).
CB_DEV bool track_lost_chunk(SeedTrack &tr, int audiosize, int Fs);
CB_DEV bool track_lost_frame(SeedTrack &tr, int audiosize, int Fs) {
    const int F20 = Fs / 50;
    do {
        const int a = imin(audiosize, F20);
        if (!track_lost_chunk(tr, a, Fs)) return false;
        audiosize -= a;
    } while (audiosize > 0);
    return true;
}
CB_DEV bool track_lost_chunk(SeedTrack &tr, int audiosize, int Fs) {
    const int N = audiosize * (48000 / Fs);
    int LM;
    for (LM = 0; LM <= kMaxLM; LM++)
        if (kShortMdct << LM == N) break;
    if (LM > kMaxLM) return false;   // celt_decode_with_ec rejects the size: OPUS_BAD_ARG, state untouched
    if (tr.streak >= 5) {
        const int effEnd = imin(bandwidth_to_endband(tr.bw_last), kNbEBands);
        const int steps = tr.channels * (kEBands[effEnd] << LM);
        unsigned sd = tr.seed;
        CB_NOUNROLL for (int k = 0; k < steps; k++) sd = lcg_rand(sd);
        tr.seed = sd;
    }
    tr.streak++;
    return true;
}
// The frame size opus_decode_frame conceals when `room` samples are left (opus_decoder.c:244-300); 0 = nothing is concealed.
CB_DEV int lost_audiosize(const SeedTrack &tr, int room, int Fs) {
    const int F20 = Fs / 50, F10 = F20 >> 1, F5 = F10 >> 1, F2_5 = F5 >> 1;
    if (room < F2_5) return -1;
    int audiosize = imin(imin(room, Fs / 25 * 3), tr.fs_last);
    if (tr.have_mode && audiosize < F20) {
        if (audiosize > F10) audiosize = F10;
        else if (audiosize > F5 && audiosize < F10) audiosize = F5;
    }
    return audiosize;
}
// A wholly lost packet: opus_decode_native's concealment loop over `cap` samples (opus_decoder.c:613-627)
CB_DEV void track_lost_packet(SeedTrack &tr, int cap, int Fs) {
    int pcm_count = 0;
    do {
        const int a = lost_audiosize(tr, cap - pcm_count, Fs);
        if (a <= 0) break;
        if (tr.have_mode && !track_lost_frame(tr, a, Fs)) break;
        pcm_count += a;
    } while (pcm_count < cap);
}

CB_DEV void opus_parse_packet(const uint8_t *data, int len, int cap, int Fs, int decode_fec, int kmax, SeedTrack &tr,
                              CbPacketIR &pk, CbFrameIR *fr, int16_t *Xarea, ParseScratch &ps, bool dry) {
    pk.count = 0; pk.lost = 0; pk.frame_size = 0; pk.mode = 0; pk.bandwidth = 0; pk.stream_channels = 0; pk.reserved = 0;
    if (decode_fec < 0 || decode_fec > 1) { pk.ret = OPUS_BAD_ARG_; return; }
    if ((decode_fec || len == 0 || data == nullptr) && cap % (Fs / 400) != 0) { pk.ret = OPUS_BAD_ARG_; return; }
    if (len == 0 || data == nullptr) { pk.ret = 0; pk.lost = 1; track_lost_packet(tr, cap, Fs); return; }
    if (len < 0) { pk.ret = OPUS_BAD_ARG_; return; }
    const int packet_mode = pkt_mode(data);
    const int packet_frame_size = pkt_samples_per_frame(data, Fs);
    int16_t size[48];
    int offset;
    uint8_t toc;
    const int count = pkt_parse(data, len, 0, &toc, size, &offset, nullptr);
    if (count < 0) { pk.ret = count; return; }
    // scope edge: no SILK / hybrid decoding.  A SILK / hybrid TOC whose frames are all empty (<= 1 byte) is still ours: every
    // frame is concealed from the CELT state (opus_decoder.c:246-252) — the "PLC frames" an encoder emits when max_data_bytes
    // leaves no room for a frame (opus_encoder.c:1240-1270) look like that.
    if (packet_mode != CB_MODE_CELT_ONLY) {
        bool all_empty = !decode_fec;
        for (int i = 0; i < count; i++) all_empty &= size[i] <= 1;
        if (!all_empty) { pk.ret = OPUS_UNIMPLEMENTED_; return; }
    }
    if (decode_fec) { pk.ret = 0; pk.lost = 1; track_lost_packet(tr, cap, Fs); return; }   // CELT carries no in-band FEC (opus_decoder.c:655-657)
    if (count * packet_frame_size > cap) { pk.ret = OPUS_BUFFER_TOO_SMALL_; return; }
    if (count > kmax) { pk.ret = OPUS_BUFFER_TOO_SMALL_; return; }   // cannot happen when kmax = cap / (Fs/400)
    pk.ret = 0;
    pk.count = (int16_t)count;
    pk.frame_size = packet_frame_size;
    pk.mode = (int16_t)packet_mode;
    pk.bandwidth = (int16_t)pkt_bandwidth(data);
    pk.stream_channels = (int16_t)pkt_nb_channels(data);
    tr.fs_last = packet_frame_size;
    tr.bw_last = pk.bandwidth;
    const int C = pk.stream_channels;
    const int end = bandwidth_to_endband(pk.bandwidth);
    const int N = packet_frame_size * (48000 / Fs);   // CELT frame length at 48 kHz
    int LM = 0;
    while ((kShortMdct << LM) != N && LM < kMaxLM) LM++;   // (a 40 / 60 ms SILK-TOC frame has no LM: its frames are all concealed)
    const uint8_t *p = data + offset;
    int xoff = 0;
    int nb_samples = 0;
    CB_NOUNROLL for (int i = 0; i < count; i++) {
        CbFrameIR &ir = fr[i];
        ir.x_off = xoff;
        if (size[i] <= 1) {
            // concealment frame inside a received packet (opus_decoder.c:246-252): no symbols
            ir.flags = CB_IR_LOST;
            ir.len = size[i];
            ir.LM = (uint8_t)LM; ir.C = (uint8_t)C; ir.end = (uint8_t)end;
            ir.rng_final = 0; ir.seed_bands = tr.seed;
            const int a = lost_audiosize(tr, cap - nb_samples, Fs);
            if (a > 0 && tr.have_mode) track_lost_frame(tr, a, Fs);
            nb_samples += a > 0 ? a : 0;
        } else {
            celt_parse_frame(p, size[i], LM, C, end, &tr.seed, ir, Xarea + xoff, ps, dry);
            tr.streak = 0;
            tr.have_mode = 1;
            nb_samples += packet_frame_size;
        }
        xoff += N * C;
        p += size[i];
    }
}

// What stage A reads of a stream's state: the context at the START OF THE CALL.  It is a snapshot (taken on the call's stream
// before any stage runs) and not the live state, because a call is pipelined over chunks: stage A of chunk c+1 runs while stage B
// of chunk c is already advancing rng / loss_count / frame_size in the state.  (Found by tools/parity_sweep.py: a loss run that
// covers the start of a call through its whole first chunk made a second-chunk run fall back to a half-updated state.)
struct CbCallCtx {
    uint32_t rng;
    int32_t loss_count, frame_size, bandwidth, prev_mode, channels, Fs;
};
CB_DEV void call_ctx_from_state(CbCallCtx &c, const CbDecState *st) {
    c.rng = st->rng; c.loss_count = st->loss_count; c.frame_size = st->frame_size; c.bandwidth = st->bandwidth;
    c.prev_mode = st->prev_mode; c.channels = st->channels; c.Fs = st->Fs;
}

// One stage-A work item: packets [first, last) of one stream (a "run").  st = the stream's state as it was when the launch
// began (stage B of this launch has not run yet), call_f0 = first packet of the call, f0 = first packet of the chunk whose IR
// arrays pk_s / fr_s / X_s (already offset to this stream) are being filled: packet f goes to slot f - f0.
// Two phases around ONE call site of the (large) packet parser.  Seeking: f walks BACK from first-1 to the nearest packet
// that holds a received frame (it fixes seed, loss streak, frame size and bandwidth: SeedTrack), dry-parsing into the run's
// first IR slot as scratch; then f walks FORWARD — dry over the lost / rejected packets up to the run, for real over the run.
CB_DEV void opus_parse_run(const CbCallCtx &st0, const uint8_t *data, const int64_t *offs_s, const int32_t *lens_s, int call_f0, int first,
                           int last, int cap, int decode_fec, int kmax, int xstride, CbPacketIR *pk_s, CbFrameIR *fr_s, int16_t *X_s, int f0,
                           ParseScratch &ps) {
    const CbCallCtx *st = &st0;
    const int Fs = st->Fs;
    const size_t slot0 = (size_t)(first - f0);
    SeedTrack tr;
    tr.seed = st->rng; tr.streak = st->loss_count; tr.fs_last = st->frame_size; tr.bw_last = st->bandwidth;
    tr.have_mode = st->prev_mode != 0; tr.channels = st->channels;
    bool seeking = first > call_f0;
    int f = seeking ? first - 1 : first;
    while (f < last) {
        if (seeking && f < call_f0) {   // nothing received since the call began: the state is the context
            tr.seed = st->rng; tr.streak = st->loss_count; tr.fs_last = st->frame_size; tr.bw_last = st->bandwidth;
            tr.have_mode = st->prev_mode != 0;
            seeking = false;
            f = call_f0;
            continue;
        }
        const bool dry = f < first;
        const size_t slot = dry ? slot0 : (size_t)(f - f0);
        const int len = lens_s[f];
        const uint8_t *p = len > 0 ? data + offs_s[f] : nullptr;
        CbPacketIR pk;   // built in registers, stored once
        opus_parse_packet(p, len, cap, Fs, dry ? 0 : decode_fec, kmax, tr, pk, fr_s + slot * kmax, X_s + slot * xstride, ps, dry);
        if (!dry) pk_s[slot] = pk;
        if (seeking) {
            bool parsed = false;   // did any frame of this packet run the range decoder?
            if (pk.ret >= 0 && !pk.lost)
                for (int i = 0; i < pk.count; i++) parsed |= !(fr_s[slot * kmax + i].flags & CB_IR_LOST);
            if (parsed) {
                seeking = false;   // tr now holds the context after this packet: walk forward from the next one
                f++;
            } else {
                f--;
            }
        } else {
            f++;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Stage B: consume one packet's IR in stream order.
// ---------------------------------------------------------------------------------------------------

// What stage B tells stage C about one packet: which 48 kHz sample range of the packet's staging area holds signal
// that still has to be de-emphasised (PCM outside it was written by stage B itself: leading zero frames).
struct CbSigRange {
    int32_t begin, end;
};

// Concealment of one lost frame of frame_size samples, out of line (rare path; keeps the synthesis loop's registers free).  More
// than 20 ms — st->frame_size set by a SILK / hybrid TOC whose frames are all empty — goes in 20 ms pieces, each one a complete
// opus_decode_frame(NULL) call of the reference (opus_decoder.c:257-268).  Returns < 0 when the first piece failed.
template <class TM>
CB_DEV_NOINLINE int conceal_frame(TM tm, CbDecState *st, SynthScratch &S, PlcScratch &P, int *const *sig, int frame_size, int F20, int mode) {
    int done = 0, r = 0;
    do {
        const int chunk = imin(frame_size - done, F20);
        int *sig2[2] = {sig[0] + done * st->downsample, sig[1] + done * st->downsample};
        r = celt_decode_lost_frame(tm, st, S, P, sig2, chunk, bandwidth_to_endband(st->bandwidth));
        if (r < 0) break;
        done += chunk;
        if (done < frame_size) {   // what the reference's inner call leaves behind before the next piece
            if (tm.lane() == 0) { st->rangeFinal = 0; st->prev_mode = mode; st->prev_redundancy = 0; }
            tm.sync();
        }
    } while (done < frame_size);
    return done > 0 ? done : r;
}
// Zero-fill n samples of a PCM row (rejected packets, row tails), out of line for the same reason.
template <class TM>
CB_DEV_NOINLINE void zero_pcm(TM tm, int16_t *p, int n) {
    CB_TEAM_FOR(i, n, tm) p[i] = 0;
    tm.sync();
}

// opus_decode_frame remainder for one frame (opus_decoder.c:246-252,265-272,452-596).  sig[c] points at this frame's
// position in the packet's staging area; *staged is set when the frame left signal there.
template <class TM>
CB_DEV int opus_synth_frame(TM tm, CbDecState *st, SynthScratch &S, PlcScratch &P, const CbFrameIR &ir, int16_t *X, int16_t *pcm, int room,
                            int *const *sig, bool *staged) {
    const int F20 = st->Fs / 50, F10 = F20 >> 1, F5 = F10 >> 1, F2_5 = F5 >> 1;
    *staged = false;
    if (room < F2_5) return OPUS_BUFFER_TOO_SMALL_;
    int frame_size = imin(room, st->Fs / 25 * 3);
    const bool lost = (ir.flags & CB_IR_LOST) != 0;
    int audiosize, mode;
    if (!lost) {
        audiosize = st->frame_size;
        mode = st->mode;
    } else {
        frame_size = imin(frame_size, st->frame_size);
        audiosize = frame_size;
        mode = st->prev_mode;
        if (mode == 0) {
            CB_TEAM_FOR(i, audiosize * st->channels, tm) pcm[i] = 0;
            tm.sync();
            return audiosize;
        }
        // more than 20 ms (st->frame_size set by a SILK / hybrid TOC whose frames are all empty) is concealed in 20 ms pieces
        // below, each one a complete opus_decode_frame(NULL) call of the reference (opus_decoder.c:257-268)
        if (audiosize < F20) {
            if (audiosize > F10) audiosize = F10;
            else if (mode != CB_MODE_SILK_ONLY && audiosize > F5 && audiosize < F10) audiosize = F5;
        }
    }
    if (audiosize > frame_size) return OPUS_BAD_ARG_;
    frame_size = audiosize;
    int celt_ret;
    if (lost) {
        celt_ret = conceal_frame(tm, st, S, P, sig, frame_size, F20, mode);
        *staged = celt_ret >= 0;
    } else {
        celt_ret = celt_synth_frame(tm, st, S, ir, X, sig);
        *staged = true;
    }
    // decode gain (opus_decoder.c:567-577) is applied by stage C when it writes the PCM
    if (tm.lane() == 0) {
        st->rangeFinal = (lost || ir.len <= 1) ? 0 : ir.rng_final;
        st->prev_mode = mode;
        st->prev_redundancy = 0;
    }
    tm.sync();
    return celt_ret < 0 ? celt_ret : audiosize;
}

// opus_decode_native remainder for one packet (opus_decoder.c:613-627,682-708).  Returns what opus_decode returns.
// sigbase: the packet's staging area, channel c at sigbase + c*cap48 (cap48 = cap * downsample samples).
template <class TM>
CB_DEV int opus_synth_packet(TM tm, CbDecState *st, SynthScratch &S, PlcScratch &P, const CbPacketIR &pk, const CbFrameIR *fr, int16_t *Xarea,
                             int16_t *pcm, int cap, int *sigbase, CbSigRange *range) {
    const int ds = st->downsample;
    const int cap48 = cap * ds;
    int sb = 0, se = 0;   // staged range in 48 kHz samples
    bool any = false;
    int result;
    if (pk.ret < 0) {
        // rejected packet: the state stays untouched; its PCM row is zero-filled so that batch / span calls never hand stale
        // buffer contents back (libopus leaves the caller's buffer unspecified on error)
        result = pk.ret;
        zero_pcm(tm, pcm, cap * st->channels);
    } else if (pk.lost) {
        // conceal `cap` samples, frame by frame
        CbFrameIR lostir;
        lostir.flags = CB_IR_LOST; lostir.len = 0; lostir.rng_final = 0;
        int pcm_count = 0;
        result = 0;
        do {
            int *sig[2] = {sigbase + pcm_count * ds, sigbase + cap48 + pcm_count * ds};
            bool staged;
            int ret = opus_synth_frame(tm, st, S, P, lostir, nullptr, pcm + pcm_count * st->channels, cap - pcm_count, sig, &staged);
            if (ret < 0) { result = ret; break; }
            if (staged) { if (!any) sb = pcm_count * ds; any = true; se = (pcm_count + ret) * ds; }
            pcm_count += ret;
        } while (pcm_count < cap);
        if (result >= 0) {
            if (tm.lane() == 0) st->last_packet_duration = pcm_count;
            tm.sync();
            result = pcm_count;
        }
    } else {
        if (tm.lane() == 0) {
            st->mode = pk.mode;
            st->bandwidth = pk.bandwidth;
            st->frame_size = pk.frame_size;
            st->stream_channels = pk.stream_channels;
        }
        tm.sync();
        int nb_samples = 0;
        result = 0;
        CB_NOUNROLL for (int i = 0; i < pk.count; i++) {
            int *sig[2] = {sigbase + nb_samples * ds, sigbase + cap48 + nb_samples * ds};
            bool staged;
            int ret = opus_synth_frame(tm, st, S, P, fr[i], Xarea + fr[i].x_off, pcm + nb_samples * st->channels, cap - nb_samples, sig,
                                       &staged);
            if (staged) { if (!any) sb = nb_samples * ds; any = true; se = (nb_samples + (ret < 0 ? pk.frame_size : ret)) * ds; }
            if (ret < 0) { result = ret; break; }
            nb_samples += ret;
        }
        if (result >= 0) {
            if (tm.lane() == 0) st->last_packet_duration = nb_samples;
            tm.sync();
            result = nb_samples;
        }
    }
    if (tm.lane() == 0) { range->begin = any ? sb : 0; range->end = any ? se : 0; }
    // a packet shorter than the row's capacity: the rest of the row is zero-filled, so batch / span calls hand back
    // deterministic buffers (stage C writes [0, result) only)
    if (result >= 0 && result < cap) zero_pcm(tm, pcm + result * st->channels, (cap - result) * st->channels);
    return result;
}

// Stage C for one packet and one channel: de-emphasis + decode gain + PCM store (celt_decoder.c:185-275).
CB_DEV void opus_deemph_packet(CbDecState *st, int c, const int *sigbase, const CbSigRange &rg, int16_t *pcm, int cap) {
    if (rg.end <= rg.begin) return;
    const int ds = st->downsample, CC = st->channels;
    const int gain = st->decode_gain ? celt_exp2(s16(mul16_16_p15(21771, st->decode_gain))) : -1;   // QCONST16(6.48814081e-4f, 25); -1 = no gain stage (a very negative gain gives 0 = silence, opus_decoder.c:700-711)
    const int *x = sigbase + c * (cap * ds) + rg.begin;
    int16_t *y = pcm + (rg.begin / ds) * CC + c;
    st->preemph_memD[c] = deemphasis_channel(x, rg.end - rg.begin, y, CC, ds, st->preemph_memD[c], gain);
}

}  // namespace cb
