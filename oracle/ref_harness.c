/* oracle/ref_harness.c — TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * Thin drivers around the UNMODIFIED reference (opus-fix FIXED_POINT libopus 1.1.2) public C API
 * (opus-fix/include/opus.h:171-512).  Linked into oracle/_ref/libopus_ref.so next to the reference
 * objects and called through ctypes by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.  It plays the role CSharp/ParityTest/TestDriver.cs:17-40 plays in the
 * reference: run the C build on an input, keep every packet byte / PCM sample / final range.
 *
 * Packed stream layout shared with the product's batch API:
 *   packets of stream s, frame f live at data[offs[s*F+f] .. +lens[s*F+f]).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <time.h>
#include "opus.h"

/* Signal of opus-fix/tests/test_opus_encode.c:59-90 (generate_music) driven by the MWC generator of
 * tests/test_opus_common.h:55-62, restated with an explicit seed (the reference seeds Rz=Rw=iseed). */
void ref_generate_music(int16_t *buf, int32_t len, uint32_t seed)
{
    uint32_t Rz = seed, Rw = seed;
    int32_t a1 = 0, b1 = 0, a2 = 0, b2 = 0, c1 = 0, c2 = 0, d1 = 0, d2 = 0, i, j = 0;
    for (i = 0; i < len; i++) {
        uint32_t r;
        int32_t v1, v2;
        v1 = v2 = (((j * ((j >> 12) ^ ((j >> 10 | j >> 12) & 26 & j >> 7))) & 128) + 128) << 15;
        Rz = 36969 * (Rz & 65535) + (Rz >> 16); Rw = 18000 * (Rw & 65535) + (Rw >> 16);
        r = (Rz << 16) + Rw; v1 += r & 65535; v1 -= r >> 16;
        Rz = 36969 * (Rz & 65535) + (Rz >> 16); Rw = 18000 * (Rw & 65535) + (Rw >> 16);
        r = (Rz << 16) + Rw; v2 += r & 65535; v2 -= r >> 16;
        b1 = v1 - a1 + ((b1 * 61 + 32) >> 6); a1 = v1;
        b2 = v2 - a2 + ((b2 * 61 + 32) >> 6); a2 = v2;
        c1 = (30 * (c1 + b1 + d1) + 32) >> 6; d1 = b1;
        c2 = (30 * (c2 + b2 + d2) + 32) >> 6; d2 = b2;
        v1 = (c1 + 128) >> 8;
        v2 = (c2 + 128) >> 8;
        buf[i * 2] = v1 > 32767 ? 32767 : (v1 < -32768 ? -32768 : v1);
        buf[i * 2 + 1] = v2 > 32767 ? 32767 : (v2 < -32768 ? -32768 : v2);
        if (i % 6 == 0) j++;
    }
}

typedef struct {
    int application, bitrate, vbr, cvbr, complexity, max_bytes, force_channels, bandwidth;
} ref_enc_cfg;

static OpusEncoder *make_encoder(int Fs, int channels, const ref_enc_cfg *c)
{
    int err = 0;
    OpusEncoder *e = opus_encoder_create(Fs, channels, c->application, &err);
    if (!e || err != OPUS_OK) return NULL;
    opus_encoder_ctl(e, OPUS_SET_BITRATE(c->bitrate));
    opus_encoder_ctl(e, OPUS_SET_VBR(c->vbr));
    opus_encoder_ctl(e, OPUS_SET_VBR_CONSTRAINT(c->cvbr));
    opus_encoder_ctl(e, OPUS_SET_COMPLEXITY(c->complexity));
    if (c->force_channels) opus_encoder_ctl(e, OPUS_SET_FORCE_CHANNELS(c->force_channels));
    if (c->bandwidth) opus_encoder_ctl(e, OPUS_SET_BANDWIDTH(c->bandwidth));
    return e;
}

/* Encode F frames of one stream.  out: F slots of `stride` bytes.  Returns 0 or a negative opus error. */
int ref_encode_stream(const int16_t *pcm, int F, int frame_size, int channels, int Fs,
                      const ref_enc_cfg *cfg, uint8_t *out, int stride, int32_t *lens, uint32_t *ranges)
{
    OpusEncoder *e = make_encoder(Fs, channels, cfg);
    int f, rc = 0;
    if (!e) return OPUS_ALLOC_FAIL;
    for (f = 0; f < F; f++) {
        int n = opus_encode(e, pcm + (size_t)f * frame_size * channels, frame_size,
                            out + (size_t)f * stride, cfg->max_bytes < stride ? cfg->max_bytes : stride);
        lens[f] = n;
        if (n < 0) { rc = n; break; }
        if (ranges) opus_encoder_ctl(e, OPUS_GET_FINAL_RANGE(&ranges[f]));
    }
    opus_encoder_destroy(e);
    return rc;
}

/* Encode F frames with a ctl script: before frame f the K (request, value) pairs script[(f*K+k)*2 .. +2) are applied through
 * opus_encoder_ctl (request 0 = nothing, request OPUS_RESET_STATE takes no value) — the shape of the reference's own encoder fuzz
 * loop (opus-fix/tests/test_opus_encode.c:236-330), restricted to integer-valued requests. */
int ref_encode_stream_script(const int16_t *pcm, int F, int frame_size, int channels, int Fs, const ref_enc_cfg *cfg,
                             const int32_t *script, int K, uint8_t *out, int stride, int32_t *lens, uint32_t *ranges)
{
    OpusEncoder *e = make_encoder(Fs, channels, cfg);
    int f, k;
    if (!e) return OPUS_ALLOC_FAIL;
    for (f = 0; f < F; f++) {
        for (k = 0; k < K; k++) {
            int req = script[(f * K + k) * 2], val = script[(f * K + k) * 2 + 1];
            if (req == 0) continue;
            if (req == OPUS_RESET_STATE) opus_encoder_ctl(e, OPUS_RESET_STATE);
            else opus_encoder_ctl(e, req, val);
        }
        lens[f] = opus_encode(e, pcm + (size_t)f * frame_size * channels, frame_size, out + (size_t)f * stride,
                              cfg->max_bytes < stride ? cfg->max_bytes : stride);
        if (ranges) opus_encoder_ctl(e, OPUS_GET_FINAL_RANGE(&ranges[f]));
    }
    opus_encoder_destroy(e);
    return 0;
}

/* Decode F packets of one stream (packed layout).  rets[f] = opus_decode return value. */
int ref_decode_stream(const uint8_t *data, const int64_t *offs, const int32_t *lens, int F,
                      int frame_size, int channels, int Fs, int16_t *pcm, uint32_t *ranges, int32_t *rets)
{
    int err = 0, f;
    OpusDecoder *d = opus_decoder_create(Fs, channels, &err);
    if (!d) return err;
    for (f = 0; f < F; f++) {
        const uint8_t *p = lens[f] > 0 ? data + offs[f] : NULL;
        int n = opus_decode(d, p, lens[f], pcm + (size_t)f * frame_size * channels, frame_size, 0);
        if (rets) rets[f] = n;
        if (ranges) opus_decoder_ctl(d, OPUS_GET_FINAL_RANGE(&ranges[f]));
    }
    opus_decoder_destroy(d);
    return 0;
}

/* The same with OPUS_SET_GAIN(gain_q8) applied before the first packet and OPUS_SET_GAIN(gain2_q8) before packet `switch_at`
 * (src/opus_decoder.c:836-846; applied in opus_decode_native :700-711). */
int ref_decode_stream_gain(const uint8_t *data, const int64_t *offs, const int32_t *lens, int F,
                           int frame_size, int channels, int Fs, int gain_q8, int gain2_q8, int switch_at,
                           int16_t *pcm, uint32_t *ranges, int32_t *rets)
{
    int err = 0, f;
    OpusDecoder *d = opus_decoder_create(Fs, channels, &err);
    if (!d) return err;
    opus_decoder_ctl(d, OPUS_SET_GAIN(gain_q8));
    for (f = 0; f < F; f++) {
        const uint8_t *p = lens[f] > 0 ? data + offs[f] : NULL;
        int n;
        if (f == switch_at) opus_decoder_ctl(d, OPUS_SET_GAIN(gain2_q8));
        n = opus_decode(d, p, lens[f], pcm + (size_t)f * frame_size * channels, frame_size, 0);
        if (rets) rets[f] = n;
        if (ranges) opus_decoder_ctl(d, OPUS_GET_FINAL_RANGE(&ranges[f]));
    }
    opus_decoder_destroy(d);
    return 0;
}

/* Decode with the getters read after every packet and OPUS_RESET_STATE issued before packet `reset_at` (-1 = never):
 * info[f*4 + {0,1,2,3}] = OPUS_GET_PITCH, OPUS_GET_LAST_PACKET_DURATION, OPUS_GET_BANDWIDTH, OPUS_GET_SAMPLE_RATE. */
int ref_decode_stream_info(const uint8_t *data, const int64_t *offs, const int32_t *lens, int F,
                           int frame_size, int channels, int Fs, int reset_at, int16_t *pcm, uint32_t *ranges,
                           int32_t *rets, int32_t *info)
{
    int err = 0, f;
    OpusDecoder *d = opus_decoder_create(Fs, channels, &err);
    if (!d) return err;
    for (f = 0; f < F; f++) {
        const uint8_t *p = lens[f] > 0 ? data + offs[f] : NULL;
        opus_int32 v = 0;
        if (f == reset_at) opus_decoder_ctl(d, OPUS_RESET_STATE);
        rets[f] = opus_decode(d, p, lens[f], pcm + (size_t)f * frame_size * channels, frame_size, 0);
        opus_decoder_ctl(d, OPUS_GET_FINAL_RANGE(&ranges[f]));
        opus_decoder_ctl(d, OPUS_GET_PITCH(&v)); info[f * 4 + 0] = v;
        opus_decoder_ctl(d, OPUS_GET_LAST_PACKET_DURATION(&v)); info[f * 4 + 1] = v;
        opus_decoder_ctl(d, OPUS_GET_BANDWIDTH(&v)); info[f * 4 + 2] = v;
        opus_decoder_ctl(d, OPUS_GET_SAMPLE_RATE(&v)); info[f * 4 + 3] = v;
    }
    opus_decoder_destroy(d);
    return 0;
}

/* ---- one-stream-per-thread pool (BASELINE.md §3 "Driver") ------------------------------------------ */
typedef struct {
    int kind; /* 0 = decode, 1 = encode */
    int S, F, frame_size, channels, Fs, stride;
    const uint8_t *data; const int64_t *offs; const int32_t *lens;   /* decode in / encode lens out */
    int16_t *pcm; uint32_t *ranges; int32_t *rets;
    const ref_enc_cfg *cfg; uint8_t *out; int32_t *olens;
    int pcm_shared; /* decode: all streams write the same scratch PCM slot per thread (bench); 0 = full output */
    volatile int next; pthread_mutex_t mu;
} pool_job;

static void *pool_worker(void *arg)
{
    pool_job *j = (pool_job *)arg;
    int16_t *scratch = NULL;
    if (j->kind == 0 && j->pcm_shared)
        scratch = (int16_t *)malloc((size_t)j->F * j->frame_size * j->channels * sizeof(int16_t));
    for (;;) {
        int s;
        pthread_mutex_lock(&j->mu); s = j->next++; pthread_mutex_unlock(&j->mu);
        if (s >= j->S) break;
        size_t fo = (size_t)s * j->F;
        if (j->kind == 0) {
            int16_t *dst = scratch ? scratch : j->pcm + fo * j->frame_size * j->channels;
            ref_decode_stream(j->data, j->offs + fo, j->lens + fo, j->F, j->frame_size, j->channels, j->Fs,
                              dst, j->ranges ? j->ranges + fo : NULL, j->rets ? j->rets + fo : NULL);
        } else {
            ref_encode_stream(j->pcm + fo * j->frame_size * j->channels, j->F, j->frame_size, j->channels, j->Fs,
                              j->cfg, j->out + fo * j->stride, j->stride, j->olens + fo,
                              j->ranges ? j->ranges + fo : NULL);
        }
    }
    free(scratch);
    return NULL;
}

static double run_pool(pool_job *j, int threads)
{
    struct timespec t0, t1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * threads);
    int i;
    j->next = 0; pthread_mutex_init(&j->mu, NULL);
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (i = 0; i < threads; i++) pthread_create(&th[i], NULL, pool_worker, j);
    for (i = 0; i < threads; i++) pthread_join(th[i], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    pthread_mutex_destroy(&j->mu); free(th);
    return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

/* Decode S streams x F frames on `threads` host threads; returns wall seconds of the decode loop.
 * pcm may be NULL (each thread then decodes into a private scratch: timing only). */
double ref_decode_streams_mt(int S, int F, int threads, const uint8_t *data, const int64_t *offs, const int32_t *lens,
                             int frame_size, int channels, int Fs, int16_t *pcm, uint32_t *ranges, int32_t *rets)
{
    pool_job j; memset(&j, 0, sizeof j);
    j.kind = 0; j.S = S; j.F = F; j.frame_size = frame_size; j.channels = channels; j.Fs = Fs;
    j.data = data; j.offs = offs; j.lens = lens; j.pcm = pcm; j.ranges = ranges; j.rets = rets;
    j.pcm_shared = (pcm == NULL);
    return run_pool(&j, threads);
}

/* Encode S streams x F frames; out has S*F slots of `stride` bytes. Returns wall seconds. */
double ref_encode_streams_mt(int S, int F, int threads, const int16_t *pcm, int frame_size, int channels, int Fs,
                             const ref_enc_cfg *cfg, uint8_t *out, int stride, int32_t *lens, uint32_t *ranges)
{
    pool_job j; memset(&j, 0, sizeof j);
    j.kind = 1; j.S = S; j.F = F; j.frame_size = frame_size; j.channels = channels; j.Fs = Fs;
    j.pcm = (int16_t *)pcm; j.cfg = cfg; j.out = out; j.stride = stride; j.olens = lens; j.ranges = ranges;
    return run_pool(&j, threads);
}
