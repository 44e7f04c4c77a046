// opus_repacketizer.cu — host half of the C ABI that only moves bytes: the repacketizer and opus_packet_pad / opus_packet_unpad
// for any packet (frame-count codes 0-3).  Interface: opus-fix/include/opus.h:628-750, behaviour: src/repacketizer.c:37-273
// (same return codes, same byte layout of the packets it writes).  No device code: packets are parsed with csrc/opus_packet.h,
// the parser the decoder kernels use.
#include <cstdlib>
#include <cstring>

#include "../../include/opus_b200.h"
#include "celt_simt.cuh"
#include "opus_packet.h"

// Frames collected by opus_repacketizer_cat.  The frames point into the caller's packets (opus.h: "the repacketizer state contains
// pointers to the submitted packets"), so those must stay valid until the next init.
struct OpusRepacketizer {
    unsigned char toc;
    int nb_frames;
    const unsigned char *frames[48];
    opus_int16 len[48];
    int framesize;   // samples per frame at 8 kHz: 120 ms = 960
};

namespace {

// a frame length as it is written in a packet header (one or two bytes, src/opus.c:148-167 read side)
inline int put_size(int size, unsigned char *p) {
    if (size < 252) {
        p[0] = (unsigned char)size;
        return 1;
    }
    p[0] = (unsigned char)(252 + (size & 3));
    p[1] = (unsigned char)((size - p[0]) >> 2);
    return 2;
}

// How frames [begin, end) are laid out: the cheapest code that can carry them (0: one frame, 1: two equal, 2: two different,
// 3: any count, CBR or VBR) — code 3 is forced when padding is wanted and there is room for it.
struct Layout {
    int code, vbr, header, total, pad;
};

int plan_layout(const opus_int16 *len, int count, opus_int32 maxlen, int want_pad, Layout &L) {
    int payload = 0;
    for (int i = 0; i < count; i++) payload += len[i];
    L.pad = 0;
    L.vbr = 0;
    if (count == 1) { L.code = 0; L.header = 1; }
    else if (count == 2 && len[0] == len[1]) { L.code = 1; L.header = 1; }
    else if (count == 2) { L.code = 2; L.header = 1 + 1 + (len[0] >= 252); }
    else L.code = 3;
    if (L.code != 3) {
        L.total = L.header + payload;
        if (L.total > maxlen) return OPUS_BUFFER_TOO_SMALL;
        if (!(want_pad && L.total < maxlen)) return OPUS_OK;
        L.code = 3;   // padding needs the code-3 header
    }
    for (int i = 1; i < count; i++)
        if (len[i] != len[0]) { L.vbr = 1; break; }
    L.header = 2;
    if (L.vbr)
        for (int i = 0; i < count - 1; i++) L.header += 1 + (len[i] >= 252);
    L.total = L.header + payload;
    if (L.total > maxlen) return OPUS_BUFFER_TOO_SMALL;
    if (want_pad && L.total < maxlen) {
        L.pad = maxlen - L.total;   // includes the padding-length bytes themselves
        L.total = maxlen;
    }
    return OPUS_OK;
}

// opus_repacketizer_out_range_impl (repacketizer.c:102-227) without self-delimited framing.  `data` may overlap the frames as
// long as every frame lies at or after the place it is moved to (in-place pad / unpad): frames are moved front to back.
opus_int32 emit_range(const OpusRepacketizer *rp, int begin, int end, unsigned char *data, opus_int32 maxlen, int want_pad) {
    if (begin < 0 || begin >= end || end > rp->nb_frames) return OPUS_BAD_ARG;
    const int count = end - begin;
    const opus_int16 *len = rp->len + begin;
    const unsigned char *const *frames = rp->frames + begin;
    Layout L;
    const int rc = plan_layout(len, count, maxlen, want_pad, L);
    if (rc != OPUS_OK) return rc;
    unsigned char *p = data;
    const unsigned char cfg = (unsigned char)(rp->toc & 0xFC);
    if (L.code < 3) {
        *p++ = (unsigned char)(cfg | L.code);
        if (L.code == 2) p += put_size(len[0], p);
    } else {
        *p++ = (unsigned char)(cfg | 3);
        *p++ = (unsigned char)(count | (L.vbr ? 0x80 : 0) | (L.pad ? 0x40 : 0));
        if (L.pad) {
            const int n255 = (L.pad - 1) / 255;
            for (int i = 0; i < n255; i++) *p++ = 255;
            *p++ = (unsigned char)(L.pad - 255 * n255 - 1);
        }
        if (L.vbr)
            for (int i = 0; i < count - 1; i++) p += put_size(len[i], p);
    }
    for (int i = 0; i < count; i++) {
        memmove(p, frames[i], (size_t)len[i]);
        p += len[i];
    }
    if (want_pad)
        while (p < data + maxlen) *p++ = 0;
    return L.total;
}

}  // namespace

extern "C" {

int opus_repacketizer_get_size(void) { return (int)sizeof(OpusRepacketizer); }

OpusRepacketizer *opus_repacketizer_init(OpusRepacketizer *rp) {
    rp->nb_frames = 0;
    return rp;
}

OpusRepacketizer *opus_repacketizer_create(void) {
    OpusRepacketizer *rp = (OpusRepacketizer *)malloc(sizeof(OpusRepacketizer));
    return rp ? opus_repacketizer_init(rp) : nullptr;
}

void opus_repacketizer_destroy(OpusRepacketizer *rp) { free(rp); }

// repacketizer.c:62-97
int opus_repacketizer_cat(OpusRepacketizer *rp, const unsigned char *data, opus_int32 len) {
    if (len < 1) return OPUS_INVALID_PACKET;
    if (rp->nb_frames == 0) {
        rp->toc = data[0];
        rp->framesize = cb::pkt_samples_per_frame(data, 8000);
    } else if ((rp->toc & 0xFC) != (data[0] & 0xFC)) {
        return OPUS_INVALID_PACKET;
    }
    const int incoming = opus_packet_get_nb_frames(data, len);
    if (incoming < 1) return OPUS_INVALID_PACKET;
    if ((incoming + rp->nb_frames) * rp->framesize > 960) return OPUS_INVALID_PACKET;   // 120 ms at most
    unsigned char toc;
    int offset = 0;
    const int n = cb::pkt_parse(data, len, 0, &toc, rp->len + rp->nb_frames, &offset, nullptr);
    if (n < 1) return n;
    const unsigned char *p = data + offset;
    for (int i = 0; i < n; i++) {
        rp->frames[rp->nb_frames + i] = p;
        p += rp->len[rp->nb_frames + i];
    }
    rp->nb_frames += incoming;
    return OPUS_OK;
}

int opus_repacketizer_get_nb_frames(OpusRepacketizer *rp) { return rp->nb_frames; }

opus_int32 opus_repacketizer_out_range(OpusRepacketizer *rp, int begin, int end, unsigned char *data, opus_int32 maxlen) {
    return emit_range(rp, begin, end, data, maxlen, 0);
}

opus_int32 opus_repacketizer_out(OpusRepacketizer *rp, unsigned char *data, opus_int32 maxlen) {
    return emit_range(rp, 0, rp->nb_frames, data, maxlen, 0);
}

// repacketizer.c:239-258: the packet is first moved to the end of the buffer, then re-emitted at its start with padding
int opus_packet_pad(unsigned char *data, opus_int32 len, opus_int32 new_len) {
    if (len < 1) return OPUS_BAD_ARG;
    if (len == new_len) return OPUS_OK;
    if (len > new_len) return OPUS_BAD_ARG;
    OpusRepacketizer rp;
    opus_repacketizer_init(&rp);
    memmove(data + new_len - len, data, (size_t)len);
    opus_repacketizer_cat(&rp, data + new_len - len, len);
    const opus_int32 ret = emit_range(&rp, 0, rp.nb_frames, data, new_len, 1);
    return ret > 0 ? OPUS_OK : ret;
}

// repacketizer.c:260-273
opus_int32 opus_packet_unpad(unsigned char *data, opus_int32 len) {
    if (len < 1) return OPUS_BAD_ARG;
    OpusRepacketizer rp;
    opus_repacketizer_init(&rp);
    const int rc = opus_repacketizer_cat(&rp, data, len);
    if (rc < 0) return rc;
    return emit_range(&rp, 0, rp.nb_frames, data, len, 0);
}

}  // extern "C"
