// celt_ec.cuh — the Opus range coder (decoder and encoder) and the Laplace energy model.
//
// Restates opus-fix/celt/entcode.h:63-121, celt/entcode.c:69-99 (ec_tell_frac), celt/entdec.c:93-316
// and celt/laplace.c:44-134 as member functions of one small struct that lives in the registers of the
// lane that walks the bitstream.  Range bytes are consumed front-to-back, raw bits back-to-front, both
// straight from the packet in global memory (a packet is <= 1275 B and is touched once).
#pragma once
#include "celt_arith.cuh"

namespace cb {

enum {
    kEcSymBits = 8, kEcCodeBits = 32, kEcSymMax = 255, kEcCodeShift = 23,
    kEcCodeExtra = 7, kEcUintBits = 8, kEcWindow = 32,
};
#define CB_EC_CODE_TOP 0x80000000u
#define CB_EC_CODE_BOT 0x00800000u

// Heavy symbol readers are called from ~50 sites: they are kept out of line (CB_MEM_NOINLINE) so that stage A fits the
// instruction cache; the short ones (bit_logp, tell, tell_frac) stay inline.

struct EcDec {
    const uint8_t *buf;
    unsigned storage;      // bytes available to the coder (may shrink: redundancy / raw-bit reservations)
    unsigned end_offs;     // raw-bit bytes consumed from the tail
    unsigned end_window;
    int nend_bits;
    int nbits_total;
    unsigned offs;         // next range byte
    unsigned rng, val, ext;
    int rem;
    int error;

    CB_MEM int read_byte() { return offs < storage ? buf[offs++] : 0; }
    CB_MEM int read_byte_from_end() { return end_offs < storage ? buf[storage - ++end_offs] : 0; }

    // entdec.c:104-131
    CB_MEM void normalize() {
        while (rng <= CB_EC_CODE_BOT) {
            nbits_total += kEcSymBits;
            rng <<= kEcSymBits;
            int sym = rem;
            rem = read_byte();
            sym = (sym << kEcSymBits | rem) >> (kEcSymBits - kEcCodeExtra);
            val = ((val << kEcSymBits) + (kEcSymMax & ~sym)) & (CB_EC_CODE_TOP - 1);
        }
    }
    // entdec.c:133-153
    CB_MEM void init(const uint8_t *b, unsigned n) {
        buf = b; storage = n; end_offs = 0; end_window = 0; nend_bits = 0;
        nbits_total = kEcCodeBits + 1 - ((kEcCodeBits - kEcCodeExtra) / kEcSymBits) * kEcSymBits;
        offs = 0;
        rng = 1u << kEcCodeExtra;
        rem = read_byte();
        val = rng - 1 - (rem >> (kEcSymBits - kEcCodeExtra));
        ext = 0; error = 0;
        normalize();
    }
    // entcode.h:114-121, entcode.c:69-99
    CB_MEM int tell() const { return nbits_total - ec_ilog(rng); }
    CB_MEM unsigned tell_frac() const {
        unsigned nbits = (unsigned)nbits_total << kBitResEc;
        int l = ec_ilog(rng);
        unsigned r = rng >> (l - 16);
        unsigned b = (r >> 12) - 8;
        // correction[] thresholds of entcode.c:75-77
        const unsigned corr = b == 0 ? 35733u : b == 1 ? 38967u : b == 2 ? 42495u : b == 3 ? 46340u :
                              b == 4 ? 50535u : b == 5 ? 55109u : b == 6 ? 60097u : 65535u;
        b += r > corr;
        l = (l << 3) + (int)b;
        return nbits - (unsigned)l;
    }
    // entdec.c:155-172
    CB_MEM_NOINLINE unsigned decode(unsigned ft) {
        ext = rng / ft;
        unsigned s = val / ext;
        return ft - imin_u(s + 1, ft);
    }
    CB_MEM unsigned decode_bin(unsigned bits) {
        ext = rng >> bits;
        unsigned s = val / ext;
        return (1u << bits) - imin_u(s + 1u, 1u << bits);
    }
    // entdec.c:181-200
    CB_MEM_NOINLINE void update(unsigned fl, unsigned fh, unsigned ft) {
        unsigned s = ext * (ft - fh);
        val -= s;
        rng = fl > 0 ? ext * (fh - fl) : rng - s;
        normalize();
    }
    // entdec.c:203-216
    CB_MEM int bit_logp(unsigned logp) {
        unsigned r = rng, d = val, s = r >> logp;
        int ret = d < s;
        if (!ret) val = d - s;
        rng = ret ? s : r - s;
        normalize();
        return ret;
    }
    // entdec.c:218-236
    CB_MEM_NOINLINE int icdf(const uint8_t *tab, unsigned ftb) {
        unsigned s = rng, d = val, r = s >> ftb, t;
        int ret = -1;
        do {
            t = s;
            s = r * tab[++ret];
        } while (d < s);
        val = d - s;
        rng = t - s;
        normalize();
        return ret;
    }
    // entdec.c:284-316
    CB_MEM_NOINLINE unsigned bits(unsigned nb) {
        unsigned window = end_window;
        int available = nend_bits;
        if ((unsigned)available < nb) {
            do {
                window |= (unsigned)read_byte_from_end() << available;
                available += kEcSymBits;
            } while (available <= kEcWindow - kEcSymBits);
        }
        unsigned ret = window & ((1u << nb) - 1u);
        window >>= nb;
        available -= nb;
        end_window = window;
        nend_bits = available;
        nbits_total += nb;
        return ret;
    }
    // entdec.c:238-282
    CB_MEM_NOINLINE unsigned uint_(unsigned ft_in) {
        unsigned ft = ft_in - 1;
        int ftb = ec_ilog(ft);
        if (ftb > kEcUintBits) {
            ftb -= kEcUintBits;
            unsigned f = (ft >> ftb) + 1;
            unsigned s = decode(f);
            update(s, s + 1, f);
            unsigned t = s << ftb | bits(ftb);
            if (t <= ft) return t;
            error = 1;
            return ft;
        } else {
            ft++;
            unsigned s = decode(ft);
            update(s, s + 1, ft);
            return s;
        }
    }

    // laplace.c:44-49, :94-134.  fs = P(0) in Q15, decay in Q14.
    CB_MEM_NOINLINE int laplace(unsigned fs, int decay) {
        int v = 0;
        unsigned fm = decode_bin(15);
        unsigned fl = 0;
        if (fm >= fs) {
            v++;
            fl = fs;
            fs = laplace_freq1(fs, decay) + kLaplaceMinP;
            while (fs > kLaplaceMinP && fm >= fl + 2 * fs) {
                fs *= 2;
                fl += fs;
                fs = ((fs - 2 * kLaplaceMinP) * (int)decay) >> 15;
                fs += kLaplaceMinP;
                v++;
            }
            if (fs <= kLaplaceMinP) {
                int di = (fm - fl) >> (kLaplaceLogMinP + 1);
                v += di;
                fl += 2 * di * kLaplaceMinP;
            }
            if (fm < fl + fs) v = -v;
            else fl += fs;
        }
        unsigned fh = fl + fs;
        update(fl, fh < 32768u ? fh : 32768u, 32768u);
        return v;
    }

    enum { kBitResEc = 3, kLaplaceLogMinP = 0, kLaplaceMinP = 1, kLaplaceNMin = 16 };
    static CB_MEM unsigned imin_u(unsigned a, unsigned b) { return a < b ? a : b; }
    static CB_MEM unsigned laplace_freq1(unsigned fs0, int decay) {
        unsigned ft = 32768 - kLaplaceMinP * (2 * kLaplaceNMin) - fs0;
        return (ft * (unsigned)(16384 - decay)) >> 15;
    }
};

// ---------------------------------------------------------------------------------------------------
// Range ENCODER (opus-fix/celt/entenc.c:62-508) and Laplace encode (celt/laplace.c:51-92).
// Same register-resident struct style; `buf` is the packet under construction (global memory).
// ---------------------------------------------------------------------------------------------------
struct EcEnc {
    uint8_t *buf;
    unsigned storage, end_offs, end_window;
    int nend_bits, nbits_total;
    unsigned offs, rng, val, ext;
    int rem, error;

    CB_MEM int write_byte(unsigned v) {
        if (offs + end_offs >= storage) return -1;
        buf[offs++] = (uint8_t)v;
        return 0;
    }
    CB_MEM int write_byte_at_end(unsigned v) {
        if (offs + end_offs >= storage) return -1;
        buf[storage - ++end_offs] = (uint8_t)v;
        return 0;
    }
    // entenc.c:111-128
    CB_MEM void carry_out(int c) {
        if (c != kEcSymMax) {
            int carry = c >> kEcSymBits;
            if (rem >= 0) error |= write_byte((unsigned)(rem + carry));
            if (ext > 0) {
                unsigned sym = (unsigned)(kEcSymMax + carry) & kEcSymMax;
                CB_NOUNROLL do error |= write_byte(sym);
                while (--ext > 0);
            }
            rem = c & kEcSymMax;
        } else {
            ext++;
        }
    }
    // entenc.c:145-152
    CB_MEM void normalize_inl() {
        CB_NOUNROLL while (rng <= CB_EC_CODE_BOT) {
            carry_out((int)(val >> kEcCodeShift));
            val = (val << kEcSymBits) & (CB_EC_CODE_TOP - 1);
            rng <<= kEcSymBits;
            nbits_total += kEcSymBits;
        }
    }
    // the band walk's symbols (encode / bit_logp / uint) share one copy of the loop under CB_TINY_CODE; the header symbols of the
    // thread-per-stream stages (encode_bin / icdf / laplace) keep it inline: there a call is latency, not instruction-cache space
    CB_MEM_TINY void normalize() { normalize_inl(); }
    // entenc.c:170-184
    CB_MEM void init(uint8_t *b, unsigned size) {
        buf = b; end_offs = 0; end_window = 0; nend_bits = 0;
        nbits_total = kEcCodeBits + 1;
        offs = 0; rng = CB_EC_CODE_TOP; rem = -1; val = 0; ext = 0; storage = size; error = 0;
    }
    CB_MEM int tell() const { return nbits_total - ec_ilog(rng); }
    CB_MEM_TINY unsigned tell_frac() const {
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
    // entenc.c:187-216
    CB_MEM_NOINLINE void encode(unsigned fl, unsigned fh, unsigned ft) {
        unsigned r = rng / ft;
        if (fl > 0) {
            val += rng - r * (ft - fl);
            rng = r * (fh - fl);
        } else {
            rng -= r * (ft - fh);
        }
        normalize();
    }
    CB_MEM_NOINLINE void encode_bin(unsigned fl, unsigned fh, unsigned bits) {
        unsigned r = rng >> bits;
        if (fl > 0) {
            val += rng - r * ((1u << bits) - fl);
            rng = r * (fh - fl);
        } else {
            rng -= r * ((1u << bits) - fh);
        }
        normalize_inl();
    }
    // entenc.c:249-277
    CB_MEM_NOINLINE void bit_logp(int v, unsigned logp) {
        unsigned r = rng, l = val, s = r >> logp;
        r -= s;
        if (v) val = l + r;
        rng = v ? s : r;
        normalize();
    }
    // entenc.c:279-311
    CB_MEM_NOINLINE void icdf(int s, const uint8_t *tab, unsigned ftb) {
        unsigned r = rng >> ftb;
        if (s > 0) {
            val += rng - r * tab[s - 1];
            rng = r * (unsigned)(tab[s - 1] - tab[s]);
        } else {
            rng -= r * tab[s];
        }
        normalize_inl();
    }
    // entenc.c:346-384
    CB_MEM_NOINLINE void bits(unsigned fl, unsigned nb) {
        unsigned window = end_window;
        int used = nend_bits;
        if (used + (int)nb > kEcWindow) {
            CB_NOUNROLL do {
                error |= write_byte_at_end(window & kEcSymMax);
                window >>= kEcSymBits;
                used -= kEcSymBits;
            } while (used >= kEcSymBits);
        }
        window |= fl << used;
        used += nb;
        end_window = window;
        nend_bits = used;
        nbits_total += nb;
    }
    // entenc.c:313-344
    CB_MEM_NOINLINE void uint_(unsigned fl, unsigned ft_in) {
        unsigned ft = ft_in - 1;
        int ftb = ec_ilog(ft);
        if (ftb > kEcUintBits) {
            ftb -= kEcUintBits;
            unsigned f = (ft >> ftb) + 1;
            unsigned l = fl >> ftb;
            encode(l, l + 1, f);
            bits(fl & ((1u << ftb) - 1u), (unsigned)ftb);
        } else {
            encode(fl, fl + 1, ft + 1);
        }
    }
    // entenc.c:386-425
    CB_MEM void patch_initial_bits(unsigned v, unsigned nbits) {
        int shift = kEcSymBits - (int)nbits;
        unsigned mask = ((1u << nbits) - 1) << shift;
        if (offs > 0) buf[0] = (uint8_t)((buf[0] & ~mask) | v << shift);
        else if (rem >= 0) rem = (int)(((unsigned)rem & ~mask) | v << shift);
        else if (rng <= (CB_EC_CODE_TOP >> nbits)) val = (val & ~(mask << kEcCodeShift)) | v << (kEcCodeShift + shift);
        else error = -1;
    }
    // entenc.c:427-445: move the raw-bit tail down to the new end
    CB_MEM void shrink(unsigned size) {
        // memmove semantics: destination is below the source
        CB_NOUNROLL for (unsigned i = 0; i < end_offs; i++) buf[size - end_offs + i] = buf[storage - end_offs + i];
        storage = size;
    }
    // entenc.c:447-508
    CB_MEM_NOINLINE void done() {
        int l = kEcCodeBits - ec_ilog(rng);
        unsigned msk = (CB_EC_CODE_TOP - 1) >> l;
        unsigned end = (val + msk) & ~msk;
        if ((end | msk) >= val + rng) {
            l++;
            msk >>= 1;
            end = (val + msk) & ~msk;
        }
        while (l > 0) {
            carry_out((int)(end >> kEcCodeShift));
            end = (end << kEcSymBits) & (CB_EC_CODE_TOP - 1);
            l -= kEcSymBits;
        }
        if (rem >= 0 || ext > 0) carry_out(0);
        unsigned window = end_window;
        int used = nend_bits;
        while (used >= kEcSymBits) {
            error |= write_byte_at_end(window & kEcSymMax);
            window >>= kEcSymBits;
            used -= kEcSymBits;
        }
        if (!error) {
            CB_NOUNROLL for (unsigned i = offs; i < storage - end_offs; i++) buf[i] = 0;
            if (used > 0) {
                if (end_offs >= storage) error = -1;
                else {
                    l = -l;
                    if (offs + end_offs >= storage && l < used) {
                        window &= (1u << l) - 1;
                        error = -1;
                    }
                    buf[storage - end_offs - 1] |= (uint8_t)window;
                }
            }
        }
    }
    // ec_laplace_encode (laplace.c:51-92); *value may be clamped
    CB_MEM_NOINLINE void laplace(int *value, unsigned fs, int decay) {
        unsigned fl = 0;
        int v = *value;
        if (v) {
            int s = -(v < 0);
            v = (v + s) ^ s;
            fl = fs;
            fs = EcDec::laplace_freq1(fs, decay);
            int i;
            CB_NOUNROLL for (i = 1; fs > 0 && i < v; i++) {
                fs *= 2;
                fl += fs + 2 * 1;
                fs = (fs * (unsigned)decay) >> 15;
            }
            if (!fs) {
                int ndi_max = (int)((32768 - fl + 1 - 1) >> 0);
                ndi_max = (ndi_max - s) >> 1;
                int di = v - i < ndi_max - 1 ? v - i : ndi_max - 1;
                fl += (unsigned)((2 * di + 1 + s) * 1);
                fs = 1u < 32768 - fl ? 1u : 32768 - fl;
                *value = (i + di + s) ^ s;
            } else {
                fs += 1;
                fl += fs & ~(unsigned)s;
            }
        }
        encode_bin(fl, fl + fs, 15);
    }
};

}  // namespace cb
