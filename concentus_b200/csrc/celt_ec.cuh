// celt_ec.cuh — the Opus range coder (decoder half) and the Laplace energy model.
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

}  // namespace cb
