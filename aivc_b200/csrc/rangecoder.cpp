// Host range coder of the AIVC bitstream (product side).
//
// Byte-compatible with the arithmetic coder AIVC reaches through torchac
// (src/real_life/bitstream.py:281 encode_float_cdf, :454/:482 decode_float_cdf): 32-bit
// low/high interval, 16-bit CDFs, MSB-first bit packing, underflow ("pending") bits, one
// terminating bit and zero padding.  Unlike the bit-at-a-time reference loop this coder
// renormalises in bulk: all settled leading bits are shifted out with one clz, then all
// underflow positions with another, and bits are staged in a 64-bit accumulator.
// The Laplace decoder evaluates the integer CDF on the fly (laplace_cdf.h) instead of
// reading the reference's [C,H,W,514] float table (bitstream.py:127-154).
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include "../../include/aivc_b200.h"
#include "laplace_cdf.h"

void aivc_set_error(const char *fmt, ...);

namespace {

struct BitWriter {
    uint8_t *out;
    size_t cap, len = 0;
    uint64_t acc = 0;
    int nbits = 0;
    bool overflow = false;

    BitWriter(uint8_t *o, size_t c) : out(o), cap(c) {}

    inline void drain() {
        while (nbits >= 8) {
            nbits -= 8;
            if (len < cap) out[len] = (uint8_t)(acc >> nbits);
            else overflow = true;
            ++len;
        }
    }
    // n <= 32 bits, MSB first
    inline void put(uint32_t bits, int n) {
        if (n == 0) return;
        acc = (acc << n) | (uint64_t)(n == 32 ? bits : (bits & ((1u << n) - 1u)));
        nbits += n;
        drain();
    }
    inline void put_run(int bit, uint64_t count) {
        const uint32_t pat = bit ? 0xFFFFFFFFu : 0u;
        while (count >= 32) { put(pat, 32); count -= 32; }
        put(pat, (int)count);
    }
    inline void finish() {
        if (nbits > 0) put(0, 8 - nbits);   // zero padding to a byte boundary
    }
};

struct Encoder {
    uint32_t low = 0, high = 0xFFFFFFFFu;
    uint64_t pending = 0;
    BitWriter bw;
    Encoder(uint8_t *o, size_t c) : bw(o, c) {}

    inline void step(uint32_t c_low, uint32_t c_high) {
        const uint64_t span = (uint64_t)high - (uint64_t)low + 1;
        high = (low - 1) + (uint32_t)((span * c_high) >> 16);
        low = low + (uint32_t)((span * c_low) >> 16);
        const uint32_t diff = low ^ high;
        if (!(diff & 0x80000000u)) {
            // n leading bits agree: they are final.  First one carries the pending run.
            const int n = diff ? __builtin_clz(diff) : 32;
            const int first = (int)(low >> 31);
            bw.put((uint32_t)first, 1);
            if (pending) { bw.put_run(!first, pending); pending = 0; }
            if (n > 1) bw.put(n == 32 ? low : (low >> (32 - n)), n - 1);
            if (n == 32) { low = 0; high = 0xFFFFFFFFu; }
            else { low <<= n; high = (high << n) | ((1u << n) - 1u); }
        }
        // now low = 0..., high = 1...; strip positions where low continues 1.. and high 0..
        const uint32_t t = (low & ~high) << 1;
        const int m = (~t) ? __builtin_clz(~t) : 32;
        if (m > 0) {
            pending += (uint64_t)m;
            low = (low << m) & 0x7FFFFFFFu;
            high = (high << m) | 0x80000000u | ((1u << m) - 1u);
        }
    }
    inline void finish() {
        pending += 1;
        const int bit = low < 0x40000000u ? 0 : 1;
        bw.put((uint32_t)bit, 1);
        bw.put_run(!bit, pending);
        bw.finish();
    }
};

struct BitReader {
    const uint8_t *in;
    size_t len, pos = 0;
    uint64_t acc = 0;     // holds `avail` unread bits in its low part
    int avail = 0;
    BitReader(const uint8_t *i, size_t l) : in(i), len(l) {}
    inline uint32_t get(int n) {          // n <= 32; bits past the end read as zero
        if (n == 0) return 0;
        if (avail < n) {                  // refill: four bytes at once while they exist, bytewise at the tail
            if (avail <= 32 && pos + 4 <= len) {
                uint32_t w;
                memcpy(&w, in + pos, 4);
                acc = (acc << 32) | (uint64_t)__builtin_bswap32(w);
                pos += 4;
                avail += 32;
            } else {
                while (avail < n) {
                    acc = (acc << 8) | (pos < len ? in[pos] : 0);
                    ++pos;
                    avail += 8;
                }
            }
        }
        avail -= n;
        const uint64_t v = acc >> avail;
        return (uint32_t)(n == 32 ? v : (v & ((1ull << n) - 1ull)));
    }
};

struct Decoder {
    uint32_t low = 0, high = 0xFFFFFFFFu, value;
    BitReader br;
    Decoder(const uint8_t *i, size_t l) : br(i, l) { value = br.get(32); }

    inline uint32_t target() const {
        const uint64_t span = (uint64_t)high - (uint64_t)low + 1;
        return (uint32_t)((((uint64_t)value - (uint64_t)low + 1) * 0x10000ull - 1) / span) & 0xFFFFu;
    }
    inline void consume(uint32_t c_low, uint32_t c_high) {
        const uint64_t span = (uint64_t)high - (uint64_t)low + 1;
        high = (low - 1) + (uint32_t)((span * c_high) >> 16);
        low = low + (uint32_t)((span * c_low) >> 16);
        const uint32_t diff = low ^ high;
        if (!(diff & 0x80000000u)) {
            const int n = diff ? __builtin_clz(diff) : 32;
            if (n == 32) { low = 0; high = 0xFFFFFFFFu; value = br.get(32); }
            else {
                low <<= n; high = (high << n) | ((1u << n) - 1u);
                value = (value << n) | br.get(n);
            }
        }
        const uint32_t t = (low & ~high) << 1;
        const int m = (~t) ? __builtin_clz(~t) : 32;
        if (m > 0) {
            low = (low << m) & 0x7FFFFFFFu;
            high = (high << m) | 0x80000000u | ((1u << m) - 1u);
            value = (value & 0x80000000u) | ((value << m) & 0x7FFFFFFFu) | br.get(m);
        }
    }
};

// largest s in [0, 512] with row[s] <= target  (rows are strictly increasing)
inline int table_search(const uint16_t *row, uint32_t target) {
    int lo = 0, hi = AIVC_AC_LP - 1;      // invariant: row[lo] <= target < row[hi] (virtually)
    while (lo + 1 < hi) {
        const int mid = (lo + hi) >> 1;
        if (row[mid] <= target) lo = mid; else hi = mid;
    }
    return lo;
}

// same search on the analytic Laplace CDF; starts at the mode (q = 0) and walks outwards a
// few steps before bisecting, because almost all symbols sit within a few bins of zero.
inline int laplace_search(float b, uint32_t target, uint32_t *c_low, uint32_t *c_high) {
    int s = AIVC_AC_MAX_VAL;
    uint32_t lo = aivc_laplace_cdf_int(b, s);
    if (target >= lo) {
        uint32_t hi = aivc_laplace_cdf_int(b, s + 1);
        int steps = 0;
        while (target >= hi && steps < 3) {
            ++s; ++steps; lo = hi; hi = aivc_laplace_cdf_int(b, s + 1);
        }
        if (target >= hi) {                       // bisect (s, 512]
            int l = s + 1, h = AIVC_AC_LP - 1;    // cdf(l) <= target < cdf(h) (virtually)
            while (l + 1 < h) {
                const int mid = (l + h) >> 1;
                if (aivc_laplace_cdf_int(b, mid) <= target) l = mid; else h = mid;
            }
            s = l; lo = aivc_laplace_cdf_int(b, s); hi = aivc_laplace_cdf_int(b, s + 1);
        }
        *c_low = lo; *c_high = hi;
        return s;
    }
    uint32_t hi = lo;
    int steps = 0;
    --s; lo = aivc_laplace_cdf_int(b, s);
    while (target < lo && steps < 3 && s > 0) {
        --s; ++steps; hi = lo; lo = aivc_laplace_cdf_int(b, s);
    }
    if (target < lo) {                            // bisect [0, s)
        int l = 0, h = s;                         // cdf(l) <= target < cdf(h)
        while (l + 1 < h) {
            const int mid = (l + h) >> 1;
            if (aivc_laplace_cdf_int(b, mid) <= target) l = mid; else h = mid;
        }
        s = l; lo = aivc_laplace_cdf_int(b, s); hi = aivc_laplace_cdf_int(b, s + 1);
    }
    *c_low = lo; *c_high = hi;
    return s;
}

}  // namespace

extern "C" {

size_t aivc_rc_bound(size_t n) { return 2 * n + n / 4 + 64; }   // <= 16 bits/symbol + slack

int aivc_rc_encode_bounds(const uint32_t *bounds, size_t n, uint8_t *out, size_t cap, size_t *out_len) {
    Encoder e(out, cap);
    for (size_t i = 0; i < n; ++i) {
        const uint32_t b = bounds[i];
        e.step(b & 0xFFFFu, b >> 16);
    }
    e.finish();
    if (e.bw.overflow) { aivc_set_error("rc_encode_bounds: output buffer too small"); return 1; }
    *out_len = e.bw.len;
    return 0;
}

int aivc_rc_encode_table(const uint16_t *table, const int16_t *sym, int c, size_t hw, uint8_t *out,
                         size_t cap, size_t *out_len) {
    Encoder e(out, cap);
    for (int ch = 0; ch < c; ++ch) {
        const uint16_t *row = table + (size_t)ch * AIVC_AC_LP;
        const int16_t *s = sym + (size_t)ch * hw;
        for (size_t i = 0; i < hw; ++i) {
            const int v = (int)s[i] + AIVC_AC_MAX_VAL;
            if (v < 0 || v > AIVC_AC_LP - 2) { aivc_set_error("rc_encode_table: symbol %d out of range", (int)s[i]); return 1; }
            e.step(row[v], v == AIVC_AC_LP - 2 ? 0x10000u : row[v + 1]);
        }
    }
    e.finish();
    if (e.bw.overflow) { aivc_set_error("rc_encode_table: output buffer too small"); return 1; }
    *out_len = e.bw.len;
    return 0;
}

int aivc_rc_decode_table(const uint16_t *table, const uint8_t *in, size_t in_len, int c, size_t hw,
                         int16_t *sym) {
    Decoder d(in, in_len);
    for (int ch = 0; ch < c; ++ch) {
        const uint16_t *row = table + (size_t)ch * AIVC_AC_LP;
        int16_t *s = sym + (size_t)ch * hw;
        for (size_t i = 0; i < hw; ++i) {
            const int v = table_search(row, d.target());
            s[i] = (int16_t)(v - AIVC_AC_MAX_VAL);
            d.consume(row[v], v == AIVC_AC_LP - 2 ? 0x10000u : row[v + 1]);
        }
    }
    return 0;
}

int aivc_rc_decode_laplace(const float *b, const uint8_t *in, size_t in_len, size_t n, int16_t *sym) {
    Decoder d(in, in_len);
    for (size_t i = 0; i < n; ++i) {
        uint32_t lo, hi;
        const int v = laplace_search(b[i], d.target(), &lo, &hi);
        sym[i] = (int16_t)(v - AIVC_AC_MAX_VAL);
        d.consume(lo, v == AIVC_AC_LP - 2 ? 0x10000u : hi);
    }
    return 0;
}


// As aivc_rc_decode_laplace, with the device-evaluated CDF window win[8 n] (entries 253..260 of every
// symbol): the analytic search only runs for symbols outside q = -3..+3.
int aivc_rc_decode_laplace_win(const float *b, const uint16_t *win, const uint8_t *in, size_t in_len, size_t n,
                               int16_t *sym) {
    Decoder d(in, in_len);
    for (size_t i = 0; i < n; ++i) {
        const uint16_t *w = win + 8 * i;
        // target >= c  <=>  floor((X - 1) / span) >= c  <=>  X > c * span   with X = (value - low + 1) << 16:
        // the window is searched with multiplications; the 64-bit division only runs for outliers
        const uint64_t span = (uint64_t)d.high - (uint64_t)d.low + 1;
        const uint64_t X = ((uint64_t)d.value - (uint64_t)d.low + 1) << 16;
        uint32_t lo, hi;
        int v;
        // branch-free: how many of the eight window entries lie at or below the target (symbols are close to
        // random, so a data-dependent walk would mispredict every other symbol)
        int cnt = 0;
#pragma GCC unroll 8
        for (int j = 0; j < 8; ++j) cnt += (X > w[j] * span) ? 1 : 0;
        if (cnt >= 1 && cnt <= 7) {
            const int j = cnt - 1;
            v = AIVC_WIN_FIRST + j; lo = w[j]; hi = w[j + 1];
        } else {
            v = laplace_search(b[i], d.target(), &lo, &hi);
        }
        sym[i] = (int16_t)(v - AIVC_AC_MAX_VAL);
        d.consume(lo, v == AIVC_AC_LP - 2 ? 0x10000u : hi);
    }
    return 0;
}

}  // extern "C"
