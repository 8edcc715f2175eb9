/* ORACLE (test infrastructure only).
 *
 * Plain-C restatement of the arithmetic coder AIVC reaches through the third-party
 * `torchac` package (`import torchac`, src/real_life/bitstream.py:10; installed
 * un-pinned by `pip install torchac`, README.md:140 -- latest upstream at the time: 0.9.3).
 * Call sites restated: bitstream.py:281 (encode_float_cdf), :454 and :482
 * (decode_float_cdf).  torchac is NOT in /root/reference and not installable here,
 * so this follows the published algorithm of its backend (torchac_backend.cpp):
 * a 32-bit low/high binary arithmetic coder with 16-bit CDF precision, MSB
 * renormalisation, a pending-bit (underflow) counter, bits packed MSB-first, a final
 * disambiguating bit and zero padding to a byte boundary; the decoder reads missing
 * trailing bits as zeros and finds each symbol by binary search over its Lp-entry CDF.
 *
 *   parity unpinned: no torchac golden bitstream exists in the reference tree.
 *   What IS pinned by tests: self-consistency (decode(encode(x)) == x), hand-derived
 *   known answers (tests/test_rangecoder.py) and equality with the product coder.
 *
 * The CDF is given either as a full table (cdf[i*Lp + s], uint16, "pmf"/generic mode)
 * or, for encoding only, as per-symbol (c_low, c_high) pairs.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    uint8_t *buf;
    size_t len, cap;
    uint8_t cache;
    int count;
} bitout_t;

static void bo_append(bitout_t *o, int bit) {
    o->cache = (uint8_t)((o->cache << 1) | (bit & 1));
    o->count++;
    if (o->count == 8) {
        if (o->len == o->cap) {
            o->cap = o->cap ? 2 * o->cap : 1024;
            o->buf = (uint8_t *)realloc(o->buf, o->cap);
        }
        o->buf[o->len++] = o->cache;
        o->count = 0;
        o->cache = 0;
    }
}

static void bo_bit_and_pending(bitout_t *o, int bit, uint64_t *pending) {
    bo_append(o, bit);
    while (*pending > 0) {
        bo_append(o, !bit);
        (*pending)--;
    }
}

static void bo_flush(bitout_t *o) {
    while (o->count != 0) bo_append(o, 0);
}

/* one coding step shared by both entry points */
static void enc_step(bitout_t *o, uint32_t *low, uint32_t *high, uint64_t *pending,
                     uint32_t c_low, uint32_t c_high) {
    const uint64_t span = (uint64_t)(*high) - (uint64_t)(*low) + 1;
    *high = (*low - 1) + (uint32_t)((span * (uint64_t)c_high) >> 16);
    *low = (*low) + (uint32_t)((span * (uint64_t)c_low) >> 16);
    for (;;) {
        if (*high < 0x80000000U) {
            bo_bit_and_pending(o, 0, pending);
            *low <<= 1;
            *high = (*high << 1) | 1;
        } else if (*low >= 0x80000000U) {
            bo_bit_and_pending(o, 1, pending);
            *low <<= 1;
            *high = (*high << 1) | 1;
        } else if (*low >= 0x40000000U && *high < 0xC0000000U) {
            (*pending)++;
            *low = (*low << 1) & 0x7FFFFFFFU;
            *high = (*high << 1) | 0x80000001U;
        } else {
            break;
        }
    }
}

static size_t enc_finish(bitout_t *o, uint32_t low, uint64_t pending, uint8_t **out) {
    pending += 1;
    bo_bit_and_pending(o, low < 0x40000000U ? 0 : 1, &pending);
    bo_flush(o);
    *out = o->buf;
    return o->len;
}

/* Encode n symbols against a full CDF table. Returns byte count; *out is malloc'ed. */
size_t tac_ref_encode_table(const uint16_t *cdf, int lp, const int16_t *sym, size_t n,
                            uint8_t **out) {
    bitout_t o = {0};
    uint32_t low = 0, high = 0xFFFFFFFFU;
    uint64_t pending = 0;
    const int max_symbol = lp - 2;
    for (size_t i = 0; i < n; ++i) {
        const int s = sym[i];
        const uint32_t c_low = cdf[i * (size_t)lp + s];
        const uint32_t c_high = (s == max_symbol) ? 0x10000U : cdf[i * (size_t)lp + s + 1];
        enc_step(&o, &low, &high, &pending, c_low, c_high);
    }
    return enc_finish(&o, low, pending, out);
}

/* Same coder, CDF supplied as per-symbol bounds (c_high as uint32: may be 0x10000). */
size_t tac_ref_encode_bounds(const uint32_t *c_low, const uint32_t *c_high, size_t n,
                             uint8_t **out) {
    bitout_t o = {0};
    uint32_t low = 0, high = 0xFFFFFFFFU;
    uint64_t pending = 0;
    for (size_t i = 0; i < n; ++i) enc_step(&o, &low, &high, &pending, c_low[i], c_high[i]);
    return enc_finish(&o, low, pending, out);
}

void tac_ref_free(uint8_t *p) { free(p); }

typedef struct {
    const uint8_t *in;
    size_t len, pos;
    uint8_t cache;
    int cached_bits;
} bitin_t;

static void bi_get(bitin_t *b, uint32_t *value) {
    if (b->cached_bits == 0) {
        if (b->pos == b->len) {
            *value <<= 1;
            return;
        }
        b->cache = b->in[b->pos++];
        b->cached_bits = 8;
    }
    *value = (*value << 1) | ((b->cache >> (b->cached_bits - 1)) & 1);
    b->cached_bits--;
}

static uint16_t binsearch(const uint16_t *cdf, uint16_t target, uint16_t max_sym) {
    uint16_t left = 0, right = (uint16_t)(max_sym + 1);
    while (left + 1 < right) {
        const uint16_t m = (uint16_t)((left + right) / 2);
        const uint16_t v = cdf[m];
        if (v < target) left = m;
        else if (v > target) right = m;
        else return m;
    }
    return left;
}

/* Decode n symbols against a full CDF table. */
void tac_ref_decode_table(const uint16_t *cdf, int lp, const uint8_t *in, size_t in_len,
                          int16_t *sym, size_t n) {
    bitin_t b = {in, in_len, 0, 0, 0};
    uint32_t low = 0, high = 0xFFFFFFFFU, value = 0;
    const int max_symbol = lp - 2;
    for (int i = 0; i < 32; ++i) bi_get(&b, &value);
    for (size_t i = 0; i < n; ++i) {
        const uint64_t span = (uint64_t)high - (uint64_t)low + 1;
        const uint16_t count =
            (uint16_t)((((uint64_t)value - (uint64_t)low + 1) * 0x10000ULL - 1) / span);
        const uint16_t *row = cdf + i * (size_t)lp;
        const int s = binsearch(row, count, (uint16_t)max_symbol);
        sym[i] = (int16_t)s;
        if (i == n - 1) break;
        const uint32_t c_low = row[s];
        const uint32_t c_high = (s == max_symbol) ? 0x10000U : row[s + 1];
        high = (low - 1) + (uint32_t)((span * (uint64_t)c_high) >> 16);
        low = low + (uint32_t)((span * (uint64_t)c_low) >> 16);
        for (;;) {
            if (low >= 0x80000000U || high < 0x80000000U) {
                low <<= 1;
                high = (high << 1) | 1;
                bi_get(&b, &value);
            } else if (low >= 0x40000000U && high < 0xC0000000U) {
                low = (low << 1) & 0x7FFFFFFFU;
                high = (high << 1) | 0x80000001U;
                value -= 0x40000000U;
                bi_get(&b, &value);
            } else {
                break;
            }
        }
    }
}
