// hostz.cpp -- host-side DEFLATE decoder and CRC-32 for the BGZF readers (see hostz.h).
//
// The host decode of a BAM file is the end-to-end bottleneck of `cmd coverage`: the device bins 5e11
// reads/s, zlib inflates about 1 GB/s per core and spends another quarter of that time on the CRC.  A BGZF block is
// at most 64 KiB with its size known beforehand, which is the easy case for a decoder: one output buffer, no
// window to wrap, no streaming state.  So: a 64-bit bit buffer refilled by unaligned loads, one table lookup per
// symbol (10-bit primary table, subtables behind it for the longer codes), literals in runs, matches copied eight
// bytes at a time -- and a careful byte-wise loop for the last 300 bytes of the output.  Whatever this decoder does
// not like, it refuses: the caller repeats the block with zlib, and every block is checked against its CRC-32
// either way, so a bug here can cost time but not correctness.
#include "hostz.h"

#include <immintrin.h>
#include <zlib.h>

#include <cstring>

namespace cnvb {
namespace {

// ---- CRC-32 -----------------------------------------------------------------------------------------
// Folding with carry-less multiplication (Gopal, Ozturk, Guilford et al., "Fast CRC Computation for Generic
// Polynomials Using PCLMULQDQ Instruction", Intel 2009), bit-reflected form for polynomial 0x04C11DB7: four 128-bit
// lanes are folded across 64 bytes per step (x^(512+64), x^512 mod P), then to one lane (x^(128+64), x^128), then
// reduced 128 -> 64 -> 32 bits (Barrett).  crc is the running register (zlib's value, inverted).
__attribute__((target("pclmul,sse4.1"))) uint32_t crc32_clmul(const uint8_t *buf, size_t len, uint32_t crc) {
    alignas(16) static const uint64_t k1k2[2] = {0x0154442bd4ull, 0x01c6e41596ull};
    alignas(16) static const uint64_t k3k4[2] = {0x01751997d0ull, 0x00ccaa009eull};
    alignas(16) static const uint64_t k5k0[2] = {0x0163cd6124ull, 0x0000000000ull};
    alignas(16) static const uint64_t poly[2] = {0x01db710641ull, 0x01f7011641ull};
    __m128i x0, x1, x2, x3, x4, x5, x6, x7, x8, y5, y6, y7, y8;
    x1 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(buf + 0x00));
    x2 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(buf + 0x10));
    x3 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(buf + 0x20));
    x4 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(buf + 0x30));
    x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)crc));
    x0 = _mm_load_si128(reinterpret_cast<const __m128i *>(k1k2));
    buf += 64;
    len -= 64;
    while (len >= 64) {
        x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
        x6 = _mm_clmulepi64_si128(x2, x0, 0x00);
        x7 = _mm_clmulepi64_si128(x3, x0, 0x00);
        x8 = _mm_clmulepi64_si128(x4, x0, 0x00);
        x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
        x2 = _mm_clmulepi64_si128(x2, x0, 0x11);
        x3 = _mm_clmulepi64_si128(x3, x0, 0x11);
        x4 = _mm_clmulepi64_si128(x4, x0, 0x11);
        y5 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(buf + 0x00));
        y6 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(buf + 0x10));
        y7 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(buf + 0x20));
        y8 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(buf + 0x30));
        x1 = _mm_xor_si128(_mm_xor_si128(x1, x5), y5);
        x2 = _mm_xor_si128(_mm_xor_si128(x2, x6), y6);
        x3 = _mm_xor_si128(_mm_xor_si128(x3, x7), y7);
        x4 = _mm_xor_si128(_mm_xor_si128(x4, x8), y8);
        buf += 64;
        len -= 64;
    }
    x0 = _mm_load_si128(reinterpret_cast<const __m128i *>(k3k4));
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x3), x5);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x4), x5);
    while (len >= 16) {
        x2 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(buf));
        x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
        x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
        x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
        buf += 16;
        len -= 16;
    }
    x2 = _mm_clmulepi64_si128(x1, x0, 0x10);
    x3 = _mm_setr_epi32(~0, 0, ~0, 0);
    x1 = _mm_srli_si128(x1, 8);
    x1 = _mm_xor_si128(x1, x2);
    x0 = _mm_loadl_epi64(reinterpret_cast<const __m128i *>(k5k0));
    x2 = _mm_srli_si128(x1, 4);
    x1 = _mm_and_si128(x1, x3);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_xor_si128(x1, x2);
    x0 = _mm_load_si128(reinterpret_cast<const __m128i *>(poly));
    x2 = _mm_and_si128(x1, x3);
    x2 = _mm_clmulepi64_si128(x2, x0, 0x10);
    x2 = _mm_and_si128(x2, x3);
    x2 = _mm_clmulepi64_si128(x2, x0, 0x00);
    x1 = _mm_xor_si128(x1, x2);
    return (uint32_t)_mm_extract_epi32(x1, 1);
}

// ---- DEFLATE ----------------------------------------------------------------------------------------
// table entry: value << 16 | kind << 12 | extra bits << 8 | bits to drop (the code; for a length / distance entry the
// code and its extra bits together)
//   literal: value = the byte;  length / distance: value = base;  subtable: value = its offset, extra = its index bits
// (Entries holding TWO literals where both codes fit the primary index were built and measured: no gain on BAM-like
// data -- a sequencer's bases and binned qualities come out of zlib as short matches, not as runs of literals -- and a
// loss on text, where the pairing pass over the table costs more than it saves.  Removed.)
// (kinds are one bit each so that the hot loop tests them with one instruction; kBad has none)
enum : uint32_t { kLit = 1, kBase = 2, kEnd = 4, kSub = 8, kBad = 0 };
constexpr uint32_t entry(uint32_t value, uint32_t kind, uint32_t extra, uint32_t len) {
    return value << 16 | kind << 12 | extra << 8 | len;
}
constexpr uint32_t kIsLit = kLit << 12, kIsBase = kBase << 12, kIsEnd = kEnd << 12, kIsSub = kSub << 12;
constexpr uint32_t kLitBits = 11, kDistBits = 8;
constexpr uint32_t kLitSize = (1u << kLitBits) + 288 * 16, kDistSize = (1u << kDistBits) + 32 * 128;

const uint16_t kLenBase[29] = {3,  4,  5,  6,  7,  8,  9,  10, 11,  13,  15,  17,  19,  23, 27,
                               31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kDistBase[30] = {1,   2,   3,   4,   5,   7,    9,    13,   17,   25,   33,   49,   65,    97,    129,
                                193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

inline uint32_t reverse_bits(uint32_t code, uint32_t len) {
    uint32_t r = 0;
    for (uint32_t i = 0; i < len; ++i) r |= ((code >> i) & 1u) << (len - 1 - i);
    return r;
}

// what a symbol of the alphabet decodes to (without its code length)
inline uint32_t litlen_entry(uint32_t sym) {
    if (sym < 256) return entry(sym, kLit, 0, 0);
    if (sym == 256) return entry(0, kEnd, 0, 0);
    if (sym < 286) return entry(kLenBase[sym - 257], kBase, kLenExtra[sym - 257], 0);
    return entry(0, kBad, 0, 0);
}
inline uint32_t dist_entry(uint32_t sym) {
    if (sym < 30) return entry(kDistBase[sym], kBase, kDistExtra[sym], 0);
    return entry(0, kBad, 0, 0);
}

// Canonical Huffman code of n symbols with the given lengths (0 = unused, <= 15) -> lookup table with `bits` primary
// index bits.  false: over-subscribed code.  An incomplete code leaves kBad entries (decoding one refuses the block).
template <typename F>
bool build_table(const uint8_t *lens, uint32_t n, uint32_t bits, uint32_t *table, uint32_t table_size, F sym_entry) {
    uint32_t count[16] = {0}, next[16];
    for (uint32_t s = 0; s < n; ++s) ++count[lens[s]];
    count[0] = 0;
    uint32_t code = 0, left = 1;
    for (uint32_t l = 1; l <= 15; ++l) {
        left <<= 1;
        if (count[l] > left) return false;
        left -= count[l];
        code = (code + count[l - 1]) << 1;
        next[l] = code;
    }
    const uint32_t primary = 1u << bits;
    for (uint32_t i = 0; i < primary; ++i) table[i] = entry(0, kBad, 0, 1);
    // subtables: one per primary prefix that longer codes share, as wide as the longest of them needs
    uint8_t sub_bits[1u << kLitBits];  // (the widest primary table)
    uint32_t used = primary;
    bool have_long = false;
    for (uint32_t l = bits + 1; l <= 15; ++l) have_long = have_long || count[l];
    if (have_long) {
        memset(sub_bits, 0, primary);
        uint32_t nx[16];
        memcpy(nx, next, sizeof nx);
        for (uint32_t s = 0; s < n; ++s) {
            const uint32_t l = lens[s];
            if (l <= bits) {
                if (l) ++nx[l];
                continue;
            }
            const uint32_t prefix = reverse_bits(nx[l]++, l) & (primary - 1);
            if (l - bits > sub_bits[prefix]) sub_bits[prefix] = (uint8_t)(l - bits);
        }
        for (uint32_t p = 0; p < primary; ++p)
            if (sub_bits[p]) {
                if (used + (1u << sub_bits[p]) > table_size) return false;
                table[p] = entry(used, kSub, sub_bits[p], 0);
                for (uint32_t i = 0; i < (1u << sub_bits[p]); ++i) table[used + i] = entry(0, kBad, 0, 15);
                used += 1u << sub_bits[p];
            }
    }
    for (uint32_t s = 0; s < n; ++s) {
        const uint32_t l = lens[s];
        if (!l) continue;
        const uint32_t rev = reverse_bits(next[l]++, l);
        const uint32_t se = sym_entry(s);
        const uint32_t e = se | (l + ((se & kIsBase) ? ((se >> 8) & 15u) : 0u));  // base entries: code + extra bits
        if (l <= bits) {
            for (uint32_t i = rev; i < primary; i += 1u << l) table[i] = e;
        } else {
            const uint32_t sub = table[rev & (primary - 1)];
            const uint32_t base = sub >> 16, sb = (sub >> 8) & 15u;
            for (uint32_t i = rev >> bits; i < (1u << sb); i += 1u << (l - bits)) table[base + i] = e;
        }
    }
    return true;
}

struct Bits {
    const uint8_t *ip, *in, *in_end;
    uint64_t bb = 0;
    uint32_t bc = 0;
    // after this: at least 56 valid bits (zeros or the slack's bytes past the end of the input: the consumed
    // length is checked when the stream ends).  false: further past the end than the slack allows.
    inline bool refill() {
        if (ip > in_end + 8) return false;
        uint64_t w;
        memcpy(&w, ip, 8);
        bb |= w << bc;
        ip += (63 - bc) >> 3;
        bc |= 56;
        return true;
    }
    inline void refill_unchecked() {  // (the caller knows that ip + 8 <= in_end)
        uint64_t w;
        memcpy(&w, ip, 8);
        bb |= w << bc;
        ip += (63 - bc) >> 3;
        bc |= 56;
    }
    inline uint32_t peek(uint32_t n) const { return (uint32_t)bb & ((1u << n) - 1u); }
    inline void drop(uint32_t n) {
        bb >>= n;
        bc -= n;
    }
    inline uint32_t take(uint32_t n) {
        const uint32_t v = peek(n);
        drop(n);
        return v;
    }
    inline size_t bits_used() const { return (size_t)(ip - in) * 8 - bc; }
};

// the extra bits of a length / distance entry: they follow its code in the bit buffer
inline uint32_t extra_bits(uint64_t bb, uint32_t e) {
    const uint32_t x = (e >> 8) & 15u;
    return (uint32_t)(bb >> ((e & 255u) - x)) & ((1u << x) - 1u);
}

inline uint32_t lookup(const uint32_t *table, uint32_t bits, uint64_t bb) {
    uint32_t e = table[(uint32_t)bb & ((1u << bits) - 1u)];
    if (__builtin_expect(e & kIsSub, 0)) e = table[(e >> 16) + (((uint32_t)(bb >> bits)) & ((1u << ((e >> 8) & 15u)) - 1u))];
    return e;
}

}  // namespace

uint32_t crc32_fast(const uint8_t *p, size_t n) {
    static const bool have_clmul = __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1");
    uint32_t crc = 0;
    if (have_clmul && n >= 64) {
        const size_t body = n & ~(size_t)15;
        crc = ~crc32_clmul(p, body, ~crc);
        p += body;
        n -= body;
    }
    while (n) {  // (zlib takes uInt lengths)
        const uInt step = n > (1u << 30) ? (1u << 30) : (uInt)n;
        crc = (uint32_t)::crc32(crc, p, step);
        p += step;
        n -= step;
    }
    return crc;
}

namespace {
// (always inlined into the two entry points below, so that one of them is compiled with BMI2: the variable shifts of
// the bit buffer are SHRX / SHLX there instead of the three-uop shifts by CL -- 8 to 15 % of the decoder's time.  A
// shuffle-based copy for matches less than 16 back was measured too: they are 0.5 % of a BAM file's matches.)
__attribute__((always_inline)) inline bool inflate_impl(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len) {
    Bits b;
    b.ip = b.in = in;
    b.in_end = in + in_len;
    uint8_t *op = out, *const out_end = out + out_len;
    uint32_t lit[kLitSize], dist[kDistSize];
    uint8_t lens[288 + 32];
    bool last = false;
    while (!last) {
        if (!b.refill()) return false;
        last = b.take(1);
        const uint32_t type = b.take(2);
        if (type == 0) {  // stored: the rest of the byte is skipped, LEN, ~LEN, the bytes
            b.drop(b.bc & 7u);
            if (!b.refill()) return false;
            const uint32_t len = b.take(16), nlen = b.take(16);
            if ((len ^ nlen) != 0xFFFFu) return false;
            const uint8_t *src = b.ip - (b.bc >> 3);  // bytes still in the bit buffer go back
            if (src > b.in_end || len > (size_t)(b.in_end - src) || len > (size_t)(out_end - op)) return false;
            memcpy(op, src, len);
            op += len;
            b.ip = src + len;
            b.bb = 0;
            b.bc = 0;
            continue;
        }
        if (type == 3) return false;
        if (type == 1) {  // fixed code (RFC 1951 3.2.6)
            memset(lens, 8, 144);
            memset(lens + 144, 9, 112);
            memset(lens + 256, 7, 24);
            memset(lens + 280, 8, 8);
            memset(lens + 288, 5, 32);
            if (!build_table(lens, 288, kLitBits, lit, kLitSize, litlen_entry)) return false;
            if (!build_table(lens + 288, 32, kDistBits, dist, kDistSize, dist_entry)) return false;
        } else {  // dynamic code (3.2.7)
            const uint32_t hlit = b.take(5) + 257, hdist = b.take(5) + 1, hclen = b.take(4) + 4;
            if (hlit > 286 || hdist > 30) return false;
            static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
            uint8_t pre_lens[19] = {0};
            for (uint32_t i = 0; i < hclen; ++i) {
                if (b.bc < 3 && !b.refill()) return false;
                pre_lens[order[i]] = (uint8_t)b.take(3);
            }
            uint32_t pre[1u << 7];
            if (!build_table(pre_lens, 19, 7, pre, 1u << 7, [](uint32_t s) { return entry(s, kLit, 0, 0); })) return false;
            uint32_t n = 0;
            while (n < hlit + hdist) {
                if (!b.refill()) return false;
                const uint32_t e = pre[b.peek(7)];
                if (!(e & kIsLit)) return false;
                b.drop(e & 255u);
                const uint32_t sym = e >> 16;
                if (sym < 16) {
                    lens[n++] = (uint8_t)sym;
                    continue;
                }
                uint32_t rep, val = 0;
                if (sym == 16) {
                    if (n == 0) return false;
                    val = lens[n - 1];
                    rep = 3 + b.take(2);
                } else if (sym == 17) {
                    rep = 3 + b.take(3);
                } else {
                    rep = 11 + b.take(7);
                }
                if (n + rep > hlit + hdist) return false;
                memset(lens + n, (int)val, rep);
                n += rep;
            }
            if (lens[256] == 0) return false;
            uint8_t dl[32];
            memcpy(dl, lens + hlit, hdist);
            if (!build_table(lens, hlit, kLitBits, lit, kLitSize, litlen_entry)) return false;
            if (!build_table(dl, hdist, kDistBits, dist, kDistSize, dist_entry)) return false;
        }
        // ---- the symbols of the block
        for (;;) {
            // fast path: room for the longest match plus the copy's overshoot, input left for two refills
            while ((size_t)(out_end - op) >= 258 + 32 + 3 && b.ip + 16 <= b.in_end) {
                b.refill_unchecked();
                uint32_t e = lookup(lit, kLitBits, b.bb);
                // up to three literals on one refill (3 x 15 bits), then the bits a match needs (15 + 5 + 15 + 13)
                if (e & kIsLit) {
                    *op++ = (uint8_t)(e >> 16);
                    b.drop(e & 255u);
                    e = lookup(lit, kLitBits, b.bb);
                    if (e & kIsLit) {
                        *op++ = (uint8_t)(e >> 16);
                        b.drop(e & 255u);
                        e = lookup(lit, kLitBits, b.bb);
                        if (e & kIsLit) {
                            *op++ = (uint8_t)(e >> 16);
                            b.drop(e & 255u);
                            continue;
                        }
                    }
                    b.refill_unchecked();
                }
                if (__builtin_expect(!(e & kIsBase), 0)) {
                    if (e & kIsEnd) {
                        b.drop(e & 255u);
                        goto block_done;
                    }
                    return false;
                }
                // (an entry's low byte counts the code AND its extra bits: one shift on the bit buffer's dependency
                // chain; the extra bits are cut out of a copy beside it)
                const uint32_t len = (e >> 16) + extra_bits(b.bb, e);
                b.drop(e & 255u);
                const uint32_t d = lookup(dist, kDistBits, b.bb);
                if (__builtin_expect(!(d & kIsBase), 0)) return false;
                const uint32_t back = (d >> 16) + extra_bits(b.bb, d);
                b.drop(d & 255u);
                if (__builtin_expect(back > (size_t)(op - out), 0)) return false;
                const uint8_t *src = op - back;
                uint8_t *dst = op;
                op += len;
                if (back >= 16) {  // sixteen bytes at a time, up to 15 past the match (the room is there)
                    do {
                        _mm_storeu_si128(reinterpret_cast<__m128i *>(dst), _mm_loadu_si128(reinterpret_cast<const __m128i *>(src)));
                        src += 16;
                        dst += 16;
                    } while (dst < op);
                } else if (back >= 8) {
                    do {
                        uint64_t w;
                        memcpy(&w, src, 8);
                        memcpy(dst, &w, 8);
                        src += 8;
                        dst += 8;
                    } while (dst < op);
                } else if (back == 1) {
                    memset(dst, *src, len);
                } else {
                    do *dst++ = *src++; while (dst < op);
                }
            }
            // careful path: one symbol with every bound checked
            if (!b.refill()) return false;
            const uint32_t e = lookup(lit, kLitBits, b.bb);
            if (e & kIsLit) {
                if (op == out_end) return false;
                *op++ = (uint8_t)(e >> 16);
                b.drop(e & 255u);
                continue;
            }
            if (e & kIsEnd) {
                b.drop(e & 255u);
                break;
            }
            if (!(e & kIsBase)) return false;
            const uint32_t len = (e >> 16) + extra_bits(b.bb, e);
            b.drop(e & 255u);
            const uint32_t d = lookup(dist, kDistBits, b.bb);
            if (!(d & kIsBase)) return false;
            const uint32_t back = (d >> 16) + extra_bits(b.bb, d);
            b.drop(d & 255u);
            if (back > (size_t)(op - out) || len > (size_t)(out_end - op)) return false;
            const uint8_t *src = op - back;
            for (uint32_t i = 0; i < len; ++i) op[i] = src[i];
            op += len;
        }
    block_done:;
    }
    // all of the output, and the input used up to its last byte (no bits from past its end)
    return op == out_end && (b.bits_used() + 7) / 8 == in_len;
}
__attribute__((target("bmi2"))) bool inflate_bmi2(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len) {
    return inflate_impl(in, in_len, out, out_len);
}
bool inflate_plain(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len) {
    return inflate_impl(in, in_len, out, out_len);
}
}  // namespace

bool inflate_raw_fast(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len) {
    static const bool have_bmi2 = __builtin_cpu_supports("bmi2");
    return have_bmi2 ? inflate_bmi2(in, in_len, out, out_len) : inflate_plain(in, in_len, out, out_len);
}

}  // namespace cnvb
