// inflate_fast.hpp -- a DEFLATE (RFC 1951) / zlib-stream (RFC 1950) decoder for the MINC 2 reader's chunks.
//
// Why: reading a study FROM FILES is inflate-bound (one zlib thread turns a 180x252x156 uint16 volume around in
// 0.08-0.13 s, and every other step of mincLm from files is faster than that), and zlib's inflate decodes one symbol per
// iteration through a 32-bit bit buffer.  This decoder keeps 64 bits of input in a register (one unaligned 8-byte load per refill),
// resolves literals / lengths through an 11-bit primary table and distances through an 8-bit one (second-level tables for longer
// codes), emits up to three literals per refill and copies matches 8 bytes at a time.  It is written from the RFCs; the test suite
// checks it against zlib's own compressor output at every level and strategy, on noise, images, runs and empty input, and against
// truncated / damaged streams (tests/test_minc2_reader.py).  The reader falls back to zlib's uncompress if this decoder
// reports an error, so a stream it cannot handle costs time, never correctness.
// Measured and dropped: a 12-bit primary table whose entries carry two literals where both codes fit the index -- the per-block
// pairing pass and the wider entries cost more than the saved look-ups (154 against 182 MB/s on noise-like uint16 samples).
#pragma once
#include <cstdint>
#include <cstring>

namespace rmg_inflate {

constexpr int kLitBits = 11, kDistBits = 8;
constexpr int kMaxLitEntries = (1 << kLitBits) + 1400;      // primary + second-level tables (generous bound)
constexpr int kMaxDistEntries = (1 << kDistBits) + 600;

// table entry: bits 0..7 = code length to consume (primary) / sub-table index bits (pointer); bits 8..15 = kind + extra bits;
// bits 16..31 = literal / base value / sub-table offset
constexpr uint32_t kLiteral = 0x8000, kEob = 0x4000, kSub = 0x2000, kInvalid = 0x1000;

struct Tables {
    uint32_t lit[kMaxLitEntries];
    uint32_t dist[kMaxDistEntries];
};

static const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

inline uint32_t reverse_bits(uint32_t v, int n)
{
    uint32_t r = 0;
    for (int i = 0; i < n; ++i) { r = (r << 1) | (v & 1); v >>= 1; }
    return r;
}

// Canonical Huffman code -> lookup tables.  sym_entry(s) gives the payload (kind, extra bits, value) of symbol s.
// Returns false for an over-subscribed code, or an incomplete one with more than one symbol (zlib's rule: a single code of length
// 1 is allowed for distances).
template <typename F>
bool build_table(const uint8_t *lens, int nsym, int primary_bits, uint32_t *table, int max_entries, F sym_entry)
{
    int count[16] = {0};
    for (int s = 0; s < nsym; ++s) count[lens[s]]++;
    count[0] = 0;
    int left = 1, used = 0, maxlen = 0;
    for (int l = 1; l <= 15; ++l) {
        left = (left << 1) - count[l];
        if (left < 0) return false;
        used += count[l];
        if (count[l]) maxlen = l;
    }
    if (left > 0 && !(used == 1 && maxlen == 1) && used != 0) return false;
    uint32_t next[16];
    uint32_t code = 0;
    for (int l = 1; l <= 15; ++l) { code = (code + count[l - 1]) << 1; next[l] = code; }
    const int psize = 1 << primary_bits;
    for (int i = 0; i < psize; ++i) table[i] = kInvalid | 1;      // holes of an incomplete code
    int sub_end = psize;
    // second-level tables: one per primary prefix that has longer codes; size = 2^(longest code under the prefix - primary_bits)
    int sub_bits[1 << kLitBits];
    if (maxlen > primary_bits) {
        for (int i = 0; i < psize; ++i) sub_bits[i] = 0;
        uint32_t nc[16];
        std::memcpy(nc, next, sizeof nc);
        for (int s = 0; s < nsym; ++s) {
            const int l = lens[s];
            if (l <= primary_bits) { if (l) nc[l]++; continue; }
            const uint32_t rev = reverse_bits(nc[l]++, l);
            const int prefix = rev & (psize - 1);
            if (l - primary_bits > sub_bits[prefix]) sub_bits[prefix] = l - primary_bits;
        }
        for (int i = 0; i < psize; ++i) {
            if (!sub_bits[i]) continue;
            if (sub_end + (1 << sub_bits[i]) > max_entries) return false;
            table[i] = ((uint32_t)sub_end << 16) | kSub | ((uint32_t)sub_bits[i] << 8) | (uint32_t)primary_bits;
            for (int k = 0; k < (1 << sub_bits[i]); ++k) table[sub_end + k] = kInvalid | 1;
            sub_end += 1 << sub_bits[i];
        }
    }
    for (int s = 0; s < nsym; ++s) {
        const int l = lens[s];
        if (!l) continue;
        const uint32_t rev = reverse_bits(next[l]++, l);
        const uint32_t payload = sym_entry(s);
        if (l <= primary_bits) {
            for (uint32_t i = rev; i < (uint32_t)psize; i += 1u << l) table[i] = payload | (uint32_t)l;
        } else {
            const int prefix = rev & (psize - 1);
            const uint32_t ptr = table[prefix];
            const int sb = (ptr >> 8) & 0x1f, base = ptr >> 16;
            const int rl = l - primary_bits;
            for (uint32_t i = rev >> primary_bits; i < (1u << sb); i += 1u << rl) table[base + i] = payload | (uint32_t)rl;
        }
    }
    return true;
}

inline uint32_t lit_entry(int s)
{
    if (s < 256) return ((uint32_t)s << 16) | kLiteral;
    if (s == 256) return kEob;
    if (s > 285) return kInvalid;
    return ((uint32_t)kLenBase[s - 257] << 16) | ((uint32_t)kLenExtra[s - 257] << 8);
}
inline uint32_t dist_entry(int s)
{
    if (s > 29) return kInvalid;
    return ((uint32_t)kDistBase[s] << 16) | ((uint32_t)kDistExtra[s] << 8);
}

struct BitReader {
    const uint8_t *in, *end;
    uint64_t buf = 0;
    int cnt = 0;
    int overrun = 0;       // zero bytes supplied past the end of the input
    inline void refill()
    {
        if (end - in >= 8) {
            uint64_t w;
            std::memcpy(&w, in, 8);
            buf |= w << cnt;
            in += (63 - cnt) >> 3;
            cnt |= 56;
        } else {
            while (cnt <= 56) {
                if (in < end) buf |= (uint64_t)*in++ << cnt;
                else ++overrun;
                cnt += 8;
            }
        }
    }
    inline uint32_t peek(int n) const { return (uint32_t)(buf & ((1ull << n) - 1)); }
    inline void drop(int n) { buf >>= n; cnt -= n; }
    inline uint32_t take(int n) { const uint32_t v = peek(n); drop(n); return v; }
};

// One block's symbols.  FAST: the caller guarantees >= 16 input bytes and >= 3 maximal matches of output space; the loop
// re-checks both every iteration and returns 1 ("go on in the other mode") when they no longer hold.  Returns 0 at the
// end-of-block code, -1 on a malformed stream or a full output buffer.
template <bool FAST>
inline int decode_symbols(BitReader &br_, uint8_t *&out_, uint8_t *dst, uint8_t *out_end, const uint32_t *__restrict lit,
                          const uint32_t *__restrict dtab)
{
    BitReader br = br_;
    uint8_t *out = out_;
    int rc;
    for (;;) {
        if (FAST) {
            if (br.end - br.in < 16 || out_end - out < 3 * 258 + 16) { rc = 1; break; }
            uint64_t w;
            std::memcpy(&w, br.in, 8);
            br.buf |= w << br.cnt;
            br.in += (63 - br.cnt) >> 3;
            br.cnt |= 56;
        } else {
            br.refill();
        }
        uint32_t e = lit[br.peek(kLitBits)];
        if (e & kLiteral) {
            // up to three literals per refill (a primary-table literal takes at most 11 bits; >= 56 are in the buffer)
            if (!FAST && out_end - out < 3) {
                if (out >= out_end) { rc = -1; break; }
                *out++ = (uint8_t)(e >> 16);
                br.drop(e & 0xff);
                continue;
            }
            *out++ = (uint8_t)(e >> 16);
            br.drop(e & 0xff);
            e = lit[br.peek(kLitBits)];
            if (e & kLiteral) {
                *out++ = (uint8_t)(e >> 16);
                br.drop(e & 0xff);
                e = lit[br.peek(kLitBits)];
                if (e & kLiteral) {
                    *out++ = (uint8_t)(e >> 16);
                    br.drop(e & 0xff);
                    continue;
                }
            }
            // e is a non-literal entry looked up with >= 34 bits left: handled below
        }
        if (e & kSub) {
            br.drop(e & 0xff);
            e = lit[(e >> 16) + br.peek((e >> 8) & 0x1f)];
            if (e & kLiteral) {
                if (!FAST && out >= out_end) { rc = -1; break; }
                *out++ = (uint8_t)(e >> 16);
                br.drop(e & 0xff);
                continue;
            }
        }
        if (e & kInvalid) { rc = -1; break; }
        br.drop(e & 0xff);
        if (e & kEob) { rc = 0; break; }
        // length (<= 5 extra bits), then refill for the distance code (<= 15 bits) and its extra bits (<= 13)
        const uint32_t len = (e >> 16) + br.take((e >> 8) & 0x1f);
        br.refill();
        uint32_t d = dtab[br.peek(kDistBits)];
        if (d & kSub) {
            br.drop(d & 0xff);
            d = dtab[(d >> 16) + br.peek((d >> 8) & 0x1f)];
        }
        if (d & kInvalid) { rc = -1; break; }
        br.drop(d & 0xff);
        const uint32_t dist = (d >> 16) + br.take((d >> 8) & 0x1f);
        if (dist > (size_t)(out - dst) || (!FAST && len > (size_t)(out_end - out))) { rc = -1; break; }
        const uint8_t *from = out - dist;
        if (dist >= 8 && (FAST || (size_t)(out_end - out) >= len + 8)) {
            uint8_t *o = out;
            const uint8_t *f = from;
            uint8_t *stop = out + len;
            do {
                uint64_t w8;
                std::memcpy(&w8, f, 8);
                std::memcpy(o, &w8, 8);
                o += 8;
                f += 8;
            } while (o < stop);
            out = stop;
        } else if (dist == 1) {
            std::memset(out, *from, len);
            out += len;
        } else {
            for (uint32_t k = 0; k < len; ++k) out[k] = from[k];
            out += len;
        }
    }
    br_ = br;
    out_ = out;
    return rc;
}

// Raw DEFLATE.  Returns the number of bytes written, or -1 on a malformed / truncated stream or when dst is too small.
// *consumed (nullable) receives the number of input bytes used.
inline int64_t inflate_raw(const uint8_t *src, size_t src_len, uint8_t *dst, size_t dst_cap, size_t *consumed, Tables &T)
{
    BitReader br{src, src + src_len};
    uint8_t *out = dst, *const out_end = dst + dst_cap;
    for (;;) {
        br.refill();
        const uint32_t final = br.take(1), type = br.take(2);
        if (type == 0) {                                        // stored block
            br.drop(br.cnt & 7);
            br.refill();
            const uint32_t len = br.take(16), nlen = br.take(16);
            if ((len ^ nlen) != 0xffff || br.overrun * 8 > br.cnt) return -1;
            // give the whole (real) bytes still in the bit buffer back to the input
            const uint8_t *p = br.in - ((br.cnt - 8 * br.overrun) >> 3);
            br.buf = 0;
            br.cnt = 0;
            br.overrun = 0;
            if ((size_t)(br.end - p) < len || (size_t)(out_end - out) < len) return -1;
            std::memcpy(out, p, len);
            out += len;
            br.in = p + len;
        } else if (type == 1 || type == 2) {
            uint8_t lens[288 + 32];
            int nlit, ndist;
            if (type == 1) {
                nlit = 288;
                ndist = 32;                                      // 30 and 31 never occur: their entries are invalid
                for (int i = 0; i < 144; ++i) lens[i] = 8;
                for (int i = 144; i < 256; ++i) lens[i] = 9;
                for (int i = 256; i < 280; ++i) lens[i] = 7;
                for (int i = 280; i < 288; ++i) lens[i] = 8;
                for (int i = 0; i < 32; ++i) lens[288 + i] = 5;
            } else {
                br.refill();
                nlit = (int)br.take(5) + 257;
                ndist = (int)br.take(5) + 1;
                const int nclen = (int)br.take(4) + 4;
                if (nlit > 286 || ndist > 30) return -1;
                static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
                uint8_t cl[19] = {0};
                for (int i = 0; i < nclen; ++i) {
                    if (br.cnt < 3) br.refill();
                    cl[order[i]] = (uint8_t)br.take(3);
                }
                uint32_t ct[1 << 7];
                if (!build_table(cl, 19, 7, ct, 1 << 7, [](int s) { return (uint32_t)s << 16; })) return -1;
                int i = 0;
                while (i < nlit + ndist) {
                    br.refill();
                    const uint32_t e = ct[br.peek(7)];
                    if (e & kInvalid) return -1;
                    br.drop(e & 0xff);
                    const int sym = (int)(e >> 16);
                    if (sym < 16) { lens[i++] = (uint8_t)sym; continue; }
                    int rep;
                    uint8_t val = 0;
                    if (sym == 16) {
                        if (i == 0) return -1;
                        val = lens[i - 1];
                        rep = 3 + (int)br.take(2);
                    } else if (sym == 17) rep = 3 + (int)br.take(3);
                    else rep = 11 + (int)br.take(7);
                    if (i + rep > nlit + ndist) return -1;
                    while (rep--) lens[i++] = val;
                }
                if (br.overrun * 8 > br.cnt) return -1;
                if (lens[256] == 0) return -1;                  // no end-of-block code
                // the distance lengths follow the literal / length ones: move them to a fixed place
                uint8_t dl[32];
                std::memcpy(dl, lens + nlit, ndist);
                std::memset(lens + nlit, 0, 288 - nlit);
                std::memcpy(lens + 288, dl, ndist);
            }
            if (!build_table(lens, type == 1 ? 288 : nlit, kLitBits, T.lit, kMaxLitEntries, lit_entry)) return -1;
            if (!build_table(lens + 288, ndist, kDistBits, T.dist, kMaxDistEntries, dist_entry)) return -1;
            // ---- symbols: the fast loop while 16 input bytes and a maximal match (+ the copy's overshoot) fit, then the careful one
            int st = 1;
            while (st == 1) {
                if (br.end - br.in >= 16 && out_end - out >= 3 * 258 + 16) st = decode_symbols<true>(br, out, dst, out_end, T.lit, T.dist);
                else st = decode_symbols<false>(br, out, dst, out_end, T.lit, T.dist);
            }
            if (st < 0) return -1;
        } else {
            return -1;
        }
        if (br.overrun * 8 > br.cnt) return -1;                  // bits past the end of the input were consumed
        if (final) break;
    }
    // whole real bytes the bit buffer still holds were not consumed
    const size_t unused = (size_t)((br.cnt - 8 * br.overrun) >> 3);
    if (consumed) *consumed = (size_t)(br.in - src) - unused;
    return out - dst;
}

inline uint32_t adler32(const uint8_t *p, size_t n)
{
    uint32_t a = 1, b = 0;
    while (n) {
        size_t k = n < 5552 ? n : 5552;
        n -= k;
        while (k >= 8) {
            a += p[0]; b += a; a += p[1]; b += a; a += p[2]; b += a; a += p[3]; b += a;
            a += p[4]; b += a; a += p[5]; b += a; a += p[6]; b += a; a += p[7]; b += a;
            p += 8;
            k -= 8;
        }
        while (k--) { a += *p++; b += a; }
        a %= 65521;
        b %= 65521;
    }
    return (b << 16) | a;
}

// zlib stream (2-byte header, DEFLATE, Adler-32 of the output).  Returns bytes written or -1.
inline int64_t zlib_uncompress(const uint8_t *src, size_t src_len, uint8_t *dst, size_t dst_cap, Tables &T)
{
    if (src_len < 6) return -1;
    const uint32_t cmf = src[0], flg = src[1];
    if ((cmf & 0x0f) != 8 || (cmf >> 4) > 7 || ((cmf << 8) | flg) % 31 != 0 || (flg & 0x20)) return -1;
    size_t used = 0;
    const int64_t n = inflate_raw(src + 2, src_len - 2, dst, dst_cap, &used, T);
    if (n < 0 || 2 + used + 4 > src_len) return -1;
    const uint8_t *t = src + 2 + used;
    const uint32_t want = ((uint32_t)t[0] << 24) | ((uint32_t)t[1] << 16) | ((uint32_t)t[2] << 8) | t[3];
    if (adler32(dst, (size_t)n) != want) return -1;
    return n;
}

}  // namespace rmg_inflate
