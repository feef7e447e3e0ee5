// Checker of cnvetti_b200/csrc/hostz.cpp (fast DEFLATE decoder + carry-less-multiplication CRC-32) against zlib: built and
// run by tests/test_hostz.py with the address and undefined-behaviour sanitizers.  argv[1] = number of rounds.
//   * CRC-32 equals zlib's for lengths around every folding boundary and unaligned starts;
//   * every raw DEFLATE stream zlib writes (levels 0 / 1 / 6 / 9; default, fixed-code, Huffman-only and RLE strategies;
//     six kinds of data, sizes 0 .. 65536) is decoded to the source bytes, with canaries on both sides of the output;
//   * a wrong output size is refused; streams with a flipped bit or cut short never write outside the output and are
//     never accepted when cut short.
#include "hostz.h"
#include <zlib.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <random>
#include <chrono>
using namespace cnvb;
static std::vector<uint8_t> deflate_raw(const std::vector<uint8_t>& src, int level, int strategy = Z_DEFAULT_STRATEGY) {
    z_stream zs; memset(&zs, 0, sizeof zs);
    deflateInit2(&zs, level, Z_DEFLATED, -15, 8, strategy);
    std::vector<uint8_t> out(deflateBound(&zs, src.size()) + 64);
    zs.next_in = (Bytef*)src.data(); zs.avail_in = src.size(); zs.next_out = out.data(); zs.avail_out = out.size();
    int r = deflate(&zs, Z_FINISH); if (r != Z_STREAM_END) { printf("deflate failed\n"); exit(1); }
    out.resize(zs.total_out); deflateEnd(&zs); return out;
}
int main(int argc, char **argv) {
    const int rounds = argc > 1 ? atoi(argv[1]) : 3000;
    std::mt19937_64 rng(1);
    // crc
    for (size_t n : {0, 1, 15, 16, 63, 64, 65, 127, 128, 1000, 65536, 65280, 1000003}) {
        std::vector<uint8_t> d(n + 3); for (auto& x : d) x = rng();
        for (int off = 0; off < 3; ++off) {
            uint32_t a = crc32_fast(d.data() + off, n), b = crc32(0, d.data() + off, n);
            if (a != b) { printf("CRC MISMATCH n=%zu off=%d %08x %08x\n", n, off, a, b); return 1; }
        }
    }
    printf("crc ok\n");
    int cases = 0, fallback = 0;
    for (int it = 0; it < rounds; ++it) {
        size_t n = it < 20 ? it : (rng() % 65536) + 1;
        if (it % 50 == 0) n = 65280;
        std::vector<uint8_t> src(n);
        int kind = it % 6;
        for (size_t i = 0; i < n; ++i) {
            switch (kind) {
            case 0: src[i] = rng(); break;                         // incompressible
            case 1: src[i] = "ACGT"[rng() & 3]; break;             // 2 bits/byte
            case 2: src[i] = (uint8_t)(i / 7); break;              // runs
            case 3: src[i] = i % 100 < 90 ? (uint8_t)(i % 100) : rng(); break;  // periodic with noise
            case 4: src[i] = 0; break;
            default: src[i] = (rng() % 10 < 8 && i > 300) ? src[i - 1 - rng() % 300] : rng() % 40; break;
            }
        }
        for (int level : {0, 1, 6, 9}) for (int strat : {Z_DEFAULT_STRATEGY, Z_FIXED, Z_HUFFMAN_ONLY, Z_RLE}) {
            auto comp = deflate_raw(src, level, strat);
            std::vector<uint8_t> in(comp.size() + kInflateSlack, 0xA5);
            memcpy(in.data(), comp.data(), comp.size());
            std::vector<uint8_t> out(n + 64, 0xEE);
            bool ok = inflate_raw_fast(in.data(), comp.size(), out.data() + 32, n);
            ++cases;
            if (!ok) { ++fallback; printf("REFUSED valid stream it=%d n=%zu level=%d strat=%d\n", it, n, level, strat); return 1; }
            if (memcmp(out.data() + 32, src.data(), n)) { printf("DATA MISMATCH it=%d n=%zu level=%d\n", it, n, level); return 1; }
            for (int i = 0; i < 32; ++i) if (out[i] != 0xEE || out[32 + n + i] != 0xEE) { printf("CANARY it=%d\n", it); return 1; }
            // wrong out_len must be refused
            if (n > 1) {
                std::vector<uint8_t> o2(n + 64, 0xEE);
                if (inflate_raw_fast(in.data(), comp.size(), o2.data() + 32, n - 1)) { printf("ACCEPTED short out it=%d\n", it); return 1; }
                for (int i = 0; i < 32; ++i) if (o2[i] != 0xEE || o2[32 + n - 1 + i] != 0xEE) { printf("CANARY2 it=%d\n", it); return 1; }
                std::vector<uint8_t> o3(n + 1 + 64, 0xEE);
                if (inflate_raw_fast(in.data(), comp.size(), o3.data() + 32, n + 1)) { printf("ACCEPTED long out it=%d\n", it); return 1; }
            }
            // corrupt streams: no crash, no out-of-bounds write; accepted ones must carry the right CRC or differ (checked by CRC upstream)
            if (it % 10 == 0 && comp.size() > 4) {
                for (int f = 0; f < 20; ++f) {
                    std::vector<uint8_t> bad = in;
                    bad[rng() % comp.size()] ^= 1u << (rng() % 8);
                    std::vector<uint8_t> o4(n + 64, 0xEE);
                    inflate_raw_fast(bad.data(), comp.size(), o4.data() + 32, n);
                    for (int i = 0; i < 32; ++i) if (o4[i] != 0xEE || o4[32 + n + i] != 0xEE) { printf("CANARY3 it=%d\n", it); return 1; }
                    size_t cut = rng() % comp.size();
                    std::vector<uint8_t> tr(cut + kInflateSlack, 0x5A); memcpy(tr.data(), comp.data(), cut);
                    if (inflate_raw_fast(tr.data(), cut, o4.data() + 32, n) && cut < comp.size() - 1) { /* may happen only if stream ends earlier: impossible */ printf("ACCEPTED truncated it=%d cut=%zu of %zu\n", it, cut, comp.size()); return 1; }
                }
            }
        }
    }
    printf("inflate ok: %d cases\n", cases);
    return 0;
}
