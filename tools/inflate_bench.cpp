// Single-thread inflate + CRC-32 of every block of a BGZF file: zlib against csrc/hostz.cpp (the numbers of DESIGN 4.0's
// host-decode paragraph).  Minimum of seven passes, and the ratio inside one run (the container's clock wanders).
//   g++ -O2 -I cnvetti_b200/csrc cnvetti_b200/csrc/hostz.cpp tools/inflate_bench.cpp -lz -o /tmp/inflate_bench
//   /tmp/inflate_bench file.bam
#include "hostz.h"
#include <zlib.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <chrono>
#include <algorithm>
using namespace cnvb;
struct Blk { size_t off, clen, isize; uint32_t crc; };
int main(int argc, char** argv) {
    FILE* f = fopen(argv[1], "rb"); fseek(f, 0, SEEK_END); size_t n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> buf(n + 64); fread(buf.data(), 1, n, f); fclose(f);
    std::vector<Blk> blocks; size_t p = 0, utotal = 0;
    while (p + 18 <= n) {
        const uint8_t* h = buf.data() + p; uint32_t xlen = h[10] | h[11] << 8; uint32_t bsize = (h[16] | h[17] << 8) + 1;
        Blk b; b.off = p + 12 + xlen; b.clen = bsize - 12 - xlen - 8; memcpy(&b.crc, h + bsize - 8, 4); uint32_t is; memcpy(&is, h + bsize - 4, 4); b.isize = is;
        blocks.push_back(b); utotal += is; p += bsize;
    }
    std::vector<uint8_t> out(utotal + 64);
    using clk = std::chrono::steady_clock;
    double bz=1e9,bf=1e9;
    for (int rep = 0; rep < 7; ++rep) {
        auto t0 = clk::now(); size_t o = 0; int bad = 0;
        z_stream zs; memset(&zs, 0, sizeof zs); inflateInit2(&zs, -15);
        for (auto& b : blocks) { if (!b.isize) continue; inflateReset(&zs); zs.next_in = buf.data() + b.off; zs.avail_in = b.clen; zs.next_out = out.data() + o; zs.avail_out = b.isize;
            if (inflate(&zs, Z_FINISH) != Z_STREAM_END) bad++; o += b.isize; }
        auto t1 = clk::now(); o = 0;
        for (auto& b : blocks) { if ((uint32_t)crc32(0, out.data() + o, b.isize) != b.crc) bad++; o += b.isize; }
        auto t2 = clk::now(); o = 0;
        for (auto& b : blocks) { if (!b.isize) continue; if (!inflate_raw_fast(buf.data() + b.off, b.clen, out.data() + o, b.isize)) bad++; o += b.isize; }
        auto t3 = clk::now(); o = 0;
        for (auto& b : blocks) { if (crc32_fast(out.data() + o, b.isize) != b.crc) bad++; o += b.isize; }
        auto t4 = clk::now();
        auto ms0 = [](auto a, auto b) { return std::chrono::duration<double>(b - a).count() * 1e3; }; bz = std::min(bz, ms0(t0,t1)); bf = std::min(bf, ms0(t2,t3)); if (rep==6) printf("MIN zlib %.1f ms (%.0f MB/s)  fast %.1f ms (%.0f MB/s)  ratio %.2f\n", bz, utotal/1e3/bz, bf, utotal/1e3/bf, bz/bf);
        auto ms = [](auto a, auto b) { return std::chrono::duration<double>(b - a).count() * 1e3; };
        printf("%zu blocks %.1f MB -> %.1f MB: zlib inflate %.1f ms (%.0f MB/s) crc %.1f ms | fast inflate %.1f ms (%.0f MB/s) crc %.1f ms  bad=%d\n", blocks.size(), n / 1e6, utotal / 1e6,
               ms(t0, t1), utotal / 1e3 / ms(t0, t1), ms(t1, t2), ms(t2, t3), utotal / 1e3 / ms(t2, t3), ms(t3, t4), bad);
    }
}
