// bam.cu -- host BGZF/BAM decoder: file -> per-contig structure-of-arrays read batches.
//
// Replaces what the reference gets from rust-htslib for the coverage path: `bam::IndexedReader`
// + `fetch(tid, 0, len)` + `read(&mut record)` per region (lib-coverage/src/lib.rs:533-547,
// agg.rs:224-228) and the per-record derivations of the aggregators (agg.rs:124-133, 174-198,
// 312-320).  `cmd coverage` visits every contig of the file in turn, so the decoder makes ONE
// pass over the file instead of one index seek per region: BGZF blocks are inflated on host
// threads (hostz.cpp's DEFLATE decoder and carry-less-multiplication CRC-32, zlib behind them; one block per task),
// record boundaries are walked once, and the
// threads derive the SoA columns of disjoint record ranges.  The columns of a contig are what
// `fetch` on that contig yields, in file (= coordinate) order: every record whose refID is the
// contig, including unmapped reads placed at their mate's position; records without a
// reference (refID < 0) are counted and dropped.
//
// Pure host code.  The columns are page-locked on request (cudaHostRegister, once the decode is complete) so
// that they can be handed to cnvb_cov_put_records / cnvb_coverage_run and copied by DMA.
#include <sys/mman.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"
#include "hostz.h"

using namespace cnvb;

namespace {

// Fresh anonymous memory costs a page fault per 4 KiB the first time it is touched -- for a cold decode (one reader
// per file: the real case) that was more than the decode itself (2 M reads, 8 threads: 154 ms cold against 66 ms
// with buffers reused).  Transparent huge pages are opt-in on these boxes (madvise): ask for them on every large
// buffer, one fault per 2 MiB then.
inline void advise_huge(void *p, uint64_t bytes) {
#ifdef MADV_HUGEPAGE
    static const bool off = getenv("CNVB_BAM_NO_THP") != nullptr;  // (A/B measurements)
    if (off || bytes < (4ull << 20)) return;
    const uintptr_t a = (reinterpret_cast<uintptr_t>(p) + 4095u) & ~uintptr_t(4095), e = (reinterpret_cast<uintptr_t>(p) + bytes) & ~uintptr_t(4095);
    if (e > a) madvise(reinterpret_cast<void *>(a), e - a, MADV_HUGEPAGE);
#else
    (void)p, (void)bytes;
#endif
}

template <typename T>
struct HostColumn {  // growing array; page-locked in place once the decode is complete
    T *p = nullptr;
    uint64_t n = 0, cap = 0;
    bool pinned = false;
    bool reserve(uint64_t want, bool /*pin*/) {
        if (want <= cap) return true;
        unpin();
        const uint64_t ncap = std::max<uint64_t>(want, cap ? cap * 2 : 1 << 16);
        T *q = static_cast<T *>(realloc(p, ncap * sizeof(T)));  // (grows in place or by remapping: no pinned copies)
        if (!q) return false;
        advise_huge(q, ncap * sizeof(T));
        p = q;
        cap = ncap;
        return true;
    }
    // cudaHostRegister pins the pages where they are (about 0.1 ms per MB); allocating page-locked memory up
    // front and copying on every growth step cost 80 ms per 2 M reads
    void pin() {
        if (pinned || !p || !n) return;
        // (the rows in use, not the capacity the doubling left behind: page-locking costs by the page)
        const uint64_t bytes = std::min<uint64_t>(cap * sizeof(T), (n * sizeof(T) + 4095u) & ~uint64_t(4095));
        if (cudaHostRegister(p, bytes, cudaHostRegisterDefault) == cudaSuccess) pinned = true; else cudaGetLastError();
    }
    void unpin() {
        if (pinned) cudaHostUnregister(p);
        pinned = false;
    }
    void release() {
        unpin();
        free(p);
        p = nullptr;
        cap = 0;
    }
};

struct RefReads {
    HostColumn<int32_t> pos, end, mid, frag_end;
    HostColumn<uint8_t> mapq;
    HostColumn<uint16_t> flags;
    uint64_t n = 0;
    bool reserve(uint64_t want, bool pin) {
        return pos.reserve(want, pin) && end.reserve(want, pin) && mid.reserve(want, pin) &&
               frag_end.reserve(want, pin) && mapq.reserve(want, pin) && flags.reserve(want, pin);
    }
    void release() {
        pos.release(); end.release(); mid.release(); frag_end.release(); mapq.release(); flags.release();
    }
    void pin() {
        pos.pin(); end.pin(); mid.pin(); frag_end.pin(); mapq.pin(); flags.pin();
    }
};

struct RawBuf {  // byte buffer without value-initialisation (std::vector::resize would memset every chunk)
    uint8_t *p = nullptr;
    uint64_t n = 0, cap = 0;
    bool resize(uint64_t want, bool keep) {
        if (want > cap) {
            uint64_t ncap = std::max<uint64_t>(want, cap + cap / 2);
            const bool huge = !keep && ncap >= (4ull << 20);  // (2 MiB aligned, so that every part of it can be a huge page)
            if (huge) ncap = (ncap + (2ull << 20) - 1) & ~((2ull << 20) - 1);
            uint8_t *q = static_cast<uint8_t *>(keep ? realloc(p, ncap) : huge ? aligned_alloc(2ull << 20, ncap) : malloc(ncap));
            if (!q) return false;
            advise_huge(q, ncap);
            if (!keep) free(p);
            p = q;
            cap = ncap;
        }
        n = want;
        return true;
    }
    uint8_t *data() { return p; }
    uint64_t size() const { return n; }
    ~RawBuf() { free(p); }
};

struct Block {  // one BGZF block inside the file image
    uint64_t off;     // start of the deflate stream
    uint32_t clen;    // compressed payload length
    uint32_t isize;   // uncompressed length (trailer)
    uint32_t crc;
    uint64_t uoff;    // offset in the chunk's uncompressed buffer
};

inline uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline uint16_t rd16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return v; }

// the per-record derivation shared with cnvb_pack_reads (ctx.cu)
inline void derive(int32_t pos, int32_t isize, uint16_t f, const uint8_t *cigar, uint32_t n_cigar, float min_unclipped,
                   int32_t &o_end, int32_t &o_mid, int32_t &o_frag_end, uint16_t &o_flags) {
    uint32_t ref_len = 0, clipped = 0, unclipped = 0;
    for (uint32_t c = 0; c < n_cigar; ++c) {
        const uint32_t w = rd32(cigar + 4 * c), op = w & 0xFu, len = w >> 4;
        switch (op) {
        case 0: case 7: case 8: ref_len += len; unclipped += len; break;  // M = X
        case 1: unclipped += len; break;                                  // I
        case 2: case 3: ref_len += len; break;                            // D N
        case 4: case 5: clipped += len; break;                            // S H
        default: break;                                                   // P
        }
    }
    const int32_t end_pos = pos + (int32_t)ref_len;  // CigarStringView::end_pos()
    uint32_t rlen = 1;                               // htslib bam_endpos()
    if (!(f & 0x4) && n_cigar > 0) rlen = ref_len ? ref_len : 1;
    o_end = pos + (int32_t)rlen;
    const bool paired = f & 0x1;
    o_mid = paired ? pos + isize / 2 : pos + end_pos;          // agg.rs:124-133
    o_frag_end = paired ? pos + isize : end_pos;               // agg.rs:312-320
    uint16_t x = f & 0x0FFF;
    const float ratio = (float)unclipped / (float)(clipped + unclipped);  // agg.rs:174-193 (f32; NaN compares false)
    if (ratio < min_unclipped) x |= CNVB_FLAG_CLIPFAIL;
    if (isize <= 0) x |= CNVB_FLAG_ISIZE_NONPOS;               // agg.rs:196-198
    o_flags = x;
}

void run_parallel(int n_threads, uint64_t n_items, const std::function<void(uint64_t, uint64_t, int)> &fn) {
    if (n_threads <= 1 || n_items < 2) {
        fn(0, n_items, 0);
        return;
    }
    std::vector<std::thread> th;
    const uint64_t per = (n_items + n_threads - 1) / n_threads;
    for (int t = 0; t < n_threads; ++t) {
        const uint64_t a = std::min<uint64_t>(n_items, per * t), b = std::min<uint64_t>(n_items, a + per);
        if (a < b) th.emplace_back(fn, a, b, t);
    }
    for (auto &x : th) x.join();
}

}  // namespace

struct BaiRef {  // one reference of a .bai index
    std::unordered_map<uint32_t, std::vector<std::pair<uint64_t, uint64_t>>> bins;
    std::vector<uint64_t> linear;
};

struct FetchRange {  // an indexed fetch: start reading at a virtual offset, keep what overlaps [beg, end) on tid
    uint64_t coff;   // file offset of the first BGZF block
    uint32_t uoff;   // offset of the first record inside it
    int32_t tid, beg, end;
};

struct cnvb_bam {
    std::vector<BaiRef> bai;
    bool bai_loaded = false;
    std::string path, error, text;
    std::vector<std::string> ref_names;
    std::vector<uint32_t> ref_lens;
    std::vector<RefReads> refs;
    int n_threads = 1;
    uint64_t n_records = 0, n_unplaced = 0, bytes_compressed = 0, bytes_uncompressed = 0;
    double t_inflate = 0.0, t_parse = 0.0, t_total = 0.0;
    bool decoded = false;
    RawBuf cbuf, ubuf, ubuf2;  // chunk buffers, kept across chunks and calls (no page faults after the first chunks)
    ~cnvb_bam() {
        for (auto &r : refs) r.release();
    }
};

namespace {

int fail_to(std::string &dst, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    dst = buf;
    return code;
}

struct Run {  // consecutive records of one refID inside a chunk
    uint64_t first;  // index of its first record in the chunk
    int32_t tid;
    uint64_t dst;    // row in the contig's columns where the run starts
};

// One stretch of the record walk: the records that start in [q, limit) and lie completely inside the n bytes at u.
// Appends their offsets and the runs of equal refID (first = index into offs); returns where the walk stopped (the
// first record at / past the limit, or the incomplete one at the end of the data).  bad: a block_size below 32.
struct WalkOut {
    std::vector<uint64_t> offs;
    std::vector<Run> runs;
    uint64_t end = 0, bad_size = 0;
    bool bad = false;
};
void walk_records(const uint8_t *u, uint64_t n, uint64_t q, uint64_t limit, WalkOut &o) {
    int32_t run_tid = 0;
    while (q < limit && q + 4 <= n) {
        const uint64_t bs = rd32(u + q);
        if (bs < 32) {
            o.bad = true;
            o.bad_size = bs;
            break;
        }
        if (q + 4 + bs > n) break;
        // the walk is a chain of dependent loads, one cache line per record (~55 ns each when cold): records
        // of a file have similar sizes, so the line some records ahead can be requested now
        __builtin_prefetch(u + q + 12 * (4 + bs));
        __builtin_prefetch(u + q + 12 * (4 + bs) + 64);
        const int32_t tid = (int32_t)rd32(u + q + 4);
        if (o.runs.empty() || tid != run_tid) {
            o.runs.push_back(Run{o.offs.size(), tid, 0});
            run_tid = tid;
        }
        o.offs.push_back(q);
        q += 4 + bs;
    }
    o.end = q;
}

// Does a chain of records start at q?  The fixed fields of up to four consecutive records must be consistent with
// one another and with the header (a guess: walk_parallel keeps a stretch only if the walk before it arrives at
// exactly this offset).
bool plausible_record_chain(const uint8_t *u, uint64_t n, uint64_t q, int64_t n_ref) {
    for (int k = 0; k < 4; ++k) {
        if (q + 36 > n) return k > 0;
        const uint64_t bs = rd32(u + q);
        const int64_t tid = (int32_t)rd32(u + q + 4), pos = (int32_t)rd32(u + q + 8);
        const uint64_t l_name = u[q + 12], n_cigar = rd16(u + q + 16), l_seq = rd32(u + q + 20);
        const int64_t mtid = (int32_t)rd32(u + q + 24), mpos = (int32_t)rd32(u + q + 28);
        if (bs < 32 || bs > (1u << 28) || tid < -1 || tid >= n_ref || pos < -1 || mtid < -1 || mtid >= n_ref || mpos < -1 ||
            l_name == 0 || l_seq > (1u << 28))
            return false;
        if (32 + l_name + 4 * n_cigar + (l_seq + 1) / 2 + l_seq > bs) return false;
        if (q + 36 + l_name <= n && u[q + 36 + l_name - 1] != 0) return false;  // read_name is NUL-terminated
        q += 4 + bs;
    }
    return true;
}

// The record walk of a chunk on all threads.  The chain of block_size fields is serial, so every thread but the
// first GUESSES where a record starts in its share of the bytes (the first offset from which four records in a row
// look like records) and walks from there; the stretches are then joined front to back, and a stretch is kept only
// if the walk before it ends exactly where the stretch began -- otherwise that share is walked again from the true
// offset.  The result is the serial walk's, whatever the guesses were.
void walk_parallel(const uint8_t *u, uint64_t n, uint64_t q, int64_t n_ref, int n_threads, WalkOut &out) {
    const uint64_t span = n - q;
    uint64_t share = 1ull << 20;  // at least 1 MiB each (tests: CNVB_BAM_WALK_SHARE makes the shares small)
    if (const char *e = getenv("CNVB_BAM_WALK_SHARE")) share = std::max<uint64_t>(64, (uint64_t)atoll(e));
    const int parts = (int)std::min<uint64_t>((uint64_t)std::max(1, n_threads), span / share);
    if (parts <= 1) {
        walk_records(u, n, q, n, out);
        return;
    }
    std::vector<WalkOut> seg(parts);
    std::vector<uint64_t> lo(parts + 1), guess(parts, UINT64_MAX);
    for (int i = 0; i <= parts; ++i) lo[i] = i == parts ? n : q + span * (uint64_t)i / parts;
    run_parallel(parts, (uint64_t)parts, [&](uint64_t a, uint64_t b, int) {
        for (uint64_t i = a; i < b; ++i) {
            uint64_t start = lo[i];
            if (i > 0) {
                const uint64_t stop = std::min<uint64_t>(lo[i + 1], lo[i] + (4ull << 20));  // (records longer than this: walked serially)
                while (start < stop && !plausible_record_chain(u, n, start, n_ref)) ++start;
                if (start >= stop) continue;  // (no record starts in this share, or none that looks like one)
            }
            guess[i] = start;
            walk_records(u, n, start, lo[i + 1], seg[i]);
        }
    });
    uint64_t cur = q;
    for (int i = 0; i < parts && !out.bad; ++i) {
        if (cur >= lo[i + 1]) continue;  // a record spans the whole share
        WalkOut redo;
        WalkOut *w = &seg[i];
        if (guess[i] != cur) {
            walk_records(u, n, cur, lo[i + 1], redo);
            w = &redo;
        }
        const uint64_t base = out.offs.size();
        for (const Run &r : w->runs)
            if (out.runs.empty() || out.runs.back().tid != r.tid) out.runs.push_back(Run{base + r.first, r.tid, 0});
        out.offs.insert(out.offs.end(), w->offs.begin(), w->offs.end());
        out.bad = w->bad;
        out.bad_size = w->bad_size;
        const bool stopped_short = w->end < lo[i + 1];  // incomplete record at the end of the data (or a bad one)
        cur = w->end;
        if (stopped_short) break;
    }
    out.end = cur;
}

// The whole decode: one pass over the file in chunks of compressed bytes.  Two stages per chunk --
// (A) read, cut into BGZF blocks, inflate on all threads; (B) header / record walk / SoA derivation -- and
// stage B of chunk k runs on a helper thread while stage A of chunk k+1 proceeds (two uncompressed
// buffers).  Stage B is sequential from chunk to chunk (it hands the partial record at the end of a chunk
// to the next one); its serial part, the record walk, is what the overlap hides.
int decode(cnvb_bam *b, float min_unclipped, bool pinned, bool header_only, const FetchRange *range = nullptr) {
    using clk = std::chrono::steady_clock;
    const auto t_begin = clk::now();
    FILE *f = fopen(b->path.c_str(), "rb");
    if (!f) return fail_to(b->error, CNVB_ERR_INVALID, "Could not open BAM file %s", b->path.c_str());
    // compressed bytes per chunk (an indexed fetch starts small and doubles: short intervals stay short reads)
    uint64_t kChunkBytes = header_only ? (1ull << 20) : range ? (256ull << 10) : (8ull << 20);
    const bool grow_chunks = range != nullptr && !getenv("CNVB_BAM_CHUNK");
    if (const char *e = getenv("CNVB_BAM_CHUNK")) {  // (tests: small chunks exercise the chunk seams)
        const long v = atol(e);
        if (v >= 64) kChunkBytes = (uint64_t)v;
    }
    const bool zlib_only = getenv("CNVB_BAM_ZLIB") != nullptr;  // (A/B measurements, tests of the fallback)
    constexpr uint64_t kHead = 1ull << 16;  // room in front of a chunk's bytes for the carried-over partial record
    RawBuf &cbuf = b->cbuf;
    RawBuf *ubufs[2] = {&b->ubuf, &b->ubuf2};
    std::vector<Block> blocks;
    std::vector<uint8_t> left;  // compressed bytes of an incomplete block at the end of the last read
    uint64_t file_pos = 0;
    bool eof = false;
    int rc = CNVB_OK;
    if (range) {
        if (fseeko(f, (off_t)range->coff, SEEK_SET) != 0) {
            fclose(f);
            return fail_to(b->error, CNVB_ERR_INVALID, "Could not seek to region in %s", b->path.c_str());
        }
        file_pos = range->coff;
    }

    // ---- stage B state (touched by the helper thread only while it runs)
    std::vector<uint8_t> carry, joined;
    std::vector<uint64_t> rec_off;
    std::vector<Run> runs;
    bool header_done = range != nullptr, stop_after_header = false;  // (an indexed fetch starts at a record)
    bool first_parse = true;
    std::string werr;
    int wrc = CNVB_OK;
    double t_parse = 0.0;
    auto parse = [&](RawBuf *ub, uint64_t n_data, bool last) {
        const auto t1 = clk::now();
        const uint8_t *u;
        uint64_t un;
        if (carry.size() <= kHead) {
            u = ub->data() + kHead - carry.size();
            if (!carry.empty()) memcpy(const_cast<uint8_t *>(u), carry.data(), carry.size());
            un = carry.size() + n_data;
        } else {  // (a header or record longer than the head-room: join in a buffer of its own)
            joined.assign(carry.begin(), carry.end());
            joined.insert(joined.end(), ub->data() + kHead, ub->data() + kHead + n_data);
            u = joined.data();
            un = joined.size();
        }
        uint64_t q = 0;
        if (range && first_parse) q = range->uoff;
        first_parse = false;
        if (!header_done) {
            bool complete = false;
            do {
                if (un < 12) break;
                if (memcmp(u, "BAM\1", 4) != 0) {
                    wrc = fail_to(werr, CNVB_ERR_INVALID, "%s: not a BAM file (bad magic)", b->path.c_str());
                    return;
                }
                const uint64_t l_text = rd32(u + 4);
                if (un < 12 + l_text) break;
                const uint32_t n_ref = rd32(u + 8 + l_text);
                uint64_t w = 12 + l_text;
                std::vector<std::string> names;
                std::vector<uint32_t> lens;
                bool ok = true;
                for (uint32_t r = 0; r < n_ref; ++r) {
                    if (un < w + 4) { ok = false; break; }
                    const uint64_t l_name = rd32(u + w);
                    if (un < w + 4 + l_name + 4) { ok = false; break; }
                    names.emplace_back(reinterpret_cast<const char *>(u + w + 4), l_name ? l_name - 1 : 0);
                    lens.push_back(rd32(u + w + 4 + l_name));
                    w += 8 + l_name;
                }
                if (!ok) break;
                b->text.assign(reinterpret_cast<const char *>(u + 8), l_text);
                while (!b->text.empty() && b->text.back() == '\0') b->text.pop_back();
                b->ref_names = names;
                b->ref_lens = lens;
                b->refs.resize(n_ref);
                q = w;
                complete = true;
            } while (false);
            if (!complete) {  // header spans more than this chunk: read on
                if (last) {
                    wrc = fail_to(werr, CNVB_ERR_INVALID, "%s: truncated BAM header", b->path.c_str());
                    return;
                }
                carry.assign(u, u + un);
                return;
            }
            header_done = true;
            if (header_only) {
                stop_after_header = true;
                return;
            }
        }
        // record boundaries, and the runs of equal refID among them (the refID sits next to block_size)
        rec_off.clear();
        runs.clear();
        if (!range) {
            WalkOut w;
            if (getenv("CNVB_BAM_SERIAL_WALK")) walk_records(u, un, q, un, w);  // (A/B measurements, tests)
            else walk_parallel(u, un, q, (int64_t)b->refs.size(), b->n_threads, w);
            if (w.bad) {
                wrc = fail_to(werr, CNVB_ERR_INVALID, "%s: corrupt BAM record (block_size %llu)", b->path.c_str(),
                              (unsigned long long)w.bad_size);
                return;
            }
            rec_off.swap(w.offs);
            runs.swap(w.runs);
            q = w.end;
        }
        int32_t run_tid = 0;
        while (range && q + 4 <= un) {
            const uint64_t bs = rd32(u + q);
            if (bs < 32) {
                wrc = fail_to(werr, CNVB_ERR_INVALID, "%s: corrupt BAM record (block_size %llu)", b->path.c_str(),
                              (unsigned long long)bs);
                return;
            }
            if (q + 4 + bs > un) break;
            // the walk is a chain of dependent loads, one cache line per record (~55 ns each when cold): records
            // of a file have similar sizes, so the line some records ahead can be requested now
            __builtin_prefetch(u + q + 12 * (4 + bs));
            __builtin_prefetch(u + q + 12 * (4 + bs) + 64);
            const int32_t tid = (int32_t)rd32(u + q + 4);
            if (range) {
                // htslib's iterator: the records from the start offset on until one lies on another reference or
                // starts at / after the end; of those, the ones whose [pos, bam_endpos) overlaps the interval
                const int32_t pos = (int32_t)rd32(u + q + 8);
                if (tid != range->tid || pos >= range->end) {
                    stop_after_header = true;  // (ends the read loop)
                    break;
                }
                const uint32_t l_name = u[q + 12], n_cigar = rd16(u + q + 16);
                const uint16_t flag = rd16(u + q + 18);
                uint32_t ref_len = 0;
                if (36ull + l_name + 4ull * n_cigar <= 4ull + bs)
                    for (uint32_t c = 0; c < n_cigar; ++c) {
                        const uint32_t w = rd32(u + q + 36 + l_name + 4 * c), op = w & 0xFu;
                        if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) ref_len += w >> 4;
                    }
                const int32_t endpos = pos + (int32_t)((!(flag & 0x4) && n_cigar > 0 && ref_len) ? ref_len : 1);
                if (endpos <= range->beg) {
                    q += 4 + bs;
                    continue;
                }
            }
            if (runs.empty() || tid != run_tid) {
                runs.push_back(Run{rec_off.size(), tid, 0});
                run_tid = tid;
            }
            rec_off.push_back(q);
            q += 4 + bs;
        }
        const uint64_t nr = rec_off.size();
        // room in the contigs' columns first (file order is kept: a run is appended where its contig ends),
        // then the threads derive disjoint record ranges straight into them
        for (size_t k = 0; k < runs.size(); ++k) {
            Run &run = runs[k];
            const uint64_t m = (k + 1 < runs.size() ? runs[k + 1].first : nr) - run.first;
            if (run.tid < 0) {
                b->n_unplaced += m;
            } else if ((uint64_t)run.tid >= b->refs.size()) {
                wrc = fail_to(werr, CNVB_ERR_INVALID, "%s: record with refID %d but the header has %zu references",
                              b->path.c_str(), run.tid, b->refs.size());
                return;
            } else {
                RefReads &rr = b->refs[run.tid];
                if (!rr.reserve(rr.n + m, pinned)) {
                    wrc = fail_to(werr, CNVB_ERR_NOMEM, "out of host memory while decoding %s", b->path.c_str());
                    return;
                }
                run.dst = rr.n;
                rr.n += m;
                rr.pos.n = rr.end.n = rr.mid.n = rr.frag_end.n = rr.mapq.n = rr.flags.n = rr.n;
            }
        }
        std::atomic<int> badrec{0};
        run_parallel(b->n_threads, nr, [&](uint64_t lo, uint64_t hi, int) {
            // the run holding record lo: last run with first <= lo
            size_t k = (size_t)(std::upper_bound(runs.begin(), runs.end(), lo,
                                                 [](uint64_t v, const Run &r) { return v < r.first; }) - runs.begin()) - 1;
            for (uint64_t i = lo; i < hi; ++i) {
                while (k + 1 < runs.size() && runs[k + 1].first <= i) ++k;
                const Run &run = runs[k];
                if (run.tid < 0) continue;
                const uint8_t *r = u + rec_off[i];
                const uint32_t bs = rd32(r);
                const int32_t pos = (int32_t)rd32(r + 8);
                const uint32_t l_name = r[12], mapq = r[13], n_cigar = rd16(r + 16);
                const uint16_t flag = rd16(r + 18);
                const int32_t tlen = (int32_t)rd32(r + 32);
                if (36ull + l_name + 4ull * n_cigar > 4ull + bs) {
                    badrec = 1;
                    continue;
                }
                RefReads &rr = b->refs[run.tid];
                const uint64_t at = run.dst + (i - run.first);
                rr.pos.p[at] = pos;
                rr.mapq.p[at] = (uint8_t)mapq;
                derive(pos, tlen, flag, r + 36 + l_name, n_cigar, min_unclipped, rr.end.p[at], rr.mid.p[at],
                       rr.frag_end.p[at], rr.flags.p[at]);
            }
        });
        if (badrec) {
            wrc = fail_to(werr, CNVB_ERR_INVALID, "%s: corrupt BAM record (CIGAR past the end of the record)", b->path.c_str());
            return;
        }
        b->n_records += nr;
        if (stop_after_header) carry.clear();
        else carry.assign(u + q, u + un);  // (may alias `joined`: assign from a range of another vector is fine)
        t_parse += std::chrono::duration<double>(clk::now() - t1).count();
    };

    std::thread worker;
    auto join_worker = [&]() {
        if (worker.joinable()) worker.join();
    };
    for (uint64_t k = 0; !eof && rc == CNVB_OK; ++k) {
        // ---- stage A: read compressed bytes, cut into whole blocks
        const uint64_t have = left.size();
        if (!cbuf.resize(have + kChunkBytes + kInflateSlack, false)) {  // (the decoder's loads may run past a block's end)
            rc = fail_to(b->error, CNVB_ERR_NOMEM, "out of host memory while decoding %s", b->path.c_str());
            break;
        }
        if (have) memcpy(cbuf.data(), left.data(), have);
        const size_t got = fread(cbuf.data() + have, 1, kChunkBytes, f);
        if (grow_chunks && kChunkBytes < (8ull << 20)) kChunkBytes *= 2;
        cbuf.n = have + got;
        file_pos += got;
        if (got == 0) eof = true;
        blocks.clear();
        uint64_t p = 0, utotal = 0;
        while (p + 18 <= cbuf.size()) {
            const uint8_t *h = cbuf.data() + p;
            if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) {
                rc = fail_to(b->error, CNVB_ERR_INVALID, "%s: not a BGZF file (bad block header at compressed offset %llu)",
                             b->path.c_str(), (unsigned long long)(file_pos - cbuf.size() + p));
                break;
            }
            const uint32_t xlen = rd16(h + 10);
            if (p + 12 + xlen > cbuf.size()) break;
            uint32_t bsize = 0;
            for (uint32_t x = 0; x + 4 <= xlen;) {  // extra subfields: find BC
                const uint8_t *e = h + 12 + x;
                const uint32_t slen = rd16(e + 2);
                if (e[0] == 'B' && e[1] == 'C' && slen == 2) bsize = (uint32_t)rd16(e + 4) + 1;
                x += 4 + slen;
            }
            if (bsize < 12 + xlen + 8) {
                rc = fail_to(b->error, CNVB_ERR_INVALID, "%s: BGZF block without a valid BC subfield", b->path.c_str());
                break;
            }
            if (p + bsize > cbuf.size()) break;  // incomplete block: next read
            Block bl;
            bl.off = p + 12 + xlen;
            bl.clen = bsize - (12 + xlen) - 8;
            bl.crc = rd32(h + bsize - 8);
            bl.isize = rd32(h + bsize - 4);
            bl.uoff = utotal;
            utotal += bl.isize;
            blocks.push_back(bl);
            p += bsize;
        }
        if (rc != CNVB_OK) break;
        left.assign(cbuf.data() + p, cbuf.data() + cbuf.size());
        if (eof && !left.empty()) {
            rc = fail_to(b->error, CNVB_ERR_INVALID, "%s: truncated BGZF block at the end of the file", b->path.c_str());
            break;
        }
        b->bytes_compressed += p;
        b->bytes_uncompressed += utotal;
        // ---- inflate the blocks of the chunk in parallel (the helper may still be parsing the previous chunk)
        const auto t0 = clk::now();
        RawBuf *ub = ubufs[k & 1];
        if (!ub->resize(kHead + utotal, false)) {
            rc = fail_to(b->error, CNVB_ERR_NOMEM, "out of host memory while decoding %s", b->path.c_str());
            break;
        }
        std::atomic<int> bad{0};
        // (blocks are handed out four at a time from a shared cursor: their decode times differ, and the helper
        // thread's parse of the previous chunk takes cores away unevenly)
        std::atomic<uint64_t> cursor{0};
        const uint64_t n_blocks = blocks.size();
        run_parallel(b->n_threads, (uint64_t)std::max(1, b->n_threads), [&](uint64_t, uint64_t, int) {
            // hostz.cpp's decoder first (a BGZF block is one buffer of known size); zlib for whatever it refuses
            // or gets wrong by the block's CRC-32 -- zlib's verdict is the one that stands
            z_stream zs;
            bool zs_init = false;
            for (uint64_t i0 = cursor.fetch_add(4); i0 < n_blocks; i0 = cursor.fetch_add(4))
            for (uint64_t i = i0; i < std::min(n_blocks, i0 + 4); ++i) {
                const Block &bl = blocks[i];
                if (bl.isize == 0) continue;
                uint8_t *dst = ub->data() + kHead + bl.uoff;
                if (!zlib_only && inflate_raw_fast(cbuf.data() + bl.off, bl.clen, dst, bl.isize) &&
                    crc32_fast(dst, bl.isize) == bl.crc)
                    continue;
                if (!zs_init) {
                    memset(&zs, 0, sizeof zs);
                    if (inflateInit2(&zs, -15) != Z_OK) {
                        bad = 1;
                        return;
                    }
                    zs_init = true;
                } else {
                    inflateReset(&zs);
                }
                zs.next_in = cbuf.data() + bl.off;
                zs.avail_in = bl.clen;
                zs.next_out = dst;
                zs.avail_out = bl.isize;
                const int r = inflate(&zs, Z_FINISH);
                if (r != Z_STREAM_END || zs.avail_out != 0 || crc32_fast(dst, bl.isize) != bl.crc) bad = 1;
            }
            if (zs_init) inflateEnd(&zs);
        });
        b->t_inflate += std::chrono::duration<double>(clk::now() - t0).count();
        if (bad) {
            rc = fail_to(b->error, CNVB_ERR_INVALID, "%s: corrupt BGZF block (inflate or CRC failure)", b->path.c_str());
            break;
        }
        // ---- stage B of this chunk, after stage B of the previous one
        join_worker();
        if (wrc != CNVB_OK || stop_after_header) break;
        const bool last = eof;
        worker = std::thread(parse, ub, utotal, last);
        if (header_only) {  // opening reads no further than the header needs
            join_worker();
            if (wrc != CNVB_OK || stop_after_header) break;
        }
    }
    join_worker();
    fclose(f);
    if (rc == CNVB_OK && wrc != CNVB_OK) {
        rc = wrc;
        b->error = werr;
    }
    b->t_parse += t_parse;
    if (rc == CNVB_OK && pinned && !header_only)
        for (auto &r : b->refs) r.pin();
    if (rc == CNVB_OK && !header_done) rc = fail_to(b->error, CNVB_ERR_INVALID, "%s: empty or truncated BAM file", b->path.c_str());
    if (rc == CNVB_OK && !header_only && !carry.empty())
        rc = fail_to(b->error, CNVB_ERR_INVALID, "%s: truncated BAM record at the end of the file", b->path.c_str());
    b->t_total += std::chrono::duration<double>(clk::now() - t_begin).count();
    return rc;
}

}  // namespace

extern "C" int cnvb_bam_open(const char *path, int n_threads, cnvb_bam **out) {
    if (!path || !out) return CNVB_ERR_INVALID;
    cnvb_bam *b = new (std::nothrow) cnvb_bam();
    if (!b) return CNVB_ERR_NOMEM;
    b->path = path;
    const unsigned hw = std::thread::hardware_concurrency();
    b->n_threads = n_threads > 0 ? n_threads : (hw ? (int)hw : 1);
    *out = b;
    return decode(b, 0.0f, false, true);  // header only; the handle is returned even on failure (error text)
}

extern "C" const char *cnvb_bam_error(const cnvb_bam *b) { return b ? b->error.c_str() : "null BAM handle"; }
extern "C" uint32_t cnvb_bam_n_refs(const cnvb_bam *b) { return b ? (uint32_t)b->ref_names.size() : 0; }
extern "C" const char *cnvb_bam_ref_name(const cnvb_bam *b, uint32_t tid) {
    return b && tid < b->ref_names.size() ? b->ref_names[tid].c_str() : nullptr;
}
extern "C" uint32_t cnvb_bam_ref_len(const cnvb_bam *b, uint32_t tid) {
    return b && tid < b->ref_lens.size() ? b->ref_lens[tid] : 0;
}
extern "C" const char *cnvb_bam_header_text(const cnvb_bam *b, uint64_t *len) {
    if (!b) return nullptr;
    if (len) *len = b->text.size();
    return b->text.c_str();
}

extern "C" int cnvb_bam_decode(cnvb_bam *b, float min_unclipped, int pinned) {
    if (!b) return CNVB_ERR_INVALID;
    if (!b->error.empty()) return CNVB_ERR_STATE;
    for (auto &r : b->refs) {
        r.n = 0;
        r.pos.n = r.end.n = r.mid.n = r.frag_end.n = r.mapq.n = r.flags.n = 0;
    }
    b->n_records = b->n_unplaced = b->bytes_compressed = b->bytes_uncompressed = 0;
    b->t_inflate = b->t_parse = b->t_total = 0.0;
    const int rc = decode(b, min_unclipped, pinned != 0, false);
    b->decoded = rc == CNVB_OK;
    return rc;
}

namespace {

// the .bai beside the file: <path>.bai, else <path without .bam>.bai
int load_bai(cnvb_bam *b) {
    if (b->bai_loaded) return CNVB_OK;
    std::string p1 = b->path + ".bai", p2;
    if (b->path.size() > 4 && b->path.compare(b->path.size() - 4, 4, ".bam") == 0) p2 = b->path.substr(0, b->path.size() - 4) + ".bai";
    FILE *f = fopen(p1.c_str(), "rb");
    if (!f && !p2.empty()) f = fopen(p2.c_str(), "rb");
    if (!f) return fail_to(b->error, CNVB_ERR_INVALID, "Could not open BAI-indexed BAM file %s: no index", b->path.c_str());
    std::vector<uint8_t> d;
    {
        fseek(f, 0, SEEK_END);
        const long sz = ftell(f);
        fseek(f, 0, SEEK_SET);
        d.resize(sz > 0 ? (size_t)sz : 0);
        const size_t got = d.empty() ? 0 : fread(d.data(), 1, d.size(), f);
        fclose(f);
        if (got != d.size()) return fail_to(b->error, CNVB_ERR_INVALID, "Could not read the index of %s", b->path.c_str());
    }
    auto bad = [&]() { return fail_to(b->error, CNVB_ERR_INVALID, "%s: corrupt .bai index", b->path.c_str()); };
    if (d.size() < 8 || memcmp(d.data(), "BAI\1", 4) != 0) return bad();
    const uint32_t n_ref = rd32(d.data() + 4);
    uint64_t at = 8;
    std::vector<BaiRef> refs(n_ref);
    auto rd64 = [&](uint64_t o) { uint64_t v; memcpy(&v, d.data() + o, 8); return v; };
    for (uint32_t r = 0; r < n_ref; ++r) {
        if (at + 4 > d.size()) return bad();
        const uint32_t n_bin = rd32(d.data() + at);
        at += 4;
        for (uint32_t k = 0; k < n_bin; ++k) {
            if (at + 8 > d.size()) return bad();
            const uint32_t bin = rd32(d.data() + at), n_chunk = rd32(d.data() + at + 4);
            at += 8;
            if (at + 16ull * n_chunk > d.size()) return bad();
            auto &v = refs[r].bins[bin];
            for (uint32_t c = 0; c < n_chunk; ++c) v.emplace_back(rd64(at + 16ull * c), rd64(at + 16ull * c + 8));
            at += 16ull * n_chunk;
        }
        if (at + 4 > d.size()) return bad();
        const uint32_t n_intv = rd32(d.data() + at);
        at += 4;
        if (at + 8ull * n_intv > d.size()) return bad();
        refs[r].linear.resize(n_intv);
        for (uint32_t i = 0; i < n_intv; ++i) refs[r].linear[i] = rd64(at + 8ull * i);
        at += 8ull * n_intv;
    }
    b->bai.swap(refs);
    b->bai_loaded = true;
    return CNVB_OK;
}

}  // namespace

// `fetch(tid, start, end)` of bam::IndexedReader (lib-coverage/src/lib.rs:533-543) through the .bai index: the bins
// overlapping the interval give the candidate chunks, the linear index bounds them from below, the decoder seeks to
// the first of them and reads until a record lies past the interval; of the records met, those whose alignment
// overlaps [start, end) -- what htslib's iterator yields, in file order.  The records replace the columns of `tid`.
extern "C" int cnvb_bam_fetch(cnvb_bam *b, uint32_t tid, uint32_t start, uint32_t end, float min_unclipped, int pinned) {
    if (!b) return CNVB_ERR_INVALID;
    if (tid >= b->refs.size()) return fail_to(b->error, CNVB_ERR_INVALID, "no such reference: %u", tid);
    CNVB_TRY(load_bai(b));
    RefReads &rr = b->refs[tid];
    rr.n = 0;
    rr.pos.n = rr.end.n = rr.mid.n = rr.frag_end.n = rr.mapq.n = rr.flags.n = 0;
    b->decoded = true;  // (the columns of this reference are valid after the call; others keep what they had)
    if (end <= start || tid >= b->bai.size()) return CNVB_OK;
    const BaiRef &ix = b->bai[tid];
    const uint64_t w = (uint64_t)start >> 14;
    uint64_t min_off = 0;
    if (!ix.linear.empty()) min_off = w < ix.linear.size() ? ix.linear[w] : ix.linear.back();
    uint64_t first = ~0ull;
    {   // reg2bins(start, end) of the BAM specification (min_shift 14, depth 5)
        const uint32_t e = end - 1;
        uint32_t t = 0;
        for (int l = 0, s = 29; l <= 5; ++l, s -= 3) {
            const uint32_t lo = t + (start >> s), hi = t + (e >> s);
            for (uint32_t bin = lo; bin <= hi; ++bin) {
                auto it = ix.bins.find(bin);
                if (it == ix.bins.end()) continue;
                for (auto &c : it->second)
                    if (c.second > min_off) first = std::min(first, c.first);
            }
            t += 1u << (3 * l);
        }
    }
    if (first == ~0ull) return CNVB_OK;
    FetchRange range{first >> 16, (uint32_t)(first & 0xffff), (int32_t)tid, (int32_t)start, (int32_t)end};
    b->n_records = b->n_unplaced = b->bytes_compressed = b->bytes_uncompressed = 0;
    b->t_inflate = b->t_parse = b->t_total = 0.0;
    return decode(b, min_unclipped, pinned != 0, false, &range);
}

extern "C" int cnvb_bam_ref_reads(const cnvb_bam *b, uint32_t tid, cnvb_read_soa *out) {
    if (!b || !out || tid >= b->refs.size() || !b->decoded) return CNVB_ERR_INVALID;
    const RefReads &r = b->refs[tid];
    memset(out, 0, sizeof *out);
    out->n = r.n;
    out->pos = r.pos.p;
    out->end = r.end.p;
    out->mid = r.mid.p;
    out->frag_end = r.frag_end.p;
    out->mapq = r.mapq.p;
    out->flags = r.flags.p;
    out->mem = CNVB_MEM_HOST;
    return CNVB_OK;
}

extern "C" int cnvb_bam_stats(const cnvb_bam *b, uint64_t counts[4], double seconds[3]) {
    if (!b) return CNVB_ERR_INVALID;
    if (counts) {
        counts[0] = b->n_records;
        counts[1] = b->n_unplaced;
        counts[2] = b->bytes_compressed;
        counts[3] = b->bytes_uncompressed;
    }
    if (seconds) {
        seconds[0] = b->t_total;
        seconds[1] = b->t_inflate;
        seconds[2] = b->t_parse;
    }
    return CNVB_OK;
}

extern "C" void cnvb_bam_close(cnvb_bam *b) { delete b; }
