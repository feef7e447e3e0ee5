// vcf.cu -- host writer of the coverage VCF (`cmd coverage` / `cmd normalize` output).
//
// The reference assembles one htslib record per window (lib-coverage/src/lib.rs:560-673: CHROM, POS, ID
// "chrom:pos+1-end", REF N, ALT <WINDOW>|<TARGET>, INFO END [GAP] [GC] MEAN_MAPQ, FORMAT GT RCV [RCVSD]
// LCV [LCVSD] [FRM] MQ [FT]; lib-normalize adds FORMAT CV [CV2] [CVSD] and INFO FEW_GCWINDOWS) under the
// header of lib-coverage/src/lib.rs:67-166, and lets the file extension choose the container
// (lib-shared/src/bcf_utils.rs:47-82).  This writer produces the text forms, `.vcf` and BGZF-compressed
// `.vcf.gz`, from the per-window columns of the region loop: rows are formatted on host threads, BGZF
// blocks are deflated on host threads, the file is written in order.  Binary `.bcf` (and the CSI / TBI
// index) is not implemented.  Number formatting follows printf %g, which is what htslib's text writer
// produces for the value ranges of this path; nothing in the reference pins it.
//
// Pure host code.
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

namespace {

void put_float(std::string &s, float v) {
    uint32_t bits;
    memcpy(&bits, &v, 4);
    if (bits == CNVB_F32_MISSING_BITS) {  // bcf_float_missing
        s += '.';
        return;
    }
    char buf[32];
    const int n = snprintf(buf, sizeof buf, "%g", (double)v);
    s.append(buf, (size_t)n);
}
void put_int(std::string &s, long long v) {
    char buf[24];
    const int n = snprintf(buf, sizeof buf, "%lld", v);
    s.append(buf, (size_t)n);
}

std::string bgzf_block(const uint8_t *data, size_t n, int level) {
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    std::string out(18 + compressBound((uLong)n) + 8, '\0');
    deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
    zs.next_in = const_cast<Bytef *>(data);
    zs.avail_in = (uInt)n;
    zs.next_out = reinterpret_cast<Bytef *>(&out[18]);
    zs.avail_out = (uInt)(out.size() - 18);
    deflate(&zs, Z_FINISH);
    const size_t clen = zs.total_out;
    deflateEnd(&zs);
    const uint8_t head[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
    memcpy(&out[0], head, 16);
    const uint16_t bsize = (uint16_t)(clen + 25);
    memcpy(&out[16], &bsize, 2);
    const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), data, (uInt)n), isize = (uint32_t)n;
    memcpy(&out[18 + clen], &crc, 4);
    memcpy(&out[18 + clen + 4], &isize, 4);
    out.resize(18 + clen + 8);
    return out;
}

void parallel_for(int n_threads, uint64_t n, const std::function<void(uint64_t, uint64_t)> &fn) {
    if (n_threads <= 1 || n < 2) {
        fn(0, n);
        return;
    }
    std::vector<std::thread> th;
    const uint64_t per = (n + n_threads - 1) / n_threads;
    for (int t = 0; t < n_threads; ++t) {
        const uint64_t a = std::min<uint64_t>(n, per * t), b = std::min<uint64_t>(n, a + per);
        if (a < b) th.emplace_back(fn, a, b);
    }
    for (auto &x : th) x.join();
}

struct Sink {  // plain or BGZF file
    FILE *f = nullptr;
    bool gz = false;
    int n_threads = 1;
    std::string pending;  // text not yet cut into blocks
    bool write_all(const std::string &s) { return fwrite(s.data(), 1, s.size(), f) == s.size(); }
    bool put(const std::string &text, bool final) {
        if (!gz) return write_all(text);
        pending += text;
        constexpr size_t kBlock = 0xff00;
        const size_t n_full = pending.size() / kBlock;
        const size_t n_blocks = n_full + ((final && pending.size() % kBlock) ? 1 : 0);
        std::vector<std::string> out(n_blocks);
        parallel_for(n_threads, n_blocks, [&](uint64_t a, uint64_t b) {
            for (uint64_t i = a; i < b; ++i) {
                const size_t off = i * kBlock, len = std::min(kBlock, pending.size() - off);
                out[i] = bgzf_block(reinterpret_cast<const uint8_t *>(pending.data()) + off, len, 6);
            }
        });
        for (auto &o : out)
            if (!write_all(o)) return false;
        pending.erase(0, std::min(pending.size(), n_blocks * kBlock));
        if (final) return write_all(bgzf_block(nullptr, 0, 6));  // EOF marker
        return true;
    }
};

int vfail(char *err, uint64_t err_len, int code, const std::string &msg) {
    if (err && err_len) snprintf(err, (size_t)err_len, "%s", msg.c_str());
    return code;
}

}  // namespace

extern "C" int cnvb_vcf_write_coverage(const cnvb_vcf_opts *o, const cnvb_vcf_columns *c, uint64_t n_rows, char *err,
                                       uint64_t err_len) {
    if (!o || !c || !o->path || !o->sample) return vfail(err, err_len, CNVB_ERR_INVALID, "cnvb_vcf_write_coverage: null argument");
    if (n_rows && (!c->contig || !c->start || !c->end || !c->rcv || !c->lcv))
        return vfail(err, err_len, CNVB_ERR_INVALID, "cnvb_vcf_write_coverage: contig, start, end, rcv and lcv columns are required");
    const std::string path = o->path;
    const bool gz = path.size() > 3 && path.compare(path.size() - 3, 3, ".gz") == 0;
    if (path.size() > 4 && path.compare(path.size() - 4, 4, ".bcf") == 0)
        return vfail(err, err_len, CNVB_ERR_INVALID, "binary BCF output is not implemented: use .vcf or .vcf.gz");
    Sink sink;
    sink.f = fopen(path.c_str(), "wb");
    if (!sink.f) return vfail(err, err_len, CNVB_ERR_INVALID, "Could not open " + path + " for writing");
    sink.gz = gz;
    const unsigned hw = std::thread::hardware_concurrency();
    sink.n_threads = o->n_threads > 0 ? o->n_threads : (hw ? (int)hw : 1);
    // ---- header: the tag set of lib-coverage/src/lib.rs:67-166 (IDs, Number, Type as there)
    std::string h = "##fileformat=VCFv4.2\n";
    char date[16];
    if (o->file_date) {
        snprintf(date, sizeof date, "%s", o->file_date);
    } else {
        const time_t now = time(nullptr);
        struct tm tmv;
        gmtime_r(&now, &tmv);
        strftime(date, sizeof date, "%Y%m%d", &tmv);
    }
    h += std::string("##fileDate=") + date + "\n";
    h += "##cnvetti_cmdCoverageVersion=0.1.0\n";
    h += std::string("##cnvetti_cmdCoverageCommand=") + (o->command_line ? o->command_line : "cnvetti cmd coverage") + "\n";
    for (uint32_t i = 0; i < o->n_contigs; ++i) {
        h += "##contig=<ID=" + std::string(o->contig_names[i]) + ",length=";
        put_int(h, o->contig_lens[i]);
        h += ">\n";
    }
    h += "##ALT=<ID=WINDOW,Description=\"Window for read or coverage counting\">\n"
         "##ALT=<ID=TARGET,Description=\"Target region for fragment counting\">\n"
         "##INFO=<ID=END,Number=1,Type=Integer,Description=\"Window end\">\n"
         "##INFO=<ID=GC,Number=1,Type=Float,Description=\"Reference GC fraction of the window\">\n"
         "##INFO=<ID=MEAN_MAPQ,Number=1,Type=Float,Description=\"Mean MAPQ across samples\">\n"
         "##INFO=<ID=FEW_GCWINDOWS,Number=0,Type=Flag,Description=\"Too few windows with this GC content for the GC correction\">\n"
         "##INFO=<ID=REF_MEAN,Number=1,Type=Flag,Description=\"Reference mean coverage\">\n"
         "##INFO=<ID=REF_STD_DEV,Number=1,Type=Flag,Description=\"Reference coverage standard deviation\">\n"
         "##FILTER=<ID=PILE,Description=\"Piles cover too large a fraction of the window\">\n"
         "##FILTER=<ID=LOWCOV,Description=\"Too low raw coverage (targeted data)\">\n"
         "##FILTER=<ID=NO_REF,Description=\"No reference targets for this target\">\n"
         "##INFO=<ID=GAP,Number=0,Type=Flag,Description=\"Window overlaps N in the reference\">\n"
         "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n"
         "##FORMAT=<ID=MQ,Number=1,Type=Float,Description=\"Mean read MAPQ of the region\">\n"
         "##FORMAT=<ID=RCV,Number=1,Type=Float,Description=\"Raw coverage value\">\n"
         "##FORMAT=<ID=RCVSD,Number=1,Type=Float,Description=\"Raw coverage standard deviation\">\n"
         "##FORMAT=<ID=LCV,Number=1,Type=Float,Description=\"Length-normalized coverage value\">\n"
         "##FORMAT=<ID=LCVSD,Number=1,Type=Float,Description=\"Length-normalized coverage standard deviation\">\n"
         "##FORMAT=<ID=NCV,Number=1,Type=Float,Description=\"Further normalized coverage value\">\n"
         "##FORMAT=<ID=CV,Number=1,Type=Float,Description=\"Finalized coverage value\">\n"
         "##FORMAT=<ID=CVSD,Number=1,Type=Float,Description=\"Finalized coverage standard deviation\">\n"
         "##FORMAT=<ID=CV2,Number=1,Type=Float,Description=\"Log2 of the finalized coverage value\">\n"
         "##FORMAT=<ID=CVZ,Number=1,Type=Float,Description=\"Finalized coverage as Z-score of the reference targets\">\n"
         "##FORMAT=<ID=FRM,Number=1,Type=Float,Description=\"Fraction of the window masked in the sample\">\n"
         "##FORMAT=<ID=FT,Number=1,Type=String,Description=\"Filter values of the genotype\">\n";
    h += std::string("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t") + o->sample + "\n";
    bool ok = sink.put(h, n_rows == 0);
    // ---- records, in batches (bounded memory), formatted on threads
    constexpr uint64_t kBatch = 1u << 20;
    const char *alt = o->is_targets ? "<TARGET>" : "<WINDOW>";
    for (uint64_t b0 = 0; ok && b0 < n_rows; b0 += kBatch) {
        const uint64_t nb = std::min<uint64_t>(kBatch, n_rows - b0);
        const int nt = sink.n_threads;
        std::vector<std::string> parts((size_t)nt);
        const uint64_t per = (nb + nt - 1) / nt;
        std::vector<std::thread> th;
        std::atomic<bool> bad_contig{false};
        for (int t = 0; t < nt; ++t) {
            const uint64_t lo = std::min<uint64_t>(nb, per * t), hi = std::min<uint64_t>(nb, lo + per);
            if (lo >= hi) continue;
            th.emplace_back([&, t, lo, hi]() {
                std::string &s = parts[(size_t)t];
                s.reserve((hi - lo) * 128);
                for (uint64_t r = b0 + lo; r < b0 + hi; ++r) {
                    if (c->contig[r] >= o->n_contigs) {
                        bad_contig = true;
                        return;
                    }
                    const char *chrom = o->contig_names[c->contig[r]];
                    const long long pos = c->start[r], wend = c->end[r];
                    s += chrom;
                    s += '\t';
                    put_int(s, pos + 1);
                    s += '\t';
                    s += chrom;  // ID chrom:pos+1-end (lib.rs:586-588)
                    s += ':';
                    put_int(s, pos + 1);
                    s += '-';
                    put_int(s, wend);
                    s += "\tN\t";
                    s += alt;
                    s += "\t.\t.\tEND=";
                    put_int(s, wend);
                    if (c->gc) {  // lib.rs:597-606: GAP flag, then GC
                        if (c->gap && c->gap[r]) s += ";GAP";
                        s += ";GC=";
                        put_float(s, c->gc[r]);
                    }
                    s += ";MEAN_MAPQ=";
                    put_float(s, c->mean_mapq ? c->mean_mapq[r] : 60.0f);
                    const uint8_t nf = c->norm_flags ? c->norm_flags[r] : 0;
                    if (nf & CNVB_N2_FEW_GCWINDOWS) s += ";FEW_GCWINDOWS";
                    const uint8_t ft = c->ft ? c->ft[r] : 0;
                    const bool has_cv = c->cv && (!c->norm_flags || (nf & CNVB_N2_CV));
                    const bool has_cv2 = c->cv2 && (!c->norm_flags || (nf & CNVB_N2_CV2));
                    const bool has_cvsd = c->cvsd && (!c->norm_flags || (nf & CNVB_N2_CVSD));
                    s += "\tGT:RCV";
                    if (c->rcvsd) s += ":RCVSD";
                    s += ":LCV";
                    if (c->lcvsd) s += ":LCVSD";
                    if (c->frm) s += ":FRM";
                    s += ":MQ";
                    if (ft) s += ":FT";
                    if (has_cv) s += ":CV";
                    if (has_cv2) s += ":CV2";
                    if (has_cvsd) s += ":CVSD";
                    s += "\t.:";
                    put_float(s, c->rcv[r]);
                    if (c->rcvsd) { s += ':'; put_float(s, c->rcvsd[r]); }
                    s += ':';
                    put_float(s, c->lcv[r]);
                    if (c->lcvsd) { s += ':'; put_float(s, c->lcvsd[r]); }
                    if (c->frm) { s += ':'; put_float(s, c->frm[r]); }
                    s += ':';
                    put_float(s, c->mean_mapq ? c->mean_mapq[r] : 60.0f);
                    if (ft) {  // lib.rs:659-664: joined with ';'
                        s += ':';
                        if (ft & CNVB_FT_LOWCOV) s += "LOWCOV";
                        if ((ft & CNVB_FT_LOWCOV) && (ft & CNVB_FT_PILE)) s += ';';
                        if (ft & CNVB_FT_PILE) s += "PILE";
                    }
                    if (has_cv) { s += ':'; put_float(s, c->cv[r]); }
                    if (has_cv2) { s += ':'; put_float(s, c->cv2[r]); }
                    if (has_cvsd) { s += ':'; put_float(s, c->cvsd[r]); }
                    s += '\n';
                }
            });
        }
        for (auto &x : th) x.join();
        if (bad_contig) {
            fclose(sink.f);
            return vfail(err, err_len, CNVB_ERR_INVALID, "a row names a contig index beyond the contig list");
        }
        std::string batch;
        size_t tot = 0;
        for (auto &p : parts) tot += p.size();
        batch.reserve(tot);
        for (auto &p : parts) batch += p;
        ok = sink.put(batch, b0 + nb == n_rows);
    }
    if (fclose(sink.f) != 0) ok = false;
    if (!ok) return vfail(err, err_len, CNVB_ERR_INVALID, "Could not write " + path);
    return CNVB_OK;
}
