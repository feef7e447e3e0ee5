// run.cu -- the region loop of `lib_coverage::run` (lib-coverage/src/lib.rs:768-781) and
// `process_region` (:422-676) as ONE native call over in-memory regions.
//
// The reference walks the regions one by one: load the contig for GC/gap statistics (:440-454),
// construct the aggregator (:510-529), feed the fetched records (:547), drain get_stats per
// window (:560-561) and assemble the FORMAT values (:623-668).  Here every stage of every region
// is enqueued back to back on the ctx stream, the per-window columns of all regions land in one
// table (one column per BCF tag), and the host waits exactly once, for the tallies and
// reference-panic flags of all regions.  Doing the loop natively keeps the stream fed: a
// 50 Mbp contig is ~45 us of GPU work, less than an interpreter spends per region.
#include <vector>

#include "common.cuh"

using namespace cnvb;

namespace {

uint64_t rows_of(const cnvb_cov_opts *o, int kind, const cnvb_region *r) {
    if (kind == CNVB_FRAGMENTS_TARGET_REGIONS) return r->n_tgt;
    const uint64_t wl = o->window_length;
    return wl ? ((uint64_t)r->region_end + wl - 1) / wl : 0;
}

struct Cols {  // device views
    uint32_t *start = nullptr, *end = nullptr;
    float *rcv = nullptr, *rcvsd = nullptr, *lcv = nullptr, *lcvsd = nullptr, *frm = nullptr, *gc = nullptr,
          *mq = nullptr;
    uint8_t *gap = nullptr, *ft = nullptr;
};

}  // namespace

extern "C" uint64_t cnvb_coverage_rows(const cnvb_cov_opts *opts, int kind, const cnvb_region *regions,
                                       uint32_t n_regions) {
    if (!opts || (n_regions && !regions)) return 0;
    uint64_t total = 0;
    for (uint32_t i = 0; i < n_regions; ++i) total += rows_of(opts, kind, &regions[i]);
    return total;
}

// What a deferred run leaves behind: the aggregators (their tallies arrive in pinned slots, in stream order)
// and the pooled temporaries, handed back by cnvb_coverage_wait.
struct cnvb_cov_pending {
    cnvb_ctx *ctx = nullptr;
    std::vector<cnvb_cov *> aggs;
    std::vector<void *> temps;
};

namespace {

int coverage_complete(cnvb_ctx *ctx, std::vector<cnvb_cov *> &aggs, uint64_t *num_processed, uint64_t *num_skipped) {
    // the single synchronisation point: tallies + reference-panic checks of every region, in order
    for (uint32_t i = 0; i < aggs.size(); ++i) {
        const int rc = cnvb_cov_finish(aggs[i]);
        if (rc) {
            const std::string inner = ctx->err;
            return fail(ctx, rc, "region %u\ncaused by: %s", i, inner.c_str());
        }
        if (num_processed) num_processed[i] = cnvb_cov_num_processed(aggs[i]);
        if (num_skipped) num_skipped[i] = cnvb_cov_num_skipped(aggs[i]);
    }
    return CNVB_OK;
}

int coverage_run_impl(cnvb_ctx *ctx, const cnvb_cov_opts *opts, int kind, const cnvb_region *regions, uint32_t n_regions,
                      const cnvb_cov_columns *out, uint64_t *num_processed, uint64_t *num_skipped, int mem,
                      cnvb_cov_pending **deferred) {
    if (!ctx) return CNVB_ERR_INVALID;
    if (!opts || !out || (n_regions && !regions))
        return fail(ctx, CNVB_ERR_INVALID, "cnvb_coverage_run: null argument");
    if (kind != CNVB_FRAGMENTS_GENOME_WIDE && kind != CNVB_FRAGMENTS_TARGET_REGIONS && kind != CNVB_COVERAGE)
        return fail(ctx, CNVB_ERR_INVALID, "Invalid combination of coverage/on-target regions");
    const bool is_cov = kind == CNVB_COVERAGE, is_tgt = kind == CNVB_FRAGMENTS_TARGET_REGIONS;
    if ((out->rcvsd || out->lcvsd) && !is_cov)
        return fail(ctx, CNVB_ERR_INVALID, "RCVSD/LCVSD exist only for count kind Coverage (lib.rs:641-645)");
    if (out->frm && kind != CNVB_FRAGMENTS_GENOME_WIDE)
        return fail(ctx, CNVB_ERR_INVALID, "FRM exists only for genome-wide fragment counts (lib.rs:647-654)");
    if ((out->gc || out->gap) && is_tgt)
        return fail(ctx, CNVB_ERR_INVALID, "GC/GAP are computed per window, not per target (lib.rs:440)");
    const uint64_t total = cnvb_coverage_rows(opts, kind, regions, n_regions);
    if (total && (!out->rcv || !out->start || !out->end || !out->lcv))
        return fail(ctx, CNVB_ERR_INVALID, "cnvb_coverage_run: START/END/RCV/LCV columns are required");
    CNVB_CUDA(ctx, cudaSetDevice(ctx->device), "selecting device");

    // device columns: the caller's (DEVICE) or pooled temporaries (HOST)
    Cols d;
    std::vector<void *> temps;
    std::vector<cnvb_cov *> aggs;
    bool forked = false;
    auto cleanup = [&](int rc) {
        if (rc != 0 && forked)  // an error path leaves work in flight on the side streams: wait before buffers are recycled
            for (int q = 0; q < 3; ++q) cudaStreamSynchronize(ctx->side[q]);
        for (cnvb_cov *a : aggs) cnvb_cov_destroy(a);
        for (void *p : temps) pool_release(ctx, p);
        return rc;
    };
    auto column = [&](auto *user, size_t elem, auto **dev) -> cudaError_t {
        using T = std::remove_pointer_t<decltype(user)>;
        if (!user) return cudaSuccess;
        if (mem == CNVB_MEM_DEVICE) {
            *dev = user;
            return cudaSuccess;
        }
        void *p = nullptr;
        cudaError_t e = pool_alloc(ctx, (total ? total : 1) * elem, &p);
        if (e == cudaSuccess) {
            temps.push_back(p);
            *dev = static_cast<T *>(p);
        }
        return e;
    };
    cudaError_t e;
    if ((e = column(out->start, 4, &d.start)) != cudaSuccess || (e = column(out->end, 4, &d.end)) != cudaSuccess ||
        (e = column(out->rcv, 4, &d.rcv)) != cudaSuccess || (e = column(out->rcvsd, 4, &d.rcvsd)) != cudaSuccess ||
        (e = column(out->lcv, 4, &d.lcv)) != cudaSuccess || (e = column(out->lcvsd, 4, &d.lcvsd)) != cudaSuccess ||
        (e = column(out->frm, 4, &d.frm)) != cudaSuccess || (e = column(out->gc, 4, &d.gc)) != cudaSuccess ||
        (e = column(out->mean_mapq, 4, &d.mq)) != cudaSuccess || (e = column(out->gap, 1, &d.gap)) != cudaSuccess ||
        (e = column(out->ft, 1, &d.ft)) != cudaSuccess)
        return cleanup(fail_cuda(ctx, e, "allocating the coverage table"));
    if (is_cov && d.lcvsd && !d.rcvsd) {  // LCVSD is derived from RCVSD
        void *p = nullptr;
        if ((e = pool_alloc(ctx, (total ? total : 1) * 4, &p)) != cudaSuccess)
            return cleanup(fail_cuda(ctx, e, "allocating the coverage table"));
        temps.push_back(p);
        d.rcvsd = static_cast<float *>(p);
    }

    aggs.reserve(n_regions);
    // Device-resident inputs: consecutive regions go to different side streams (fork / join with events),
    // so the tail of one region's persistent kernels overlaps the head of the next region's and the small
    // per-region operations (memsets, window statistics, tally copies) run beside the big ones.  Host
    // inputs share one staging buffer and stay on the context stream.
    // (Per-kernel timing mode keeps everything on one stream: a kernel's duration is only meaningful
    // when it has the device to itself.)
    bool fan = n_regions > 1 && !ctx->timing;
    for (uint32_t i = 0; i < n_regions && fan; ++i)
        fan = regions[i].reads.mem == CNVB_MEM_DEVICE && (!regions[i].seq || regions[i].seq_mem == CNVB_MEM_DEVICE);
    cudaStream_t main_stream = ctx->stream;
    if (fan) {
        for (int q = 0; q < 3 && fan; ++q) {
            if (!ctx->side[q] && cudaStreamCreateWithFlags(&ctx->side[q], cudaStreamNonBlocking) != cudaSuccess) fan = false;
            if (fan && !ctx->side_join[q] && cudaEventCreateWithFlags(&ctx->side_join[q], cudaEventDisableTiming) != cudaSuccess)
                fan = false;
        }
        if (fan && !ctx->side_fork && cudaEventCreateWithFlags(&ctx->side_fork, cudaEventDisableTiming) != cudaSuccess) fan = false;
    }
    if (fan) {
        if (d.gc) {
            const int rc = refstats_prepare(ctx);  // the shared table is built before the streams fork
            if (rc) return cleanup(rc);
        }
        if ((e = cudaEventRecord(ctx->side_fork, main_stream)) != cudaSuccess) return cleanup(fail_cuda(ctx, e, "forking streams"));
        for (int q = 0; q < 3; ++q)
            if ((e = cudaStreamWaitEvent(ctx->side[q], ctx->side_fork, 0)) != cudaSuccess)
                return cleanup(fail_cuda(ctx, e, "forking streams"));
        forked = true;
    }
    struct StreamGuard {  // whatever happens, the context gets its own stream back
        cnvb_ctx *c;
        cudaStream_t s;
        ~StreamGuard() { c->stream = s; }
    } stream_guard{ctx, main_stream};
    // lib.rs:440-454 for all regions at once when every sequence is in device memory: one launch instead of
    // one per region (each costs ~12 us before its first byte; refstats.cu).  It runs on the third side
    // stream beside the binning of the first regions.
    bool ref_batched = false;
    if (d.gc && d.gap && opts->window_length) {
        std::vector<const uint8_t *> bs;
        std::vector<uint64_t> bl, bo;
        bool all_dev = true;
        uint64_t o = 0;
        for (uint32_t i = 0; i < n_regions; ++i) {
            const cnvb_region *r = &regions[i];
            const uint64_t rows = rows_of(opts, kind, r);
            if (r->seq) {
                const uint64_t wl = opts->window_length;
                if (r->seq_mem != CNVB_MEM_DEVICE || (r->seq_len + wl - 1) / wl != rows) all_dev = false;
                bs.push_back(r->seq);
                bl.push_back(r->seq_len);
                bo.push_back(o);
            }
            o += rows;
        }
        static const bool batch_on = [] {
            const char *e = getenv("CNVB_REF_BATCH");
            return !(e && atoi(e) == 0);
        }();
        if (all_dev && bs.size() > 1 && batch_on) {
            if (fan) ctx->stream = ctx->side[2];
            const int rc = refstats_batch_device(ctx, (uint32_t)bs.size(), bs.data(), bl.data(), bo.data(),
                                                 opts->window_length, d.gc, d.gap, fan, &temps);
            ctx->stream = main_stream;
            if (rc < 0) return cleanup(rc);
            ref_batched = rc == 0;
        }
    }
    // FragmentsGenomeWide over device-resident regions: the regions are taken in a few GROUPS of consecutive
    // regions with about the same number of records.  Per group: the aggregators are constructed, ONE launch bins
    // the records of all its regions (coverage.cu: frag_bin_gw_batch_kernel -- one fill and one tail per group
    // instead of one per region), then the per-region statistics follow; consecutive groups alternate between two
    // side streams, so the small operations of one group run beside the binning of the next.
    // OFF unless CNVB_GW_BATCH = 1.  Measured on config 2 (22 regions): the binning alone drops from 1.10 to 0.93 ms
    // (one group; 0.96 with four) -- 6.8 TB/s, above the copy-measured HBM peak -- but the STEP grows from 1.65 to
    // 1.73 - 2.0 ms: the per-region form hides every region's small operations (two allocations, three memsets,
    // the tally copy, the window statistics: ~13 us per region) behind the binning of its neighbours on the other
    // stream, and a long persistent launch beside the persistent reference-statistics kernel leaves half of its
    // CTAs waiting for a free slot.  The batch form pays off only once those small operations are batched too.
    bool gw_try = kind == CNVB_FRAGMENTS_GENOME_WIDE && n_regions > 1 && [] {
        const char *e = getenv("CNVB_GW_BATCH");
        return e && atoi(e) == 1;
    }();
    for (uint32_t i = 0; i < n_regions && gw_try; ++i) gw_try = regions[i].reads.mem == CNVB_MEM_DEVICE;
    std::vector<uint32_t> gstart;  // first region of every group, then n_regions
    if (gw_try) {
        const uint32_t want_groups = [] {
            const char *e = getenv("CNVB_GW_GROUPS");
            const int v = e ? atoi(e) : 0;
            return (uint32_t)(v > 0 ? v : 4);
        }();
        const uint32_t G = std::max<uint32_t>(1, std::min<uint32_t>(want_groups, n_regions / 2));
        uint64_t total_reads = 0, seen = 0;
        for (uint32_t i = 0; i < n_regions; ++i) total_reads += regions[i].reads.n;
        gstart.push_back(0);
        for (uint32_t i = 0; i < n_regions; ++i) {
            // a group ends where the running share of the records passes the next multiple of 1 / G
            if (i > 0 && gstart.size() < G && seen * G >= (uint64_t)gstart.size() * total_reads) gstart.push_back(i);
            seen += regions[i].reads.n;
        }
        gstart.push_back(n_regions);
        aggs.resize(n_regions, nullptr);
    }
    std::vector<uint8_t> batched(n_regions, 0);
    uint32_t g_cur = 0;
    uint64_t off = 0;
    for (uint32_t i = 0; i < n_regions; ++i) {
        const cnvb_region *r = &regions[i];
        const uint64_t rows = rows_of(opts, kind, r);
        if (fan) ctx->stream = ctx->side[ref_batched ? i % 2 : i % 3];
        if (gw_try) {
            while (i >= gstart[g_cur + 1]) ++g_cur;
            if (fan) ctx->stream = ctx->side[g_cur % 2];
            if (i == gstart[g_cur]) {  // the group's aggregators, then its records in one launch
                const uint32_t g0 = gstart[g_cur], g1 = gstart[g_cur + 1];
                for (uint32_t q = g0; q < g1; ++q) {
                    const cnvb_region *rq = &regions[q];
                    const int rc = cnvb_cov_create(ctx, opts, kind, rq->region_end, rq->tgt_start, rq->tgt_end, rq->n_tgt,
                                                   rq->pile_start, rq->pile_end, rq->n_pile, &aggs[q]);
                    if (rc) return cleanup(rc);
                }
                static const int beside = [] {
                    const char *e = getenv("CNVB_GW_BESIDE");
                    return e ? atoi(e) : 2;
                }();
                const int rc = gw_batch_put(ctx, aggs.data() + g0, &regions[g0].reads, sizeof(cnvb_region), g1 - g0,
                                            fan && ref_batched ? beside : 0);
                if (rc < 0) return cleanup(rc);
                for (uint32_t q = g0; q < g1; ++q) batched[q] = rc == 0;
            }
        }
        // lib.rs:440-454: GC content and gap flag per window from the contig bytes
        if (r->seq && (d.gc || d.gap) && !ref_batched) {
            if (!d.gc || !d.gap)
                return cleanup(fail(ctx, CNVB_ERR_INVALID, "GC and GAP columns come together"));
            const uint64_t wl = opts->window_length;
            if ((r->seq_len + wl - 1) / wl != rows)
                return cleanup(fail(ctx, CNVB_ERR_INVALID,
                                    "region %u: sequence length %llu does not match the region end %u", i,
                                    (unsigned long long)r->seq_len, r->region_end));
            const uint8_t *dseq = r->seq;
            void *stage = nullptr;
            if (r->seq_mem != CNVB_MEM_DEVICE) {
                if ((e = pool_alloc(ctx, r->seq_len ? r->seq_len : 1, &stage)) != cudaSuccess ||
                    (e = cudaMemcpyAsync(stage, r->seq, r->seq_len, cudaMemcpyHostToDevice, ctx->stream)) !=
                        cudaSuccess) {
                    pool_release(ctx, stage);
                    return cleanup(fail_cuda(ctx, e, "uploading the contig sequence"));
                }
                dseq = static_cast<const uint8_t *>(stage);
            }
            const int rc = cnvb_refstats(ctx, dseq, r->seq_len, opts->window_length, d.gc + off, d.gap + off,
                                         CNVB_MEM_DEVICE);
            pool_release(ctx, stage);  // stream-ordered reuse
            if (rc) return cleanup(rc);
        }
        // lib.rs:510-529, :547, :560-561
        cnvb_cov *agg = gw_try ? aggs[i] : nullptr;
        int rc = 0;
        if (!gw_try) {
            rc = cnvb_cov_create(ctx, opts, kind, r->region_end, r->tgt_start, r->tgt_end, r->n_tgt,
                                 r->pile_start, r->pile_end, r->n_pile, &agg);
            if (rc) return cleanup(rc);
            aggs.push_back(agg);
        }
        if ((!batched[i] && (rc = cnvb_cov_put_records(agg, &r->reads)) != 0) || (rc = cnvb_cov_finish_async(agg)) != 0 ||
            (rc = cnvb_cov_get_stats(agg, d.rcv + off, d.rcvsd ? d.rcvsd + off : nullptr, d.start + off,
                                     d.end + off, d.frm ? d.frm + off : nullptr, d.mq ? d.mq + off : nullptr,
                                     CNVB_MEM_DEVICE)) != 0)
            return cleanup(rc);
        off += rows;
    }
    ctx->stream = main_stream;
    if (fan) {
        for (int q = 0; q < 3; ++q) {
            if ((e = cudaEventRecord(ctx->side_join[q], ctx->side[q])) != cudaSuccess ||
                (e = cudaStreamWaitEvent(main_stream, ctx->side_join[q], 0)) != cudaSuccess)
                return cleanup(fail_cuda(ctx, e, "joining streams"));
        }
    }
    // lib.rs:623-668 for all rows at once
    if (total) {
        const int rc = cnvb_cov_records(ctx, total, d.rcv, d.rcvsd, d.start, d.end, d.frm, is_tgt ? 1 : 0, opts,
                                        d.lcv, d.lcvsd, d.ft, CNVB_MEM_DEVICE);
        if (rc) return cleanup(rc);
    }
    if (mem != CNVB_MEM_DEVICE && total) {
        auto back = [&](void *dst, const void *src, size_t bytes) -> cudaError_t {
            return dst ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream) : cudaSuccess;
        };
        if ((e = back(out->start, d.start, total * 4)) != cudaSuccess || (e = back(out->end, d.end, total * 4)) != cudaSuccess ||
            (e = back(out->rcv, d.rcv, total * 4)) != cudaSuccess || (e = back(out->rcvsd, d.rcvsd, total * 4)) != cudaSuccess ||
            (e = back(out->lcv, d.lcv, total * 4)) != cudaSuccess || (e = back(out->lcvsd, d.lcvsd, total * 4)) != cudaSuccess ||
            (e = back(out->frm, d.frm, total * 4)) != cudaSuccess || (e = back(out->gc, d.gc, total * 4)) != cudaSuccess ||
            (e = back(out->mean_mapq, d.mq, total * 4)) != cudaSuccess || (e = back(out->gap, d.gap, total)) != cudaSuccess ||
            (e = back(out->ft, d.ft, total)) != cudaSuccess)
            return cleanup(fail_cuda(ctx, e, "copying the coverage table to the host"));
    }
    if (deferred) {  // the caller collects tallies and checks later (cnvb_coverage_wait); nothing waits here
        cnvb_cov_pending *p = new (std::nothrow) cnvb_cov_pending();
        if (!p) return cleanup(fail(ctx, CNVB_ERR_NOMEM, "out of host memory"));
        p->ctx = ctx;
        p->aggs.swap(aggs);
        p->temps.swap(temps);
        *deferred = p;
        return CNVB_OK;
    }
    {
        const int rc = coverage_complete(ctx, aggs, num_processed, num_skipped);
        if (rc) return cleanup(rc);
    }
    if (mem != CNVB_MEM_DEVICE)
        if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess)
            return cleanup(fail_cuda(ctx, e, "copying the coverage table to the host"));
    return cleanup(CNVB_OK);
}

}  // namespace

extern "C" int cnvb_coverage_run(cnvb_ctx *ctx, const cnvb_cov_opts *opts, int kind,
                                 const cnvb_region *regions, uint32_t n_regions,
                                 const cnvb_cov_columns *out, uint64_t *num_processed,
                                 uint64_t *num_skipped, int mem) {
    return coverage_run_impl(ctx, opts, kind, regions, n_regions, out, num_processed, num_skipped, mem, nullptr);
}

extern "C" int cnvb_coverage_run_async(cnvb_ctx *ctx, const cnvb_cov_opts *opts, int kind,
                                       const cnvb_region *regions, uint32_t n_regions,
                                       const cnvb_cov_columns *out, cnvb_cov_pending **pending) {
    if (!ctx) return CNVB_ERR_INVALID;
    if (!pending) return fail(ctx, CNVB_ERR_INVALID, "cnvb_coverage_run_async: null argument");
    *pending = nullptr;
    for (uint32_t i = 0; i < n_regions && regions; ++i)
        if (regions[i].reads.mem != CNVB_MEM_DEVICE || (regions[i].seq && regions[i].seq_mem != CNVB_MEM_DEVICE))
            return fail(ctx, CNVB_ERR_INVALID, "cnvb_coverage_run_async: inputs must be device-resident (region %u)", i);
    return coverage_run_impl(ctx, opts, kind, regions, n_regions, out, nullptr, nullptr, CNVB_MEM_DEVICE, pending);
}

extern "C" int cnvb_coverage_wait(cnvb_cov_pending *p, uint64_t *num_processed, uint64_t *num_skipped) {
    if (!p) return CNVB_ERR_INVALID;
    cnvb_ctx *ctx = p->ctx;
    const int rc = coverage_complete(ctx, p->aggs, num_processed, num_skipped);
    for (cnvb_cov *a : p->aggs) cnvb_cov_destroy(a);
    for (void *t : p->temps) pool_release(ctx, t);
    delete p;
    return rc;
}

// ---- the region loop as a CUDA graph --------------------------------------------------------------------
// Repeated runs over the same device buffers (a pipeline that copies every sample into fixed device columns, the
// strong-scaling step of bench.py) are bound by the HOST once the regions are small: enqueueing the ~25 operations
// of a three-region run takes 0.25 ms of CPU, the GPU needs 0.15 ms for them (tools/scale_probe.py on 8 GPUs).  A
// plan runs the loop once under stream capture -- side streams, memsets, kernels, tally copies and all -- and
// replays the instantiated graph with ONE launch.  Everything the graph touches is owned by the plan: the
// aggregators, the pooled temporaries, the blocks released during the capture, the pinned tables and tally slots.
struct cnvb_cov_plan {
    cnvb_ctx *ctx = nullptr;
    cnvb_cov_pending *pending = nullptr;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    cudaEvent_t done = nullptr;  // recorded behind every launch: plan_wait waits for the plan's own work only
    uint64_t kernels = 0;        // kernel launches one replay stands for (cnvb_ctx_launch_count counts those)
    std::vector<void *> hold, pinned;
};

extern "C" void cnvb_coverage_plan_destroy(cnvb_cov_plan *p) {
    if (!p) return;
    cnvb_ctx *ctx = p->ctx;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (p->exec) cudaGraphExecDestroy(p->exec);
    if (p->graph) cudaGraphDestroy(p->graph);
    if (p->done) cudaEventDestroy(p->done);
    if (p->pending) {
        for (cnvb_cov *a : p->pending->aggs) cnvb_cov_destroy(a);
        for (void *t : p->pending->temps) pool_release(ctx, t);
        delete p->pending;
    }
    for (void *b : p->hold) pool_release(ctx, b);
    for (void *h : p->pinned) cudaFreeHost(h);
    delete p;
}

extern "C" int cnvb_coverage_plan_create(cnvb_ctx *ctx, const cnvb_cov_opts *opts, int kind, const cnvb_region *regions,
                                         uint32_t n_regions, const cnvb_cov_columns *out, cnvb_cov_plan **plan) {
    if (!ctx) return CNVB_ERR_INVALID;
    if (!plan) return fail(ctx, CNVB_ERR_INVALID, "cnvb_coverage_plan_create: null argument");
    *plan = nullptr;
    if (kind != CNVB_FRAGMENTS_GENOME_WIDE)
        return fail(ctx, CNVB_ERR_INVALID, "a coverage plan exists for genome-wide fragment counts: the other kinds size "
                                          "their launches from the records (targets tables, the longest alignment)");
    if (ctx->timing) return fail(ctx, CNVB_ERR_STATE, "a plan cannot be built in per-kernel timing mode");
    if (ctx->capturing) return fail(ctx, CNVB_ERR_STATE, "a plan is already being captured on this context");
    // one eager run first: pooled blocks of every size the loop asks for, function attributes, the classification
    // table, the side streams -- nothing is created while the capture is open
    {
        cnvb_cov_pending *warm = nullptr;
        CNVB_TRY(cnvb_coverage_run_async(ctx, opts, kind, regions, n_regions, out, &warm));
        CNVB_TRY(cnvb_coverage_wait(warm, nullptr, nullptr));
    }
    CNVB_CUDA(ctx, cudaDeviceSynchronize(), "before the capture");
    pool_quiesce(ctx);
    cnvb_cov_plan *p = new (std::nothrow) cnvb_cov_plan();
    if (!p) return fail(ctx, CNVB_ERR_NOMEM, "out of host memory");
    p->ctx = ctx;
    // the capture runs on a stream of its own (the context's may be the legacy default stream, which cannot be
    // captured); the graph is launched on the context's stream afterwards
    cudaStream_t cap = nullptr, user = ctx->stream;
    cudaError_t e = cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamBeginCapture(cap, cudaStreamCaptureModeRelaxed);
    if (e != cudaSuccess) {
        if (cap) cudaStreamDestroy(cap);
        delete p;
        return fail_cuda(ctx, e, "opening the capture");
    }
    ctx->capturing = true;
    ctx->stream = cap;
    const uint64_t launches_before = ctx->launches;
    const int rc = coverage_run_impl(ctx, opts, kind, regions, n_regions, out, nullptr, nullptr, CNVB_MEM_DEVICE, &p->pending);
    ctx->stream = user;
    ctx->capturing = false;
    p->kernels = ctx->launches - launches_before;
    ctx->launches = launches_before;  // (captured, not launched)
    p->hold.swap(ctx->capture_hold);
    p->pinned.swap(ctx->capture_pinned);
    e = cudaStreamEndCapture(cap, &p->graph);
    cudaStreamDestroy(cap);
    if (rc != CNVB_OK || e != cudaSuccess || !p->graph) {
        const std::string inner = ctx->err;
        // (the blocks named in pool_live stay valid; an unfinished capture leaves no work behind)
        cnvb_coverage_plan_destroy(p);
        if (rc != CNVB_OK) return fail(ctx, rc, "capturing the region loop\ncaused by: %s", inner.c_str());
        return fail_cuda(ctx, e != cudaSuccess ? e : cudaErrorUnknown, "closing the capture");
    }
    e = cudaGraphInstantiate(&p->exec, p->graph, 0);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->done, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        cnvb_coverage_plan_destroy(p);
        return fail_cuda(ctx, e, "instantiating the graph");
    }
    *plan = p;
    return CNVB_OK;
}

// one run of the planned loop, enqueued on the context's stream; returns at once
extern "C" int cnvb_coverage_plan_launch(cnvb_cov_plan *p) {
    if (!p || !p->exec) return CNVB_ERR_INVALID;
    cnvb_ctx *ctx = p->ctx;
    CNVB_CUDA(ctx, cudaGraphLaunch(p->exec, ctx->stream), "launching the planned region loop");
    CNVB_CUDA(ctx, cudaEventRecord(p->done, ctx->stream), "marking the end of the planned region loop");
    ctx->launches += p->kernels;
    return CNVB_OK;
}

// waits for the plan's launches so far (not for later work on the stream); tallies and reference-panic checks of
// the last one (nullable outputs)
extern "C" int cnvb_coverage_plan_wait(cnvb_cov_plan *p, uint64_t *num_processed, uint64_t *num_skipped) {
    if (!p || !p->pending) return CNVB_ERR_INVALID;
    cnvb_ctx *ctx = p->ctx;
    CNVB_CUDA(ctx, cudaEventSynchronize(p->done), "waiting for the planned region loop");
    std::vector<cnvb_cov *> &aggs = p->pending->aggs;
    for (uint32_t i = 0; i < aggs.size(); ++i) {
        const int rc = cov_collect(aggs[i]);
        if (rc) {
            const std::string inner = ctx->err;
            return fail(ctx, rc, "region %u\ncaused by: %s", i, inner.c_str());
        }
        if (num_processed) num_processed[i] = cnvb_cov_num_processed(aggs[i]);
        if (num_skipped) num_skipped[i] = cnvb_cov_num_skipped(aggs[i]);
    }
    return CNVB_OK;
}
