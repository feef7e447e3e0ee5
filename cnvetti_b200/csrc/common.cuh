// common.cuh -- context, error chain and small device helpers shared by all kernels.
// Target: sm_100a (B200) only.
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <initializer_list>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/cnvetti_b200.h"

struct cnvb_span {  // one timed kernel launch (only while timing is switched on)
    int name_id;
    cudaEvent_t a, b;
};

struct cnvb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    uint64_t launches = 0;
    std::string err;
    // grow-only device scratch (outputs staged for CNVB_MEM_HOST callers, partial sums, ...)
    void *scratch = nullptr;
    size_t scratch_bytes = 0;
    // grow-only device staging for CNVB_MEM_HOST inputs
    void *stage = nullptr;
    size_t stage_bytes = 0;
    // caching device allocator: aggregators are created and destroyed once per region, and
    // cudaMalloc/cudaFree would serialise the device every time
    // (stream-aware: a free block remembers the stream it was released on; handing it to another
    // stream first orders that stream behind the releasing one, see pool_alloc)
    struct pool_block {
        void *p;
        cudaStream_t released_on;
    };
    std::map<size_t, std::vector<pool_block>> pool_free;
    std::unordered_map<void *, size_t> pool_live;
    cudaEvent_t pool_event = nullptr;
    // pinned host slots for asynchronous read-back of per-aggregator tallies
    uint32_t *tally_slots = nullptr;  // kTallySlots x 4 u32
    uint32_t tally_next = 0;
    // resident CTAs per SM of each persistent kernel (cudaOccupancy... is not free: cached)
    std::unordered_map<const void *, int> occupancy;
    uint64_t wis_stats[4] = {0, 0, 0, 0};
    // refstats: the byte-pair classification table (128 KB), built once per context
    void *ref_lut = nullptr;
    // side streams of the region loop (cnvb_coverage_run): consecutive regions run on different
    // streams so that the ramp-down of one region's kernels overlaps the ramp-up of the next
    cudaStream_t side[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t side_fork = nullptr, side_join[3] = {nullptr, nullptr, nullptr};
    // optional per-kernel timing with CUDA events on the launching stream (bench.py roofline)
    bool timing = false;
    std::vector<std::string> span_names;
    std::vector<cnvb_span> spans;
    std::vector<cudaEvent_t> event_pool;
    // a coverage plan is being captured into a CUDA graph (run.cu: cnvb_coverage_plan_create): blocks released
    // meanwhile are HELD (the graph's nodes keep using them at every replay), host tables the graph reads and
    // the tally slots live in pinned memory the plan owns
    bool capturing = false;
    std::vector<void *> capture_hold;     // device blocks released during the capture
    std::vector<void *> capture_pinned;   // cudaMallocHost blocks
};

namespace cnvb {

int refstats_prepare(cnvb_ctx *ctx);  // refstats.cu: builds the classification table on ctx->stream if missing
// refstats.cu: reference statistics of all regions in one launch; 1 = not applicable (go region by region)
// FragmentsGenomeWide records of all regions of a run in one launch (coverage.cu); 1 = does not apply
// coverage.cu: the tallies and reference-panic checks of an aggregator whose stream work is known to be complete
// (cnvb_cov_finish without the event wait: a plan synchronises once for all its aggregators)
int cov_collect(cnvb_cov *a);
int gw_batch_put(cnvb_ctx *ctx, cnvb_cov *const *aggs, const cnvb_read_soa *batches, size_t batch_stride, uint32_t n,
                 int ctas_per_sm);
int refstats_batch_device(cnvb_ctx *ctx, uint32_t n, const uint8_t *const *seq, const uint64_t *len, const uint64_t *out_off,
                          uint32_t window_length, float *gc, uint8_t *has_gap, bool shares_sms,
                          std::vector<void *> *defer_release);

// error chain in the style of error-chain's "error: ... / caused by: ..." (cnvetti/src/main.rs)
inline int fail(cnvb_ctx *ctx, int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    return code;
}

inline int fail_cuda(cnvb_ctx *ctx, cudaError_t e, const char *what) {
    return fail(ctx, CNVB_ERR_CUDA, "%s\ncaused by: CUDA error %d (%s): %s", what, (int)e,
                cudaGetErrorName(e), cudaGetErrorString(e));
}

#define CNVB_CUDA(ctx, call, what)                                        \
    do {                                                                  \
        cudaError_t e__ = (call);                                         \
        if (e__ != cudaSuccess) return cnvb::fail_cuda((ctx), e__, what); \
    } while (0)

#define CNVB_TRY(expr)            \
    do {                          \
        int rc__ = (expr);        \
        if (rc__ != 0) return rc__; \
    } while (0)

// RAII: brackets a kernel launch with two events when ctx->timing is on
struct KernelSpan {
    cnvb_ctx *ctx;
    cudaEvent_t b = nullptr;
    static cudaEvent_t get_event(cnvb_ctx *c) {
        if (!c->event_pool.empty()) {
            cudaEvent_t e = c->event_pool.back();
            c->event_pool.pop_back();
            return e;
        }
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        return e;
    }
    KernelSpan(cnvb_ctx *c, const char *name) : ctx(c) {
        if (!c->timing) return;
        int id = -1;
        for (size_t i = 0; i < c->span_names.size(); ++i)
            if (c->span_names[i] == name) id = (int)i;
        if (id < 0) {
            c->span_names.push_back(name);
            id = (int)c->span_names.size() - 1;
        }
        cudaEvent_t a = get_event(c);
        b = get_event(c);
        cudaEventRecord(a, c->stream);
        c->spans.push_back({id, a, b});
    }
    ~KernelSpan() {
        if (b) cudaEventRecord(b, ctx->stream);
    }
};

// after a kernel launch: count it and surface launch errors
#define CNVB_LAUNCHED(ctx, name)                                              \
    do {                                                                      \
        (ctx)->launches += 1;                                                 \
        cudaError_t e__ = cudaGetLastError();                                 \
        if (e__ != cudaSuccess) return cnvb::fail_cuda((ctx), e__, "launch of " name); \
    } while (0)

inline int ensure_scratch(cnvb_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->scratch_bytes) return 0;
    if (ctx->scratch) {
        CNVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "sync before scratch regrow");
        cudaFree(ctx->scratch);
        ctx->scratch = nullptr;
        ctx->scratch_bytes = 0;
    }
    size_t want = bytes + (bytes >> 2) + 4096;
    CNVB_CUDA(ctx, cudaMalloc(&ctx->scratch, want), "allocating device scratch");
    ctx->scratch_bytes = want;
    return 0;
}

inline int ensure_stage(cnvb_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->stage_bytes) return 0;
    if (ctx->stage) {
        CNVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "sync before staging regrow");
        cudaFree(ctx->stage);
        ctx->stage = nullptr;
        ctx->stage_bytes = 0;
    }
    size_t want = bytes + 4096;
    CNVB_CUDA(ctx, cudaMalloc(&ctx->stage, want), "allocating device staging");
    ctx->stage_bytes = want;
    return 0;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Grid size of a persistent (grid-stride) kernel: exactly one wave of resident CTAs, so that no
// partial second wave leaves SMs idle at the end (148 SMs x resident CTAs per SM).
template <typename K>
inline unsigned persistent_grid(cnvb_ctx *ctx, K kernel, int threads, size_t smem, uint64_t work_blocks) {
    const void *key = reinterpret_cast<const void *>(kernel);
    auto it = ctx->occupancy.find(key);
    int per_sm;
    if (it == ctx->occupancy.end()) {
        per_sm = 1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1)
            per_sm = 1;
        ctx->occupancy[key] = per_sm;
    } else {
        per_sm = it->second;
    }
    uint64_t grid = (uint64_t)ctx->sm_count * (uint64_t)per_sm;
    if (work_blocks < grid) grid = work_blocks;
    return grid ? (unsigned)grid : 1u;
}

constexpr uint32_t kTallySlots = 4096;

// pooled device allocation (size classes: next power of two, >= 512 B).
// Reuse is stream-ordered: a block is released while the work that uses it may still be in flight on
// the stream current at release time (ctx->stream, which the region loop points at its side streams).
// A block released on the stream that now asks is reused as is; a block released on ANOTHER stream is
// handed out only after the asking stream has been ordered behind everything enqueued on the
// releasing stream so far (event record + wait, no host synchronisation).
// (a block tagged kPoolQuiescent was released before the last device-wide synchronisation: no ordering needed)
static const cudaStream_t kPoolQuiescent = reinterpret_cast<cudaStream_t>(~(uintptr_t)0);
inline cudaError_t pool_alloc(cnvb_ctx *ctx, size_t bytes, void **out) {
    size_t cls = 512;
    while (cls < bytes) cls <<= 1;
    auto it = ctx->pool_free.find(cls);
    if (it != ctx->pool_free.end() && !it->second.empty()) {
        auto &v = it->second;
        size_t pick = v.size() - 1;
        for (size_t i = v.size(); i-- > 0;)
            if (v[i].released_on == ctx->stream) {
                pick = i;
                break;
            }
        const cnvb_ctx::pool_block b = v[pick];
        if (b.released_on != ctx->stream && b.released_on != kPoolQuiescent) {
            cudaError_t e = cudaSuccess;
            if (!ctx->pool_event) e = cudaEventCreateWithFlags(&ctx->pool_event, cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventRecord(ctx->pool_event, b.released_on);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, ctx->pool_event, 0);
            if (e != cudaSuccess) return e;
        }
        v.erase(v.begin() + (long)pick);
        *out = b.p;
    } else {
        cudaError_t e = cudaMalloc(out, cls);
        if (e != cudaSuccess) return e;
    }
    ctx->pool_live[*out] = cls;
    return cudaSuccess;
}
template <typename T>
inline cudaError_t pool_alloc_t(cnvb_ctx *ctx, size_t count, T **out) {
    void *p = nullptr;
    cudaError_t e = pool_alloc(ctx, (count ? count : 1) * sizeof(T), &p);
    *out = static_cast<T *>(p);
    return e;
}
// Returns the block to the pool, tagged with the stream its last use was enqueued on (= ctx->stream
// at the time of the call: release a block BEFORE switching ctx->stream away from the stream that used it).
inline void pool_release(cnvb_ctx *ctx, void *p) {
    if (!p) return;
    auto it = ctx->pool_live.find(p);
    if (it == ctx->pool_live.end()) return;
    if (ctx->capturing) {  // the captured work uses the block at every replay: it stays out of circulation
        ctx->capture_hold.push_back(p);
        return;
    }
    ctx->pool_free[it->second].push_back({p, ctx->stream});
    ctx->pool_live.erase(it);
}
// after a device-wide synchronisation: every free block may go to any stream without an ordering event
inline void pool_quiesce(cnvb_ctx *ctx) {
    for (auto &kv : ctx->pool_free)
        for (auto &b : kv.second) b.released_on = kPoolQuiescent;
}
// pinned host memory that lives as long as the plan under capture (nullptr when none is being captured)
inline void *capture_pin(cnvb_ctx *ctx, const void *src, size_t bytes) {
    if (!ctx->capturing) return nullptr;
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    if (src) memcpy(p, src, bytes); else memset(p, 0, bytes);
    ctx->capture_pinned.push_back(p);
    return p;
}
inline void pool_destroy(cnvb_ctx *ctx) {
    for (auto &kv : ctx->pool_free)
        for (auto &b : kv.second) cudaFree(b.p);
    for (auto &kv : ctx->pool_live) cudaFree(kv.first);
    ctx->pool_free.clear();
    ctx->pool_live.clear();
    if (ctx->pool_event) cudaEventDestroy(ctx->pool_event);
    ctx->pool_event = nullptr;
}

// Carves aligned sub-buffers out of one allocation.
struct Carver {
    char *base;
    size_t off = 0;
    explicit Carver(void *b) : base(static_cast<char *>(b)) {}
    template <typename T>
    T *take(size_t count) {
        off = align_up(off, 256);
        T *p = reinterpret_cast<T *>(base + off);
        off += count * sizeof(T);
        return p;
    }
    static size_t need(std::initializer_list<size_t> sizes) {
        size_t o = 0;
        for (size_t s : sizes) o = align_up(o, 256) + s;
        return o + 256;
    }
};

// ---- device helpers ---------------------------------------------------------------------

__device__ __forceinline__ float f32_missing() { return __uint_as_float(CNVB_F32_MISSING_BITS); }

// streaming 128/64/32-bit loads that do not pollute L1 (data is touched exactly once)
__device__ __forceinline__ int4 ld_stream(const int4 *p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ld_stream(const uint2 *p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];"
                 : "=r"(r.x), "=r"(r.y)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ld_stream(const uint32_t *p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

// Exact u32 division by a runtime-invariant divisor: q = n / d with m = min(2^32/d, 2^32-1).
// floor(n*m / 2^32) is q or q-1, so one correction step suffices.
struct FastDiv {
    uint32_t d, m;
};
inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    f.d = d;
    uint64_t m = (1ull << 32) / d;
    f.m = m > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)m;
    return f;
}
__device__ __forceinline__ uint32_t fastdiv(uint32_t n, FastDiv f) {
    uint32_t q = __umulhi(n, f.m);
    uint32_t r = n - q * f.d;
    return q + (r >= f.d ? 1u : 0u);
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

}  // namespace cnvb
