// ctx.cu -- context lifecycle and the host-side read packer.
#include "common.cuh"

static thread_local std::string g_create_error;

extern "C" const char *cnvb_version(void) { return "cnvetti_b200 0.1.0 (sm_100a)"; }

extern "C" const char *cnvb_create_error(void) { return g_create_error.c_str(); }

extern "C" int cnvb_ctx_create(int device, void *stream, cnvb_ctx **out) {
    if (!out) return CNVB_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_create_error = std::string("no CUDA device available (there is no CPU fallback)") +
                         "\ncaused by: " + cudaGetErrorString(e);
        return CNVB_ERR_CUDA;
    }
    if (device < 0 || device >= count) {
        g_create_error = "device index out of range";
        return CNVB_ERR_INVALID;
    }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        g_create_error = std::string("cudaSetDevice failed\ncaused by: ") + cudaGetErrorString(e);
        return CNVB_ERR_CUDA;
    }
    cnvb_ctx *ctx = new cnvb_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    if (stream) {
        ctx->stream = static_cast<cudaStream_t>(stream);
        ctx->own_stream = false;
    } else {
        e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            g_create_error = std::string("cudaStreamCreate failed\ncaused by: ") + cudaGetErrorString(e);
            delete ctx;
            return CNVB_ERR_CUDA;
        }
        ctx->own_stream = true;
    }
    e = cudaHostAlloc(reinterpret_cast<void **>(&ctx->tally_slots), cnvb::kTallySlots * 16, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        g_create_error = std::string("cudaHostAlloc failed\ncaused by: ") + cudaGetErrorString(e);
        if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
        delete ctx;
        return CNVB_ERR_CUDA;
    }
    *out = ctx;
    return CNVB_OK;
}

extern "C" void cnvb_ctx_destroy(cnvb_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto &s : ctx->spans) {
        cudaEventDestroy(s.a);
        cudaEventDestroy(s.b);
    }
    for (auto e : ctx->event_pool) cudaEventDestroy(e);
    cnvb::pool_destroy(ctx);
    if (ctx->tally_slots) cudaFreeHost(ctx->tally_slots);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->ref_lut) cudaFree(ctx->ref_lut);
    for (int i = 0; i < 3; ++i) {
        if (ctx->side[i]) cudaStreamDestroy(ctx->side[i]);
        if (ctx->side_join[i]) cudaEventDestroy(ctx->side_join[i]);
    }
    if (ctx->side_fork) cudaEventDestroy(ctx->side_fork);
    if (ctx->stage) cudaFree(ctx->stage);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int cnvb_ctx_synchronize(cnvb_ctx *ctx) {
    if (!ctx) return CNVB_ERR_INVALID;
    CNVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "synchronising the context stream");
    return CNVB_OK;
}

extern "C" int cnvb_ctx_release_memory(cnvb_ctx *ctx) {
    if (!ctx) return CNVB_ERR_INVALID;
    CNVB_CUDA(ctx, cudaSetDevice(ctx->device), "selecting device");
    CNVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "synchronising the context stream");
    for (int q = 0; q < 3; ++q)
        if (ctx->side[q]) CNVB_CUDA(ctx, cudaStreamSynchronize(ctx->side[q]), "synchronising the side streams");
    for (auto &kv : ctx->pool_free)
        for (auto &b : kv.second) cudaFree(b.p);
    ctx->pool_free.clear();
    if (ctx->scratch) cudaFree(ctx->scratch);
    ctx->scratch = nullptr;
    ctx->scratch_bytes = 0;
    if (ctx->stage) cudaFree(ctx->stage);
    ctx->stage = nullptr;
    ctx->stage_bytes = 0;
    return CNVB_OK;
}

extern "C" const char *cnvb_last_error(const cnvb_ctx *ctx) { return ctx ? ctx->err.c_str() : ""; }

extern "C" uint64_t cnvb_ctx_launch_count(const cnvb_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int cnvb_ctx_set_timing(cnvb_ctx *ctx, int on) {
    if (!ctx) return CNVB_ERR_INVALID;
    ctx->timing = on != 0;
    return CNVB_OK;
}

extern "C" int cnvb_ctx_timing_reset(cnvb_ctx *ctx) {
    if (!ctx) return CNVB_ERR_INVALID;
    CNVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "synchronising before timing reset");
    for (auto &s : ctx->spans) {
        ctx->event_pool.push_back(s.a);
        ctx->event_pool.push_back(s.b);
    }
    ctx->spans.clear();
    return CNVB_OK;
}

extern "C" int cnvb_ctx_timing_query(cnvb_ctx *ctx, const char *kernel, double *total_ms, uint64_t *count) {
    if (!ctx || !kernel) return CNVB_ERR_INVALID;
    CNVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "synchronising before timing query");
    double tot = 0.0;
    uint64_t cnt = 0;
    for (auto &s : ctx->spans) {
        if (ctx->span_names[s.name_id] != kernel) continue;
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) {
            tot += ms;
            cnt += 1;
        }
    }
    if (total_ms) *total_ms = tot;
    if (count) *count = cnt;
    return CNVB_OK;
}

// ---- host packer ---------------------------------------------------------------------------
// What the reference derives per bam::Record, restated once for the SoA the kernels consume:
//   end      = htslib bam_endpos()                       (pileup interval, agg.rs:524)
//   mid      = GenomeWide centre: pos + isize/2 (paired, i32 trunc) | pos + end_pos (unpaired,
//              sic)                                       (agg.rs:124-133)
//   frag_end = TargetRegions end: pos + isize (paired) | end_pos (unpaired)  (agg.rs:312-320)
//   flags    = FLAG | CLIPFAIL (agg.rs:174-193, f32 ratio) | ISIZE_NONPOS (agg.rs:196-198)
extern "C" int cnvb_pack_reads(uint64_t n, const int32_t *pos, const int32_t *isize,
                               const uint16_t *flag, const uint64_t *cigar_off,
                               const uint32_t *cigar, float min_unclipped, int32_t *out_end,
                               int32_t *out_mid, int32_t *out_frag_end, uint16_t *out_flags) {
    if (n && (!pos || !isize || !flag || !cigar_off)) return CNVB_ERR_INVALID;
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t ref_len = 0, clipped = 0, unclipped = 0;
        const uint64_t c0 = cigar_off[i], c1 = cigar_off[i + 1];
        for (uint64_t c = c0; c < c1; ++c) {
            const uint32_t op = cigar[c] & 0xFu, len = cigar[c] >> 4;
            switch (op) {
            case 0: case 7: case 8: ref_len += len; unclipped += len; break;  // M = X
            case 1: unclipped += len; break;                                  // I
            case 2: case 3: ref_len += len; break;                            // D N
            case 4: case 5: clipped += len; break;                            // S H
            default: break;                                                   // P
            }
        }
        const int32_t end_pos = pos[i] + (int32_t)ref_len;  // CigarStringView::end_pos()
        const uint16_t f = flag[i];
        if (out_end) {
            uint32_t rlen = 1;
            if (!(f & 0x4) && c1 > c0) rlen = ref_len ? ref_len : 1;
            out_end[i] = pos[i] + (int32_t)rlen;
        }
        const bool paired = f & 0x1;
        if (out_mid) out_mid[i] = paired ? pos[i] + isize[i] / 2 : pos[i] + end_pos;
        if (out_frag_end) out_frag_end[i] = paired ? pos[i] + isize[i] : end_pos;
        if (out_flags) {
            uint16_t x = f & 0x0FFF;
            const float ratio = (float)unclipped / (float)(clipped + unclipped);
            if (ratio < min_unclipped) x |= CNVB_FLAG_CLIPFAIL;
            if (isize[i] <= 0) x |= CNVB_FLAG_ISIZE_NONPOS;
            out_flags[i] = x;
        }
    }
    return CNVB_OK;
}
