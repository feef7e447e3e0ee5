// wis_tensor.cuh -- placeholder, replaced by the tcgen05 engine
#pragma once
namespace {
int wis_distances_tensor(cnvb_ctx *ctx, const float *, const uint32_t *, uint32_t, uint32_t, uint32_t, uint32_t,
                         uint32_t, float *, uint32_t *) {
    return fail(ctx, CNVB_ERR_INVALID, "tensor engine not built");
}
}  // namespace
