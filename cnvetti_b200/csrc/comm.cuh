// comm.cuh -- the one-process-per-GPU communicator behind the C ABI (cnvb_comm): NCCL over
// NVLink / NVSwitch, resolved at run time (dlopen "libnccl.so.2": inside a torch process that is the
// NCCL torch itself loaded; a Rust / C host gets the system library).  All operations are enqueued on
// the context's stream; buffers are device memory.
#pragma once

#include "common.cuh"

struct cnvb_comm {
    cnvb_ctx *ctx = nullptr;
    void *nccl = nullptr;  // ncclComm_t
    int rank = 0, world = 1;
    bool own = false;      // destroy the NCCL communicator with the handle
};

namespace cnvb {

// rank / world of a nullable communicator (NULL = single process)
inline int comm_rank(const cnvb_comm *c) { return c ? c->rank : 0; }
inline int comm_world(const cnvb_comm *c) { return c ? c->world : 1; }

// in-place sum over ranks (count elements of u32 / f64)
int comm_all_reduce_u32(cnvb_comm *c, uint32_t *buf, size_t count);
int comm_all_reduce_f64(cnvb_comm *c, double *buf, size_t count);
// rank r contributes bytes[r] bytes; recv holds the contributions back to back in rank order
// (send may alias its own slot of recv).  One grouped launch of per-rank broadcasts, so blocks
// may differ in size.
int comm_all_gather_v(cnvb_comm *c, const void *send, void *recv, const uint64_t *bytes);
// send_off / recv_off: world + 1 byte offsets into send / recv; block q of send goes to rank q
int comm_all_to_all_v(cnvb_comm *c, const void *send, const uint64_t *send_off, void *recv, const uint64_t *recv_off);

}  // namespace cnvb
