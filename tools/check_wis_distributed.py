"""Ad-hoc multi-GPU check (run under torchrun, NCCL): build-model-wis with targets sharded by row
block must equal the single-GPU model bit for bit.  Not a pytest (needs >= 2 GPUs)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cnvetti_b200 as cb  # noqa: E402
from cnvetti_b200 import model_wis as mw, parallel, synth  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = cb.Context(local)
    T, S = 40_000, 100
    X, tids = synth.synth_panel(3, T, S)
    opts = mw.BuildModelWisOptions()
    lo, hi = parallel.shard_rows(T, rank, world)
    dev = "cuda:%d" % local
    Xb = torch.from_numpy(X[lo:hi]).to(dev)
    tb = torch.from_numpy(tids[lo:hi].view(np.int32)).to(dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    parallel.wis_run_distributed(Xb, tb, opts, backend=parallel.CudaWisBackend(ctx))
    dist.barrier()
    torch.cuda.synchronize()
    ev0.record()
    m = parallel.wis_run_distributed(Xb, tb, opts, backend=parallel.CudaWisBackend(ctx))
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    full = mw.run(torch.from_numpy(X).to(dev), torch.from_numpy(tids.view(np.int32)).to(dev), opts, ctx=ctx)
    ok = (torch.equal(m["ref_idx"], full.ref_idx[lo:hi]) and torch.equal(m["ref_len"], full.ref_len[lo:hi])
          and torch.equal(m["num_calls"], full.num_calls[lo:hi]) and torch.equal(m["filt"], full.filt[lo:hi])
          and torch.equal(m["ref_dist"].view(torch.int32), full.ref_dist[lo:hi].view(torch.int32))
          and np.float32(m["threshold"]).tobytes() == np.float32(full.threshold).tobytes())
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("wis distributed world=%d rows/rank=%d identical=%s ms=%.2f" % (world, hi - lo, bool(flag.item()), ms))
    dist.destroy_process_group()
    return 0 if flag.item() else 1


if __name__ == "__main__":
    sys.exit(main())
