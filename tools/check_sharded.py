"""Multi-GPU parity check (run under torchrun, >= 2 GPUs; NCCL inside the library through cnvb_comm):

  1. one genome, contigs sharded over the ranks (pipeline.shard_regions / coverage_run_sharded) +
     MedianGcBinned and TotalCovSum with the statistics exchanged inside the library  ==  the single-GPU
     run over all contigs, bit for bit (counts, GC, medians, CV, CV2, flags), also after gather_table;
  2. cnvb_wis_build_sharded (uneven row blocks)  ==  cnvb_wis_build;
  3. the cohort chain (pipeline.wis_call_cohort with comm=)  ==  the single-process chain.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cnvetti_b200 as cb  # noqa: E402
from cnvetti_b200 import model_wis as mw, parallel, pipeline, synth, synth_torch  # noqa: E402


def bits(x):
    return x.contiguous().view(torch.int32) if x.dtype == torch.float32 else x


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    dist.init_process_group("nccl", device_id=torch.device(dev))
    ctx = cb.Context(local)
    comm = parallel.Communicator(ctx, rank, world)
    ok = True

    def check(name, cond):
        nonlocal ok
        flag = torch.tensor([1 if cond else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print("%-58s %s" % (name, "identical" if flag.item() else "DIFFERENT"))
        ok = ok and bool(flag.item())

    # ---- 1. one genome on N GPUs -------------------------------------------------------------------------
    lens = [3_100_000, 1_200_000, 2_400_017, 900_000, 1_800_500, 650_000, 2_000_000]
    opts = cb.CoverageOptions(window_length=500, min_mapq=20)

    def region(i):
        L = lens[i]
        soa = synth_torch.synth_soa_contig(100 + i, L, L // 20, dev)
        seq = synth_torch.synth_reference_torch(100 + i, L, dev)
        return pipeline.Region("c%d" % i, L, cb.ReadBatch(mid=soa["mid"].contiguous(), mapq=soa["mapq"].contiguous(),
                                                         flags=soa["flags"].contiguous()), seq)

    owner, per_rank = pipeline.shard_regions(lens, world)
    t = pipeline.coverage_run_sharded([region(i) for i in per_rank[rank]], opts, comm)
    pipeline.normalize_run(t, "MedianGcBinned", comm=comm)
    full_sh = pipeline.gather_table(t, owner, comm)
    one = pipeline.normalize_run(pipeline.coverage_run([region(i) for i in range(len(lens))], opts, ctx=ctx),
                                 "MedianGcBinned", ctx=ctx)
    check("coverage sharded by contig + MedianGcBinned: medians", np.array_equal(t.medians.view(np.uint32), one.medians.view(np.uint32)))
    for col in ("start", "end", "rcv", "lcv", "gc", "gap", "ft", "cv", "cv2", "norm_flags"):
        check("  gathered column %s" % col, torch.equal(bits(getattr(full_sh, col)), bits(getattr(one, col))))
    check("  gathered offsets", full_sh.offsets == one.offsets)
    t2 = pipeline.normalize_run(t, "TotalCovSum", comm=comm)
    cv_sh = comm.all_gather_rows(t2.cv)
    one2 = pipeline.normalize_run(one, "TotalCovSum", ctx=ctx)
    order = torch.cat([one2.cv[one.offsets[i]:one.offsets[i + 1]] for q in range(world) for i in per_rank[q]])
    check("TotalCovSum sharded (double-double pairs exchanged)", torch.equal(bits(cv_sh), bits(order)))

    # ---- 2. build-model-wis, uneven row blocks ---------------------------------------------------------
    T, S = 30_000, 100
    X, tids = synth.synth_panel(3, T, S)
    cuts = [0] + [T * (q + 1) // world + (53 if q % 2 == 0 and q + 1 < world else 0) for q in range(world)]
    lo, hi = cuts[rank], cuts[rank + 1]
    o = mw.BuildModelWisOptions(engine=mw.ENGINE_TENSOR)
    m, (rb, re_), Tall = mw.run_sharded(torch.from_numpy(X[lo:hi]).to(dev), torch.from_numpy(tids[lo:hi].view(np.int32)).to(dev), o, comm)
    full = mw.run(torch.from_numpy(X).to(dev), torch.from_numpy(tids.view(np.int32)).to(dev), o, ctx=ctx)
    same = (rb, re_, Tall) == (lo, hi, T) and np.float32(m.threshold).tobytes() == np.float32(full.threshold).tobytes()
    same = same and torch.equal(bits(m.ref_dist), bits(full.ref_dist[lo:hi])) and torch.equal(m.ref_len, full.ref_len[lo:hi])
    same = same and torch.equal(m.num_calls, full.num_calls[lo:hi]) and torch.equal(m.filt, full.filt[lo:hi])
    keep = torch.arange(m.ref_idx.shape[1], device=dev)[None, :] < m.ref_len[:, None]
    same = same and torch.equal(m.ref_idx[keep], full.ref_idx[lo:hi][keep])
    check("cnvb_wis_build_sharded (uneven row blocks) vs cnvb_wis_build", same)

    # ---- 3. the cohort chain -------------------------------------------------------------------------------
    Tc, Sc = 40_000, 160
    ranges = [parallel.shard_rows(Sc, r, world) for r in range(world)]
    a, b = ranges[rank]
    lcv, tid_c = synth_torch.synth_cohort_lcv(5, Tc, a, b, dev)
    r = pipeline.wis_call_cohort(lcv, tid_c, sample_ranges=ranges, sigma=0.08, comm=comm)
    lcv_all, _ = synth_torch.synth_cohort_lcv(5, Tc, 0, Sc, dev)
    r1 = pipeline.wis_call_cohort(lcv_all, tid_c, sigma=0.08, ctx=ctx)
    rl, rh = r["rows"]
    same = torch.equal(r["states"], r1["states"][a:b]) and torch.equal(bits(r["cv_rel"]), bits(r1["cv_rel"][a:b]))
    same = same and torch.equal(bits(r["model"]["ref_dist"]), bits(r1["model"]["ref_dist"][rl:rh]))
    same = same and torch.equal(r["model"]["filt"], r1["model"]["filt"][rl:rh])
    check("cohort chain (normalise, merge-cov, model, mod-coverage, HMM)", same)
    if rank == 0:
        print("check_sharded world=%d: %s" % (world, "ALL IDENTICAL" if ok else "FAILED"))
    comm.close()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
