#!/usr/bin/env python3
"""Where the step time of the contig-sharded headline workload goes at N ranks (run under torchrun):
the whole step / the region loop alone / the normaliser alone, each with and without the clock sampler,
plus the host time of one step's enqueue (wall clock until the calls return)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    sys.argv = [sys.argv[0], "--gpus", os.environ.get("WORLD_SIZE", "1"), "--workload", "coverage", "--quick"]
    args = bench.parse_args()
    env = bench.Env(args)
    torch = env.torch
    import cnvetti_b200 as cb
    from cnvetti_b200 import pipeline, synth_torch
    regs = bench.contigs(1.0)
    n_pairs = [max(1000, int(round(L * bench.READS_PER_BP / 2))) for _, L in regs]
    opts = cb.CoverageOptions(window_length=bench.WINDOW, min_mapq=bench.MIN_MAPQ)
    owner, per_rank = pipeline.shard_regions([L for _, L in regs], env.world)
    out = []
    for ci in per_rank[env.rank]:
        name, L = regs[ci]
        soa = synth_torch.synth_soa_contig(2000 + ci, L, n_pairs[ci], env.dev)
        seq = synth_torch.synth_reference_torch(2000 + ci, L, env.dev)
        out.append(pipeline.Region(name, L, cb.ReadBatch(mid=soa["mid"].contiguous(), mapq=soa["mapq"].contiguous(),
                                                        flags=soa["flags"].contiguous()), seq))
    prep = pipeline.PreparedRegions(out, opts)
    shard = env.comm if env.world > 1 else None
    table = pipeline.coverage_run(prep, opts, ctx=env.ctx)

    def both():
        t = pipeline.coverage_run(prep, opts, ctx=env.ctx)
        pipeline.normalize_run(t, "MedianGcBinned", ctx=env.ctx, comm=shard, want_medians=False)

    def cov():
        pipeline.coverage_run(prep, opts, ctx=env.ctx)

    flight = []

    def both_deferred():
        t = pipeline.coverage_run(prep, opts, ctx=env.ctx, defer=True)
        pipeline.normalize_run(t, "MedianGcBinned", ctx=env.ctx, comm=shard, want_medians=False)
        flight.append(t)
        if len(flight) > 1:
            flight.pop(0).wait()

    def flush():
        while flight:
            flight.pop(0).wait()
    both_deferred.flush = flush

    def norm():
        pipeline.normalize_run(table, "MedianGcBinned", ctx=env.ctx, comm=shard, want_medians=False)

    res = {"world": env.world}
    for name, fn in (("both", both), ("both_deferred", both_deferred), ("cov", cov), ("norm", norm)):
        for with_sampler in (False, True):
            s = bench.ClockSampler(env.local) if with_sampler else None
            ms, _, _ = bench.timed_steps(env, fn, 50, 5, s)
            res["%s%s" % (name, "+sampler" if with_sampler else "")] = round(ms, 4)
        # host time of the enqueue alone
        torch.cuda.synchronize()
        env.barrier()
        t0 = time.perf_counter()
        for _ in range(20):
            fn()
        if hasattr(fn, "flush"):
            fn.flush()
        host = (time.perf_counter() - t0) / 20
        torch.cuda.synchronize()
        res[name + "_host_ms"] = round(host * 1e3, 4)
    if env.rank == 0:
        print(json.dumps(res))
    env.close()


if __name__ == "__main__":
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, "w", buffering=1)
    main()
