"""Generates tests/golden/path_vectors.npz: small seeded inputs for every row of the hot path together
with the outputs of the CPU restatement (oracle/), so that both the oracle (CPU suite) and the CUDA
path (GPU suite) are held against COMMITTED vectors.  The reference itself cannot be run here (Rust,
no toolchain), so these vectors pin the restatement against drift; the reference-derived known-answer
tests are in stats_kat.json.

    python tests/golden/make_path_vectors.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from cnvetti_b200 import synth  # noqa: E402


def main():
    oracle.build()
    out = {}
    L, wl = 60_000, 500
    d = synth.synth_bam_records(7, L, 1500, unpaired_frac=0.0)
    L = max(L, int(d["pos"].max()) + 2000)
    for k in ("pos", "isize", "flag", "mapq", "cigar_off", "cigar"):
        out["reads_" + k] = d[k]
    out["L"] = np.int64(L)
    out["wl"] = np.int64(wl)
    rb = oracle.ReadBatch(d["pos"], d["isize"], d["flag"], d["mapq"], d["cigar_off"], d["cigar"])
    oo = oracle.cov_opts(window_length=wl, min_mapq=20)
    counts, proc, skip = oracle.frag_gw(rb, oo, L)                       # A2
    out["gw_counts"], out["gw_tally"] = counts, np.array([proc, skip], np.int64)
    ts, te = synth.synth_targets(8, L, 60)                               # A3
    tc, tp, tsk = oracle.frag_targets(rb, oracle.cov_opts(window_length=0, min_mapq=20), ts, te)
    out["tgt_start"], out["tgt_end"], out["tgt_counts"], out["tgt_tally"] = ts, te, tc, np.array([tp, tsk], np.int64)
    mean, sd = oracle.coverage(rb, oo, L)                                # A4
    out["cov_mean"], out["cov_sd"] = mean, sd
    seq = synth.synth_reference(9, L)                                    # A5
    gc, gap = oracle.refstats(seq, wl)
    out["seq"], out["gc"], out["gap"] = seq, gc, gap
    cov, start, end, frm = oracle.frag_gw_stats(counts, oo, L)           # A6
    lcv, _, ft = oracle.cov_records(cov, None, start, end, frm, False, oo)
    out["lcv"], out["ft"] = lcv, ft
    s, cv, _ = oracle.norm_total(lcv)                                    # N1
    out["n1_sum"], out["n1_cv"] = np.float64(s), cv
    rng = np.random.default_rng(10)                                      # N2 (enough windows per GC bin)
    n2_gc = np.clip(rng.normal(0.42, 0.08, 6000), 0, 1).astype(np.float32)
    n2_gc[::97] = np.frombuffer(np.uint32(0x7F800001).tobytes(), np.float32)[0]
    n2_lcv = (rng.poisson(60 * (0.5 + np.nan_to_num(n2_gc.astype(np.float64)).clip(0, 1))) / 500.0).astype(np.float32)
    cv2_cv, cv2, _, flags, med = oracle.norm_gc_median(n2_lcv, n2_gc)
    out["n2_lcv"], out["n2_gc"], out["n2_cv"], out["n2_cv2"], out["n2_flags"], out["n2_med"] = n2_lcv, n2_gc, cv2_cv, cv2, flags, med
    X, tids = synth.synth_panel(11, 600, 24, n_contigs=5)                # W2-W5
    dist, idx = oracle.wis_distances(X, tids, max_ref_targets=30)
    out["wis_X"], out["wis_tid"], out["wis_dist"], out["wis_idx"] = X, tids, dist, idx
    thr, pidx, plen = oracle.wis_prune(dist, idx, 5.64)
    calls = oracle.wis_calls(X, pidx, plen, 10, 5.64, 0.35)
    filt = oracle.wis_filters(plen, calls, 10, 4)
    out["wis_thr"], out["wis_ref_idx"], out["wis_ref_len"], out["wis_calls"], out["wis_filt"] = np.float32(thr), pidx, plen, calls, filt
    mc = oracle.modcov(X[:, 0].copy(), (filt != 0).astype(np.uint8), pidx, plen)     # M1
    for k, v in mc.items():
        out["modcov_" + k] = v
    cvl, _ = synth.synth_cv_lanes(12, 5000, n_cnv=6)                     # H
    t = oracle.hmm_tables(0.08)
    out["hmm_cv"], out["hmm_state"] = cvl, oracle.hmm_viterbi(cvl, t)
    ps, pe, psz = oracle.piles_from_depths(oracle.coverage(rb, oracle.cov_opts(window_length=64, min_mapq=20), L, debug=True)[3][:L], 20)
    out["pile_start"], out["pile_end"], out["pile_size"] = ps, pe, psz
    path = os.path.join(ROOT, "tests", "golden", "path_vectors.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    main()
