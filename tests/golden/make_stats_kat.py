#!/usr/bin/env python3
"""Extract the reference's own known-answer tests into JSON fixtures.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):

    python tests/golden/make_stats_kat.py

Sources:
  * lib-shared/src/stats.rs:377-907  -- 17 `Stats` KATs (vectors generated from R upstream)
  * lib-coverage/src/lib.rs:797-807  -- `test_compute_threshold` (pile depth threshold == 7)

Only the test DATA (input vectors and expected values) is extracted; no code is copied.
"""
import json
import os
import re

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def parse_stats_kats(path):
    src = open(path).read()
    tests_src = src[src.index("mod tests"):]
    out = []
    num = r"[-+]?(?:\d+\.\d*(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+|\d+)"
    for m in re.finditer(r"fn (test_\w+)\(\) \{(.*?)\n    \}", tests_src, re.S):
        name, body = m.group(1), m.group(2)
        vm = re.search(r"let val = &\[(.*?)\];", body, re.S)
        sm = re.search(r"let summ = &Summary \{(.*?)\};", body, re.S)
        if not (vm and sm):
            continue
        vals = [float(x) for x in re.findall(num, vm.group(1))]
        summ = {}
        for fm in re.finditer(r"(\w+): (\(.*?\)|" + num + r"),", sm.group(1), re.S):
            key, v = fm.group(1), fm.group(2)
            if v.startswith("("):
                summ[key] = [float(x) for x in re.findall(num, v)]
            else:
                summ[key] = float(v)
        out.append({"name": name, "val": vals, "summary": summ})
    return out


def main():
    stats_rs = os.path.join(REF, "lib-shared/src/stats.rs")
    kats = parse_stats_kats(stats_rs)
    fixture = {
        "source": "lib-shared/src/stats.rs:377-907",
        "exact_fields": ["sum", "min", "max", "mean", "median", "quartiles", "iqr"],
        "approx_fields": ["var", "std_dev", "std_dev_pct", "median_abs_dev", "median_abs_dev_pct"],
        "approx_abs_tol": 1.0e-6,
        "summaries": kats,
        # stats.rs:378-383, 901-907 (asserted exactly in the reference)
        "min_max_nan": {"val": [1.0, 2.0, "nan", 3.0, 4.0], "min": 1.0, "max": 4.0},
        "sum_cases": [
            {"val": [0.5, 3.2321, 1.5678], "sum": 5.2999},
            {"val": [1e30, 1.2, -1e30], "sum": 1.2},
        ],
        # lib-coverage/src/lib.rs:797-807
        "pile_threshold": {"y": [0, 10, 8, 7, 6, 5, 3], "fdr": 0.001, "expected": 7},
    }
    with open(os.path.join(HERE, "stats_kat.json"), "w") as f:
        json.dump(fixture, f, indent=1)
    print("wrote %d summary KATs" % len(kats))


if __name__ == "__main__":
    main()
