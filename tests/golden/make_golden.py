"""Generates tests/golden/ref_outputs.json by running the UNMODIFIED reference headers
(oracle/_ref/libradix_ref.so, built from /root/reference by oracle/Makefile) on the seeded
inputs of tests/cases.py.  Run in the build container (the reference tree is absent on the
GPU box):

    python tests/golden/make_golden.py

For every case it stores the SHA-256 of the reference's output bytes and which buffer the
reference returned; small cases also keep the first/last elements for debugging.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402

import pyoracle  # noqa: E402
from cases import GOLDEN_CASES, case_id, digest, make_input  # noqa: E402


def main():
    ref = pyoracle.Ref()
    out = {"_generator": "tests/golden/make_golden.py", "_source": "reference radix_sort.hpp / radix_sort_rank.hpp via oracle/ref_shim.cpp",
           "sort": {}, "sort_desc": {}, "rank_as_shipped": {}}
    for c in GOLDEN_CASES:
        t = pyoracle.TYPES[c[0]]
        data = make_input(c[0], c[1], 1234, c[2], c[3], c[4])
        res, in_aux = ref.radix_sort(t, data)
        out["sort"][case_id(c)] = {"sha256": digest(res), "result_in_aux": int(in_aux)}
        if c[1] in (257, 5000, 70001):
            res, in_aux = ref.radix_sort(t, data, descending=True)
            out["sort_desc"][case_id(c)] = {"sha256": digest(res), "result_in_aux": int(in_aux)}
        # The rank header as shipped (radix_sort_rank.hpp:82) -- recorded so that the tests can
        # show where it agrees with the intended semantics (<= 1 live column, early exit).
        if c[1] in (257, 1000, 5000):
            ranks, in_aux, _ = ref.radix_sort_rank(t, data, np.uint32)
            out["rank_as_shipped"][case_id(c)] = {"sha256": digest(ranks), "result_in_aux": int(in_aux)}
    with open(os.path.join(HERE, "ref_outputs.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("cases:", len(out["sort"]), len(out["sort_desc"]), len(out["rank_as_shipped"]))


if __name__ == "__main__":
    main()
