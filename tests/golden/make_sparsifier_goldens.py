#!/usr/bin/env python3
"""Writes tests/golden/sparsifier_goldens.json: (CoB, Res) of the whole sparsifier pipeline computed by the INDEPENDENT Python
restatement tests/py_sparsifier.py (fractions / integers mod p; shares no code with the C++ oracle or the product) on matrices of
the reference's data/ directory (fixture tests/golden/all_matrices.json) in the configurations of bin/FDT.sh:64-66 (`-c 5` over Q,
`-q 7 -c 5`) and of BASELINE configs 1 and 3 (`-c 4` on 2x2x2_7_DPS-smallrat-12.2034_L; `-c 11` on 4x4x4_48_rational_L over Q and
mod 2^31-1).   python tests/golden/make_sparsifier_goldens.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib as O  # noqa: E402  (only for the matrix fixtures)
import py_sparsifier as PS  # noqa: E402

CASES = [(name, q, 5) for name in ("2x2x2_7_DPS-smallrat-12.2034_L", "2x2x2_7_Winograd_R", "2x2x2_7_DPS-accurate_P", "3x3x3_23_58_L", "3x3x6_40_R",
                                   "4x4x4_48_rational_L", "4x4x4_48_rational_R", "4x4x4_48_accurate_L", "3x4x7_63_rational_L", "3x4x7_63_rational_R",
                                   "3x4x7_63_rational_P", "4x4x4_49_156_P", "6x3x3_40_L") for q in (0, 7)]
CASES += [("2x2x2_7_DPS-smallrat-12.2034_L", 0, 4), ("4x4x4_48_rational_L", 0, 11), ("4x4x4_48_rational_L", 2147483647, 11), ("3x4x7_63_rational_R", 2147483647, 11)]


def main():
    out = []
    for name, q, c in CASES:
        M = O.dense_fractions(name)
        F = PS.QQ() if q == 0 else PS.Zp(q)
        CoB, Res = PS.block_sparsifier(F, M, 4, c, True)
        out.append({"matrix": name, "q": q, "c": c, "CoB": [[str(v) for v in r] for r in CoB], "Res": [[str(v) for v in r] for r in Res]})
        print(name, q, c, "nnz(Res) =", sum(1 for r in Res for v in r if v != 0))
    with open(os.path.join(HERE, "sparsifier_goldens.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))


if __name__ == "__main__":
    main()
