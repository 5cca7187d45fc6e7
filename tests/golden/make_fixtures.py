#!/usr/bin/env python3
"""Generate tests/golden/hm_matrices.json from the reference's data/ directory.

Run ONCE in the build container (where /root/reference is mounted); the JSON it
writes is committed and is what tests / bench / smoke read at run time (the GPU
box has no /root/reference).  Matrices are the mathematical inputs of the hot
path (Hopcroft-Musinski L/R/P triples and ALT/CoB factor pairs), stored as
(rows, cols, [[i, j, "num/den"], ...]) with 0-based indices, entries sorted
row-major (the storage order of LinBox SparseSeq rows).

Source format: data/README.md:10-17 of the reference (SMS: optional '#' lines,
header 'm n R|M', 1-based 'i j value' lines, terminator '0 0 0').
"""
import json, os, sys
from fractions import Fraction

REF = os.environ.get("PLINOPT_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hm_matrices.json")

STEMS = [
    "2x2x2_7_Strassen", "2x2x2_7_Winograd", "2x2x2_7_DPS-accurate",
    "2x2x2_7_DPS-smallrat-12.2034", "2x2x2_7_DPS-evenpow-12.2034",
    "2x2x2_7_DPS-integral-12.0662", "2x2x2_7_DPS-intermediate-12.0695",
    "3x3x3_23_58", "3x3x3_23_Grey-221", "3x3x6_40", "3x6x3_40", "6x3x3_40",
    "4x4x4_48_rational", "4x4x4_48_rational-ALT", "4x4x4_48_rational-CoB",
    "4x4x4_48_accurate", "4x4x4_48_accurate-ALT", "4x4x4_48_accurate-CoB",
    "4x4x4_49_156",
    "3x4x7_63_rational", "3x4x7_63_rational-ALT", "3x4x7_63_rational-CoB",
]
SINGLES = ["cyclic"]
# straight-line programs the reference ships next to the .sms they were generated from (data/Makefile:31-32):
# golden pairs for the SLP -> matrix builder
ALL_OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "all_matrices.json")
SLP_OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "slp_programs.json")
SLP_STEMS = ["2x2x2_7_Winograd", "3x3x3_23_58", "3x4x7_63_rational", "3x4x7_63_rational-ALT", "3x4x7_63_rational-CoB",
             "4x4x4_48_rational", "4x4x4_48_rational-CoB", "4x4x4_48_accurate", "4x4x4_49_156"]


def read_sms(path):
    rows = cols = None
    ent = {}
    header_comments = []
    with open(path) as f:
        for line in f:
            s = line.strip()
            if not s:
                continue
            if s.startswith("#"):
                header_comments.append(s)
                continue
            t = s.split()
            if rows is None:
                rows, cols = int(t[0]), int(t[1])
                continue
            i, j = int(t[0]), int(t[1])
            if i == 0 and j == 0:
                break
            v = Fraction(t[2])
            if v != 0:
                ent[(i - 1, j - 1)] = v
    entries = [[i, j, str(v)] for (i, j), v in sorted(ent.items())]
    return {"rows": rows, "cols": cols, "entries": entries, "comments": header_comments}


def main():
    out = {}
    for stem in STEMS:
        for x in "LRP":
            p = os.path.join(REF, "data", f"{stem}_{x}.sms")
            out[f"{stem}_{x}"] = read_sms(p)
    for s in SINGLES:
        out[s] = read_sms(os.path.join(REF, "data", f"{s}.sms"))
    with open(OUT, "w") as f:
        json.dump(out, f, separators=(",", ":"), sort_keys=True)
    print("wrote", OUT, len(out), "matrices", os.path.getsize(OUT), "bytes")
    # every rational .sms of data/ (bin/FDT.sh and `make mmcheck` run over the whole directory); the three polynomial
    # files *-X_{L,R,P}.sms are out of scope
    allm = {}
    for fn in sorted(os.listdir(os.path.join(REF, "data"))):
        if fn.endswith(".sms"):
            try:
                allm[fn[:-4]] = read_sms(os.path.join(REF, "data", fn))
            except ValueError:
                pass
    with open(ALL_OUT, "w") as f:
        json.dump(allm, f, separators=(",", ":"), sort_keys=True)
    print("wrote", ALL_OUT, len(allm), "matrices", os.path.getsize(ALL_OUT), "bytes")
    slp = {f"{stem}_{x}": open(os.path.join(REF, "data", f"{stem}_{x}.slp")).read() for stem in SLP_STEMS for x in "LRP"}
    with open(SLP_OUT, "w") as f:
        json.dump(slp, f, separators=(",", ":"), sort_keys=True)
    print("wrote", SLP_OUT, len(slp), "programs", os.path.getsize(SLP_OUT), "bytes")


if __name__ == "__main__":
    sys.exit(main())
