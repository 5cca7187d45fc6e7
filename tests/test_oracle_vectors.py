"""CPU: the oracle re-derives the committed golden vectors (tests/golden/oracle_vectors.json, made by
tests/golden/make_oracle_vectors.py): sparsifier traces of C1 / C3, the exhaustive 48^3 orbit table of C2, Philox decodes,
Factorizer and dependency samples; plus the engine's HOST functions (no GPU) against the same vectors."""
import importlib.util
import json
import os

import numpy as np

import oracle_lib as O
from plinopt_b200 import capi

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "oracle_vectors.json")))
spec = importlib.util.spec_from_file_location("make_oracle_vectors", os.path.join(HERE, "golden", "make_oracle_vectors.py"))
gen = importlib.util.module_from_spec(spec)
spec.loader.exec_module(gen)


def test_oracle_reproduces_every_golden_vector():
    now = json.loads(json.dumps(gen.build(), sort_keys=True))
    assert sorted(now) == sorted(GOLD)
    for key in GOLD:
        assert now[key] == GOLD[key], key


def test_c1_trace_shape():
    """BASELINE config 1 (`sparsifier -c 4`): SparseFactor runs c = 3, 7, 11, ... first (Q8) and -c only in the second pass; every step
    records (block, num, rl, cl, index, c) and the factorisation is consistent (bin/FDT.sh:64-66)."""
    g = GOLD["C1_c4"]
    assert g["consistent"] and len(g["trace"]) >= 4
    assert {t["c"] for t in g["trace"]} >= {3, 4}
    assert all(t["index"] < t["c"] ** 4 for t in g["trace"] if t["index"] >= 0)


def test_c2_table_summary():
    g = GOLD["C2_exhaustive"]
    assert sum(g["nnz_hist"].values()) == 110592 and g["nnz_min"] == 36
    assert abs(g["g2_min"] - (12 + 2 * 2 ** 0.5)) < 1e-12  # the Winograd orbit contains a 14.83-point (Strassen-like growth factor)
    assert g["best_g2"][0] == g["g2_min_index"]


def test_host_decodes_match_golden():
    """plo_orbit_decode / plo_factor_decode are host functions of the engine (same templated code as the device decode)."""
    for key, (U, V, W) in GOLD["orbit_decode_philox"].items():
        shape, idx = key.split(":")
        m, k, n = (int(x) for x in shape.split("x"))
        gU, gV, gW = capi.orbit_decode(m, k, n, 1, gen.SEED, int(idx))
        assert [np.asarray(gU).reshape(-1).tolist(), np.asarray(gV).reshape(-1).tolist(), np.asarray(gW).reshape(-1).tolist()] == [U, V, W]
    assert capi.factor_decode(48, gen.SEED, 17).tolist() == GOLD["factor_4x4x4_L_k16"]["order_17"]
