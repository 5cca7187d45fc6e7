#!/usr/bin/env python3
"""Records the per-kernel constants bench.py may quote from an ncu capture:
    python tools/ncu_inputs.py RAW_CSV KEY UNITS_PER_LAUNCH SOURCE_NOTE
KEY = "<kernel name>|<m>x<k>x<n>|<nnz|G2>" (what plo_orbit_plan_kernel reports for the plan), UNITS_PER_LAUNCH = candidates of the
captured launch.  Writes profiles/ncu_inputs.json: thread instructions per candidate (smsp__inst_executed.sum x 32 / candidates) and
DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum), shared-memory wavefronts per candidate.  bench.py looks the running plan's kernel up in this file and
quotes nothing when it is absent -- a changed kernel selection cannot reuse a stale profile."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "ncu_inputs.json")
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    raw, key, units, note = sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4]
    rows = list(csv.reader(open(raw)))
    hdr, un, r = rows[0], rows[1], rows[2]
    get = lambda k: (float(r[hdr.index(k)].replace(",", "")), un[hdr.index(k)])
    inst, _ = get("smsp__inst_executed.sum")
    rd, ru = get("dram__bytes_read.sum")
    wr, wu = get("dram__bytes_write.sum")
    try:
        smem, _ = get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
    except ValueError:
        smem = None
    name = r[hdr.index("Kernel Name")]
    assert key.split("|")[0].split("+")[0] in name, (key, name)
    try:
        table = json.load(open(OUT))
    except Exception:
        table = {}
    table[key] = {"inst_per_candidate": inst * 32 / units, "dram_bytes_per_launch": rd * UNIT[ru] + wr * UNIT[wu], "candidates_per_launch": units,
                  "smem_wavefronts_per_candidate": (smem / units) if smem is not None else None,
                  "kernel": name, "source": note}
    json.dump(table, open(OUT, "w"), indent=1, sort_keys=True)
    print(key, table[key])


if __name__ == "__main__":
    main()
