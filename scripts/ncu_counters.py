#!/usr/bin/env python
"""Reduces an ncu `--set full` raw CSV (or .ncu-rep) of `bench.py` to the per-launch counter table that bench.py
divides by LIVE CUDA-event durations (profiles/r2_kernel_counters.json):

    python scripts/ncu_counters.py <workload> <raw.csv | file.ncu-rep> [--out profiles/r2_kernel_counters.json]

Per kernel family (blend_fwd, blend_bwd, preprocess_fwd, preprocess_bwd, tile_place, ...): the MEDIAN over the
captured launches of instructions executed, shared-memory wavefronts, DRAM bytes (read + write), issue-active %
and the (cold, serialised) duration ncu saw.  The counts are properties of the workload (same seed and views as
the bench), which is what lets the bench turn them into live rates."""
import csv
import io
import json
import os
import re
import statistics
import subprocess
import sys

FAMILIES = [("blend_fwd", r"blend_fwd_kernel"), ("blend_bwd", r"blend_bwd_kernel"),
            ("preprocess_fwd", r"preprocess_fwd_kernel"),
            # <3, *> = deferred SH gradient (the timed region of the bench); <1, *> / <2, *> = the row-writing variants of the module path
            ("preprocess_bwd", r"preprocess_bwd_kernel<3"), ("preprocess_bwd_rows", r"preprocess_bwd_kernel<[12]"),
            ("tile_place", r"tile_place_kernel"), ("tile_count", r"tile_count_kernel"),
            ("sh_grad_expand", r"sh_grad_expand_kernel"), ("radix_scatter", r"radix_scatter_kernel"), ("onesweep_pass", r"onesweep_pass_kernel"),
            ("bind_preprocess_fwd", r"bind_preprocess_fwd_kernel"), ("bind_preprocess_bwd", r"bind_preprocess_bwd_kernel")]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "msecond": 1e-3,
        "nsecond": 1e-9, "ms": 1e-3, "second": 1.0}


def main():
    workload, src = sys.argv[1], sys.argv[2]
    out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else "profiles/r2_kernel_counters.json"
    txt = open(src).read() if src.endswith(".csv") else subprocess.run(
        ["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}

    def val(r, name):
        if name not in idx:
            return None
        try:
            f = float(r[idx[name]].replace(",", ""))
        except ValueError:
            return None
        return f * UNIT.get(units[idx[name]], 1.0)

    fam = {}
    for r in rows[2:]:
        kname = r[idx["Kernel Name"]]
        for f, pat in FAMILIES:
            if re.search(pat, kname):
                rd, wr = val(r, "dram__bytes_read.sum") or 0.0, val(r, "dram__bytes_write.sum") or 0.0
                fam.setdefault(f, []).append({
                    "inst_executed": val(r, "smsp__inst_executed.sum"),
                    "smem_wavefronts": val(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
                    "dram_bytes": rd + wr,
                    "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                    "ncu_duration_us": (val(r, "gpu__time_duration.sum") or 0.0) * 1e6,
                    "registers": val(r, "launch__registers_per_thread"),
                })
                break
    table = {}
    for f, launches in fam.items():
        med = {}
        for k in launches[0]:
            vs = [l[k] for l in launches if l[k] is not None]
            med[k] = statistics.median(vs) if vs else None
        med["launches_in_capture"] = len(launches)
        table[f] = med
    # (warp square, list entry) pairs that survive the culling = trips of the blend kernels' inner loops
    # (scripts/blend_stats.py, backward range; the forward walks a little further, to each square's saturation)
    stats = os.path.join(os.path.dirname(out), "r2_blend_stats.json")
    if os.path.exists(stats):
        with open(stats) as fh:
            st = json.load(fh).get(workload)
        if st and "blend_bwd" in table:
            table["blend_bwd"]["warp_entry_pairs"] = st["8x8"]["survivors"]
            table["blend_bwd"]["pixel_entry_hits"] = st["8x8"]["pixel_hits"]
    doc = {"source": None, "workloads": {}}
    if os.path.exists(out):
        with open(out) as fh:
            doc = json.load(fh)
    doc["source"] = doc.get("source") or {}
    if not isinstance(doc["source"], dict):
        doc["source"] = {}
    doc["source"][workload] = os.path.relpath(src)
    doc["workloads"][workload] = table
    with open(out, "w") as fh:
        json.dump(doc, fh, indent=1)
        fh.write("\n")
    print(json.dumps(table, indent=1))


if __name__ == "__main__":
    main()
