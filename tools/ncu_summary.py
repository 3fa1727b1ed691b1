#!/usr/bin/env python3
"""Summarise an .ncu-rep: headline metrics + stall samples per SASS region.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--regions]"""
import csv, io, subprocess, sys

def run(args):
    return subprocess.run(["ncu", "-i", *args], capture_output=True, text=True).stdout

def main():
    rep = sys.argv[1]
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "lts__t_sector_hit_rate.pct", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
            "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum",
            "smsp__sass_inst_executed_op_global_ld.sum", "smsp__sass_inst_executed_op_global_st.sum",
            "smsp__inst_executed_op_tma_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]
    for i, h in enumerate(hdr):
        if h in want or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
            vals = [r[i] for r in data]
            print(f"{h} [{units[i]}]: {', '.join(v[:60] for v in vals)}")
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv"]))))
    # first kernel block only
    hdr = src[1]
    ix = {h: i for i, h in enumerate(hdr)}
    body = []
    for r in src[2:]:
        if r and r[0] == "Kernel Name":
            break
        if len(r) >= len(hdr) - 5 and r[0] != "Address":
            body.append(r)
    tot = sum(int(r[ix["# Samples"]]) for r in body)
    print("total samples", tot, "SASS instructions", len(body))
    keys = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    agg = {k: sum(int(r[ix[k]]) for r in body) for k in keys}
    print("stalls overall:", {k[6:]: round(100 * v / tot, 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.01 * tot})
    # regions split at BAR / SYNCS / branch-back markers
    marks = [i for i, r in enumerate(body) if any(t in r[ix["Source"]] for t in ("BAR.SYNC", "SYNCS.PHASECHK", "UTMALDG", "EXIT"))]
    print("markers:", [(i, body[i][ix["Source"]].strip()[:40]) for i in marks][:40])
    top = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]]))[:16]
    for i in top:
        r = body[i]
        st = {k[6:]: int(r[ix[k]]) for k in keys if int(r[ix[k]]) > 0.15 * max(1, int(r[ix["# Samples"]]))}
        print(f"{i:5d} {r[ix['Source']].strip()[:58]:58s} {r[ix['# Samples']]:>6s} {st}")
    if "--regions" in sys.argv:
        edges = [0] + marks + [len(body)]
        for a, b in zip(edges[:-1], edges[1:]):
            s = sum(int(r[ix["# Samples"]]) for r in body[a:b])
            ex = sum(int(r[ix["Instructions Executed"]]) for r in body[a:b])
            if s > 0.01 * tot:
                st = {k[6:]: sum(int(r[ix[k]]) for r in body[a:b]) for k in keys}
                print(f"[{a:5d},{b:5d}) samples {s:6d} ({100*s/tot:4.1f}%) inst {ex:10d}", {k: v for k, v in st.items() if v > 0.08 * s})

if __name__ == "__main__":
    main()
