#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): key raw metrics per launch and stall samples per SASS opcode.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [--sass N]    # N = print the N hottest SASS lines
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ldgsts.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "smsp__inst_executed.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "smsp__average_warps_issue_stalled"]


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    nsass = int(sys.argv[sys.argv.index("--sass") + 1]) if "--sass" in sys.argv else 0
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print("launches:", len(data))
    name_i = hdr.index("Kernel Name")
    for r in data:
        print("  ", r[name_i][:110])
    for i, h in enumerate(hdr):
        if any(h == k or (k.startswith("smsp__average_warps_issue_stalled") and h.startswith(k) and h.endswith("per_issue_active.ratio")) for k in KEYS):
            vals = [r[i] for r in data]
            if all(v in ("0", "") for v in vals):
                continue
            print(f"{h:100s} {units[i]:12s} {vals}")
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv", "--print-source", "sass"]))))
    k = 0
    while k < len(src):
        if src[k] and src[k][0] == "Kernel Name":
            kname = src[k][1][:100]
            h = src[k + 1]
            idx = {x: i for i, x in enumerate(h)}
            k += 2
            body = []
            while k < len(src) and not (src[k] and src[k][0] == "Kernel Name"):
                if len(src[k]) >= len(h) - 2:
                    body.append(src[k])
                k += 1
            stalls = [x for x in h if x.startswith("stall_") and "Not Issued" not in x]
            byop, execd, st = collections.Counter(), collections.Counter(), collections.Counter()
            for r in body:
                toks = r[idx["Source"]].split()
                if not toks:
                    continue
                op = (toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]).split(".")[0]
                n = int(r[idx["# Samples"]] or 0)
                byop[op] += n
                execd[op] += int(r[idx["Instructions Executed"]] or 0)
                for s in stalls:
                    if r[idx[s]]:
                        st[(op, s)] += int(r[idx[s]])
            tot = sum(byop.values()) or 1
            print(f"\n== {kname}: {len(body)} SASS instructions, {tot} samples")
            for op, n in byop.most_common(12):
                print(f"  {op:8s} {100 * n / tot:5.1f}%  executed {execd[op]:13d}  " + " ".join(f"{s[6:]}={st[(op, s)]}" for s in stalls if st[(op, s)] > 0.04 * n))
            if nsass:
                hot = sorted(body, key=lambda r: -int(r[idx["# Samples"]] or 0))[:nsass]
                for r in hot:
                    print(f"    {r[idx['Address']][-6:]} {r[idx['Source']][:80]:80s} samples={r[idx['# Samples']]} exec={r[idx['Instructions Executed']]}")
        else:
            k += 1


if __name__ == "__main__":
    main()
