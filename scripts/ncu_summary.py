"""Summarise an .ncu-rep (read on the CPU box): key counters, stall reasons, instruction mix.
usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep [particles_per_launch]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
npart = float(sys.argv[2]) if len(sys.argv) > 2 else None


def page(name, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"] + list(extra), capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


rows = page("raw")
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    u = dict(zip(hdr, units))
    print("== kernel:", d.get("Kernel Name", "?")[:60], " grid", d.get("Grid Size"), "block", d.get("Block Size"))
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "smsp__inst_executed.sum", "sm__cycles_elapsed.avg.per_second",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum",
            "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
            "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
            "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "smsp__average_warp_latency_per_inst_issued.ratio"]
    for k in keys:
        if k in d:
            print("  %-62s %s %s" % (k, d[k], u.get(k, "")))
    if npart and "smsp__inst_executed.sum" in d:
        print("  warp-inst per 32 particles: %.1f" % (float(d["smsp__inst_executed.sum"]) / (npart / 32)))
    st = {k: float(v) for k, v in d.items() if "issue_stalled" in k and k.endswith(".ratio") and v}
    print("  stalls (warps per issue):", ", ".join("%s %.2f" % (k.split("issue_stalled_")[1].split("_per_")[0], v)
                                                    for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
    pipes = {k: float(v) for k, v in d.items() if k.startswith("sm__inst_executed_pipe_") and k.endswith("avg.pct_of_peak_sustained_active") and v}
    print("  pipes %:", ", ".join("%s %.1f" % (k.split("pipe_")[1].split(".")[0], v) for k, v in sorted(pipes.items(), key=lambda kv: -kv[1])[:8]))

src = page("source", ["--print-source", "sass"])
if len(src) > 2:
    h = src[1]
    iS, iE, iW = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
    byop, stall, data, tot, totw = collections.Counter(), collections.Counter(), [], 0, 0
    for r in src[2:]:
        try:
            e, w = int(r[iE]), int(r[iW])
        except (ValueError, IndexError):
            continue
        toks = r[iS].split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        byop[op] += e
        stall[op] += w
        tot += e
        totw += w
        data.append((e, w, r[iS].strip()))
    print("== SASS: %d static instructions, %d executed warp-inst, %d stall samples" % (len(data), tot, totw))
    scale = (npart / 32) if npart else 1.0
    print("  op mix (per 32 particles):", ", ".join("%s %.1f" % (o, c / scale) for o, c in byop.most_common(22)))
    print("  top stall sites:")
    for e, w, s in sorted(data, key=lambda x: -x[1])[:16]:
        print("    %5.1f%%  %9d  %s" % (100.0 * w / max(totw, 1), e, s[:80]))
