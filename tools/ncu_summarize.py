"""Condense an .ncu-rep into a small text summary (key raw metrics + hottest SASS lines) for profiles/."""
import csv, subprocess, sys, io

rep, out = sys.argv[1], sys.argv[2]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "sm__cycles_elapsed.max", "lts__t_bytes.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__mio_inst_issued.avg.pct_of_peak_sustained_active", "smsp__average_warp_latency_issue_stalled_barrier.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
with open(out, "w") as f:
    f.write(f"# summary of {rep} (ncu --set full --clock-control none; cold-cache serialised replays)\n")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        f.write(f"\n## {d.get('Kernel Name','?')[:100]}  (ID {d.get('ID','?')})\n")
        for k in KEYS:
            if k in d and d[k] != "":
                f.write(f"{k:85s} {d[k]} {units[hdr.index(k)]}\n")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    if len(srows) > 3:
        sh = srows[1]; ix = {h: i for i, h in enumerate(sh)}
        data = [r for r in srows[2:] if len(r) == len(sh)]
        tot = sum(int(r[ix["# Samples"]] or 0) for r in data) or 1
        stalls = [h for h in sh if h.startswith("stall_") and "Not Issued" not in h]
        f.write(f"\n## hottest SASS instructions (warp-state samples, total {tot})\n")
        for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:25]:
            st = " ".join(f"{h[6:]}={r[ix[h]]}" for h in stalls if int(r[ix[h]] or 0) > tot * 0.002)
            f.write(f"{int(r[ix['# Samples']]):7d} {100*int(r[ix['# Samples']])/tot:5.1f}%  exec={r[ix['Instructions Executed']]:>9s}  {r[ix['Source']][:80]:80s} {st}\n")
        sass = " ".join(r[ix["Source"]] for r in data)
        f.write("\n## Blackwell-native evidence in SASS: " + ", ".join(f"{m}:{sass.count(m)}" for m in ["UTCHMMA", "UTCMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS"]) + "\n")
print(open(out).read()[:1500])
