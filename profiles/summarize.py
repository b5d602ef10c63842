"""Turns the ncu captures under gpurun_out/ (scratch) into the tracked summaries under profiles/.
Usage: python profiles/summarize.py r1      (expects gpurun_out/<tag>_launches.csv and <tag>_step_kernel.ncu-rep)"""
import csv, io, json, subprocess, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
CMD = ("python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1" if tag == "r1" else
       "python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-workloads --no-her --e2e-steps 1 --e2e-fused 8")
raw = [l for l in open(f"gpurun_out/{tag}_launches.csv") if not l.startswith("==")]
rows = list(csv.DictReader(io.StringIO("".join(raw))))
agg = {}
for r in rows:
    k = r["Kernel Name"].split("(")[0][:90]
    agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += float(r["Metric Value"].replace(",", ""))
tot = sum(v[1] for v in agg.values())
lines = [f"# ncu --metrics gpu__time_duration.sum --clock-control none : {CMD}",
         "# per-launch times are cold-cache and serialised: compare SHARES, not absolutes", "kernel,launches,total_ms,share"]
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    lines.append('"%s",%d,%.3f,%.4f' % (k, v[0], v[1] / 1e6, v[1] / tot))
open(f"profiles/{tag}_launch_summary.csv", "w").write("\n".join(lines) + "\n")
open(f"profiles/{tag}_launches.csv", "w").write("".join(raw))

out = subprocess.run(f"ncu -i gpurun_out/{tag}_step_kernel.ncu-rep --page raw --csv", shell=True, capture_output=True, text=True).stdout
rr = list(csv.reader(out.splitlines()))
hdr, units, val = rr[0], rr[1], rr[2]
keep = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        # instruction delivery (round 2): SM instruction cache and the GPC-level cache behind it
        "sm__icc_requests.sum", "sm__icc_request_hit_rate.pct", "gcc__cache_requests_type_instruction.sum",
        "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
d = {h: (v, u) for h, u, v in zip(hdr, units, val) if h in keep}
txt = [f"# ncu --set full --clock-control none --import-source on -k regex:step_kernel_async -s 3 -c 1 : {CMD}",
       "# one launch = 1,048,576 BlocksTouch-v0 envs x K = 64 fused steps = 67,108,864 env-steps"]
txt += ["%s = %s %s" % (k, d[k][0], d[k][1]) for k in keep if k in d]


def num(k):
    v, u = d[k]
    return float(v.replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)


tr = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
steps = 67108864
txt.append("derived: dram bytes per env-step = %.1f (algorithmic 252.4); warp instructions per env-step = %.1f" %
           (tr / steps, float(d["smsp__inst_executed.sum"][0].replace(",", "")) / steps))
open(f"profiles/{tag}_step_kernel_ncu_summary.txt", "w").write("\n".join(txt) + "\n")
json.dump({"dram_bytes_per_launch": tr, "dram_read_bytes": num("dram__bytes_read.sum"), "dram_write_bytes": num("dram__bytes_write.sum"),
           "env_steps_per_launch": steps, "dram_bytes_per_env_step": tr / steps,
           "source": f"profiles/{tag}_step_kernel_ncu_summary.txt (ncu --set full, one launch of step_kernel_async<1,4>)"},
          open("profiles/roofline_traffic.json", "w"), indent=1)
print("\n".join(txt))

# ---- SASS-level table of the same capture: where the issue slots go and at how many active lanes
src = subprocess.run(f"ncu -i gpurun_out/{tag}_step_kernel.ncu-rep --page source --csv --print-source sass", shell=True, capture_output=True, text=True).stdout
rs = list(csv.reader(src.splitlines()))
if len(rs) > 3:
    h = rs[1]
    ix = {k: i for i, k in enumerate(h)}
    body = [r for r in rs[2:] if len(r) > 10]
    ex = [int(r[ix["Instructions Executed"]]) for r in body]
    th = [int(r[ix["Thread Instructions Executed"]]) for r in body]
    sm = [int(r[ix["# Samples"]]) for r in body]
    tot, tth, tsm = sum(ex), sum(th), sum(sm)
    low = sum(e for e, t in zip(ex, th) if e and t / e <= 4.0)
    out = [f"# ncu --page source --print-source sass of gpurun_out/{tag}_step_kernel.ncu-rep ({len(body)} SASS instructions = {len(body) * 16 // 1024} KB)",
           "# warp-instructions executed %d = %.1f per env-step; average active lanes %.1f; share executed at <= 4 active lanes %.1f %%" % (
               tot, tot / steps, tth / tot, 100.0 * low / tot),
           "# columns: first SASS index of the 120-instruction chunk, share of executed warp-instructions, average active lanes, share of stall samples, "
           "no_instruction / wait / branch_resolving / long_scoreboard shares of the chunk's samples, most frequent opcodes",
           "chunk,exec_share_pct,avg_lanes,sample_share_pct,no_inst_pct,wait_pct,branch_pct,long_sb_pct,opcodes"]
    N = 120
    for i in range(0, len(body), N):
        ch = body[i:i + N]
        e, t, s_ = sum(ex[i:i + N]), sum(th[i:i + N]), sum(sm[i:i + N])
        if e == 0:
            continue
        ops = {}
        for r in ch:
            parts = r[ix["Source"]].split()
            op = parts[1] if parts[0].startswith("@") else parts[0]
            ops[op.split(".")[0]] = ops.get(op.split(".")[0], 0) + 1
        top = " ".join("%s:%d" % kv for kv in sorted(ops.items(), key=lambda kv: -kv[1])[:5])
        st = lambda k: 100.0 * sum(int(r[ix[k]]) for r in ch) / max(s_, 1)
        out.append("%d,%.2f,%.1f,%.2f,%.1f,%.1f,%.1f,%.1f,%s" % (i, 100.0 * e / tot, t / e, 100.0 * s_ / tsm, st("stall_no_inst"), st("stall_wait"),
                                                           st("stall_branch_resolving"), st("stall_long_sb"), top))
    open(f"profiles/{tag}_step_kernel_sass_regions.csv", "w").write("\n".join(out) + "\n")
    print("\n".join(out[:3]))
