"""Markdown table of the gemm_tc_kernel launches in an `ncu --set full` report:
python tools/summarize_ncu_gemm.py gpurun_out/prof_gemm_step.ncu-rep [traffic.json]"""
import csv, io, json, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
def col(name): return h.index(name)
cols = {"name": col("Kernel Name"), "us": col("gpu__time_duration.sum"), "rd": col("dram__bytes_read.sum"), "wr": col("dram__bytes_write.sum"),
        "tensor": col("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"), "ghz": col("sm__cycles_elapsed.avg.per_second"),
        "lts": col("lts__throughput.avg.pct_of_peak_sustained_elapsed"), "issue": col("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "l1": col("l1tex__throughput.avg.pct_of_peak_sustained_active"), "regs": col("launch__registers_per_thread"), "cluster": col("launch__cluster_size")}
units = rows[1]
def mb(v, u): return float(v) * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[u]
print("| # | kernel instance | µs | DRAM read MB | DRAM write MB | tensor pipe active % | SM clock GHz | L2 % | L1tex % | issue % | regs | cluster |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
tr = []
for i, r in enumerate(rows[2:]):
    name = r[cols["name"]].split("(")[0].replace("void ", "").replace("theia::", "")
    rd, wr = mb(r[cols["rd"]], units[cols["rd"]]), mb(r[cols["wr"]], units[cols["wr"]])
    tr.append(rd + wr)
    print(f"| {i} | `{name}` | {float(r[cols['us']]):.1f} | {rd:.1f} | {wr:.1f} | {float(r[cols['tensor']]):.1f} | {float(r[cols['ghz']]):.3f} | "
          f"{float(r[cols['lts']]):.1f} | {float(r[cols['l1']]):.1f} | {float(r[cols['issue']]):.1f} | {r[cols['regs']]} | {r[cols['cluster']]} |")
print(f"\nmean DRAM traffic per launch: {sum(tr)/len(tr):.1f} MB over {len(tr)} launches")
if len(sys.argv) > 2:
    json.dump({"kernel": "gemm_tc_kernel", "traffic_bytes_per_launch_mean": int(sum(tr) / len(tr) * 1e6), "launches": len(tr),
               "source": "ncu --set full --clock-control none, consecutive gemm_tc_kernel launches of one ViT block forward inside the base+cdiv step"},
              open(sys.argv[2], "w"), indent=1)
