"""Turn the final round-2 GPU run's scratch output (gpurun_out/) into the tracked artefacts under profiles/:
ncu summaries per kernel, traffic.json (DRAM bytes per launch of the headline kernels), bench / config lines."""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def raw_row(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    return {h: (v, u) for h, v, u in zip(hdr, rows[2], units)}


def fnum(v):
    return float(v.replace(",", ""))


def to_bytes(v, u):
    return fnum(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


def main():
    traffic = {}
    for tag, out in (("c2_f32", "r2_trace_f32_ncu_summary.txt"), ("c2_f64", "r2_trace_f64_ncu_summary.txt"),
                     ("zern_f32", "r2_zernike_f32_ncu_after.txt"), ("zern_f64", "r2_zernike_f64_ncu_after.txt"),
                     ("c5pol_f32", "r2_c5pol_f32_ncu_after.txt"), ("c5pol_f64", "r2_c5pol_f64_ncu_after.txt"),
                     ("bwd_f32", "r2_bwd_f32_ncu_summary.txt"), ("bwd_f64", "r2_bwd_f64_ncu_summary.txt")):
        raw = os.path.join(G, f"r2f_{tag}.raw.csv")
        if not os.path.exists(raw) or os.path.getsize(raw) < 100:
            print("missing", raw)
            continue
        txt = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), raw], capture_output=True, text=True).stdout
        d = raw_row(raw)
        extra = []
        for k in ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
                  "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
                  "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.avg.per_cycle_active",
                  "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
                  "smsp__thread_inst_executed_per_inst_executed.ratio", "derived__smsp__sass_thread_inst_executed_op_spill", "smsp__inst_executed_op_local_ld.sum",
                  "smsp__inst_executed_op_local_st.sum"):
            if k in d:
                extra.append("   %-62s %s %s" % (k, d[k][0], d[k][1]))
        open(os.path.join(P, out), "w").write(txt + "\n".join(extra) + "\n")
        if tag in ("c2_f32", "c2_f64"):
            r, w = d["dram__bytes_read.sum"], d["dram__bytes_write.sum"]
            traffic[tag[3:]] = int(to_bytes(*r) + to_bytes(*w))
    if traffic:
        traffic["_source"] = ("dram__bytes_read.sum + dram__bytes_write.sum of trace_kernel<float,4,0> / <double,1,0> on the headline workload "
                              "(Double-Gauss, 10 M rays, 13 record rows), one `ncu --set full` capture each: profiles/r2_trace_f32_ncu_summary.txt, "
                              "r2_trace_f64_ncu_summary.txt (scripts/r2_final_measure.sh)")
        json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
        print("traffic", traffic)
    for src, dst in (("r2_bench_n1.json", "r2_bench_n1.json"), ("r2_bench_ref.json", "r2_bench_reference_arm.json"),
                     ("r2_configs_final.jsonl", "r2_configs.jsonl"), ("r2_launches_bench.csv", "r2_launches_bench.csv"),
                     ("r2_gputests_final.log", "r2_gputests_final.log")):
        p = os.path.join(G, src)
        if os.path.exists(p) and dst:
            shutil.copy(p, os.path.join(P, dst))
    print("done")


if __name__ == "__main__":
    main()
