"""Text summary of one kernel in an .ncu-rep (run where ncu is): launch geometry, time, DRAM/L2/tensor/issue metrics, top stall reasons.
usage: python tools/ncu_summary.py report.ncu-rep [kernel-index ...]"""
import csv
import subprocess
import sys

WANT = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.max", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "sm__inst_executed_pipe_tensor.sum", "sm__sass_thread_inst_executed_op_ffma_pred_on.sum"]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(l for l in raw.splitlines() if l.startswith('"')))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    which = [int(x) for x in sys.argv[2:]] or list(range(len(data)))
    for k in which:
        r = data[k]
        print("=" * 100)
        print("%-90s %s" % ("Kernel Name", r[idx["Kernel Name"]]))
        for w in WANT:
            if w in idx:
                print("%-90s %s %s" % (w, r[idx[w]], units[idx[w]]))
        stalls = [(float(r[i].replace(",", "") or 0), h) for h, i in idx.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
        print("\ntop stall reasons (warps stalled per issue-active cycle):")
        for v, h in sorted(stalls, reverse=True)[:8]:
            print("%-90s %f" % (h, v))


if __name__ == "__main__":
    main()
