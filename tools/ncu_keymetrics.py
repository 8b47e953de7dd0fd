"""Print the roofline-relevant metrics of every kernel in an ncu report (`ncu --set full`).
  python tools/ncu_keymetrics.py gpurun_out/x.ncu-rep"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "l1tex__data_bank_conflicts_pipe_lsu.sum",
        "smsp__inst_executed.sum"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")][:140])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print("  %-68s %s %s" % (w, r[i], units[i]))


if __name__ == "__main__":
    main(sys.argv[1])
