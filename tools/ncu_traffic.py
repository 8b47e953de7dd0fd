"""Per-launch table and the roofline `traffic` record from an `ncu --set full` report of one prediction().
  python tools/ncu_traffic.py gpurun_out/x.ncu-rep profiles/r02_ncu_conv_per_launch.txt profiles/r02_conv_traffic.json
The JSON is what bench.py reads for roofline.traffic: DRAM and L2 bytes per conv launch, mean over the conv launches of
the step (cold-cache capture: ncu flushes caches and serialises the kernels)."""
import csv
import json
import re
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "dur"), ("dram__bytes_read.sum", "dram_r"), ("dram__bytes_write.sum", "dram_w"),
        ("lts__t_bytes.sum", "lts"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"), ("launch__grid_size", "grid"),
        ("launch__shared_mem_per_block_dynamic", "smem"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_pct")]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6,
        "ns": 1e-3, "us": 1.0, "ms": 1e3}


def main(rep, txt_out, json_out):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {short: (hdr.index(name) if name in hdr else -1) for name, short in COLS}
    name_i = hdr.index("Kernel Name")
    recs = []
    for r in rows[2:]:
        rec = {"kernel": re.sub(r"\(.*", "", r[name_i]).replace("void ", "").replace("b200::", "")}
        for name, short in COLS:
            i = idx[short]
            if i < 0 or not r[i]:
                rec[short] = None
                continue
            v = float(r[i].replace(",", ""))
            rec[short] = v * UNIT.get(units[i], 1.0)
        recs.append(rec)
    convs = [r for r in recs if r["kernel"].startswith("conv_")]
    with open(txt_out, "w") as f:
        f.write("ncu --set full --clock-control none, one eager prediction(): every launch (cold caches, serialised)\n")
        f.write("%-3s %-44s %8s %9s %9s %9s %8s %6s %7s %5s\n" % ("#", "kernel", "us", "dram_rd_KB", "dram_wr_KB", "L2_KB",
                                                                  "tensor%", "grid", "smem_KB", "regs"))
        for i, r in enumerate(recs):
            f.write("%-3d %-44s %8.2f %9.0f %9.0f %9.0f %8s %6d %7.1f %5d\n" % (
                i, r["kernel"][:44], r["dur"], (r["dram_r"] or 0) / 1e3, (r["dram_w"] or 0) / 1e3, (r["lts"] or 0) / 1e3,
                ("%.1f" % r["tensor_pct"]) if r["tensor_pct"] is not None else "-", int(r["grid"] or 0),
                (r["smem"] or 0) / 1e3, int(r["regs"] or 0)))
        tot = sum(r["dur"] for r in recs)
        f.write("total %.1f us over %d launches; conv kernels %.1f us (%d launches)\n" % (
            tot, len(recs), sum(r["dur"] for r in convs), len(convs)))
    if convs:
        n = len(convs)
        j = {"source": rep.split("/")[-1], "conv_launches": n,
             "dram_bytes_per_launch": sum((r["dram_r"] or 0) + (r["dram_w"] or 0) for r in convs) / n,
             "dram_read_bytes_per_prediction": sum(r["dram_r"] or 0 for r in convs),
             "dram_write_bytes_per_prediction": sum(r["dram_w"] or 0 for r in convs),
             "lts_bytes_per_launch": sum(r["lts"] or 0 for r in convs) / n,
             "conv_us_cold": sum(r["dur"] for r in convs),
             "note": "ncu --set full --clock-control none of one eager ResNet-50 INT8 b8 prediction(); caches flushed per kernel"}
        with open(json_out, "w") as f:
            json.dump(j, f, indent=1)
        print(json.dumps(j))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3])
