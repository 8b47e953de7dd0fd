"""Phase timeline of the conv kernels inside a real prediction() (debug build, -DB200_TIMELINE).
  python tools/timeline.py build                   # here (CPU): compile anakin_b200/lib_tl
  python tools/timeline.py run [--batch 8]         # on the GPU box: graph replays + per-launch phase table
Slots (SM clock, per CTA): 0 entry, 1 prologue done, 2 producer passed the grid-dependency wait,
3 first operand stage landed, 4 last MMA issued, 5 accumulators complete, 6 output staged (before the TMA
store), 7 store read out of smem.  gt0 / gt1 = %globaltimer at entry / exit (ns), which orders launches.
"""
import argparse
import ctypes as C
import glob
import os
import shutil
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB_TL = os.path.join(ROOT, "anakin_b200", "lib_tl")


def build():
    from anakin_b200 import build as B
    B.build_all()
    os.makedirs(LIB_TL, exist_ok=True)
    objs = []
    for src in sorted(glob.glob(os.path.join(B.CSRC, "*.cu"))):
        obj = os.path.join(LIB_TL, os.path.basename(src) + ".o")
        subprocess.check_call([B.NVCC] + B.NVCC_FLAGS + ["-DB200_TIMELINE", "-diag-suppress", "68", "-c", src, "-o", obj])
        objs.append(obj)
    subprocess.check_call([B.NVCC] + B.ARCH + ["-shared", "-o", os.path.join(LIB_TL, "libb200saber.so")] + objs +
                          ["-ccbin", "g++"])
    for o in objs:
        os.remove(o)
    shutil.copy(os.path.join(B.LIBDIR, "libanakin_b200.so"), LIB_TL)
    print("built", LIB_TL)


REC = np.dtype([("gt0", "<u8"), ("gt1", "<u8"), ("clk", "<i8", (8,)), ("bx", "<u4"), ("by", "<u4"), ("bz", "<u4"),
                ("smid", "<u4"), ("K", "<u4"), ("KS", "<u4"), ("bn", "<u4"), ("stages", "<u4")])


def run(batch, model, precision, flush):
    os.environ["ANAKIN_B200_LIBDIR"] = LIB_TL
    import torch
    from anakin_b200 import anakin_bin, api, modelzoo, saber_abi
    g = modelzoo.build(model, batch=batch, precision=precision if precision == "int8" else "fp32")
    G = api.Graph.from_bytes(anakin_bin.dumps(g))
    G.ResetBatchSize("input_0", batch)
    G.Optimize()
    net = api.Net(G, precision)
    net.set_input("input_0", modelzoo.synthetic_input(batch, 224))
    for _ in range(5):
        net.prediction()
    net.sync()
    lib = C.CDLL(os.path.join(LIB_TL, "libb200saber.so"))
    lib.b200_debug_timeline.argtypes = [C.c_void_p, C.c_int]
    lib.b200_debug_timeline(None, 0)          # reset
    fl = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if flush else None
    if fl is not None:
        fl.zero_()
        torch.cuda.synchronize()
    net.prediction()
    net.sync()
    buf = np.zeros(1 << 15, REC)
    n = lib.b200_debug_timeline(buf.ctypes.data_as(C.c_void_p), len(buf))
    recs = buf[:n]
    # group CTAs into launches: same (K, KS, bn) and overlapping time; order by first entry
    order = np.argsort(recs["gt0"], kind="stable")
    recs = recs[order]
    launches = []
    for r in recs:
        key = (int(r["K"]), int(r["KS"]), int(r["bn"]))
        if launches and launches[-1]["key"] == key and r["gt0"] < launches[-1]["gt1"] + 200:
            launches[-1]["recs"].append(r)
            launches[-1]["gt1"] = max(launches[-1]["gt1"], int(r["gt1"]))
        else:
            # a CTA of an earlier launch can enter after the next launch's first CTA: search back a little
            for L in launches[-3:]:
                if L["key"] == key and r["gt0"] < L["gt1"] + 200:
                    L["recs"].append(r)
                    L["gt1"] = max(L["gt1"], int(r["gt1"]))
                    break
            else:
                launches.append({"key": key, "recs": [r], "gt1": int(r["gt1"])})
    t_first = int(recs["gt0"].min())
    clk_ghz = None
    print("%d CTA records, %d launches; times in us relative to the first conv CTA entry" % (n, len(launches)))
    print("%-4s %-16s %5s | %8s %8s %8s %7s %7s | per-CTA medians (us): %7s %7s %6s %6s %6s %6s %6s" %
          ("#", "K/KS/BN", "ctas", "entry", "release", "end", "handoff", "crit", "prolog", "depwait", "load", "mma",
           "drain", "epi", "store"))
    prev_end = None
    tot_handoff = tot_crit = 0.0
    for i, L in enumerate(launches):
        rr = np.array(L["recs"], dtype=REC)
        clk = rr["clk"].astype(np.float64)
        dt_ns = (rr["gt1"].astype(np.float64) - rr["gt0"].astype(np.float64))
        dclk = clk[:, 7] - clk[:, 0]
        ok = (dt_ns > 1500) & (rr["bz"] == 0)
        if ok.any():
            clk_ghz = float(np.median(dclk[ok] / dt_ns[ok]))
        f = 1e-3 / (clk_ghz or 1.9)     # clocks -> us
        med = lambda a: float(np.median(a)) if len(a) else float("nan")
        r0 = rr["bz"] == 0
        ph = [med((clk[:, 1] - clk[:, 0]) * f), med((clk[:, 2] - clk[:, 1]) * f), med((clk[:, 3] - clk[:, 2]) * f),
              med((clk[:, 4] - clk[:, 3]) * f), med((clk[:, 5] - clk[:, 4]) * f),
              med((clk[r0][:, 6] - clk[r0][:, 5]) * f), med((clk[r0][:, 7] - clk[r0][:, 6]) * f)]
        entry = (int(rr["gt0"].min()) - t_first) / 1e3
        rel_each = (rr["gt0"].astype(np.float64) - t_first) / 1e3 + (clk[:, 2] - clk[:, 0]) * f
        release = float(rel_each.min())
        end = (int(rr[r0]["gt1"].max()) - t_first) / 1e3
        handoff = (release - prev_end) if prev_end is not None else 0.0
        crit = end - release
        tot_handoff += handoff
        tot_crit += crit
        print("%-4d %-16s %5d | %8.2f %8.2f %8.2f %7.2f %7.2f | %21s %7.2f %7.2f %6.2f %6.2f %6.2f %6.2f %6.2f" %
              (i, "%d/%d/%d" % L["key"], len(rr), entry, release, end, handoff, crit, "", *ph))
        prev_end = end
        if rr["bz"].max() > 0:
            # split-K launch: absolute phase times per rank, relative to the launch's first dependency release
            base = (rr["gt0"].astype(np.float64) - t_first) / 1e3
            for z in range(int(rr["bz"].max()) + 1):
                m = rr["bz"] == z
                absu = lambda k: float(np.median(base[m] + (clk[m][:, k] - clk[m][:, 0]) * f)) - release
                print("       rank %d: entry %+6.2f prolog %+6.2f release %+6.2f first-stage %+6.2f last-mma %+6.2f acc-done %+6.2f%s" %
                      (z, float(np.median(base[m])) - release, absu(1), absu(2), absu(3), absu(4), absu(5),
                       (" staged %+6.2f stored %+6.2f" % (absu(6), absu(7))) if z == 0 else ""))
    print("sum handoff (prev end -> dependency released) %.1f us; sum crit (release -> last exit) %.1f us" % (tot_handoff, tot_crit))
    print("SM clock estimate %.3f GHz" % (clk_ghz or 0))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("cmd", choices=["build", "run"])
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--model", default="resnet50")
    ap.add_argument("--precision", default="int8")
    ap.add_argument("--flush", type=int, default=1)
    a = ap.parse_args()
    build() if a.cmd == "build" else run(a.batch, a.model, a.precision, a.flush)
