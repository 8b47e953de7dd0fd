"""Tile / split-K tuning experiment over every distinct conv shape of ResNet-50 INT8: times single
layers (CUDA-graph replay of a 20-launch PDL chain, so host launch cost is excluded) under forced
BN / split settings and reports, weighted by how often each shape occurs in the net, what the
plan heuristic leaves on the table.
  python tools/tile_tune.py [batch]        # runs the grid in subprocesses, prints the summary
  python tools/tile_tune.py one <batch>    # one configuration from the environment
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# name: (count, h, w, c, k, r, stride, pad, residual, relu)
LAYERS = [
    ("s2 1x1 64>256 br1", 1, 56, 56, 64, 256, 1, 1, 0, False, False),
    ("s2 1x1 64>64", 1, 56, 56, 64, 64, 1, 1, 0, False, True),
    ("s2 3x3 64", 3, 56, 56, 64, 64, 3, 1, 1, False, True),
    ("s2 1x1 64>256 +res", 3, 56, 56, 64, 256, 1, 1, 0, True, True),
    ("s2 1x1 256>64", 2, 56, 56, 256, 64, 1, 1, 0, False, True),
    ("s3 1x1/2 256>512 br1", 1, 56, 56, 256, 512, 1, 2, 0, False, False),
    ("s3 1x1/2 256>128", 1, 56, 56, 256, 128, 1, 2, 0, False, True),
    ("s3 3x3 128", 4, 28, 28, 128, 128, 3, 1, 1, False, True),
    ("s3 1x1 128>512 +res", 4, 28, 28, 128, 512, 1, 1, 0, True, True),
    ("s3 1x1 512>128", 3, 28, 28, 512, 128, 1, 1, 0, False, True),
    ("s4 1x1/2 512>1024 br1", 1, 28, 28, 512, 1024, 1, 2, 0, False, False),
    ("s4 1x1/2 512>256", 1, 28, 28, 512, 256, 1, 2, 0, False, True),
    ("s4 3x3 256", 6, 14, 14, 256, 256, 3, 1, 1, False, True),
    ("s4 1x1 256>1024 +res", 6, 14, 14, 256, 1024, 1, 1, 0, True, True),
    ("s4 1x1 1024>256", 5, 14, 14, 1024, 256, 1, 1, 0, False, True),
    ("s5 1x1/2 1024>2048 br1", 1, 14, 14, 1024, 2048, 1, 2, 0, False, False),
    ("s5 1x1/2 1024>512", 1, 14, 14, 1024, 512, 1, 2, 0, False, True),
    ("s5 3x3 512", 3, 7, 7, 512, 512, 3, 1, 1, False, True),
    ("s5 1x1 512>2048 +res", 3, 7, 7, 512, 2048, 1, 1, 0, True, True),
    ("s5 1x1 2048>512", 2, 7, 7, 2048, 512, 1, 1, 0, False, True),
    ("fc 2048>1000", 1, 1, 1, 2048, 1000, 1, 1, 0, False, False),
]


def one(batch):
    import torch
    from anakin_b200 import saber_abi as A
    from gpu_util import ConvRunner, dev
    rng = np.random.default_rng(0)
    for name, cnt, h, w, c, k, r, stride, pad, has_res, relu in LAYERS:
        n = batch
        x = rng.integers(0, 255, (n, h, w, c)).astype(np.uint8)
        wt = rng.integers(-127, 128, (k, c, r, r)).astype(np.int8)
        fc = name.startswith("fc")
        odt = A.FLOAT if fc else (A.UINT8 if relu else A.INT8)
        ldc = k if fc else (k + 15) // 16 * 16
        try:
            run = ConvRunner(A.MATH_I8, x.shape, A.UINT8, wt, np.zeros(k, np.float32), np.full(k, 1e-4, np.float32),
                             odt, res_dtype=(A.UINT8 if has_res else -1), stride=(stride, stride), pad=(pad, pad),
                             relu=relu, ldc=ldc, sum_scale=0.5)
        except Exception as e:
            print("%-24s plan failed: %s" % (name, e))
            continue
        xd = dev(x)
        res = dev(rng.integers(0, 255, (n, run.ho, run.wo, ldc)).astype(np.uint8)) if has_res else None
        out = run.run(xd, res)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for _ in range(3):
                run.run(xd, res, out_dev=out)
            s.synchronize()
            with torch.cuda.graph(g, stream=s):
                for _ in range(20):
                    run.run(xd, res, out_dev=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g.replay()
        best = 1e9
        for _ in range(3):
            e0.record()
            for _ in range(5):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 100 * 1e3)
        info = run.info()
        print("%-24s|%d|BN=%-3d grid=%dx%dx%d smem=%-6d|%7.2f" % (name, cnt, info["block_n"], info["grid_x"], info["grid_y"],
                                                             info.get("split", 0), info["smem"], best))


def main(batch):
    table = {}   # layer -> {cfg: us}
    counts = {}
    cfgs = [("auto", "")] + [(bn, sp) for bn, sps in (("32", "12"), ("64", "124"), ("128", "1248"), ("256", "1")) for sp in sps]
    for bn, split in cfgs:
        env = dict(os.environ)
        if bn != "auto":
            env["B200_SABER_FORCE_BN"] = bn
            env["B200_SABER_FORCE_SPLIT"] = split
        cfg = "auto" if bn == "auto" else "%s/%s" % (bn, split)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "one", str(batch)], env=env,
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
        for ln in r.stdout.splitlines():
            parts = ln.split("|")
            if len(parts) == 4:
                name = parts[0].strip()
                counts[name] = int(parts[1])
                table.setdefault(name, {})[cfg] = (float(parts[3]), parts[2].strip())
            elif "failed" in ln:
                pass
    names = [l[0] for l in LAYERS if l[0] in table]
    cols = [c for c in ["auto"] + ["%s/%s" % c for c in cfgs[1:]]]
    print("batch %d; us per launch in a 20-launch PDL chain (graph replay); columns BN/split" % batch)
    print("%-24s %2s " % ("layer", "x") + " ".join("%7s" % c for c in cols) + "   best")
    tot_auto = tot_best = 0.0
    for nm in names:
        row = table[nm]
        # forced settings the plan could not honour fall back silently: only trust distinct realised plans
        best_cfg = min((c for c in cols if c in row), key=lambda c: row[c][0])
        print("%-24s %2d " % (nm, counts[nm]) + " ".join("%7.2f" % row[c][0] if c in row else "%7s" % "-" for c in cols) +
              "   %s (%s)" % (best_cfg, row[best_cfg][1]))
        tot_auto += counts[nm] * row["auto"][0]
        tot_best += counts[nm] * row[best_cfg][0]
    print("sum over the net: heuristic %.1f us, per-layer best %.1f us (%.1f us left)" % (tot_auto, tot_best, tot_auto - tot_best))
    print("heuristic plans:")
    for nm in names:
        print("  %-24s %s" % (nm, table[nm]["auto"][1]))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        one(int(sys.argv[2]) if len(sys.argv) > 2 else 8)
    else:
        main(int(sys.argv[1]) if len(sys.argv) > 1 else 8)
