"""Tile / split-K tuning experiment: times single conv layers (CUDA-graph replay of 20 launches, so
host launch cost is excluded) under forced BN / split settings.
  python tools/tile_tune.py            # runs the grid in subprocesses
  python tools/tile_tune.py one        # one configuration from the environment
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# name: n,h,w,c,k,r,stride,pad, residual
LAYERS = {
    "res5_2b 3x3 512 7x7 b8": (8, 7, 7, 512, 512, 3, 1, 1, False),
    "res4_2b 3x3 256 14x14 b8": (8, 14, 14, 256, 256, 3, 1, 1, False),
    "res5_2a 1x1 2048>512 b8": (8, 7, 7, 2048, 512, 1, 1, 0, False),
    "res4_2c 1x1 256>1024 b8": (8, 14, 14, 256, 1024, 1, 1, 0, True),
    "res2_2c 1x1 64>256 56 b8": (8, 56, 56, 64, 256, 1, 1, 0, True),
    "res5_2b b1": (1, 7, 7, 512, 512, 3, 1, 1, False),
    "fc 2048>1000 b8": (8, 1, 1, 2048, 1000, 1, 1, 0, False),
}


def one():
    import torch
    from anakin_b200 import saber_abi as A
    from gpu_util import ConvRunner, dev
    rng = np.random.default_rng(0)
    for name, (n, h, w, c, k, r, stride, pad, has_res) in LAYERS.items():
        x = rng.integers(0, 255, (n, h, w, c)).astype(np.uint8)
        wt = rng.integers(-127, 128, (k, c, r, r)).astype(np.int8)
        fc = name.startswith("fc")
        odt = A.FLOAT if fc else A.UINT8
        ldc = k if fc else (k + 15) // 16 * 16
        try:
            run = ConvRunner(A.MATH_I8, x.shape, A.UINT8, wt, np.zeros(k, np.float32), np.full(k, 1e-4, np.float32),
                             odt, res_dtype=(A.UINT8 if has_res else -1), stride=(stride, stride), pad=(pad, pad),
                             relu=not fc, ldc=ldc, sum_scale=0.5)
        except Exception as e:
            print("%-26s plan failed: %s" % (name, e))
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
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        info = run.info()
        print("%-26s BN=%-3d grid=%dx%d smem=%-6d %7.2f us" % (name, info["block_n"], info["grid_x"], info["grid_y"],
                                                                info["smem"], e0.elapsed_time(e1) / 100 * 1e3))


def main():
    for bn in ("", "32", "64", "128"):
        for split in ("1", "2", "4"):
            env = dict(os.environ)
            if bn:
                env["B200_SABER_FORCE_BN"] = bn
            env["B200_SABER_FORCE_SPLIT"] = split
            print("== FORCE_BN=%s FORCE_SPLIT=%s" % (bn or "auto", split), flush=True)
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "one"], env=env, stdout=subprocess.PIPE,
                               stderr=subprocess.STDOUT, text=True, timeout=300)
            print(r.stdout, end="", flush=True)


if __name__ == "__main__":
    one() if len(sys.argv) > 1 else main()
