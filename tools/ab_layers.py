"""A/B of the conv kernel variants on single layers (CUDA events around back-to-back launches through the C ABI):
per-tile im2col kernel vs persistent tile-pipelined kernel vs slab kernel.  python tools/ab_layers.py [batch]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from anakin_b200 import saber_abi as A  # noqa: E402
from gpu_util import ConvRunner, dev  # noqa: E402

# name, h, c, k, r, stride, pad, residual
LAYERS = [
    ("s2 2a 1x1 256>64", 56, 256, 64, 1, 1, 0, False), ("s2 2b 3x3 64", 56, 64, 64, 3, 1, 1, False),
    ("s2 2c 1x1 64>256 +res", 56, 64, 256, 1, 1, 0, True),
    ("s3 2a 1x1 512>128", 28, 512, 128, 1, 1, 0, False), ("s3 2b 3x3 128", 28, 128, 128, 3, 1, 1, False),
    ("s3 2c 1x1 128>512 +res", 28, 128, 512, 1, 1, 0, True),
    ("s4 2a 1x1 1024>256", 14, 1024, 256, 1, 1, 0, False), ("s4 2b 3x3 256", 14, 256, 256, 3, 1, 1, False),
    ("s4 2c 1x1 256>1024 +res", 14, 256, 1024, 1, 1, 0, True),
    ("s5 2a 1x1 2048>512", 7, 2048, 512, 1, 1, 0, False), ("s5 2b 3x3 512", 7, 512, 512, 3, 1, 1, False),
    ("s5 2c 1x1 512>2048 +res", 7, 512, 2048, 1, 1, 0, True),
]
VARIANTS = [("per-tile", {"B200_SABER_PERSISTENT": "0", "B200_SABER_SLAB": "0"}),
            ("persistent", {"B200_SABER_PERSISTENT": "2", "B200_SABER_SLAB": "0"}),
            ("slab", {"B200_SABER_PERSISTENT": "0", "B200_SABER_SLAB": "2"}),
            ("auto", {})]


def time_layer(n, h, c, k, r, stride, pad, with_res, env):
    for key in ("B200_SABER_PERSISTENT", "B200_SABER_SLAB"):
        os.environ.pop(key, None)
    os.environ.update(env)
    rng = np.random.default_rng(0)
    x = rng.integers(0, 256, (n, h, h, c)).astype(np.uint8)
    wt = rng.integers(-127, 128, (k, c, r, r)).astype(np.int8)
    run = ConvRunner(A.MATH_I8, x.shape, A.UINT8, wt, np.zeros(k, np.float32), np.full(k, 1e-4, np.float32), A.UINT8,
                     res_dtype=(A.UINT8 if with_res else -1), stride=(stride, stride), pad=(pad, pad), relu=True)
    info = run.info()
    xd = dev(x)
    res = dev(rng.integers(0, 256, (n, run.ho, run.wo, k)).astype(np.uint8)) if with_res else None
    out = run.run(xd, res)
    for _ in range(5):
        run.run(xd, res, out_dev=out)
    torch.cuda.synchronize()
    reps = 40
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run.run(xd, res, out_dev=out)
    e1.record()
    torch.cuda.synchronize()
    tag = "slab" if info["slab"] else ("pers" if info["persistent"] else "tile")
    return e0.elapsed_time(e1) / reps * 1e3, "%s/BN%d/%dx%d" % (tag, info["block_n"], info["grid_x"], info["grid_y"])


if __name__ == "__main__":
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    print("batch %d; us per launch in a back-to-back PDL chain" % batch)
    print("%-26s " % "layer" + " ".join("%-28s" % v for v, _ in VARIANTS))
    for name, h, c, k, r, stride, pad, wr in LAYERS:
        cells = []
        for vname, env in VARIANTS:
            try:
                us, what = time_layer(batch, h, c, k, r, stride, pad, wr, env)
                cells.append("%7.2f %-20s" % (us, what))
            except Exception as e:   # a variant that does not apply to the layer
                cells.append("%-28s" % ("n/a " + str(e)[:20]))
        print("%-26s " % name + " ".join(cells))
