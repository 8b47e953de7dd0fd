"""Generate the committed fixtures under tests/golden/ with the CPU oracle (run HERE, in the
build container; the GPU box only reads the results):

  <model>_calib.json   max-abs calibration table {node: scale} from the fp32 oracle over 8
                       synthetic images (seeds 1000..1007), CalibrationAlgoType::MAXABS
  <model>_golden.npz   oracle outputs for the bench / parity inputs (seed 42+i): fp32 logits
                       + probabilities, int8 logits + probabilities, top-1 indices

  python tools/make_golden.py [tiny_resnet resnet50 resnet101 vgg16 mobilenet_v1]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from anakin_b200 import modelzoo as Z  # noqa: E402
from oracle import model_walker as W  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
HW = {"tiny_resnet": 32}
LOGITS = {"tiny_resnet": "fc", "resnet50": "fc1000", "resnet101": "fc1000", "vgg16": "fc8", "mobilenet_v1": "fc7"}
INT8_MODELS = {"tiny_resnet", "resnet50", "resnet101"}


def main(models):
    os.makedirs(GOLD, exist_ok=True)
    for name in models:
        hw = HW.get(name, 224)
        nb = 4 if name != "vgg16" else 2
        g = Z.BUILDERS[name](batch=1)
        t0 = time.time()
        x = Z.synthetic_input(nb, hw, seed=42)
        out = {}
        fp32 = W.run_fp32(g, x)
        out["prob_fp32"] = fp32["prob_out"].astype(np.float32)
        out["top1_fp32"] = fp32["prob_out"].argmax(1).astype(np.int32)
        if name in INT8_MODELS:
            cal_x = Z.synthetic_input(8, hw, seed=1000)
            scales = W.calibrate(g, cal_x)
            with open(os.path.join(GOLD, "%s_calib.json" % name), "w") as f:
                json.dump({"model": name, "algo": "maxabs", "images": 8, "seed": 1000,
                           "edge_scales": {k: float(np.float32(v)) for k, v in scales.items()}}, f, indent=0)
            scales32 = {k: float(np.float32(v)) for k, v in scales.items()}
            q, trace = W.run_int8(g, x, scales32, return_intermediate=True)
            out["prob_int8"] = q["prob_out"].astype(np.float32)
            out["top1_int8"] = q["prob_out"].argmax(1).astype(np.int32)
            lg = trace[LOGITS[name]][0]
            out["logits_int8"] = lg.reshape(nb, -1).astype(np.float32)
        np.savez_compressed(os.path.join(GOLD, "%s_golden.npz" % name), **out)
        print("%s: %.1fs top1 fp32 %s%s" % (name, time.time() - t0, out["top1_fp32"],
                                           (" int8 %s" % out["top1_int8"]) if "top1_int8" in out else ""))


if __name__ == "__main__":
    main(sys.argv[1:] or ["tiny_resnet", "resnet50", "resnet101", "vgg16", "mobilenet_v1"])
