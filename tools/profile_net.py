"""Per-op device-time table of a model (eager, CUDA-event pair per op) -- the reference's
ENABLE_OP_TIMER view (framework/core/net/net.cpp:445-449,494-506).
  python tools/profile_net.py --model resnet50 --precision int8 --batch 8
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="resnet50")
    ap.add_argument("--precision", default="int8")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    from anakin_b200 import anakin_bin, api, modelzoo
    hw = 32 if a.model == "tiny_resnet" else 224
    g = modelzoo.build(a.model, batch=a.batch, precision=a.precision if a.precision == "int8" else "fp32")
    G = api.Graph.from_bytes(anakin_bin.dumps(g))
    G.ResetBatchSize("input_0", a.batch)
    G.Optimize()
    net = api.Net(G, a.precision)
    net.set_input("input_0", modelzoo.synthetic_input(a.batch, hw))
    net.prediction()
    net.sync()
    prof = net.profile_ops(a.iters, a.reps)
    nodes = {n[0]: n for n in G.describe()}
    total = sum(ms for _, _, ms in prof)
    print("%-22s %-26s %9s  %-18s %-18s" % ("node", "op", "us", "in", "out"))
    for name, op, ms in prof:
        o = net.tensor_info(name)
        ins = nodes[name][2]
        i = net.tensor_info(ins[0]) if ins else None
        print("%-22s %-26s %9.2f  %-18s %-18s" % (name, op, ms * 1e3, i["dims"] if i else "", o["dims"]))
    print("total eager op time %.1f us over %d ops; activations %.1f MB" %
          (total * 1e3, len(prof), net.activation_bytes() / 1e6))


if __name__ == "__main__":
    main()
