"""Repeatability probe of the Worker e2e path: several timed bursts per thread count.
  python tools/worker_probe.py [--batch 8] [--threads 2,4,6,8] [--bursts 5] [--requests 300]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--threads", default="2,4,6,8")
    ap.add_argument("--bursts", type=int, default=5)
    ap.add_argument("--requests", type=int, default=300)
    a = ap.parse_args()
    import torch
    from anakin_b200 import anakin_bin, api, modelzoo
    print("host cpus:", os.cpu_count(), "affinity:", len(os.sched_getaffinity(0)))
    blob = anakin_bin.dumps(modelzoo.build("resnet50", batch=a.batch, precision="int8"))
    d = os.path.join(ROOT, ".bench_tmp")
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, "probe.anakin.bin")
    open(path, "wb").write(blob)
    x = modelzoo.synthetic_input(a.batch, 224)
    for T in [int(t) for t in a.threads.split(",")]:
        W = api.Worker(path, "int8", threads=T, devices=[0], batch=a.batch)
        W.wait_ready()
        depth = 2 * T
        xin = [torch.from_numpy(x).pin_memory() for _ in range(depth)]
        xout = [torch.empty(a.batch * 1000, dtype=torch.float32).pin_memory() for _ in range(depth)]

        def serve(n):
            inflight = 0
            for i in range(n):
                if inflight == depth:
                    W.async_get_result()
                    inflight -= 1
                j = i % depth
                W.async_prediction_ptr(xin[j].data_ptr(), xin[j].numel(), xout[j].data_ptr(), xout[j].numel())
                inflight += 1
            while inflight:
                W.async_get_result()
                inflight -= 1

        serve(6 * T)
        res = []
        for _ in range(a.bursts):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            serve(a.requests)
            dt = time.perf_counter() - t0
            res.append(a.requests * a.batch / dt)
        print("T=%d  img/s per burst: %s   (us/request: %s)" % (
            T, " ".join("%7.0f" % r for r in res), " ".join("%6.1f" % (a.batch / r * 1e6) for r in res)), flush=True)
        del W


if __name__ == "__main__":
    main()
