"""Side-by-side table of several tools/profile_net.py outputs (one column of per-op microseconds per file), for in-net
A/B runs of one model under different environment switches on ONE box.
  python tools/perop_compare.py a.txt b.txt ..."""
import os
import sys


def read(path):
    rows, total = [], None
    for line in open(path):
        parts = line.split()
        if len(parts) >= 3 and parts[0] not in ("node", "total"):
            try:
                rows.append((parts[0], float(parts[2])))
            except ValueError:
                pass
        if line.startswith("total eager op time"):
            total = float(parts[4])
    return rows, total


def main(paths):
    cols = [read(p) for p in paths]
    names = [os.path.splitext(os.path.basename(p))[0][-14:] for p in paths]
    print("%-24s" % "node" + "".join("%15s" % n for n in names))
    base = cols[0][0]
    for i, (node, _) in enumerate(base):
        vals = []
        for rows, _ in cols:
            d = dict(rows)
            vals.append("%15.2f" % d[node] if node in d else "%15s" % "-")
        print("%-24s" % node + "".join(vals))
    print("%-24s" % "total" + "".join("%15.1f" % (t if t is not None else float("nan")) for _, t in cols))


if __name__ == "__main__":
    main(sys.argv[1:])
