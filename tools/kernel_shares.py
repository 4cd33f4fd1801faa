"""Per-kernel share of one steady-state epoch from an ncu launch list (gpu__time_duration.sum, --csv).

    python tools/kernel_shares.py profiles/r02_launches_bench.csv [epochs_from_the_end=2] > profiles/r02_epoch_kernel_shares.txt

An epoch = the launches between two consecutive occurrences of the training step's first kernel; the last complete
epochs of the list are averaged (they are CUDA-graph replays of the steady state)."""
import csv
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    n_ep = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    head = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[head]
    ki, vi, mi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name")
    ui = h.index("Metric Unit")
    launches = []
    for r in rows[head + 1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(r[ui], 1.0)
        name = r[ki].split("(")[0].replace("tgcn::", "").replace("void ", "")
        launches.append((name, v))
    # an epoch ends with the masked NLL reductions of the eval pass; find the period as the distance between the last
    # two launches of the fused-Adam SpMM epilogue kernel (once per train step)
    marks = [i for i, (n, _) in enumerate(launches) if "k_adam_prepare" in n or "k_increment" in n]
    if len(marks) < n_ep + 1:
        print("not enough epochs in the list", len(marks))
        return
    lo, hi = marks[-(n_ep + 1)], marks[-1]
    per = OrderedDict()
    for n, v in launches[lo:hi]:
        c, t = per.get(n, (0, 0.0))
        per[n] = (c + 1, t + v)
    tot = sum(t for _, t in per.values())
    print(f"# {path}: {n_ep} steady-state epochs, launches {lo}..{hi}, {tot / n_ep:.1f} us of kernel time per epoch "
          f"(serialised, cold caches under ncu: compare SHARES)")
    print(f"{'kernel':60s} {'launches/epoch':>14s} {'us/epoch':>10s} {'share':>7s}")
    for n, (c, t) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        print(f"{n[:60]:60s} {c / n_ep:14.1f} {t / n_ep:10.1f} {100 * t / tot:6.1f}%")


if __name__ == "__main__":
    main()
