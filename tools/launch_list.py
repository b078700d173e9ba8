"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` log into a per-kernel table (markdown on stdout)."""
import csv
import re
import sys
from collections import OrderedDict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rd:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        unit = r[iu]
        ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v
        name = re.sub(r"^void ", "", r[ik])
        name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name)
        name = re.sub(r"\(.*$", "", name)
        rows.append((name, ms))
    agg = OrderedDict()
    for n, ms in rows:
        c, t = agg.get(n, (0, 0.0))
        agg[n] = (c + 1, t + ms)
    step = {k: v for k, v in agg.items() if "peak_kernel" not in k}
    tot = sum(t for _, t in step.values())
    print("| kernel | launches | total ms | share of the step kernels |")
    print("|---|---|---|---|")
    for k, (c, t) in sorted(step.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %d | %.3f | %.1f %% |" % (k, c, t, 100 * t / tot))
    print("| **all step kernels** | %d | %.2f | 100 %% |" % (sum(c for c, _ in step.values()), tot))
    pk = [(c, t) for k, (c, t) in agg.items() if "peak_kernel" in k]
    if pk:
        print("\nFP64-peak micro-benchmark kernels (measurement only, excluded): %.1f ms in %d launches." % (sum(t for _, t in pk), sum(c for c, _ in pk)))


if __name__ == "__main__":
    main(sys.argv[1])
