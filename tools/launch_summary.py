"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: share of total time per kernel name."""
import collections, csv, re, sys

def main(path, title=""):
    rows = [r for r in csv.reader(l for l in open(path, errors="replace") if not l.startswith("=="))]
    hdr = rows[0]; ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if len(r) <= iv or r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
            continue
        us = float(r[iv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(r[iu], 1e-3)
        name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").strip()[:70]
        agg[name][0] += 1; agg[name][1] += us
    tot = sum(v[1] for v in agg.values()); n = sum(v[0] for v in agg.values())
    if title:
        print(title)
    print(f"total {tot:.1f} us over {n} launches")
    for name, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
        print(f"{100 * us / tot:5.1f}%  n={c:4d}  avg {us / c:8.1f} us  {name}")

if __name__ == "__main__":
    main(sys.argv[1], " ".join(sys.argv[2:]))
