"""Summarise an `ncu --csv` launch list (gpu__time_duration.sum [+ dram__bytes_read/write.sum]) per kernel name:
    python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches_summary.txt
Also writes <csv>.json with the per-launch averages of the tcgen05 GEMM family (read by bench.py for roofline.traffic)."""
import collections
import csv
import io
import json
import re
import sys


def unit_scale(unit):
    return {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6,
            "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}.get(unit, 1.0)


def main(path, out_json=None):
    lines = open(path).read().splitlines()
    hi = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    per = collections.OrderedDict()
    for r in csv.DictReader(io.StringIO("\n".join(lines[hi:]))):
        k = per.setdefault(r["ID"], {"name": r["Kernel Name"]})
        k[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * unit_scale(r["Metric Unit"])
    agg = collections.OrderedDict()
    for k in per.values():
        name = re.sub(r"\(.*", "", k["name"]).replace("void ", "")[:80]
        a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += k.get("gpu__time_duration.sum", 0.0)
        a[2] += k.get("dram__bytes_read.sum", 0.0)
        a[3] += k.get("dram__bytes_write.sum", 0.0)
    tot = sum(a[1] for a in agg.values())
    print("total %.1f us over %d launches (per-launch times are cold-cache and serialised: compare SHARES)" % (tot, len(per)))
    print("%10s %6s %6s %9s %12s %12s  kernel" % ("time us", "share", "n", "avg us", "dram rd MB/l", "dram wr MB/l"))
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%10.1f %5.1f%% %6d %9.2f %12.3f %12.3f  %s" % (a[1], 100 * a[1] / tot, a[0], a[1] / a[0], a[2] / a[0] / 1e6,
                                                            a[3] / a[0] / 1e6, name))
    fam = [a for n, a in agg.items() if "gemm_tc_kernel" in n or "vistok_kernel" in n or "vistok_pg_kernel" in n]
    n = sum(a[0] for a in fam)
    info = {"family": "gemm_tc_kernel + vistok_kernel + vistok_pg_kernel", "launches": n, "share_of_kernel_time": sum(a[1] for a in fam) / tot,
            "avg_us_per_launch": sum(a[1] for a in fam) / max(n, 1),
            "dram_bytes_per_launch": (sum(a[2] for a in fam) + sum(a[3] for a in fam)) / max(n, 1),
            "source": path}
    print("\ntcgen05 GEMM family: %d launches, %.1f%% of the kernel time, %.2f us and %.3f MB of DRAM traffic per launch"
          % (n, 100 * info["share_of_kernel_time"], info["avg_us_per_launch"], info["dram_bytes_per_launch"] / 1e6))
    if out_json:
        json.dump(info, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
