"""Turn gpurun_out/*.ncu-rep and launch lists into the small tracked summaries under profiles/.
usage: python profiles/summarize.py <prof.ncu-rep> <launches.csv> <tag>"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]


def short(name):
    name = name.replace("void ", "").replace("nsvd::", "")
    return name.split("(")[0][:70]


def main(rep, launches, tag):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    seen = set()
    with open(f"profiles/{tag}_ncu_full.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + [f"{k} [{units[idx[k]]}]" for k in KEYS if k in idx])
        for r in rows[2:]:
            k = short(r[idx["Kernel Name"]])
            if k in seen:
                continue
            seen.add(k)
            w.writerow([k] + [r[idx[m]] for m in KEYS if m in idx])
    # launch list: aggregate per kernel
    agg, cnt = collections.OrderedDict(), collections.Counter()
    text = open(launches).read()
    start = text.index('"ID"')
    rd = csv.DictReader(io.StringIO(text[start:]))
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = short(r["Kernel Name"])
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v = v / 1000.0 if u in ("ns", "nsecond") else (v * 1000.0 if u in ("ms", "msecond") else v)   # -> us
        agg[k] = agg.get(k, 0.0) + v
        cnt[k] += 1
    tot = sum(agg.values())
    with open(f"profiles/{tag}_launches.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_us", "share_pct"])
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
            w.writerow([k, cnt[k], f"{v:.1f}", f"{100 * v / tot:.2f}"])
    print(open(f"profiles/{tag}_launches.csv").read())
    print(open(f"profiles/{tag}_ncu_full.csv").read())


if __name__ == "__main__":
    main(*sys.argv[1:4])
