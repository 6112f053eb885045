"""Turn ncu artefacts from gpurun_out/ into the small text summaries committed under profiles/.
  python tools/summarize_ncu.py launches <csv> > profiles/xxx.txt      (per-kernel time shares of a launch list)
  python tools/summarize_ncu.py full <ncu-rep> > profiles/yyy.txt       (key metrics of a --set full capture)"""
import collections
import csv
import subprocess
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hi]
    ci = {n: i for i, n in enumerate(h)}
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) < len(h) or r[ci["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[ci["Kernel Name"]].split("(")[0][:44] + " grid=" + r[ci["Grid Size"]] + " blk=" + r[ci["Block Size"]]
        v = float(r[ci["Metric Value"]])
        if r[ci["Metric Unit"]] == "ns":
            v /= 1000
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# source: {path}\n# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)")
    print(f"# total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:78s} n={v[0]:5d} total={v[1]:9.1f}us avg={v[1]/v[0]:7.2f}us share={v[1]/tot*100:5.1f}%")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    print(f"# source: {path}\n# ncu --set full --clock-control none --import-source on")
    for r in rows[2:]:
        print("--- kernel", r[h.index("Kernel Name")][:60])
        for w in WANT:
            if w in h:
                print(f"{w:86s} {r[h.index(w)]:>16s} {units[h.index(w)]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
