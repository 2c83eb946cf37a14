"""Turn `ncu --set full` reports into the tracked summaries under profiles/.

    python tests/tools/ncu_summary.py <report.ncu-rep> <out.csv> [--json-key conv_traffic|gs_traffic --pick N]

Writes one CSV row per profiled launch with the counters the roofline discussion uses, and (optionally)
records dram read+write bytes of launch N in profiles/ncu_summary.json under the given key (bench.py copies
that into `roofline.traffic`)."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
COLS = [
    ("kernel", "Kernel Name"), ("grid", "Grid Size"), ("block", "Block Size"),
    ("time_us", "gpu__time_duration.sum"),
    ("dram_read", "dram__bytes_read.sum"), ("dram_write", "dram__bytes_write.sum"),
    ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l2_to_sm_bytes", "l1tex__m_xbar2l1tex_read_bytes.sum"),
    ("l2_to_sm_per_s", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second"),
    ("lts_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l1tex_pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor_pipe_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("tensor_pipe_pct_rt", "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("regs", "launch__registers_per_thread"), ("smem_dyn", "launch__shared_mem_per_block_dynamic"),
    ("occ_limit_smem", "launch__occupancy_limit_shared_mem"),
]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    recs = []
    with open(out, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow([c for c, m in COLS if m in idx] + ["dram_total_bytes"])
        for r in rows[2:]:
            line = []
            for c, m in COLS:
                if m not in idx:
                    continue
                v = r[idx[m]]
                if c == "kernel":
                    v = v.split("(")[0].replace("void ", "").replace("<unnamed>::", "")
                elif units[idx[m]]:
                    v = v + " " + units[idx[m]]
                line.append(v)
            tot = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + \
                to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
            line.append("%d" % tot)
            recs.append(tot)
            w.writerow(line)
    print("wrote", out, len(recs), "launches")
    if "--json-key" in sys.argv:
        key = sys.argv[sys.argv.index("--json-key") + 1]
        pick = int(sys.argv[sys.argv.index("--pick") + 1]) if "--pick" in sys.argv else 0
        path = os.path.join(ROOT, "profiles", "ncu_summary.json")
        d = json.load(open(path)) if os.path.exists(path) else {}
        d[key] = recs[pick]
        d[key + "_source"] = "%s launch %d (dram__bytes_read.sum + dram__bytes_write.sum)" % (os.path.basename(out), pick)
        json.dump(d, open(path, "w"), indent=1)
        print("recorded", key, recs[pick])


if __name__ == "__main__":
    main()
