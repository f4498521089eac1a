"""Summarise an `ncu --set full` report: per captured launch, the metrics the DESIGN / bench roofline cite.

    python profiles/summarize_ncu.py gpurun_out/r02q_prof.ncu-rep > profiles/r02q_ncu_full.md
    (also rewrites profiles/ncu_traffic.json when given --traffic <name>)
"""
import csv
import io
import json
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
           "launch__block_size", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active"]
CLASS = {"gru_persist_fwd": "gru_persistent_fwd", "gru_persist_bwd": "gru_persistent_bwd",
         "dec_persist_fwd": "decoder_persistent_fwd", "dec_persist_bwd": "decoder_persistent_bwd"}


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def main(rep, traffic_name=None):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(head)}
    traffic = {}
    print("# ncu --set full captures (%s)\n" % rep)
    for r in body:
        name = r[col["Kernel Name"]]
        print("## `%s`\n\n| metric | value |\n|---|---|" % name)
        for m in METRICS:
            if m in col:
                print("| `%s` | %s %s |" % (m, r[col[m]], units[col[m]]))
        print()
        rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
        wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        for k, c in CLASS.items():
            if k in name and c not in traffic:
                traffic[c] = rd + wr
    if traffic_name:
        json.dump({"source": "profiles/%s (ncu --set full --clock-control none; dram__bytes_read.sum + dram__bytes_write.sum per "
                             "launch; report %s)" % (traffic_name, rep), "kernels": traffic},
                  open("profiles/ncu_traffic.json", "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[3] if len(sys.argv) > 3 and sys.argv[2] == "--traffic" else None)
