"""Summarise an .ncu-rep (ncu --set full) into the JSON kept under profiles/: python profiles/summarize_ncu.py rep out.json"""
import csv, io, json, subprocess, sys

KEYS = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__cycles_elapsed.max", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "lts__t_sector_hit_rate.pct"]


def main(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        e = {"kernel": d.get("Kernel Name", "")}
        for k in KEYS:
            if k in d:
                e[k] = f"{d[k]} {u[k]}".strip()
        st = {}
        for k, v in d.items():
            if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
                try:
                    st[k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = round(float(v.replace(",", "")), 4)
                except ValueError:
                    pass
        e["stalls_per_issue"] = dict(sorted(st.items(), key=lambda kv: -kv[1])[:10])
        res.append(e)
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
