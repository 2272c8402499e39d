"""Extract the headline counters of every launch in an .ncu-rep (`ncu --set full`) into JSON.
   python tools/ncu_metrics.py gpurun_out/x.ncu-rep > profiles/x_metrics.json"""
import csv, io, json, subprocess, sys
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.max", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_bytes.sum", "smsp__inst_executed.sum",
        "launch__shared_mem_per_block_dynamic"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
res = []
for r in rows[2:]:
    d = dict(zip(hdr, r))
    e = {"kernel": d.get("Kernel Name", "").split("(")[0].replace("vlg::", "")}
    for k in WANT:
        if k in d:
            e[k] = f"{d[k]} {units[hdr.index(k)]}".strip()
    res.append(e)
print(json.dumps(res, indent=1))
