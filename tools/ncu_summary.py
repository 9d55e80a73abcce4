"""ncu report -> compact per-kernel summary CSV (metric,unit,value per captured launch) for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [out.csv]
Reads `ncu -i <rep> --page raw --csv` and keeps the counters the roofline discussion needs."""
import csv
import io
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__inst_executed.avg.per_cycle_active", "sm__cycles_elapsed.avg", "sm__cycles_active.avg"]


def main():
    rep = sys.argv[1]
    out = sys.argv[2] if len(sys.argv) > 2 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, check=True).stdout.decode()
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = ["metric,unit," + ",".join("launch{}".format(i) for i in range(len(rows) - 2))]
    name_i = hdr.index("Kernel Name")
    lines.append("Kernel Name,," + ",".join('"{}"'.format(r[name_i]) for r in rows[2:]))
    for k in KEEP:
        cols = [i for i, h in enumerate(hdr) if h == k or h.endswith("." + k)]
        if not cols:
            continue
        i = cols[0]
        lines.append("{},{},{}".format(k, units[i], ",".join(r[i] for r in rows[2:])))
    text = "\n".join(lines) + "\n"
    if out:
        open(out, "w").write(text)
    sys.stdout.write(text)


if __name__ == "__main__":
    main()
