"""Text summary of an .ncu-rep (run where ncu is installed; no GPU needed):
    python scripts/ncu_summary.py gpurun_out/full_cold.ncu-rep [more.ncu-rep ...] > profiles/rN_ncu_full_summary.txt
One block per profiled launch with the metrics the roofline discussion uses."""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]


def summarise(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"==== {path} ====")
    for r in rows[2:]:
        print(r[col["Kernel Name"]][:150])
        for m in METRICS:
            if m in col and r[col[m]] != "":
                print(f"    {m} [{units[col[m]]}] = {r[col[m]]}")


if __name__ == "__main__":
    for p in sys.argv[1:]:
        summarise(p)
