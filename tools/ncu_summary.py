"""Summarise an .ncu-rep (read on the CPU box with `ncu -i`) into the handful of metrics the roofline needs.
usage: python tools/ncu_summary.py gpurun_out/prof_gemm.ncu-rep [more.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum",
    "gpc__cycles_elapsed.avg.per_second",
    "sm__cycles_active.avg",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "lts__t_bytes.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "launch__shared_mem_per_block_dynamic",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
]


def main():
    for path in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        print(f"== {path}")
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            print(f"-- {name[:100]}")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    print(f"   {k:95s} {r[i]:>18s} {units[i]}")
            extra = [h for h in hdr if ("pipe_tensor" in h and "pct" in h) or "ops_path_tensor" in h and h.endswith("pct_of_peak_sustained_elapsed")]
            for k in extra:
                i = hdr.index(k)
                if r[i] not in ("0", "0.00", ""):
                    print(f"   {k:95s} {r[i]:>18s} {units[i]}")


if __name__ == "__main__":
    main()
