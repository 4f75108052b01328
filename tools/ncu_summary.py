"""Summarise an .ncu-rep (read here, no GPU needed) into the per-kernel table committed under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.csv
"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct_of_ncu_peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_pct_active"),
    ("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "dmma_pipe_pct_active"),
    ("sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_elapsed", "fp64_shared_pipe_pct_elapsed"),
    ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor_pipe_pct_elapsed"),
    ("sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active", "imma_inst_pct"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "tensor_core_smem_reads_pct"),
    ("l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "tma_load_bytes"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("smsp__inst_executed.sum", "warp_instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math_pipe"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {}
    for i, h in enumerate(hdr):          # some metrics appear twice (a triage copy with a section prefix): keep the plain name
        idx.setdefault(h.split(".", 2)[-1] if h.startswith(("TPC.", "SM_A.", "LTS.")) else h, i)
        idx[h] = i
    out = csv.writer(sys.stdout)
    out.writerow(["kernel"] + [f"{short} [{units[idx[m]]}]" if m in idx and units[idx[m]] else short for m, short in METRICS])
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].replace("void <unnamed>::", "").split("(")[0]
        out.writerow([name] + [r[idx[m]] if m in idx else "" for m, _ in METRICS])


if __name__ == "__main__":
    main(sys.argv[1])
