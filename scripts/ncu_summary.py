"""Key counters of an ncu report as text: `python scripts/ncu_summary.py report.ncu-rep [kernel-regex]`
(run where ncu is installed; the report comes from `ncu --set full --clock-control none --import-source on ...`)."""
import csv, io, re, subprocess, sys

KEYS = [
    "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "launch__block_size", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "lts__t_sector_hit_rate.pct", "sm__icc_request_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
]
EXTRA = ["sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
         "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
         "sm__ops_path_tensor_op_hmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main():
    rep = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    name_col = hdr.index("Kernel Name")
    n = 0
    for r in rows[2:]:
        if pat and not pat.search(r[name_col]):
            continue
        n += 1
        print(f"--- {r[name_col].split('(')[0]} launch {n}")
        vals = {}
        for k in KEYS + EXTRA:
            cols = [c for c in hdr if c == k or c.endswith("." + k)]
            if cols:
                i = hdr.index(cols[0])
                vals[k] = (r[i], units[i])
                print(f"{k:90s} {r[i]} {units[i]}")
        try:
            rd, wr = vals["dram__bytes_read.sum"], vals["dram__bytes_write.sum"]
            tot = float(rd[0].replace(",", "")) * UNIT.get(rd[1], 1.0) + float(wr[0].replace(",", "")) * UNIT.get(wr[1], 1.0)
            print(f"{'derived: dram bytes per launch':90s} {tot / 1e9:.3f} GB")
            bc = float(vals["l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"][0].replace(",", ""))
            wf = float(vals["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"][0].replace(",", ""))
            if wf:
                print(f"{'derived: shared-memory bank-conflict wavefronts':90s} {100 * bc / wf:.1f} %")
        except (KeyError, ValueError):
            pass


if __name__ == "__main__":
    main()
