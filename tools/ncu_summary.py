"""Print the ncu metrics that matter for the roofline discussion from a .ncu-rep (one block per launch).
Usage: python tools/ncu_summary.py file.ncu-rep [--last-of-each] [...]"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % (occupancy)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    # pipes
    ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active %"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "tensor hmma sub-pipe active cycles (avg/SM)"),
    ("sm__ops_path_tensor_op_hmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed", "tensor ops fp16->fp32 % of peak"),
    ("sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed", "tensor ops bf16->fp32 % of peak"),
    ("sm__ops_path_tensor_op_hmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed", "tensor ops tf32->fp32 % of peak"),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor memory (TMEM) cycles active %"),
    ("sm__cycles_elapsed.avg", "SM cycles elapsed (avg)"),
    ("sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", "tensor (tc) pipe inst %"),
    ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "tmem pipe inst %"),
    ("sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active", "tma pipe inst %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma pipe cycles %"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "fma-heavy pipe cycles %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu pipe cycles %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu (MUFU) pipe inst %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu pipe inst %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe cycles %"),
    # L1
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "L1 global-load requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "L1 global-load sectors"),
    ("l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum", "L1 global-load wavefronts"),
    ("l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "L1 global-load hit rate %"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "L1 global-store requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "L1 global-store sectors"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 LSU data-pipe wavefronts %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem LSU wavefronts %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem->TC wavefronts %"),
    ("l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "TMA load bytes"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed", "xbar->SM read %"),
    # L2 / DRAM
    ("lts__t_sector_hit_rate.pct", "L2 sector hit rate %"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "L2 sectors read by SMs"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
]
STALL_PREFIX, STALL_SUFFIX = "smsp__average_warps_issue_stalled_", "_per_issue_active.ratio"


def fnum(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    last_only = "--last-of-each" in sys.argv
    for path in args:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        body = rows[2:]
        if last_only:
            seen = {}
            for vals in body:
                seen[vals[hdr.index("Kernel Name")]] = vals
            body = list(seen.values())
        for vals in body:
            name = vals[hdr.index("Kernel Name")][:90]
            print(f"== {path.split('/')[-1]}: {name}")
            got = {}
            for key, label in KEYS:
                # some metrics carry a section prefix ("TPC.TriageCompute.<metric>"): match on the suffix
                cand = [h for h in hdr if h == key or h.endswith("." + key)]
                if cand:
                    i = hdr.index(cand[0])
                    got[key] = fnum(vals[i])
                    print(f"   {label:32s} {vals[i]:>18s} {units[i]}")
            rq, sc = got.get("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"), got.get(
                "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum")
            wf = got.get("l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum")
            if rq and sc:
                print(f"   {'-> sectors / load request':32s} {sc / rq:18.2f}")
            if rq and wf:
                print(f"   {'-> wavefronts / load request':32s} {wf / rq:18.2f}")
            stalls = []
            for i, h in enumerate(hdr):
                if h.startswith(STALL_PREFIX) and h.endswith(STALL_SUFFIX) and "not_issued" not in h:
                    v = fnum(vals[i])
                    if v is not None and v > 0.2:
                        stalls.append((v, h[len(STALL_PREFIX):-len(STALL_SUFFIX)]))
            if stalls:
                print("   stalls (warps per issue): " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)))


if __name__ == "__main__":
    main()
