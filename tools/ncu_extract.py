"""Selected raw-page metrics of an ncu report -> CSV lines (profiles/*_ncu_raw_extract.csv).  usage: python tools/ncu_extract.py rep.ncu-rep [out.csv]"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__cluster", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum", "sm__pipe_tensor_cycles_active", "sm__pipe_fp64_cycles_active",
        "sm__ops_path_tensor_op_utcimma", "sm__inst_executed_pipe_tensor", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_active.avg", "sm__cycles_elapsed.avg.per_second", "lts__cycles_elapsed.avg.per_second", "dram__cycles_elapsed.avg.per_second",
        "smsp__warp_issue_stalled", "smsp__average_warp", "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank", "smsp__inst_executed.sum",
        "sm__sass_inst_executed_op_shared", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "lts__d_sectors_fill", "lts__t_sectors_srcunit_ltcfabric",
        "sm__inst_executed_pipe_fp64", "smsp__inst_executed_pipe_fp64", "sm__pipe_fma_cycles_active", "sm__pipe_alu_cycles_active", "sm__inst_executed_pipe_uniform",
        "lts__t_sectors_srcnode_gpc", "lts__average_t_sector", "l1tex__m_xbar2l1tex_read_sectors"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, universal_newlines=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = []
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""
        for h, u, v in zip(hdr, units, r):
            if any(w in h for w in WANT) and v not in ("", "0"):
                lines.append("%s,%s,%s,%s" % (name.split("(")[0][-60:], h, u, v))
    text = "\n".join(lines) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    else:
        sys.stdout.write(text)


if __name__ == "__main__":
    main()
