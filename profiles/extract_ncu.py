"""Condense an .ncu-rep (ncu --set full) into the per-launch numbers DESIGN.md / bench.py quote.
usage: python profiles/extract_ncu.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import csv
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "time"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_%act"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_pipe_%elapsed"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
        ("lts__t_bytes.sum", "l2_bytes"), ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("launch__block_size", "block"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_%"),
        ("smsp__inst_executed.sum", "warp_insts"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%")]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("# %s" % path)
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].replace("void ", "").replace("<unnamed>::", "").split("(")[0]
        parts = []
        for k, short in KEYS:
            if k in idx:
                parts.append("%s=%s%s" % (short, r[idx[k]], units[idx[k]] if units[idx[k]] not in ("", "%") else ""))
        print("%-36s %s" % (name[:36], "  ".join(parts)))


if __name__ == "__main__":
    main(sys.argv[1])
