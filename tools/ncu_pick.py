"""Condense an `ncu --page raw --csv` dump to the columns the roofline discussion needs, one line per launch.
Usage: python tools/ncu_pick.py raw.csv"""
import csv
import sys

COLS = [
    ("gpu__time_duration.sum", "dur"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "hmma%"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__occupancy_limit_registers", "occ_regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
]


def main(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rows = list(csv.reader(lines))
    if len(rows) < 3:
        print("no data in", path)
        return
    header, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(header)}
    name_i = idx.get("Kernel Name")
    print("kernel".ljust(44), " ".join(short.rjust(12) for _, short in COLS))
    for r in rows[2:]:
        out = []
        for full, _ in COLS:
            i = idx.get(full)
            out.append((r[i] + (units[i] if units[i] in ("us", "ms", "ns", "Gbyte", "Mbyte", "Kbyte", "byte") else "")) if i is not None and i < len(r) else "-")
        print(r[name_i][:44].ljust(44), " ".join(o.replace(",", "").rjust(12) for o in out))


if __name__ == "__main__":
    main(sys.argv[1])
