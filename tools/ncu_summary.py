"""Summarise `ncu -i X.ncu-rep --page raw --csv` output: one block of the metrics DESIGN.md / profiles/ quote per kernel.
usage: ncu -i rep.ncu-rep --page raw --csv | python tools/ncu_summary.py [cells]"""
import csv
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX (shared-memory wavefront) pipe % busy"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % busy"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % active"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots % busy"),
    ("sm__inst_executed.avg.per_cycle_elapsed", "IPC (per SM)"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__cycles_elapsed.max", "SM cycles"),
]


def main():
    rows = list(csv.reader(l for l in sys.stdin if not l.startswith("==")))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cells = float(sys.argv[1]) if len(sys.argv) > 1 else None
    tot_r = tot_w = 0.0
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("unnamed>::", "")
        print(f"== {name}")
        for key, label in WANT:
            if key in idx and r[idx[key]] != "":
                print(f"   {label:48s} {r[idx[key]]:>18s} {units[idx[key]]}")
        def gb(key):
            v, u = float(r[idx[key]]), units[idx[key]]
            return v * {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9}[u]
        rd, wr = gb("dram__bytes_read.sum"), gb("dram__bytes_write.sum")
        tot_r += rd
        tot_w += wr
        t = float(r[idx["gpu__time_duration.sum"]]) * {"ms": 1e-3, "us": 1e-6, "s": 1.0, "ns": 1e-9}[units[idx["gpu__time_duration.sum"]]]
        print(f"   {'DRAM read + write / duration':48s} {(rd + wr) / t:18.1f} GB/s")
    print(f"== sum over the listed launches: DRAM read {tot_r:.3f} GB + write {tot_w:.3f} GB = {tot_r + tot_w:.3f} GB", end="")
    if cells:
        print(f" = {(tot_r + tot_w) * 1e9 / cells / 1e3:.1f} KB per cell ({int(cells)} cells)")
    else:
        print()


if __name__ == "__main__":
    main()
