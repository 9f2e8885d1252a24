"""Condense `ncu --page raw --csv` output (one row per profiled launch) into the few numbers DESIGN.md / profiles/ quote.

usage: python scripts/ncu_summary.py raw.csv [kernel-substring] > profiles/rNN_<name>.md
"""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 (LTS) throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (elapsed)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (active)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__cluster_dim_x", "cluster x"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
]
STALLS = "smsp__pcsamp_warps_issue_stalled_"


def main():
    path = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    rows = list(csv.reader(open(path)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    print(f"# ncu summary of `{path}`\n")
    for r in body:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        name = d.get("Kernel Name", "?")
        if want and want not in name:
            continue
        print(f"## {name}  (launch id {d.get('ID', '?')})\n")
        print("| metric | value | unit |\n|---|---|---|")
        for k, label in KEYS:
            if k in d and d[k] != "":
                print(f"| {label} (`{k}`) | {d[k]} | {u.get(k, '')} |")
        try:
            rd, wr = float(d["dram__bytes_read.sum"]), float(d["dram__bytes_write.sum"])
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            tot = rd * scale[u["dram__bytes_read.sum"]] + wr * scale[u["dram__bytes_write.sum"]]
            print(f"| **DRAM traffic (read + write)** | {tot / 1e6:.1f} | MB |")
        except (KeyError, ValueError):
            pass
        st = [(k[len(STALLS):], float(d[k])) for k in hdr if k.startswith(STALLS) and not k.endswith("_not_issued") and d[k] not in ("", "0")]
        tot = sum(v for _, v in st)
        if tot > 0:
            st.sort(key=lambda kv: -kv[1])
            print("\nwarp-state samples: " + ", ".join(f"{k} {100 * v / tot:.0f}%" for k, v in st[:7]))
        print()


if __name__ == "__main__":
    main()
