"""Summarises an .ncu-rep (read here, without a GPU) into the few numbers DESIGN.md and
profiles/ quote.  Usage: python tools/ncu_summary.py report.ncu-rep [--source]"""
import collections
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__issue_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'sm__cycles_elapsed.avg',
        'smsp__cycles_active.avg', 'sm__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio']


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def main():
    rep = sys.argv[1]
    rows = page(rep, "raw")
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        print("kernel:", d.get("Kernel Name"))
        for k in KEYS:
            if k in d:
                print("  %-82s %14s %s" % (k, d[k], units[hdr.index(k)]))
    if "--source" in sys.argv:
        rows = page(rep, "source")
        h = rows[1]
        ix = {n: i for i, n in enumerate(h)}
        data = rows[2:]

        def f(r, k):
            try:
                return float(r[ix[k]])
            except Exception:
                return 0.0
        tot = sum(f(r, '# Samples') for r in data)
        stalls = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
        agg = {s: sum(f(r, s) for r in data) for s in stalls}
        print("warp-state samples: %d; share by state:" % tot)
        for k, v in sorted(agg.items(), key=lambda x: -x[1])[:9]:
            print("  %-24s %.3f" % (k, v / tot))
        c, n = collections.Counter(), collections.Counter()
        for r in data:
            src = [o for o in r[ix['Source']].split() if not o.startswith('@')]
            if not src:
                continue
            op = src[0].split('.')[0]
            c[op] += f(r, '# Samples')
            n[op] += f(r, 'Instructions Executed')
        print("samples / executed warp-instructions by opcode:")
        for k, v in c.most_common(14):
            print("  %-8s samples %5.1f%%   executed %5.1f%%" % (k, 100 * v / tot, 100 * n[k] / sum(n.values())))


if __name__ == "__main__":
    main()
