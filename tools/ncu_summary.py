"""Summarise an .ncu-rep (raw + source pages) into text: key metrics, instruction mix, hottest SASS block."""
import collections, csv, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__shared_mem_per_block_dynamic',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.max', 'smsp__warps_eligible.avg.per_cycle_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__sass_inst_executed_op_shared_ld.sum', 'smsp__sass_inst_executed_op_global_ld.sum']
print(f"# {rep}")
for i, h in enumerate(hdr):
    if h in want or ('average_warps_issue_stalled' in h and h.endswith('per_issue_active.ratio')):
        try:
            v = float(vals[i])
            if 'issue_stalled' in h and v < 0.05:
                continue
        except ValueError:
            pass
        print(f"{h:88s} {units[i]:16s} {vals[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
h2 = rows[hi]
ci, si = h2.index('Instructions Executed'), h2.index('Source')
byop, tot, counts = collections.Counter(), 0, collections.Counter()
for r in rows[hi + 1:]:
    try:
        n = int(r[ci])
    except Exception:
        continue
    toks = r[si].split()
    op = toks[1] if toks and toks[0].startswith('@') else (toks[0] if toks else '?')
    byop[op.rstrip(';').split('.')[0]] += n
    tot += n
    counts[round(n / 1e6)] += 1
print(f"\ninstruction mix (warp-level, total {tot}):")
for op, n in byop.most_common(18):
    print(f"  {op:10s} {n:14d} {100 * n / tot:5.1f}%")
print("\nSASS instructions grouped by execution count (M executions x #instructions):")
for k in sorted(counts, reverse=True)[:8]:
    print(f"  {k:6d} M x {counts[k]}")
