"""Per-instruction listing of the quadrature's hot loop from an ncu --set full report (source page): offset, executions
relative to the loop header, share of the kernel's warp-state samples, top stall reason, SASS.
usage: ncu_hot_loop.py REP "title line" > profiles/..._hot_loop_sass.txt"""
import csv, subprocess, sys

rep, title = sys.argv[1], sys.argv[2]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]
ci, si, ai = h.index("Instructions Executed"), h.index("Source"), h.index("Address")
smp = h.index("# Samples") if "# Samples" in h else h.index("Warp Stall Sampling (All Samples)")
stall_cols = [(i, c[len("stall_"):]) for i, c in enumerate(h) if c.startswith("stall_")]
ins = []
for r in rows[hi + 1:]:
    try:
        n = int(r[ci])
    except Exception:
        continue
    s = int(r[smp] or 0)
    top = max(stall_cols, key=lambda ic: int(r[ic[0]] or 0))[1] if s else "-"
    ins.append((r[ai], n, s, top, r[si]))
tot_n, tot_s = sum(x[1] for x in ins), sum(x[2] for x in ins)
mx = max(x[1] for x in ins)
# the hot loop: the longest run of instructions executed at least 70 % as often as the most executed one
best, cur = (0, 0), None
for i, x in enumerate(ins):
    if x[1] >= 0.7 * mx:
        cur = (cur[0], i + 1) if cur else (i, i + 1)
        if cur[1] - cur[0] > best[1] - best[0]:
            best = cur
    elif x[1] < 0.05 * mx:      # edge-term blocks inside the loop run rarely: skip them without closing the run
        continue
    else:
        cur = None
lo, hi2 = best
body = [x for x in ins[lo:hi2] if x[1] >= 0.05 * mx]
n_b, s_b = sum(x[1] for x in body), sum(x[2] for x in body)
print(f"# {title}")
print("# the hot loop of stage A: one iteration = 32 fine bins x 7 Gauss nodes of one radius.  Columns: offset, executions relative to the")
print("# loop header, share of all warp-state samples of the kernel (%), top stall reason of the instruction, SASS.  Blocks executed by < 5 % of")
print("# the iterations (edge terms) are left out.")
print(f"# {len(body)} instructions, {100.0 * n_b / tot_n:.1f} % of all executed warp instructions, {100.0 * s_b / max(tot_s, 1):.1f} % of all samples")
ref = body[0][1]
for a, n, s, top, t in body:
    print(f"{a[-6:]}  {n / ref:5.3f}  {100.0 * s / max(tot_s, 1):5.2f}  {top:18s} {t}")
