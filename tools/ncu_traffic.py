"""Per-spectrum figures of one k_quad launch from an .ncu-rep (ncu --set full): DRAM bytes, warp instructions, FP32-datapath
cycles -> the JSON bench.py reads (profiles/r01_quad_traffic.json).  usage: ncu_traffic.py REP SETS_PER_LAUNCH SOURCE_NOTE > out.json"""
import csv, json, subprocess, sys

import os
rep, sets, note = sys.argv[1], int(sys.argv[2]), sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]


def get(name):
    i = hdr.index(name)
    v = float(vals[i].replace(",", ""))
    u = units[i]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
    return v * scale


inst = get("smsp__inst_executed.sum")
cyc = get("smsp__cycles_active.avg")
fma = get("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active")
n_smsp = 148 * 4
# the hot-loop fingerprint of the build the capture was taken from (profiles/hot_loop_fingerprint.json is re-recorded together
# with every measured build): bench.py quotes the per-instruction figures only for a build with this very loop
fp_rec = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "hot_loop_fingerprint.json")
out = {
    "kernel": vals[hdr.index("Kernel Name")],
    "hot_loop_fingerprint": json.load(open(fp_rec))["fingerprint"] if os.path.exists(fp_rec) else None,
    "source": note,
    "sets_per_launch": sets,
    "dram_bytes_read": get("dram__bytes_read.sum"),
    "dram_bytes_write": get("dram__bytes_write.sum"),
    "warp_instructions_per_set": inst / sets,
    "warp_instructions_source": "smsp__inst_executed.sum of the same capture",
    "issue_active_pct": get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "fma_pipe_cycles_active_pct": fma,
    "alu_pipe_cycles_active_pct": get("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
    "fma_pipe_cycles_per_set": fma / 100.0 * cyc * n_smsp / sets,
    "fma_pipe_cycles_source": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active x smsp__cycles_active.avg x 592 sub-partitions / sets "
                              "(FFMA2/FMUL2/FADD2 occupy the FP32 datapath for two cycles)",
    "gpu_time_ms_under_ncu": get("gpu__time_duration.sum") / (1e6 if units[hdr.index("gpu__time_duration.sum")] in ("ns", "nsecond") else 1.0),
    "note": "DRAM traffic is the deposit scratch (evaluate-anywhere / add-in-order): 4 B written and read back per candidate (radius, bin); "
            "the table itself stays in L2. Option deposit_ring keeps them in an L2-resident ring with four consumer warps instead (no DRAM reads, a sixth of the traffic, ~5 % slower: DESIGN.md section 4).",
}
print(json.dumps(out, indent=1))
