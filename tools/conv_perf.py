"""kconv throughput, reference-order vs cumulative-profile convolution (option conv_mode)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gpu_perf import run
import lineprofiles_b200 as klp
from lineprofiles_b200 import synthetic as syn

gt = klp.Table.from_arrays(syn.make_table(), 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
for mode in (0, 1):
    run(gt, "kconv", B, "f32", n_bins=4096, log=True, opts={"conv_mode": mode})
    run(gt, "kconv10", B, "f32", n_bins=4096, log=True, opts={"conv_mode": mode})
    run(gt, "kconv", B, "f32", n_bins=1000, log=False, opts={"conv_mode": mode})
