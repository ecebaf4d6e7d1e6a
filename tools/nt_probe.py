"""CTA size of k_quad (512 vs 1024 threads) across batch sizes, f32 kline5 and kconv. usage: nt_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gpu_perf import run
import lineprofiles_b200 as klp
from lineprofiles_b200 import synthetic as syn

if __name__ == "__main__":
    gt = klp.Table.from_arrays(syn.make_table(), 0)
    for B in (1, 37, 296, 2000, 8388, 16384):
        for nt in (512, 1024):
            run(gt, "kline5", B, "f32", opts={"quad_threads": nt}, reps=3)
    for nt in (512, 1024):
        run(gt, "kconv", 2000, "f32", n_bins=4096, log=True, opts={"quad_threads": nt})
