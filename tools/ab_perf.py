"""A/B probe for one library build (KLP_LIBRARY): kline5 f32 throughput over CTA sizes. usage: ab_perf.py B nt,nt,..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gpu_perf import run
import lineprofiles_b200 as klp
from lineprofiles_b200 import synthetic as syn

if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    nts = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [512]
    gt = klp.Table.from_arrays(syn.make_table(), 0)
    print("library", klp.api.library_path())
    for nt in nts:
        run(gt, "kline5", B, "f32", opts={"quad_threads": nt})
