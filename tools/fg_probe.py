import os, sys
sys.path.insert(0, '/root/repo/tools'); sys.path.insert(0, '/root/repo')
from gpu_perf import run
import lineprofiles_b200 as klp
from lineprofiles_b200 import synthetic as syn
gt = klp.Table.from_arrays(syn.make_table(), 0)
for fg in (0, 1, 0, 1):
    run(gt, "kline5", 50000, "f32", opts={"frame_global": fg})
