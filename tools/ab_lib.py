"""A/B of library builds on ONE box, independent of the Python binding's symbol list: for every .so given, kline5 f32 on 50 000
device-resident sets, k_quad milliseconds from the library's own CUDA events.  usage: ab_lib.py lib1.so lib2.so ... [--reps 3]"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lineprofiles_b200 import synthetic as syn

libs = [a for a in sys.argv[1:] if a.endswith(".so")]
B = 50000
tab = syn.make_table()
k = {n: np.ascontiguousarray(tab[n], np.float32) for n in ("spins", "mus", "radii", "gmin", "gmax", "upper", "lower")}
fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
E = syn.energy_grid_linear(1000)
p = syn.random_params("kline5", B, seed=1)
dp, dE = torch.from_numpy(p).cuda(), torch.from_numpy(E).cuda()
out = torch.empty((B, 1000), dtype=torch.float64, device="cuda")
res = {}
for rnd in range(3):
    for path in libs:
        L = C.CDLL(os.path.abspath(path))
        h = C.c_void_p()
        L.klp_table_create.argtypes = [C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int] + [C.POINTER(C.c_float)] * 5 + [C.c_int, C.POINTER(C.c_void_p)]
        assert L.klp_table_create(fp(k["spins"]), 30, fp(k["mus"]), 20, 150, 30, fp(k["radii"]), fp(k["gmin"]), fp(k["gmax"]), fp(k["upper"]), fp(k["lower"]), 0, C.byref(h)) == 0
        L.klp_eval_batch_device.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_long, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.klp_set_timing.argtypes = [C.c_void_p, C.c_int]
        L.klp_get_timing.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_long)]
        L.klp_table_destroy.argtypes = [C.c_void_p]
        st = torch.cuda.current_stream().cuda_stream
        best = None
        for r in range(3):
            L.klp_set_timing(h, 1)
            assert L.klp_eval_batch_device(h, 1, 0, B, dp.data_ptr(), dE.data_ptr(), 1000, None, 0, out.data_ptr(), st) == 0
            torch.cuda.synchronize()
            ms = (C.c_double * 4)(); n = (C.c_long * 4)()
            L.klp_get_timing(h, ms, n)
            if r > 0: best = ms[1] if best is None else min(best, ms[1])
        L.klp_table_destroy(h)
        res.setdefault(path, []).append(best)
        print(f"round {rnd} {os.path.basename(path):40s} k_quad {best:8.2f} ms  ({B / best * 1e3:9.0f} spectra/s)  checksum {float(out.sum()):.6f}", flush=True)
for path, v in res.items():
    print(f"{os.path.basename(path):40s} best {min(v):8.2f} ms  median {sorted(v)[len(v)//2]:8.2f} ms")
