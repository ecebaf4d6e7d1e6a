"""One-correction against two-correction division (libraries built with / without -DKLP_DIV_TWO_CORRECTIONS): every model in both
precisions on N random sets - the two builds must agree BIT FOR BIT (both divisions are correctly rounded), and k_quad times.
usage: div_ab.py LIB_ONE LIB_TWO [N]"""
import ctypes as C, os, subprocess, sys, json
import numpy as np

if len(sys.argv) > 3 and sys.argv[1] == "--child":
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import lineprofiles_b200 as klp
    from lineprofiles_b200 import synthetic as syn
    n = int(sys.argv[2])
    gt = klp.Table.from_arrays(syn.make_table(), 0)
    E, El = syn.energy_grid_linear(1000), syn.energy_grid_log(512)
    spec = syn.powerlaw_continuum(El)
    out = {}
    for prec in ("f32", "f64"):
        for model in ("kline", "kline5", "kconv", "kconv5", "kconv10"):
            conv = model.startswith("kconv")
            p = syn.random_params(model, n, seed=300)
            r = gt.eval_batch(model, p, El if conv else E, spectrum=spec if conv else None, precision=prec)
            out[f"{prec} {model}"] = r
    np.savez(sys.argv[3], **out)
    sys.exit(0)

one, two = sys.argv[1], sys.argv[2]
n = sys.argv[3] if len(sys.argv) > 3 else "300"
res = []
for k, lib in enumerate((one, two)):
    f = f"/tmp/div_ab_{k}.npz"
    subprocess.check_call([sys.executable, __file__, "--child", n, f], env=dict(os.environ, KLP_LIBRARY=os.path.abspath(lib)))
    res.append(np.load(f))
ok = True
for key in res[0].files:
    same = np.array_equal(res[0][key], res[1][key], equal_nan=True)
    ok &= same
    print(f"{key:14s} {n} sets: one correction == two corrections bit for bit: {same}")
print("ALL IDENTICAL" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
