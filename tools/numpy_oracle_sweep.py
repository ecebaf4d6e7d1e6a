"""Sweep: random parameter sets of all five models through the numpy restatement (oracle/numpy_check.py) and the C++ oracle,
f32, compared bit for bit.  CPU only; ~4 min.  Last run: 196 sets, 0 mismatches."""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np
from oracle import numpy_check as nc, oracle_py
from lineprofiles_b200 import synthetic as syn
import make_numpy_golden as mg
warnings.filterwarnings("ignore")
tab=syn.make_table()
ot=oracle_py.OracleTable(tab)
E=syn.energy_grid_linear(1000); El=syn.energy_grid_log(128); spec=syn.powerlaw_continuum(El)
bad=0; n=0
for model,B,seed in (("kline5",120,1001),("kline",40,1002),("kconv",12,1003),("kconv5",12,1004),("kconv10",12,1005)):
    P=syn.random_params(model,B,seed=seed)
    conv=model.startswith("kconv")
    ref=ot.eval_batch(model,P,El if conv else E,spectrum=spec if conv else None,precision="f32")["out"]
    for s,p in enumerate(P):
        out=mg.spectrum(tab,model,p,El if conv else E,spec)
        ok=bool(np.all((out==ref[s])|(np.isnan(out)&np.isnan(ref[s]))))
        n+=1
        if not ok:
            bad+=1; print("MISMATCH",model,s,list(p),flush=True)
    print(model,"done, mismatches so far",bad,"of",n,flush=True)
print("total",n,"mismatches",bad)
