"""Decode the reference's bundled datasets (R/data.R:1-65; data/*.rda) into the package data files
microclimc_b200/data/{weather.npz,dailyprecip.npy,soilparams.json,climvars.json}.
Run once in the build container (needs /root/reference); outputs are committed."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(__file__))
from rda_reader import read_rda

REF = "/root/reference/data"
OUT = os.path.join(os.path.dirname(__file__), "..", "microclimc_b200", "data")

w = read_rda(f"{REF}/weather.rda")["weather"]
cols = ["temp", "relhum", "pres", "swrad", "difrad", "skyem", "windspeed", "winddir"]
np.savez_compressed(f"{OUT}/weather.npz", obs_time=np.array(w["obs_time"]),
                    **{c: np.asarray(w[c], dtype=np.float64) for c in cols})
np.save(f"{OUT}/dailyprecip.npy", np.asarray(read_rda(f"{REF}/dailyprecip.rda")["dailyprecip"], dtype=np.float64))
sp = read_rda(f"{REF}/soilparams.rda")["soilparams"]
json.dump({k: v for k, v in sp.items() if k != "__attrs__"}, open(f"{OUT}/soilparams.json", "w"), indent=1)
cv = read_rda(f"{REF}/climvars.rda")["climvars"]
json.dump({k: (None if v[0] is None else v[0]) for k, v in cv.items() if k != "__attrs__"},
          open(f"{OUT}/climvars.json", "w"), indent=1)
print("ok")
