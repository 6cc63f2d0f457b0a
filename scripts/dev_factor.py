"""dev: factorisation timings (symbolic / numeric / pack, CUDA events) under different options.
usage: dev_factor.py <problem> <kind> [key=value ...]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ilupp_b200 as ilupp
from ilupp_b200 import problems as P

name, kind = sys.argv[1], sys.argv[2]
for kv in sys.argv[3:]:
    k, v = kv.split("="); ilupp.set_option(k, int(v))
A = {"p2d256": lambda: P.poisson2d(256), "p3d128": lambda: P.poisson3d(128), "p3d64": lambda: P.poisson3d(64),
     "cd3d160": lambda: P.convdiff3d(160), "irr1m": lambda: P.irregular_powerlaw()}[name]()
cls = ilupp.IChol0Preconditioner if kind == "ichol0" else ilupp.ILU0Preconditioner
for rep in range(3):
    t = time.perf_counter()
    B = cls(A)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t
    print("%s %s %s: create wall %.1f ms; symbolic %.2f numeric %.3f pack %.3f total %.2f ms" % ((name, kind, sys.argv[3:], wall * 1e3) + tuple(B.timings_ms)), flush=True)
a = torch.from_numpy(A.data).cuda()
for rep in range(3):
    B.refactor(a)
    print("  refactor: numeric %.3f pack %.3f ms" % (B.timings_ms[1], B.timings_ms[2]))
