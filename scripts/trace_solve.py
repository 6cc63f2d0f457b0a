"""dev tool: per-chunk timeline of the sync-free solve (option trace=1)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ilupp_b200 as ilupp
from ilupp_b200 import problems as P
from scripts.bench_apply import PROBLEMS

prob, kind = sys.argv[1], sys.argv[2]
for kv in sys.argv[3:]:
    k, v = kv.split("=")
    ilupp.set_option(k, int(v))
ilupp.set_option("trace", 1)
A = PROBLEMS[prob]()
n = A.shape[0]
B = (ilupp.IChol0Preconditioner if kind == "ichol0" else ilupp.ILU0Preconditioner)(A)
x = torch.from_numpy(P.rhs(n, 0)).cuda()
for _ in range(3):
    B.apply(x)
torch.cuda.synchronize()
lev, nlev = B.levels(0)
rows, tr = B.pr.debug_chunks(0)
rows = np.asarray(rows).reshape(-1, 32)
tr = np.asarray(tr).reshape(-1, 4).astype(np.int64)
st = B.stats
print("nchunks", rows.shape[0], "parts", st["nparts"], "warps", st["warps"], "nlev", nlev)
# which CTA: find chunk boundaries by level resets: use rows->level
valid = rows >= 0
lv = np.where(valid, lev[np.clip(rows, 0, n - 1)], -1)
lvmin = np.where(valid, lv, 1 << 30).min(axis=1)
lvmax = lv.max(axis=1)
# split into CTAs where lvmin decreases
starts = [0] + [i for i in range(1, len(lvmin)) if lvmin[i] < lvmin[i - 1]] + [len(lvmin)]
print("CTAs found", len(starts) - 1)
for ci in (0, (len(starts) - 1) // 2):
    a, b = starts[ci], starts[ci + 1]
    t = tr[a:b]
    t0 = t[:, 0].min()
    print("CTA", ci, "chunks", b - a, "duration cycles", t[:, 3].max() - t0)
    d_done = np.diff(np.sort(t[:, 3]))
    print("  completion spacing: median %.0f mean %.0f p90 %.0f max %.0f cycles" % (np.median(d_done), d_done.mean(), np.percentile(d_done, 90), d_done.max()))
    print("  per chunk: wait-for-gate median %.0f | stage wait median %.0f | resolve median %.0f cycles" % (
        np.median(t[:, 1] - t[:, 0]), np.median(t[:, 2] - t[:, 1]), np.median(t[:, 3] - t[:, 2])))
    # level completion times: time when last chunk containing level L done
    nl = lvmax[a:b].max() + 1
    tl = np.zeros(nl, dtype=np.int64)
    for k in range(b - a):
        for L in range(lvmin[a + k], lvmax[a + k] + 1):
            tl[L] = max(tl[L], t[k, 3])
    tl = tl[tl > 0]
    dl = np.diff(tl)
    print("  level-to-level: median %.0f mean %.0f p90 %.0f cycles over %d levels" % (np.median(dl), dl.mean(), np.percentile(dl, 90), len(tl)))
    k0 = (b - a) // 2
    print("  sample chunks (rel. cycles: start gate staged done | lvmin lvmax nrows):")
    for k in range(k0, min(k0 + 24, b - a)):
        print("   %5d: %8d %8d %8d %8d | %4d %4d %3d" % (k, *(t[k] - t0), lvmin[a + k], lvmax[a + k], valid[a + k].sum()))
