"""Spatial (Morton) processing order for the transposed conv backward: L1 hit rates of the `bwd` schedule of l1sim.py with the
points of a CTA round taken in Morton order of their coordinates (CPU, numpy; results quoted in DESIGN.md section 9).
    python profiles/l1sim_spatial.py        (~4 minutes)"""
import sys, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import l1sim as L
N, K = L.N, L.K
# rebuild the graph, keeping xyz (same seed as l1sim.build_graph)
rng = np.random.default_rng(1234)
xyz = rng.random((N, 3), dtype=np.float32)
idx, cnt, filt = L.build_graph()
mask = np.arange(K)[None, :] < cnt[:, None]
m_of = np.broadcast_to(np.arange(N)[:, None], (N, K))[mask]
n_of, f_of = idx[mask], filt[mask]
indeg = np.bincount(n_of, minlength=N)
srt = np.argsort(n_of, kind="stable")
ms, fs = m_of[srt], f_of[srt]
u, first, counts = np.unique(n_of[srt], return_index=True, return_counts=True)
bypt = {int(a): (ms[b:b + c], fs[b:b + c]) for a, b, c in zip(u, first, counts)}

def morton(p, bits=5):
    q = np.minimum((p * (1 << bits)).astype(np.int64), (1 << bits) - 1)
    code = np.zeros(len(p), np.int64)
    for b in range(bits):
        for d in range(3):
            code |= ((q[:, d] >> b) & 1) << (3 * b + d)
    return code

def bwd(order, per_round, cap_lines, stride=1, G=4, entry_sort="bin_row"):
    lru = L.LRU(cap_lines)
    rounds = list(range(0, len(order), per_round))[::stride]
    for r0 in rounds:
        streams = []
        for h in order[r0:r0 + per_round]:
            mm, ff = bypt.get(int(h), (np.array([], np.int64),) * 2)
            for c in range(G):
                s = (ff % G) == c
                if entry_sort == "bin_row":
                    o = np.lexsort((mm[s], ff[s]))
                else:
                    o = np.argsort(mm[s], kind="stable")
                streams.append(mm[s][o].tolist())
        for m in L.interleave(streams):
            for l in range(4):
                lru.acc(m * 4 + l)
    return lru.rate

active = np.nonzero(indeg > 0)[0]
mo = active[np.argsort(morton(xyz[active]), kind="stable")]
print("as built, index order, every 37th round:           %.3f" % bwd(np.arange(N), 6, 800, stride=37))
print("morton order, consecutive points per CTA (6):      %.3f" % bwd(mo, 6, 800, stride=7))
# hubs only (the 256 lowest indices carry half the edges): morton among hubs
hubs = np.arange(256)
hm = hubs[np.argsort(morton(xyz[hubs]), kind="stable")]
print("hubs (256 lowest) in index order, 6 per CTA:       %.3f" % bwd(hubs, 6, 800))
print("hubs in morton order, 6 per CTA:                   %.3f" % bwd(hm, 6, 800))
print("hubs in morton order, 6 per CTA, 200 KB L1:        %.3f" % bwd(hm, 6, 1600))
print("hubs in morton order, 6 per CTA, entries by row:   %.3f" % bwd(hm, 6, 800, entry_sort="row"))
