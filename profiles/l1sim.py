"""L1 hit-rate simulator for the gather kernels (CPU, numpy; no GPU, no oracle).

Builds ONE cloud of the Cfg-T graph (N = 10 000 uniform points, K = 64, saturating radius, the reference's growing-radius
chain and first-K-by-index rule restated in numpy -- exactness is irrelevant for a cache study) and replays the address
stream of a CTA through an LRU cache of 128-byte lines, for the schedules that were considered in round 2:

  fwd      conv forward: 128-row chunks, 32 warps, rows of 512 B (4 lines) vs 64- / 32-channel columns (2 / 1 lines)
  bwd      transposed backward as built: 6 points x 4 bin classes per CTA, points dealt round-robin, 100 KB of L1
  bwd-deg  the same with the points in degree order, consecutive points per CTA
  quad     streaming form: 4 lists per warp, 32-channel column, lists in degree order, `per_round` lists per CTA round,
           entries sorted by (row block, bin, row) / (bin, row) / row

    python profiles/l1sim.py            (~2 minutes; results of the round are in profiles/r2_l1sim.txt)
"""
from collections import OrderedDict

import numpy as np

N, K = 10000, 64


def build_graph(seed=1234):
    rng = np.random.default_rng(seed)
    xyz = rng.random((N, 3), dtype=np.float32)
    radius = np.float32((3.0 * 2 * K / (4.0 * np.pi * N)) ** (1.0 / 3.0))
    idx = np.zeros((N, K), np.int64); cnt = np.zeros(N, np.int64); filt = np.zeros((N, K), np.int64)
    for j0 in range(0, N, 1000):
        q = xyz[j0:j0 + 1000]
        d = np.sqrt(((xyz[None, :, :] - q[:, None, :]) ** 2).sum(-1))
        r = radius + np.float32(0.05) * ((np.arange(j0, j0 + len(q)) // 1024).astype(np.float32))     # chain step = j // 1024 (B <= 32)
        inr = d < r[:, None]
        for i in range(len(q)):
            nb = np.nonzero(inr[i])[0][:K]
            m = j0 + i
            cnt[m] = len(nb); idx[m, :len(nb)] = nb
            v = xyz[nb] - xyz[m]
            az = ((np.arctan2(v[:, 1], v[:, 0]) + np.pi) * 8 / (2 * np.pi)).astype(int).clip(0, 7)
            el = ((np.arctan2(v[:, 2], np.hypot(v[:, 0], v[:, 1])) + np.pi / 2) * 2 / np.pi).astype(int).clip(0, 1)
            rad = np.minimum(1, (np.sqrt(d[i, nb]) * 2 / (radius + 1e-6)).astype(int))
            filt[m, :len(nb)] = np.where(d[i, nb] < 1e-6, 0, rad * 16 + el * 8 + az + 1)
    return idx, cnt, filt


class LRU:
    def __init__(self, cap):
        self.cap, self.d, self.hit, self.miss = cap, OrderedDict(), 0, 0

    def acc(self, line):
        if line in self.d:
            self.d.move_to_end(line); self.hit += 1
        else:
            self.miss += 1; self.d[line] = 1
            if len(self.d) > self.cap:
                self.d.popitem(last=False)

    @property
    def rate(self):
        return self.hit / max(1, self.hit + self.miss)


def interleave(streams):
    for i in range(max((len(x) for x in streams), default=0)):
        for x in streams:
            if i < len(x):
                yield x[i]


def main():
    idx, cnt, filt = build_graph()
    mask = np.arange(K)[None, :] < cnt[:, None]
    m_of = np.broadcast_to(np.arange(N)[:, None], (N, K))[mask]
    n_of, f_of = idx[mask], filt[mask]
    E = len(n_of)
    indeg = np.bincount(n_of, minlength=N)
    print("E %d, edges into the 64 / 256 / 1000 lowest-index points: %.2f / %.2f / %.2f" % (
        E, (n_of < 64).mean(), (n_of < 256).mean(), (n_of < 1000).mean()))
    print("(point, bin) segments: %d (%.2f edges each)" % (len(np.unique(n_of * 64 + f_of)), E / len(np.unique(n_of * 64 + f_of))))

    # ---- forward: capacity in rows
    for lines_per_row, cap_rows in ((4, 390), (2, 780), (1, 1560)):
        hit = tot = 0
        for start in range(5):
            lru = LRU(cap_rows)
            for ch in range(start, (N + 127) // 128, 5):
                streams = [[] for _ in range(32)]
                for i, r in enumerate(range(ch * 128, min(N, ch * 128 + 128))):
                    o = np.argsort(filt[r, :cnt[r]], kind="stable")
                    streams[i % 32].extend(idx[r, :cnt[r]][o].tolist())
                for n in interleave(streams):
                    lru.acc(n)
            hit += lru.hit; tot += lru.hit + lru.miss
        print("fwd  %d-byte row pieces, L1 holds %4d of them: hit %.3f" % (128 * lines_per_row, cap_rows, hit / tot))

    srt = np.argsort(n_of, kind="stable")
    ms, fs = m_of[srt], f_of[srt]
    u, first, counts = np.unique(n_of[srt], return_index=True, return_counts=True)
    bypt = {int(a): (ms[b:b + c], fs[b:b + c]) for a, b, c in zip(u, first, counts)}
    by_degree = np.argsort(-indeg, kind="stable")[:int((indeg > 0).sum())]

    def bwd(order, per_round, cap_lines, stride=1, G=4):
        lru = LRU(cap_lines)
        rounds = list(range(0, len(order), per_round))[::stride]
        for r0 in rounds:
            streams = []
            for h in order[r0:r0 + per_round]:
                mm, ff = bypt.get(int(h), (np.array([], np.int64),) * 2)
                for c in range(G):
                    s = (ff % G) == c
                    streams.append(mm[s][np.lexsort((mm[s], ff[s]))].tolist())
            for m in interleave(streams):
                for l in range(4):
                    lru.acc(m * 4 + l)
        return lru.rate

    print("bwd  as built (6 points x 4 classes per CTA, every 37th round, 100 KB L1): hit %.3f" % bwd(np.arange(N), 6, 800, stride=37))
    print("bwd  degree order, consecutive points per CTA, 100 KB L1:                  hit %.3f" % bwd(by_degree, 6, 800))
    print("bwd  degree order, 200 KB L1:                                              hit %.3f" % bwd(by_degree, 6, 1600))

    def quad(points, per_round, cap_lines, order, MB=None):
        lru = LRU(cap_lines); g = padded = 0
        for r0 in range(0, len(points), per_round):
            streams = []
            rp = points[r0:r0 + per_round]
            for q in range(0, len(rp), 4):
                lists = []
                for h in rp[q:q + 4]:
                    mm, ff = bypt.get(int(h), (np.array([], np.int64),) * 2)
                    o = {"bin_row": np.lexsort((mm, ff)), "row": np.argsort(mm, kind="stable"),
                         "block_bin_row": np.lexsort((mm, ff, mm // (MB or N)))}[order]
                    lists.append(mm[o])
                L = max(len(x) for x in lists); padded += 4 * L; g += sum(len(x) for x in lists)
                streams.append([int(x[i]) for i in range(L) for x in lists if i < len(x)])
            for m in interleave(streams):
                lru.acc(m)
        return lru.rate, padded / g, lru.miss / g

    for per_round, order, MB in ((128, "bin_row", None), (128, "row", None), (128, "block_bin_row", 1024), (64, "block_bin_row", 1024)):
        print("quad %3d lists per round, entries by %-14s %s: hit %.3f, lock-step padding %.2f, L2 lines per edge %.3f"
              % ((per_round, order, MB or "") + quad(by_degree, per_round, 1600, order, MB)))


if __name__ == "__main__":
    main()
