"""Array formulation of DistributeOctTree (ORBextractor.cc:1006-1287) that the CUDA kernel implements:
nodes live in LIST ORDER (slot == list position), keys never move (they carry a node label), each
'round' divides a set of nodes selected by rank and rebuilds the list as
    reverse(children in creation order) ++ surviving old nodes in old order.
Checked against the C oracle by tests/test_quadtree_proto.py; kept as executable documentation."""
import numpy as np


def order_key(x, y, wcell, hcell):
    """reference raw order: cell-row-major (i then j), then y, then x inside the cell"""
    return (((y // hcell) * 4096 + (x // wcell)) * 4096 + y) * 4096 + x


def distribute(x, y, score, okey, minX, maxX, minY, maxY, N):
    """x, y: int window coords; score: int; okey: unique order key (smaller = earlier in the raw list).
    returns winners (indices into the input) in list order."""
    n = len(x)
    if n == 0:
        return []
    f32 = np.float32
    nIni = int(np.floor(f32(maxX - minX) / f32(maxY - minY) + f32(0.5)))  # C round() for positive values
    if nIni < 1:
        return []
    hX = f32(maxX - minX) / f32(nIni)
    # roots, in list order
    box = [[int(hX * f32(i)), 0, int(hX * f32(i + 1)), maxY - minY] for i in range(nIni)]  # x0,y0,x1,y1
    label = np.array([min(max(int(f32(xx) / hX), 0), nIni - 1) for xx in x], np.int64)
    cnt = np.bincount(label, minlength=nIni)
    # pass 0: drop empty roots
    keep = [i for i in range(nIni) if cnt[i] > 0]
    remap = -np.ones(nIni, np.int64); remap[keep] = np.arange(len(keep))
    label = remap[label]; box = [box[i] for i in keep]; cnt = cnt[keep]
    size = len(box)
    cprev = size

    def round_(ranked):
        """ranked: slots to divide, in processing order. returns (#children created, #children with >1 keys)"""
        nonlocal box, cnt, label, size
        isdiv = np.zeros(size, bool)
        childpos = -np.ones((size, 4), np.int64)
        newbox, newcnt = [], []
        t = 0
        created = []
        for s in ranked:
            x0, y0, x1, y1 = box[s]
            hx = int(np.ceil(f32(x1 - x0) / f32(2))); hy = int(np.ceil(f32(y1 - y0) / f32(2)))
            sel = label == s
            q = (x[sel] >= x0 + hx).astype(np.int64) + 2 * (y[sel] >= y0 + hy).astype(np.int64)  # 0:n1 1:n2 2:n3 3:n4
            cc = np.bincount(q, minlength=4)
            cb = [[x0, y0, x0 + hx, y0 + hy], [x0 + hx, y0, x1, y0 + hy], [x0, y0 + hy, x0 + hx, y1], [x0 + hx, y0 + hy, x1, y1]]
            isdiv[s] = True
            for k in range(4):
                if cc[k] > 0:
                    childpos[s, k] = t; t += 1
                    created.append((cb[k], cc[k]))
        C = t
        surv = [s for s in range(size) if not isdiv[s]]
        newpos = -np.ones(size, np.int64)
        for r, s in enumerate(surv):
            newpos[s] = C + r
        nb = [None] * (C + len(surv)); nc = np.zeros(C + len(surv), np.int64)
        for ti, (b, c) in enumerate(created):
            nb[C - 1 - ti] = b; nc[C - 1 - ti] = c
        for s in surv:
            nb[newpos[s]] = box[s]; nc[newpos[s]] = cnt[s]
        # relabel keys
        newlabel = label.copy()
        for i in range(n):
            s = label[i]
            if isdiv[s]:
                x0, y0, x1, y1 = box[s]
                hx = int(np.ceil(f32(x1 - x0) / f32(2))); hy = int(np.ceil(f32(y1 - y0) / f32(2)))
                q = int(x[i] >= x0 + hx) + 2 * int(y[i] >= y0 + hy)
                newlabel[i] = C - 1 - childpos[s, q]
            else:
                newlabel[i] = newpos[s]
        label = newlabel; box = nb; cnt = nc; size = len(nb)
        return C, int(sum(1 for (_, c) in created if c > 1))

    finish = False
    while not finish:
        prev = size
        ranked = [s for s in range(size) if cnt[s] > 1]
        cprev, nexp = round_(ranked)
        if size >= N or size == prev:
            finish = True
        elif size + 3 * nexp > N:
            while not finish:
                prev = size
                cand = [s for s in range(cprev) if cnt[s] > 1]
                cand.sort(key=lambda s: (-cnt[s], s))      # size desc, later-created (= nearer the front) first
                # speculative cut: first rank where the running size reaches N
                run = size
                cut = len(cand)
                for r, s in enumerate(cand):
                    x0, y0, x1, y1 = box[s]
                    hx = int(np.ceil(f32(x1 - x0) / f32(2))); hy = int(np.ceil(f32(y1 - y0) / f32(2)))
                    sel = label == s
                    q = (x[sel] >= x0 + hx).astype(np.int64) + 2 * (y[sel] >= y0 + hy).astype(np.int64)
                    run += len(np.unique(q)) - 1
                    if run >= N:
                        cut = r + 1
                        break
                cprev, _ = round_(cand[:cut])
                if size >= N or size == prev:
                    finish = True
    # winners: max score, ties -> earliest in raw order
    out = []
    for s in range(size):
        idx = np.nonzero(label == s)[0]
        best = max(idx, key=lambda i: (score[i], -okey[i]))
        out.append(int(best))
    return out
