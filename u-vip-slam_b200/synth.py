"""Deterministic synthetic inputs for the ORB front-end.

Integer-only, counter-based SplitMix64 so every language produces the same bytes.  Used by
tests/ and bench.py; nothing here touches the CPU oracle."""
import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
_GOLD = np.uint64(0x9E3779B97F4A7C15)
_C1 = np.uint64(0xBF58476D1CE4E5B9)
_C2 = np.uint64(0x94D049BB133111EB)


def draw(seed, k):
    """SplitMix64 draw number k (k = 1, 2, ...; scalar or array) of stream `seed` -> uint64."""
    with np.errstate(over='ignore'):
        z = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + np.asarray(k, np.uint64) * _GOLD
        z = (z ^ (z >> np.uint64(30))) * _C1
        z = (z ^ (z >> np.uint64(27))) * _C2
        return z ^ (z >> np.uint64(31))


def synth_frame(seed, W, H, dx=0, dy=0, noise_seed=None):
    """Smooth random background + random gray rectangles (corners) + +-3 noise.  uint8 (H, W)."""
    if noise_seed is None:
        noise_seed = seed
    Wc, Hc = W + 64, H + 64
    gh, gw = Hc // 16 + 2, Wc // 16 + 2
    g = (draw(seed, np.arange(1, gh * gw + 1, dtype=np.uint64)) % np.uint64(256)).astype(np.int64).reshape(gh, gw)
    ys = np.arange(Hc); xs = np.arange(Wc)
    cy, wy = ys // 16, (ys % 16)[:, None]
    cx, wx = xs // 16, (xs % 16)[None, :]
    g00 = g[cy][:, cx]; g01 = g[cy][:, cx + 1]; g10 = g[cy + 1][:, cx]; g11 = g[cy + 1][:, cx + 1]
    canvas = (g00 * (16 - wx) * (16 - wy) + g01 * wx * (16 - wy) + g10 * (16 - wx) * wy + g11 * wx * wy + 128) >> 8
    canvas = canvas.astype(np.uint8)
    R = (Wc * Hc) // 1500
    base = gh * gw + 1
    d = draw(seed, np.arange(base, base + 5 * R, dtype=np.uint64)).reshape(R, 5)
    rx = (d[:, 0] % np.uint64(Wc)).astype(np.int64); ry = (d[:, 1] % np.uint64(Hc)).astype(np.int64)
    rw = 6 + (d[:, 2] % np.uint64(43)).astype(np.int64); rh = 6 + (d[:, 3] % np.uint64(43)).astype(np.int64)
    gray = (d[:, 4] % np.uint64(256)).astype(np.uint8)
    for i in range(R):
        canvas[ry[i]:min(ry[i] + rh[i], Hc), rx[i]:min(rx[i] + rw[i], Wc)] = gray[i]
    crop = canvas[32 + dy:32 + dy + H, 32 + dx:32 + dx + W].astype(np.int16)
    noise = (draw(noise_seed ^ 0xA5A5A5A5, np.arange(1, W * H + 1, dtype=np.uint64)) % np.uint64(7)).astype(np.int16) - 3
    return np.clip(crop + noise.reshape(H, W), 0, 255).astype(np.uint8)


def synth_batch(seed0, n, W, H):
    return np.stack([synth_frame(seed0 + i, W, H) for i in range(n)])


def random_descriptors(seed, n):
    """n x 32 uniform random bytes: byte j of row i = draw(seed, 32 i + j + 1) mod 256."""
    k = np.arange(1, 32 * n + 1, dtype=np.uint64)
    return (draw(seed, k) % np.uint64(256)).astype(np.uint8).reshape(n, 32)


def flip_bits(desc, seed, counts):
    """Flip counts[i] distinct bits of row i (positions from stream `seed`)."""
    out = desc.copy()
    for i, c in enumerate(counts):
        if c == 0:
            continue
        pos = []
        k = 1
        while len(pos) < c:
            p = int(draw(seed + i, k)) % 256
            k += 1
            if p not in pos:
                pos.append(p)
        for p in pos:
            out[i, p >> 3] ^= np.uint8(1 << (p & 7))
    return out


def knn_database(nt, nq, seed_t=42, seed_q=43):
    """config 4 inputs, SURVEY Appendix B row 4.  Train T: byte j of row i = draw(42, 32 i + j + 1) mod 256.  Queries: even i =
    T[draw(43, 2 i + 1) mod nt] with k_i = draw(43, 2 i + 2) mod 64 DISTINCT bit flips, the positions being the first k_i distinct
    values of draw(44, 128 i + j) mod 256, j = 1, 2, ... (a block of 128 draws of stream 44 per query: 63 distinct positions out of
    256 need more than 128 draws with probability < 1e-9; asserted); odd i = uniform, byte j = draw(45, 32 i + j + 1) mod 256.
    Unrelated pairs are ~Binomial(256, 1/2): many exact distance ties, which exercise the (distance, global index) rule."""
    T = random_descriptors(seed_t, nt)
    Q = random_descriptors(seed_q + 2, nq)
    ev = np.arange(0, nq, 2)
    src = (draw(seed_q, 2 * ev.astype(np.uint64) + np.uint64(1)) % np.uint64(nt)).astype(np.int64)
    kf = (draw(seed_q, 2 * ev.astype(np.uint64) + np.uint64(2)) % np.uint64(64)).astype(np.int64)
    Q[ev] = T[src]
    mask = np.zeros((len(ev), 32), np.uint8)                  # bit positions flipped so far (= the XOR mask)
    cnt = np.zeros(len(ev), np.int64)
    base = ev.astype(np.uint64) * np.uint64(128)
    rows = np.arange(len(ev))
    for j in range(1, 129):
        if not (cnt < kf).any():
            break
        p = (draw(seed_q + 1, base + np.uint64(j)) % np.uint64(256)).astype(np.int64)
        bit = (1 << (p & 7)).astype(np.uint8)
        take = (cnt < kf) & ((mask[rows, p >> 3] & bit) == 0)
        mask[rows[take], p[take] >> 3] |= bit[take]
        cnt += take
    assert (cnt == kf).all(), 'a query needed more than 128 draws for its distinct bit positions'
    Q[ev] ^= mask
    return T, Q


def projection_case(seed_f=3, seed_p=4, nk=2000, nq=10000, W=752, H=480):
    """cfg3: a 2000-keypoint frame and 10 000 projected map points."""
    quota = np.array([434, 362, 302, 251, 209, 175, 145, 122], np.int64)
    cum = np.cumsum(quota)
    k = np.arange(nk, dtype=np.uint64)
    d = lambda s, j, n, m: draw(s, np.arange(n, dtype=np.uint64) * np.uint64(m) + np.uint64(j + 1))
    kx = 16 + (d(seed_f, 0, nk, 8) % np.uint64(W - 32)).astype(np.float32) + (d(seed_f, 1, nk, 8) % np.uint64(1000)).astype(np.float32) / np.float32(1000)
    ky = 16 + (d(seed_f, 2, nk, 8) % np.uint64(H - 32)).astype(np.float32) + (d(seed_f, 3, nk, 8) % np.uint64(1000)).astype(np.float32) / np.float32(1000)
    octave = np.searchsorted(cum, (d(seed_f, 4, nk, 8) % np.uint64(cum[-1])).astype(np.int64), side='right').astype(np.int32)
    kangle = ((d(seed_f, 5, nk, 8) % np.uint64(36000)).astype(np.float32) / np.float32(100)).astype(np.float32)
    kdesc = random_descriptors(seed_f + 1000, nk)
    # map points
    u = np.zeros(nq, np.float32); v = np.zeros(nq, np.float32); lvl = np.zeros(nq, np.int32)
    qangle = np.zeros(nq, np.float32); qdesc = random_descriptors(seed_p + 1000, nq)
    r = d(seed_p, 0, nq, 16)
    near = (r % np.uint64(10)) < 6
    src = (d(seed_p, 1, nq, 16) % np.uint64(nk)).astype(np.int64)
    offx = ((d(seed_p, 2, nq, 16) % np.uint64(6001)).astype(np.float32) - 3000) / np.float32(1000)
    offy = ((d(seed_p, 3, nq, 16) % np.uint64(6001)).astype(np.float32) - 3000) / np.float32(1000)
    u[:] = np.where(near, kx[src] + offx, (d(seed_p, 4, nq, 16) % np.uint64(W)).astype(np.float32))
    v[:] = np.where(near, ky[src] + offy, (d(seed_p, 5, nq, 16) % np.uint64(H)).astype(np.float32))
    up = (d(seed_p, 6, nq, 16) % np.uint64(2)).astype(np.int32)
    lvl[:] = np.where(near, np.minimum(octave[src] + up, 7), (d(seed_p, 7, nq, 16) % np.uint64(8)).astype(np.int32))
    qangle[:] = np.where(near, kangle[src] + ((d(seed_p, 8, nq, 16) % np.uint64(21)).astype(np.float32) - 10),
                         (d(seed_p, 9, nq, 16) % np.uint64(36000)).astype(np.float32) / np.float32(100))
    qangle = np.mod(qangle, np.float32(360)).astype(np.float32)
    nflip = (d(seed_p, 10, nq, 16) % np.uint64(31)).astype(np.int64)
    nd = kdesc[src].copy()
    for j in range(30):
        sel = near & (nflip > j)
        p = (draw(seed_p + 7, np.arange(nq, dtype=np.uint64) * np.uint64(32) + np.uint64(j + 1)) % np.uint64(256)).astype(np.int64)
        rows = np.nonzero(sel)[0]
        nd[rows, p[rows] >> 3] ^= (1 << (p[rows] & 7)).astype(np.uint8)
    qdesc[near] = nd[near]
    view_cos = np.where(np.arange(nq) % 2 == 0, np.float32(0.9990), np.float32(0.9)).astype(np.float32)
    return dict(kx=kx.astype(np.float32), ky=ky.astype(np.float32), octave=octave, kangle=kangle, kdesc=kdesc,
                u=u, v=v, level=lvl, qangle=qangle, qdesc=qdesc, view_cos=view_cos,
                bounds=(0.0, float(W), 0.0, float(H)))


def synthetic_vocabulary(k=10, L=3, seed=77, ragged=False):
    """a DBoW2-shaped vocabulary tree with random 256-bit node descriptors (children = parent with ~40 bit flips, so that
    descents are meaningful) and idf-like weights; ragged=True drops some children and ends some branches early.
    Returns the flat tree (see frontend.ORBVocabulary) and the lines of the equivalent DBoW2 text file."""
    rng = np.random.default_rng(seed)
    parent, level, desc = [0], [0], [rng.integers(0, 256, 32, dtype=np.uint8)]
    frontier = [0]
    for lv in range(1, L + 1):
        nxt = []
        for p in frontier:
            nk = k if not ragged else int(rng.integers(max(1, k // 2), k + 1))
            if ragged and lv > 1 and rng.random() < 0.15:
                continue                                   # this branch ends early: p stays a leaf
            for _ in range(nk):
                d = desc[p].copy()
                for b in rng.integers(0, 256, 40):
                    d[b >> 3] ^= np.uint8(1 << (b & 7))
                parent.append(p); level.append(lv); desc.append(d); nxt.append(len(parent) - 1)
        frontier = nxt
    n = len(parent)
    children = [[] for _ in range(n)]
    for i in range(1, n):
        children[parent[i]].append(i)
    cs = np.zeros(n + 1, np.int32); cs[1:] = np.cumsum([len(c) for c in children])
    ci = np.array([c for ch in children for c in ch], np.int32)
    leaf = np.array([len(c) == 0 for c in children]); leaf[0] = False
    word = np.full(n, -1, np.int32); word[leaf] = np.arange(int(leaf.sum()))
    weight = np.where(leaf, np.round(rng.uniform(0.0, 9.0, n), 5), 0.0)
    weight[leaf & (rng.random(n) < 0.03)] = 0.0               # a few stopped words
    lines = ['%d %d 0 0' % (k, L)]
    for i in range(1, n):
        lines.append('%d %d %s %s' % (parent[i], int(leaf[i]), ' '.join(str(int(v)) for v in desc[i]), repr(float(weight[i]))))
    tree = dict(k=k, L=L, scoring=0, weighting=0, child_start=cs, child_ids=ci, desc=np.stack(desc), weight=weight.astype(np.float64), word=word)
    return tree, lines
