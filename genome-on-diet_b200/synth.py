"""Seeded synthetic workloads of the shapes BASELINE.json names. numpy only.

Sequences are byte codes as the reference's DP call sites produce them (0..3 = ACGT, 4 = N,
7 = reverse-complemented N; GDiet-ShortReads/map.c:737-757) or ASCII for the sketch entry points.
"""
import numpy as np

ACGTN = np.frombuffer(b"ACGTN", np.uint8)

# scoring presets (GDiet-ShortReads/options.c:45-47,106,134-136)
SCORING = {
    "sr": dict(a=2, b=8, q=12, e=2, q2=24, e2=1, zdrop=100, end_bonus=10),
    "map-hifi": dict(a=1, b=4, q=6, e=2, q2=26, e2=1, zdrop=400, end_bonus=0),
    "map-ont": dict(a=2, b=4, q=4, e=2, q2=24, e2=1, zdrop=400, end_bonus=0),
}


def score_matrix(a, b):
    """5x5 matrix exactly as the live call site builds it (GDiet-ShortReads/map.c:861-865)."""
    bb = -abs(int(b))
    m = np.full((5, 5), bb, np.int8)
    np.fill_diagonal(m, a)
    m[4, :] = 0
    m[:, 4] = 0
    return np.ascontiguousarray(m.reshape(-1))


def random_codes(rng, n):
    return rng.integers(0, 4, n, dtype=np.uint8)


def mutate_codes(rng, src, rate, sub=0.6, dele=0.2):
    """Apply `rate` edits per base: 60 % substitutions, 20 % deletions, 20 % insertions."""
    n = len(src)
    x = rng.random(n)
    is_sub = x < rate * sub
    is_del = (x >= rate * sub) & (x < rate * (sub + dele))
    is_ins = (x >= rate * (sub + dele)) & (x < rate)
    out = src.copy()
    out[is_sub] = (out[is_sub] + rng.integers(1, 4, int(is_sub.sum()), dtype=np.uint8)) & 3
    reps = np.ones(n, np.int64)
    reps[is_del] = 0
    reps[is_ins] = 2
    out = np.repeat(out, reps)
    # the second copy of an inserted base becomes a random base
    ins_pos = np.cumsum(reps)[is_ins] - 1
    out[ins_pos] = rng.integers(0, 4, len(ins_pos), dtype=np.uint8)
    return out


def ksw_pairs(n, qlen=150, tlen=200, edit=0.05, seed=3, n_every=50, high_edit_frac=0.0, high_edit=0.30):
    """Config 2 of BASELINE.json: targets of `tlen` random bases; queries of exactly `qlen` bases derived
    from the target prefix with `edit` edits; every `n_every`-th query gets one N. Returns a dict of
    flat arrays (qbuf, qoff, qlen, tbuf, toff, tlen)."""
    rng = np.random.default_rng(seed)
    tbuf = random_codes(rng, n * tlen)
    qbuf = np.empty(n * qlen, np.uint8)
    n_hi = int(n * high_edit_frac)
    for i in range(n):
        t = tbuf[i * tlen:(i + 1) * tlen]
        rate = high_edit if i < n_hi else edit
        q = mutate_codes(rng, t, rate)
        if len(q) < qlen:
            q = np.concatenate([q, random_codes(rng, qlen - len(q))])
        q = q[:qlen]
        if n_every and i % n_every == n_every - 1:
            q = q.copy()
            q[int(rng.integers(0, qlen))] = 4
        qbuf[i * qlen:(i + 1) * qlen] = q
    return dict(
        n=n,
        qbuf=qbuf, qoff=np.arange(n, dtype=np.int64) * qlen, qlen=np.full(n, qlen, np.int32),
        tbuf=tbuf, toff=np.arange(n, dtype=np.int64) * tlen, tlen=np.full(n, tlen, np.int32),
    )


def ksw_pairs_fast(n, qlen=150, tlen=200, edit=0.05, seed=3, n_every=50):
    """Vectorised variant of ksw_pairs for millions of pairs: substitutions are applied per base; one
    deletion and one insertion event are placed per pair with probability proportional to `edit`, which
    gives the same mix of band-touching alignments without a Python loop."""
    rng = np.random.default_rng(seed)
    t = rng.integers(0, 4, (n, tlen), dtype=np.uint8)
    q = t[:, :qlen + 8].copy()
    sub = rng.random((n, qlen + 8)) < edit * 0.6
    q[sub] = (q[sub] + rng.integers(1, 4, int(sub.sum()), dtype=np.uint8)) & 3
    # indels: shift the tail of the query left (deletion) or right (insertion) at a random position
    n_del = rng.binomial(qlen, edit * 0.2, n)
    n_ins = rng.binomial(qlen, edit * 0.2, n)
    idx = np.arange(qlen + 8)[None, :]
    for k in range(int(max(n_del.max(initial=0), n_ins.max(initial=0)))):
        pos = rng.integers(1, qlen - 1, n)[:, None]
        do_del = (n_del > k)[:, None]
        src = np.where(do_del & (idx >= pos), np.minimum(idx + 1, qlen + 7), idx)
        q = np.take_along_axis(q, src, 1)
        pos = rng.integers(1, qlen - 1, n)[:, None]
        do_ins = (n_ins > k)[:, None]
        src = np.where(do_ins & (idx > pos), idx - 1, idx)
        q = np.take_along_axis(q, src, 1)
        newb = rng.integers(0, 4, n, dtype=np.uint8)
        rows = np.nonzero(do_ins[:, 0])[0]
        q[rows, pos[rows, 0]] = newb[rows]
    q = np.ascontiguousarray(q[:, :qlen])
    if n_every:
        rows = np.arange(n_every - 1, n, n_every)
        q[rows, rng.integers(0, qlen, len(rows))] = 4
    return dict(
        n=n,
        qbuf=q.reshape(-1), qoff=np.arange(n, dtype=np.int64) * qlen, qlen=np.full(n, qlen, np.int32),
        tbuf=t.reshape(-1), toff=np.arange(n, dtype=np.int64) * tlen, tlen=np.full(n, tlen, np.int32),
    )


def ragged_pairs(n, seed=11, max_len=260, with_n=True):
    """Mixed-length pairs for parity tests: lengths 1..max_len, edits 2-30 %, N and code-7 bases."""
    rng = np.random.default_rng(seed)
    qs, ts = [], []
    for _ in range(n):
        tl = int(rng.integers(1, max_len))
        t = random_codes(rng, tl)
        src = t[: max(1, int(tl * rng.uniform(0.5, 1.0)))]
        q = mutate_codes(rng, src, float(rng.choice([0.02, 0.05, 0.3])))
        if len(q) == 0:
            q = np.array([1], np.uint8)
        if with_n:
            if rng.random() < 0.3:
                q[int(rng.integers(0, len(q)))] = 4
            if rng.random() < 0.2:
                t = t.copy()
                t[int(rng.integers(0, tl))] = 4
            if rng.random() < 0.1:
                q[int(rng.integers(0, len(q)))] = 7
        qs.append(q)
        ts.append(t)
    return pack_pairs(qs, ts)


def pack_pairs(qs, ts):
    qlen = np.array([len(x) for x in qs], np.int32)
    tlen = np.array([len(x) for x in ts], np.int32)
    qoff = np.zeros(len(qs), np.int64)
    toff = np.zeros(len(ts), np.int64)
    if len(qs) > 1:
        qoff[1:] = np.cumsum(qlen[:-1])
        toff[1:] = np.cumsum(tlen[:-1])
    qbuf = np.concatenate(qs).astype(np.uint8) if len(qs) else np.zeros(0, np.uint8)
    tbuf = np.concatenate(ts).astype(np.uint8) if len(ts) else np.zeros(0, np.uint8)
    return dict(n=len(qs), qbuf=qbuf, qoff=qoff, qlen=qlen, tbuf=tbuf, toff=toff, tlen=tlen)


def long_pairs(n, qlen, edit, seed, tlen_extra=0.0):
    """HiFi / ONT-like gap-filling shapes: query of `qlen` bases, target = the unmutated source region."""
    rng = np.random.default_rng(seed)
    qs, ts = [], []
    for _ in range(n):
        tl = int(qlen * (1.0 + tlen_extra))
        t = random_codes(rng, tl)
        q = mutate_codes(rng, t, edit, sub=0.4, dele=0.3)[:qlen]
        qs.append(q)
        ts.append(t)
    return pack_pairs(qs, ts)


def random_genome(n, seed, n_frac=0.0):
    rng = np.random.default_rng(seed)
    codes = rng.integers(0, 4, n, dtype=np.uint8)
    if n_frac > 0:
        codes[rng.random(n) < n_frac] = 4
    return ACGTN[codes]


def sample_reads(genome, n_reads, read_len, seed, sub=0.01, indel=0.002):
    """Config 1/5 reads: uniform positions, 50 % reverse-complemented, substitutions (+ rare indels handled
    as substitutions of the flanking base so that every read stays exactly read_len long)."""
    rng = np.random.default_rng(seed)
    start = rng.integers(0, len(genome) - read_len, n_reads)
    idx = start[:, None] + np.arange(read_len)[None, :]
    reads = genome[idx]
    lut = np.zeros(256, np.uint8)
    lut[ACGTN] = np.arange(5)
    codes = lut[reads]
    m = (rng.random(codes.shape) < (sub + indel)) & (codes < 4)
    codes[m] = (codes[m] + rng.integers(1, 4, int(m.sum()), dtype=np.uint8)) & 3
    rev = rng.random(n_reads) < 0.5
    rc = np.where(codes[rev] < 4, 3 - codes[rev], 4)[:, ::-1]
    codes[rev] = rc
    return ACGTN[codes]
