"""TEST INFRASTRUCTURE: numpy restatement of gnn-mwvc_b200/csrc/gvc_px.cuh -- the reference's sequential
fp32 neighbour sum (src/gnn_inference.cpp:33-36) computed batch-parallel and still bit for bit: per-batch
approximate prefix sums predict the binade of the running sum, every addend is rounded to that binade's
grid (exact, order-free sums), and a final in-order pass verifies each prediction and falls back to the
element-wise chain for the batches that need it.  tests/test_px_model.py checks it against the chain."""
import numpy as np
f32 = np.float32

def seq_sum(v):
    acc = f32(0)
    for x in v: acc = f32(acc + x)
    return acc

def expo(x):   # floor(log2(x)) for positive normal float32, from the bits
    return ((np.asarray(x, f32).view(np.uint32) >> 23) & 0xFF).astype(np.int32) - 127

def px_sum(v, BATCH=64, delta=1.0 / 4096, stats=None):
    v = np.asarray(v, f32)
    n = len(v); nb = (n + BATCH - 1) // BATCH
    # pass A: approximate batch sums (any order), prefix in double
    S = np.array([v[b*BATCH:(b+1)*BATCH].astype(np.float64).sum() for b in range(nb)])
    P = np.concatenate([[0.0], np.cumsum(S)])          # predicted entry value of batch b = P[b]
    # pass B: per batch, under the predicted entry binade
    D = np.zeros(nb, f32); clean = np.zeros(nb, bool); E = np.zeros(nb, np.int32)
    for b in range(nb):
        vb = v[b*BATCH:(b+1)*BATCH]
        if not np.any(vb != 0):                           # an all-zero batch adds nothing whatever the running sum is
            D[b] = 0; clean[b] = True; E[b] = -999
            continue
        lo, hi = P[b] * (1 - delta), P[b+1] * (1 + delta)
        if not (lo >= 2.0**-125 and hi >= lo and hi < 2.0**127):              # zero / tiny / NaN / decreasing prefix: no prediction
            continue
        e = int(np.floor(np.log2(lo)))
        if int(np.floor(np.log2(hi))) != e:              # a power of two inside the (widened) interval
            continue
        M = f32(2.0**e); u_half = f32(2.0**(e-24))
        s = (M + vb).astype(f32)                          # RN(M + v)
        d = (s - M).astype(f32)
        t = (vb - d).astype(f32)
        ok = np.all((vb >= 0) & (vb < M)) and not np.any(np.abs(t) == u_half)
        if not ok: continue
        D[b] = d.astype(np.float64).sum()                 # exact: multiples of u, total < 2^24 u  (checked at composition)
        clean[b] = True; E[b] = e
    # pass C: sequential composition with verification
    acc = f32(0); slow = 0
    for b in range(nb):
        good = clean[b] and (E[b] == -999 or (acc > 0 and expo(acc) == E[b]))
        if good:
            nxt = f32(acc + D[b])
            good = E[b] == -999 or nxt < f32(2.0**(E[b]+1))
        if good:
            acc = nxt
        else:
            slow += 1
            for x in v[b*BATCH:(b+1)*BATCH]: acc = f32(acc + x)
    if stats is not None: stats.append((nb, slow))
    return acc

