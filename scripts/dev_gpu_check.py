"""Developer smoke script (run under gpurun): correctness of the scan/top-k kernels against numpy on
small seeded inputs plus a first throughput number. Not part of the product path."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "visual-rag-toolkit_b200"))
from visual_rag_b200.corpus import GpuCorpus  # noqa: E402


def maxsim_np(q, d, normalize=True):
    q = q.astype(np.float32)
    d = d.astype(np.float32)
    if normalize:
        q = q / (np.linalg.norm(q, axis=1, keepdims=True) + 1e-8)
        d = d / (np.linalg.norm(d, axis=1, keepdims=True) + 1e-8)
    return float((q @ d.T).max(axis=1).sum())


def make_rows(rng, n):
    x = rng.standard_normal((n, 128)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    x *= rng.uniform(0.5, 2.0, size=(n, 1)).astype(np.float32)  # non-unit norms exercise the scale path
    return x.astype(np.float16)


def check(name, got, want, tol=2e-5):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    err = np.abs(got - want) / np.maximum(np.abs(want), 1e-6)
    ok = bool(np.all(err < tol))
    print(f"[{'OK' if ok else 'FAIL'}] {name}: max rel err {err.max():.3e} (n={got.size})", flush=True)
    return ok


def main():
    rng = np.random.default_rng(0)
    ok = True
    c = GpuCorpus(0)
    for Q in (20, 1, 8, 33, 100):
        q = rng.standard_normal((Q, 128)).astype(np.float32)
        # LARGE fixed
        for T in (300, 1030, 129):
            n = 37
            rows = make_rows(rng, n * T)
            c.add_store("s", rows, fixed_rows=T)
            got = c.score("s", q)
            want = [maxsim_np(q, rows[i * T:(i + 1) * T]) for i in range(n)]
            ok &= check(f"large fixed T={T} Q={Q}", got, want)
        # LARGE variable
        lens = rng.integers(1, 700, size=53)
        lens[5] = 640
        off = np.concatenate([[0], np.cumsum(lens)])
        rows = make_rows(rng, int(off[-1]))
        c.add_store("s", rows, page_offsets=off)
        got = c.score("s", q)
        want = [maxsim_np(q, rows[off[i]:off[i + 1]]) for i in range(len(lens))]
        ok &= check(f"large variable Q={Q}", got, want)
        got_nn = c.score("s", q, normalize=False)
        want_nn = [maxsim_np(q, rows[off[i]:off[i + 1]], normalize=False) for i in range(len(lens))]
        ok &= check(f"large variable no-normalize Q={Q}", got_nn, want_nn, tol=1e-4)
        # candidates on LARGE
        cand = rng.permutation(len(lens))[:17]
        got = c.score("s", q, candidate_ids=cand)
        ok &= check(f"large candidates Q={Q}", got, [want[i] for i in cand])
        # PACKED fixed
        for R in (32, 13, 76, 1, 34, 128, 64):
            n = 1001
            rows = make_rows(rng, n * R)
            c.add_store("p", rows, fixed_rows=R)
            got = c.score("p", q)
            want = [maxsim_np(q, rows[i * R:(i + 1) * R]) for i in range(n)]
            ok &= check(f"packed fixed R={R} Q={Q}", got, want)
            cand = rng.permutation(n)[:77]
            got = c.score("p", q, candidate_ids=cand)
            ok &= check(f"packed candidates R={R} Q={Q}", got, [want[i] for i in cand])
            gp = c.score("p", q, pool_query=True)
            qb = q.mean(axis=0, keepdims=True)
            wantp = [maxsim_np(qb, rows[i * R:(i + 1) * R]) for i in range(n)]
            ok &= check(f"packed pooled-query R={R} Q={Q}", gp, wantp)
        # PACKED variable
        lens = rng.integers(1, 33, size=777)
        off = np.concatenate([[0], np.cumsum(lens)])
        rows = make_rows(rng, int(off[-1]))
        c.add_store("p", rows, page_offsets=off)
        got = c.score("p", q)
        want = np.array([maxsim_np(q, rows[off[i]:off[i + 1]]) for i in range(len(lens))])
        ok &= check(f"packed variable Q={Q}", got, want)
        # top-k
        for k in (1, 10, 256, 777, 1000):
            s, ids = c.search("p", q, k)
            order = np.lexsort((np.arange(len(want)), -got))[:k]
            good = np.array_equal(ids, order) and np.array_equal(s, got[order])
            print(f"[{'OK' if good else 'FAIL'}] topk k={k} Q={Q}", flush=True)
            ok &= good
    # large top-k (multi-level)
    n = 300000
    c.add_synthetic_store("g", n, fixed_rows=1, seed=1)
    q = rng.standard_normal((1, 128)).astype(np.float32)
    sc = c.score("g", q)
    rows = c.read_rows("g", 0, n)
    want = (rows.astype(np.float32) / (np.linalg.norm(rows.astype(np.float32), axis=1, keepdims=True) + 1e-8)) @ (
        q[0] / (np.linalg.norm(q[0]) + 1e-8))
    ok &= check("global store 300k", sc, want, tol=1e-4)
    for k in (10, 1000, 4096):
        s, ids = c.search("g", q, k)
        order = np.lexsort((np.arange(n), -sc))[:k]
        good = np.array_equal(ids, order) and np.array_equal(s, sc[order])
        print(f"[{'OK' if good else 'FAIL'}] topk-large k={k}", flush=True)
        ok &= good
    # multistage
    n = 5000
    c.add_synthetic_store("initial", n, fixed_rows=300, seed=2)
    c.add_synthetic_store("mean_pooling", n, fixed_rows=32, seed=3)
    q = rng.standard_normal((20, 128)).astype(np.float32)
    st = c.search_multistage([("mean_pooling", False, 256), ("initial", False, 10)], q)
    s1 = c.score("mean_pooling", q)
    o1 = np.lexsort((np.arange(n), -s1))[:256]
    s2 = c.score("initial", q, candidate_ids=o1)
    o2 = np.lexsort((np.arange(256), -s2))[:10]
    good = np.array_equal(st[0][1], o1) and np.array_equal(st[1][1], o1[o2]) and np.allclose(st[1][0], s2[o2])
    print(f"[{'OK' if good else 'FAIL'}] multistage", flush=True)
    ok &= good

    # throughput
    for n_pages, T in ((100000, 1030), (100000, 768), (1000000, 32)):
        c.add_synthetic_store("big", n_pages, fixed_rows=T, seed=7)
        q = rng.standard_normal((20, 128)).astype(np.float32)
        for _ in range(3):
            c.search("big", q, 10)
        ts = []
        for _ in range(10):
            c.search("big", q, 10)
            ts.append(c.last_timing_ms())
        tot = np.median([t[0] for t in ts])
        ker = np.median([t[1] for t in ts])
        gb = n_pages * T * 256 / 1e9
        print(f"scan {n_pages}x{T}: total {tot:.3f} ms, kernel {ker:.3f} ms, {gb / ker * 1e3:.0f} GB/s "
              f"({gb / ker * 1e3 / 6549.8:.2%} of measured HBM peak), {n_pages / ker * 1e3 / 1e6:.2f} M pages/s", flush=True)
        c.drop_store("big")
    print("ALL OK" if ok else "SOME FAILED", flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
