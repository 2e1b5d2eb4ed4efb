"""The closed-form permutation the GPU builder uses for the reference's in-place swap partition (build_sah.cu
partition_dest) equals the sequential loop of tiny_bvh.h:2414-2422.  CPU only (numpy model of the kernel's formula)."""
import numpy as np


def sequential(A, left):
    A = A.copy()
    n, s, j = len(A), 0, len(A)
    for _ in range(n):
        if left[A[s]]:
            s += 1
        else:
            j -= 1
            A[s], A[j] = A[j], A[s]
    return A, s


def closed_form(A, left):
    """Scatter form, element by element, exactly as k_scatter / k_build_small evaluate it."""
    n = len(A)
    fl = left[A]
    L = int(fl.sum())
    scan = np.concatenate([[0], np.cumsum(fl)])
    pos_bl = np.zeros(n + 1, np.int64)
    for p in range(L, n):
        if fl[p]:
            pos_bl[scan[n] - scan[p + 1]] = p
    lefts_in_F = scan[L]
    m = L - lefts_in_F
    extra = L < n and not fl[L]
    out = np.full(n, -1, A.dtype)
    for rel in range(n):
        before = scan[rel]
        if rel < L:
            if fl[rel]:
                out[rel] = A[rel]
            else:
                k = rel - before
                out[rel] = A[pos_bl[k]]
                out[n - 1 if k == 0 else pos_bl[k - 1] - 1] = A[rel]
        elif fl[rel]:
            pass
        elif rel == L:
            out[n - 1 if m == 0 else pos_bl[m - 1] - 1] = A[rel]
        else:
            l = L - before
            mx = m + (1 if extra else 0)
            out[n - 1 - (min(l + 1, mx) + (n - 1 - rel) - l)] = A[rel]
    return out, L


def test_closed_form_equals_sequential_partition():
    rng = np.random.default_rng(0)
    for _ in range(30000):
        n = int(rng.integers(1, 70))
        A = rng.permutation(128)[:n]
        left = rng.random(128) < rng.random()
        a, s = sequential(A, left)
        b, L = closed_form(A, left)
        assert s == L and np.array_equal(a, b)
