#!/usr/bin/env python
"""Exhaustive random check that the closed-form permutation used by the GPU builder (build_sah.cu partition_dest)
equals the reference's sequential swap partition (tiny_bvh.h:2414-2422)."""
import numpy as np
def seq(A, left):
    A=A.copy(); n=len(A); s=0; j=n
    for _ in range(n):
        if left[A[s]]: s+=1
        else:
            j-=1; A[s],A[j]=A[j],A[s]
    return A, s
def par(A, left):
    n=len(A); fl = left[A]
    L = int(fl.sum())
    out = np.empty_like(A)
    F = np.arange(L); B = np.arange(n-1, L-1, -1)
    fr = F[~fl[F]]; bl = B[fl[B]]
    m = len(fr); assert len(bl)==m
    out[F[fl[F]]] = A[F[fl[F]]]
    out[fr] = A[bl]
    extra = (L < n) and (not fl[L])      # FR_m = position L when it holds a right
    frx = np.concatenate([fr, [L]]) if extra else fr
    mx = len(frx)
    if mx:
        dest = np.empty(mx, int); dest[0]=n-1
        dest[1:] = bl[:mx-1]-1
        out[dest] = A[frx]
    bidx = np.arange(len(B)); isl = fl[B]
    l_before = np.cumsum(isl) - isl
    br = ~isl
    if extra: br[-1] = False
    r = np.minimum(l_before+1, mx) + bidx - l_before
    out[(n-1-r)[br]] = A[B[br]]
    return out, L
rng=np.random.default_rng(0)
for trial in range(200000):
    n = rng.integers(1,40)
    A = rng.permutation(100)[:n]
    left = rng.random(100) < rng.random()
    a,s = seq(A,left); b,L = par(A,left)
    assert s==L and np.array_equal(a,b), (A,left[A],a,b)
print("parallel partition formula == sequential swap partition")
