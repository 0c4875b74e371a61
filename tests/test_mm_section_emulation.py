"""CPU restatements (numpy) of the index logic of the round-2d kernels -- no GPU, no oracle: they pin the schedules the CUDA code follows.

* k_chol_cluster (csrc/dense_mm.cuh): right-looking blocked Cholesky with in-place panels and trailing blocks updated in ANY order, then
  the explicit inverse by recursive doubling on 64 x 64 blocks with the work-item decode (heaviest first), the transposed W scratch and
  the L^-T blocks written beside L^-1;
* k_mm64 (csrc/gemm_mm64.cuh): the folded grid is a permutation of the work items, k ranges clipped at 64 only skip exact zeros;
* the fragment-layout moments epilogue of k_gemm_i8 (csrc/gemm_i8.cuh): the tcgen05.ld.16x256b register layout covers the warp's 32 x 32
  patch exactly once and the DMMA k-permutation (column 8 j + 2 q + e plays k = q) reproduces W @ Phi.
"""
import numpy as np
import pytest

NB = 8   # block size of the emulation (the kernels use 64; the index logic does not depend on it)


def _spd(n, seed):
    r = np.random.default_rng(seed).standard_normal((n, n + 5))
    return r @ r.T / n + np.eye(n)


def _blk(a, i, j):
    return a[i * NB:(i + 1) * NB, j * NB:(j + 1) * NB]


def cluster_factor(A, rng):
    """k_chol_cluster, factorisation part: returns (A with L in the lower blocks, T = block inverses)."""
    A = A.copy()
    nblk = A.shape[0] // NB
    T = [None] * nblk

    def potf2(k):
        L = np.linalg.cholesky(np.tril(_blk(A, k, k)) + np.tril(_blk(A, k, k), -1).T)
        _blk(A, k, k)[:] = L
        T[k] = np.linalg.inv(L)

    potf2(0)
    for k in range(nblk - 1):
        for i in range(k + 1, nblk):                       # panels, in place
            _blk(A, i, k)[:] = _blk(A, i, k) @ T[k].T
        nb = nblk - k - 1
        items = list(range(1, nb * (nb + 1) // 2))
        rng.shuffle(items)                                  # whoever draws a block from the atomic counter
        for w in [0] + items:                               # block 0 (the next diagonal block) first, by rank 0
            bi = 0
            while (bi + 1) * (bi + 2) // 2 <= w:
                bi += 1
            bj = w - bi * (bi + 1) // 2
            i, j = k + 1 + bi, k + 1 + bj
            _blk(A, i, j)[:] -= _blk(A, i, k) @ _blk(A, j, k).T
            if w == 0:
                potf2(k + 1)
    return A, T


def cluster_inverse(A, T):
    """k_chol_cluster, inverse part: Linv, LinvT from the factored A and the block inverses, item decode as in the kernel."""
    n = A.shape[0]
    nblk = n // NB
    A = np.tril(A)
    Linv, LinvT, Wk = np.zeros((n, n)), np.zeros((n, n)), np.full((n, n), np.nan)
    for k in range(nblk):
        _blk(Linv, k, k)[:] = T[k]
        _blk(LinvT, k, k)[:] = T[k].T
    sb = 1
    while sb < nblk:
        npairs, items = nblk // (2 * sb), (nblk // (2 * sb)) * sb * sb
        for phase in (0, 1):
            weights = []
            for w in range(items):
                pr, r = w % npairs, w // npairs
                b0 = 2 * sb * pr
                i = r % sb if phase == 0 else sb - 1 - r // sb
                j = r // sb if phase == 0 else r % sb
                k_lo, k_hi = (j, sb - 1) if phase == 0 else (0, i)
                weights.append(k_hi - k_lo + 1)
                acc = np.zeros((NB, NB))
                for k in range(k_lo, k_hi + 1):
                    if phase == 0:   # sa = L(b0+sb+i, b0+k), sb[c][kk] = LinvT(b0+j, b0+k)
                        acc += _blk(A, b0 + sb + i, b0 + k) @ _blk(LinvT, b0 + j, b0 + k).T
                    else:            # sa = Linv(b0+sb+i, b0+sb+k), sb[c][kk] = WkT(b0+j, b0+sb+k)
                        acc += _blk(Linv, b0 + sb + i, b0 + sb + k) @ _blk(Wk, b0 + j, b0 + sb + k).T
                if phase == 0:
                    _blk(Wk, b0 + j, b0 + sb + i)[:] = acc.T
                else:
                    _blk(Linv, b0 + sb + i, b0 + j)[:] = -acc
                    _blk(LinvT, b0 + j, b0 + sb + i)[:] = -acc.T
            per_slice = weights[::npairs]
            assert per_slice == sorted(per_slice, reverse=True)     # heaviest first (ties in any order)
        sb *= 2
    return Linv, LinvT


@pytest.mark.parametrize("nblk", [4, 8, 16])
def test_cluster_resident_cholesky_and_block_inverse(nblk):
    rng = np.random.default_rng(nblk)
    A0 = _spd(nblk * NB, seed=nblk)
    A, T = cluster_factor(A0, rng)
    L = np.tril(A)
    assert np.abs(L @ L.T - A0).max() < 1e-12 * np.abs(A0).max()
    assert np.abs(L - np.linalg.cholesky(A0)).max() < 1e-11
    Linv, LinvT = cluster_inverse(A, T)
    assert np.abs(Linv @ L - np.eye(nblk * NB)).max() < 1e-10
    assert np.array_equal(LinvT, Linv.T)
    assert np.abs(np.triu(Linv, 1)).max() == 0.0


@pytest.mark.parametrize("total,nsm", [(256, 148), (149, 148), (148, 148), (10, 148), (600, 148)])
def test_mm64_grid_fold_is_a_permutation(total, nsm):
    fold = nsm if total > nsm else 0
    seen = []
    for blk in range(total):
        w = blk
        if fold > 0 and w >= fold:
            w = total - 1 - (w - fold)
        seen.append(w)
    assert sorted(seen) == list(range(total))
    if fold:   # the first CTA of the second wave (same SM as CTA 0, the heaviest item) takes the lightest item
        assert seen[fold] == total - 1


def test_mm64_clipped_k_ranges_only_skip_zeros():
    KM_A_LOWER, KM_A_UPPER, KM_B_LOWER, KM_B_UPPER = 1, 2, 4, 8
    T, n = 4, 20                                   # tile size of the emulation, ragged matrix size
    rng = np.random.default_rng(0)
    F = rng.standard_normal((n, n))
    Lo, Up = np.tril(F), np.triu(F)
    nt = -(-n // T)
    for a, b, km in ((Lo, F, KM_A_LOWER), (Up, F, KM_A_UPPER), (F, Lo, KM_B_LOWER), (F, Up, KM_B_UPPER), (Up, Up, KM_A_UPPER | KM_B_UPPER),
                     (Lo, Lo, KM_A_LOWER | KM_B_LOWER)):
        C = np.zeros((n, n))
        for tm in range(nt):
            for tn in range(nt):
                k_lo, k_hi = 0, n
                if km & KM_A_LOWER: k_hi = min(k_hi, (tm + 1) * T)
                if km & KM_B_LOWER: k_hi = min(k_hi, (tn + 1) * T)
                if km & KM_A_UPPER: k_lo = max(k_lo, tm * T)
                if km & KM_B_UPPER: k_lo = max(k_lo, tn * T)
                if k_hi <= k_lo:
                    continue
                r, c = slice(tm * T, min(n, (tm + 1) * T)), slice(tn * T, min(n, (tn + 1) * T))
                C[r, c] = a[r, k_lo:k_hi] @ b[c, k_lo:k_hi].T
        assert np.abs(C - a @ b.T).max() < 1e-13, km


def test_fragment_layout_moments_reproduce_the_matrix_product():
    """tcgen05.ld.16x256b.x4 at lane offsets 0 and 16: thread 4 g + q holds, for h in {0, 1}, column group j, r2 in {0, 1}, e in {0, 1},
    element (row 16 h + 8 r2 + g, column 8 j + 2 q + e) of the warp's 32 x 32 patch at register index 16 h + 4 j + 2 r2 + e.  The moments
    feed those registers to mma.m8n8k4 as the A fragment (row g of row block G4 = 2 h + r2, k = q <-> column 8 j + 2 q + e) against the B
    fragment Phi[column 8 j + 2 q + e][moment 8 B + g]."""
    rng = np.random.default_rng(3)
    W = rng.standard_normal((32, 32))
    d = 8
    Phi = rng.standard_normal((32, 2 * d))                       # moments 1 .. 2 d of the warp's 32 columns
    regs = np.zeros((32, 32))                                     # [lane][register]
    cover = np.zeros((32, 32), dtype=int)
    for lane in range(32):
        g, q = lane >> 2, lane & 3
        for h in range(2):
            for j in range(4):
                for r2 in range(2):
                    for e in range(2):
                        row, col = 16 * h + 8 * r2 + g, 8 * j + 2 * q + e
                        regs[lane, 16 * h + 4 * j + 2 * r2 + e] = W[row, col]
                        cover[row, col] += 1
    assert (cover == 1).all()
    # DMMA m8n8k4: D[g][2 q', 2 q' + 1] += sum_k A[g][k] B[k][n]; lane 4 g + q supplies A[g][k = q] and B[k = q][n = g]
    mom = np.zeros((32, 2 * d))
    for G4 in range(4):
        for B in range(2 * d // 8):
            for j in range(4):
                for e in range(2):
                    Afrag = np.zeros((8, 4)); Bfrag = np.zeros((4, 8))
                    for lane in range(32):
                        g, q = lane >> 2, lane & 3
                        Afrag[g, q] = regs[lane, 16 * (G4 >> 1) + 4 * j + 2 * (G4 & 1) + e]
                        Bfrag[q, g] = Phi[8 * j + 2 * q + e, 8 * B + g]
                    mom[8 * G4:8 * G4 + 8, 8 * B:8 * B + 8] += Afrag @ Bfrag
    assert np.abs(mom - W @ Phi).max() < 1e-12
