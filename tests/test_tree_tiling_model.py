"""CPU model of tree_cta_kernel's tiling (parcompfin_b200/csrc/tree_kernels.cu): the same geometry formulas -- kOwn,
kC, stride, the halo exchange between warps, the d-power window index kbase - t >= -1 -- executed with numpy in place
of warps, against a plain backward induction. It pins the index arithmetic of the kernel on the CPU; the kernel itself
is checked bit for bit against the compiled reference by the -m gpu tests."""
import numpy as np
import pytest


def node(lo, hi, p, q, R):
    return (p * hi + q * lo) / R   # reference operation order, binom_vanilla_eur.cpp:35


def plain_tree(N, p, q, R, v, pu, pd, S0, sgn, nE, amer):
    for n in range(N - 1, -1, -1):
        nv = node(v[:n + 1], v[1:n + 2], p, q, R)
        if amer:
            i = np.arange(n + 1)
            nv = np.maximum(nv, np.maximum(sgn * ((S0 * pu[i]) * pd[n - i]) + nE, 0.0))
        v = nv
    return v[0]


def cta_launch(vin, n0, steps, kR, kH, kW, kK, p, q, R, pu, pd, S0, sgn, nE, amer):
    kL = 32 * kR
    kOwn = kL - kH
    kC = kW * kOwn + kH
    stride = kC - kK
    halo_lanes = kH // kR
    assert kH % kR == 0 and kC > kK >= kH
    n_out = n0 - steps
    vout = np.full(n_out + 1, np.nan)
    for b in range((n0 - steps + 1 + stride - 1) // stride):
        base = b * stride
        lo = n0 - steps - base - kC + 1
        win = steps + kC - 1
        s_pd = np.zeros(win + 1)                       # entry k at [k + 1]; [0] = 0 stands for k = -1
        idx = lo + np.arange(win)
        ok = (idx >= 0) & (idx <= n0)
        s_pd[1:][ok] = pd[idx[ok]]
        c0 = (np.arange(kW)[:, None] * kOwn + np.arange(32)[None, :] * kR)          # [warp, lane]
        i = base + c0[:, :, None] + np.arange(kR)[None, None, :]
        v = np.where(i <= n0, vin[np.minimum(i, n0)], 0.0)
        A = np.where(i <= n0, S0 * pu[np.minimum(i, n0)], 0.0)
        kbase = steps - 1 + kC - 1 - c0
        assert (kbase - (steps - 1 + kR) >= -1).all()  # the window index never drops below the stored zero
        W = np.stack([s_pd[1 + kbase - t] for t in range(kR)], axis=2)
        s_x = np.zeros((2, kW, kH))
        rounds = (steps + kH - 1) // kH
        for r in range(rounds):
            for ss in range(kH):
                s = r * kH + ss
                if s >= steps:
                    break
                halo = np.concatenate([v[:, 1:, 0], v[:, 31:, 0]], axis=1)           # shfl_down by one lane
                hi = np.concatenate([v[:, :, 1:], halo[:, :, None]], axis=2)
                nv = node(v, hi, p, q, R)
                if amer:
                    Wj = np.stack([W[:, :, (ss + j) % kR] for j in range(kR)], axis=2)
                    nv = np.maximum(nv, sgn * (A * Wj) + nE)                         # single max: nv >= 0 when p, q >= 0
                    W[:, :, ss % kR] = s_pd[1 + kbase - (s + kR)]
                v = nv
            if r + 1 < rounds:
                s_x[r & 1] = v[:, :halo_lanes, :].reshape(kW, kH)
                v[:-1, 32 - halo_lanes:, :] = s_x[r & 1][1:].reshape(kW - 1, halo_lanes, kR)
        own = (np.arange(32)[None, :, None] * kR + np.arange(kR)[None, None, :] < kOwn) & \
              (c0[:, :, None] + np.arange(kR)[None, None, :] < stride) & (i <= n_out)
        assert np.isnan(vout[i[own]]).all()            # every node is written by exactly one thread
        vout[i[own]] = v[own]
    assert not np.isnan(vout).any()
    return vout


@pytest.mark.parametrize("shape", [(2, 8, 2, 16), (1, 4, 3, 8), (4, 8, 2, 32), (3, 9, 2, 40), (6, 12, 2, 64), (8, 8, 1, 64)])
def test_cta_tiling_reproduces_the_plain_tree(shape):
    kR, kH, kW, kK = shape
    T, r, sig, S0, E = 1.0, 0.05, 0.2, 100.0, 100.0
    for N in (1, 5, 31, 97, 300, 533):
        dt = T / N
        beta = 0.5 * (np.exp(-r * dt) + np.exp((r + sig * sig) * dt))
        u, d = beta + np.sqrt(beta * beta - 1), beta - np.sqrt(beta * beta - 1)
        R = np.exp(r * dt)
        p = (R - d) / (u - d)
        q = 1 - p
        pu, pd = u ** np.arange(N + 1), d ** np.arange(N + 1)
        for amer in (False, True):
            for sgn in (1.0, -1.0):
                nE = -sgn * E
                i = np.arange(N + 1)
                v0 = np.maximum(sgn * ((S0 * pu[i]) * pd[N - i]) + nE, 0.0)
                want = plain_tree(N, p, q, R, v0.copy(), pu, pd, S0, sgn, nE, amer)
                v, n = v0, N
                while n > 0:
                    steps = min(kK, n)
                    v = cta_launch(v, n, steps, kR, kH, kW, kK, p, q, R, pu, pd, S0, sgn, nE, amer)
                    n -= steps
                assert v[0] == want, (shape, N, amer, sgn)
