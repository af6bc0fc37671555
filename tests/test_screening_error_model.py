"""CPU: a numpy FP32 model of the screening arithmetic of k_exh_screen / k_cand_screen (unit-normalised columns,
products accumulated along two half-row chains per descriptor row, rows combined by a tree, one division by the number of
valid column pairs) against the oracle's FP64 distDirectSC at the same shift.  The kernels select rescoring candidates
with |d32 - d| <= EXH_EPS = 1e-5 ("error budget of d32"); this model -- without FMA, i.e. with one more
rounding per product than the kernel -- must stay well inside it.  It is a check of the budget, not of the kernel
(the kernel is compared with the exact path on the device by tests/test_gpu_parity.py)."""
import numpy as np
import pytest

EXH_EPS = 1e-5


def _screen_model(a, b, R, S, shift):
    """a, b: column-major R*S float64 descriptors -> FP32 model of distDirectSC(a, circshift(b, shift))."""
    f32 = np.float32
    A = a.reshape(S, R).T.astype(f32)                                  # [R][S]
    B = np.roll(b.reshape(S, R), shift, axis=0).T.astype(f32)         # candidate shifted: column j <- column j - shift
    na = np.sqrt((a.reshape(S, R) ** 2).sum(axis=1))                   # FP64 norms, as stored
    nb = np.roll(np.sqrt((b.reshape(S, R) ** 2).sum(axis=1)), shift)
    ia = np.where(na == 0, 0, 1.0 / np.where(na == 0, 1, na)).astype(f32)
    ib = np.where(nb == 0, 0, 1.0 / np.where(nb == 0, 1, nb)).astype(f32)
    Ah, Bh = (A * ia[None, :]).astype(f32), (B * ib[None, :]).astype(f32)
    H = S // 2
    tot = f32(0)
    rows = []
    for r in range(R):
        c0, c1 = f32(0), f32(0)
        for p in range(H):                                             # two chains per row, as the FFMA2 halves
            c0 = f32(c0 + f32(Ah[r, p] * Bh[r, p]))
            c1 = f32(c1 + f32(Ah[r, p + H] * Bh[r, p + H]))
        rows.append(f32(c0 + c1))
    while len(rows) > 1:                                               # tree over the rows (shuffle reduction)
        rows = [f32(rows[i] + rows[i + 1]) if i + 1 < len(rows) else rows[i] for i in range(0, len(rows), 2)]
    tot = rows[0]
    n = int(((na != 0) & (nb != 0)).sum())
    if n == 0:
        return np.nan
    return float(f32(1) - f32(tot / f32(n)))


@pytest.mark.parametrize("R,S", [(20, 60), (40, 120)])
def test_fp32_screening_model_stays_inside_the_selection_margin(R, S):
    from oracle import oracle as orc
    port = orc.Port(orc.Params(R=R, S=S))
    rng = np.random.default_rng(2024)
    worst = 0.0
    for trial in range(24):
        fill = (0.05, 0.3, 0.7, 1.0)[trial % 4]                        # from almost empty to dense descriptors
        a = np.where(rng.random(R * S) < fill, rng.uniform(0, 12, R * S), 0.0).astype(np.float32).astype(np.float64)
        b = a.copy() if trial % 3 == 0 else np.where(rng.random(R * S) < fill, rng.uniform(0, 12, R * S), 0.0).astype(np.float32).astype(np.float64)
        if trial % 3 == 1:                                             # a noisy revisit: small perturbation of a rolled copy
            b = np.roll(a.reshape(S, R), 7, axis=0).ravel() + np.where(rng.random(R * S) < 0.1, 0.5, 0.0)
        for shift in (0, 1, S // 2, S - 3):
            d64 = port.dist_direct_shifted(a, b, shift)
            d32 = _screen_model(a, b, R, S, shift)
            if np.isnan(d64):
                assert np.isnan(d32)
                continue
            worst = max(worst, abs(d32 - d64))
    assert worst < 0.5 * EXH_EPS, worst
