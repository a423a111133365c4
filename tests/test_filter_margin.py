"""Host-side check of the inclusive margin of the tensor-core pair filter (csrc/acsf2.cu, `mma_tf32_16x8x8`).

The kernel keeps a neighbour pair (l, s) of a centre when
    -l~.s~ + S~_s + trunc(S~_l - thr/2) < 0,        S~ = tf32(|v~|^2 / 2),  v~ = tf32(v),  thr = rc^2 (1 + 1e-4) + 1e-4 + margin
with margin = 5e-3 r_list^2 + 1e-3, evaluated with exact products and FP32 accumulation.  This file re-states that
arithmetic in numpy (round-to-nearest TF32 for the staged vectors, truncation to TF32 for the row constant, float32
sums in several orders) and asserts on random and adversarial neighbourhoods that no pair with an exact r_jk < rc is
ever dropped, and that the margin admits well under 2 % extra pairs.  The same for the Gaussian-screening test."""
import numpy as np
import pytest

RC = 12.0


def tf32_rna(x):
    """cvt.rna.tf32.f32: round to nearest (ties away) at 10 explicit mantissa bits."""
    b = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    b = (b + 0x1000) & 0xFFFFE000
    return b.astype(np.uint32).view(np.float32)


def tf32_trunc(x):
    """what the tensor core does with an FP32 operand declared .tf32: the low 13 mantissa bits are ignored."""
    return (np.asarray(x, np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def filter_keep(vec, thr, order):
    """survivor matrix of the cutoff test for all (l, s) pairs of one neighbourhood; vec in float64"""
    v = tf32_rna(vec.astype(np.float32))                                   # staged (x, y, z)
    s_half = tf32_rna(np.float32(0.5) * (v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1] + v[:, 2] * v[:, 2]))
    l1 = tf32_trunc(s_half - np.float32(0.5) * np.float32(thr))           # row constant, truncated by the MMA
    # k = 0..4 products: exact in FP32 (two 11-bit significands), then accumulated in FP32 in the given order
    terms = np.stack([-(v[:, None, 0] * v[None, :, 0]), -(v[:, None, 1] * v[None, :, 1]), -(v[:, None, 2] * v[None, :, 2]),
                      np.broadcast_to(s_half[None, :], (len(v), len(v))), np.broadcast_to(l1[:, None], (len(v), len(v)))]).astype(np.float32)
    acc = np.zeros((len(v), len(v)), np.float32)
    for k in order:
        acc = (acc + terms[k]).astype(np.float32)
    return acc < 0


def margin(r_list):
    return 5.0e-3 * r_list * r_list + 1e-3


def neighbourhood(rng, n, kind):
    if kind == "ball":      # uniform in the cutoff sphere
        v = rng.standard_normal((n, 3))
        v *= (RC * rng.random(n) ** (1 / 3) / np.linalg.norm(v, axis=1))[:, None]
    elif kind == "shell":   # everything close to the cutoff radius: largest magnitudes, largest rounding
        v = rng.standard_normal((n, 3))
        v *= (RC * (1 - 1e-3 * rng.random(n)) / np.linalg.norm(v, axis=1))[:, None]
    else:                   # pairs placed at r_jk = rc (1 - eps) exactly on purpose
        base = neighbourhood(rng, n // 2, "ball") * 0.45
        d = rng.standard_normal(base.shape)
        d *= (RC * (1 - 10.0 ** rng.uniform(-12, -3, len(base))) / np.linalg.norm(d, axis=1))[:, None]
        v = np.concatenate([base, base + d])
        v = v[np.linalg.norm(v, axis=1) < RC]
    return v


@pytest.mark.parametrize("kind", ["ball", "shell", "boundary"])
def test_no_live_pair_is_dropped_and_few_dead_ones_are_kept(kind):
    rng = np.random.default_rng({"ball": 1, "shell": 2, "boundary": 3}[kind])
    thr = RC * RC * 1.0001 + 1e-4 + margin(RC)
    live_total = kept_total = 0
    for _ in range(60):
        v = neighbourhood(rng, 120, kind)
        d2 = ((v[:, None, :] - v[None, :, :]) ** 2).sum(-1)
        live = d2 < RC * RC
        for order in ((0, 1, 2, 3, 4), (4, 3, 2, 1, 0), (3, 4, 0, 1, 2)):
            keep = filter_keep(v, thr, order)
            assert not (live & ~keep).any(), "a pair inside the cutoff was dropped"
        off = ~np.eye(len(v), dtype=bool)
        live_total += int((live & off).sum())
        kept_total += int((keep & off).sum())
    if kind == "ball":  # the realistic case: the margin costs little
        assert kept_total <= 1.02 * live_total


def test_gaussian_screening_test_is_inclusive_too():
    """second accumulator: (r_jk^2 + r_s^2 + r_l^2 - r2max) / 2 with |s~|^2 = 2 S~_s in the column and the truncated row
    constant 2 S~_l - r2max / 2; r2max carries 2 x margin.  No pair with an exact sum below the un-inflated bound may go."""
    rng = np.random.default_rng(7)
    for _ in range(40):
        v = neighbourhood(rng, 120, "ball")
        r2 = (v ** 2).sum(1)
        d2 = ((v[:, None, :] - v[None, :, :]) ** 2).sum(-1)
        bound = rng.uniform(150.0, 420.0)                      # ref + T / eta of the kernel, before its margins
        r2max = np.float32(bound * 1.001 + 2 * margin(RC))
        vt = tf32_rna(v.astype(np.float32))
        s_half = tf32_rna(np.float32(0.5) * (vt[:, 0] ** 2 + vt[:, 1] ** 2 + vt[:, 2] ** 2))
        l2 = tf32_trunc(np.float32(2.0) * s_half - np.float32(0.5) * r2max)
        dot = (vt[:, None, 0] * vt[None, :, 0]).astype(np.float32)
        dot = (dot + vt[:, None, 1] * vt[None, :, 1]).astype(np.float32)
        dot = (dot + vt[:, None, 2] * vt[None, :, 2]).astype(np.float32)
        acc = ((-dot + (np.float32(2.0) * s_half)[None, :]).astype(np.float32) + l2[:, None]).astype(np.float32)
        keep = acc < 0
        must = (d2 + r2[:, None] + r2[None, :]) < bound
        assert not (must & ~keep).any()
