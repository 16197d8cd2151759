"""Property tests (tier 3): shard-split invariance, permutation equivariance,
random shapes.  The CPU half exercises the oracle and the host logic; the GPU half the kernels."""
import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st, HealthCheck

from oracle import hamming_oracle as ho
from oracle import c_oracle as co

SETTINGS = dict(deadline=None, suppress_health_check=list(HealthCheck))


def _desc(rng, n, hi):
    return rng.integers(0, hi, (n, 32), dtype=np.uint8)


@settings(max_examples=40, **SETTINGS)
@given(seed=st.integers(0, 2**31 - 1), nq=st.integers(1, 40), nt=st.integers(1, 200), hi=st.sampled_from([2, 4, 256]),
       ncuts=st.integers(0, 5))
def test_oracle_shard_split_invariance(seed, nq, nt, hi, ncuts):
    """Top-2 over a union of shards == top-2 of the per-shard top-2's, for any split (ties included)."""
    rng = np.random.default_rng(seed)
    q, t = _desc(rng, nq, hi), _desc(rng, nt, hi)
    cuts = sorted(set([0, nt] + rng.integers(0, nt + 1, ncuts).tolist()))
    parts = [co.knn2_keys(q, t[a:b], train_base=a) if b > a else np.full((nq, 2), ho.NO_MATCH_KEY, np.uint64)
             for a, b in zip(cuts[:-1], cuts[1:])]
    assert np.array_equal(ho.merge_top2_keys(np.stack(parts)), ho.knn2_keys(q, t))


@settings(max_examples=40, **SETTINGS)
@given(seed=st.integers(0, 2**31 - 1), nq=st.integers(1, 30), nt=st.integers(2, 150))
def test_oracle_permutation_equivariance(seed, nq, nt):
    """Permuting train rows permutes trainIdx wherever the top-2 distances are tie-free."""
    rng = np.random.default_rng(seed)
    q, t = _desc(rng, nq, 256), _desc(rng, nt, 256)
    perm = rng.permutation(nt)
    idx, dist = ho.knn(q, t, 2)
    pidx, pdist = ho.knn(q, t[perm], 2)
    assert np.array_equal(dist, pdist)
    d = ho.hamming_matrix(q, t)
    srt = np.sort(d, axis=1)
    free = (srt[:, 0] != srt[:, 1]) & ((srt[:, 1] != srt[:, 2]) if nt > 2 else True)
    assert np.array_equal(perm[pidx[free]], idx[free])


@pytest.mark.gpu
@settings(max_examples=30, **SETTINGS)
@given(seed=st.integers(0, 2**31 - 1), nq=st.integers(1, 700), nt=st.integers(1, 1500), hi=st.sampled_from([2, 3, 256]),
       variant=st.sampled_from(["popc", "i8", "f4"]), base=st.sampled_from([0, 1, 123456789]))
def test_gpu_random_shapes_vs_oracle(seed, nq, nt, hi, variant, base):
    from slam_experiments_b200 import _native as nat
    rng = np.random.default_rng(seed)
    q, t = _desc(rng, nq, hi), _desc(rng, nt, hi)
    got = nat.knn2_keys(torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda(), train_base=base, variant=variant)
    assert np.array_equal(got.cpu().numpy().view(np.uint64), co.knn2_keys(q, t, train_base=base))


@pytest.mark.gpu
@settings(max_examples=12, **SETTINGS)
@given(seed=st.integers(0, 2**31 - 1), nq=st.integers(1, 400), nt=st.integers(1, 900), batch=st.integers(1, 5),
       variant=st.sampled_from(["popc", "i8", "f4"]), ratio=st.sampled_from([None, 0.7, 0.9]), cross=st.booleans())
def test_gpu_fused_pipeline_random(seed, nq, nt, batch, variant, ratio, cross):
    from slam_experiments_b200 import _native as nat
    rng = np.random.default_rng(seed)
    t = rng.integers(0, 256, (batch, nt, 32), dtype=np.uint8)
    q = rng.integers(0, 256, (batch, nq, 32), dtype=np.uint8)
    k = min(nq, nt)
    q[:, : k // 2] = t[:, : k // 2] ^ (rng.random((batch, k // 2, 32)) < 0.05).astype(np.uint8)   # matchable half
    oq, ot, od, cnt = nat.match_fused(torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda(), ratio=ratio,
                                      cross_check=cross, variant=variant)
    oq, ot, od, cnt = oq.cpu().numpy(), ot.cpu().numpy(), od.cpu().numpy(), cnt.cpu().numpy()
    for b in range(batch):
        eq, et, ed = co.pipeline(q[b], t[b], ratio=ratio, cross_check=cross)
        n = int(cnt[b])
        assert n == len(eq)
        assert np.array_equal(oq[b, :n], eq) and np.array_equal(ot[b, :n], et) and np.array_equal(od[b, :n], ed)
