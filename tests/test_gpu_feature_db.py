"""Feature-DB builder vs the CPU oracle: batched encode of a window set into the matcher's
layout (fp32 rows, bf16 rows + norms), sharded by rows; collect_CVAE_feature_action.py:167-190,
compute_cnt_norm.py:157-179, test_fullframework.py:266-272,:293."""
import numpy as np
import pytest
import torch

from mocha_oracle import clip, matching
from mocha_sigasia2023_b200 import feature_db, weights, workload
from mocha_sigasia2023_b200.balltree import BallTree
from mocha_sigasia2023_b200.sharded import shard_bounds

pytestmark = pytest.mark.gpu

N = 70    # ragged against the batch of 32


@pytest.fixture(scope="module")
def setup():
    gen_sd = weights.generator_state_dict(1777)
    st = workload.driver_stats()
    X = workload.pose_windows(N, 4242)
    enc, cnt = clip.encode_windows({k: v.numpy() for k, v in gen_sd.items()}, X)
    rows = ((cnt - st.cnt_mean[None]) / st.cnt_std[None]).reshape(N, -1).astype(np.float32)
    return gen_sd, st, X, enc, cnt, rows


def test_fp32_db_vs_oracle(setup):
    gen_sd, st, X, enc, cnt, rows = setup
    fdb = feature_db.build_feature_db(gen_sd, weights.DEFAULT_MODEL_CFG, st.cnt_mean, st.cnt_std, X, batch=32,
                                      keep_cnt=True)
    assert (fdb.lo, fdb.hi, fdb.n_total) == (0, N, N)
    for got, want in ((fdb.encoded, enc), (fdb.cnt, cnt), (fdb.rows32, rows)):
        got = got.cpu().numpy()
        assert np.abs(got - want).max() / np.abs(want).max() < 1e-4
    # the bf16 rows are the fp32 rows minus the mean row, rounded to nearest-even; the norms those of the ROUNDED rows
    assert torch.allclose(fdb.center, fdb.rows32.mean(0), rtol=1e-5, atol=1e-5)
    r16 = (fdb.rows32 - fdb.center).to(torch.bfloat16)
    assert torch.equal(fdb.rows16, r16)
    np.testing.assert_allclose(fdb.norms.cpu().numpy(), (r16.double() ** 2).sum(1).cpu().numpy(), rtol=1e-5)
    mean, std = fdb.cnt_statistics()
    np.testing.assert_allclose(mean.cpu().numpy(), cnt.mean(0), rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(std.cpu().numpy(), cnt.std(0), rtol=1e-3, atol=1e-4)


def test_sharded_db_equals_slices_and_matches(setup):
    gen_sd, st, X, enc, cnt, rows = setup
    whole = feature_db.build_feature_db(gen_sd, weights.DEFAULT_MODEL_CFG, st.cnt_mean, st.cnt_std, X, batch=32)
    q = rows[::7] + 0.01
    wd, wi = matching.knn(rows, q, 2)
    ds, ids = [], []
    for r in range(3):
        part = feature_db.build_feature_db(gen_sd, weights.DEFAULT_MODEL_CFG, st.cnt_mean, st.cnt_std, X, batch=32,
                                           shard=(r, 3), keep_encoded=False)
        lo, hi = shard_bounds(N, 3, r)
        assert (part.lo, part.hi) == (lo, hi) and part.encoded is None
        # batch composition differs between the whole build and a shard: GEMM tiles see other neighbours, fp32 sums
        # are order-independent per row, so the rows agree to fp32 round-off
        np.testing.assert_allclose(part.rows32.cpu().numpy(), whole.rows32[lo:hi].cpu().numpy(), rtol=1e-5, atol=1e-5)
        d, i = BallTree.from_feature_db(part).query_device(torch.from_numpy(q).cuda(), k=2)
        ds.append(d); ids.append(i + lo)
    from mocha_sigasia2023_b200.sharded import merge_topk_cuda
    d, i = merge_topk_cuda(torch.stack(ds), torch.stack(ids), 2)
    got_d, got_i = d.cpu().numpy(), i.cpu().numpy()
    ok = (wd[:, 1] - wd[:, 0]) > 1e-3      # GPU rows differ from the oracle's by ~1e-4 relative
    np.testing.assert_array_equal(got_i[ok, 0], wi[ok, 0])


def test_bf16_encoder_db(setup):
    gen_sd, st, X, enc, cnt, rows = setup
    fdb = feature_db.build_feature_db(gen_sd, weights.DEFAULT_MODEL_CFG, st.cnt_mean, st.cnt_std, X, batch=64,
                                      precision="bf16")
    err = np.abs(fdb.encoded.cpu().numpy() - enc).max() / np.abs(enc).max()
    assert err < 2e-2, err
