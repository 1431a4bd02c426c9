"""Tracking bookkeeping: oracle sanity on CPU, exact parity of the C ABI on the GPU.
vdo_tracklets_build is host-only, so its parity test runs without a GPU as well."""
import numpy as np
import pytest

from oracle import tracking_ops as T
from vdo_slam_b200 import capi


def _assoc_rows(rng, n_rows, n_feat, p_match=0.8, dup=True):
    rows, labs, prev_n = [], [], n_feat
    for i in range(n_rows):
        n = int(rng.integers(n_feat // 2, n_feat + 1))
        a = np.full(n, -1, np.int32)
        m = rng.random(n) < p_match
        a[m] = rng.integers(0, prev_n, int(m.sum()))           # duplicates on purpose: two features may claim the same parent
        if not dup:
            perm = rng.permutation(prev_n)[:n]
            a = np.where(m[:len(perm)], perm, -1).astype(np.int32) if len(perm) == n else a
        rows.append(a); labs.append(rng.integers(1, 6, n).astype(np.int32)); prev_n = n
    return rows, labs


def test_oracle_tracklets_hand_case():
    # frames 0..3; feature ids chosen by hand
    rows = [np.array([0, -1, 2]), np.array([2, 0, -1, 1]), np.array([-1, 3, 0])]
    trk, _ = T.tracklets_build(rows)
    # row0: new (0,0)-(1,0) id0; new (0,2)-(1,2) id1.  row1: j0 parent 2 -> id1 extend; j1 parent 0 -> id0 extend; j3 parent 1 (untracked) -> new id2
    # row2: j1 parent 3 -> id2 extend; j2 parent 0 -> id1 extend
    assert trk == [[(0, 0), (1, 0), (2, 1)], [(0, 2), (1, 2), (2, 0), (3, 2)], [(1, 1), (2, 3), (3, 1)]]


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_tracklets_capi_matches_oracle(seed):
    rng = np.random.default_rng(seed)
    rows, labs = _assoc_rows(rng, 12, 300)
    t_ref, o_ref = T.tracklets_build(rows, labs)
    t, o = capi.tracklets_build(rows, labs)
    assert t == t_ref and o == o_ref
    t2, o2 = capi.tracklets_build(rows)
    assert t2 == t_ref and o2 == []


def test_tracklets_empty_and_ragged():
    assert capi.tracklets_build([]) == ([], [])
    rows = [np.zeros(0, np.int32), np.array([-1, -1], np.int32)]
    assert capi.tracklets_build(rows) == ([], [])


def _mask_case(rng, w=320, h=128, n_obj=4):
    mask_last = np.zeros((h, w), np.int32)
    for o in range(1, n_obj + 1):
        x0, y0 = int(rng.integers(10, w - 90)), int(rng.integers(10, h - 50))
        mask_last[y0:y0 + 40, x0:x0 + 70] = o
    flow_last = (rng.normal(0, 1.0, (h, w, 2)) + np.array([6.5, -2.5])).astype(np.float32)
    # current mask: objects moved by the flow; some objects are missing (label 0 -> to be recovered)
    mask_cur = np.zeros_like(mask_last)
    for o in range(1, n_obj + 1):
        if o % 2 == 0:
            continue
        ys, xs = np.nonzero(mask_last == o)
        x2, y2 = np.clip(xs + 6, 0, w - 1), np.clip(ys - 2, 0, h - 1)
        mask_cur[y2, x2] = o
    ys, xs = np.nonzero(mask_last > 0)
    sel = rng.permutation(len(ys))[:3000]
    ys, xs = ys[sel], xs[sel]
    sem = mask_last[ys, xs].astype(np.int32)
    cor = np.stack([xs + flow_last[ys, xs, 0], ys + flow_last[ys, xs, 1]], 1).astype(np.float32)
    cor[::97] = [-3.0, 5.0]                                    # a few out-of-image correspondences
    return mask_cur, mask_last, flow_last, sem, cor


def test_oracle_update_mask_recovers_missing():
    rng = np.random.default_rng(3)
    mc, ml, fl, sem, cor = _mask_case(rng)
    m, warped = T.update_mask(mc, ml, fl, sem, cor)
    assert warped == [2, 4]
    assert (m == 2).sum() > 1000 and (m == 4).sum() > 1000 and (mc == 2).sum() == 0


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [3, 4])
def test_update_mask_gpu(seed):
    rng = np.random.default_rng(seed)
    mc, ml, fl, sem, cor = _mask_case(rng, w=1242 if seed == 4 else 320, h=375 if seed == 4 else 128, n_obj=6)
    m_ref, w_ref = T.update_mask(mc, ml, fl, sem, cor)
    ctx = capi.Context()
    h, w = mc.shape
    cur, last = capi.Frame(ctx, w, h), capi.Frame(ctx, w, h)
    cur.upload(mask=mc); last.upload(mask=ml, flow=fl)
    m, wl = capi.update_mask(cur, last, sem, cor)
    assert wl == w_ref
    assert np.array_equal(m, m_ref)
    # second application on the updated mask is idempotent (nothing left to recover)
    m2, wl2 = capi.update_mask(cur, last, sem, cor)
    m2_ref, w2_ref = T.update_mask(m_ref, ml, fl, sem, cor)
    assert wl2 == w2_ref and np.array_equal(m2, m2_ref)


def _dyn_case(rng, n_obj=7, rows=375, cols=1242):
    sem, keys, depth, f3, seml, lab = [], [], [], [], [], []
    for o in range(1, n_obj + 1):
        n = int(rng.integers(100, 900))
        kind = o % 5                  # 0: boundary, 1: static-like flow, 2: far, 3/4: kept
        cx = 30 if kind == 0 else rng.uniform(200, cols - 200)
        cy = rng.uniform(80, rows - 80)
        k = np.stack([rng.normal(cx, 25, n), rng.normal(cy, 15, n)], 1)
        d = rng.normal(60 if kind == 2 else 15, 2, n)
        mag = 0.05 if kind == 1 else 0.6
        f = rng.normal(0, 1, (n, 3)) * np.array([mag, 5.0, mag])
        sem += [o] * n; keys.append(k); depth.append(d); f3.append(f)
        seml += list(np.where(rng.random(n) < 0.8, o, rng.integers(0, n_obj + 1, n)))
        lab += list(np.where(rng.random(n) < 0.05, -1, o))
    perm = rng.permutation(len(sem))
    A = lambda x, dt: np.asarray(x, dt)[perm]
    return (A(sem, np.int32), A(lab, np.int32), np.concatenate(keys).astype(np.float32)[perm], np.concatenate(depth).astype(np.float32)[perm],
            np.concatenate(f3).astype(np.float32)[perm], A(seml, np.int32))


def test_oracle_dyn_obj_tracking_ids():
    rng = np.random.default_rng(5)
    sem, lab, keys, depth, f3, seml = _dyn_case(rng)
    ol, objs, ml, sp, mid = T.dyn_obj_tracking(sem, lab, keys, depth, f3, seml, [3, 4], [1, 1], [7, 9], 375, 1242, 25, 50, 0.12, 0.7, 40.0, 5, 12)
    assert set(sp) <= {3, 4, 6, 7}
    for s, m in zip(sp, ml):
        if s == 3: assert m == 7
        if s == 4: assert m == 9
    assert mid == 12 + sum(1 for s in sp if s not in (3, 4))
    # first frame: ids restart from 1
    ol1, objs1, ml1, sp1, mid1 = T.dyn_obj_tracking(sem, lab, keys, depth, f3, seml, [], [], [], 375, 1242, 25, 50, 0.12, 0.7, 40.0, 1, 99)
    assert ml1 == list(range(1, len(ml1) + 1)) and mid1 == len(ml1) + 1


@pytest.mark.gpu
@pytest.mark.parametrize("seed,f_id", [(5, 5), (6, 1), (7, 9)])
def test_dyn_obj_tracking_gpu(seed, f_id):
    rng = np.random.default_rng(seed)
    sem, lab, keys, depth, f3, seml = _dyn_case(rng, n_obj=12)
    last = ([3, 4, 8, 9], [1, 0, 1, 1], [7, 9, 2, 5])
    ref = T.dyn_obj_tracking(sem, lab, keys, depth, f3, seml, *last, 375, 1242, 25, 50, 0.12, 0.7, 40.0, f_id, 12)
    ctx = capi.Context()
    got = capi.dyn_obj_tracking(ctx, sem, lab, keys, depth, f3, seml, *last, 375, 1242, 25, 50, 0.12, 0.7, 40.0, f_id, 12)
    assert np.array_equal(got[0], ref[0])
    assert got[1] == ref[1] and got[2] == ref[2] and got[3] == ref[3] and got[4] == ref[4]


def _renew_case(rng, w=640, h=240, n_obj=3, max_sta=300, max_obj=200):
    mask = np.zeros((h, w), np.int32)
    boxes = []
    for o in range(1, n_obj + 2):                 # one more semantic region than tracked objects: a brand-new object
        x0, y0 = 40 + (o - 1) * 140, int(rng.integers(40, h - 110))
        mask[y0:y0 + 70, x0:x0 + 110] = o; boxes.append((x0, y0))
    depth = rng.uniform(3, 60, (h, w)).astype(np.float32)
    depth[mask > 0] = rng.uniform(5, 30, int((mask > 0).sum())).astype(np.float32)
    depth[rng.random((h, w)) < 0.02] = 0
    flow = rng.normal(0, 2, (h, w, 2)).astype(np.float32)
    flow[rng.random((h, w)) < 0.03] = 0
    n_stat = 400
    stat_keys = np.stack([rng.uniform(-5, w + 5, n_stat), rng.uniform(-5, h + 5, n_stat)], 1).astype(np.float32)
    tm = rng.permutation(n_stat)[:330].astype(np.int32); tm[rng.random(330) < 0.2] = -1
    samp = np.stack([rng.uniform(0, w, 900), rng.uniform(0, h, 900)], 1).astype(np.float32)
    samp[:40] = stat_keys[tm[tm >= 0][:40]] + 0.3                      # near-duplicates of inliers
    # object side: previous object keys (float, propagated), inlier lists, fresh raster samples
    obj_keys, obj_label, inl = [], [], []
    for o in range(1, n_obj + 1):
        x0, y0 = boxes[o - 1]; n = 260
        k = np.stack([rng.uniform(x0 - 8, x0 + 118, n), rng.uniform(y0 - 8, y0 + 78, n)], 1)
        base = sum(len(x) for x in obj_keys)
        obj_keys.append(k); obj_label += [10 + o] * n
        inl.append((base + rng.permutation(n)[:int(rng.integers(60, 230))]).astype(np.int32))
    obj_keys = np.concatenate(obj_keys).astype(np.float32)
    ys, xs = np.mgrid[0:h:4, 0:w:4]
    ys, xs = ys.ravel(), xs.ravel()
    sel = mask[ys, xs] > 0
    ys, xs = ys[sel], xs[sel]
    tmp_keys = np.stack([xs, ys], 1).astype(np.float32); tmp_depth = depth[ys, xs]; tmp_sem = mask[ys, xs]
    tmp_flow = flow[ys, xs]; tmp_corres = tmp_keys + tmp_flow
    obj_stat = np.ones(n_obj, np.uint8); obj_stat[n_obj - 1] = 0          # one failed object: re-enters as new with label -2
    sem_pos = np.arange(1, n_obj + 1, dtype=np.int32); mod_lab = (10 + sem_pos).astype(np.int32)
    K4 = np.array([721.5377, 721.5377, 320.0, 120.0], np.float32)
    Twc = np.eye(4, dtype=np.float32); Twc[:3, 3] = [0.3, -0.1, 2.0]; Twc[0, 1] = 0.01; Twc[1, 0] = -0.01
    return dict(mask=mask, depth=depth, flow=flow, tm_sta=tm, stat_keys=stat_keys, samp_keys=samp, max_num_sta=max_sta, obj_inliers=inl, obj_stat=obj_stat,
                sem_position=sem_pos, mod_label=mod_lab, obj_keys=obj_keys, obj_label=np.array(obj_label, np.int32), tmp_keys=tmp_keys, tmp_depth=tmp_depth,
                tmp_sem=tmp_sem, tmp_flow=tmp_flow, tmp_corres=tmp_corres, max_num_obj=max_obj, K4=K4, Twc=Twc)


def test_oracle_renew_frame_info_quotas():
    c = _renew_case(np.random.default_rng(11))
    S, O = T.renew_frame_info(**c)
    assert len(S["keys"]) == c["max_num_sta"]                           # enough ORB candidates to fill the quota exactly
    n_inl = int((S["inlier_id"] >= 0).sum())
    assert 0 < n_inl < c["max_num_sta"] and (S["inlier_id"][n_inl:] == -1).all()
    assert (O["label"] == -2).sum() > 0                                 # failed + brand-new object re-enter with label -2
    for lab in (11, 12):
        assert (O["label"] == lab).sum() <= c["max_num_obj"] + 230


@pytest.mark.gpu
@pytest.mark.parametrize("seed,max_sta,max_obj", [(11, 300, 200), (12, 100, 50), (13, 2000, 5000)])
def test_renew_frame_info_gpu(seed, max_sta, max_obj):
    c = _renew_case(np.random.default_rng(seed), max_sta=max_sta, max_obj=max_obj)
    S_ref, O_ref = T.renew_frame_info(**c)
    ctx = capi.Context()
    h, w = c["mask"].shape
    fr = capi.Frame(ctx, w, h)
    fr.upload(depth=c["depth"], flow=c["flow"], mask=c["mask"])
    args = {k: v for k, v in c.items() if k not in ("mask", "depth", "flow")}
    S, O = capi.renew_frame_info(fr, **args)
    for k in S_ref:
        assert np.array_equal(S[k], S_ref[k]), k
    for k in O_ref:
        assert np.array_equal(O[k], O_ref[k]), k
