"""CPU restatement of the tracking bookkeeping stages (TEST INFRASTRUCTURE ONLY -- never imported by the product).

Parity unpinned: the reference has no tests or golden vectors for these functions and cannot be built here; the
restatement follows the cited lines statement by statement (sequential loops, float32 arithmetic where the reference
uses float).

  tracklets_build     <- Tracking::GetStaticTrack / GetDynamicTrackNew   src/Tracking.cc:2201-2307 / 2309-2421
  update_mask         <- Tracking::UpdateMask                            src/Tracking.cc:2997-3110
  dyn_obj_tracking    <- Tracking::DynObjTracking                        src/Tracking.cc:1366-1612
"""
import numpy as np


def tracklets_build(assoc_rows, label_rows=None):
    """assoc_rows[i][j]: index in row i-1's frame of feature j of frame i+1 (or -1).  Returns (tracklets, obj_ids) with
    tracklets[t] = [(frame, feat), ...] in creation order (Tracking.cc:2212-2276)."""
    tracklets, obj_ids = [], []
    pre = []
    for i, row in enumerate(assoc_rows):
        cur = [-1] * len(row)
        for j, a in enumerate(row):
            a = int(a)
            if a == -1:
                continue
            if i > 0 and pre[a] != -1:                     # :2249-2253 extend
                tracklets[pre[a]].append((i + 1, j))
                cur[j] = pre[a]
            else:                                          # :2225-2237 / :2256-2268 new tracklet of two entries
                tracklets.append([(i, a), (i + 1, j)])
                if label_rows is not None:
                    obj_ids.append(int(label_rows[i][j]))  # :2345 / :2374
                cur[j] = len(tracklets) - 1
        pre = cur
    return tracklets, obj_ids


def _majority(values):
    """std::map count, then std::sort by count descending (SortPairInt, Tracking.cc:39-43).  For <= 16 distinct labels
    std::sort is an insertion sort, i.e. ties keep ascending-label order."""
    dups = {}
    for k in values:
        dups[int(k)] = dups.get(int(k), 0) + 1
    best, cnt = 0, -1
    for k in sorted(dups):
        if dups[k] > cnt:
            best, cnt = k, dups[k]
    return best


def update_mask(mask_cur, mask_last, flow_last, sem_label_last, corres):
    """Tracking.cc:2997-3065.  mask_* (h, w) int32, flow_last (h, w, 2) f32, corres (n, 2) f32 = mvObjCorres.
    Returns (updated copy of mask_cur, list of recovered labels)."""
    m = mask_cur.copy()
    h, w = m.shape
    uni = sorted(set(int(x) for x in sem_label_last))
    warped = []
    cu = corres[:, 0].astype(np.float32).astype(np.int32)   # const int u = pt.x  (truncation)
    cv = corres[:, 1].astype(np.float32).astype(np.int32)
    for lab in uni:
        idx = [i for i in range(len(sem_label_last)) if sem_label_last[i] == lab]
        tmp = [int(m[cv[i], cu[i]]) for i in idx if 0 < cu[i] < w and 0 < cv[i] < h]
        if len(tmp) < 100:
            continue
        if _majority(tmp) != 0:
            continue
        js, ks = np.nonzero(mask_last == lab)                 # row-major scan order like the reference's double loop
        fx = flow_last[js, ks, 0].astype(np.int32)
        fy = flow_last[js, ks, 1].astype(np.int32)
        x, y = ks + fx, js + fy
        ok = (x < w) & (x > 0) & (y < h) & (y > 0)
        m[y[ok], x[ok]] = lab
        warped.append(lab)
    return m, warped


def dyn_obj_tracking(sem_label, obj_label, keys, depth, flow3d, sem_label_last, last_sem_position, last_obj_stat,
                     last_mod_label, rows, cols, shrink_row, shrink_col, sf_mg_thres, sf_ds_thres, th_depth_obj, f_id, max_id):
    """Tracking.cc:1366-1612.  Returns (obj_label', objects [list of index lists], mod_label, sem_position, max_id')."""
    f32 = np.float32
    obj_label = np.array(obj_label, dtype=np.int32).copy()
    uni = sorted(set(int(x) for x in sem_label))
    posi = [[] for _ in uni]
    pos_of = {l: j for j, l in enumerate(uni)}
    for i in range(len(sem_label)):
        if obj_label[i] == -1:                              # :1388
            continue
        posi[pos_of[int(sem_label[i])]].append(i)
    objs, sems = [], []
    for j, p in enumerate(posi):
        if not p:
            continue
        sz = f32(len(p))
        count = f32(0)
        for i in p:                                         # :1414-1420
            u, v = f32(keys[i][0]), f32(keys[i][1])
            if v < f32(shrink_row) or v > f32(rows - shrink_row) or u < f32(shrink_col) or u > f32(cols - shrink_col):
                count = f32(count + f32(1))
        if f32(count / sz) > f32(0.5):                      # :1421
            obj_label[p] = -1
            continue
        dsum, sfc = f32(0), f32(0)
        for i in p:                                         # :1445-1452
            dsum = f32(dsum + f32(depth[i]))
            fx, fz = f32(flow3d[i][0]), f32(flow3d[i][2])
            nrm = np.sqrt(f32(f32(fx * fx) + f32(fz * fz)), dtype=f32)
            if nrm < f32(sf_mg_thres):
                sfc = f32(sfc + f32(1))
        if f32(sfc / sz) > f32(sf_ds_thres):                # :1490
            obj_label[p] = 0
            continue
        if f32(dsum / sz) > f32(th_depth_obj) or len(p) < 150:   # :1497
            obj_label[p] = -1
            continue
        objs.append(p)
        sems.append(uni[j])
    if f_id == 1:                                           # :1548
        max_id = 1
    lab_id = []
    for p in objs:
        new_lab = _majority([sem_label_last[k] for k in p])  # :1555-1563
        ident = -1
        if max_id == 1:                                     # :1565
            ident = max_id
            max_id += 1
        else:
            for k in range(len(last_sem_position)):         # :1575-1585
                if last_sem_position[k] == new_lab and last_obj_stat[k]:
                    ident = int(last_mod_label[k])
                    break
            if ident == -1:
                ident = max_id
                max_id += 1
        obj_label[p] = ident
        lab_id.append(ident)
    return obj_label, objs, lab_id, sems, max_id


def motion_model_inliers(obj, img, T_mm, K4, thr=0.4):
    """Tracking.cc:1676-1691 / 1783-1798: reprojection test of the constant-motion model in float arithmetic
    (cv::Mat float gemm = double accumulation rounded once; 1.0/z in double rounded to float)."""
    f32 = np.float32
    obj = np.asarray(obj, f32); img = np.asarray(img, f32); T = np.asarray(T_mm, f32).reshape(4, 4); K = np.asarray(K4, f32)
    Xc = (obj.astype(np.float64) @ T[:3, :3].astype(np.float64).T + T[:3, 3].astype(np.float64)).astype(f32)
    invz = (1.0 / Xc[:, 2].astype(np.float64)).astype(f32)
    u = (K[0] * Xc[:, 0]) * invz + K[2]
    v = (K[1] * Xc[:, 1]) * invz + K[3]
    u_, v_ = img[:, 0] - u, img[:, 1] - v
    rpe = np.sqrt(u_ * u_ + v_ * v_, dtype=f32)
    return np.nonzero(rpe.astype(np.float64) < thr)[0].astype(np.int32)


def init_model(obj, img, K4, T_mm=None, iters=500, thr=0.4, conf=0.98):
    """GetInitModelCam / GetInitModelObj (Tracking.cc:1614-1715, 1717-1849): RANSAC model vs constant-motion model.
    Returns (T_init 4x4 f32, chosen local inlier indices, info dict)."""
    from . import pyoracle as po
    r = po.pnp_ransac(obj, img, np.asarray(K4, np.float32).astype(np.float64), iters, thr, conf)
    Mod = np.eye(4, dtype=np.float32)
    inl = np.zeros(0, np.int32)
    if r is not None:
        Mod[:3, :3] = r["Rt"][:9].reshape(3, 3).astype(np.float32)
        Mod[:3, 3] = r["Rt"][9:].astype(np.float32)
        inl = r["inliers"]
    info = dict(n_ransac=len(inl), n_mm=0, used_mm=False, ransac=r)
    if T_mm is not None:
        mm = motion_model_inliers(obj, img, T_mm, K4, thr)
        info["n_mm"] = len(mm)
        if not (len(inl) > len(mm)):
            info["used_mm"] = True
            return np.asarray(T_mm, np.float32).reshape(4, 4).copy(), mm, info
    return Mod, inl, info
