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


def get3d_world(key, depth, K4, Twc):
    """Optimizer::Get3DinWorld (src/Optimizer.cc:2974-2993): float back-projection, Rwc*x + twc as a float gemm."""
    f32 = np.float32
    K = np.asarray(K4, f32); T = np.asarray(Twc, f32).reshape(4, 4)
    invfx, invfy = f32(1.0) / K[0], f32(1.0) / K[1]
    z = f32(depth)
    x = f32(f32(f32(key[0]) - K[2]) * z) * invfx
    y = f32(f32(f32(key[1]) - K[3]) * z) * invfy
    v = np.array([x, y, z], np.float64)
    return (T[:3, :3].astype(np.float64) @ v + T[:3, 3].astype(np.float64)).astype(f32)


def renew_frame_info(mask, depth, flow, tm_sta, stat_keys, samp_keys, max_num_sta, obj_inliers, obj_stat, sem_position, mod_label,
                     obj_keys, obj_label, tmp_keys, tmp_depth, tmp_sem, tmp_flow, tmp_corres, max_num_obj, K4, Twc):
    """Tracking::RenewFrameInfo (src/Tracking.cc:2660-2995), statement by statement.  Images: mask (h,w) i32, depth (h,w) f32, flow (h,w,2) f32."""
    f32 = np.float32
    h, w = mask.shape
    stat_keys = np.asarray(stat_keys, f32).reshape(-1, 2); samp_keys = np.asarray(samp_keys, f32).reshape(-1, 2)
    obj_keys = np.asarray(obj_keys, f32).reshape(-1, 2); tmp_keys = np.asarray(tmp_keys, f32).reshape(-1, 2)
    tmp_flow = np.asarray(tmp_flow, f32).reshape(-1, 2); tmp_corres = np.asarray(tmp_corres, f32).reshape(-1, 2)
    S = dict(keys=[], corres=[], flow=[], inlier_id=[], depth=[])

    def try_static(k, ident):
        x, y = int(k[0]), int(k[1])                                      # :2684-2685 truncation
        if x >= w or y >= h or x <= 0 or y <= 0:
            return False
        if mask[y, x] != 0:
            return False
        d = depth[y, x]
        if d > 40 or d <= 0:
            return False
        fx, fy = flow[y, x, 0], flow[y, x, 1]
        if fx != 0 and fy != 0:
            cx, cy = f32(k[0] + fx), f32(k[1] + fy)
            if cx < w and cy < h and cx > 0 and cy > 0:
                S["keys"].append((k[0], k[1])); S["corres"].append((cx, cy)); S["flow"].append((fx, fy)); S["inlier_id"].append(ident); S["depth"].append(d)
                return True
        return False

    for t in tm_sta:                                                      # (1) :2677-2708
        if t == -1:
            continue
        try_static(stat_keys[t], int(t))
        if len(S["keys"]) > max_num_sta:
            break
    check = np.array(S["keys"], f32).reshape(-1, 2)                       # snapshot :2714
    tot, start_id, step = len(S["keys"]), 0, 20
    while tot < max_num_sta:                                              # (2) :2719-2790
        if start_id == step:
            break
        for i in range(start_id, len(samp_keys), step):
            k = samp_keys[i]
            if len(check):
                dx, dy = check[:, 0] - k[0], check[:, 1] - k[1]
                if (np.sqrt(dx * dx + dy * dy, dtype=f32) < f32(1.0)).any():
                    continue
            if try_static(k, -1):
                tot += 1
            if tot >= max_num_sta:
                break
        start_id += 1
    S["p3d"] = [get3d_world(k, d, K4, Twc) for k, d in zip(S["keys"], S["depth"])]

    O = dict(keys=[], depth=[], corres=[], flow=[], sem=[], inlier_id=[], label=[])
    n_obj = len(obj_inliers)
    fea = [0] * n_obj
    for i in range(n_obj):                                                # (1) :2831-2869
        if not obj_stat[i]:
            fea[i] = -1
            continue
        cnt = 0
        for idx in obj_inliers[i]:
            x, y = int(obj_keys[idx][0]), int(obj_keys[idx][1])
            if x >= w or y >= h or x <= 0 or y <= 0:
                continue
            if mask[y, x] != 0 and depth[y, x] < 25 and depth[y, x] > 0:
                fx, fy = flow[y, x, 0], flow[y, x, 1]
                cx, cy = f32(f32(x) + fx), f32(f32(y) + fy)
                if cx < w and cy < h and cx > 0 and cy > 0:
                    O["keys"].append((f32(x), f32(y))); O["depth"].append(depth[y, x]); O["sem"].append(int(mask[y, x])); O["flow"].append((fx, fy))
                    O["corres"].append((cx, cy)); O["inlier_id"].append(int(idx)); O["label"].append(int(obj_label[idx]))
                    cnt += 1
        fea[i] = cnt
    check = np.array(O["keys"], f32).reshape(-1, 2)                       # snapshot :2874

    def push_tmp(j, lab):
        O["keys"].append(tuple(tmp_keys[j])); O["depth"].append(tmp_depth[j]); O["sem"].append(int(tmp_sem[j])); O["flow"].append(tuple(tmp_flow[j]))
        O["corres"].append(tuple(tmp_corres[j])); O["inlier_id"].append(-1); O["label"].append(lab)

    for i in range(n_obj):                                                # (2) :2875-2927
        if not obj_stat[i]:
            continue
        sem, tot, start_id, step = sem_position[i], fea[i], 0, 15
        while tot < max_num_obj:
            if start_id == step:
                break
            for j in range(start_id, len(tmp_sem), step):
                if tmp_sem[j] != sem:
                    continue
                if len(check):
                    dx, dy = check[:, 0] - tmp_keys[j][0], check[:, 1] - tmp_keys[j][1]
                    if (np.sqrt(dx * dx + dy * dy, dtype=f32) < f32(1.0)).any():
                        continue
                push_tmp(j, int(mod_label[i]))
                tot += 1
                if tot >= max_num_obj:
                    break
            start_id += 1
    uni = sorted(set(int(x) for x in tmp_sem))                            # (3) :2929-2972
    known = [any(sem_position[i] == u and obj_stat[i] for i in range(n_obj)) for u in uni]
    for u, kn in zip(uni, known):
        if kn:
            continue
        for j in range(len(tmp_sem)):
            if tmp_sem[j] == u:
                push_tmp(j, -2)
    O["p3d"] = [get3d_world(k, d, K4, Twc) for k, d in zip(O["keys"], O["depth"])]
    A = lambda l, dt, sh: np.array(l, dt).reshape(sh)
    return (dict(keys=A(S["keys"], f32, (-1, 2)), corres=A(S["corres"], f32, (-1, 2)), flow=A(S["flow"], f32, (-1, 2)), inlier_id=A(S["inlier_id"], np.int32, -1),
                 depth=A(S["depth"], f32, -1), p3d=A(S["p3d"], f32, (-1, 3))),
            dict(keys=A(O["keys"], f32, (-1, 2)), depth=A(O["depth"], f32, -1), corres=A(O["corres"], f32, (-1, 2)), flow=A(O["flow"], f32, (-1, 2)),
                 sem=A(O["sem"], np.int32, -1), inlier_id=A(O["inlier_id"], np.int32, -1), label=A(O["label"], np.int32, -1), p3d=A(O["p3d"], f32, (-1, 3))))
