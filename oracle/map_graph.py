"""CPU restatement of the graph construction of the batch optimisers (TEST INFRASTRUCTURE ONLY):
Optimizer::FullBatchOptimization (src/Optimizer.cc:1232-1767) and Optimizer::PartialBatchOptimization (:42-805), turning the
tracker's Map slices into the arrays oracle/ba_lm.c optimises.  Parity unpinned (no fixtures in the reference).

Map fields used (include/Map.h:34-84): vpFeatSta, vfDepSta, vp3DPointSta, vnAssoSta, vpFeatDyn, vfDepDyn, vp3DPointDyn, vnAssoDyn,
vnFeatLabel, vmCameraPose, vmRigidMotion, vnRMLabel; TrackletSta / TrackletDyn / nObjID come from tracking_ops.tracklets_build.
Situations in which the reference dereferences a null vertex (a track whose previous position was never given a vertex) are
skipped here and in the product the same way: the edge is not created."""
import numpy as np

from . import pyoracle as po
from . import tracking_ops as to

FULL = dict(sigma2_cam=0.001, sigma2_3d_sta=80.0, sigma2_obj_smo=0.001, sigma2_obj=100.0, sigma2_3d_dyn=80.0, prior_w=100000.0, static_only=False,
            max_iters=300, gain=1e-4)
PARTIAL = dict(sigma2_cam=0.0001, sigma2_3d_sta=16.0, sigma2_obj_smo=0.1, sigma2_obj=20.0, sigma2_3d_dyn=16.0, prior_w=1.0 / 0.0000001, static_only=True,
               max_iters=100, gain=1e-3)
HUBER = 0.0001


def to_iso(T):
    """Converter::toSE3Quat (src/Converter.cc:25-35) followed by the SE3Quat -> Isometry3d cast of VertexSE3::setEstimate."""
    T = np.asarray(T, np.float32).astype(np.float64)
    out = np.zeros(12)
    R = np.ascontiguousarray(T[:3, :3]).reshape(9); t = np.ascontiguousarray(T[:3, 3])
    import ctypes as C
    po.lib().vdo_oracle_iso_from_Rt_via_quat(R.ctypes.data_as(C.POINTER(C.c_double)), t.ctypes.data_as(C.POINTER(C.c_double)), out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def to_iso_matrix(T12):
    """getEstimateData -> Quaterniond -> rotation matrix (src/Optimizer.cc:2094-2110): rotation part re-normalised through the quaternion."""
    import ctypes as C
    T12 = np.ascontiguousarray(T12, np.float64)
    out = np.zeros(12)
    R = np.ascontiguousarray(T12[:9]); t = np.ascontiguousarray(T12[9:])
    po.lib().vdo_oracle_iso_from_Rt_via_quat(R.ctypes.data_as(C.POINTER(C.c_double)), t.ctypes.data_as(C.POINTER(C.c_double)), out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def get3d_camera(key, depth, K4):
    f32 = np.float32
    K = np.asarray(K4, f32)
    invfx, invfy = f32(1.0) / K[0], f32(1.0) / K[1]
    z = f32(depth)
    return np.array([f32(f32(f32(key[0]) - K[2]) * z) * invfx, f32(f32(f32(key[1]) - K[3]) * z) * invfy, z], np.float64)


def build_graph(m, K4, mode="full", window=20):
    """m: dict with the Map fields (lists per frame).  Returns (graph dict in vdo_slam_b200.synth.make_batch_graph layout, meta)."""
    c = FULL if mode == "full" else PARTIAL
    N = len(m["featSta"])
    sta_tracks, _ = to.tracklets_build(m["assoSta"])
    dyn_tracks, obj_id = to.tracklets_build(m["assoDyn"], m["featLabel"])
    labS = [np.full(len(m["featSta"][i]), -1, np.int64) for i in range(N)]; makS = [np.full(len(m["featSta"][i]), -1, np.int64) for i in range(N)]
    labD = [np.full(len(m["featDyn"][i]), -1, np.int64) for i in range(N)]; makD = [np.full(len(m["featDyn"][i]), -1, np.int64) for i in range(N)]
    for ti, tr in enumerate(sta_tracks):
        if len(tr) < 3:
            continue
        for (f, j) in tr:
            labS[f][j] = ti
    for ti, tr in enumerate(dyn_tracks):
        if len(tr) < 3:
            continue
        for (f, j) in tr:
            labD[f][j] = ti
    se3, pt = [], []
    prior_v, prior_Z, prior_w = [], [], []
    se3e_ij, se3e_Z, se3e_w = [], [], []
    obs_cp, obs_z, obs_w = [], [], []
    ter_pph, ter_w = [], []
    cam_vid = [-1] * N
    mot_vid = [[-1] * len(m["rmLabel"][i]) for i in range(N - 1)]
    start = 0 if mode == "full" else N - window
    ident = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], np.float64)

    def add_obs(cam, p, key, dep, w):
        obs_cp.append((cam, p)); obs_z.append(get3d_camera(key, dep, K4)); obs_w.append(w)

    pre = -1
    for i in range(start, N):
        cur = len(se3); se3.append(to_iso(m["cameraPose"][i])); cam_vid[i] = cur
        if cur == 0 and (mode == "full" or N == window):
            prior_v.append(cur); prior_Z.append(to_iso(m["cameraPose"][i])); prior_w.append(c["prior_w"])
        if i != start:
            se3e_ij.append((pre, cur)); se3e_Z.append(to_iso(m["rigidMotion"][i - 1][0])); se3e_w.append(1.0 / float(np.float32(c["sigma2_cam"])))
        # ---- static points ----
        for j in range(len(labS[i])):
            tid = labS[i][j]
            if tid == -1:
                continue
            tr = sta_tracks[tid]
            pos = next((k for k, (f, q) in enumerate(tr) if f == i and q == j), -1)
            if pos == -1:
                continue
            w = 1.0 / float(np.float32(c["sigma2_3d_sta"]))
            if pos == 0:
                p = len(pt); pt.append(np.asarray(m["p3dSta"][i][j], np.float32).astype(np.float64))
                add_obs(cur, p, m["featSta"][i][j], m["depSta"][i][j], w); makS[i][j] = p
            else:
                pf, pj = tr[pos - 1]
                p = makS[pf][pj]
                if p == -1:
                    continue
                add_obs(cur, p, m["featSta"][i][j], m["depSta"][i][j], w); makS[i][j] = p
        # ---- object motions and dynamic points ----
        if not c["static_only"] and i > 0:
            obj_uid = []
            for j in range(1, len(m["rigidMotion"][i - 1])):
                v = len(se3); se3.append(ident.copy())
                if i > 2:
                    trace = next((k for k in range(len(m["rmLabel"][i - 2])) if m["rmLabel"][i - 2][k] == m["rmLabel"][i - 1][j]), -1)
                    if trace != -1 and mot_vid[i - 2][trace] != -1:
                        se3e_ij.append((mot_vid[i - 2][trace], v)); se3e_Z.append(ident.copy()); se3e_w.append(1.0 / float(np.float32(c["sigma2_obj_smo"])))
                obj_uid.append(v); mot_vid[i - 1][j] = v
            for j in range(len(labD[i])):
                tid = labD[i][j]
                if tid == -1:
                    continue
                tr = dyn_tracks[tid]
                pos = next((k for k, (f, q) in enumerate(tr) if f == i and q == j), -1)
                if pos == -1:
                    continue
                objv = -1
                for k in range(1, len(m["rmLabel"][i - 1])):
                    if m["rmLabel"][i - 1][k] == obj_id[tid]:
                        objv = obj_uid[k - 1]; break
                if objv == -1 and pos != 0:
                    continue
                p = len(pt); pt.append(np.asarray(m["p3dDyn"][i][j], np.float32).astype(np.float64))
                add_obs(cur, p, m["featDyn"][i][j], m["depDyn"][i][j], 1.0 / float(np.float32(c["sigma2_3d_dyn"])))
                if pos != 0:
                    pf, pj = tr[pos - 1]
                    if makD[pf][pj] != -1:
                        ter_pph.append((makD[pf][pj], p, objv)); ter_w.append(1.0 / float(np.float32(c["sigma2_obj"])))
                makD[i][j] = p
        elif not c["static_only"] and i == 0:
            for j in range(len(labD[i])):
                if labD[i][j] == -1:
                    continue
                p = len(pt); pt.append(np.asarray(m["p3dDyn"][i][j], np.float32).astype(np.float64))
                add_obs(cur, p, m["featDyn"][i][j], m["depDyn"][i][j], 1.0 / float(np.float32(c["sigma2_3d_dyn"]))); makD[i][j] = p
        pre = cur
    A = lambda l, dt, sh: np.ascontiguousarray(np.array(l, dt).reshape(sh))
    g = dict(se3=A(se3, np.float64, (-1, 12)), pt=A(pt, np.float64, (-1, 3)),
             prior_v=A(prior_v, np.int32, -1), prior_Z=A(prior_Z, np.float64, (-1, 12)), prior_w=A(prior_w, np.float64, -1),
             se3e_ij=A(se3e_ij, np.int32, (-1, 2)), se3e_Z=A(se3e_Z, np.float64, (-1, 12)), se3e_w=A(se3e_w, np.float64, -1),
             se3e_delta=np.full(len(se3e_w), np.float64(np.float32(HUBER))),
             obs_cp=A(obs_cp, np.int32, (-1, 2)), obs_z=A(obs_z, np.float64, (-1, 3)), obs_w=A(obs_w, np.float64, -1),
             obs_delta=np.full(len(obs_w), np.float64(np.float32(HUBER))),
             ter_pph=A(ter_pph, np.int32, (-1, 3)), ter_w=A(ter_w, np.float64, -1), ter_delta=np.full(len(ter_w), np.float64(np.float32(HUBER))))
    meta = dict(cam_vid=cam_vid, mot_vid=mot_vid, makS=makS, makD=makD, max_iters=c["max_iters"], gain=c["gain"])
    return g, meta
