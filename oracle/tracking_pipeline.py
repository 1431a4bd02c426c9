"""CPU restatement of the per-frame path as a whole (TEST INFRASTRUCTURE ONLY -- never imported by the product):
Tracking::GrabImageRGBD + Tracking::Track (src/Tracking.cc:164-648, 650-1212) sequenced over the stage oracles of this directory.
Parity unpinned (the reference cannot be built here and ships no fixtures).  Ground-truth error metrics, drawing and file output
are not restated; the ground-truth object list only gates which objects get a motion estimate, as in the reference (:767-810)."""
import numpy as np

from . import image_ops as io
from . import pyoracle as po
from . import tracking_ops as to

f32 = np.float32


def mul4(A, B):
    """cv::Mat A * B of two 4x4 CV_32F: OpenCV's own gemm (its small-matrix branch: float accumulation, left to right)."""
    import cv2
    return cv2.gemm(np.ascontiguousarray(A, f32), np.ascontiguousarray(B, f32), 1.0, None, 0.0)


def inv4(T):
    """Converter::toInvMatrix (src/Converter.cc:151-166)."""
    T = np.asarray(T, f32)
    I = np.eye(4, dtype=f32)
    I[:3, :3] = T[:3, :3].T
    import cv2
    I[:3, 3:4] = cv2.gemm(np.ascontiguousarray(T[:3, :3]), np.ascontiguousarray(T[:3, 3:4]), -1.0, None, 0.0, flags=cv2.GEMM_1_T)   # -R.t()*t: generic branch
    return I


class Frame:
    pass


class OracleTracker:
    def __init__(self, width=1242, height=375, K4=(721.5377, 721.5377, 609.5593, 172.8540), bf=387.5744, depth_factor=256.0, th_depth_bg=40.0,
                 th_depth_obj=25.0, max_track_bg=1200, max_track_obj=800, sf_mg_thres=0.12, sf_ds_thres=0.3, n_features=2500, scale_factor=1.2, n_levels=8,
                 ini_th_fast=20, min_th_fast=7, is_kitti=True, quirk=1, window_size=20, overlap_size=4, local_batch=True):
        self.w, self.h = width, height
        self.K4 = np.asarray(K4, f32)
        self.bf, self.factor = f32(bf), f32(depth_factor)
        self.th_bg, self.th_obj = f32(th_depth_bg), f32(th_depth_obj)
        self.max_bg, self.max_obj = max_track_bg, max_track_obj
        self.sf_mg, self.sf_ds = f32(sf_mg_thres), f32(sf_ds_thres)
        self.orb = io.OrbParams(n_features, scale_factor, n_levels, ini_th_fast, min_th_fast)
        self.kitti, self.quirk = is_kitti, quirk
        self.window, self.overlap, self.local_batch = window_size, overlap_size, local_batch
        self.local_ba = []
        self.first = True
        self.f_id, self.max_id = 0, 1
        self.velocity = None
        self.last = None
        self.cur = None
        self.map = dict(cameraPose_RF=[], rigidMotion_RF=[], cameraPose=[], assoSta=[], assoDyn=[], featLabel=[], featSta=[], depSta=[], p3dSta=[], featDyn=[], depDyn=[], p3dDyn=[],
                        rigidMotion=[], rmLabel=[])

    # ---- Frame::Frame (src/Frame.cc:61-260) ----
    def _build_frame(self, gray, depth, flow, mask):
        F = Frame()
        o = io.orb_extract(gray, self.orb, with_angle=False)
        F.keys = np.stack([o["x"], o["y"]], 1).astype(f32).reshape(-1, 2)
        keep, cx, cy, fu, fv, dep = io.filter_static(o["x"], o["y"], mask, depth, flow, self.th_bg)
        F.statKeysTmp = F.keys[keep]
        F.corres = np.stack([cx, cy], 1).astype(f32).reshape(-1, 2)
        F.flowNext = np.stack([fu, fv], 1).astype(f32).reshape(-1, 2)
        F.statDepthTmp = dep
        s = io.sample_objects(mask, depth, flow, self.th_obj, 4)
        F.objKeys = np.stack([s["x"], s["y"]], 1).astype(f32).reshape(-1, 2)
        F.objCorres = np.stack([s["cx"], s["cy"]], 1).astype(f32).reshape(-1, 2)
        F.objFlowNext = np.stack([s["fx"], s["fy"]], 1).astype(f32).reshape(-1, 2)
        F.objDepth = s["depth"].astype(f32); F.semObjLabel = s["label"].astype(np.int32)
        F.Tcw = np.eye(4, dtype=f32)
        F.nModLabel, F.nSemPosition, F.bObjStat, F.vObjMod = [], [], [], []
        return F

    def _flow(self, L, ids, objects, T_init):
        K = L.objKeys if objects else L.statKeys
        D = L.objDepth if objects else L.statDepth
        Fl = L.objFlowNext if objects else L.flowNext
        p = dict(pts=K[ids], depth=D[ids], flow=Fl[ids], K=self.K4, Tcw_last=L.Tcw, T_init=T_init)
        return po.flow2(p, mode=1 if objects else 0, quirk=self.quirk)

    def track(self, gray, depth_raw, flow, mask, gt_ids):
        """Returns Tcw (4x4 f32).  depth_raw / mask are not modified; the prepared depth and propagated mask are kept in self.depth / self.mask."""
        depth = io.depth_prep(depth_raw, self.bf, self.factor)                      # Tracking.cc:180-204
        mask = np.array(mask, np.int32)
        flow = np.asarray(flow, f32)
        L = self.last
        if not self.first:
            mask, self.warped = to.update_mask(mask, self.mask_last, self.flow_last, L.semObjLabel, L.objCorres)   # :236-245
        C = self._build_frame(gray, depth, flow, mask)
        if not self.first:                                                             # :254-312
            C.statKeys = L.corres.copy()
            n = len(C.statKeys)
            C.statDepth = np.full(n, -1, f32)
            for i in range(n):
                u, v = int(C.statKeys[i, 0]), int(C.statKeys[i, 1])
                if 0 < u < self.w - 1 and 0 < v < self.h - 1 and depth[v, u] > 0:
                    C.statDepth[i] = depth[v, u]
            self.tmp = dict(keys=C.objKeys, depth=C.objDepth, sem=C.semObjLabel, flow=C.objFlowNext, corres=C.objCorres)
            C.objKeys = L.objCorres.copy()
            n = len(C.objKeys)
            C.objDepth = np.full(n, 0.1, f32); C.semObjLabel = np.zeros(n, np.int32)
            for i in range(n):
                u, v = int(C.objKeys[i, 0]), int(C.objKeys[i, 1])
                if 0 < u < self.w - 1 and 0 < v < self.h - 1 and depth[v, u] < self.th_obj and depth[v, u] > 0:
                    C.objDepth[i] = depth[v, u]; C.semObjLabel[i] = mask[v, u]
        C.semPosiGt = list(gt_ids)
        C.objLabel = np.full(len(C.objKeys), -2, np.int32)
        if self.first:                                                                 # Initialization (:1215-1276)
            C.Tcw = np.eye(4, dtype=f32)
            C.stat3DTmp = np.array([to.get3d_world(k, d, self.K4, np.eye(4, dtype=f32)) for k, d in zip(C.statKeysTmp, C.statDepthTmp)], f32).reshape(-1, 3)
            C.obj3D = np.array([to.get3d_world(k, d, self.K4, np.eye(4, dtype=f32)) for k, d in zip(C.objKeys, C.objDepth)], f32).reshape(-1, 3)
            self.map["cameraPose"].append(np.eye(4, dtype=f32)); self.map["cameraPose_RF"].append(np.eye(4, dtype=f32))
            self.first = False
        else:
            self._track(C, L, depth, flow, mask)
        self.map["featSta"].append(C.statKeysTmp.copy()); self.map["featDyn"].append(C.objKeys.copy())
        self.map["depSta"].append(C.statDepthTmp.copy()); self.map["depDyn"].append(C.objDepth.copy())
        self.map["p3dSta"].append(np.asarray(C.stat3DTmp, f32).reshape(-1, 3).copy()); self.map["p3dDyn"].append(np.asarray(C.obj3D, f32).reshape(-1, 3).copy())
        C.statKeys = C.statKeysTmp; C.statDepth = C.statDepthTmp                      # :1006-1014
        self.last, self.cur = C, C
        self.mask_last, self.flow_last, self.depth, self.mask = mask, flow, depth, mask
        if (self.local_batch and self.window > self.overlap and (self.f_id - self.overlap + 1) % (self.window - self.overlap) == 0
                and self.f_id >= self.window - 1):                                   # Tracking.cc:1150-1160
            self.batch_optimize("partial")
        self.f_id += 1
        return C.Tcw.copy()

    def _track(self, C, L, depth, flow, mask):
        Ns = len(C.statKeys)
        # GetInitModelCam (:1614-1715)
        pre3 = io.unproject_world(L.statKeys[:, 0], L.statKeys[:, 1], L.statDepth, self.K4, L.Tcw).reshape(-1, 3)
        mm = L.Tcw.copy() if self.velocity is None else mul4(self.velocity, L.Tcw)
        T0, sub, info = to.init_model(pre3, C.statKeys, self.K4, mm)
        self.cam_info = info
        C.Tcw = T0
        tm_sub = np.asarray(sub, np.int32).copy()
        if len(tm_sub) >= 3:                                                           # PoseOptimizationFlow2Cam (src/Optimizer.cc:2333-2542)
            r = self._flow(L, tm_sub, False, C.Tcw)
            C.Tcw = r["T"].astype(f32)
            for i, idx in enumerate(tm_sub.copy()):
                if r["inlier"][i]:
                    C.statKeys[idx] = (L.statKeys[idx].astype(np.float64) + r["flow"][i]).astype(f32)
                else:
                    tm_sub[i] = -1
        self.tm_sub = tm_sub
        self.velocity = mul4(C.Tcw, inv4(L.Tcw))                                       # :700-706
        # GetSceneFlowObj (:1278-1364)
        No = len(C.objKeys)
        if No:
            C.flow3d, valid = io.scene_flow(L.objKeys[:, 0], L.objKeys[:, 1], L.objDepth, L.Tcw, C.objKeys[:, 0], C.objKeys[:, 1], C.objDepth, C.Tcw, self.K4,
                                            L.semObjLabel, C.semObjLabel)
            C.objLabel[~valid] = -1
        else:
            C.flow3d = np.zeros((0, 3), f32)
        # DynObjTracking (:1366-1612)
        C.objLabel, objs, C.nModLabel, C.nSemPosition, self.max_id = to.dyn_obj_tracking(
            C.semObjLabel, C.objLabel, C.objKeys, C.objDepth, C.flow3d, L.semObjLabel, L.nSemPosition, L.bObjStat, L.nModLabel, self.h, self.w,
            25 if self.kitti else 0, 50 if self.kitti else 0, self.sf_mg, self.sf_ds, self.th_obj, self.f_id, self.max_id)
        nobj = len(objs)
        C.bObjStat = [True] * nobj; C.vObjMod = [np.eye(4, dtype=f32) for _ in range(nobj)]
        C.vnObjInlierID = [[] for _ in range(nobj)]
        self.obj_info = []
        Twc = inv4(C.Tcw)
        for i in range(nobj):                                                          # :760-1003
            sem = C.nSemPosition[i]
            if sem not in L.semPosiGt or sem not in C.semPosiGt:
                C.bObjStat[i] = False; C.vnObjInlierID[i] = list(objs[i]); continue
            ids = np.asarray(objs[i], np.int32)
            pre3 = io.unproject_world(L.objKeys[ids, 0], L.objKeys[ids, 1], L.objDepth[ids], self.K4, L.Tcw).reshape(-1, 3)
            pre = next((k for k in range(len(L.nModLabel)) if L.nModLabel[k] == C.nModLabel[i]), -1)
            mm = mul4(C.Tcw, L.vObjMod[pre]) if pre != -1 else None
            Ti, sub, info = to.init_model(pre3, C.objKeys[ids], self.K4, mm)           # GetInitModelObj (:1717-1849)
            self.obj_info.append(info)
            kept = np.zeros(len(ids), bool); kept[sub] = True
            C.objLabel[ids[~kept]] = -1
            ids_in = ids[sub]
            if len(ids_in) < 50:
                C.bObjStat[i] = False; C.vnObjInlierID[i] = ids_in.tolist(); continue
            r = self._flow(L, ids_in, True, Ti)                                        # PoseOptimizationFlow2 (src/Optimizer.cc:2755-2972)
            inl = []
            for q, idx in enumerate(ids_in):
                if r["inlier"][q]:
                    C.objKeys[idx] = (L.objKeys[idx].astype(np.float64) + r["flow"][q]).astype(f32); inl.append(int(idx))
                else:
                    C.objLabel[idx] = -1
            C.vObjMod[i] = mul4(Twc, r["T"].astype(f32))
            C.vnObjInlierID[i] = inl
        # RenewFrameInfo (:2660-2995)
        S, O = to.renew_frame_info(mask, depth, flow, tm_sub, C.statKeys, C.keys, self.max_bg, C.vnObjInlierID, C.bObjStat, C.nSemPosition, C.nModLabel,
                                   C.objKeys, C.objLabel, self.tmp["keys"], self.tmp["depth"], self.tmp["sem"], self.tmp["flow"], self.tmp["corres"], self.max_obj,
                                   self.K4, Twc)
        C.statKeysTmp, C.corres, C.flowNext, C.staInlierID, C.statDepthTmp, C.stat3DTmp = S["keys"], S["corres"], S["flow"], S["inlier_id"], S["depth"], S["p3d"]
        C.objKeys, C.objDepth, C.objCorres, C.objFlowNext = O["keys"], O["depth"], O["corres"], O["flow"]
        C.semObjLabel, C.dynInlierID, C.objLabel, C.obj3D = O["sem"], O["inlier_id"], O["label"], O["p3d"]
        self.map["cameraPose"].append(Twc); self.map["cameraPose_RF"].append(Twc.copy()); self.map["assoSta"].append(C.staInlierID.copy()); self.map["assoDyn"].append(C.dynInlierID.copy())
        self.map["featLabel"].append(C.objLabel.copy())
        self.map["rigidMotion"].append([inv4(self.velocity)] + [C.vObjMod[i] for i in range(nobj) if C.bObjStat[i]])
        self.map["rigidMotion_RF"].append([T.copy() for T in self.map["rigidMotion"][-1]])
        self.map["rmLabel"].append([0] + [C.nModLabel[i] for i in range(nobj) if C.bObjStat[i]])

    def batch_optimize(self, mode):
        """PartialBatchOptimization / FullBatchOptimization on the map, with the write-back of src/Optimizer.cc:1058-1101 (partial: into
        vmCameraPose / vmRigidMotion) / :2094-2172 (full: into vmCameraPose_RF[i + 1] / vmRigidMotion_RF, the estimates themselves stay)."""
        from . import map_graph as mg
        g, meta = mg.build_graph(self.map, self.K4, mode, self.window)
        r = po.ba_optimize(g, max_iters=meta["max_iters"], gain_threshold=meta["gain"])
        m = self.map

        def from_iso(T):
            out = np.eye(4, dtype=f32)
            out[:3, :3] = mg.to_iso_matrix(T)[:9].reshape(3, 3).astype(f32); out[:3, 3] = T[9:].astype(f32)
            return out
        cam_out = m["cameraPose"] if mode == "partial" else m["cameraPose_RF"]
        mot_out = m["rigidMotion"] if mode == "partial" else m["rigidMotion_RF"]
        for i, v in enumerate(meta["cam_vid"]):
            if v != -1 and (mode == "partial" or i > 0):
                cam_out[i] = from_iso(r["se3"][v])
        for i in range(len(m["p3dSta"])):
            for j, pidx in enumerate(meta["makS"][i]):
                if pidx != -1:
                    m["p3dSta"][i][j] = r["pt"][pidx].astype(f32)
            for j, pidx in enumerate(meta["makD"][i]):
                if pidx != -1:
                    m["p3dDyn"][i][j] = r["pt"][pidx].astype(f32)
        for i in range(len(m["cameraPose"]) - 1):
            if mode == "partial" and meta["cam_vid"][i] != -1 and meta["cam_vid"][i + 1] != -1:
                m["rigidMotion"][i][0] = mul4(inv4(m["cameraPose"][i]), m["cameraPose"][i + 1])
            for j in range(1, len(meta["mot_vid"][i])):
                if meta["mot_vid"][i][j] != -1:
                    mot_out[i][j] = from_iso(r["se3"][meta["mot_vid"][i][j]])
        self.local_ba.append(r["iters"])
        return r
