"""Host-side mirror of the Hamming matching core of ORB_SLAM3::ORBmatcher
(R/include/cloud_edge_slam_lib/ORBmatcher.h:36-103, R/lib_src/ORBmatcher.cc:31-33, :1830-1844) over the C ABI.

The GPU returns the raw (best index, best distance, second-best distance) triple of the scan every reference
matcher runs; the call-site specific acceptance tests (TH_LOW / TH_HIGH, '<' vs '<=', which side of the ratio
is cast to float) stay here on the host exactly as written at each call site.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import KP_DTYPE, check, ptr


class ORBmatcher:
    TH_HIGH = 100        # ORBmatcher.cc:31
    TH_LOW = 50          # ORBmatcher.cc:32
    HISTO_LENGTH = 30    # ORBmatcher.cc:33

    def __init__(self, nnratio=0.6, checkOri=True, device=0):
        self._L = _lib.lib()
        self._m = C.c_void_p()
        check(self._L.rumi_match_create(C.byref(self._m), int(device)))
        self.mfNNratio = np.float32(nnratio)
        self.mbCheckOrientation = bool(checkOri)
        self.device = device

    def close(self):
        if getattr(self, "_m", None) is not None and self._m:
            self._L.rumi_match_destroy(self._m)
            self._m = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def last_path(self):
        """Kernel of the last top-2 call: 'popc' (LOP3+POPC) or 'umma' (tcgen05 int8 tensor cores, TMEM accumulators)."""
        return {1: "popc", 3: "umma"}.get(int(self._L.rumi_match_last_path(self._m)), "none")

    @staticmethod
    def DescriptorDistance(a, b):
        """ORBmatcher::DescriptorDistance (ORBmatcher.cc:1830-1844) for one pair of 32-byte rows."""
        a = np.ascontiguousarray(a, np.uint8).reshape(32)
        b = np.ascontiguousarray(b, np.uint8).reshape(32)
        return _lib.lib().rumi_descriptor_distance(ptr(a), ptr(b))

    # ---- raw top-2 (host buffers) ----
    def top2(self, Q, T):
        Q = np.ascontiguousarray(Q, np.uint8).reshape(-1, 32)
        T = np.ascontiguousarray(T, np.uint8).reshape(-1, 32)
        i1 = np.zeros(len(Q), np.int32)
        d1 = np.zeros(len(Q), np.uint16)
        d2 = np.zeros(len(Q), np.uint16)
        check(self._L.rumi_hamming_top2(self._m, ptr(Q), len(Q), ptr(T), len(T), ptr(i1), ptr(d1), ptr(d2)))
        return i1, d1, d2

    # ---- raw top-2 (CUDA torch tensors; train indices offset by t_base for a train shard) ----
    def top2_device(self, Q, T, t_base=0, out=None, sync=True):
        import torch
        assert Q.is_cuda and T.is_cuda and Q.dtype == torch.uint8 and T.dtype == torch.uint8
        nq, nt = Q.shape[0], T.shape[0]
        if out is None:
            out = (torch.empty(nq, dtype=torch.int32, device=Q.device),
                   torch.empty(nq, dtype=torch.int16, device=Q.device),
                   torch.empty(nq, dtype=torch.int16, device=Q.device))
        i1, d1, d2 = out
        st = _lib.torch_stream()                       # Q / T were produced on torch's stream, the results are used there
        check(self._L.rumi_match_wait_stream(self._m, st))
        check(self._L.rumi_hamming_top2_device(self._m, ptr(Q), nq, ptr(T), nt, int(t_base), ptr(i1), ptr(d1), ptr(d2),
                                               1 if sync else 0))
        if not sync:
            check(self._L.rumi_match_signal_stream(self._m, st))
        return i1, d1, d2

    # ---- multi-GPU all-pairs top-2: train set sharded over the ranks, NCCL candidate all-gather ----
    def comm_init(self, unique_id, rank, nranks):
        """unique_id: the 128 bytes rank 0 got from ORBmatcher.nccl_unique_id(), distributed by the caller."""
        buf = np.frombuffer(bytes(unique_id), np.uint8).copy()
        assert buf.size == 128
        check(self._L.rumi_match_comm_init(self._m, ptr(buf), int(rank), int(nranks)))
        self.rank, self.nranks = int(rank), int(nranks)

    @staticmethod
    def nccl_unique_id():
        buf = np.zeros(128, np.uint8)
        check(_lib.lib().rumi_nccl_unique_id(ptr(buf)))
        return buf.tobytes()

    def top2_sharded(self, Q, T_local, t_base, out=None, sync=False):
        """Q: all queries, replicated on every rank; T_local: this rank's contiguous train range starting at global row
        t_base.  Every rank gets the global (idx1, d1, d2).  Scan, exchange and fold are enqueued on the matcher's
        stream; torch's current stream is ordered behind them."""
        import torch
        assert Q.is_cuda and T_local.is_cuda and Q.dtype == torch.uint8 and T_local.dtype == torch.uint8
        nq, nt = Q.shape[0], T_local.shape[0]
        if out is None:
            out = (torch.empty(nq, dtype=torch.int32, device=Q.device),
                   torch.empty(nq, dtype=torch.int16, device=Q.device),
                   torch.empty(nq, dtype=torch.int16, device=Q.device))
        i1, d1, d2 = out
        st = _lib.torch_stream()
        check(self._L.rumi_match_wait_stream(self._m, st))
        check(self._L.rumi_hamming_top2_sharded(self._m, ptr(Q), nq, ptr(T_local), nt, int(t_base), ptr(i1), ptr(d1),
                                                ptr(d2), 1 if sync else 0))
        if not sync:
            check(self._L.rumi_match_signal_stream(self._m, st))
        return i1, d1, d2

    def pack_device(self, i1, d1, d2, out=None, sync=True):
        import torch
        nq = i1.shape[0]
        if out is None:
            out = torch.empty(nq, dtype=torch.int64, device=i1.device)
        st = _lib.torch_stream()
        check(self._L.rumi_match_wait_stream(self._m, st))
        check(self._L.rumi_top2_pack_device(self._m, ptr(i1), ptr(d1), ptr(d2), nq, ptr(out), 1 if sync else 0))
        if not sync:
            check(self._L.rumi_match_signal_stream(self._m, st))
        return out

    def merge_device(self, packed, nshards, nq, out=None, sync=True):
        import torch
        if out is None:
            out = (torch.empty(nq, dtype=torch.int32, device=packed.device),
                   torch.empty(nq, dtype=torch.int16, device=packed.device),
                   torch.empty(nq, dtype=torch.int16, device=packed.device))
        i1, d1, d2 = out
        st = _lib.torch_stream()                       # `packed` may come from a collective on torch's stream
        check(self._L.rumi_match_wait_stream(self._m, st))
        check(self._L.rumi_top2_merge_device(self._m, ptr(packed), int(nshards), int(nq), ptr(i1), ptr(d1), ptr(d2),
                                             1 if sync else 0))
        if not sync:
            check(self._L.rumi_match_signal_stream(self._m, st))
        return i1, d1, d2

    def timer_start(self):
        check(self._L.rumi_match_timer_start(self._m))

    def timer_stop(self):
        ms = C.c_float(0)
        check(self._L.rumi_match_timer_stop(self._m, C.byref(ms)))
        return ms.value

    def launch_count(self, reset=False):
        return int(self._L.rumi_match_launch_count(self._m, 1 if reset else 0))

    # ---- acceptance rules of the reference call sites, applied to the raw triple ----
    def accept_bow(self, d1, d2, th=None):
        """SearchByBoW KF->F (ORBmatcher.cc:290-291): best1 <= TH_LOW and (float)best1 < ratio*(float)best2."""
        th = self.TH_LOW if th is None else th
        d1 = np.asarray(d1).astype(np.int64)
        d2 = np.asarray(d2).astype(np.int64)
        return (d1 <= th) & (d1.astype(np.float32) < self.mfNNratio * d2.astype(np.float32))

    @staticmethod
    def accept_knn_ratio(d1, d2, ratio=0.7):
        """ComputeStereoFishEyeMatches (Frame.cc:1146): d0 < d1 * 0.7 (float distance, double constant)."""
        d1 = np.asarray(d1).astype(np.float32).astype(np.float64)
        d2 = np.asarray(d2).astype(np.float32).astype(np.float64)
        return d1 < d2 * ratio

    def match_bow(self, Q, T):
        i1, d1, d2 = self.top2(Q, T)
        ok = self.accept_bow(d1, d2) & (i1 >= 0)
        return np.where(ok, i1, -1), d1, d2

    # ---- descriptor association of many key-frame pairs in one launch (submap merge) ----
    def top2_pairs(self, descs_a, descs_b):
        """descs_a[i], descs_b[i]: descriptor matrices of the i-th matched key-frame pair (the pairs
        CloudMerging.cc:503-551 walks).  Returns a list of (idx1, d1, d2) per pair, idx1 = feature index inside
        descs_b[i]; identical to calling top2(descs_a[i], descs_b[i]) pair by pair."""
        assert len(descs_a) == len(descs_b)
        A = [np.ascontiguousarray(a, np.uint8).reshape(-1, 32) for a in descs_a]
        B = [np.ascontiguousarray(b, np.uint8).reshape(-1, 32) for b in descs_b]
        qo = np.concatenate([[0], np.cumsum([len(a) for a in A])]).astype(np.int64)
        to = np.concatenate([[0], np.cumsum([len(b) for b in B])]).astype(np.int64)
        Q = np.concatenate(A) if A else np.zeros((0, 32), np.uint8)
        T = np.concatenate(B) if B else np.zeros((0, 32), np.uint8)
        segs = np.array([[qo[i], len(A[i]), to[i], len(B[i])] for i in range(len(A))], np.int32).reshape(-1, 4)
        i1 = np.zeros(len(Q), np.int32)
        d1 = np.zeros(len(Q), np.uint16)
        d2 = np.zeros(len(Q), np.uint16)
        check(self._L.rumi_hamming_top2_pairs(self._m, ptr(Q), len(Q), ptr(T), len(T), ptr(segs), len(segs), ptr(i1),
                                              ptr(d1), ptr(d2)))
        return [(i1[qo[i]:qo[i + 1]], d1[qo[i]:qo[i + 1]], d2[qo[i]:qo[i + 1]]) for i in range(len(A))]

    def match_keyframe_pairs(self, descs_a, descs_b, th=None):
        """Per pair: matches12[q] = feature of key frame 2 accepted by the SearchByBoW rule (ORBmatcher.cc:290-291:
        best1 <= TH_LOW and best1 < mfNNratio * best2), else -1."""
        out = []
        for i1, d1, d2 in self.top2_pairs(descs_a, descs_b):
            ok = self.accept_bow(d1, d2, th) & (i1 >= 0)
            out.append(np.where(ok, i1, -1))
        return out

    # ---- candidate-list matching (the form SearchByProjection / SearchForInitialization use) ----
    def candidates(self, Q, T, cand_off, cand_idx, top2=False):
        """dist[k] = DescriptorDistance(Q[q], T[cand_idx[k]]) for every entry k of query q's list
        cand_idx[cand_off[q]:cand_off[q + 1]]; with top2 also (idx1, d1, idx2, d2) per query in list order."""
        Q = np.ascontiguousarray(Q, np.uint8).reshape(-1, 32)
        T = np.ascontiguousarray(T, np.uint8).reshape(-1, 32)
        off = np.ascontiguousarray(cand_off, np.int32)
        idx = np.ascontiguousarray(cand_idx, np.int32)
        dist = np.zeros(max(len(idx), 1), np.uint16)
        nq = len(Q)
        if top2:
            i1, i2 = np.zeros(nq, np.int32), np.zeros(nq, np.int32)
            d1, d2 = np.zeros(nq, np.uint16), np.zeros(nq, np.uint16)
            check(self._L.rumi_hamming_candidates(self._m, ptr(Q), nq, ptr(T), len(T), ptr(off), ptr(idx), ptr(dist), ptr(i1),
                                                  ptr(d1), ptr(i2), ptr(d2)))
            return dist[:len(idx)], (i1, d1, i2, d2)
        check(self._L.rumi_hamming_candidates(self._m, ptr(Q), nq, ptr(T), len(T), ptr(off), ptr(idx), ptr(dist), None, None,
                                              None, None))
        return dist[:len(idx)]

    # ---- ORBmatcher::SearchForInitialization (ORBmatcher.cc:581-680) ----
    def SearchForInitialization(self, keys1, desc1, keys2, desc2, bounds, vbPrevMatched, windowSize=10):
        """keys*: mvKeysUn of F1 / F2 (KP_DTYPE), bounds = F2's (mnMinX, mnMinY, mnMaxX, mnMaxY).  Returns (nmatches,
        vnMatches12, updated vbPrevMatched).  The GPU computes the distance of every (level-0 key point of F1, window
        candidate of F2) pair in one launch; the acceptance -- vMatchedDistance skip (:617), re-assignment of a taken
        feature (:631-634), rotation histogram -- is replayed here in the reference's order."""
        k1, k2 = np.ascontiguousarray(keys1, KP_DTYPE), np.ascontiguousarray(keys2, KP_DTYPE)
        prev = np.array(vbPrevMatched, np.float32).reshape(-1, 2)
        n1, n2 = len(k1), len(k2)
        m12 = np.full(n1, -1, np.int32)
        q = np.flatnonzero(k1["octave"] <= 0)                                 # level1 > 0: continue (:596-597)
        grid = FrameGrid(k2, bounds)
        off, idx = grid.candidate_lists(prev[q], float(windowSize), k1["octave"][q], k1["octave"][q])
        dist = self.candidates(np.ascontiguousarray(desc1, np.uint8).reshape(-1, 32)[q], desc2, off, idx)
        matched_dist = np.full(n2, np.iinfo(np.int32).max, np.int64)
        m21 = np.full(n2, -1, np.int32)
        rot_hist = [[] for _ in range(self.HISTO_LENGTH)]
        nmatches = 0
        ratio = np.float32(self.mfNNratio)
        big = np.iinfo(np.int32).max
        for qi, i1 in enumerate(q):
            best, best2, best_idx = big, big, -1
            for p in range(off[qi], off[qi + 1]):
                i2, d = int(idx[p]), int(dist[p])
                if matched_dist[i2] <= d:
                    continue
                if d < best:
                    best2, best, best_idx = best, d, i2
                elif d < best2:
                    best2 = d
            if best <= self.TH_LOW and np.float32(best) < np.float32(best2) * ratio:
                if m21[best_idx] >= 0:
                    m12[m21[best_idx]] = -1
                    nmatches -= 1
                m12[i1] = best_idx
                m21[best_idx] = i1
                matched_dist[best_idx] = best
                nmatches += 1
                if self.mbCheckOrientation:
                    rot_hist[self._rot_bin(k1["angle"][i1], k2["angle"][best_idx])].append(int(i1))
        if self.mbCheckOrientation:
            keep = _three_maxima(rot_hist)
            for i in range(self.HISTO_LENGTH):
                if i in keep:
                    continue
                for i1 in rot_hist[i]:
                    if m12[i1] >= 0:
                        m12[i1] = -1
                        nmatches -= 1
        ok = m12 >= 0
        prev[ok, 0] = k2["x"][m12[ok]]
        prev[ok, 1] = k2["y"][m12[ok]]
        return nmatches, m12, prev

    # ---- ORBmatcher::SearchByProjection(Frame&, vpMapPoints, th, ...) (ORBmatcher.cc:39-118, mono frame) ----
    def SearchByProjection(self, keysF, descF, scale_factors, bounds, proj, level, view_cos, descMP, has_obs, th=3.0, *,
                           occupied=None, u_right=None, proj_r=None, keys_right=None, in_view=None, in_view_r=None,
                           level_r=None, view_cos_r=None, l2r=None, r2l=None):
        """Map points in view (mbTrackInView) with their projection (mTrackProjX/Y), predicted level, viewing cosine,
        descriptor and Observations() > 0 flag.  Returns (nmatches, frame_match[j] = map point index or -1).
        Keyword extras = the rest of the reference function (ORBmatcher.cc:39-189), see _search_by_projection_full:
        occupied (features that already hold a point with observations), u_right + proj_r (mvuRight / mTrackProjXR:
        right-image gate of rectified stereo and RGB-D frames), keys_right ... r2l (stereo-fisheye rigs)."""
        if any(a is not None for a in (occupied, u_right, keys_right, in_view, in_view_r)):
            return self._search_by_projection_full(keysF, descF, scale_factors, bounds, proj, level, view_cos, descMP, has_obs, th,
                                                   occupied, u_right, proj_r, keys_right, in_view, in_view_r, level_r,
                                                   view_cos_r, l2r, r2l)
        kF = np.ascontiguousarray(keysF, KP_DTYPE)
        sf = np.asarray(scale_factors, np.float32)
        proj = np.asarray(proj, np.float32).reshape(-1, 2)
        level = np.asarray(level, np.int32)
        r = np.where(np.asarray(view_cos, np.float32) > 0.998, np.float32(2.5), np.float32(4.0)).astype(np.float32)
        if th != 1.0:
            r = (r * np.float32(th)).astype(np.float32)
        grid = FrameGrid(kF, bounds)
        off, idx = grid.candidate_lists(proj, (r * sf[level]).astype(np.float32), level - 1, level)
        dist = self.candidates(descMP, descF, off, idx)
        fm = np.full(len(kF), -1, np.int32)
        ratio = np.float32(self.mfNNratio)
        nmatches = 0
        for i in range(len(proj)):
            best, best_level, best2, best_level2, best_idx = 256, -1, 256, -1, -1
            for p in range(off[i], off[i + 1]):
                j, d = int(idx[p]), int(dist[p])
                if fm[j] >= 0 and has_obs[fm[j]]:
                    continue
                if d < best:
                    best2, best, best_level2, best_level, best_idx = best, d, best_level, int(kF["octave"][j]), j
                elif d < best2:
                    best_level2, best2 = int(kF["octave"][j]), d
            if best <= self.TH_HIGH and best_idx >= 0:
                if best_level == best_level2 and np.float32(best) > ratio * np.float32(best2):
                    continue
                fm[best_idx] = i
                nmatches += 1
        return nmatches, fm

    def _search_by_projection_full(self, keysF, descF, scale_factors, bounds, proj, level, view_cos, descMP, has_obs, th,
                                   occupied, u_right, proj_r, keys_right, in_view, in_view_r, level_r, view_cos_r, l2r, r2l):
        """The whole ORBmatcher::SearchByProjection(Frame&, vpMapPoints, th) (ORBmatcher.cc:39-189) as
        Tracking::SearchLocalPoints meets it.  keysF = mvKeysUn (mvKeys of the left camera on a stereo-fisheye rig),
        keys_right = mvKeysRight (then descF holds the left rows followed by the right rows, Nleft = len(keysF)).  Every
        DescriptorDistance of both cameras' windows runs in ONE launch; the scan, whose skips depend on what earlier map
        points were assigned, is replayed in the reference's order."""
        f32 = np.float32
        kL = np.ascontiguousarray(keysF, KP_DTYPE)
        kR = None if keys_right is None else np.ascontiguousarray(keys_right, KP_DTYPE)
        nL, nR = len(kL), 0 if kR is None else len(kR)
        fisheye = nR > 0
        sf = np.asarray(scale_factors, f32)
        proj = np.asarray(proj, f32).reshape(-1, 2)
        nMP = len(proj)
        level = np.asarray(level, np.int32)
        inv = np.ones(nMP, bool) if in_view is None else np.asarray(in_view, bool)
        invr = np.zeros(nMP, bool) if (in_view_r is None or not fisheye) else np.asarray(in_view_r, bool)
        r = np.where(np.asarray(view_cos, f32) > 0.998, f32(2.5), f32(4.0)).astype(f32)
        if th != 1.0:
            r = (r * f32(th)).astype(f32)
        rad = (r * sf[level]).astype(f32)
        qL = np.flatnonzero(inv)
        offL, idxL = FrameGrid(kL, bounds).candidate_lists(proj[qL], rad[qL], level[qL] - 1, level[qL])
        if fisheye:
            pr = np.asarray(proj_r, f32).reshape(-1, 2)
            lr = np.asarray(level_r, np.int32)
            rr = np.where(np.asarray(view_cos_r, f32) > 0.998, f32(2.5), f32(4.0)).astype(f32)     # no th factor (:129)
            qR = np.flatnonzero(invr & (lr != -1))
            radR = (rr[qR] * sf[lr[qR]]).astype(f32)
            offR, idxR = FrameGrid(kR, bounds).candidate_lists(pr[qR], radR, lr[qR] - 1, lr[qR])
        else:
            pr = None if proj_r is None else np.asarray(proj_r, f32).reshape(-1, 2)
            qR, offR, idxR = np.zeros(0, np.int64), np.zeros(1, np.int32), np.zeros(0, np.int32)
        # one launch: the left windows, then the right windows (descriptor rows nL + i)
        dMP = np.ascontiguousarray(descMP, np.uint8).reshape(-1, 32)
        off = np.concatenate([offL, offL[-1] + offR[1:]]).astype(np.int32)
        idx = np.concatenate([idxL, idxR + nL]).astype(np.int32)
        dist = self.candidates(np.concatenate([dMP[qL], dMP[qR]]), descF, off, idx)
        slotL = np.full(nMP, -1, np.int64); slotL[qL] = np.arange(len(qL))
        slotR = np.full(nMP, -1, np.int64); slotR[qR] = len(qL) + np.arange(len(qR))
        has_obs = np.asarray(has_obs, bool)
        ur = None if (u_right is None or fisheye) else np.asarray(u_right, f32)
        st = np.full(nL + nR, -1, np.int64)                   # -2: held a point with observations on entry
        if occupied is not None:
            st[np.asarray(occupied, bool)] = -2
        octL = kL["octave"]
        octR = None if kR is None else kR["octave"]
        ratio = f32(self.mfNNratio)
        nmatches = 0

        def taken(j):
            return st[j] == -2 or (st[j] >= 0 and has_obs[st[j]])

        for i in range(nMP):
            if not inv[i] and not invr[i]:
                continue
            skip_right = False
            if inv[i]:
                q = slotL[i]
                best, best_level, best2, best_level2, best_idx = 256, -1, 256, -1, -1
                for p in range(off[q], off[q + 1]):
                    j, d = int(idx[p]), int(dist[p])
                    if taken(j):
                        continue
                    if ur is not None and ur[j] > 0:
                        if abs(f32(pr[i, 0] - ur[j])) > rad[i]:
                            continue
                    if d < best:
                        best2, best, best_level2, best_level, best_idx = best, d, best_level, int(octL[j]), j
                    elif d < best2:
                        best_level2, best2 = int(octL[j]), d
                if best <= self.TH_HIGH and best_idx >= 0:
                    if best_level == best_level2 and f32(best) > ratio * f32(best2):
                        skip_right = True                      # the reference's `continue` leaves the whole iteration
                    else:
                        st[best_idx] = i
                        if fisheye and l2r is not None and l2r[best_idx] != -1:
                            st[int(l2r[best_idx]) + nL] = i
                            nmatches += 1
                        nmatches += 1
            if skip_right or not fisheye or not invr[i] or slotR[i] < 0:
                continue
            q = slotR[i]
            best, best_level, best2, best_level2, best_idx = 256, -1, 256, -1, -1
            for p in range(off[q], off[q + 1]):
                j, d = int(idx[p]), int(dist[p])              # j = nL + index among the right key points
                if taken(j):
                    continue
                if d < best:
                    best2, best, best_level2, best_level, best_idx = best, d, best_level, int(octR[j - nL]), j - nL
                elif d < best2:
                    best_level2, best2 = int(octR[j - nL]), d
            if best <= self.TH_HIGH and best_idx >= 0:
                if best_level == best_level2 and f32(best) > ratio * f32(best2):
                    continue
                if r2l is not None and r2l[best_idx] != -1:
                    st[int(r2l[best_idx])] = i
                    nmatches += 1
                st[best_idx + nL] = i
                nmatches += 1
        return nmatches, np.where(st >= 0, st, -1).astype(np.int32)

    # ---- ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) (ORBmatcher.cc:1498-1684), Nleft == -1 ----
    def SearchByProjectionLastFrame(self, keysC, descC, scale_factors, bounds, valid, uv, invz, octave, angle_last, descMP,
                                    mp_has_obs, th=15.0, u_right=None, occupied=None, mbf=0.0, forward=False, backward=False):
        """The matcher of Tracking::TrackWithMotionModel on flattened frames.  Last frame, per feature i: valid[i] = has a
        map point and is not an outlier, uv[i] / invz[i] = projection of that point into the current frame and 1 / depth
        (the caller's pose and camera model, :1522-1534), octave[i] / angle_last[i] of its key point, descMP[i] =
        GetDescriptor(), mp_has_obs[i] = Observations() > 0.  Current frame: mvKeysUn, mDescriptors, grid bounds, mvuRight
        (None: mono), occupied[j] = mvpMapPoints[j] already holds a point with observations.  forward / backward =
        bForward / bBackward (:1513-1514).  Returns (nmatches, cur_match[j] = last-frame feature or -1).
        All DescriptorDistance calls of all windows run in one launch; the acceptance is replayed in the reference's order
        (it depends on earlier assignments, :1563-1565)."""
        f32 = np.float32
        kC = np.ascontiguousarray(keysC, KP_DTYPE)
        sf = np.asarray(scale_factors, f32)
        uv = np.asarray(uv, f32).reshape(-1, 2)
        invz, octave = np.asarray(invz, f32), np.asarray(octave, np.int32)
        ok = np.asarray(valid, bool) & ~(invz < 0) & ~(uv[:, 0] < f32(bounds[0])) & ~(uv[:, 0] > f32(bounds[2])) \
            & ~(uv[:, 1] < f32(bounds[1])) & ~(uv[:, 1] > f32(bounds[3]))
        q = np.flatnonzero(ok)
        radius = (f32(th) * sf[octave[q]]).astype(f32)
        if forward:
            lo, hi = octave[q], np.full(len(q), -1, np.int32)
        elif backward:
            lo, hi = np.zeros(len(q), np.int32), octave[q]
        else:
            lo, hi = octave[q] - 1, octave[q] + 1
        off, idx = FrameGrid(kC, bounds).candidate_lists(uv[q], radius, lo, hi)
        dist = self.candidates(np.ascontiguousarray(descMP, np.uint8).reshape(-1, 32)[q], descC, off, idx)
        occ = np.zeros(len(kC), bool) if occupied is None else np.asarray(occupied, bool)
        ur = None if u_right is None else np.asarray(u_right, f32)
        has_obs = np.asarray(mp_has_obs, bool)
        cm = np.full(len(kC), -1, np.int32)
        rot_hist = [[] for _ in range(self.HISTO_LENGTH)]
        nmatches = 0
        for qi, i in enumerate(q):
            best, best_idx = 256, -1
            for p in range(off[qi], off[qi + 1]):
                j = int(idx[p])
                if occ[j] or (cm[j] >= 0 and has_obs[cm[j]]):
                    continue
                if ur is not None and ur[j] > 0:
                    if abs(f32(uv[i, 0] - f32(f32(mbf) * invz[i])) - ur[j]) > radius[qi]:
                        continue
                d = int(dist[p])
                if d < best:
                    best, best_idx = d, j
            if best <= self.TH_HIGH:
                cm[best_idx] = i
                nmatches += 1
                if self.mbCheckOrientation:
                    rot_hist[self._rot_bin(angle_last[i], kC["angle"][best_idx])].append(best_idx)
        if self.mbCheckOrientation:
            keep = _three_maxima(rot_hist)
            for b in range(self.HISTO_LENGTH):
                if b not in keep:
                    for j in rot_hist[b]:                       # (:1672-1675: no "still set" test, duplicates count twice)
                        cm[j] = -1
                        nmatches -= 1
        return nmatches, cm

    # ---- the same with a stereo-fisheye CURRENT frame (CurrentFrame.Nleft != -1, ORBmatcher.cc:1602-1656 in addition) ----
    def SearchByProjectionLastFrameFisheye(self, keysC, keys_right, descC, scale_factors, bounds, valid, uv, uv_r, invz, octave,
                                           angle_last, descMP, mp_has_obs, th=15.0, occupied=None, forward=False,
                                           backward=False):
        """keysC = CurrentFrame.mvKeys (left camera, Nleft of them), keys_right = mvKeysRight, descC = the left rows followed
        by the right rows; uv_r[i] = projection of last-frame point i into the right camera (GetRelativePoseTrl() * x3Dc
        through the caller's camera model).  Every last-frame point whose LEFT window is not empty is also searched, best-1,
        in the right camera; both searches vote in one rotation histogram.  occupied / cur_match: Nleft + len(keys_right)."""
        f32 = np.float32
        kC, kR = np.ascontiguousarray(keysC, KP_DTYPE), np.ascontiguousarray(keys_right, KP_DTYPE)
        nC, nR = len(kC), len(kR)
        sf = np.asarray(scale_factors, f32)
        uv, uvr = np.asarray(uv, f32).reshape(-1, 2), np.asarray(uv_r, f32).reshape(-1, 2)
        invz, octave = np.asarray(invz, f32), np.asarray(octave, np.int32)
        ok = np.asarray(valid, bool) & ~(invz < 0) & ~(uv[:, 0] < f32(bounds[0])) & ~(uv[:, 0] > f32(bounds[2])) \
            & ~(uv[:, 1] < f32(bounds[1])) & ~(uv[:, 1] > f32(bounds[3]))
        q = np.flatnonzero(ok)
        radius = (f32(th) * sf[octave[q]]).astype(f32)
        if forward:
            lo, hi = octave[q], np.full(len(q), -1, np.int32)
        elif backward:
            lo, hi = np.zeros(len(q), np.int32), octave[q]
        else:
            lo, hi = octave[q] - 1, octave[q] + 1
        offL, idxL = FrameGrid(kC, bounds).candidate_lists(uv[q], radius, lo, hi)
        offR, idxR = FrameGrid(kR, bounds).candidate_lists(uvr[q], radius, lo, hi)
        dMP = np.ascontiguousarray(descMP, np.uint8).reshape(-1, 32)
        off = np.concatenate([offL, offL[-1] + offR[1:]]).astype(np.int32)          # left windows, then right windows
        idx = np.concatenate([idxL, idxR + nC]).astype(np.int32)
        dist = self.candidates(np.concatenate([dMP[q], dMP[q]]), descC, off, idx)
        occ = np.zeros(nC + nR, bool) if occupied is None else np.asarray(occupied, bool)
        has_obs = np.asarray(mp_has_obs, bool)
        cm = np.full(nC + nR, -1, np.int32)
        rot_hist = [[] for _ in range(self.HISTO_LENGTH)]
        nmatches = 0
        nq = len(q)
        for qi, i in enumerate(q):
            if off[qi] == off[qi + 1]:
                continue                                        # empty left window: the right camera is not searched either
            for side, ql, ang in ((0, qi, kC["angle"]), (1, nq + qi, kR["angle"])):
                best, best_idx = 256, -1
                for p in range(off[ql], off[ql + 1]):
                    j = int(idx[p])
                    if occ[j] or (cm[j] >= 0 and has_obs[cm[j]]):
                        continue
                    d = int(dist[p])
                    if d < best:
                        best, best_idx = d, j
                if best <= self.TH_HIGH:
                    cm[best_idx] = i
                    nmatches += 1
                    if self.mbCheckOrientation:
                        rot_hist[self._rot_bin(angle_last[i], ang[best_idx - side * nC])].append(best_idx)
        if self.mbCheckOrientation:
            keep = _three_maxima(rot_hist)
            for b in range(self.HISTO_LENGTH):
                if b not in keep:
                    for j in rot_hist[b]:
                        cm[j] = -1
                        nmatches -= 1
        return nmatches, cm

    # ---- ORBmatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) (ORBmatcher.cc:1685-1794) ----
    def SearchByProjectionKeyFrame(self, keysC, descC, scale_factors, bounds, valid, uv, dist3d, min_dist, max_dist, level,
                                   angle_kf, descMP, th=10.0, orb_dist=100, occupied=None):
        """The matcher of Tracking::Relocalization on flattened inputs.  Key frame, per feature i: valid[i] = map point
        present, not bad, not in sAlreadyFound; uv[i] = its projection; dist3d[i] = |x3Dw - Ow| with the invariance window
        [min_dist, max_dist]; level[i] = PredictScale; angle_kf[i]; descMP[i].  A current-frame feature holding ANY map point
        (occupied[j] or assigned earlier in this call) is skipped.  Returns (nmatches, cur_match[j] = key-frame feature)."""
        f32 = np.float32
        kC = np.ascontiguousarray(keysC, KP_DTYPE)
        sf = np.asarray(scale_factors, f32)
        uv = np.asarray(uv, f32).reshape(-1, 2)
        d3, lo_d, hi_d = np.asarray(dist3d, f32), np.asarray(min_dist, f32), np.asarray(max_dist, f32)
        level = np.asarray(level, np.int32)
        ok = np.asarray(valid, bool) & ~(uv[:, 0] < f32(bounds[0])) & ~(uv[:, 0] > f32(bounds[2])) & ~(uv[:, 1] < f32(bounds[1])) \
            & ~(uv[:, 1] > f32(bounds[3])) & ~(d3 < lo_d) & ~(d3 > hi_d)
        q = np.flatnonzero(ok)
        radius = (f32(th) * sf[level[q]]).astype(f32)
        off, idx = FrameGrid(kC, bounds).candidate_lists(uv[q], radius, level[q] - 1, level[q] + 1)
        dist = self.candidates(np.ascontiguousarray(descMP, np.uint8).reshape(-1, 32)[q], descC, off, idx)
        occ = np.zeros(len(kC), bool) if occupied is None else np.asarray(occupied, bool)
        cm = np.full(len(kC), -1, np.int32)
        rot_hist = [[] for _ in range(self.HISTO_LENGTH)]
        nmatches = 0
        for qi, i in enumerate(q):
            best, best_idx = 256, -1
            for p in range(off[qi], off[qi + 1]):
                j = int(idx[p])
                if occ[j] or cm[j] >= 0:
                    continue
                d = int(dist[p])
                if d < best:
                    best, best_idx = d, j
            if best <= orb_dist:
                cm[best_idx] = i
                nmatches += 1
                if self.mbCheckOrientation:
                    rot_hist[self._rot_bin(angle_kf[i], kC["angle"][best_idx])].append(best_idx)
        if self.mbCheckOrientation:
            keep = _three_maxima(rot_hist)
            for b in range(self.HISTO_LENGTH):
                if b not in keep:
                    for j in rot_hist[b]:
                        cm[j] = -1
                        nmatches -= 1
        return nmatches, cm

    # ---- ORBmatcher::Fuse(pKF, vpMapPoints, th, bRight = false) (ORBmatcher.cc:1015-1181): the matching core ----
    def FuseSearch(self, keysK, descK, scale_factors, inv_level_sigma2, bounds, u_right, valid, uv, ur, dist3d, min_dist,
                   max_dist, level, descMP, th=3.0, th_dist=None):
        """The matcher of LocalMapping::SearchInNeighbors up to the fuse decision (:1147).  Per map point i: valid[i] =
        present, not bad, not already in pKF, depth >= 0, viewing-angle test passed; uv[i] = projection into pKF, ur[i] =
        uv.x - bf * invz; dist3d[i] against [min_dist, max_dist]; level[i] = PredictScale.  Key frame: mvKeysUn, mDescriptors,
        mvuRight (< 0: mono feature), mvInvLevelSigma2, grid bounds.  Returns (nFused, best_idx[i] = key-frame feature or -1,
        best_dist[i]); the caller then runs its own Replace / AddObservation loop (:1148-1160) -- map edits that do not feed
        back into the search.  The level window and the chi-square gates (5.99 mono / 7.8 stereo) filter the candidate
        lists on the host; all descriptor distances and the per-list best run in one launch."""
        f32 = np.float32
        kK = np.ascontiguousarray(keysK, KP_DTYPE)
        sf, inv_s2 = np.asarray(scale_factors, f32), np.asarray(inv_level_sigma2, f32)
        uv = np.asarray(uv, f32).reshape(-1, 2)
        ur, d3 = np.asarray(ur, f32), np.asarray(dist3d, f32)
        level = np.asarray(level, np.int32)
        u_right = np.asarray(u_right, f32)
        n = len(uv)
        ok = np.asarray(valid, bool) & (uv[:, 0] >= f32(bounds[0])) & (uv[:, 0] < f32(bounds[2])) & (uv[:, 1] >= f32(bounds[1])) \
            & (uv[:, 1] < f32(bounds[3])) & ~(d3 < np.asarray(min_dist, f32)) & ~(d3 > np.asarray(max_dist, f32))
        q = np.flatnonzero(ok)
        off, idx = FrameGrid(kK, bounds).candidate_lists(uv[q], (f32(th) * sf[level[q]]).astype(f32))
        # gates of :1107-1141 on every (map point, candidate) pair, vectorised
        owner = np.repeat(np.arange(len(q)), np.diff(off))
        i = q[owner]
        kp = kK[idx]
        lv = kp["octave"]
        ex, ey = uv[i, 0] - kp["x"], uv[i, 1] - kp["y"]
        er = ur[i] - u_right[idx]
        stereo = u_right[idx] >= 0
        e2 = np.where(stereo, ((ex * ex + ey * ey).astype(f32) + er * er).astype(f32), (ex * ex + ey * ey).astype(f32))
        gate = (e2 * inv_s2[lv]).astype(f32) > np.where(stereo, 7.8, 5.99)          # float * float -> float, compared with a double
        keep = (lv >= level[i] - 1) & (lv <= level[i]) & ~gate
        off2 = np.zeros(len(q) + 1, np.int32)
        off2[1:] = np.cumsum(np.bincount(owner[keep], minlength=len(q)))
        _, (i1, d1, _, _) = self.candidates(np.ascontiguousarray(descMP, np.uint8).reshape(-1, 32)[q], descK, off2, idx[keep], top2=True)
        best_idx, best_dist = np.full(n, -1, np.int32), np.full(n, 256, np.int32)
        best_dist[q] = d1
        fused = d1 <= (self.TH_LOW if th_dist is None else th_dist)
        best_idx[q[fused]] = i1[fused]
        return int(fused.sum()), best_idx, best_dist

    def FuseSearchSim3(self, keysK, descK, scale_factors, bounds, valid, uv, dist3d, min_dist, max_dist, level, descMP, th=3.0,
                       th_dist=None):
        """Matching core of ORBmatcher::Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) (ORBmatcher.cc:1182-1292, LoopClosing):
        FuseSearch without the reprojection gates.  valid[i] = not bad, not already a map point of pKF, depth >= 0,
        viewing-angle test passed.  The caller fills vpReplacePoint / AddObservation from best_idx (:1268-1280)."""
        n, k = len(np.asarray(uv).reshape(-1, 2)), len(keysK)
        return self.FuseSearch(keysK, descK, scale_factors, np.zeros(len(scale_factors), np.float32), bounds,
                               np.full(k, -1.0, np.float32), valid, uv, np.zeros(n, np.float32), dist3d, min_dist, max_dist, level,
                               descMP, th, th_dist)

    # ---- ORBmatcher::SearchBySim3(pKF1, pKF2, vpMatches12, S12, th) (ORBmatcher.cc:1293-1497) ----
    def SearchBySim3(self, keys1, desc1, keys2, desc2, scale_factors, bounds, valid1, uv12, dist12, min1, max1, level12, valid2,
                     uv21, dist21, min2, max2, level21, th=7.5):
        """LoopClosing's Sim3-guided matcher.  valid1[i1] = feature i1 of pKF1 has a map point that is not bad and not matched
        yet (vbAlreadyMatched1), depth >= 0 in camera 2; uv12 / dist12 / level12 = its projection into image 2, |p3Dc2| against
        [min1, max1], PredictScale; the *2 / *21 arguments likewise for pKF2's points in image 1 (valid2 excludes
        vbAlreadyMatched2).  Both key frames share scale factors and image bounds.  Each direction is the gate-free fuse search
        with TH_HIGH; a pair is kept when both directions agree (:1483-1494).  Returns (nFound, match12)."""
        _, vn1, _ = self.FuseSearchSim3(keys2, desc2, scale_factors, bounds, valid1, uv12, dist12, min1, max1, level12, desc1, th,
                                        self.TH_HIGH)
        _, vn2, _ = self.FuseSearchSim3(keys1, desc1, scale_factors, bounds, valid2, uv21, dist21, min2, max2, level21, desc2, th,
                                        self.TH_HIGH)
        i1 = np.arange(len(vn1))
        ok = (vn1 >= 0) & (vn2[np.maximum(vn1, 0)] == i1)
        return int(ok.sum()), np.where(ok, vn1, -1).astype(np.int32)

    # ---- ORBmatcher::SearchByProjection(pKF, Scw, vpPoints, [vpPointsKFs,] vpMatched, [vpMatchedKF,] th, ratioHamming) ----
    def SearchByProjectionSim3(self, keysK, descK, scale_factors, bounds, occupied, valid, uv, dist3d, min_dist, max_dist, level,
                               descMP, th=3, ratio_hamming=1.0):
        """The matchers of LoopClosing (ORBmatcher.cc:372-471 and :473-580; same search).  Per candidate map point i: valid[i] =
        not bad, not in vpMatched on entry, depth >= 0, viewing-angle test passed; uv[i] = projection through Scw; dist3d[i]
        against [min_dist, max_dist]; level[i] = PredictScale.  occupied[j] = vpMatched[j] != NULL on entry.  Returns
        (nmatches, kf_match[j] = candidate stored in vpMatched[j] by this call, or -1); for the second overload
        vpMatchedKF[j] = vpPointsKFs[kf_match[j]]."""
        f32 = np.float32
        kK = np.ascontiguousarray(keysK, KP_DTYPE)
        sf = np.asarray(scale_factors, f32)
        uv = np.asarray(uv, f32).reshape(-1, 2)
        d3 = np.asarray(dist3d, f32)
        level = np.asarray(level, np.int32)
        ok = np.asarray(valid, bool) & (uv[:, 0] >= f32(bounds[0])) & (uv[:, 0] < f32(bounds[2])) & (uv[:, 1] >= f32(bounds[1])) \
            & (uv[:, 1] < f32(bounds[3])) & ~(d3 < np.asarray(min_dist, f32)) & ~(d3 > np.asarray(max_dist, f32))
        q = np.flatnonzero(ok)
        off, idx = FrameGrid(kK, bounds).candidate_lists(uv[q], (f32(int(th)) * sf[level[q]]).astype(f32))
        dist = self.candidates(np.ascontiguousarray(descMP, np.uint8).reshape(-1, 32)[q], descK, off, idx)
        occ = np.asarray(occupied, bool)
        km = np.full(len(kK), -1, np.int32)
        limit = f32(f32(self.TH_LOW) * f32(ratio_hamming))
        nmatches = 0
        for qi, i in enumerate(q):
            best, best_idx = 256, -1
            for p in range(off[qi], off[qi + 1]):
                j = int(idx[p])
                if occ[j] or km[j] >= 0:
                    continue
                lv = int(kK["octave"][j])
                if lv < level[i] - 1 or lv > level[i]:
                    continue
                d = int(dist[p])
                if d < best:
                    best, best_idx = d, j
            if f32(best) <= limit:
                km[best_idx] = i
                nmatches += 1
        return nmatches, km

    # ---- descriptor-based key-point association of matched key-frame pairs (submap merge) ----
    def AssociateSubmap(self, extractor, images1, keys1, valid1, images2, keys2, valid2, th=None):
        """For every matched key-frame pair p (the pairs CloudMerging.cc:503-551 walks): real descriptors for the cloud key
        frames (CloudFrameComputeDescriptors, ORBextractor.cc:989-1011, one batched call per side), all pairs' top-2 in one
        launch, SearchByBoW's acceptance (best <= TH_LOW and best < ratio * second).  Only key points with a map point
        (valid*) take part, like the `vpMap1MapPoints[i] && vpMap2MapPoints[j]` test of the reference.
        images*: [npairs, h, w] u8; keys*: list of KP_DTYPE arrays; valid*: list of bool arrays.
        Returns (match12 list: key point of key frame 2 or -1 per key point of key frame 1, number of matches per pair)."""
        def describe(images, keys):
            off = np.zeros(len(keys) + 1, np.int32)
            off[1:] = np.cumsum([len(k) for k in keys])
            allk = np.concatenate([np.ascontiguousarray(k, KP_DTYPE) for k in keys]) if off[-1] else np.zeros(0, KP_DTYPE)
            d = extractor.CloudFrameComputeDescriptorsBatch(images, allk, off)
            return [d[off[i]:off[i + 1]] for i in range(len(keys))]
        d1, d2 = describe(images1, keys1), describe(images2, keys2)
        sel1 = [np.flatnonzero(v) for v in valid1]
        sel2 = [np.flatnonzero(v) for v in valid2]
        res = self.top2_pairs([d[s] for d, s in zip(d1, sel1)], [d[s] for d, s in zip(d2, sel2)])
        out, counts = [], []
        for p, (i1, e1, e2) in enumerate(res):
            ok = self.accept_bow(e1, e2, th) & (i1 >= 0)
            m = np.full(len(keys1[p]), -1, np.int32)
            m[sel1[p][ok]] = sel2[p][i1[ok]]
            out.append(m)
            counts.append(int(ok.sum()))
        return out, counts

    # ---- stereo row-band best-1 (Frame.cc:828-905) ----
    def stereo_best1(self, Lk, Ld, Rk, Rd, scale_factors, n_rows, min_d, max_d):
        Lk = np.ascontiguousarray(Lk, KP_DTYPE)
        Rk = np.ascontiguousarray(Rk, KP_DTYPE)
        Ld = np.ascontiguousarray(Ld, np.uint8)
        Rd = np.ascontiguousarray(Rd, np.uint8)
        sf = np.ascontiguousarray(scale_factors, np.float32)
        best = np.zeros(len(Lk), np.int32)
        dist = np.zeros(len(Lk), np.uint16)
        check(self._L.rumi_stereo_best1(self._m, ptr(Lk), ptr(Ld), len(Lk), ptr(Rk), ptr(Rd), len(Rk), ptr(sf),
                                        len(sf), int(n_rows), float(min_d), float(max_d), ptr(best), ptr(dist)))
        return best, dist

    # ---- Frame::ComputeStereoMatches, complete (Frame.cc:828-985) ----
    def stereo_match(self, ex_left, ex_right, Lk, Ld, Rk, Rd, mbf, mb):
        """ex_left / ex_right: the ORBextractor objects whose LAST call produced (Lk, Ld) / (Rk, Rd); their device
        pyramids are used in place.  Returns (mvuRight, mvDepth, number of matches kept)."""
        Lk = np.ascontiguousarray(Lk, KP_DTYPE)
        Rk = np.ascontiguousarray(Rk, KP_DTYPE)
        Ld = np.ascontiguousarray(Ld, np.uint8)
        Rd = np.ascontiguousarray(Rd, np.uint8)
        u = np.zeros(len(Lk), np.float32)
        d = np.zeros(len(Lk), np.float32)
        n = C.c_int32(0)
        check(self._L.rumi_stereo_match(self._m, ex_left._h, ex_right._h, ptr(Lk), ptr(Ld), len(Lk), ptr(Rk), ptr(Rd),
                                        len(Rk), float(mbf), float(mb), ptr(u), ptr(d), C.byref(n)))
        return u, d, n.value

    # ---- MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:355-426), batched over map points ----
    def ComputeDistinctiveDescriptors(self, desc, offsets):
        """desc: [total,32] observed descriptors of all map points, point p owns rows offsets[p]:offsets[p+1].
        Returns (best_idx, best_median): per point the row (relative to its own list) with the least median distance
        to the point's other observations -- the descriptor the reference stores in mDescriptor."""
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        offsets = np.ascontiguousarray(offsets, np.int32)
        n = len(offsets) - 1
        best, med = np.zeros(max(n, 0), np.int32), np.zeros(max(n, 0), np.int32)
        if n > 0:
            check(self._L.rumi_distinctive_descriptors(self._m, ptr(desc), ptr(offsets), n, ptr(best), ptr(med)))
        return best, med

    # ---- ORBmatcher::SearchByBoW: shared machinery ----
    def _bow_blocks(self, desc_a, featvec_a, desc_b, featvec_b):
        """Distance blocks of every vocabulary node common to both feature vectors (ascending node id == the nodes
        the reference's lower_bound walk visits): [(a_indices, b_indices, dist[len(a), len(b)])]."""
        da = np.ascontiguousarray(desc_a, np.uint8).reshape(-1, 32)
        db = np.ascontiguousarray(desc_b, np.uint8).reshape(-1, 32)
        common = sorted(set(featvec_a) & set(featvec_b))
        a_idx, b_idx, segs, off = [], [], [], 0
        for nid in common:
            ia, ib = featvec_a[nid], featvec_b[nid]
            segs.append((len(a_idx), len(ia), len(b_idx), len(ib), off))
            a_idx += list(ia); b_idx += list(ib)
            off += len(ia) * len(ib)
        if not segs or off == 0:
            return []
        a_idx, b_idx = np.array(a_idx, np.int32), np.array(b_idx, np.int32)
        segs_a = np.array(segs, np.int32).reshape(-1, 5)
        dist = np.zeros(off, np.uint16)
        check(self._L.rumi_bow_node_distances(self._m, ptr(da), len(da), ptr(db), len(db), ptr(a_idx), len(a_idx),
                                              ptr(b_idx), len(b_idx), ptr(segs_a), len(segs_a), ptr(dist), off))
        return [(a_idx[a0:a0 + ac], b_idx[b0:b0 + bc], dist[o:o + ac * bc].reshape(ac, bc)) for (a0, ac, b0, bc, o) in segs]

    def _rot_bin(self, angle_a, angle_b):
        rot = np.float32(angle_a) - np.float32(angle_b)
        if rot < 0.0:
            rot = np.float32(rot + np.float32(360.0))
        b = _c_round(float(np.float32(rot * (np.float32(1.0) / np.float32(self.HISTO_LENGTH)))))
        return 0 if b == self.HISTO_LENGTH else b

    # ---- ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches) (ORBmatcher.cc:198-370, F.Nleft == -1) ----
    def SearchByBoW(self, desc_kf, angle_kf, kf_valid, featvec_kf, desc_f, angle_f, featvec_f, n_left=-1):
        """desc_*: [n,32] descriptors; angle_*: keypoint angles (mvKeysUn[i].angle / mvKeys[i].angle); kf_valid[i]:
        the keyframe feature has a map point that is not bad (:227-233); featvec_*: FeatureVector {node: [indices]}.
        n_left = F.Nleft: -1 for a mono / rectified frame; for a stereo-fisheye frame features [0, n_left) are the left and
        [n_left, N) the right camera (:258-340: a best / second best per side; the right match needs no ratio test but is
        only taken when the left one passed TH_LOW).
        Returns (nmatches, match_f) with match_f[j] = keyframe feature whose map point was assigned to frame feature
        j, or -1.  The GPU computes every distance of every common vocabulary node; the acceptance (TH_LOW, ratio,
        the "already matched" skip :249 and the rotation histogram) is replayed here in the reference's order."""
        nf = len(np.asarray(desc_f).reshape(-1, 32))
        match_f = np.full(nf, -1, np.int32)
        nmatches = 0
        rot_hist = [[] for _ in range(self.HISTO_LENGTH)]
        ratio = np.float32(self.mfNNratio)

        def accept(real_kf, idx_f):
            match_f[idx_f] = real_kf
            if self.mbCheckOrientation:
                rot_hist[self._rot_bin(angle_kf[real_kf], angle_f[idx_f])].append(idx_f)
            return 1
        for a_idx, b_idx, block in self._bow_blocks(desc_kf, featvec_kf, desc_f, featvec_f):
            for i in range(len(a_idx)):
                real_kf = int(a_idx[i])
                if not kf_valid[real_kf]:
                    continue
                best1, best_f, best2 = 256, -1, 256
                best1r, best_fr, best2r = 256, -1, 256
                for j in range(len(b_idx)):
                    real_f = int(b_idx[j])
                    if match_f[real_f] >= 0:
                        continue
                    d = int(block[i, j])
                    if n_left == -1 or real_f < n_left:
                        if d < best1:
                            best2, best1, best_f = best1, d, real_f
                        elif d < best2:
                            best2 = d
                    else:
                        if d < best1r:
                            best2r, best1r, best_fr = best1r, d, real_f
                        elif d < best2r:
                            best2r = d
                if best1 <= self.TH_LOW:
                    if np.float32(best1) < ratio * np.float32(best2):
                        nmatches += accept(real_kf, best_f)
                    if best1r <= self.TH_LOW:                   # ":316 ... || true": no ratio test on the right side
                        nmatches += accept(real_kf, best_fr)
        if self.mbCheckOrientation:
            keep = _three_maxima(rot_hist)
            for i in range(self.HISTO_LENGTH):
                if i in keep:
                    continue
                for j in rot_hist[i]:
                    match_f[j] = -1
                    nmatches -= 1
        return nmatches, match_f

    # ---- Frame::ComputeStereoFishEyeMatches (Frame.cc:1120-1161): the matching core ----
    def StereoFishEyeMatches(self, desc_left, mono_left, desc_right, mono_right, ratio=0.7):
        """Brute-force k = 2 match of the lapping-area descriptors (rows mono_left.. of the left, mono_right.. of the right
        image: what operator() with vLappingArea put at the back, :1122-1126) + Lowe's ratio d0 < d1 * 0.7 (:1146).  Returns
        [(left feature, right feature)] in the reference's order, indices in the full arrays; the caller triangulates
        each pair with its camera model and keeps those with positive depth (:1151-1158)."""
        dl = np.ascontiguousarray(desc_left, np.uint8).reshape(-1, 32)[mono_left:]
        dr = np.ascontiguousarray(desc_right, np.uint8).reshape(-1, 32)[mono_right:]
        if len(dl) == 0 or len(dr) < 2:                       # knnMatch returns fewer than 2 neighbours: size() >= 2 fails
            return []
        i1, d1, d2 = self.top2(dl, dr)
        ok = self.accept_knn_ratio(d1, d2, ratio) & (i1 >= 0)
        return [(int(q) + mono_left, int(i1[q]) + mono_right) for q in np.flatnonzero(ok)]

    # ---- ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) (ORBmatcher.cc:682-804, NLeft == -1) ----
    def SearchByBoW_KF(self, desc1, angle1, valid1, featvec1, desc2, angle2, valid2, featvec2):
        """Returns (nmatches, match12): match12[i] = feature of keyframe 2 whose map point is stored in vpMatches12[i],
        or -1.  valid1 / valid2: the feature has a map point that is not bad."""
        n1 = len(np.asarray(desc1).reshape(-1, 32))
        n2 = len(np.asarray(desc2).reshape(-1, 32))
        match12 = np.full(n1, -1, np.int32)
        matched2 = np.zeros(max(n2, 1), bool)
        nmatches = 0
        rot_hist = [[] for _ in range(self.HISTO_LENGTH)]
        ratio = np.float32(self.mfNNratio)
        for a_idx, b_idx, block in self._bow_blocks(desc1, featvec1, desc2, featvec2):
            for i in range(len(a_idx)):
                id1 = int(a_idx[i])
                if not valid1[id1]:
                    continue
                best1, best_2, best2 = 256, -1, 256
                for j in range(len(b_idx)):
                    id2 = int(b_idx[j])
                    if matched2[id2] or not valid2[id2]:
                        continue
                    d = int(block[i, j])
                    if d < best1:
                        best2, best1, best_2 = best1, d, id2
                    elif d < best2:
                        best2 = d
                if best1 < self.TH_LOW and np.float32(best1) < ratio * np.float32(best2):      # strict '<' here (:756)
                    match12[id1] = best_2
                    matched2[best_2] = True
                    if self.mbCheckOrientation:
                        rot_hist[self._rot_bin(angle1[id1], angle2[best_2])].append(id1)
                    nmatches += 1
        if self.mbCheckOrientation:
            keep = _three_maxima(rot_hist)
            for i in range(self.HISTO_LENGTH):
                if i in keep:
                    continue
                for j in rot_hist[i]:
                    match12[j] = -1
                    nmatches -= 1
        return nmatches, match12

    # ---- ORBmatcher::SearchForTriangulation (ORBmatcher.cc:806-1013, key frames without a second camera) ----
    def SearchForTriangulation(self, desc1, angle1, has_mp1, stereo1, featvec1, desc2, angle2, has_mp2, stereo2, keys2, featvec2,
                               scale_factors2, epipole, epipolar_ok, only_stereo=False, coarse=False):
        """The matcher of LocalMapping::CreateNewMapPoints.  has_mp*[i]: the feature already has a map point; stereo*[i]:
        mvuRight[i] >= 0; keys2: mvKeysUn of key frame 2; epipole = projection of camera centre 1 into image 2 (:815-819);
        epipolar_ok(i1, i2) -> bool = pCamera1->epipolarConstrain(pCamera2, kp1, kp2, R12, t12, sigma1, sigma2) (:957), the
        caller's geometry, called only for pairs that survive the distance tests, in the reference's order (a [n1, n2] bool
        array works too).  Returns (nmatches, match12).  The GPU computes every distance of every common vocabulary node."""
        f32 = np.float32
        n1 = len(np.asarray(desc1).reshape(-1, 32))
        k2 = np.ascontiguousarray(keys2, KP_DTYPE)
        sf2 = np.asarray(scale_factors2, f32)
        epx, epy = f32(epipole[0]), f32(epipole[1])
        ok = epipolar_ok if callable(epipolar_ok) else (lambda a, b, _t=np.asarray(epipolar_ok): bool(_t[a, b]))
        match12 = np.full(n1, -1, np.int32)
        rot_hist = [[] for _ in range(self.HISTO_LENGTH)]
        nmatches = 0
        for a_idx, b_idx, block in self._bow_blocks(desc1, featvec1, desc2, featvec2):
            for i in range(len(a_idx)):
                id1 = int(a_idx[i])
                if has_mp1[id1] or (only_stereo and not stereo1[id1]):
                    continue
                best, best2 = self.TH_LOW, -1
                for j in range(len(b_idx)):
                    id2 = int(b_idx[j])
                    if has_mp2[id2] or (only_stereo and not stereo2[id2]):
                        continue
                    d = int(block[i, j])
                    if d > self.TH_LOW or d > best:
                        continue
                    if not stereo1[id1] and not stereo2[id2]:
                        ex, ey = f32(epx - k2["x"][id2]), f32(epy - k2["y"][id2])
                        if f32(f32(ex * ex) + f32(ey * ey)) < f32(f32(100.0) * sf2[k2["octave"][id2]]):
                            continue
                    if coarse or ok(id1, id2):
                        best2, best = id2, d
                if best2 >= 0:
                    match12[id1] = best2
                    nmatches += 1
                    if self.mbCheckOrientation:
                        rot_hist[self._rot_bin(angle1[id1], angle2[best2])].append(id1)
        if self.mbCheckOrientation:
            keep = _three_maxima(rot_hist)
            for b in range(self.HISTO_LENGTH):
                if b not in keep:
                    for id1 in rot_hist[b]:
                        match12[id1] = -1
                        nmatches -= 1
        return nmatches, match12


class FrameGrid:
    """The 64 x 48 key-point grid a reference Frame / KeyFrame carries (FRAME_GRID_COLS / ROWS,
    R/include/cloud_edge_slam_lib/Frame.h:42-43): AssignFeaturesToGrid + PosInGrid (R/lib_src/Frame.cc:441-466, 752-767)
    and GetFeaturesInArea (:695-750; KeyFrame.cc:887-925 is the same walk).  Host-side index bookkeeping, as in the
    reference; its output (candidate lists in the reference's order) feeds rumi_hamming_candidates."""
    COLS, ROWS = 64, 48

    def __init__(self, kps, bounds):
        """kps: KP_DTYPE key points (mvKeysUn); bounds = (mnMinX, mnMinY, mnMaxX, mnMaxY)."""
        f32 = np.float32
        self.kps = np.ascontiguousarray(kps, KP_DTYPE)
        self.min_x, self.min_y = f32(bounds[0]), f32(bounds[1])
        self.w_inv = f32(self.COLS) / f32(bounds[2] - bounds[0])
        self.h_inv = f32(self.ROWS) / f32(bounds[3] - bounds[1])
        px = _c_round_arr((self.kps["x"] - self.min_x) * self.w_inv)
        py = _c_round_arr((self.kps["y"] - self.min_y) * self.h_inv)
        ok = (px >= 0) & (px < self.COLS) & (py >= 0) & (py < self.ROWS)
        cell = np.where(ok, px * self.ROWS + py, self.COLS * self.ROWS)
        self._order = np.argsort(cell, kind="stable").astype(np.int32)           # insertion order inside a cell
        self._start = np.searchsorted(cell[self._order], np.arange(self.COLS * self.ROWS + 1)).astype(np.int64)

    def features_in_area(self, x, y, r, min_level=-1, max_level=-1):
        f32 = np.float32
        x, y, r = f32(x), f32(y), f32(r)
        c0 = max(0, int(np.floor((x - self.min_x - r) * self.w_inv)))
        if c0 >= self.COLS:
            return np.zeros(0, np.int32)
        c1 = min(self.COLS - 1, int(np.ceil((x - self.min_x + r) * self.w_inv)))
        if c1 < 0:
            return np.zeros(0, np.int32)
        r0 = max(0, int(np.floor((y - self.min_y - r) * self.h_inv)))
        if r0 >= self.ROWS:
            return np.zeros(0, np.int32)
        r1 = min(self.ROWS - 1, int(np.ceil((y - self.min_y + r) * self.h_inv)))
        if r1 < 0:
            return np.zeros(0, np.int32)
        parts = [self._order[self._start[ix * self.ROWS + r0]:self._start[ix * self.ROWS + r1 + 1]] for ix in range(c0, c1 + 1)]
        idx = np.concatenate(parts) if parts else np.zeros(0, np.int32)
        if len(idx) == 0:
            return idx
        k = self.kps[idx]
        keep = (np.abs(k["x"] - x) < r) & (np.abs(k["y"] - y) < r)
        if min_level > 0 or max_level >= 0:                                      # bCheckLevels
            keep &= k["octave"] >= min_level
            if max_level >= 0:
                keep &= k["octave"] <= max_level
        return idx[keep]

    def candidate_lists(self, qxy, qr, qmin=None, qmax=None):
        """CSR lists (off[nq + 1], idx) of features_in_area for many queries."""
        qxy = np.asarray(qxy, np.float32).reshape(-1, 2)
        nq = len(qxy)
        qr = np.broadcast_to(np.asarray(qr, np.float32), (nq,))
        qmin = np.broadcast_to(np.asarray(-1 if qmin is None else qmin, np.int32), (nq,))
        qmax = np.broadcast_to(np.asarray(-1 if qmax is None else qmax, np.int32), (nq,))
        lists = [self.features_in_area(qxy[q, 0], qxy[q, 1], qr[q], int(qmin[q]), int(qmax[q])) for q in range(nq)]
        off = np.zeros(nq + 1, np.int32)
        off[1:] = np.cumsum([len(l) for l in lists])
        return off, (np.concatenate(lists).astype(np.int32) if nq and off[-1] else np.zeros(0, np.int32))


def _c_round_arr(v):
    """C round() of a float32 array: half away from zero (numpy rounds half to even)."""
    v = np.asarray(v, np.float32)
    return np.where(v >= 0, np.floor(v + np.float32(0.5)), -np.floor(-v + np.float32(0.5))).astype(np.int64)


def _c_round(x):
    """C round(): half away from zero (numpy rounds half to even)."""
    import math
    return int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5))


def _three_maxima(histo):
    """ORBmatcher::ComputeThreeMaxima (ORBmatcher.cc:1795-1828): indices of the three fullest bins (-1 dropped)."""
    max1 = max2 = max3 = 0
    ind1 = ind2 = ind3 = -1
    for i, h in enumerate(histo):
        s = len(h)
        if s > max1:
            max3, max2, max1 = max2, max1, s
            ind3, ind2, ind1 = ind2, ind1, i
        elif s > max2:
            max3, max2 = max2, s
            ind3, ind2 = ind2, i
        elif s > max3:
            max3, ind3 = s, i
    if max2 < np.float32(0.1) * np.float32(max1):
        ind2 = ind3 = -1
    elif max3 < np.float32(0.1) * np.float32(max1):
        ind3 = -1
    return {ind1, ind2, ind3}
