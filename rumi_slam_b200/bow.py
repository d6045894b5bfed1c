"""Host-side mirror of the bag-of-words interface the reference uses around its matchers:
ORBVocabulary = DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB> (R/include/cloud_edge_slam_lib/ORBVocabulary.h,
R/Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h) and Frame::ComputeBoW (R/lib_src/Frame.cc), over the C ABI.

The tree descent of every feature (the Hamming work) runs on the GPU; assembling the BowVector / FeatureVector maps
from the per-feature (word, weight, node) triples is the reference's own map bookkeeping and stays on the host.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, ptr

# BowVector.h enums
TF_IDF, TF, IDF, BINARY = 0, 1, 2, 3
L1_NORM, L2_NORM, CHI_SQUARE, KL, BHATTACHARYYA, DOT_PRODUCT = 0, 1, 2, 3, 4, 5


class ORBVocabulary:
    def __init__(self, k, L, parent, is_leaf, desc, weight, scoring=L1_NORM, weighting=TF_IDF, device=0):
        """Nodes in id order as loadFromTextFile reads them (node 0 = root): parent id, leaf flag, 32-byte
        descriptor, weight."""
        self._L = _lib.lib()
        self._v = C.c_void_p()
        self.m_k, self.m_L, self.m_scoring, self.m_weighting = int(k), int(L), int(scoring), int(weighting)
        parent = np.ascontiguousarray(parent, np.int32)
        is_leaf = np.ascontiguousarray(is_leaf, np.uint8)
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        weight = np.ascontiguousarray(weight, np.float64)
        assert len(parent) == len(is_leaf) == len(desc) == len(weight)
        check(self._L.rumi_vocab_create(C.byref(self._v), int(device), self.m_k, self.m_L, len(parent), ptr(parent),
                                        ptr(is_leaf), ptr(desc), ptr(weight)))

    @classmethod
    def loadFromTextFile(cls, filename, device=0):
        """TemplatedVocabulary::loadFromTextFile (:1338-1421): 'k L scoring weighting' then one node per line."""
        with open(filename) as f:
            k, L, n1, n2 = (int(x) for x in f.readline().split())
            if k < 0 or k > 20 or L < 1 or L > 10 or n1 < 0 or n1 > 5 or n2 < 0 or n2 > 3:
                raise ValueError("Vocabulary loading failure: This is not a correct text file!")
            parent, leaf, desc, weight = [0], [0], [np.zeros(32, np.uint8)], [0.0]
            for line in f:
                t = line.split()
                if not t:
                    continue
                parent.append(int(t[0])); leaf.append(1 if int(t[1]) > 0 else 0)
                desc.append(np.array([int(x) for x in t[2:34]], np.uint8)); weight.append(float(t[34]))
        return cls(k, L, parent, leaf, np.stack(desc), weight, n1, n2, device)

    def close(self):
        if getattr(self, "_v", None) is not None and self._v:
            self._L.rumi_vocab_destroy(self._v)
            self._v = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def size(self):
        return int(self._L.rumi_vocab_words(self._v))

    def launch_count(self, reset=False):
        return int(self._L.rumi_vocab_launch_count(self._v, 1 if reset else 0))

    # ---- per-feature descent (TemplatedVocabulary::transform(feature, id, weight, nid, levelsup), :1218-1258) ----
    def transform_features(self, desc, levelsup=0):
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        word, weight, node = np.zeros(n, np.int32), np.zeros(n, np.float64), np.zeros(n, np.int32)
        check(self._L.rumi_bow_transform(self._v, ptr(desc), n, int(levelsup), ptr(word), ptr(weight), ptr(node)))
        return word, weight, node

    def transform_features_device(self, desc, levelsup=0, sync=True):
        import torch
        assert desc.is_cuda and desc.dtype == torch.uint8 and desc.is_contiguous()
        n = desc.shape[0]
        word = torch.empty(n, dtype=torch.int32, device=desc.device)
        weight = torch.empty(n, dtype=torch.float64, device=desc.device)
        node = torch.empty(n, dtype=torch.int32, device=desc.device)
        st = _lib.torch_stream()                    # `desc` was produced on torch's stream; the results are used there
        check(self._L.rumi_vocab_wait_stream(self._v, st))
        check(self._L.rumi_bow_transform_device(self._v, ptr(desc), n, int(levelsup), ptr(word), ptr(weight), ptr(node),
                                                1 if sync else 0))
        if not sync:
            check(self._L.rumi_vocab_signal_stream(self._v, st))
        return word, weight, node

    # ---- transform(features, BowVector&, FeatureVector&, levelsup) (:1128-1200) == Frame::ComputeBoW ----
    def transform(self, desc, levelsup=4):
        """Returns (BowVector {word id: value}, FeatureVector {node id: [feature indices]})."""
        word, weight, node = self.transform_features(desc, levelsup)
        must = self.m_scoring != DOT_PRODUCT                     # ScoringObject.h:73-89
        l2 = self.m_scoring == L2_NORM
        bow, fv = {}, {}
        tf = self.m_weighting in (TF, TF_IDF)
        for i in range(len(word)):
            w = float(weight[i])
            if w > 0:                                            # not stopped
                wid = int(word[i])
                if tf:
                    bow[wid] = bow.get(wid, 0.0) + w             # addWeight
                elif wid not in bow:
                    bow[wid] = w                                 # addIfNotExist
                fv.setdefault(int(node[i]), []).append(i)        # addFeature
        ids = sorted(bow)                                        # std::map iteration order = ascending id
        if tf and bow and not must:
            nd = float(len(bow))
            for wid in ids:
                bow[wid] /= nd
        if must:                                                 # BowVector::normalize (BowVector.cpp:62-84)
            norm = 0.0
            if not l2:
                for wid in ids:
                    norm += abs(bow[wid])
            else:
                for wid in ids:
                    norm += bow[wid] * bow[wid]
                norm = float(np.sqrt(norm))
            if norm > 0.0:
                for wid in ids:
                    bow[wid] /= norm
        return {wid: bow[wid] for wid in ids}, {nid: fv[nid] for nid in sorted(fv)}
