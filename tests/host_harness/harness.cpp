// TEST INFRASTRUCTURE: compiles the product's host/device (RUMI_HD) arithmetic for the HOST so that the kernel
// logic can be checked against the oracle on a machine without a GPU (pytest -m "not gpu").
// It is never loaded by the product path; the shipped library exports no host compute.
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>

#include "../../rumi_slam_b200/csrc/orb_geom.h"
#include "../../rumi_slam_b200/csrc/orb_math.cuh"
#include "../../rumi_slam_b200/csrc/octree_core.cuh"

using namespace rumi;

struct HostCtx {
    int tid = 0, nthr = 1;
    void sync() {}
    void sync_warp() {}
    void sort_u64(uint64_t* a, int n) { std::sort(a, a + n); }
    void sort_tree_keys(uint64_t* a, int m, int npad, int bshift, int nb) {
        for (int i = 0; i < m; ++i)
            if ((long long)(a[i] >> bshift) >= nb) std::abort();          // every key must fall into a bucket
        std::sort(a, a + npad);
    }
    int warp_id() const { return 0; }
    int num_warps() const { return 1; }
    int lane() const { return 0; }
    int warp_size() const { return 1; }
    int atomic_add(int* p, int v) { int o = *p; *p += v; return o; }
    int exclusive_scan(int v, int*, int* total) { *total = v; return 0; }
    void mark(int) {}
    void atomic_min(int* p, int v) { if (v < *p) *p = v; }
};

extern "C" {

int hh_geom(int W, int H, int nfeatures, float scale, int nlevels, int ini, int mn, OrbConst* out) {
    return build_orb_const(*out, W, H, nfeatures, scale, nlevels, ini, mn);
}
int hh_sizeof_const() { return (int)sizeof(OrbConst); }

// level geometry as flat ints: w,h,stride,nCols,nRows,wCell,hCell,quota,nIni,treeDepth,candCap,kpBase
int hh_level_info(int W, int H, int nfeatures, float scale, int nlevels, int ini, int mn, int* info, float* finfo) {
    OrbConst oc;
    int rc = build_orb_const(oc, W, H, nfeatures, scale, nlevels, ini, mn);
    if (rc) return rc;
    for (int l = 0; l < nlevels; ++l) {
        const LevelGeom& g = oc.lv[l];
        int* p = info + 12 * l;
        p[0] = g.w; p[1] = g.h; p[2] = g.stride; p[3] = g.nCols; p[4] = g.nRows; p[5] = g.wCell; p[6] = g.hCell;
        p[7] = g.quota; p[8] = g.nIni; p[9] = g.treeDepth; p[10] = g.candCap; p[11] = g.kpBase;
        finfo[3 * l] = g.hX; finfo[3 * l + 1] = g.scale; finfo[3 * l + 2] = g.patchSize;
    }
    for (int i = 0; i < 16; ++i) info[12 * nlevels + i] = oc.umax[i];
    info[12 * nlevels + 16] = oc.kpCap; info[12 * nlevels + 17] = oc.totalCells;
    return 0;
}

void hh_resize_coef(int sn, int dn, uint16_t* ofs, int16_t* a0, int16_t* a1) {
    AxisCoef c = make_axis_coef(sn, dn);
    std::memcpy(ofs, c.ofs.data(), dn * 2); std::memcpy(a0, c.a0.data(), dn * 2); std::memcpy(a1, c.a1.data(), dn * 2);
}

int hh_fast_score(const int* d16) { return fast_score16(d16); }
float hh_atan2(float y, float x) { return fast_atan2_deg(y, x); }
void hh_sincos(float a, float* s, float* c) { glibc_sincosf(a, s, c); }

// std::sort replay check: sorts (key32 << 32 | idx) with the replay and returns the permutation.
void hh_stdsort(uint64_t* v, int n) { stdsort::sort(v, n); }
void hh_realsort(uint64_t* v, int n) {
    std::sort(v, v + n, [](const uint64_t& a, const uint64_t& b) { return (uint32_t)(a >> 32) < (uint32_t)(b >> 32); });
}

// serial std::sort replays of the last hh_octree call | (sorted final phase ran) << 16
static int g_last_replays = 0;
int hh_octree_last_replays() { return g_last_replays; }

// Quad-tree distribution of packed candidates for level `level` of the given configuration.
int hh_octree(int W, int H, int nfeatures, float scale, int nlevels, int level, const uint32_t* cand, int M,
              int N_override, uint32_t* out, int outCap) {
    OrbConst oc;
    int rc = build_orb_const(oc, W, H, nfeatures, scale, nlevels, 20, 7);
    if (rc) return rc;
    const LevelGeom& g = oc.lv[level];
    const int N = N_override >= 0 ? N_override : g.quota;
    OctreeWork w;
    const int nodeCap = std::max(N, 4 * g.nIni) + 4;
    w.nodeCap = nodeCap; w.createCap = 3 * nodeCap + 16; w.pendCap = 2 * nodeCap + 16;
    std::vector<uint64_t> keys(next_pow2(std::max(M, 8))), lkeys(next_pow2(std::max(nodeCap, 8))), pend(w.pendCap), lsort(w.pendCap);
    std::vector<uint32_t> glo(nodeCap + 1), crlo(w.createCap), crcnt(w.createCap), next(w.pendCap), next2(w.pendCap), meta(w.pendCap), qbase(w.pendCap);
    std::vector<int> hist(2 * (kMaxTreeDepth + 2)), part(2), scal(SC_COUNT);
    w.keys = keys.data(); w.lkeys = lkeys.data(); w.glo = glo.data(); w.cr_lo = crlo.data(); w.cr_cnt = crcnt.data();
    w.lsort = lsort.data();
    w.pend = pend.data(); w.next = next.data(); w.next2 = next2.data(); w.meta = meta.data(); w.qbase = qbase.data(); w.hist = hist.data(); w.part = part.data(); w.scal = scal.data();
    HostCtx ctx;
    distribute_quadtree(ctx, cand, M, N, g, w, out, outCap);
    g_last_replays = scal[SC_NREPLAY] | (scal[SC_PHASEB] << 16);
    return scal[SC_NOUT];
}

}  // extern "C"
