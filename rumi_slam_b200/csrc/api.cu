// C ABI of librumi_orb.so (declared in include/rumi_orb.h): handle management, workspace layout in HBM, stream
// orchestration.  No arithmetic of the hot path happens on the host; without a CUDA device every call fails.
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rumi_orb.h"
#include "kernels.cuh"
#include "orb_geom.h"

using namespace rumi;

static_assert(sizeof(rumi_kp) == 28 && sizeof(KeyPointRec) == 28, "cv::KeyPoint layout");

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU_TRY(expr)                                                                                     \
    do {                                                                                                 \
        cudaError_t e__ = (expr);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return fail(RUMI_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    // looked up once; a function-local static initialiser is thread safe (two extractors may be created concurrently)
    static const EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            return (EncodeTiledFn)p;
        return nullptr;
    }();
    return fn;
}

// 3-D u8 tensor map (x, y, frame) with box (boxW, boxH, 1).  Returns false when TMA's alignment rules are not met.
bool make_tmap(CUtensorMap* m, const uint8_t* base, int w, int h, int nframes, int stride, long long pitch, int boxW,
               int boxH) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return false;
    if (((uintptr_t)base & 15) || (stride & 15) || (pitch & 15) || boxW > 256 || boxH > 256 || (boxW & 15)) return false;
    cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)nframes};
    cuuint64_t strides[2] = {(cuuint64_t)stride, (cuuint64_t)pitch};
    cuuint32_t box[3] = {(cuuint32_t)boxW, (cuuint32_t)boxH, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct Workspace {
    cudaStream_t stream = nullptr;
    uint8_t *pyr = nullptr, *blur = nullptr;
    uint32_t *cand = nullptr, *candOrdered = nullptr, *sel = nullptr;
    int *countBase = nullptr;           // [chunk] frameDone counters, [4] bigCount (+ pad), then levelCount: ONE memset clears all
    int *bigList = nullptr;             // [chunk * nlevels] work list of the quad-tree's second pass
    int *levelCount = nullptr, *cellOff = nullptr, *cellCount = nullptr, *selCount = nullptr, *slot = nullptr;
    uint64_t* bigKeys = nullptr;
    KeyPointRec* kps = nullptr;
    uint8_t* desc = nullptr;
    int *nkp = nullptr, *nmono = nullptr;
    CUtensorMap tmap[kMaxLevels];      // tmap[l]: source map (internal level l-1) used to produce level l
    bool tmapOk[kMaxLevels];
    int lastFrames = 0;
    bool ready = false;                // every buffer above is allocated
    // pinned staging of the PAGEABLE host batch path (images in, results out) + "results landed" event
    uint8_t *hIn = nullptr, *hOut = nullptr;
    size_t hInCap = 0, hOutCap = 0;
    cudaEvent_t evOut = nullptr;
};

}  // namespace

constexpr int kMaxWs = 4;

struct rumi_orb {
    int device = 0;
    int nfeatures = 0, nlevels = 0, iniTh = 0, minTh = 0, chunk = 1;
    float scaleFactor = 1.2f;
    ScaleTables tables;
    bool useTMA = true;
    // geometry-dependent state
    int W = 0, H = 0;
    OrbConst oc;
    ResizeCoef* coef = nullptr;                       // all levels, x then y
    const ResizeCoef *xc[kMaxLevels], *yc[kMaxLevels];
    int boxW[kMaxLevels], boxH[kMaxLevels];
    long long pyrLevelOff[kMaxLevels], candLevelOff[kMaxLevels], bigKeysLevelOff[kMaxLevels];
    int bigKeysCap[kMaxLevels];
    long long pyrBytes = 0, candElems = 0, bigKeysElems = 0;
    int fastTilePitch = 0, fastTileRows = 0, fastScoreRows = 0;
    FastCell* fastCells = nullptr;                    // [totalCells] cell geometry of all levels
    // strip pyramid (K1): per-level column-group / row tables, and for each usable strip count the rows every strip
    // computes per level + its shared-memory budget
    bool useStrip = true, stripOk = false;
    uint8_t* pyrTables = nullptr;
    PyrStripLevel stripLv[kMaxLevels];
    struct StripVariant { int nstrips; const int2* ranges; size_t smemBytes; int buf1Offset, rowTabOffset; int rowsPerItem[kMaxLevels]; };
    std::vector<StripVariant> stripVariants;          // ascending nstrips
    int stripForce = 0;                               // RUMI_PYR_STRIPS: force a strip count (tuning)
    int smemKeys = 4096, smemKeysBig = 0, maxNodeCap = 0;
    Workspace ws[kMaxWs];          // chunk c runs on ws[c % nws]: copies and kernels of consecutive chunks overlap
    int nws = 4;
    int skipMask = 0;              // RUMI_SKIP_STAGES (timing experiments only: results are wrong) bit s = skip stage s
    bool nwsSet = false;           // RUMI_STREAMS given: use it for every path
    int nwsDefault = 4; bool nwsSetDefault = false;
    int lastWs = 0;
    // device-side error flags in mapped pinned host memory (per handle, read for free at every host sync point):
    // [0] a TMA transaction of the tile pyramid timed out, [1] a pyramid level dependency timed out
    int* errHost = nullptr; int* errDev = nullptr;
    uint8_t* descBuf = nullptr; size_t descCap = 0;   // scratch of rumi_orb_describe* (images, keypoints, descriptors)
    bool pendingSingle = false;                            // rumi_orb_extract_begin issued, _end not yet called
    // mvImagePyramid of single-frame calls: levels downloaded behind the kernels into a pinned block (rumi_orb_set_pyramid_staging)
    bool stagePyr = false, pyrStageValid = false;
    uint8_t* pyrStage = nullptr; size_t pyrStageCap = 0;
    size_t pyrStageOff[kMaxLevels] = {};
    cudaStream_t auxStream = nullptr;                      // small calls: blur runs here, beside FAST / quad-tree / slots
    cudaEvent_t evFork = nullptr, evJoinAux = nullptr;
    uint8_t* outStage = nullptr; size_t outStageCap = 0;   // pinned: results of the single-frame call land here with ONE sync
    uint8_t* dbgBuf = nullptr;     // test hook: FAST tile dump
    long long* octClk = nullptr;   // profiling hook: octree phase cycle counters
    int dbgCell = 0;
    // measurement: device-side timer on the launching streams, optional per-stage events, launch counter
    cudaEvent_t evStart = nullptr, evStop = nullptr, evJoin[kMaxWs] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t evOrder = nullptr, evSignal[kMaxWs] = {nullptr, nullptr, nullptr, nullptr};   // wait_stream / signal_stream
    cudaEvent_t evCompute[kMaxWs] = {nullptr, nullptr, nullptr, nullptr};                     // host batch path: kernels of a chunk done
    bool profile = false;
    std::vector<cudaEvent_t> evPool;
    size_t evUsed = 0;
    std::vector<std::pair<int, int>> evSpans;     // (stage, index of the first of two consecutive events)
    double stageMs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long stageLaunches[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long launches = 0;
};

namespace {

void free_workspace(Workspace& w) {
    cudaFree(w.pyr); cudaFree(w.blur); cudaFree(w.cand); cudaFree(w.candOrdered); cudaFree(w.sel);
    cudaFree(w.countBase); cudaFree(w.bigList); cudaFree(w.cellOff); cudaFree(w.cellCount); cudaFree(w.selCount); cudaFree(w.slot);
    cudaFree(w.bigKeys); cudaFree(w.kps); cudaFree(w.desc); cudaFree(w.nkp); cudaFree(w.nmono);
    if (w.hIn) cudaFreeHost(w.hIn);
    if (w.hOut) cudaFreeHost(w.hOut);
    if (w.evOut) cudaEventDestroy(w.evOut);
    cudaStream_t s = w.stream;
    w = Workspace();
    w.stream = s;
}

LevelView internal_view(const rumi_orb* h, const uint8_t* base, int l) {
    LevelView v;
    const LevelGeom& g = h->oc.lv[l];
    v.ptr = base + h->pyrLevelOff[l];
    v.stride = g.stride;
    v.pitch = (long long)g.stride * g.h;
    v.w = g.w; v.h = g.h;
    return v;
}

int alloc_workspace(rumi_orb* h, Workspace& w) {
    const OrbConst& oc = h->oc;
    const long long n = h->chunk;
    CU_TRY(cudaMalloc(&w.pyr, h->pyrBytes));
    CU_TRY(cudaMalloc(&w.blur, h->pyrBytes));
    CU_TRY(cudaMalloc(&w.cand, 4 * h->candElems));
    CU_TRY(cudaMalloc(&w.candOrdered, 4 * h->candElems));
    CU_TRY(cudaMalloc(&w.bigKeys, 8 * h->bigKeysElems));
    CU_TRY(cudaMalloc(&w.countBase, 4 * (n * (oc.nlevels + 1) + 4)));
    w.levelCount = w.countBase + n + 4;
    CU_TRY(cudaMalloc(&w.bigList, 4 * n * oc.nlevels));
    CU_TRY(cudaMalloc(&w.selCount, 4 * n * oc.nlevels));
    CU_TRY(cudaMalloc(&w.cellOff, 4 * n * oc.totalCells));
    CU_TRY(cudaMalloc(&w.cellCount, 4 * n * oc.totalCells));
    CU_TRY(cudaMalloc(&w.sel, 4 * n * oc.kpCap));
    CU_TRY(cudaMalloc(&w.slot, 4 * n * oc.kpCap));
    CU_TRY(cudaMalloc(&w.kps, sizeof(KeyPointRec) * n * oc.kpCap));
    CU_TRY(cudaMalloc(&w.desc, 32 * n * oc.kpCap));
    CU_TRY(cudaMalloc(&w.nkp, 4 * n));
    CU_TRY(cudaMalloc(&w.nmono, 4 * n));
    // the workspace streams are non-blocking: a legacy-stream memset would not be ordered before the first upload
    CU_TRY(cudaMemsetAsync(w.pyr, 0, h->pyrBytes, w.stream));
    CU_TRY(cudaMemsetAsync(w.blur, 0, h->pyrBytes, w.stream));
    for (int l = 1; l < oc.nlevels; ++l) {
        const LevelView src = internal_view(h, w.pyr, l - 1);
        w.tmapOk[l] = h->useTMA && make_tmap(&w.tmap[l], src.ptr, src.w, src.h, h->chunk, src.stride, src.pitch,
                                             h->boxW[l], h->boxH[l]);
    }
    w.ready = true;
    return RUMI_OK;
}

// Drops everything that depends on the image shape: the handle is back to "no geometry" (W = H = 0), so the next call
// rebuilds from scratch.  Also the failure path of build_geometry / alloc_workspace: a half-built geometry is never kept.
void reset_geometry(rumi_orb* h) {
    for (int i = 0; i < kMaxWs; ++i) {
        if (h->ws[i].stream) cudaStreamSynchronize(h->ws[i].stream);
        free_workspace(h->ws[i]);
    }
    cudaFree(h->coef);
    cudaFree(h->fastCells);
    cudaFree(h->pyrTables);
    h->coef = nullptr;
    h->fastCells = nullptr;
    h->pyrTables = nullptr;
    h->stripOk = false; h->stripVariants.clear();
    h->W = 0; h->H = 0;
    h->pyrBytes = 0; h->candElems = 0; h->bigKeysElems = 0;
    cudaGetLastError();                     // a failed cudaMalloc leaves a (non-sticky) error behind
}

// Builds everything that depends on the image shape (the caller resets the handle when this fails).
int build_geometry(rumi_orb* h, int W, int H) {
    OrbConst oc;
    const int rc = build_orb_const(oc, W, H, h->nfeatures, h->scaleFactor, h->nlevels, h->iniTh, h->minTh);
    if (rc) return fail(RUMI_ERR_SHAPE, "image %dx%d cannot be processed with %d levels (code %d)", W, H, h->nlevels, rc);
    for (int l = 0; l < oc.nlevels; ++l)
        if (oc.lv[l].candCap >= (1 << kOrderBits))
            return fail(RUMI_ERR_SHAPE, "level %d too large (%d candidates)", l, oc.lv[l].candCap);
    reset_geometry(h);
    h->oc = oc;
    h->W = W; h->H = H;
    // resize coefficient tables + per-level source box of a 64x32 tile
    std::vector<ResizeCoef> all;
    std::vector<size_t> xo(oc.nlevels, 0), yo(oc.nlevels, 0);
    for (int l = 1; l < oc.nlevels; ++l) {
        const LevelGeom &s = oc.lv[l - 1], &d = oc.lv[l];
        AxisCoef cx = make_axis_coef(s.w, d.w), cy = make_axis_coef(s.h, d.h);
        xo[l] = all.size();
        for (int i = 0; i < d.w; ++i) all.push_back(ResizeCoef{cx.ofs[i], cx.a0[i], cx.a1[i], 0});
        yo[l] = all.size();
        for (int i = 0; i < d.h; ++i) all.push_back(ResizeCoef{cy.ofs[i], cy.a0[i], cy.a1[i], 0});
        int bw = 0, bh = 0;
        for (int ox = 0; ox < d.w; ox += kPyrTileW) {
            const int last = std::min(ox + kPyrTileW, d.w) - 1;
            bw = std::max(bw, std::min(cx.ofs[last] + 1, s.w - 1) - (cx.ofs[ox] & ~15) + 1);   // box starts 16-aligned
        }
        for (int oy = 0; oy < d.h; oy += kPyrTileH) {
            const int last = std::min(oy + kPyrTileH, d.h) - 1;
            bh = std::max(bh, std::min(cy.ofs[last] + 1, s.h - 1) - cy.ofs[oy] + 1);
        }
        h->boxW[l] = align_up(bw, 16);
        h->boxH[l] = bh;
    }
    if (!all.empty()) {
        CU_TRY(cudaMalloc(&h->coef, all.size() * sizeof(ResizeCoef)));
        CU_TRY(cudaMemcpy(h->coef, all.data(), all.size() * sizeof(ResizeCoef), cudaMemcpyHostToDevice));
        CU_TRY(cudaDeviceSynchronize());     // pageable H2D may still be in flight; the workspace streams do not wait for it
    } else {
        CU_TRY(cudaMalloc(&h->coef, sizeof(ResizeCoef)));
    }
    for (int l = 1; l < oc.nlevels; ++l) { h->xc[l] = h->coef + xo[l]; h->yc[l] = h->coef + yo[l]; }
    // strip pyramid tables: per group of 4 destination columns the 8-byte source window (first word, alignment
    // shift, byte selectors, weights), per destination row the two source rows and the vertical weights << 16
    {
        std::vector<uint8_t> blob;
        std::vector<size_t> colOff(oc.nlevels, 0), rowOff(oc.nlevels, 0);
        std::vector<std::vector<PyrRow>> rowTab(oc.nlevels);
        bool ok = oc.nlevels > 1;
        for (int l = 1; l < oc.nlevels && ok; ++l) {
            const LevelGeom &sg = oc.lv[l - 1], &dg = oc.lv[l];
            AxisCoef cx = make_axis_coef(sg.w, dg.w), cy = make_axis_coef(sg.h, dg.h);
            const int groups = (dg.w + 3) / 4;
            std::vector<PyrColGroup> cols(groups);
            for (int g = 0; g < groups; ++g) {
                PyrColGroup& c = cols[g];
                const int s0 = cx.ofs[std::min(4 * g, dg.w - 1)];
                c.word0 = (uint16_t)(s0 >> 2); c.shift = (uint8_t)(8 * (s0 & 3)); c.pad = 0;
                uint32_t sel = 0;
                for (int k = 0; k < 4; ++k) {
                    const int x = std::min(4 * g + k, dg.w - 1);
                    const int i0 = cx.ofs[x] - s0, i1 = std::min(cx.ofs[x] + 1, sg.w - 1) - s0;
                    if (i0 < 0 || i1 > 7) ok = false;                 // scale factor > 2: tile kernel handles it
                    sel |= (uint32_t)((i0 & 15) | ((i1 & 15) << 4)) << (8 * k);
                    c.coef[k] = (uint32_t)(uint16_t)cx.a0[x] | ((uint32_t)(uint16_t)cx.a1[x] << 16);
                }
                c.sel01 = (uint16_t)(sel & 0xFFFFu); c.sel23 = (uint16_t)(sel >> 16);
            }
            std::vector<PyrRow>& rows = rowTab[l];
            rows.resize(dg.h);
            for (int y = 0; y < dg.h; ++y) {
                int s0 = cy.ofs[y], s1 = std::min(cy.ofs[y] + 1, sg.h - 1);
                uint32_t b0 = (uint32_t)cy.a0[y], b1 = (uint32_t)cy.a1[y];
                if (s1 == s0) {                                  // clamped at the bottom (weight of the second tap is 0)
                    if (b1 != 0 || s0 < 1) ok = false;
                    s0 -= 1; b1 = b0; b0 = 0;
                }
                rows[y].sy1 = s1; rows[y].b0s = b0 << 16; rows[y].b1s = b1 << 16;
                rows[y].adv = 1;
                // the kernel keeps only the last two horizontally filtered rows and reads source rows in order:
                // consecutive rows advance by 0 (bottom clamp only: first-tap weight 0), 1 or 2 source rows
                const int adv = y > 0 ? s1 - rows[y - 1].sy1 : 1;
                if (adv < 0 || adv > 2 || s1 - s0 != 1 || (adv == 0 && b0 != 0)) ok = false;
                if (y > 0) rows[y - 1].adv = adv;                // the record of row y-1 carries the advance count of row y
            }
            blob.resize((blob.size() + 15) & ~(size_t)15);
            colOff[l] = blob.size();
            blob.insert(blob.end(), (uint8_t*)cols.data(), (uint8_t*)(cols.data() + groups));
            blob.resize((blob.size() + 15) & ~(size_t)15);
            rowOff[l] = blob.size();
            blob.insert(blob.end(), (uint8_t*)rows.data(), (uint8_t*)(rows.data() + dg.h));
            PyrStripLevel& m = h->stripLv[l];
            m.groups = groups;
            m.srcLastWord = (sg.w - 1) >> 2;
            m.rowsPerItem = 8;
        }
        // Strip variants: for S strips per frame, strip s owns rows [s*h/S, (s+1)*h/S) of every level and additionally
        // computes the rows of level l that its rows of level l+1 read (1-2 halo rows per level, computed by both
        // neighbours).  Usable when the two largest adjacent levels of a strip fit the shared memory of one SM.
        std::vector<size_t> rangeOff;
        std::vector<rumi_orb::StripVariant> variants;
        const int top = oc.nlevels - 1;
        const int candidates[] = {1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 48, 64};
        for (int S : candidates) {
            if (!ok) break;
            if (S > 1 && oc.lv[top].h / S < 4) break;
            std::vector<int2> rg((size_t)S * kMaxLevels, make_int2(0, 0));
            size_t bufBytes[2] = {0, 0};
            rumi_orb::StripVariant v{};
            v.nstrips = S;
            for (int st = 0; st < S; ++st) {
                int needA = 0, needB = 0;
                for (int l = top; l >= 1; --l) {
                    const int hl = oc.lv[l].h;
                    int a = (int)((long long)st * hl / S), b = (int)((long long)(st + 1) * hl / S);
                    if (l < top) { a = std::min(a, needA); b = std::max(b, needB); }
                    rg[(size_t)st * kMaxLevels + l] = make_int2(a, b);
                    if (b > a) { needA = rowTab[l][a].sy1 - 1; needB = rowTab[l][b - 1].sy1 + 1; }    // rows of level l-1 read
                    else { needA = oc.lv[l - 1].h; needB = 0; }
                    // shared copy of a level: 16 spare bytes behind every row, one spare row behind the last (read-ahead)
                    if (l < top) bufBytes[l & 1] = std::max(bufBytes[l & 1], (size_t)(b - a + 1) * (oc.lv[l].stride + 16) + 64);
                }
            }
            int maxRows = 0;
            for (int st = 0; st < S; ++st)
                for (int l = 1; l <= top; ++l)
                    maxRows = std::max(maxRows, rg[(size_t)st * kMaxLevels + l].y - rg[(size_t)st * kMaxLevels + l].x);
            v.buf1Offset = (int)((bufBytes[0] + 127) & ~(size_t)127);
            v.rowTabOffset = v.buf1Offset + (int)((bufBytes[1] + 127) & ~(size_t)127);
            v.smemBytes = (size_t)v.rowTabOffset + 16 * (size_t)maxRows;
            if (v.smemBytes > 220 * 1024) continue;
            // rows per item: minimise rounds x per-item cost (vertical pass ~25 instr / row, horizontal ~16 per source row)
            for (int l = 1; l <= top; ++l) {
                int rowsMax = 0;
                for (int st = 0; st < S; ++st) rowsMax = std::max(rowsMax, rg[(size_t)st * kMaxLevels + l].y - rg[(size_t)st * kMaxLevels + l].x);
                const double sc = (double)oc.lv[l - 1].h / oc.lv[l].h;
                double best = 1e30; int bestR = 8;
                for (int R = 3; R <= 24; ++R) {
                    const long long items = (long long)((rowsMax + R - 1) / R) * h->stripLv[l].groups;
                    const double rounds = (double)((items + kPyrStripThreads - 1) / kPyrStripThreads);
                    const double cost = rounds * (R * 25.0 + (sc * R + 1.0) * 16.0 + 30.0);
                    if (cost < best - 1e-9) { best = cost; bestR = R; }
                }
                v.rowsPerItem[l] = bestR;
            }
            blob.resize((blob.size() + 15) & ~(size_t)15);
            rangeOff.push_back(blob.size());
            blob.insert(blob.end(), (uint8_t*)rg.data(), (uint8_t*)(rg.data() + rg.size()));
            variants.push_back(v);
        }
        ok = ok && !variants.empty();
        h->stripOk = ok;
        if (ok) {
            CU_TRY(cudaMalloc(&h->pyrTables, blob.size()));
            CU_TRY(cudaMemcpy(h->pyrTables, blob.data(), blob.size(), cudaMemcpyHostToDevice));
            CU_TRY(cudaDeviceSynchronize());
            for (int l = 1; l < oc.nlevels; ++l) {
                h->stripLv[l].cols = reinterpret_cast<const PyrColGroup*>(h->pyrTables + colOff[l]);
                h->stripLv[l].rows = reinterpret_cast<const PyrRow*>(h->pyrTables + rowOff[l]);
            }
            size_t maxSmem = 0;
            for (size_t i = 0; i < variants.size(); ++i) {
                variants[i].ranges = reinterpret_cast<const int2*>(h->pyrTables + rangeOff[i]);
                maxSmem = std::max(maxSmem, variants[i].smemBytes);
            }
            h->stripVariants = variants;
            if (pyramid_strip_prepare(maxSmem) != 0) { cudaGetLastError(); h->stripOk = false; }
        }
    }
    // level-major workspace offsets for `chunk` frames
    long long pb = 0, ce = 0, be = 0;
    int tp = 0, tr = 0, sr = 0, nodeCap = 0;
    std::vector<FastCell> cells(oc.totalCells);
    for (int l = 0; l < oc.nlevels; ++l) {
        const LevelGeom& g = oc.lv[l];
        h->pyrLevelOff[l] = pb; pb += (long long)h->chunk * g.stride * g.h;
        pb = (pb + 255) & ~255ll;
        h->candLevelOff[l] = ce; ce += (long long)h->chunk * g.candCap;
        int cap = 2;
        while (cap < g.candCap) cap <<= 1;
        h->bigKeysCap[l] = cap;
        h->bigKeysLevelOff[l] = be; be += (long long)h->chunk * cap;
        tp = std::max(tp, align_up(g.wCell + 6 + 15, 16)); tr = std::max(tr, g.hCell + 6);   // 16-B aligned staging
        sr = std::max(sr, g.hCell + 2);
        // FAST grid cells (R/lib_src/ORBextractor.cc:748-763): sub-image origin / size clipped to the level border
        const int maxBX = g.w - kMinBorder, maxBY = g.h - kMinBorder;
        for (int ci = 0; ci < g.nRows; ++ci)
            for (int cj = 0; cj < g.nCols; ++cj) {
                FastCell& c = cells[g.cellBase + ci * g.nCols + cj];
                const int iniY = kMinBorder + ci * g.hCell, iniX = kMinBorder + cj * g.wCell;
                const int maxY = std::min(iniY + g.hCell + 6, maxBY), maxX = std::min(iniX + g.wCell + 6, maxBX);
                c.iniX = (uint16_t)iniX; c.iniY = (uint16_t)iniY; c.level = (uint8_t)l;
                c.valid = !(iniY >= maxBY - 3 || iniX >= maxBX - 6) && maxX - iniX > 6 && maxY - iniY > 6;
                c.cw = (uint8_t)(c.valid ? maxX - iniX : 0); c.ch = (uint8_t)(c.valid ? maxY - iniY : 0);
            }
        const int slots = (l + 1 < oc.nlevels ? oc.lv[l + 1].kpBase : oc.kpCap) - g.kpBase;
        nodeCap = std::max(nodeCap, slots + 1);
    }
    h->pyrBytes = pb + 256; h->candElems = ce; h->bigKeysElems = be;
    h->fastTilePitch = tp; h->fastTileRows = tr; h->fastScoreRows = sr;
    CU_TRY(cudaMalloc(&h->fastCells, sizeof(FastCell) * std::max<size_t>(cells.size(), 1)));
    CU_TRY(cudaMemcpy(h->fastCells, cells.data(), sizeof(FastCell) * cells.size(), cudaMemcpyHostToDevice));
    CU_TRY(cudaDeviceSynchronize());
    h->maxNodeCap = nodeCap;
    // 2048 keys in shared memory (16 KB): five problems per SM stay resident, which is what hides the latency of the
    // serial phases; a level with more candidates sorts in the global scratch instead
    h->smemKeys = getenv("RUMI_OCT_KEYS") ? atoi(getenv("RUMI_OCT_KEYS")) : 2048;
    while (h->smemKeys > 256 && octree_smem_bytes(h->smemKeys, nodeCap, 256) > 200 * 1024) h->smemKeys >>= 1;
    // second pass (levels denser than smemKeys): as many keys as one SM's shared memory takes, up to 16384
    h->smemKeysBig = 16384;
    while (h->smemKeysBig > h->smemKeys && octree_smem_bytes(h->smemKeysBig, nodeCap, 1024) > 200 * 1024) h->smemKeysBig >>= 1;
    if (octree_smem_bytes(h->smemKeys, nodeCap, 256) > 220 * 1024)
        return fail(RUMI_ERR_CAPACITY, "nfeatures %d needs more shared memory than one SM has", h->nfeatures);
    return RUMI_OK;
}

// (Re)builds the shape-dependent state when the shape changes.  Nothing is committed on failure: the handle is left
// without geometry, so a later call with the same shape tries again instead of running on null tables.
int ensure_geometry(rumi_orb* h, int W, int H) {
    if (h->W == W && h->H == H && h->coef && h->fastCells) return RUMI_OK;
    CU_TRY(cudaSetDevice(h->device));
    const int rc = build_geometry(h, W, H);
    if (rc) {
        const std::string msg = g_err;
        reset_geometry(h);
        g_err = msg;
    }
    return rc;
}

int ensure_workspace(rumi_orb* h, int idx) {
    Workspace& w = h->ws[idx];
    if (!w.stream) CU_TRY(cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking));
    if (w.ready) return RUMI_OK;
    const int rc = alloc_workspace(h, w);
    if (rc) {                                 // partial allocation: release it, the next call starts over
        const std::string msg = g_err;
        cudaStreamSynchronize(w.stream);
        free_workspace(w);
        cudaGetLastError();
        g_err = msg;
    }
    return rc;
}

// Called after a host synchronisation: a kernel that gave up waiting (instead of hanging the GPU) becomes an error code.
int check_device_flags(rumi_orb* h) {
    if (!h->errHost) return RUMI_OK;
    const int tma = h->errHost[0], dep = h->errHost[1];
    if (!tma && !dep) return RUMI_OK;
    h->errHost[0] = h->errHost[1] = 0;                       // reported once; the handle stays usable
    return fail(RUMI_ERR_CUDA, tma ? "TMA transaction timed out (results of this call are invalid)"
                                   : "pyramid level dependency timed out (results of this call are invalid)");
}

enum { ST_PYRAMID = 0, ST_FAST = 1, ST_OCTREE = 2, ST_SLOTS = 3, ST_BLUR = 4, ST_DESCRIBE = 5, ST_H2D = 6, ST_D2H = 7,
       ST_COUNT = 8 };

cudaEvent_t prof_event(rumi_orb* h, cudaStream_t s) {
    if (h->evUsed == h->evPool.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        h->evPool.push_back(e);
    }
    cudaEvent_t e = h->evPool[h->evUsed++];
    cudaEventRecord(e, s);
    return e;
}
// closes stage `stage` (which issued `nlaunch` kernels) with an event; spans are resolved in rumi_orb_profile_read
void prof_mark(rumi_orb* h, cudaStream_t s, int stage, int nlaunch) {
    h->launches += nlaunch;
    if (!h->profile) return;
    h->stageLaunches[stage] += nlaunch;
    h->evSpans.push_back(std::make_pair(stage, (int)h->evUsed - 1));
    prof_event(h, s);
}

// Enqueues the whole extraction of `n` frames whose level 0 is `l0` (internal buffer or caller device memory).
int run_chunk(rumi_orb* h, Workspace& w, const LevelView& l0, bool l0Internal, int n, int lap0, int lap1,
              KeyPointRec* dKps, uint8_t* dDesc, int outCap, int* dNkp, int* dNmono) {
    const OrbConst& oc = h->oc;
    cudaStream_t s = w.stream;
    cudaStream_t sh = s;
    ChunkView cv;
    cv.nframes = n;
    for (int l = 0; l < oc.nlevels; ++l) {
        cv.src[l] = l == 0 ? l0 : internal_view(h, w.pyr, l);
        cv.blur[l] = internal_view(h, w.blur, l);
    }
    if (!(h->skipMask & 2)) CU_TRY(cudaMemsetAsync(w.countBase, 0, 4ull * (h->chunk + 4 + (size_t)n * oc.nlevels), s));
    if (h->profile) prof_event(h, s);
    // K1: all levels in ONE launch (a CTA carries a strip of a frame through every level in shared memory) when every
    // source row can be read as aligned 32-bit words: always true for the internal buffers, caller device memory only if
    // it is 4-byte aligned with readable row padding.  Otherwise (and for scale factors > 2) the tile kernel runs.
    bool strip = h->useStrip && h->stripOk && oc.nlevels > 1;
    if (strip && !l0Internal)
        strip = ((((uintptr_t)l0.ptr | (uintptr_t)l0.pitch | (uintptr_t)l0.stride) & 3) == 0) &&
                l0.stride >= 4 * (h->stripLv[1].srcLastWord + 1);
    int pyrLaunches = oc.nlevels - 1;
    const int skip = h->skipMask;
    if (skip & 1) {
    } else if (strip) {
        // strips per frame: enough CTAs to fill the GPU (small calls get more, thinner strips), as few as possible
        // otherwise (every strip boundary recomputes the halo rows of all levels)
        const rumi_orb::StripVariant* v = &h->stripVariants.back();
        for (const auto& c : h->stripVariants)
            if (h->stripForce ? c.nstrips >= h->stripForce : (long long)c.nstrips * n >= 128) { v = &c; break; }
        PyrStripArgs sa;
        sa.cv = cv;
        for (int l = 0; l < kMaxLevels; ++l) { sa.lv[l] = h->stripLv[l]; sa.lv[l].rowsPerItem = v->rowsPerItem[l]; }
        sa.nlevels = oc.nlevels; sa.nstrips = v->nstrips; sa.ranges = v->ranges; sa.buf1Offset = v->buf1Offset;
        sa.rowTabOffset = v->rowTabOffset;
        launch_pyramid_strip(sa, v->smemBytes, sh);
        pyrLaunches = 1;
    } else
    for (int l = 1; l < oc.nlevels; ++l) {
        PyramidLevelArgs pa;
        pa.src = cv.src[l - 1]; pa.dst = cv.src[l];
        pa.xc = h->xc[l]; pa.yc = h->yc[l];
        pa.boxW = h->boxW[l]; pa.boxH = h->boxH[l];
        pa.nframes = n; pa.err = h->errDev;
        const CUtensorMap* tm = nullptr;
        CUtensorMap dyn;
        if (h->useTMA) {
            if (l > 1 || l0Internal) {
                if (w.tmapOk[l]) tm = &w.tmap[l];
            } else if (make_tmap(&dyn, l0.ptr, l0.w, l0.h, n, l0.stride, l0.pitch, h->boxW[1], h->boxH[1])) {
                tm = &dyn;
            }
        }
        launch_pyramid_level(pa, tm, sh);
    }
    prof_mark(h, s, ST_PYRAMID, pyrLaunches);
    FastArgs fa;
    fa.cv = cv; fa.cand = w.cand; fa.levelCount = w.levelCount; fa.cellOff = w.cellOff; fa.cellCount = w.cellCount;
    fa.cells = h->fastCells;
    fa.tilePitch = h->fastTilePitch; fa.tileRows = h->fastTileRows;
    fa.scoreRows = h->fastScoreRows;
    fa.lay = fast_layout(fa.tilePitch, fa.tileRows, fa.scoreRows);
    fa.dbg = h->dbgBuf; fa.dbgCell = h->dbgCell;
    OctreeArgs oa;
    oa.nframes = n; oa.levelFirst = 0; oa.dbgClk = h->octClk; oa.cand = w.cand; oa.candOrdered = w.candOrdered; oa.levelCount = w.levelCount;
    oa.cellOff = w.cellOff; oa.cellCount = w.cellCount; oa.bigKeys = w.bigKeys; oa.sel = w.sel;
    oa.selCount = w.selCount; oa.smemKeys = h->smemKeys; oa.maxNodeCap = h->maxNodeCap;
    for (int l = 0; l < oc.nlevels; ++l) {
        // per-call frame stride inside a level block is candCap / bigKeysCap, the level offsets assume `chunk` frames
        fa.candLevelOff[l] = oa.candLevelOff[l] = h->candLevelOff[l];
        oa.bigKeysLevelOff[l] = h->bigKeysLevelOff[l];
        oa.bigKeysCap[l] = h->bigKeysCap[l];
    }
    // Calls of a few frames leave most SMs idle and are bound by the latency of the dependent launches: the blur needs
    // only the pyramid, so it runs on a second stream beside FAST -> quad-tree -> slots and joins before the descriptors
    // (single frame: 17 us off the critical path).  Large chunks fill the GPU on their own.
    static const int forkMax = getenv("RUMI_FORK_MAX") ? atoi(getenv("RUMI_FORK_MAX")) : 4;
    const bool fork = n <= forkMax && !h->profile && !(skip & 16) && getenv("RUMI_NO_FORK") == nullptr;
    if (fork) {
        if (!h->auxStream) {
            CU_TRY(cudaStreamCreateWithFlags(&h->auxStream, cudaStreamNonBlocking));
            CU_TRY(cudaEventCreateWithFlags(&h->evFork, cudaEventDisableTiming));
            CU_TRY(cudaEventCreateWithFlags(&h->evJoinAux, cudaEventDisableTiming));
        }
        CU_TRY(cudaEventRecord(h->evFork, s));
        CU_TRY(cudaStreamWaitEvent(h->auxStream, h->evFork, 0));
        launch_blur(cv, oc, h->auxStream);
        CU_TRY(cudaEventRecord(h->evJoinAux, h->auxStream));
    }
    if (!(skip & 2)) launch_fast(fa, oc, s);
    prof_mark(h, s, ST_FAST, 1);
    DescribeArgs da;
    da.cv = cv; da.sel = w.sel; da.selCount = w.selCount; da.lap0 = lap0; da.lap1 = lap1; da.slot = w.slot;
    da.kps = dKps; da.desc = dDesc; da.nkp = dNkp; da.nmono = dNmono; da.outCap = outCap;
    // the slot assignment rides in the tail of the quad-tree kernel (last CTA of a frame) unless that kernel is skipped,
    // split per level, or RUMI_SLOTS_KERNEL=1 asks for the stand-alone launch (A/B runs)
    static const bool slotsKernel = getenv("RUMI_SLOTS_KERNEL") != nullptr || getenv("RUMI_OCTREE_SPLIT") != nullptr;
    const bool fuse = !(skip & 4) && !slotsKernel;
    oa.fuseSlots = fuse ? 1 : 0;
    oa.frameDone = w.countBase;
    // Quad-tree passes.  A chunk runs ONE launch as long as this handle has never met a level with more candidates than the
    // first pass keeps in shared memory (such a level is then sorted in global memory: correct, slow, and it raises the
    // mapped flag errHost[2]); from the next chunk on dense levels are deferred to a second launch with a large key buffer.
    // Sparse workloads never pay for the second launch (0.4 % of the resident rate), dense ones pay once for learning.
    // RUMI_OCTREE_ONE_PASS=1 / RUMI_OCTREE_TWO_PASS=1 pin either behaviour (A/B runs, tests).
    static const int passMode = getenv("RUMI_OCTREE_TWO_PASS") ? 2 : getenv("RUMI_OCTREE_ONE_PASS") ? 1 : 0;
    const bool canDefer = h->smemKeysBig > h->smemKeys && passMode != 1;
    const bool twoPass = canDefer && (passMode == 2 || *(volatile int*)(h->errHost + 2) != 0);
    oa.bigCount = w.countBase + h->chunk; oa.bigList = twoPass ? w.bigList : nullptr;
    oa.denseFlag = h->errDev + 2;
    oa.smemKeysBig = h->smemKeysBig;
    // a few frames never fill the SMs: every level CTA gets the large key buffer at once
    // and 1024 threads, which hide the latencies nobody else on the SM would (RUMI_OCT_NARROW=1: 256, A/B runs)
    static const bool narrow = getenv("RUMI_OCT_NARROW") != nullptr;
    static const int wideThreads = getenv("RUMI_OCT_WIDE") ? atoi(getenv("RUMI_OCT_WIDE")) : 1024;
    oa.threads = 256;
    if (n <= 4 && canDefer) { oa.smemKeys = h->smemKeysBig; oa.bigList = nullptr; oa.threads = narrow ? 256 : wideThreads; }
    oa.slots.sel = w.sel; oa.slots.selCount = w.selCount; oa.slots.lap0 = lap0; oa.slots.lap1 = lap1; oa.slots.slot = w.slot;
    oa.slots.nkp = dNkp; oa.slots.nmono = dNmono;
    if (!(skip & 4)) launch_octree(oa, oc, sh);
    prof_mark(h, s, ST_OCTREE, oa.bigList ? 2 : 1);
    if (!fuse) launch_assign_slots(da, oc, sh);
    prof_mark(h, s, ST_SLOTS, fuse ? 0 : 1);
    if (fork) CU_TRY(cudaStreamWaitEvent(s, h->evJoinAux, 0));
    else if (!(skip & 16)) launch_blur(cv, oc, s);
    prof_mark(h, s, ST_BLUR, 1);
    if (!(skip & 32)) launch_describe(da, oc, s);
    prof_mark(h, s, ST_DESCRIBE, 1);
    CU_TRY(cudaGetLastError());
    w.lastFrames = n;
    return RUMI_OK;
}

int upload_level0(rumi_orb* h, Workspace& w, const uint8_t* imgs, int n, size_t stride, size_t framePitch) {
    const LevelGeom& g = h->oc.lv[0];
    uint8_t* dst = w.pyr + h->pyrLevelOff[0];
    if ((framePitch == stride * (size_t)g.h || n == 1) && stride == (size_t)g.stride) {
        // rows are contiguous on both sides: one linear copy (a 2-D copy pays per-row DMA descriptors)
        CU_TRY(cudaMemcpyAsync(dst, imgs, (size_t)n * g.h * g.stride, cudaMemcpyHostToDevice, w.stream));
    } else if (framePitch == stride * (size_t)g.h || n == 1) {
        CU_TRY(cudaMemcpy2DAsync(dst, g.stride, imgs, stride, g.w, (size_t)n * g.h, cudaMemcpyHostToDevice, w.stream));
    } else {
        for (int i = 0; i < n; ++i)
            CU_TRY(cudaMemcpy2DAsync(dst + (size_t)i * g.stride * g.h, g.stride, imgs + i * framePitch, stride, g.w,
                                     g.h, cudaMemcpyHostToDevice, w.stream));
    }
    return RUMI_OK;
}

}  // namespace

extern "C" {

const char* rumi_last_error(void) { return g_err.c_str(); }

int rumi_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int rumi_orb_create(rumi_orb** out, int nfeatures, float scale_factor, int nlevels, int ini_th_fast,
                    int min_th_fast, int device, int max_batch) {
    if (!out) return fail(RUMI_ERR_ARG, "out is NULL");
    *out = nullptr;
    if (nlevels < 1 || nlevels > kMaxLevels || nfeatures < 0 || scale_factor <= 1.0f || max_batch < 1)
        return fail(RUMI_ERR_ARG, "bad ORB parameters");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(RUMI_ERR_CUDA, "no CUDA device: librumi_orb has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(RUMI_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
    CU_TRY(cudaSetDevice(device));
    rumi_orb* h = new rumi_orb();
    h->device = device; h->nfeatures = nfeatures; h->scaleFactor = scale_factor; h->nlevels = nlevels;
    h->iniTh = ini_th_fast; h->minTh = min_th_fast; h->chunk = max_batch;
    h->tables = make_scale_tables(nfeatures, scale_factor, nlevels);
    const char* e = getenv("RUMI_NO_TMA");
    h->useTMA = !(e && e[0] == '1');
    const char* pm = getenv("RUMI_PYRAMID");                     // "strip" (default) | "tiles" (TMA / plain tile kernel)
    h->useStrip = !(pm && pm[0] == 't') && h->useTMA;            // RUMI_NO_TMA=1 selects the plain tile kernel
    const char* pst = getenv("RUMI_PYR_STRIPS");
    if (pst) h->stripForce = atoi(pst);
    const char* sk = getenv("RUMI_SKIP_STAGES");
    if (sk) h->skipMask = atoi(sk);
    const char* ns = getenv("RUMI_STREAMS");
    if (ns && ns[0] >= '1' && ns[0] <= '0' + kMaxWs) { h->nws = ns[0] - '0'; h->nwsSet = true; }
    h->nwsDefault = h->nws; h->nwsSetDefault = h->nwsSet;
    if (cudaHostAlloc((void**)&h->errHost, 4 * sizeof(int), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer((void**)&h->errDev, h->errHost, 0) != cudaSuccess) {
        if (h->errHost) cudaFreeHost(h->errHost);
        delete h;
        return fail(RUMI_ERR_CUDA, "cannot allocate the mapped error flags: %s", cudaGetErrorString(cudaGetLastError()));
    }
    h->errHost[0] = h->errHost[1] = h->errHost[2] = h->errHost[3] = 0;
    *out = h;
    return RUMI_OK;
}

void rumi_orb_destroy(rumi_orb* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    for (int i = 0; i < kMaxWs; ++i) {
        if (h->ws[i].stream) cudaStreamSynchronize(h->ws[i].stream);
        free_workspace(h->ws[i]);
        if (h->ws[i].stream) cudaStreamDestroy(h->ws[i].stream);
    }
    cudaFree(h->coef);
    cudaFree(h->fastCells);
    cudaFree(h->pyrTables);
    cudaFree(h->dbgBuf);
    cudaFree(h->octClk);
    cudaFree(h->descBuf);
    if (h->outStage) cudaFreeHost(h->outStage);
    if (h->pyrStage) cudaFreeHost(h->pyrStage);
    if (h->evFork) cudaEventDestroy(h->evFork);
    if (h->evJoinAux) cudaEventDestroy(h->evJoinAux);
    if (h->auxStream) cudaStreamDestroy(h->auxStream);
    if (h->errHost) cudaFreeHost(h->errHost);
    for (cudaEvent_t e : h->evPool) cudaEventDestroy(e);
    if (h->evStart) cudaEventDestroy(h->evStart);
    if (h->evStop) cudaEventDestroy(h->evStop);
    for (int i = 0; i < kMaxWs; ++i) if (h->evJoin[i]) cudaEventDestroy(h->evJoin[i]);
    if (h->evOrder) cudaEventDestroy(h->evOrder);
    for (int i = 0; i < kMaxWs; ++i) if (h->evSignal[i]) cudaEventDestroy(h->evSignal[i]);
    for (int i = 0; i < kMaxWs; ++i) if (h->evCompute[i]) cudaEventDestroy(h->evCompute[i]);
    delete h;
}

int rumi_orb_levels(const rumi_orb* h) { return h ? h->nlevels : 0; }

int rumi_orb_tables(const rumi_orb* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2, int* quota) {
    if (!h) return fail(RUMI_ERR_ARG, "handle is NULL");
    for (int l = 0; l < h->nlevels; ++l) {
        if (scale) scale[l] = h->tables.scale[l];
        if (inv_scale) inv_scale[l] = h->tables.invScale[l];
        if (sigma2) sigma2[l] = h->tables.sigma2[l];
        if (inv_sigma2) inv_sigma2[l] = h->tables.invSigma2[l];
        if (quota) quota[l] = h->tables.quota[l];
    }
    return RUMI_OK;
}

int rumi_orb_frame_capacity(rumi_orb* h, int w, int h_px) {
    if (!h) return fail(RUMI_ERR_ARG, "handle is NULL");
    OrbConst oc;
    const int rc = build_orb_const(oc, w, h_px, h->nfeatures, h->scaleFactor, h->nlevels, h->iniTh, h->minTh);
    if (rc) return fail(RUMI_ERR_SHAPE, "image %dx%d cannot be processed (code %d)", w, h_px, rc);
    return oc.kpCap;
}

namespace {
// memcpy with several host threads: one thread moves ~10 GB/s, the host links take 55 GB/s
void parallel_copy(void* dst, const void* src, size_t bytes, int nthreads) {
    if (bytes < (4u << 20) || nthreads <= 1) { std::memcpy(dst, src, bytes); return; }
    const size_t slice = ((bytes + nthreads - 1) / nthreads + 4095) & ~(size_t)4095;
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; ++t) {
        const size_t o = slice * t;
        if (o >= bytes) break;
        pool.emplace_back([=] { std::memcpy((uint8_t*)dst + o, (const uint8_t*)src + o, std::min(slice, bytes - o)); });
    }
    std::memcpy(dst, src, std::min(slice, bytes));
    for (auto& th : pool) th.join();
}

bool is_pageable(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}

int copy_threads() {
    static const int n = [] {
        const char* e = getenv("RUMI_COPY_THREADS");
        const int hw = (int)std::thread::hardware_concurrency();
        return std::max(1, e ? atoi(e) : std::min(4, std::max(1, hw / 2)));   // 4: measured best (1 / 2 / 4 / 8 / 16 threads)
    }();
    return n;
}

// Host batch path for PAGEABLE caller memory (cv::Mat images, std::vector outputs): a cudaMemcpyAsync from / to pageable
// memory is staged by the driver with one host thread (~8 GB/s: 25 k frames/s).  Here every chunk is copied into the
// workspace's own pinned block by several host threads while the GPU works on the previous chunks, travels with truly
// asynchronous copies, and its results are copied out of a pinned block the same way once their event has fired.
int extract_batch_pageable(rumi_orb* h, const std::vector<int>& sizes, const uint8_t* imgs, size_t frame_pitch, int lap0,
                           int lap1, rumi_kp* kps, uint8_t* desc, int cap_per_frame, int* n_kp, int* n_mono) {
    const OrbConst& oc = h->oc;
    const LevelGeom& g = oc.lv[0];
    const size_t frameBytes = (size_t)g.stride * g.h;
    const size_t kpRow = sizeof(rumi_kp) * (size_t)oc.kpCap, dRow = 32 * (size_t)oc.kpCap;
    const int nt = copy_threads();
    struct Pending { int f0 = 0, m = 0; bool live = false; };
    std::vector<Pending> pend(h->nws);
    auto finalize = [&](int wi) -> int {
        Pending& p = pend[wi];
        if (!p.live) return RUMI_OK;
        Workspace& ws = h->ws[wi];
        CU_TRY(cudaEventSynchronize(ws.evOut));
        const uint8_t* o = ws.hOut;
        const size_t oD = kpRow * h->chunk, oN = oD + dRow * h->chunk, oM = oN + 4 * (size_t)h->chunk;
        if (cap_per_frame == oc.kpCap) {
            parallel_copy(kps + (size_t)p.f0 * cap_per_frame, o, kpRow * p.m, nt);
            parallel_copy(desc + (size_t)p.f0 * cap_per_frame * 32, o + oD, dRow * p.m, nt);
        } else {
            for (int i = 0; i < p.m; ++i) {
                std::memcpy(kps + (size_t)(p.f0 + i) * cap_per_frame, o + kpRow * i, kpRow);
                std::memcpy(desc + (size_t)(p.f0 + i) * cap_per_frame * 32, o + oD + dRow * i, dRow);
            }
        }
        std::memcpy(n_kp + p.f0, o + oN, 4 * (size_t)p.m);
        std::memcpy(n_mono + p.f0, o + oM, 4 * (size_t)p.m);
        p.live = false;
        return RUMI_OK;
    };
    int rc, f0 = 0;
    for (size_t c = 0; c < sizes.size(); ++c) {
        const int wi = (int)(c % (size_t)h->nws), m = sizes[c];
        if ((rc = ensure_workspace(h, wi))) return rc;
        Workspace& ws = h->ws[wi];
        if ((rc = finalize(wi))) return rc;                    // the chunk that used this workspace (and its staging) before
        const size_t inNeed = frameBytes * h->chunk, outNeed = (kpRow + dRow + 8) * (size_t)h->chunk;
        if (ws.hInCap < inNeed) {
            if (ws.hIn) cudaFreeHost(ws.hIn);
            ws.hIn = nullptr; ws.hInCap = 0;
            CU_TRY(cudaHostAlloc((void**)&ws.hIn, inNeed, cudaHostAllocDefault));
            ws.hInCap = inNeed;
        }
        if (ws.hOutCap < outNeed) {
            if (ws.hOut) cudaFreeHost(ws.hOut);
            ws.hOut = nullptr; ws.hOutCap = 0;
            CU_TRY(cudaHostAlloc((void**)&ws.hOut, outNeed, cudaHostAllocDefault));
            ws.hOutCap = outNeed;
        }
        if (!ws.evOut) CU_TRY(cudaEventCreateWithFlags(&ws.evOut, cudaEventDisableTiming));
        parallel_copy(ws.hIn, imgs + (size_t)f0 * frame_pitch, frameBytes * m, nt);
        CU_TRY(cudaMemcpyAsync(ws.pyr + h->pyrLevelOff[0], ws.hIn, frameBytes * m, cudaMemcpyHostToDevice, ws.stream));
        const LevelView l0 = internal_view(h, ws.pyr, 0);
        if ((rc = run_chunk(h, ws, l0, true, m, lap0, lap1, ws.kps, ws.desc, oc.kpCap, ws.nkp, ws.nmono))) return rc;
        const size_t oD = kpRow * h->chunk, oN = oD + dRow * h->chunk, oM = oN + 4 * (size_t)h->chunk;
        CU_TRY(cudaMemcpyAsync(ws.hOut, ws.kps, kpRow * m, cudaMemcpyDeviceToHost, ws.stream));
        CU_TRY(cudaMemcpyAsync(ws.hOut + oD, ws.desc, dRow * m, cudaMemcpyDeviceToHost, ws.stream));
        CU_TRY(cudaMemcpyAsync(ws.hOut + oN, ws.nkp, 4 * (size_t)m, cudaMemcpyDeviceToHost, ws.stream));
        CU_TRY(cudaMemcpyAsync(ws.hOut + oM, ws.nmono, 4 * (size_t)m, cudaMemcpyDeviceToHost, ws.stream));
        CU_TRY(cudaEventRecord(ws.evOut, ws.stream));
        pend[wi].f0 = f0; pend[wi].m = m; pend[wi].live = true;
        h->lastWs = wi;
        f0 += m;
    }
    // drain in submission order
    for (size_t k = 0; k < (size_t)h->nws; ++k) {
        const int wi = (int)((sizes.size() + k) % (size_t)h->nws);
        if ((rc = finalize(wi))) return rc;
    }
    return check_device_flags(h);
}
}  // namespace

int rumi_orb_extract_batch(rumi_orb* h, const uint8_t* imgs, int n, int w, int h_px, size_t stride,
                           size_t frame_pitch, int lap0, int lap1, rumi_kp* kps, uint8_t* desc, int cap_per_frame,
                           int* n_kp, int* n_mono) {
    if (!h) return fail(RUMI_ERR_ARG, "handle is NULL");
    if (!imgs || w <= 0 || h_px <= 0 || n <= 0) return fail(RUMI_ERR_EMPTY, "empty image");
    if (!kps || !desc || !n_kp || !n_mono || stride < (size_t)w) return fail(RUMI_ERR_ARG, "bad output / stride");
    CU_TRY(cudaSetDevice(h->device));
    h->pyrStageValid = false;
    int rc = ensure_geometry(h, w, h_px);
    if (rc) return rc;
    const OrbConst& oc = h->oc;
    if (cap_per_frame < oc.kpCap)
        return fail(RUMI_ERR_CAPACITY, "cap_per_frame %d < frame capacity %d", cap_per_frame, oc.kpCap);
    // Chunk schedule.  Every chunk is H2D -> kernels -> D2H on its own workspace stream, consecutive chunks on different
    // streams, so the copies of one chunk hide behind the kernels of its neighbours -- except the upload of the FIRST
    // chunk and the kernels + download of the LAST one.  A long batch therefore starts and ends with short chunks
    // (1/4, 1/4, 1/2 ... 1/2, 1/4 of the chunk size): 1.1 ms of fill / drain per 1024 frames shrink to a quarter.
    std::vector<int> sizes;
    {
        const int c = h->chunk, q = std::max(c / 4, 1), hf = std::max(c / 2, 1);
        int left = n;
        std::vector<int> tail;
        if ((!h->profile || getenv("RUMI_RAMP_HEAD")) && n >= 6 * c && c >= 16) {
            std::vector<int> head = {q, q, hf}, tl = {q, hf};
            auto parse = [&](const char* e, std::vector<int>& v) {              // "16,16,32" (measurement runs only)
                if (!e) return;
                v.clear();
                for (const char* p = e; *p;) { v.push_back(std::min(c, std::max(1, atoi(p)))); while (*p && *p != ',') ++p; if (*p) ++p; }
            };
            parse(getenv("RUMI_RAMP_HEAD"), head);
            parse(getenv("RUMI_RAMP_TAIL"), tl);
            for (int sz : head) { sizes.push_back(sz); left -= sz; }
            for (int sz : tl) { tail.push_back(sz); left -= sz; }
        }
        while (left > 0) { const int sz = std::min(c, left); sizes.push_back(sz); left -= sz; }
        for (size_t i = tail.size(); i-- > 0;) sizes.push_back(tail[i]);
    }
    // pageable caller memory, rows contiguous: the library stages the chunks itself with several host threads
    if (!h->profile && stride == (size_t)oc.lv[0].stride && frame_pitch == stride * (size_t)h_px && getenv("RUMI_NO_HOST_STAGING") == nullptr &&
        is_pageable(imgs) && is_pageable(kps) && is_pageable(desc))
        return extract_batch_pageable(h, sizes, imgs, frame_pitch, lap0, lap1, kps, desc, cap_per_frame, n_kp, n_mono);
    int f0 = 0;
    for (size_t c = 0; c < sizes.size(); ++c) {
        const int wi = (int)(c % (size_t)h->nws);
        if ((rc = ensure_workspace(h, wi))) return rc;
        Workspace& ws = h->ws[wi];
        const int m = sizes[c];
        // stream order makes the reuse of workspace `wi` (nws chunks ago) safe
        if (h->profile) prof_event(h, ws.stream);
        if ((rc = upload_level0(h, ws, imgs + (size_t)f0 * frame_pitch, m, stride, frame_pitch))) return rc;
        if (h->profile) { h->evSpans.push_back(std::make_pair((int)ST_H2D, (int)h->evUsed - 1)); prof_event(h, ws.stream); }
        const LevelView l0 = internal_view(h, ws.pyr, 0);
        // RUMI_E2E_GATE=k lets at most k chunks compute at a time (A/B runs).  Round 1 needed k = 2 (more chunks thrashed the
        // L2); with the round-2 kernels the ungated pipeline is faster (150.0 -> 154.6 k frames/s end to end), so: no gate.
        static const int gate = getenv("RUMI_E2E_GATE") ? atoi(getenv("RUMI_E2E_GATE")) : 0;
        if (gate > 0 && h->nws > gate && (int)c >= gate) CU_TRY(cudaStreamWaitEvent(ws.stream, h->evCompute[(c - gate) % kMaxWs], 0));
        if ((rc = run_chunk(h, ws, l0, true, m, lap0, lap1, ws.kps, ws.desc, oc.kpCap, ws.nkp, ws.nmono))) return rc;
        if (gate > 0 && h->nws > gate) {
            if (!h->evCompute[c % kMaxWs]) CU_TRY(cudaEventCreateWithFlags(&h->evCompute[c % kMaxWs], cudaEventDisableTiming));
            CU_TRY(cudaEventRecord(h->evCompute[c % kMaxWs], ws.stream));
        }
        // results: dense [m][kpCap] blocks -> caller's [n][cap_per_frame] layout
        if (cap_per_frame == oc.kpCap) {
            CU_TRY(cudaMemcpyAsync(kps + (size_t)f0 * cap_per_frame, ws.kps, sizeof(rumi_kp) * (size_t)oc.kpCap * m,
                                   cudaMemcpyDeviceToHost, ws.stream));
            CU_TRY(cudaMemcpyAsync(desc + (size_t)f0 * cap_per_frame * 32, ws.desc, 32 * (size_t)oc.kpCap * m,
                                   cudaMemcpyDeviceToHost, ws.stream));
        } else {
            CU_TRY(cudaMemcpy2DAsync(kps + (size_t)f0 * cap_per_frame, sizeof(rumi_kp) * (size_t)cap_per_frame, ws.kps,
                                     sizeof(rumi_kp) * (size_t)oc.kpCap, sizeof(rumi_kp) * (size_t)oc.kpCap, m,
                                     cudaMemcpyDeviceToHost, ws.stream));
            CU_TRY(cudaMemcpy2DAsync(desc + (size_t)f0 * cap_per_frame * 32, 32 * (size_t)cap_per_frame, ws.desc,
                                     32 * (size_t)oc.kpCap, 32 * (size_t)oc.kpCap, m, cudaMemcpyDeviceToHost, ws.stream));
        }
        CU_TRY(cudaMemcpyAsync(n_kp + f0, ws.nkp, 4 * (size_t)m, cudaMemcpyDeviceToHost, ws.stream));
        CU_TRY(cudaMemcpyAsync(n_mono + f0, ws.nmono, 4 * (size_t)m, cudaMemcpyDeviceToHost, ws.stream));
        if (h->profile) { h->evSpans.push_back(std::make_pair((int)ST_D2H, (int)h->evUsed - 1)); prof_event(h, ws.stream); }
        h->lastWs = wi;
        f0 += m;
    }
    for (int i = 0; i < kMaxWs; ++i)
        if (h->ws[i].stream) CU_TRY(cudaStreamSynchronize(h->ws[i].stream));
    return check_device_flags(h);
}

int rumi_orb_extract_begin(rumi_orb* h, const uint8_t* img, int w, int h_px, size_t stride, int lap0, int lap1) {
    if (!h) return fail(RUMI_ERR_ARG, "handle is NULL");
    if (!img || w <= 0 || h_px <= 0) return fail(RUMI_ERR_EMPTY, "empty image");
    if (stride < (size_t)w) return fail(RUMI_ERR_ARG, "bad stride");
    CU_TRY(cudaSetDevice(h->device));
    int rc = ensure_geometry(h, w, h_px);
    if (rc) return rc;
    if ((rc = ensure_workspace(h, 0))) return rc;
    Workspace& ws = h->ws[0];
    const OrbConst& oc = h->oc;
    if ((rc = upload_level0(h, ws, img, 1, stride, stride * (size_t)h_px))) return rc;
    const LevelView l0 = internal_view(h, ws.pyr, 0);
    if ((rc = run_chunk(h, ws, l0, true, 1, lap0, lap1, ws.kps, ws.desc, oc.kpCap, ws.nkp, ws.nmono))) return rc;
    // Results: counts, key points and descriptors of the whole frame capacity (60 KB) go to a pinned staging block with four
    // back-to-back asynchronous copies; rumi_orb_extract_end synchronises ONCE and copies the rows in use to the caller's
    // (possibly pageable) buffers on the host.  (Was: two count copies, a sync, then two synchronous copies -- three round
    // trips; the single-frame latency is dominated by such fixed costs, 120 us of kernels in a 213 us call.)
    const size_t kpBytes = sizeof(rumi_kp) * (size_t)oc.kpCap, dBytes = 32 * (size_t)oc.kpCap, need = 64 + kpBytes + dBytes;
    if (need > h->outStageCap) {
        if (h->outStage) cudaFreeHost(h->outStage);
        h->outStage = nullptr; h->outStageCap = 0;
        CU_TRY(cudaHostAlloc((void**)&h->outStage, need, cudaHostAllocDefault));
        h->outStageCap = need;
    }
    int* counts = reinterpret_cast<int*>(h->outStage);
    uint8_t* sKps = h->outStage + 64;
    CU_TRY(cudaMemcpyAsync(&counts[0], ws.nkp, 4, cudaMemcpyDeviceToHost, ws.stream));
    CU_TRY(cudaMemcpyAsync(&counts[1], ws.nmono, 4, cudaMemcpyDeviceToHost, ws.stream));
    CU_TRY(cudaMemcpyAsync(sKps, ws.kps, kpBytes, cudaMemcpyDeviceToHost, ws.stream));
    CU_TRY(cudaMemcpyAsync(sKps + kpBytes, ws.desc, dBytes, cudaMemcpyDeviceToHost, ws.stream));
    h->pyrStageValid = false;
    if (h->stagePyr) {
        // the pyramid of this frame follows the results into a pinned block: level 0 is the caller's image (copied on the host
        // while the GPU works), levels 1.. come down with asynchronous copies; rumi_orb_pyramid_level then needs no round trip
        size_t total = 0;
        for (int l = 0; l < oc.nlevels; ++l) { h->pyrStageOff[l] = total; total += (size_t)oc.lv[l].stride * oc.lv[l].h; }
        if (total > h->pyrStageCap) {
            if (h->pyrStage) cudaFreeHost(h->pyrStage);
            h->pyrStage = nullptr; h->pyrStageCap = 0;
            CU_TRY(cudaHostAlloc((void**)&h->pyrStage, total, cudaHostAllocDefault));
            h->pyrStageCap = total;
        }
        for (int l = 1; l < oc.nlevels; ++l) {
            const LevelView v = internal_view(h, ws.pyr, l);
            CU_TRY(cudaMemcpyAsync(h->pyrStage + h->pyrStageOff[l], v.ptr, (size_t)v.stride * v.h, cudaMemcpyDeviceToHost, ws.stream));
        }
        const int s0 = oc.lv[0].stride;
        for (int y = 0; y < h_px; ++y) std::memcpy(h->pyrStage + (size_t)y * s0, img + (size_t)y * stride, (size_t)w);
        h->pyrStageValid = true;                       // complete once the stream has been synchronised (rumi_orb_extract_end)
    }
    h->pendingSingle = true;
    h->lastWs = 0;
    return RUMI_OK;
}

int rumi_orb_extract_end(rumi_orb* h, rumi_kp* kps, uint8_t* desc, int cap, int* n_kp, int* n_mono) {
    if (!h) return fail(RUMI_ERR_ARG, "handle is NULL");
    if (!n_kp || !n_mono) return fail(RUMI_ERR_ARG, "bad output");
    if (!h->pendingSingle) return fail(RUMI_ERR_ARG, "rumi_orb_extract_end without rumi_orb_extract_begin");
    h->pendingSingle = false;
    CU_TRY(cudaSetDevice(h->device));
    CU_TRY(cudaStreamSynchronize(h->ws[0].stream));
    int rc = check_device_flags(h);
    if (rc) return rc;
    const OrbConst& oc = h->oc;
    const int* counts = reinterpret_cast<const int*>(h->outStage);
    const uint8_t* sKps = h->outStage + 64;
    const uint8_t* sDesc = sKps + sizeof(rumi_kp) * (size_t)oc.kpCap;
    *n_kp = counts[0]; *n_mono = counts[1];
    const int m = std::min(counts[0], cap);
    if (m > 0) {
        if (!kps || !desc) return fail(RUMI_ERR_ARG, "kps/desc NULL");
        std::memcpy(kps, sKps, sizeof(rumi_kp) * (size_t)m);
        std::memcpy(desc, sDesc, 32 * (size_t)m);
    }
    return RUMI_OK;
}

int rumi_orb_extract(rumi_orb* h, const uint8_t* img, int w, int h_px, size_t stride, int lap0, int lap1,
                     rumi_kp* kps, uint8_t* desc, int cap, int* n_kp, int* n_mono) {
    if (!n_kp || !n_mono) return fail(RUMI_ERR_ARG, "bad output / stride");
    const int rc = rumi_orb_extract_begin(h, img, w, h_px, stride, lap0, lap1);
    if (rc) return rc;
    return rumi_orb_extract_end(h, kps, desc, cap, n_kp, n_mono);
}

int rumi_orb_extract_batch_device(rumi_orb* h, const uint8_t* d_imgs, int n, int w, int h_px, size_t stride,
                                  size_t frame_pitch, int lap0, int lap1, rumi_kp* d_kps, uint8_t* d_desc,
                                  int cap_per_frame, int* d_n_kp, int* d_n_mono, int sync) {
    if (!h) return fail(RUMI_ERR_ARG, "handle is NULL");
    if (!d_imgs || w <= 0 || h_px <= 0 || n <= 0) return fail(RUMI_ERR_EMPTY, "empty image");
    if (!d_kps || !d_desc || !d_n_kp || !d_n_mono || stride < (size_t)w) return fail(RUMI_ERR_ARG, "bad output / stride");
    CU_TRY(cudaSetDevice(h->device));
    h->pyrStageValid = false;
    int rc = ensure_geometry(h, w, h_px);
    if (rc) return rc;
    if (cap_per_frame < h->oc.kpCap)
        return fail(RUMI_ERR_CAPACITY, "cap_per_frame %d < frame capacity %d", cap_per_frame, h->oc.kpCap);
    const int nchunks = (n + h->chunk - 1) / h->chunk;
    // resident input: two workspaces are enough to overlap the latency-bound kernels of one chunk with the
    // throughput-bound ones of the next (measured: 2 streams 118k frames/s, 4 streams 110k -- more chunks in flight
    // only thrash L2); the host path keeps nws workspaces because it also has copies to hide
    const int nws = h->nwsSet ? h->nws : std::min(h->nws, 2);
    for (int c = 0; c < nchunks; ++c) {
        const int wi = c % nws;
        if ((rc = ensure_workspace(h, wi))) return rc;
        Workspace& ws = h->ws[wi];
        const int f0 = c * h->chunk, m = std::min(h->chunk, n - f0);
        LevelView l0;
        l0.ptr = d_imgs + (size_t)f0 * frame_pitch; l0.pitch = (long long)frame_pitch; l0.stride = (int)stride;
        l0.w = w; l0.h = h_px;
        if ((rc = run_chunk(h, ws, l0, false, m, lap0, lap1, reinterpret_cast<KeyPointRec*>(d_kps) + (size_t)f0 * cap_per_frame,
                            d_desc + (size_t)f0 * cap_per_frame * 32, cap_per_frame, d_n_kp + f0, d_n_mono + f0)))
            return rc;
        h->lastWs = wi;
    }
    if (sync) {
        for (int i = 0; i < kMaxWs; ++i)
            if (h->ws[i].stream) CU_TRY(cudaStreamSynchronize(h->ws[i].stream));
        return check_device_flags(h);
    }
    return RUMI_OK;
}

int rumi_orb_describe_batch(rumi_orb* h, const uint8_t* imgs, int nimg, int w, int h_px, size_t stride, size_t frame_pitch,
                            const rumi_kp* kps, const int32_t* kp_off, uint8_t* desc) {
    if (!h) return fail(RUMI_ERR_ARG, "handle is NULL");
    if (!imgs || w <= 0 || h_px <= 0 || nimg <= 0) return fail(RUMI_ERR_EMPTY, "empty image");
    if (stride < (size_t)w || !kp_off || kp_off[0] != 0) return fail(RUMI_ERR_ARG, "bad arguments");
    for (int f = 0; f < nimg; ++f)
        if (kp_off[f + 1] < kp_off[f]) return fail(RUMI_ERR_ARG, "kp_off must be non-decreasing");
    const int n = kp_off[nimg];
    if (n == 0) return 0;
    if (!kps || !desc) return fail(RUMI_ERR_ARG, "NULL keypoints / descriptors");
    for (int i = 0; i < n; ++i) {
        const int x = rint_f(kps[i].x), y = rint_f(kps[i].y);
        if (x < kEdge || y < kEdge || x >= w - kEdge || y >= h_px - kEdge)
            return fail(RUMI_ERR_BORDER, "keypoint %d at (%d,%d) is closer than 19 px to the border", i, x, y);
    }
    CU_TRY(cudaSetDevice(h->device));
    if (!h->ws[0].stream) CU_TRY(cudaStreamCreateWithFlags(&h->ws[0].stream, cudaStreamNonBlocking));
    cudaStream_t s = h->ws[0].stream;
    const int dstride = align_up(w, 16);
    const size_t imgBytes = (size_t)dstride * h_px;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t oK = al(imgBytes * nimg), oD = oK + al(sizeof(KeyPointRec) * (size_t)n), oO = oD + al(32 * (size_t)n),
                 need = oO + al(4 * ((size_t)nimg + 1));
    if (need > h->descCap) {                                   // scratch of the describe calls: grows, never shrinks
        cudaFree(h->descBuf); h->descBuf = nullptr; h->descCap = 0;
        CU_TRY(cudaMalloc(&h->descBuf, need));
        h->descCap = need;
    }
    uint8_t* p = h->descBuf;
    if (frame_pitch == stride * (size_t)h_px || nimg == 1) {
        CU_TRY(cudaMemcpy2DAsync(p, dstride, imgs, stride, w, (size_t)h_px * nimg, cudaMemcpyHostToDevice, s));
    } else {
        for (int f = 0; f < nimg; ++f)
            CU_TRY(cudaMemcpy2DAsync(p + f * imgBytes, dstride, imgs + f * frame_pitch, stride, w, h_px, cudaMemcpyHostToDevice, s));
    }
    CU_TRY(cudaMemcpyAsync(p + oK, kps, sizeof(KeyPointRec) * (size_t)n, cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpyAsync(p + oO, kp_off, 4 * ((size_t)nimg + 1), cudaMemcpyHostToDevice, s));
    launch_describe_given(p, w, h_px, dstride, (long long)imgBytes, nimg, reinterpret_cast<const int*>(p + oO),
                          reinterpret_cast<const KeyPointRec*>(p + oK), n, p + oD, s);
    h->launches += 1;
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(desc, p + oD, 32 * (size_t)n, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    return n;
}

int rumi_orb_describe(rumi_orb* h, const uint8_t* img, int w, int h_px, size_t stride, const rumi_kp* kps, int n,
                      uint8_t* desc) {
    if (n < 0) return fail(RUMI_ERR_ARG, "bad arguments");
    const int32_t off[2] = {0, n};
    return rumi_orb_describe_batch(h, img, 1, w, h_px, stride, stride * (size_t)(h_px > 0 ? h_px : 0), kps, off, desc);
}

static int copy_level(rumi_orb* h, const uint8_t* base, int level, uint8_t* dst, size_t dst_stride, int* w, int* hp) {
    if (!h || !h->coef) return fail(RUMI_ERR_ARG, "no extraction has run on this handle");
    if (level < 0 || level >= h->nlevels) return fail(RUMI_ERR_ARG, "level out of range");
    const LevelView v = internal_view(h, base, level);
    if (w) *w = v.w;
    if (hp) *hp = v.h;
    if (!dst) return RUMI_OK;
    if (dst_stride < (size_t)v.w) return fail(RUMI_ERR_ARG, "dst_stride too small");
    CU_TRY(cudaSetDevice(h->device));
    CU_TRY(cudaStreamSynchronize(h->ws[h->lastWs].stream));
    if (h->pyrStageValid && h->lastWs == 0 && base == h->ws[0].pyr) {          // staged behind the last single-frame call
        const uint8_t* src = h->pyrStage + h->pyrStageOff[level];
        for (int y = 0; y < v.h; ++y) std::memcpy(dst + (size_t)y * dst_stride, src + (size_t)y * v.stride, (size_t)v.w);
        return RUMI_OK;
    }
    CU_TRY(cudaMemcpy2D(dst, dst_stride, v.ptr, v.stride, v.w, v.h, cudaMemcpyDeviceToHost));
    return RUMI_OK;
}

int rumi_orb_set_pyramid_staging(rumi_orb* h, int on) {
    if (!h) return fail(RUMI_ERR_ARG, "handle is NULL");
    h->stagePyr = on != 0;
    if (!on) h->pyrStageValid = false;
    return RUMI_OK;
}

int rumi_orb_pyramid_level(rumi_orb* h, int level, uint8_t* dst, size_t dst_stride, int* w, int* h_px) {
    if (!h) return fail(RUMI_ERR_ARG, "handle is NULL");
    return copy_level(h, h->ws[h->lastWs].pyr, level, dst, dst_stride, w, h_px);
}

int rumi_orb_blurred_level(rumi_orb* h, int level, uint8_t* dst, size_t dst_stride, int* w, int* h_px) {
    if (!h) return fail(RUMI_ERR_ARG, "handle is NULL");
    return copy_level(h, h->ws[h->lastWs].blur, level, dst, dst_stride, w, h_px);
}

static int debug_list(rumi_orb* h, int level, int32_t* xyr, int cap, bool selected) {
    if (!h || !h->coef) return fail(RUMI_ERR_ARG, "no extraction has run on this handle");
    if (level < 0 || level >= h->nlevels) return fail(RUMI_ERR_ARG, "level out of range");
    CU_TRY(cudaSetDevice(h->device));
    Workspace& w = h->ws[h->lastWs];
    CU_TRY(cudaStreamSynchronize(w.stream));
    const OrbConst& oc = h->oc;
    int cnt = 0;
    const int* dcnt = (selected ? w.selCount : w.levelCount) + level;     // frame 0
    CU_TRY(cudaMemcpy(&cnt, dcnt, 4, cudaMemcpyDeviceToHost));
    std::vector<uint32_t> buf(std::max(cnt, 1));
    const uint32_t* src = selected ? w.sel + oc.lv[level].kpBase : w.candOrdered + h->candLevelOff[level];
    if (cnt > 0) CU_TRY(cudaMemcpy(buf.data(), src, 4 * (size_t)cnt, cudaMemcpyDeviceToHost));
    for (int i = 0; i < cnt && i < cap; ++i) {
        xyr[3 * i] = cand_x(buf[i]); xyr[3 * i + 1] = cand_y(buf[i]); xyr[3 * i + 2] = cand_resp(buf[i]);
    }
    return cnt;
}

int rumi_orb_timer_start(rumi_orb* h) {
    if (!h) return fail(RUMI_ERR_ARG, "handle is NULL");
    CU_TRY(cudaSetDevice(h->device));
    for (int i = 0; i < kMaxWs; ++i)
        if (!h->ws[i].stream) CU_TRY(cudaStreamCreateWithFlags(&h->ws[i].stream, cudaStreamNonBlocking));
    if (!h->evStart) {
        CU_TRY(cudaEventCreate(&h->evStart)); CU_TRY(cudaEventCreate(&h->evStop));
        for (int i = 0; i < kMaxWs; ++i) CU_TRY(cudaEventCreate(&h->evJoin[i]));
    }
    CU_TRY(cudaEventRecord(h->evStart, h->ws[0].stream));
    for (int i = 1; i < kMaxWs; ++i) CU_TRY(cudaStreamWaitEvent(h->ws[i].stream, h->evStart, 0));
    return RUMI_OK;
}

int rumi_orb_timer_stop(rumi_orb* h, float* ms) {
    if (!h || !h->evStart || !ms) return fail(RUMI_ERR_ARG, "timer not started");
    CU_TRY(cudaSetDevice(h->device));
    for (int i = 1; i < kMaxWs; ++i) {
        CU_TRY(cudaEventRecord(h->evJoin[i], h->ws[i].stream));
        CU_TRY(cudaStreamWaitEvent(h->ws[0].stream, h->evJoin[i], 0));
    }
    CU_TRY(cudaEventRecord(h->evStop, h->ws[0].stream));
    CU_TRY(cudaEventSynchronize(h->evStop));
    CU_TRY(cudaEventElapsedTime(ms, h->evStart, h->evStop));
    return check_device_flags(h);
}

// Ordering against the caller's own CUDA work (device-resident entry points): the library launches on private
// non-blocking streams, which no other stream orders itself against implicitly.
int rumi_orb_wait_stream(rumi_orb* h, void* stream) {
    if (!h) return fail(RUMI_ERR_ARG, "handle is NULL");
    CU_TRY(cudaSetDevice(h->device));
    if (!h->evOrder) CU_TRY(cudaEventCreateWithFlags(&h->evOrder, cudaEventDisableTiming));
    CU_TRY(cudaEventRecord(h->evOrder, (cudaStream_t)stream));
    for (int i = 0; i < kMaxWs; ++i) {
        if (!h->ws[i].stream) CU_TRY(cudaStreamCreateWithFlags(&h->ws[i].stream, cudaStreamNonBlocking));
        CU_TRY(cudaStreamWaitEvent(h->ws[i].stream, h->evOrder, 0));
    }
    return RUMI_OK;
}

int rumi_orb_signal_stream(rumi_orb* h, void* stream) {
    if (!h) return fail(RUMI_ERR_ARG, "handle is NULL");
    CU_TRY(cudaSetDevice(h->device));
    for (int i = 0; i < kMaxWs; ++i) {
        if (!h->ws[i].stream) continue;
        if (!h->evSignal[i]) CU_TRY(cudaEventCreateWithFlags(&h->evSignal[i], cudaEventDisableTiming));
        CU_TRY(cudaEventRecord(h->evSignal[i], h->ws[i].stream));
        CU_TRY(cudaStreamWaitEvent((cudaStream_t)stream, h->evSignal[i], 0));
    }
    return RUMI_OK;
}

int rumi_orb_debug_skip_stages(rumi_orb* h, int mask) {
    if (!h) return fail(RUMI_ERR_ARG, "handle is NULL");
    h->skipMask = mask;
    return RUMI_OK;
}

int rumi_orb_set_streams(rumi_orb* h, int n) {
    if (!h) return fail(RUMI_ERR_ARG, "handle is NULL");
    if (n < 0 || n > kMaxWs) return fail(RUMI_ERR_ARG, "streams must be 0..%d", kMaxWs);
    CU_TRY(cudaSetDevice(h->device));
    for (int i = 0; i < kMaxWs; ++i)
        if (h->ws[i].stream) CU_TRY(cudaStreamSynchronize(h->ws[i].stream));
    if (n == 0) { h->nws = h->nwsDefault; h->nwsSet = h->nwsSetDefault; }
    else { h->nws = n; h->nwsSet = true; }
    return RUMI_OK;
}

int rumi_orb_profile(rumi_orb* h, int enable) {
    if (!h) return fail(RUMI_ERR_ARG, "handle is NULL");
    h->profile = enable != 0;
    return RUMI_OK;
}

int rumi_orb_profile_read(rumi_orb* h, double* stage_ms, long long* stage_launches, int reset) {
    if (!h) return fail(RUMI_ERR_ARG, "handle is NULL");
    CU_TRY(cudaSetDevice(h->device));
    for (int i = 0; i < kMaxWs; ++i)
        if (h->ws[i].stream) CU_TRY(cudaStreamSynchronize(h->ws[i].stream));
    FILE* tl = nullptr;                                        // RUMI_TIMELINE=<file>: "stage start_ms end_ms" per span
    if (const char* path = getenv("RUMI_TIMELINE")) tl = fopen(path, "a");
    for (const auto& sp : h->evSpans) {
        float ms = 0.f;
        CU_TRY(cudaEventElapsedTime(&ms, h->evPool[sp.second], h->evPool[sp.second + 1]));
        h->stageMs[sp.first] += ms;
        if (tl) {
            float t0 = 0.f;
            cudaEventElapsedTime(&t0, h->evPool[h->evSpans.front().second], h->evPool[sp.second]);
            fprintf(tl, "%d %.4f %.4f\n", sp.first, t0, t0 + ms);
        }
    }
    if (tl) { fprintf(tl, "-1 0 0\n"); fclose(tl); }
    h->evSpans.clear();
    h->evUsed = 0;
    for (int i = 0; i < ST_COUNT; ++i) {
        if (stage_ms) stage_ms[i] = h->stageMs[i];
        if (stage_launches) stage_launches[i] = h->stageLaunches[i];
    }
    if (reset)
        for (int i = 0; i < 8; ++i) { h->stageMs[i] = 0; h->stageLaunches[i] = 0; }
    return ST_COUNT;
}

long long rumi_orb_launch_count(rumi_orb* h, int reset) {
    if (!h) return 0;
    const long long v = h->launches;
    if (reset) h->launches = 0;
    return v;
}

int rumi_orb_debug_fast_tile(rumi_orb* h, int cell, uint8_t* out, int cap, int* dims5) {
    if (!h) return fail(RUMI_ERR_ARG, "handle is NULL");
    CU_TRY(cudaSetDevice(h->device));
    if (!h->dbgBuf) { CU_TRY(cudaMalloc(&h->dbgBuf, 1 << 16)); CU_TRY(cudaMemset(h->dbgBuf, 0, 1 << 16)); }
    h->dbgCell = cell;
    if (out && h->coef) {
        const int nb = h->fastTilePitch * h->fastTileRows + h->fastTilePitch * h->fastScoreRows;
        CU_TRY(cudaDeviceSynchronize());
        CU_TRY(cudaMemcpy(out, h->dbgBuf, std::min(nb, cap), cudaMemcpyDeviceToHost));
        if (dims5) { dims5[0] = h->fastTilePitch; dims5[1] = h->fastTileRows; dims5[2] = h->fastTilePitch; dims5[3] = h->fastScoreRows; dims5[4] = 0; }
        return nb;
    }
    return 0;
}

int rumi_orb_debug_octree_clocks(rumi_orb* h, long long* out, int cap) {
    if (!h) return fail(RUMI_ERR_ARG, "handle is NULL");
    CU_TRY(cudaSetDevice(h->device));
    const int n = 16 * kMaxLevels;
    if (!h->octClk) { CU_TRY(cudaMalloc(&h->octClk, 8 * n)); CU_TRY(cudaMemset(h->octClk, 0, 8 * n)); CU_TRY(cudaDeviceSynchronize()); return 0; }
    CU_TRY(cudaDeviceSynchronize());
    if (out) CU_TRY(cudaMemcpy(out, h->octClk, 8 * (size_t)std::min(n, cap), cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemset(h->octClk, 0, 8 * n));
    CU_TRY(cudaDeviceSynchronize());
    return n;
}

int rumi_orb_debug_candidates(rumi_orb* h, int level, int32_t* xyr, int cap) { return debug_list(h, level, xyr, cap, false); }
int rumi_orb_debug_selected(rumi_orb* h, int level, int32_t* xyr, int cap) { return debug_list(h, level, xyr, cap, true); }

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// matching
struct rumi_match {
    int device = 0;
    cudaStream_t stream = nullptr;
    uint64_t* partial = nullptr; size_t partialCap = 0;
    uint8_t *dQ = nullptr, *dT = nullptr; size_t qCap = 0, tCap = 0;
    uint8_t* dOut = nullptr; size_t outCap = 0;      // idx1 (4 B) | d1 (2 B) | d2 (2 B) per query
    uint8_t* hStage = nullptr; size_t hStageCap = 0; // pinned: inputs / results of per-frame sized calls (one round trip)
    uint8_t* tx = nullptr; size_t txCap = 0;         // train set as ready-to-load UMMA operand tiles (K8-P)
    int mode = 0;                                    // 0 auto, 1 LOP3+POPC kernel only, 3 tcgen05 UMMA only
    int lastPath = 0;                                // kernel of the last top-2 call: 1 LOP3+POPC, 3 UMMA
    cudaEvent_t evStart = nullptr, evStop = nullptr;
    cudaEvent_t evOrder = nullptr;                   // ordering with the caller's streams (wait_stream / signal_stream)
    long long launches = 0;
    // multi-GPU exchange step (one process per GPU): NCCL communicator, this rank, gather buffers
    void* comm = nullptr; bool ownComm = false; int rank = 0, nranks = 1;
    uint64_t *shardSend = nullptr, *shardRecv = nullptr; size_t sendCap = 0, recvCap = 0;
};

namespace {
int grow(void** p, size_t* cap, size_t need) {
    if (need <= *cap) return RUMI_OK;
    cudaFree(*p);
    *p = nullptr; *cap = 0;
    CU_TRY(cudaMalloc(p, need));
    *cap = need;
    return RUMI_OK;
}

// Per-frame sized calls with HOST pointers are bound by round trips, not by bytes: copies from / to pageable memory are
// synchronous, one by one (1000 x 1000 top-2: 0.10 ms for ~10 us of kernel).  Below kSmallCallBytes the arguments are packed
// into ONE pinned block on the host, travel with asynchronous copies, and the results come back in one copy + one sync.
constexpr size_t kSmallCallBytes = 4u << 20;
int host_stage(rumi_match* m, size_t need) {
    if (need <= m->hStageCap) return RUMI_OK;
    if (m->hStage) cudaFreeHost(m->hStage);
    m->hStage = nullptr; m->hStageCap = 0;
    CU_TRY(cudaHostAlloc((void**)&m->hStage, need, cudaHostAllocDefault));
    m->hStageCap = need;
    return RUMI_OK;
}

// Runs the top-2 scan of nq queries against nt train rows; leaves `*slices` candidate arrays [slice][nq] in m->partial.
int top2_partials(rumi_match* m, const uint8_t* dQ, int nq, const uint8_t* dT, int nt, int tBase, int* slicesOut) {
    if (((uintptr_t)dQ | (uintptr_t)dT) & 15) return fail(RUMI_ERR_ARG, "descriptor arrays must be 16-byte aligned");
    // Large problems go to the tensor cores: tcgen05 kind::i8 UMMA with TMEM accumulators (K8-U, operands expanded
    // in-kernel from the packed rows); small ones (stereo bands, single frames) stay on the LOP3+POPC kernel, whose
    // set-up is cheaper.
    const bool big = (long long)nq * nt >= 64ll * 1024 * 1024 && nq >= 256;
    const bool fits = nt > 0 && nt <= (1 << 22);
    const bool umma = fits && (m->mode == 3 || (m->mode == 0 && big));
    int rc;
    if (umma) {
        const int slices = umma_slices(nq, nt);
        if ((rc = grow((void**)&m->partial, &m->partialCap, 8 * (size_t)slices * nq))) return rc;
        if ((rc = grow((void**)&m->tx, &m->txCap, umma_train_bytes(nt)))) return rc;
        launch_hamming_top2_umma(dQ, nq, dT, nt, m->tx, tBase, slices, m->partial, m->stream);
        m->launches += 3;                                   // train pre-pass, partial fill, the persistent UMMA kernel
        m->lastPath = 3;
        *slicesOut = slices;
    } else {
        const int slices = match_slices(nq, nt);
        if ((rc = grow((void**)&m->partial, &m->partialCap, 8 * (size_t)slices * nq))) return rc;
        launch_hamming_top2_partial(dQ, nq, dT, nt, tBase, slices, m->partial, m->stream);
        m->launches += 1;
        m->lastPath = 1;
        *slicesOut = slices;
    }
    CU_TRY(cudaGetLastError());
    return RUMI_OK;
}

int top2_device(rumi_match* m, const uint8_t* dQ, int nq, const uint8_t* dT, int nt, int tBase, int32_t* dIdx,
                uint16_t* dD1, uint16_t* dD2) {
    int slices = 1;
    const int rc = top2_partials(m, dQ, nq, dT, nt, tBase, &slices);
    if (rc) return rc;
    launch_top2_merge(m->partial, slices, nq, dIdx, dD1, dD2, m->stream);
    m->launches += 1;
    CU_TRY(cudaGetLastError());
    return RUMI_OK;
}

// ---- NCCL, resolved at run time (the library has no link-time dependency on it; a process that already loaded
//      libnccl.so.2 -- e.g. through torch -- gets that copy) ----
struct NcclApi {
    void* lib = nullptr;
    int (*getUniqueId)(void* id) = nullptr;
    int (*commDestroy)(void* comm) = nullptr;
    int (*allGather)(const void* send, void* recv, size_t count, int dtype, void* comm, cudaStream_t s) = nullptr;
    const char* (*errorString)(int) = nullptr;
    void* commInitRank = nullptr;            // ncclCommInitRank(ncclComm_t*, int, ncclUniqueId (by value, 128 B), int)
    bool ok = false;
};
struct NcclId { char internal[128]; };       // == ncclUniqueId (nccl.h: NCCL_UNIQUE_ID_BYTES 128)
constexpr int kNcclUint64 = 5;               // ncclDataType_t ncclUint64 (nccl.h)

NcclApi& nccl_api() {
    static NcclApi api = [] {                          // thread-safe one-time lookup
        NcclApi a;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names)
            if ((a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
        if (a.lib) {
            a.getUniqueId = (int (*)(void*))dlsym(a.lib, "ncclGetUniqueId");
            a.commDestroy = (int (*)(void*))dlsym(a.lib, "ncclCommDestroy");
            a.allGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(a.lib, "ncclAllGather");
            a.errorString = (const char* (*)(int))dlsym(a.lib, "ncclGetErrorString");
            a.commInitRank = dlsym(a.lib, "ncclCommInitRank");
            a.ok = a.getUniqueId && a.commDestroy && a.allGather && a.errorString && a.commInitRank;
        }
        return a;
    }();
    return api;
}
#define NCCL_TRY(expr)                                                                                  \
    do {                                                                                                \
        const int r__ = (expr);                                                                         \
        if (r__ != 0) return fail(RUMI_ERR_CUDA, "%s failed: %s", #expr, nccl_api().errorString(r__));  \
    } while (0)

// Orders a private library stream against a caller stream (NULL = the legacy default stream).
int order_after(cudaStream_t waiter, cudaStream_t producer, cudaEvent_t* ev) {
    if (!*ev) CU_TRY(cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
    CU_TRY(cudaEventRecord(*ev, producer));
    CU_TRY(cudaStreamWaitEvent(waiter, *ev, 0));
    return RUMI_OK;
}
}  // namespace

extern "C" {

int rumi_match_create(rumi_match** out, int device) {
    if (!out) return fail(RUMI_ERR_ARG, "out is NULL");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(RUMI_ERR_CUDA, "no CUDA device: librumi_orb has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(RUMI_ERR_ARG, "device %d out of range", device);
    CU_TRY(cudaSetDevice(device));
    rumi_match* m = new rumi_match();
    m->device = device;
    CU_TRY(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
    const char* mm = getenv("RUMI_MATCH");                       // "popc" / "umma": force one top-2 kernel (tests, A/B)
    if (mm) m->mode = mm[0] == 'p' ? 1 : mm[0] == 'u' ? 3 : 0;
    *out = m;
    return RUMI_OK;
}

void rumi_match_destroy(rumi_match* m) {
    if (!m) return;
    cudaSetDevice(m->device);
    if (m->stream) { cudaStreamSynchronize(m->stream); cudaStreamDestroy(m->stream); }
    cudaFree(m->partial); cudaFree(m->dQ); cudaFree(m->dT); cudaFree(m->dOut); cudaFree(m->tx);
    cudaFree(m->shardSend); cudaFree(m->shardRecv);
    if (m->hStage) cudaFreeHost(m->hStage);
    if (m->comm && m->ownComm && nccl_api().ok) nccl_api().commDestroy(m->comm);
    if (m->evStart) { cudaEventDestroy(m->evStart); cudaEventDestroy(m->evStop); }
    if (m->evOrder) cudaEventDestroy(m->evOrder);
    delete m;
}

int rumi_hamming_top2_device(rumi_match* m, const uint8_t* dQ, int nq, const uint8_t* dT, int nt, int t_base,
                             int32_t* d_idx1, uint16_t* d_d1, uint16_t* d_d2, int sync) {
    if (!m) return fail(RUMI_ERR_ARG, "context is NULL");
    if (nq < 0 || nt < 0) return fail(RUMI_ERR_ARG, "negative sizes");
    if (nq == 0) return RUMI_OK;
    CU_TRY(cudaSetDevice(m->device));
    const int rc = top2_device(m, dQ, nq, dT, nt, t_base, d_idx1, d_d1, d_d2);
    if (rc) return rc;
    if (sync) CU_TRY(cudaStreamSynchronize(m->stream));
    return RUMI_OK;
}

int rumi_hamming_top2(rumi_match* m, const uint8_t* Q, int nq, const uint8_t* T, int nt, int32_t* idx1,
                      uint16_t* d1, uint16_t* d2) {
    if (!m) return fail(RUMI_ERR_ARG, "context is NULL");
    if (nq < 0 || nt < 0) return fail(RUMI_ERR_ARG, "negative sizes");
    if (nq == 0) return RUMI_OK;
    CU_TRY(cudaSetDevice(m->device));
    int rc;
    if ((rc = grow((void**)&m->dQ, &m->qCap, 32 * (size_t)nq))) return rc;
    if ((rc = grow((void**)&m->dT, &m->tCap, 32 * (size_t)std::max(nt, 1)))) return rc;
    if ((rc = grow((void**)&m->dOut, &m->outCap, 8 * (size_t)nq))) return rc;
    int32_t* dIdx = reinterpret_cast<int32_t*>(m->dOut);
    uint16_t* dD1 = reinterpret_cast<uint16_t*>(m->dOut + 4 * (size_t)nq);
    uint16_t* dD2 = dD1 + nq;
    const size_t qBytes = 32 * (size_t)nq, tBytes = 32 * (size_t)nt, oBytes = 8 * (size_t)nq;
    if (qBytes + tBytes + oBytes <= kSmallCallBytes) {
        if ((rc = host_stage(m, qBytes + tBytes + oBytes))) return rc;
        uint8_t* hs = m->hStage;
        std::memcpy(hs, Q, qBytes);
        if (nt > 0) std::memcpy(hs + qBytes, T, tBytes);
        CU_TRY(cudaMemcpyAsync(m->dQ, hs, qBytes, cudaMemcpyHostToDevice, m->stream));
        if (nt > 0) CU_TRY(cudaMemcpyAsync(m->dT, hs + qBytes, tBytes, cudaMemcpyHostToDevice, m->stream));
        if ((rc = top2_device(m, m->dQ, nq, m->dT, nt, 0, dIdx, dD1, dD2))) return rc;
        uint8_t* ho = hs + qBytes + tBytes;
        CU_TRY(cudaMemcpyAsync(ho, m->dOut, oBytes, cudaMemcpyDeviceToHost, m->stream));
        CU_TRY(cudaStreamSynchronize(m->stream));
        std::memcpy(idx1, ho, 4 * (size_t)nq);
        std::memcpy(d1, ho + 4 * (size_t)nq, 2 * (size_t)nq);
        std::memcpy(d2, ho + 6 * (size_t)nq, 2 * (size_t)nq);
        return RUMI_OK;
    }
    CU_TRY(cudaMemcpyAsync(m->dQ, Q, qBytes, cudaMemcpyHostToDevice, m->stream));
    if (nt > 0) CU_TRY(cudaMemcpyAsync(m->dT, T, tBytes, cudaMemcpyHostToDevice, m->stream));
    if ((rc = top2_device(m, m->dQ, nq, m->dT, nt, 0, dIdx, dD1, dD2))) return rc;
    CU_TRY(cudaMemcpyAsync(idx1, dIdx, 4 * (size_t)nq, cudaMemcpyDeviceToHost, m->stream));
    CU_TRY(cudaMemcpyAsync(d1, dD1, 2 * (size_t)nq, cudaMemcpyDeviceToHost, m->stream));
    CU_TRY(cudaMemcpyAsync(d2, dD2, 2 * (size_t)nq, cudaMemcpyDeviceToHost, m->stream));
    CU_TRY(cudaStreamSynchronize(m->stream));
    return RUMI_OK;
}

int rumi_hamming_top2_pairs(rumi_match* m, const uint8_t* Q, int nq, const uint8_t* T, int nt, const int32_t* segs,
                            int nseg, int32_t* idx1, uint16_t* d1, uint16_t* d2) {
    if (!m) return fail(RUMI_ERR_ARG, "context is NULL");
    if (nq < 0 || nt < 0 || nseg < 0) return fail(RUMI_ERR_ARG, "negative sizes");
    if (nq > 0 && (!Q || !idx1 || !d1 || !d2)) return fail(RUMI_ERR_ARG, "NULL buffer");
    if (nseg > 0 && !segs) return fail(RUMI_ERR_ARG, "NULL segments");
    int maxQ = 0;
    for (int s = 0; s < nseg; ++s) {
        const int32_t* g = segs + 4 * (size_t)s;
        if (g[1] < 0 || g[3] < 0 || g[0] < 0 || g[2] < 0 || (long long)g[0] + g[1] > nq || (long long)g[2] + g[3] > nt)
            return fail(RUMI_ERR_ARG, "segment %d outside the descriptor arrays", s);
        if (g[3] >= (1 << 21)) return fail(RUMI_ERR_ARG, "segment %d: more than 2^21 train descriptors", s);
        maxQ = std::max(maxQ, g[1]);
    }
    for (int i = 0; i < nq; ++i) { idx1[i] = -1; d1[i] = 256; d2[i] = 256; }     // rows no segment covers
    if (nq == 0 || nseg == 0 || maxQ == 0) return RUMI_OK;
    CU_TRY(cudaSetDevice(m->device));
    int rc;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t oSeg = al(32 * (size_t)std::max(nt, 1));
    if ((rc = grow((void**)&m->dQ, &m->qCap, 32 * (size_t)nq))) return rc;
    if ((rc = grow((void**)&m->dT, &m->tCap, oSeg + al(16 * (size_t)nseg)))) return rc;
    if ((rc = grow((void**)&m->dOut, &m->outCap, 8 * (size_t)nq))) return rc;
    int32_t* dIdx = reinterpret_cast<int32_t*>(m->dOut);
    uint16_t* dD1 = reinterpret_cast<uint16_t*>(m->dOut + 4 * (size_t)nq);
    uint16_t* dD2 = dD1 + nq;
    CU_TRY(cudaMemcpyAsync(m->dQ, Q, 32 * (size_t)nq, cudaMemcpyHostToDevice, m->stream));
    if (nt > 0) CU_TRY(cudaMemcpyAsync(m->dT, T, 32 * (size_t)nt, cudaMemcpyHostToDevice, m->stream));
    CU_TRY(cudaMemcpyAsync(m->dT + oSeg, segs, 16 * (size_t)nseg, cudaMemcpyHostToDevice, m->stream));
    // rows outside every segment keep (-1, 256, 256): start from the host's initial values
    CU_TRY(cudaMemcpyAsync(dIdx, idx1, 4 * (size_t)nq, cudaMemcpyHostToDevice, m->stream));
    CU_TRY(cudaMemcpyAsync(dD1, d1, 2 * (size_t)nq, cudaMemcpyHostToDevice, m->stream));
    CU_TRY(cudaMemcpyAsync(dD2, d2, 2 * (size_t)nq, cudaMemcpyHostToDevice, m->stream));
    launch_hamming_top2_segments(m->dQ, m->dT, reinterpret_cast<const PairSegment*>(m->dT + oSeg), nseg, maxQ, dIdx, dD1,
                                 dD2, m->stream);
    m->launches += 1;
    m->lastPath = 1;
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(idx1, dIdx, 4 * (size_t)nq, cudaMemcpyDeviceToHost, m->stream));
    CU_TRY(cudaMemcpyAsync(d1, dD1, 2 * (size_t)nq, cudaMemcpyDeviceToHost, m->stream));
    CU_TRY(cudaMemcpyAsync(d2, dD2, 2 * (size_t)nq, cudaMemcpyDeviceToHost, m->stream));
    CU_TRY(cudaStreamSynchronize(m->stream));
    return RUMI_OK;
}

int rumi_hamming_candidates(rumi_match* m, const uint8_t* Q, int nq, const uint8_t* T, int nt, const int32_t* cand_off,
                            const int32_t* cand_idx, uint16_t* dist, int32_t* idx1, uint16_t* d1, int32_t* idx2, uint16_t* d2) {
    if (!m) return fail(RUMI_ERR_ARG, "context is NULL");
    if (nq < 0 || nt < 0) return fail(RUMI_ERR_ARG, "negative sizes");
    if (nq == 0) return RUMI_OK;
    if (!Q || !cand_off || cand_off[0] != 0) return fail(RUMI_ERR_ARG, "NULL buffer / cand_off[0] != 0");
    const bool top2 = idx1 || d1 || idx2 || d2;
    if (top2 && !(idx1 && d1 && idx2 && d2)) return fail(RUMI_ERR_ARG, "the four top-2 outputs go together");
    for (int q = 0; q < nq; ++q) {
        if (cand_off[q + 1] < cand_off[q]) return fail(RUMI_ERR_ARG, "cand_off must be non-decreasing");
        if (top2 && cand_off[q + 1] - cand_off[q] >= 0xFFFF) return fail(RUMI_ERR_ARG, "query %d: more than 65534 candidates", q);
    }
    const int ne = cand_off[nq];
    if (ne > 0 && (!cand_idx || !dist || !T)) return fail(RUMI_ERR_ARG, "NULL candidate buffers");
    for (int k = 0; k < ne; ++k)
        if (cand_idx[k] < 0 || cand_idx[k] >= nt) return fail(RUMI_ERR_ARG, "cand_idx[%d] = %d outside the train set", k, cand_idx[k]);
    CU_TRY(cudaSetDevice(m->device));
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t oT = al(32 * (size_t)nq), oOff = oT + al(32 * (size_t)std::max(nt, 1)), oIdx = oOff + al(4 * ((size_t)nq + 1)),
                 oDist = oIdx + al(4 * (size_t)std::max(ne, 1)), oRes = oDist + al(2 * (size_t)std::max(ne, 1)),
                 need = oRes + al(12 * (size_t)nq);
    int rc = grow((void**)&m->dT, &m->tCap, need);
    if (rc) return rc;
    uint8_t* p = m->dT;
    cudaStream_t s = m->stream;
    const bool small = need <= kSmallCallBytes;          // one packed block up, one block down (see host_stage)
    if (small) {
        if ((rc = host_stage(m, need))) return rc;
        uint8_t* hs = m->hStage;
        std::memcpy(hs, Q, 32 * (size_t)nq);
        if (nt > 0) std::memcpy(hs + oT, T, 32 * (size_t)nt);
        std::memcpy(hs + oOff, cand_off, 4 * ((size_t)nq + 1));
        if (ne > 0) std::memcpy(hs + oIdx, cand_idx, 4 * (size_t)ne);
        CU_TRY(cudaMemcpyAsync(p, hs, oDist, cudaMemcpyHostToDevice, s));
    } else {
        CU_TRY(cudaMemcpyAsync(p, Q, 32 * (size_t)nq, cudaMemcpyHostToDevice, s));
        if (nt > 0) CU_TRY(cudaMemcpyAsync(p + oT, T, 32 * (size_t)nt, cudaMemcpyHostToDevice, s));
        CU_TRY(cudaMemcpyAsync(p + oOff, cand_off, 4 * ((size_t)nq + 1), cudaMemcpyHostToDevice, s));
        if (ne > 0) CU_TRY(cudaMemcpyAsync(p + oIdx, cand_idx, 4 * (size_t)ne, cudaMemcpyHostToDevice, s));
    }
    int32_t* dI1 = reinterpret_cast<int32_t*>(p + oRes);
    int32_t* dI2 = dI1 + nq;
    uint16_t* dD1 = reinterpret_cast<uint16_t*>(dI2 + nq);
    uint16_t* dD2 = dD1 + nq;
    launch_hamming_candidates(p, nq, p + oT, reinterpret_cast<const int32_t*>(p + oOff), reinterpret_cast<const int32_t*>(p + oIdx),
                              reinterpret_cast<uint16_t*>(p + oDist), top2 ? dI1 : nullptr, dD1, dI2, dD2, s);
    m->launches += 1;
    m->lastPath = 1;
    CU_TRY(cudaGetLastError());
    if (small) {
        uint8_t* hs = m->hStage;
        CU_TRY(cudaMemcpyAsync(hs + oDist, p + oDist, (top2 ? need : oRes) - oDist, cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaStreamSynchronize(s));
        if (ne > 0) std::memcpy(dist, hs + oDist, 2 * (size_t)ne);
        if (top2) {
            const uint8_t* r = hs + oRes;
            std::memcpy(idx1, r, 4 * (size_t)nq);
            std::memcpy(idx2, r + 4 * (size_t)nq, 4 * (size_t)nq);
            std::memcpy(d1, r + 8 * (size_t)nq, 2 * (size_t)nq);
            std::memcpy(d2, r + 10 * (size_t)nq, 2 * (size_t)nq);
        }
        return RUMI_OK;
    }
    if (ne > 0) CU_TRY(cudaMemcpyAsync(dist, p + oDist, 2 * (size_t)ne, cudaMemcpyDeviceToHost, s));
    if (top2) {
        CU_TRY(cudaMemcpyAsync(idx1, dI1, 4 * (size_t)nq, cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaMemcpyAsync(idx2, dI2, 4 * (size_t)nq, cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaMemcpyAsync(d1, dD1, 2 * (size_t)nq, cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaMemcpyAsync(d2, dD2, 2 * (size_t)nq, cudaMemcpyDeviceToHost, s));
    }
    CU_TRY(cudaStreamSynchronize(s));
    return RUMI_OK;
}

int rumi_match_timer_start(rumi_match* m) {
    if (!m) return fail(RUMI_ERR_ARG, "context is NULL");
    CU_TRY(cudaSetDevice(m->device));
    if (!m->evStart) { CU_TRY(cudaEventCreate(&m->evStart)); CU_TRY(cudaEventCreate(&m->evStop)); }
    CU_TRY(cudaEventRecord(m->evStart, m->stream));
    return RUMI_OK;
}

int rumi_match_timer_stop(rumi_match* m, float* ms) {
    if (!m || !m->evStart || !ms) return fail(RUMI_ERR_ARG, "timer not started");
    CU_TRY(cudaSetDevice(m->device));
    CU_TRY(cudaEventRecord(m->evStop, m->stream));
    CU_TRY(cudaEventSynchronize(m->evStop));
    CU_TRY(cudaEventElapsedTime(ms, m->evStart, m->evStop));
    return RUMI_OK;
}

int rumi_match_last_path(const rumi_match* m) { return m ? m->lastPath : 0; }

long long rumi_match_launch_count(rumi_match* m, int reset) {
    if (!m) return 0;
    const long long v = m->launches;
    if (reset) m->launches = 0;
    return v;
}

int rumi_top2_pack_device(rumi_match* m, const int32_t* d_idx1, const uint16_t* d_d1, const uint16_t* d_d2, int nq,
                          uint64_t* d_packed, int sync) {
    if (!m) return fail(RUMI_ERR_ARG, "context is NULL");
    CU_TRY(cudaSetDevice(m->device));
    launch_pack_top2(d_idx1, d_d1, d_d2, nq, d_packed, m->stream);
    m->launches += 1;
    CU_TRY(cudaGetLastError());
    if (sync) CU_TRY(cudaStreamSynchronize(m->stream));
    return RUMI_OK;
}

int rumi_top2_merge_device(rumi_match* m, const uint64_t* d_packed, int nshards, int nq, int32_t* d_idx1,
                           uint16_t* d_d1, uint16_t* d_d2, int sync) {
    if (!m) return fail(RUMI_ERR_ARG, "context is NULL");
    if (nshards < 1) return fail(RUMI_ERR_ARG, "nshards < 1");
    CU_TRY(cudaSetDevice(m->device));
    launch_top2_merge(d_packed, nshards, nq, d_idx1, d_d1, d_d2, m->stream);
    m->launches += 1;
    CU_TRY(cudaGetLastError());
    if (sync) CU_TRY(cudaStreamSynchronize(m->stream));
    return RUMI_OK;
}

int rumi_match_wait_stream(rumi_match* m, void* stream) {
    if (!m) return fail(RUMI_ERR_ARG, "context is NULL");
    CU_TRY(cudaSetDevice(m->device));
    return order_after(m->stream, (cudaStream_t)stream, &m->evOrder);
}

int rumi_match_signal_stream(rumi_match* m, void* stream) {
    if (!m) return fail(RUMI_ERR_ARG, "context is NULL");
    CU_TRY(cudaSetDevice(m->device));
    return order_after((cudaStream_t)stream, m->stream, &m->evOrder);
}

int rumi_nccl_unique_id(uint8_t* id128) {
    if (!id128) return fail(RUMI_ERR_ARG, "id is NULL");
    NcclApi& n = nccl_api();
    if (!n.ok) return fail(RUMI_ERR_CUDA, "libnccl.so.2 not found: multi-GPU matching needs NCCL");
    NcclId id;
    NCCL_TRY(n.getUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return RUMI_OK;
}

int rumi_match_comm_init(rumi_match* m, const uint8_t* id128, int rank, int nranks) {
    if (!m || !id128) return fail(RUMI_ERR_ARG, "NULL argument");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(RUMI_ERR_ARG, "rank %d of %d", rank, nranks);
    NcclApi& n = nccl_api();
    if (!n.ok) return fail(RUMI_ERR_CUDA, "libnccl.so.2 not found: multi-GPU matching needs NCCL");
    CU_TRY(cudaSetDevice(m->device));
    if (m->comm && m->ownComm) n.commDestroy(m->comm);
    m->comm = nullptr;
    NcclId id;
    memcpy(&id, id128, sizeof(id));
    typedef int (*InitFn)(void**, int, NcclId, int);
    NCCL_TRY(((InitFn)n.commInitRank)(&m->comm, nranks, id, rank));
    m->ownComm = true; m->rank = rank; m->nranks = nranks;
    return RUMI_OK;
}

int rumi_match_comm_adopt(rumi_match* m, void* nccl_comm, int rank, int nranks) {
    if (!m || !nccl_comm) return fail(RUMI_ERR_ARG, "NULL argument");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(RUMI_ERR_ARG, "rank %d of %d", rank, nranks);
    NcclApi& n = nccl_api();
    if (!n.ok) return fail(RUMI_ERR_CUDA, "libnccl.so.2 not found: multi-GPU matching needs NCCL");
    if (m->comm && m->ownComm) n.commDestroy(m->comm);
    m->comm = nccl_comm; m->ownComm = false; m->rank = rank; m->nranks = nranks;
    return RUMI_OK;
}

int rumi_hamming_top2_sharded(rumi_match* m, const uint8_t* dQ, int nq, const uint8_t* dT_local, int nt_local,
                              int t_base, int32_t* d_idx1, uint16_t* d_d1, uint16_t* d_d2, int sync) {
    if (!m) return fail(RUMI_ERR_ARG, "context is NULL");
    if (nq < 0 || nt_local < 0) return fail(RUMI_ERR_ARG, "negative sizes");
    if (!m->comm) return fail(RUMI_ERR_ARG, "no communicator: call rumi_match_comm_init / rumi_match_comm_adopt first");
    if (nq == 0) return RUMI_OK;
    CU_TRY(cudaSetDevice(m->device));
    int rc, slices = 1;
    if ((rc = grow((void**)&m->shardSend, &m->sendCap, 8 * (size_t)nq))) return rc;
    if ((rc = grow((void**)&m->shardRecv, &m->recvCap, 8 * (size_t)nq * m->nranks))) return rc;
    // local scan -> packed candidates (the slice merge writes the 8-byte form the exchange sends) -> all-gather on the
    // matcher's own stream (no host synchronisation anywhere) -> fold in rank order = ascending train ranges
    if ((rc = top2_partials(m, dQ, nq, dT_local, nt_local, t_base, &slices))) return rc;
    launch_top2_merge_packed(m->partial, slices, nq, m->shardSend, m->stream);
    NCCL_TRY(nccl_api().allGather(m->shardSend, m->shardRecv, (size_t)nq, kNcclUint64, m->comm, m->stream));
    launch_top2_merge(m->shardRecv, m->nranks, nq, d_idx1, d_d1, d_d2, m->stream);
    m->launches += 2;
    CU_TRY(cudaGetLastError());
    if (sync) CU_TRY(cudaStreamSynchronize(m->stream));
    return RUMI_OK;
}

int rumi_stereo_best1(rumi_match* m, const rumi_kp* Lk, const uint8_t* Ld, int nL, const rumi_kp* Rk,
                      const uint8_t* Rd, int nR, const float* scale_factors, int nlevels, int n_rows, float min_d,
                      float max_d, int32_t* best_r, uint16_t* best_dist) {
    if (!m) return fail(RUMI_ERR_ARG, "context is NULL");
    if (nL < 0 || nR < 0 || nR >= (1 << 20) || nlevels < 1 || !scale_factors) return fail(RUMI_ERR_ARG, "bad sizes");
    if (nL == 0) return RUMI_OK;
    for (int i = 0; i < nR; ++i)
        if (Rk[i].octave < 0 || Rk[i].octave >= nlevels) return fail(RUMI_ERR_ARG, "right keypoint %d: bad octave", i);
    CU_TRY(cudaSetDevice(m->device));
    const size_t kb = sizeof(KeyPointRec);
    const size_t need = kb * ((size_t)nL + nR) + 32 * ((size_t)nL + nR) + 4 * (size_t)nlevels + 8 * (size_t)nL + 256;
    int rc = grow((void**)&m->dT, &m->tCap, need);
    if (rc) return rc;
    uint8_t* p = m->dT;                                  // 16-byte aligned carve-out
    uint8_t* dLd = p; p += 32 * (size_t)nL;
    uint8_t* dRd = p; p += 32 * (size_t)std::max(nR, 1);
    KeyPointRec* dLk = (KeyPointRec*)p; p += (kb * nL + 15) & ~(size_t)15;
    KeyPointRec* dRk = (KeyPointRec*)p; p += (kb * std::max(nR, 1) + 15) & ~(size_t)15;
    float* dSf = (float*)p; p += (4 * (size_t)nlevels + 15) & ~(size_t)15;
    int32_t* dBest = (int32_t*)p; p += 4 * (size_t)nL;
    uint16_t* dDist = (uint16_t*)p;
    cudaStream_t s = m->stream;
    CU_TRY(cudaMemcpyAsync(dLd, Ld, 32 * (size_t)nL, cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpyAsync(dLk, Lk, kb * nL, cudaMemcpyHostToDevice, s));
    if (nR > 0) {
        CU_TRY(cudaMemcpyAsync(dRd, Rd, 32 * (size_t)nR, cudaMemcpyHostToDevice, s));
        CU_TRY(cudaMemcpyAsync(dRk, Rk, kb * nR, cudaMemcpyHostToDevice, s));
    }
    CU_TRY(cudaMemcpyAsync(dSf, scale_factors, 4 * (size_t)nlevels, cudaMemcpyHostToDevice, s));
    launch_stereo_best1(dLk, dLd, nL, dRk, dRd, nR, dSf, n_rows, min_d, max_d, dBest, dDist, s);
    m->launches += 1;
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(best_r, dBest, 4 * (size_t)nL, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaMemcpyAsync(best_dist, dDist, 2 * (size_t)nL, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    return RUMI_OK;
}

int rumi_stereo_match(rumi_match* m, rumi_orb* left, rumi_orb* right, const rumi_kp* Lk, const uint8_t* Ld, int nL,
                      const rumi_kp* Rk, const uint8_t* Rd, int nR, float mbf, float mb, float* u_right, float* depth,
                      int* n_matched) {
    if (!m || !left || !right) return fail(RUMI_ERR_ARG, "NULL handle");
    if (!left->coef || !right->coef || left->W != right->W || left->H != right->H || left->nlevels != right->nlevels)
        return fail(RUMI_ERR_ARG, "both extractors must have processed an image of the same shape");
    if (left->device != m->device || right->device != m->device) return fail(RUMI_ERR_ARG, "handles on different devices");
    if (nL < 0 || nR < 0 || nR >= (1 << 20)) return fail(RUMI_ERR_ARG, "bad sizes");
    if (n_matched) *n_matched = 0;
    if (nL == 0) return RUMI_OK;
    const int nlevels = left->nlevels;
    for (int i = 0; i < nR; ++i)
        if (Rk[i].octave < 0 || Rk[i].octave >= nlevels) return fail(RUMI_ERR_ARG, "right keypoint %d: bad octave", i);
    for (int i = 0; i < nL; ++i)
        if (Lk[i].octave < 0 || Lk[i].octave >= nlevels) return fail(RUMI_ERR_ARG, "left keypoint %d: bad octave", i);
    CU_TRY(cudaSetDevice(m->device));
    const size_t kb = sizeof(KeyPointRec);
    const size_t need = (kb + 32) * ((size_t)nL + nR) + 4 * (size_t)nlevels + 20 * (size_t)nL + 512;
    int rc = grow((void**)&m->dT, &m->tCap, need);
    if (rc) return rc;
    uint8_t* p = m->dT;
    auto take = [&](size_t bytes) { uint8_t* r = p; p += (bytes + 15) & ~(size_t)15; return r; };
    uint8_t* dLd = take(32 * (size_t)nL);
    uint8_t* dRd = take(32 * (size_t)std::max(nR, 1));
    KeyPointRec* dLk = (KeyPointRec*)take(kb * nL);
    KeyPointRec* dRk = (KeyPointRec*)take(kb * std::max(nR, 1));
    float* dSf = (float*)take(4 * (size_t)nlevels);
    int32_t* dBest = (int32_t*)take(4 * (size_t)nL);
    uint16_t* dDist = (uint16_t*)take(2 * (size_t)nL);
    float* dU = (float*)take(4 * (size_t)nL);
    float* dD = (float*)take(4 * (size_t)nL);
    int* dSad = (int*)take(4 * (size_t)nL);
    cudaStream_t s = m->stream;
    // the pyramids were produced on the extractors' own streams
    CU_TRY(cudaStreamSynchronize(left->ws[left->lastWs].stream));
    CU_TRY(cudaStreamSynchronize(right->ws[right->lastWs].stream));
    // inputs (descriptors, key points, scale table) are contiguous on the device from dLd to the end of dSf, results from dU
    // to the end of dSad: per-frame sized calls pack them into one pinned block -- one copy up, one down (see host_stage)
    const size_t inBytes = (size_t)((uint8_t*)dBest - m->dT), outOff = (size_t)((uint8_t*)dU - m->dT),
                 outBytes = (size_t)(p - (uint8_t*)dU);
    const bool small = need <= kSmallCallBytes;
    if (small) {
        if ((rc = host_stage(m, need))) return rc;
        uint8_t* hs = m->hStage;
        std::memcpy(hs + ((uint8_t*)dLd - m->dT), Ld, 32 * (size_t)nL);
        std::memcpy(hs + ((uint8_t*)dLk - m->dT), Lk, kb * nL);
        if (nR > 0) {
            std::memcpy(hs + ((uint8_t*)dRd - m->dT), Rd, 32 * (size_t)nR);
            std::memcpy(hs + ((uint8_t*)dRk - m->dT), Rk, kb * nR);
        }
        std::memcpy(hs + ((uint8_t*)dSf - m->dT), left->tables.scale.data(), 4 * (size_t)nlevels);
        CU_TRY(cudaMemcpyAsync(m->dT, hs, inBytes, cudaMemcpyHostToDevice, s));
    } else {
        CU_TRY(cudaMemcpyAsync(dLd, Ld, 32 * (size_t)nL, cudaMemcpyHostToDevice, s));
        CU_TRY(cudaMemcpyAsync(dLk, Lk, kb * nL, cudaMemcpyHostToDevice, s));
        if (nR > 0) {
            CU_TRY(cudaMemcpyAsync(dRd, Rd, 32 * (size_t)nR, cudaMemcpyHostToDevice, s));
            CU_TRY(cudaMemcpyAsync(dRk, Rk, kb * nR, cudaMemcpyHostToDevice, s));
        }
        CU_TRY(cudaMemcpyAsync(dSf, left->tables.scale.data(), 4 * (size_t)nlevels, cudaMemcpyHostToDevice, s));
    }
    const float minD = 0.f, maxD = mbf / mb;                                   // Frame.cc:856-858
    launch_stereo_best1(dLk, dLd, nL, dRk, dRd, nR, dSf, left->H, minD, maxD, dBest, dDist, s);
    StereoRefineArgs ra;
    for (int l = 0; l < nlevels; ++l) {
        ra.left[l] = internal_view(left, left->ws[left->lastWs].pyr, l);
        ra.right[l] = internal_view(right, right->ws[right->lastWs].pyr, l);
        ra.scale[l] = left->tables.scale[l];
        ra.invScale[l] = left->tables.invScale[l];
    }
    ra.Lk = dLk; ra.Rk = dRk; ra.bestR = dBest; ra.bestDist = dDist; ra.nL = nL;
    ra.minD = minD; ra.maxD = maxD; ra.mbf = mbf; ra.uRight = dU; ra.depth = dD; ra.sad = dSad;
    launch_stereo_refine(ra, s);
    m->launches += 2;
    CU_TRY(cudaGetLastError());
    std::vector<int> sad(nL);
    if (small) {
        uint8_t* ho = m->hStage + outOff;
        CU_TRY(cudaMemcpyAsync(ho, dU, outBytes, cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaStreamSynchronize(s));
        std::memcpy(u_right, ho, 4 * (size_t)nL);
        std::memcpy(depth, ho + ((uint8_t*)dD - (uint8_t*)dU), 4 * (size_t)nL);
        std::memcpy(sad.data(), ho + ((uint8_t*)dSad - (uint8_t*)dU), 4 * (size_t)nL);
    } else {
        CU_TRY(cudaMemcpyAsync(u_right, dU, 4 * (size_t)nL, cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaMemcpyAsync(depth, dD, 4 * (size_t)nL, cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaMemcpyAsync(sad.data(), dSad, 4 * (size_t)nL, cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaStreamSynchronize(s));
    }
    // median-based outlier cut (Frame.cc:973-984): sequential host logic, kept as in the reference
    std::vector<std::pair<int, int>> vDistIdx;
    for (int i = 0; i < nL; ++i)
        if (sad[i] >= 0) vDistIdx.push_back(std::make_pair(sad[i], i));
    int kept = (int)vDistIdx.size();
    if (!vDistIdx.empty()) {
        std::sort(vDistIdx.begin(), vDistIdx.end());
        const float median = (float)vDistIdx[vDistIdx.size() / 2].first;
        const float thDist = 1.5f * 1.4f * median;
        for (int i = (int)vDistIdx.size() - 1; i >= 0; --i) {
            if ((float)vDistIdx[i].first < thDist) break;
            u_right[vDistIdx[i].second] = -1;
            depth[vDistIdx[i].second] = -1;
            --kept;
        }
    }
    if (n_matched) *n_matched = kept;
    return RUMI_OK;
}

int rumi_descriptor_distance(const uint8_t* a, const uint8_t* b) {
    int d = 0;
    for (int i = 0; i < 4; ++i) {
        uint64_t x, y;
        memcpy(&x, a + 8 * i, 8);
        memcpy(&y, b + 8 * i, 8);
        d += __builtin_popcountll(x ^ y);
    }
    return d;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// bag of words
struct rumi_vocab {
    int device = 0, k = 0, L = 0, nnodes = 0, nwords = 0;
    cudaStream_t stream = nullptr;
    uint8_t* blob = nullptr;           // every tree array in one allocation
    rumi::BowTreeView view;
    uint8_t* scratch = nullptr;        // host-call staging: descriptors + outputs
    size_t scratchCap = 0;
    uint8_t* hStage = nullptr;         // pinned mirror of `scratch` for per-frame sized calls (one copy up, one down)
    size_t hStageCap = 0;
    cudaEvent_t evOrder = nullptr;
    long long launches = 0;
};

extern "C" {

int rumi_vocab_create(rumi_vocab** out, int device, int k, int L, int nnodes, const int32_t* parent,
                      const uint8_t* is_leaf, const uint8_t* desc, const double* weight) {
    if (!out) return fail(RUMI_ERR_ARG, "out is NULL");
    *out = nullptr;
    if (nnodes < 2 || !parent || !is_leaf || !desc || !weight || k < 1 || L < 1) return fail(RUMI_ERR_ARG, "bad vocabulary");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(RUMI_ERR_CUDA, "no CUDA device: librumi_orb has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(RUMI_ERR_ARG, "device %d out of range", device);
    // children in the reference's push_back order: node ids ascending under each parent (loadFromTextFile :1390)
    std::vector<int32_t> cnt(nnodes, 0), start(nnodes, 0), ids(std::max(nnodes - 1, 1), 0), word(nnodes, 0);
    for (int i = 1; i < nnodes; ++i) {
        if (parent[i] < 0 || parent[i] >= i) return fail(RUMI_ERR_ARG, "node %d: parent %d must precede it", i, parent[i]);
        ++cnt[parent[i]];
    }
    if (cnt[0] == 0) return fail(RUMI_ERR_ARG, "root has no children");
    {   // limits of the descent kernel: child position travels in 16 bits, at most 63 levels
        std::vector<int32_t> depth(nnodes, 0);
        for (int i = 1; i < nnodes; ++i) {
            depth[i] = depth[parent[i]] + 1;
            if (depth[i] > 63) return fail(RUMI_ERR_ARG, "vocabulary deeper than 63 levels");
            if (cnt[i] > 65535) return fail(RUMI_ERR_ARG, "node %d has more than 65535 children", i);
        }
        if (cnt[0] > 65535) return fail(RUMI_ERR_ARG, "root has more than 65535 children");
    }
    int run = 0;
    for (int i = 0; i < nnodes; ++i) { start[i] = run; run += cnt[i]; }
    std::vector<int32_t> fill(start);
    for (int i = 1; i < nnodes; ++i) ids[fill[parent[i]]++] = i;
    int nwords = 0;
    for (int i = 1; i < nnodes; ++i)
        if (is_leaf[i]) word[i] = nwords++;                       // :1405-1411
    CU_TRY(cudaSetDevice(device));
    rumi_vocab* v = new rumi_vocab();
    v->device = device; v->k = k; v->L = L; v->nnodes = nnodes; v->nwords = nwords;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t oDesc = 0, oStart = al(32 * (size_t)nnodes), oCnt = oStart + al(4 * (size_t)nnodes),
                 oIds = oCnt + al(4 * (size_t)nnodes), oWord = oIds + al(4 * ids.size()),
                 oW = oWord + al(4 * (size_t)nnodes), total = oW + al(8 * (size_t)nnodes);
    if (cudaMalloc(&v->blob, total) != cudaSuccess) { delete v; return fail(RUMI_ERR_CUDA, "cudaMalloc(%zu) failed", total); }
    cudaMemcpy(v->blob + oDesc, desc, 32 * (size_t)nnodes, cudaMemcpyHostToDevice);
    cudaMemcpy(v->blob + oStart, start.data(), 4 * (size_t)nnodes, cudaMemcpyHostToDevice);
    cudaMemcpy(v->blob + oCnt, cnt.data(), 4 * (size_t)nnodes, cudaMemcpyHostToDevice);
    cudaMemcpy(v->blob + oIds, ids.data(), 4 * ids.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(v->blob + oWord, word.data(), 4 * (size_t)nnodes, cudaMemcpyHostToDevice);
    cudaMemcpy(v->blob + oW, weight, 8 * (size_t)nnodes, cudaMemcpyHostToDevice);
    if (cudaDeviceSynchronize() != cudaSuccess || cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking) != cudaSuccess) {
        cudaFree(v->blob); delete v;
        return fail(RUMI_ERR_CUDA, "vocabulary upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    v->view.k = k; v->view.L = L; v->view.nnodes = nnodes;
    v->view.desc = v->blob + oDesc;
    v->view.childStart = reinterpret_cast<const int32_t*>(v->blob + oStart);
    v->view.childCount = reinterpret_cast<const int32_t*>(v->blob + oCnt);
    v->view.childIds = reinterpret_cast<const int32_t*>(v->blob + oIds);
    v->view.wordId = reinterpret_cast<const int32_t*>(v->blob + oWord);
    v->view.weight = reinterpret_cast<const double*>(v->blob + oW);
    *out = v;
    return RUMI_OK;
}

void rumi_vocab_destroy(rumi_vocab* v) {
    if (!v) return;
    cudaSetDevice(v->device);
    if (v->stream) { cudaStreamSynchronize(v->stream); cudaStreamDestroy(v->stream); }
    cudaFree(v->blob); cudaFree(v->scratch);
    if (v->hStage) cudaFreeHost(v->hStage);
    if (v->evOrder) cudaEventDestroy(v->evOrder);
    delete v;
}

int rumi_vocab_words(const rumi_vocab* v) { return v ? v->nwords : 0; }

int rumi_vocab_wait_stream(rumi_vocab* v, void* stream) {
    if (!v) return fail(RUMI_ERR_ARG, "vocabulary is NULL");
    CU_TRY(cudaSetDevice(v->device));
    return order_after(v->stream, (cudaStream_t)stream, &v->evOrder);
}

int rumi_vocab_signal_stream(rumi_vocab* v, void* stream) {
    if (!v) return fail(RUMI_ERR_ARG, "vocabulary is NULL");
    CU_TRY(cudaSetDevice(v->device));
    return order_after((cudaStream_t)stream, v->stream, &v->evOrder);
}

long long rumi_vocab_launch_count(rumi_vocab* v, int reset) {
    if (!v) return 0;
    const long long n = v->launches;
    if (reset) v->launches = 0;
    return n;
}

int rumi_bow_transform_device(rumi_vocab* v, const uint8_t* d_desc, int n, int levelsup, int32_t* d_word_id,
                              double* d_weight, int32_t* d_node_id, int sync) {
    if (!v) return fail(RUMI_ERR_ARG, "vocabulary is NULL");
    if (n < 0 || levelsup < 0) return fail(RUMI_ERR_ARG, "negative argument");
    if (n == 0) return RUMI_OK;
    if (!d_desc || !d_word_id || !d_weight || !d_node_id || ((uintptr_t)d_desc & 15))
        return fail(RUMI_ERR_ARG, "NULL / unaligned device buffer");
    CU_TRY(cudaSetDevice(v->device));
    launch_bow_transform(v->view, d_desc, n, levelsup, d_word_id, d_weight, d_node_id, v->stream);
    v->launches += 1;
    CU_TRY(cudaGetLastError());
    if (sync) CU_TRY(cudaStreamSynchronize(v->stream));
    return RUMI_OK;
}

int rumi_bow_transform(rumi_vocab* v, const uint8_t* desc, int n, int levelsup, int32_t* word_id, double* weight,
                       int32_t* node_id) {
    if (!v) return fail(RUMI_ERR_ARG, "vocabulary is NULL");
    if (n < 0 || levelsup < 0) return fail(RUMI_ERR_ARG, "negative argument");
    if (n == 0) return RUMI_OK;
    if (!desc || !word_id || !weight || !node_id) return fail(RUMI_ERR_ARG, "NULL buffer");
    CU_TRY(cudaSetDevice(v->device));
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t oW = al(32 * (size_t)n), oWord = oW + al(8 * (size_t)n), oNode = oWord + al(4 * (size_t)n),
                 need = oNode + al(4 * (size_t)n);
    int rc = grow((void**)&v->scratch, &v->scratchCap, need);
    if (rc) return rc;
    double* dW = reinterpret_cast<double*>(v->scratch + oW);
    int32_t* dWord = reinterpret_cast<int32_t*>(v->scratch + oWord);
    int32_t* dNode = reinterpret_cast<int32_t*>(v->scratch + oNode);
    if (need <= kSmallCallBytes) {                       // Frame::ComputeBoW of one frame: one round trip (see host_stage)
        if (need > v->hStageCap) {
            if (v->hStage) cudaFreeHost(v->hStage);
            v->hStage = nullptr; v->hStageCap = 0;
            CU_TRY(cudaHostAlloc((void**)&v->hStage, need, cudaHostAllocDefault));
            v->hStageCap = need;
        }
        std::memcpy(v->hStage, desc, 32 * (size_t)n);
        CU_TRY(cudaMemcpyAsync(v->scratch, v->hStage, 32 * (size_t)n, cudaMemcpyHostToDevice, v->stream));
        if ((rc = rumi_bow_transform_device(v, v->scratch, n, levelsup, dWord, dW, dNode, 0))) return rc;
        CU_TRY(cudaMemcpyAsync(v->hStage + oW, v->scratch + oW, need - oW, cudaMemcpyDeviceToHost, v->stream));
        CU_TRY(cudaStreamSynchronize(v->stream));
        std::memcpy(weight, v->hStage + oW, 8 * (size_t)n);
        std::memcpy(word_id, v->hStage + oWord, 4 * (size_t)n);
        std::memcpy(node_id, v->hStage + oNode, 4 * (size_t)n);
        return RUMI_OK;
    }
    CU_TRY(cudaMemcpyAsync(v->scratch, desc, 32 * (size_t)n, cudaMemcpyHostToDevice, v->stream));
    if ((rc = rumi_bow_transform_device(v, v->scratch, n, levelsup, dWord, dW, dNode, 0))) return rc;
    CU_TRY(cudaMemcpyAsync(word_id, dWord, 4 * (size_t)n, cudaMemcpyDeviceToHost, v->stream));
    CU_TRY(cudaMemcpyAsync(weight, dW, 8 * (size_t)n, cudaMemcpyDeviceToHost, v->stream));
    CU_TRY(cudaMemcpyAsync(node_id, dNode, 4 * (size_t)n, cudaMemcpyDeviceToHost, v->stream));
    CU_TRY(cudaStreamSynchronize(v->stream));
    return RUMI_OK;
}

int rumi_bow_node_distances(rumi_match* m, const uint8_t* descA, int nA, const uint8_t* descB, int nB,
                            const int32_t* a_idx, int n_a_idx, const int32_t* b_idx, int n_b_idx,
                            const int32_t* segs, int nseg, uint16_t* dist, long long ndist) {
    if (!m) return fail(RUMI_ERR_ARG, "context is NULL");
    if (nA < 0 || nB < 0 || nseg < 0 || ndist < 0 || n_a_idx < 0 || n_b_idx < 0) return fail(RUMI_ERR_ARG, "negative sizes");
    if (nseg == 0 || ndist == 0) return RUMI_OK;
    if (!descA || !descB || !a_idx || !b_idx || !segs || !dist) return fail(RUMI_ERR_ARG, "NULL buffer");
    // validate the segment table on the host: the kernel trusts it
    for (int s = 0; s < nseg; ++s) {
        const int32_t* g = segs + 5 * (size_t)s;
        if (g[0] < 0 || g[1] < 0 || g[2] < 0 || g[3] < 0 || g[4] < 0 || (long long)g[0] + g[1] > n_a_idx ||
            (long long)g[2] + g[3] > n_b_idx || (long long)g[4] + (long long)g[1] * g[3] > ndist)
            return fail(RUMI_ERR_ARG, "segment %d out of range", s);
    }
    for (int i = 0; i < n_a_idx; ++i) if (a_idx[i] < 0 || a_idx[i] >= nA) return fail(RUMI_ERR_ARG, "a_idx[%d] out of range", i);
    for (int i = 0; i < n_b_idx; ++i) if (b_idx[i] < 0 || b_idx[i] >= nB) return fail(RUMI_ERR_ARG, "b_idx[%d] out of range", i);
    CU_TRY(cudaSetDevice(m->device));
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t oB = al(32 * (size_t)nA), oAi = oB + al(32 * (size_t)nB), oBi = oAi + al(4 * (size_t)n_a_idx),
                 oSeg = oBi + al(4 * (size_t)n_b_idx), oDist = oSeg + al(20 * (size_t)nseg),
                 need = oDist + al(2 * (size_t)ndist);
    int rc = grow((void**)&m->dT, &m->tCap, need);
    if (rc) return rc;
    uint8_t* p = m->dT;
    const bool small = need <= kSmallCallBytes;          // one key-frame pair: one copy up, one down (see host_stage)
    if (small) {
        if ((rc = host_stage(m, need))) return rc;
        uint8_t* hs = m->hStage;
        std::memcpy(hs, descA, 32 * (size_t)nA);
        std::memcpy(hs + oB, descB, 32 * (size_t)nB);
        std::memcpy(hs + oAi, a_idx, 4 * (size_t)n_a_idx);
        std::memcpy(hs + oBi, b_idx, 4 * (size_t)n_b_idx);
        std::memcpy(hs + oSeg, segs, 20 * (size_t)nseg);
        CU_TRY(cudaMemcpyAsync(p, hs, oDist, cudaMemcpyHostToDevice, m->stream));
    } else {
    CU_TRY(cudaMemcpyAsync(p, descA, 32 * (size_t)nA, cudaMemcpyHostToDevice, m->stream));
    CU_TRY(cudaMemcpyAsync(p + oB, descB, 32 * (size_t)nB, cudaMemcpyHostToDevice, m->stream));
    CU_TRY(cudaMemcpyAsync(p + oAi, a_idx, 4 * (size_t)n_a_idx, cudaMemcpyHostToDevice, m->stream));
    CU_TRY(cudaMemcpyAsync(p + oBi, b_idx, 4 * (size_t)n_b_idx, cudaMemcpyHostToDevice, m->stream));
    CU_TRY(cudaMemcpyAsync(p + oSeg, segs, 20 * (size_t)nseg, cudaMemcpyHostToDevice, m->stream));
    }
    launch_bow_node_distances(p, p + oB, reinterpret_cast<const int32_t*>(p + oAi), reinterpret_cast<const int32_t*>(p + oBi),
                              reinterpret_cast<const BowSegment*>(p + oSeg), nseg, reinterpret_cast<uint16_t*>(p + oDist),
                              m->stream);
    m->launches += 1;
    CU_TRY(cudaGetLastError());
    if (small) {
        CU_TRY(cudaMemcpyAsync(m->hStage + oDist, p + oDist, 2 * (size_t)ndist, cudaMemcpyDeviceToHost, m->stream));
        CU_TRY(cudaStreamSynchronize(m->stream));
        std::memcpy(dist, m->hStage + oDist, 2 * (size_t)ndist);
        return RUMI_OK;
    }
    CU_TRY(cudaMemcpyAsync(dist, p + oDist, 2 * (size_t)ndist, cudaMemcpyDeviceToHost, m->stream));
    CU_TRY(cudaStreamSynchronize(m->stream));
    return RUMI_OK;
}

int rumi_distinctive_descriptors(rumi_match* m, const uint8_t* desc, const int32_t* offsets, int npoints,
                                 int32_t* best_idx, int32_t* best_median) {
    if (!m) return fail(RUMI_ERR_ARG, "context is NULL");
    if (npoints < 0) return fail(RUMI_ERR_ARG, "negative sizes");
    if (npoints == 0) return RUMI_OK;
    if (!offsets || !best_idx || !best_median) return fail(RUMI_ERR_ARG, "NULL buffer");
    if (offsets[0] != 0) return fail(RUMI_ERR_ARG, "offsets[0] must be 0");
    for (int p = 0; p < npoints; ++p)
        if (offsets[p + 1] < offsets[p]) return fail(RUMI_ERR_ARG, "offsets must be non-decreasing");
    const int total = offsets[npoints];
    if (total > 0 && !desc) return fail(RUMI_ERR_ARG, "NULL descriptors");
    CU_TRY(cudaSetDevice(m->device));
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t oOff = al(32 * (size_t)std::max(total, 1)), oIdx = oOff + al(4 * ((size_t)npoints + 1)),
                 oMed = oIdx + al(4 * (size_t)npoints), need = oMed + al(4 * (size_t)npoints);
    int rc = grow((void**)&m->dT, &m->tCap, need);
    if (rc) return rc;
    uint8_t* p = m->dT;
    if (total > 0) CU_TRY(cudaMemcpyAsync(p, desc, 32 * (size_t)total, cudaMemcpyHostToDevice, m->stream));
    CU_TRY(cudaMemcpyAsync(p + oOff, offsets, 4 * ((size_t)npoints + 1), cudaMemcpyHostToDevice, m->stream));
    launch_distinctive(p, reinterpret_cast<const int32_t*>(p + oOff), npoints, reinterpret_cast<int32_t*>(p + oIdx),
                       reinterpret_cast<int32_t*>(p + oMed), m->stream);
    m->launches += 1;
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(best_idx, p + oIdx, 4 * (size_t)npoints, cudaMemcpyDeviceToHost, m->stream));
    CU_TRY(cudaMemcpyAsync(best_median, p + oMed, 4 * (size_t)npoints, cudaMemcpyDeviceToHost, m->stream));
    CU_TRY(cudaStreamSynchronize(m->stream));
    return RUMI_OK;
}

}  // extern "C"

// ===================================================================================================================
// Sparse pyramidal Lucas-Kanade flow (KFDSample::Step, R/lib_src/KFDSample.cc:131-132) -- SURVEY.md 8f rank 4
// ===================================================================================================================
struct rumi_flow {
    int device = 0;
    int win = 31, maxLevelReq = 2, maxCount = 20;
    double eps2 = 0.0009;
    float minEig = 1e-4f;
    cudaStream_t stream = nullptr;
    cudaEvent_t evStart = nullptr, evStop = nullptr, evDone = nullptr;
    int w = 0, h = 0, levels = 0;                    // geometry of the frames currently held
    int lw[kFlowMaxLevels], lh[kFlowMaxLevels], lstride[kFlowMaxLevels];
    size_t loff[kFlowMaxLevels], doff[kFlowMaxLevels], imgBytes = 0, derivBytes = 0;
    uint8_t* img[2] = {nullptr, nullptr}; size_t imgCap[2] = {0, 0};    // two pyramids: previous / next, swapped on advance
    uint8_t* deriv = nullptr; size_t derivCap = 0;   // Scharr (dx, dy) of the previous frame's pyramid
    uint8_t* pts = nullptr; size_t ptsCap = 0;       // prevPts | nextPts | err | status
    uint8_t* hpts = nullptr; size_t hptsCap = 0;     // pinned mirror of `pts`: one H2D and ONE D2H per call
    int prevSlot = 0;
    bool havePrev = false;
    long long launches = 0;
};

namespace {

void flow_geometry(rumi_flow* f, int w, int h) {
    // buildOpticalFlowPyramid: stop at the last level whose NEXT size would not exceed the window.
    f->w = w; f->h = h;
    int sw = w, sh = h, level = 0;
    size_t off = 0, doff = 0;
    for (;; ++level) {
        f->lw[level] = sw; f->lh[level] = sh; f->lstride[level] = (sw + 15) & ~15;
        f->loff[level] = off; f->doff[level] = doff;
        off += ((size_t)f->lstride[level] * sh + 255) & ~(size_t)255;
        doff += ((size_t)sw * sh * 4 + 255) & ~(size_t)255;
        sw = (sw + 1) / 2; sh = (sh + 1) / 2;
        if (level == f->maxLevelReq || sw <= f->win || sh <= f->win) break;
    }
    f->levels = level + 1;
    f->imgBytes = off; f->derivBytes = doff;
}

FlowPyramidView flow_view(const rumi_flow* f, int slot) {
    FlowPyramidView v{};
    for (int l = 0; l < f->levels; ++l) {
        v.ptr[l] = f->img[slot] + f->loff[l];
        v.w[l] = f->lw[l]; v.h[l] = f->lh[l]; v.stride[l] = f->lstride[l];
    }
    return v;
}

FlowDerivView flow_deriv_view(const rumi_flow* f) {
    FlowDerivView v{};
    for (int l = 0; l < f->levels; ++l) v.ptr[l] = reinterpret_cast<short2*>(f->deriv + f->doff[l]);
    return v;
}

int flow_upload(rumi_flow* f, int slot, const uint8_t* img, size_t stride) {
    int rc = grow((void**)&f->img[slot], &f->imgCap[slot], f->imgBytes);
    if (rc) return rc;
    CU_TRY(cudaMemcpy2DAsync(f->img[slot], f->lstride[0], img, stride, f->w, f->h, cudaMemcpyDefault, f->stream));
    for (int l = 1; l < f->levels; ++l) {
        launch_flow_pyrdown(f->img[slot] + f->loff[l - 1], f->lw[l - 1], f->lh[l - 1], f->lstride[l - 1],
                            f->img[slot] + f->loff[l], f->lstride[l], f->stream);
        f->launches += 1;
    }
    CU_TRY(cudaGetLastError());
    return RUMI_OK;
}

int flow_derivatives(rumi_flow* f, int slot) {
    int rc = grow((void**)&f->deriv, &f->derivCap, f->derivBytes);
    if (rc) return rc;
    launch_flow_scharr(flow_view(f, slot), flow_deriv_view(f), f->levels, f->stream);
    f->launches += 1;
    CU_TRY(cudaGetLastError());
    return RUMI_OK;
}

}  // namespace

extern "C" {

int rumi_flow_create(rumi_flow** out, int device, int win, int max_level, int max_count, double epsilon,
                     float min_eig_threshold) {
    if (!out) return fail(RUMI_ERR_ARG, "out is NULL");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(RUMI_ERR_CUDA, "no CUDA device: librumi_orb has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(RUMI_ERR_ARG, "device %d out of range", device);
    if (win != 31 && win != 21 && win != 15)
        return fail(RUMI_ERR_ARG, "window %d not built (KFDSample uses 31; 21 and 15 are also instantiated)", win);
    if (max_level < 0 || max_level >= kFlowMaxLevels) return fail(RUMI_ERR_ARG, "maxLevel must be in [0, %d]", kFlowMaxLevels - 1);
    CU_TRY(cudaSetDevice(device));
    rumi_flow* f = new rumi_flow();
    f->device = device;
    f->win = win;
    f->maxLevelReq = max_level;
    f->maxCount = std::min(std::max(max_count, 0), 100);           // calcOpticalFlowPyrLK clamps the criteria
    const double e = std::min(std::max(epsilon, 0.), 10.);
    f->eps2 = e * e;
    f->minEig = min_eig_threshold;
    CU_TRY(cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking));
    CU_TRY(cudaEventCreateWithFlags(&f->evDone, cudaEventDisableTiming));
    *out = f;
    return RUMI_OK;
}

void rumi_flow_destroy(rumi_flow* f) {
    if (!f) return;
    cudaSetDevice(f->device);
    if (f->stream) { cudaStreamSynchronize(f->stream); cudaStreamDestroy(f->stream); }
    cudaFree(f->img[0]); cudaFree(f->img[1]); cudaFree(f->deriv); cudaFree(f->pts);
    if (f->hpts) cudaFreeHost(f->hpts);
    if (f->evStart) { cudaEventDestroy(f->evStart); cudaEventDestroy(f->evStop); }
    if (f->evDone) cudaEventDestroy(f->evDone);
    delete f;
}

int rumi_flow_levels(const rumi_flow* f) { return f ? f->levels : 0; }
long long rumi_flow_launches(const rumi_flow* f) { return f ? f->launches : 0; }

int rumi_flow_set_prev(rumi_flow* f, const uint8_t* img, int w, int h, size_t stride) {
    if (!f) return fail(RUMI_ERR_ARG, "context is NULL");
    if (!img || w <= 0 || h <= 0) return fail(RUMI_ERR_EMPTY, "empty image");
    if (stride < (size_t)w) return fail(RUMI_ERR_ARG, "stride smaller than the width");
    CU_TRY(cudaSetDevice(f->device));
    flow_geometry(f, w, h);
    f->havePrev = false;
    int rc = flow_upload(f, f->prevSlot, img, stride);
    if (rc) return rc;
    if ((rc = flow_derivatives(f, f->prevSlot))) return rc;
    CU_TRY(cudaStreamSynchronize(f->stream));
    f->havePrev = true;
    return RUMI_OK;
}

int rumi_flow_track_next(rumi_flow* f, const uint8_t* img, size_t stride, const float* prev_pts, int n,
                         float* next_pts, uint8_t* status, float* err, int advance) {
    if (!f) return fail(RUMI_ERR_ARG, "context is NULL");
    if (!f->havePrev) return fail(RUMI_ERR_ARG, "no previous frame: call rumi_flow_set_prev first");
    if (!img) return fail(RUMI_ERR_EMPTY, "empty image");
    if (stride < (size_t)f->w) return fail(RUMI_ERR_ARG, "stride smaller than the width");
    if (n < 0) return fail(RUMI_ERR_ARG, "negative point count");
    if (n > 0 && (!prev_pts || !next_pts || !status)) return fail(RUMI_ERR_ARG, "NULL point buffers");
    CU_TRY(cudaSetDevice(f->device));
    const int nextSlot = f->prevSlot ^ 1;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t oNext = al(8 * (size_t)n), oErr = oNext + al(8 * (size_t)n), oSt = oErr + al(4 * (size_t)n),
                 total = oSt + al((size_t)n);
    int rc;
    if (n > 0) {
        if ((rc = grow((void**)&f->pts, &f->ptsCap, total))) return rc;
        if (total > f->hptsCap) {
            if (f->hpts) cudaFreeHost(f->hpts);
            f->hpts = nullptr; f->hptsCap = 0;
            CU_TRY(cudaHostAlloc((void**)&f->hpts, total, cudaHostAllocDefault));
            f->hptsCap = total;
        }
        memcpy(f->hpts, prev_pts, 8 * (size_t)n);
        CU_TRY(cudaMemcpyAsync(f->pts, f->hpts, 8 * (size_t)n, cudaMemcpyHostToDevice, f->stream));
    }
    if ((rc = flow_upload(f, nextSlot, img, stride))) return rc;
    if (n > 0) {
        FlowTrackArgs a{};
        a.I = flow_view(f, f->prevSlot);
        a.J = flow_view(f, nextSlot);
        a.D = flow_deriv_view(f);
        a.maxLevel = f->levels - 1; a.maxCount = f->maxCount; a.eps2 = f->eps2; a.minEig = f->minEig;
        a.prevPts = reinterpret_cast<const float2*>(f->pts);
        a.nextPts = reinterpret_cast<float2*>(f->pts + oNext);
        a.err = reinterpret_cast<float*>(f->pts + oErr);
        a.status = f->pts + oSt;
        a.n = n;
        if (!launch_flow_track(a, f->win, f->stream)) return fail(RUMI_ERR_ARG, "window %d not built", f->win);
        f->launches += 1;
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaMemcpyAsync(f->hpts + oNext, f->pts + oNext, oSt + (size_t)n - oNext, cudaMemcpyDeviceToHost,
                               f->stream));
    }
    CU_TRY(cudaEventRecord(f->evDone, f->stream));   // the caller's results are complete here
    if (advance) {                                   // imprvs = imnext.clone()  (KFDSample.cc:169)
        f->prevSlot = nextSlot;                      // the new previous frame's derivatives are computed behind the
        if ((rc = flow_derivatives(f, f->prevSlot))) return rc;   // caller's back; the next call queues after them
    }
    CU_TRY(cudaEventSynchronize(f->evDone));
    if (n > 0) {
        memcpy(next_pts, f->hpts + oNext, 8 * (size_t)n);
        memcpy(status, f->hpts + oSt, (size_t)n);
        if (err) memcpy(err, f->hpts + oErr, 4 * (size_t)n);
    }
    return RUMI_OK;
}

int rumi_flow_track(rumi_flow* f, const uint8_t* prev, const uint8_t* next, int w, int h, size_t stride,
                    const float* prev_pts, int n, float* next_pts, uint8_t* status, float* err) {
    int rc = rumi_flow_set_prev(f, prev, w, h, stride);
    if (rc) return rc;
    return rumi_flow_track_next(f, next, stride, prev_pts, n, next_pts, status, err, 0);
}

int rumi_flow_debug_level(rumi_flow* f, int which, int level, uint8_t* dst, int* w, int* h) {
    if (!f || !f->havePrev) return fail(RUMI_ERR_ARG, "no frames held");
    if (level < 0 || level >= f->levels) return fail(RUMI_ERR_ARG, "level out of range");
    const int slot = which ? f->prevSlot ^ 1 : f->prevSlot;
    if (!f->img[slot]) return fail(RUMI_ERR_ARG, "that frame was never uploaded");
    if (w) *w = f->lw[level];
    if (h) *h = f->lh[level];
    if (dst) {
        CU_TRY(cudaSetDevice(f->device));
        CU_TRY(cudaStreamSynchronize(f->stream));
        CU_TRY(cudaMemcpy2D(dst, f->lw[level], f->img[slot] + f->loff[level], f->lstride[level], f->lw[level],
                            f->lh[level], cudaMemcpyDeviceToHost));
    }
    return RUMI_OK;
}

int rumi_flow_debug_deriv(rumi_flow* f, int level, int16_t* dst) {
    if (!f || !f->havePrev) return fail(RUMI_ERR_ARG, "no frames held");
    if (level < 0 || level >= f->levels || !dst) return fail(RUMI_ERR_ARG, "level out of range / NULL buffer");
    CU_TRY(cudaSetDevice(f->device));
    CU_TRY(cudaStreamSynchronize(f->stream));
    CU_TRY(cudaMemcpy(dst, f->deriv + f->doff[level], (size_t)f->lw[level] * f->lh[level] * 4, cudaMemcpyDeviceToHost));
    return RUMI_OK;
}

int rumi_flow_timer_start(rumi_flow* f) {
    if (!f) return fail(RUMI_ERR_ARG, "context is NULL");
    CU_TRY(cudaSetDevice(f->device));
    if (!f->evStart) { CU_TRY(cudaEventCreate(&f->evStart)); CU_TRY(cudaEventCreate(&f->evStop)); }
    CU_TRY(cudaEventRecord(f->evStart, f->stream));
    return RUMI_OK;
}

int rumi_flow_timer_stop(rumi_flow* f, float* ms) {
    if (!f || !f->evStart || !ms) return fail(RUMI_ERR_ARG, "timer not started");
    CU_TRY(cudaEventRecord(f->evStop, f->stream));
    CU_TRY(cudaEventSynchronize(f->evStop));
    CU_TRY(cudaEventElapsedTime(ms, f->evStart, f->evStop));
    return RUMI_OK;
}

}  // extern "C"
