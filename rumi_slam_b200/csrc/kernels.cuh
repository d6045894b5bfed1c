// Kernel launch interface of the sm_100a ORB front-end (internal; the public boundary is include/rumi_orb.h).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "orb_common.h"

namespace rumi {

// Per-call view of the images of one chunk of frames.  Level 0 may alias caller memory (device-resident input).
struct LevelView {
    const uint8_t* ptr;      // frame 0, row 0
    long long pitch;         // bytes between frames
    int stride;              // bytes between rows
    int w, h;
};

struct ChunkView {
    int nframes;
    LevelView src[kMaxLevels];     // un-blurred pyramid
    LevelView blur[kMaxLevels];    // 7x7 Gaussian of each level
};

// One destination index of the fixed-point bilinear resize (cv::resize INTER_LINEAR, 8U).
struct ResizeCoef { uint16_t ofs; int16_t a0, a1; uint16_t pad; };

// Pyramid tile shape (output pixels per CTA) -- also fixes the TMA box computed on the host.
constexpr int kPyrTileW = 64, kPyrTileH = 32, kPyrThreads = 256;

struct PyramidLevelArgs {
    LevelView src, dst;
    const ResizeCoef* xc;    // [dst.w]
    const ResizeCoef* yc;    // [dst.h]
    int boxW, boxH;          // source box staged per tile (boxW multiple of 16)
    int nframes;
    int* err;                // mapped host flag: set when the TMA transaction did not complete in time
};

void launch_pyramid_level(const PyramidLevelArgs& a, const CUtensorMap* tmap /* nullptr = plain loads */,
                          cudaStream_t s);
// ---- K1, the default pyramid path (pyramid_strip.cu): all levels of a chunk in ONE launch, levels chained in shared memory ----
// Per group of 4 destination columns: where the 8-byte source window starts and how its bytes map to the outputs.
struct alignas(8) PyrColGroup {
    uint16_t word0;          // first 32-bit source word of the window
    uint8_t shift;           // 8 * (source byte offset & 3): funnel shift that aligns the window
    uint8_t pad;
    uint16_t sel01, sel23;   // PRMT selectors: (p0,p1) of outputs 0|1 and 2|3 as bytes of one register each
    uint32_t coef[4];        // a0 | a1 << 16 (11-bit fixed-point weights) per output
};
// Per destination row: vertical weights << 16 of the taps (sy1 - 1, sy1), and `adv` = how many source rows lie between
// the NEXT row's second tap and this row's (0, 1 or 2 for scale factors <= 2).  A row whose taps coincide (clamped at the bottom,
// second weight 0) is stored as (sy1 - 1, sy1) with weights (0, b0): the same value.
struct alignas(16) PyrRow { uint32_t b0s, b1s; int32_t sy1, adv; };

// One destination level of the strip kernel (source = level l-1).
struct PyrStripLevel {
    const PyrColGroup* cols;   // [groups]
    const PyrRow* rows;        // [dst.h]
    int groups;                // ceil(dst.w / 4)
    int srcLastWord;           // last readable 32-bit word of a source row
    int rowsPerItem;           // destination rows a thread marches per item (chosen per launch geometry)
};
constexpr int kPyrStripThreads = 512;
struct PyrStripArgs {
    ChunkView cv;
    PyrStripLevel lv[kMaxLevels];
    int nlevels, nstrips;
    const int2* ranges;        // [strip][kMaxLevels]: destination rows [x, y) of each level the strip computes
    int buf1Offset;            // byte offset of the second ping-pong level buffer in dynamic shared memory
    int rowTabOffset;          // byte offset of the row-record table of the current level (16 B per strip row)
};
int pyramid_strip_prepare(size_t smemBytes);      // opt in to `smemBytes` of dynamic shared memory (current device)
void launch_pyramid_strip(const PyrStripArgs& a, size_t smemBytes, cudaStream_t s);

void launch_blur(const ChunkView& cv, const OrbConst& oc, cudaStream_t s);

// One FAST grid cell (R/lib_src/ORBextractor.cc:748-763): sub-image origin and size, clipped to the level border.
struct alignas(8) FastCell { uint16_t iniX, iniY; uint8_t cw, ch, level, valid; };
static_assert(sizeof(FastCell) == 8, "FastCell is read as one 8-byte word");

struct FastLayout { int score, valid, clist, queue, warpBytes; };
FastLayout fast_layout(int tilePitch, int tileRows, int scoreRows);

struct FastArgs {
    ChunkView cv;
    const FastCell* cells;   // [totalCells], all levels
    uint32_t* cand;          // level-major: [level][frame][candCap]
    long long candLevelOff[kMaxLevels];   // element offset of (level, frame 0)
    int* levelCount;         // [frame][nlevels] (zeroed by the caller)
    int* cellOff;            // [frame][totalCells]
    int* cellCount;          // [frame][totalCells]
    int tilePitch, tileRows, scoreRows;   // per-warp shared-memory tile sizes (score tile pitch == tilePitch)
    FastLayout lay;          // byte offsets of the per-warp shared-memory regions (fast_layout)
    uint8_t* dbg;            // optional: image + score tile of (frame 0, dbgCell) for tests
    int dbgCell;
};
void launch_fast(const FastArgs& a, const OrbConst& oc, cudaStream_t s);

// ORBextractor::operator() output order (:1077-1085): what the slot assignment of one frame needs
struct SlotArgs {
    const uint32_t* sel; const int* selCount;      // [frame][kpCap], [frame][nlevels]
    int lap0, lap1;
    int* slot;                                     // [frame][kpCap] output slot of each level-ordered keypoint
    int* nkp; int* nmono;                          // [frame]
};

struct OctreeArgs {
    int nframes;
    const uint32_t* cand;    // as written by FAST
    uint32_t* candOrdered;   // same layout, reference insertion order
    long long candLevelOff[kMaxLevels];
    const int* levelCount; const int* cellOff; const int* cellCount;
    uint64_t* bigKeys;       // global sort scratch for problems that do not fit shared memory
    long long bigKeysLevelOff[kMaxLevels];   // element offset of (level, frame 0)
    int bigKeysCap[kMaxLevels];              // pow2 capacity per (level, frame)
    uint32_t* sel;           // [frame][kpCap] packed selected candidates, level order
    int* selCount;           // [frame][nlevels]
    long long* dbgClk;       // optional [nlevels][16] phase cycle counters of frame 0 (profiling hook)
    int levelFirst;          // level of blockIdx.y == 0 (0 unless the launch is split per level)
    int smemKeys;            // keys that fit the shared-memory sort buffer
    int maxNodeCap;
    // fused slot assignment: the LAST level CTA of a frame to finish (frameDone counter) assigns the output slots of that
    // frame, which saves the separate one-CTA-per-frame launch (fuseSlots = 0: launch_assign_slots runs instead)
    int fuseSlots;
    int* frameDone;          // [frame], zeroed per call
    SlotArgs slots;
    // levels with more candidates than smemKeys: appended to bigList by the first launch, processed by a second one with a
    // key buffer of smemKeysBig keys (bigList == nullptr: everything in the first launch, global scratch for big levels)
    int threads;             // CTA width of the first launch: 256, or 1024 for calls of a few frames (0 = 256)
    int* bigCount;           // zeroed per call
    int* denseFlag;          // mapped host int: set when a level had to be sorted in global memory (no second pass this call)
    int* bigList;            // [frame * nlevels] (frame << 8 | level)
    int smemKeysBig;
};
void launch_octree(const OctreeArgs& a, const OrbConst& oc, cudaStream_t s);
size_t octree_smem_bytes(int smemKeys, int maxNodeCap, int nthreads);

struct DescribeArgs {
    ChunkView cv;
    const uint32_t* sel; const int* selCount;
    int lap0, lap1;
    int* slot;               // [frame][kpCap] output slot of each level-ordered keypoint
    KeyPointRec* kps;        // [frame][outCap]
    uint8_t* desc;           // [frame][outCap][32]
    int* nkp; int* nmono;    // [frame]
    int outCap;
};
void launch_assign_slots(const DescribeArgs& a, const OrbConst& oc, cudaStream_t s);
void launch_describe(const DescribeArgs& a, const OrbConst& oc, cudaStream_t s);

// CloudFrameComputeDescriptors: descriptors of given keypoints on a given (un-pyramided, un-blurred) image.
void launch_describe_given(const uint8_t* imgs, int w, int h, int stride, long long pitch, int nimg, const int* kpOff,
                           const KeyPointRec* kps, int n, uint8_t* desc, cudaStream_t s);

// ---- matching ----
// partial[slice][nq] = {d1:16 | d2:16 | global train index:32} of train rows [slice*sliceRows, ...) (+tBase)
int match_slices(int nq, int nt);
void launch_hamming_top2_partial(const uint8_t* Q, int nq, const uint8_t* T, int nt, int tBase, int slices,
                                 uint64_t* partial, cudaStream_t s);
// K8-U (match_umma.cu): tcgen05.mma kind::i8 with TMEM accumulators, operands expanded in-kernel from the packed rows
int umma_slices(int nq, int nt);
size_t umma_train_bytes(int nt);        // scratch for the train set as ready-to-load operand tiles
void launch_hamming_top2_umma(const uint8_t* Q, int nq, const uint8_t* T, int nt, uint8_t* trainTiles, int tBase, int slices,
                              uint64_t* partial, cudaStream_t s);
// merge per-shard candidates {d1,d2,idx} gathered in shard order: cand[shard][nq]
struct PairSegment { int32_t qStart, qCount, tStart, tCount; };   // one (query set, train set) pair of K8-S
void launch_hamming_top2_segments(const uint8_t* Q, const uint8_t* T, const PairSegment* segs, int nseg, int maxQ,
                                  int32_t* idx1, uint16_t* d1, uint16_t* d2, cudaStream_t s);
void launch_hamming_candidates(const uint8_t* Q, int nq, const uint8_t* T, const int32_t* off, const int32_t* idx,
                               uint16_t* dist, int32_t* idx1, uint16_t* d1, int32_t* idx2, uint16_t* d2, cudaStream_t s);
void launch_top2_merge(const uint64_t* packed, int nshards, int nq, int32_t* idx1, uint16_t* d1, uint16_t* d2,
                       cudaStream_t s);
void launch_top2_merge_packed(const uint64_t* packed, int nshards, int nq, uint64_t* out, cudaStream_t s);
void launch_pack_top2(const int32_t* idx1, const uint16_t* d1, const uint16_t* d2, int nq, uint64_t* packed,
                      cudaStream_t s);

void launch_stereo_best1(const KeyPointRec* Lk, const uint8_t* Ld, int nL, const KeyPointRec* Rk, const uint8_t* Rd,
                         int nR, const float* scaleFactors, int nRows, float minD, float maxD, int32_t* bestR,
                         uint16_t* bestDist, cudaStream_t s);

// ---- bag of words (bow.cu) ----
struct BowTreeView {            // DBoW2 vocabulary tree in CSR form, node 0 = root
    int k, L, nnodes;
    const int32_t* childStart;   // [nnodes] first entry of the node's children in childIds
    const int32_t* childCount;   // [nnodes] 0 = leaf
    const int32_t* childIds;     // children in the reference's push_back order (loadFromTextFile :1390)
    const uint8_t* desc;         // [nnodes][32]
    const int32_t* wordId;       // [nnodes] word id of a leaf
    const double* weight;        // [nnodes]
};
struct BowSegment { int32_t aStart, aCount, bStart, bCount, outOff; };
void launch_bow_transform(const BowTreeView& t, const uint8_t* desc, int n, int levelsup, int32_t* word, double* weight,
                          int32_t* node, cudaStream_t s);
void launch_bow_node_distances(const uint8_t* A, const uint8_t* B, const int32_t* aIdx, const int32_t* bIdx,
                               const BowSegment* segs, int nseg, uint16_t* dist, cudaStream_t s);

void launch_distinctive(const uint8_t* desc, const int32_t* offsets, int npoints, int32_t* bestIdx, int32_t* bestMedian,
                        cudaStream_t s);

struct StereoRefineArgs {
    LevelView left[kMaxLevels], right[kMaxLevels];     // frame 0 of the two extractors' pyramids
    float scale[kMaxLevels], invScale[kMaxLevels];
    const KeyPointRec* Lk; const KeyPointRec* Rk;
    const int32_t* bestR; const uint16_t* bestDist;
    int nL;
    float minD, maxD, mbf;
    float* uRight; float* depth; int* sad;
};
void launch_stereo_refine(const StereoRefineArgs& a, cudaStream_t s);

// ---- sparse pyramidal Lucas-Kanade flow of KFDSample::Step (flow.cu) ----
constexpr int kFlowMaxLevels = 8;
struct FlowPyramidView { const uint8_t* ptr[kFlowMaxLevels]; int w[kFlowMaxLevels], h[kFlowMaxLevels], stride[kFlowMaxLevels]; };
struct FlowDerivView { short2* ptr[kFlowMaxLevels]; };       // [h][w] (dx, dy) per level, zero outside the image
struct FlowTrackArgs {
    FlowPyramidView I, J;        // previous / next frame
    FlowDerivView D;             // Scharr derivatives of the previous frame
    int maxLevel, maxCount;
    double eps2;                 // criteria.epsilon squared
    float minEig;
    const float2* prevPts; float2* nextPts; uint8_t* status; float* err;
    int n;
};
void launch_flow_pyrdown(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dstride, cudaStream_t s);
void launch_flow_scharr(const FlowPyramidView& img, const FlowDerivView& der, int levels, cudaStream_t s);
bool launch_flow_track(const FlowTrackArgs& a, int win, cudaStream_t s);   // false: window size not built

}  // namespace rumi
