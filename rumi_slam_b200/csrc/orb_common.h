// Shared declarations for the sm_100a ORB front-end (host + device).
// Geometry follows the reference exactly; every table cites the reference line it restates.
// R/ = reference src/rumi-slam/.
#pragma once
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define RUMI_HD __host__ __device__ __forceinline__
#else
#define RUMI_HD inline
#endif

namespace rumi {

constexpr int kMaxLevels = 16;
constexpr int kEdge = 19;            // EDGE_THRESHOLD   R/lib_src/ORBextractor.cc:71
constexpr int kMinBorder = 16;       // EDGE_THRESHOLD-3 R/lib_src/ORBextractor.cc:732
constexpr int kHalfPatch = 15;       // HALF_PATCH_SIZE  R/lib_src/ORBextractor.cc:70
constexpr int kMaxTreeDepth = 13;    // quad-tree path digits kept per key (images up to 8192 px)
constexpr int kRootBits = 4;         // up to 16 root nodes (nIni = round(width/height))
constexpr int kOrderBits = 22;        // FAST candidates per (frame, level) < 2^22: insertion index field of the quad-tree keys
constexpr int kOverQuota = 3;        // DistributeOctTree can exceed N by at most 3  (:695-696)

// Per-level static geometry (computed once per (shape, ORB params) on the host).
struct LevelGeom {
    int w, h;                // level size                       R/lib_src/ORBextractor.cc:1095-1096
    int stride;              // bytes per row of the internal level buffers (multiple of 16)
    int nCols, nRows;        // FAST grid                        :743-744
    int wCell, hCell;        //                                  :745-746
    int cellBase;            // index of this level's first cell in the all-level cell table
    int quota;               // mnFeaturesPerLevel[level]        :428-438
    int candCap;             // candidate capacity per frame for this level (worst case after 3x3 NMS)
    int kpBase;              // first keypoint slot of this level inside a frame's level-ordered keypoint block
    int nIni;                // octree roots                     :541
    int treeDepth;           // digits needed so every leaf is a single pixel
    float hX;                // root width                       :543
    float scale;             // mvScaleFactor[level]             :412-416
    float patchSize;         // (float)(int)(31*scale)           :816
    long long pyrOff;        // byte offset of frame 0 of this level in the pyramid workspace
    long long candOff;       // element offset of frame 0's candidate block for this level
};

struct OrbConst {
    int nlevels, nfeatures, iniTh, minTh;
    int W, H;
    int totalCells;          // sum over levels of nCols*nRows
    int kpCap;               // per-frame keypoint capacity = sum(quota + kOverQuota)
    int umax[16];            // R/lib_src/ORBextractor.cc:446-460
    LevelGeom lv[kMaxLevels];
};

// cv::KeyPoint layout (28 B): pt.x pt.y size angle response octave class_id.
struct KeyPointRec {
    float x, y, size, angle, response;
    int32_t octave, class_id;
};

// Packed FAST candidate: x_rel (13 bits) | y_rel (13 bits) << 13 | response (6..8 bits) << 26 ... kept simple:
// x in [0,8191], y in [0,8191] relative to (16,16); response 1..255.
RUMI_HD uint32_t pack_cand(int x, int y, int resp) { return (uint32_t)x | ((uint32_t)y << 12) | ((uint32_t)resp << 24); }
RUMI_HD int cand_x(uint32_t c) { return (int)(c & 0xFFFu); }
RUMI_HD int cand_y(uint32_t c) { return (int)((c >> 12) & 0xFFFu); }
RUMI_HD int cand_resp(uint32_t c) { return (int)(c >> 24); }
constexpr int kMaxLevelDim = 4096 + 32;   // pack_cand keeps 12 bits per coordinate

}  // namespace rumi
