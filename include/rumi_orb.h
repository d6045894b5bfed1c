/* rumi_orb.h -- C ABI of the B200-native ORB front-end (librumi_orb.so).
 *
 * Drop-in boundary for RUMI-SLAM's feature front-end.  Each entry point names the reference interface it
 * replaces (R/ = src/rumi-slam/ of Changfei-Fu/RUMI-SLAM).  Plain C: opaque handles, plain pointers and sizes,
 * int return codes (0 = ok, < 0 = error, text via rumi_last_error()).  No exceptions cross this boundary and
 * there is NO CPU fallback: without a usable CUDA device every call fails with RUMI_ERR_CUDA.
 *
 * Threading: one handle = one CUDA stream + one workspace; a handle is single-threaded (like a reference
 * ORBextractor instance, which mutates mvImagePyramid), distinct handles may be used concurrently
 * (R/lib_src/Frame.cc:116-119 runs the left/right extractors in two threads).
 */
#ifndef RUMI_ORB_H
#define RUMI_ORB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RUMI_OK 0
#define RUMI_ERR_EMPTY (-1)      /* empty image: ORBextractor::operator() returns -1 (ORBextractor.cc:1017) */
#define RUMI_ERR_ARG (-2)
#define RUMI_ERR_SHAPE (-3)      /* shape the reference itself cannot process (level smaller than one 35-px cell, ...) */
#define RUMI_ERR_CUDA (-4)
#define RUMI_ERR_CAPACITY (-5)
#define RUMI_ERR_BORDER (-6)     /* keypoint closer than 19 px to the image border (reference reads out of bounds) */

/* == cv::KeyPoint, 28 bytes (pt.x, pt.y, size, angle [deg], response, octave, class_id) */
typedef struct rumi_kp {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} rumi_kp;

typedef struct rumi_orb rumi_orb;
typedef struct rumi_match rumi_match;

const char* rumi_last_error(void);
int rumi_device_count(void);

/* ORBextractor::ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)   R/lib_src/ORBextractor.cc:405-461
 * `max_batch` = frames processed per internal chunk (workspace is sized for it); w/h are learned at first use. */
int rumi_orb_create(rumi_orb** out, int nfeatures, float scale_factor, int nlevels, int ini_th_fast,
                    int min_th_fast, int device, int max_batch);
void rumi_orb_destroy(rumi_orb* h);

/* GetLevels / GetScaleFactor(s) / GetInverseScaleFactors / GetScaleSigmaSquares / GetInverseScaleSigmaSquares
 * R/include/cloud_edge_slam_lib/ORBextractor.h:62-84.  Each table has nlevels floats; mnFeaturesPerLevel in quota. */
int rumi_orb_levels(const rumi_orb* h);
int rumi_orb_tables(const rumi_orb* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2, int* quota);

/* Upper bound of keypoints per frame for a w x h image (sum over levels of quota + 3): size kps/desc with it. */
int rumi_orb_frame_capacity(rumi_orb* h, int w, int h_px);

/* ORBextractor::operator()(image, mask, keypoints, descriptors, vLappingArea)   R/lib_src/ORBextractor.cc:1014-1091
 * img: HOST pointer, 8-bit gray, `stride` bytes per row.  lap0/lap1 = vLappingArea.  Writes up to `cap` keypoints
 * (28 B each) and descriptors (32 B each) in the reference's output order; *n_kp = total, *n_mono = return value
 * of operator() (monoIndex).  NULL/0-sized image -> RUMI_ERR_EMPTY (reference returns -1). */
int rumi_orb_extract(rumi_orb* h, const uint8_t* img, int w, int h_px, size_t stride, int lap0, int lap1,
                     rumi_kp* kps, uint8_t* desc, int cap, int* n_kp, int* n_mono);

/* Same for n frames (frame i at imgs + i*frame_pitch); frame-sharded work unit of the back-submap rebuild
 * (SURVEY.md 8e).  HOST pointers; H2D, kernels and D2H of consecutive chunks overlap on two streams.
 * kps: [n][cap_per_frame], desc: [n][cap_per_frame][32], n_kp/n_mono: [n].  cap_per_frame >= frame capacity. */
int rumi_orb_extract_batch(rumi_orb* h, const uint8_t* imgs, int n, int w, int h_px, size_t stride,
                           size_t frame_pitch, int lap0, int lap1, rumi_kp* kps, uint8_t* desc,
                           int cap_per_frame, int* n_kp, int* n_mono);

/* Same with every buffer already resident in device memory (inputs stay in HBM, results stay in HBM).
 * Asynchronous on the handle's stream unless `sync` != 0. */
int rumi_orb_extract_batch_device(rumi_orb* h, const uint8_t* d_imgs, int n, int w, int h_px, size_t stride,
                                  size_t frame_pitch, int lap0, int lap1, rumi_kp* d_kps, uint8_t* d_desc,
                                  int cap_per_frame, int* d_n_kp, int* d_n_mono, int sync);

/* Stream ordering for the device-resident entry points.  The library launches on PRIVATE non-blocking streams, which
 * no caller stream is ordered against implicitly.  `stream` is the caller's cudaStream_t (NULL = legacy default stream):
 *   wait_stream    work the library enqueues after this call starts only after everything enqueued on `stream` so far
 *                  (the caller produced the inputs / zero-filled the outputs there);
 *   signal_stream  work the caller enqueues on `stream` after this call starts only after everything the library has
 *                  enqueued so far (the caller consumes the results there).
 * Neither call blocks the host.  (The same pair exists for rumi_match and rumi_vocab below.) */
int rumi_orb_wait_stream(rumi_orb* h, void* stream);
int rumi_orb_signal_stream(rumi_orb* h, void* stream);

/* The same call in two halves, for callers that run several extractors from ONE host thread (the reference starts two
 * std::threads for the left and the right image of a stereo frame, R/lib_src/Frame.cc:116-119, because its extraction is CPU
 * bound; on the GPU the two frames only have to be IN FLIGHT together): _begin enqueues upload, kernels and the download
 * into the handle's pinned staging block and returns without waiting; _end waits and fills the outputs exactly like
 * rumi_orb_extract.  One frame in flight per handle. */
int rumi_orb_extract_begin(rumi_orb* h, const uint8_t* img, int w, int h_px, size_t stride, int lap0, int lap1);
int rumi_orb_extract_end(rumi_orb* h, rumi_kp* kps, uint8_t* desc, int cap, int* n_kp, int* n_mono);

/* ORBextractor::CloudFrameComputeDescriptors(image, keypoints, descriptors)   R/lib_src/ORBextractor.cc:989-1011
 * HOST pointers.  Returns n (like the reference) or an error code. */
int rumi_orb_describe(rumi_orb* h, const uint8_t* img, int w, int h_px, size_t stride, const rumi_kp* kps, int n,
                      uint8_t* desc);
/* The same for MANY key frames of one shape in one call (one upload, one launch): frame f = imgs + f * frame_pitch, its
 * keypoints are kps[kp_off[f] .. kp_off[f + 1]) (kp_off[0] == 0), desc[k] belongs to kps[k].  This is how the cloud key
 * frames of a submap merge get real descriptors (they carry zero descriptors in the reference,
 * R/src/cloud_edge_main.cpp:937).  Returns kp_off[nimg] or an error code. */
int rumi_orb_describe_batch(rumi_orb* h, const uint8_t* imgs, int nimg, int w, int h_px, size_t stride, size_t frame_pitch,
                            const rumi_kp* kps, const int32_t* kp_off, uint8_t* desc);

/* mvImagePyramid[level] of the LAST rumi_orb_extract call (R/include/cloud_edge_slam_lib/ORBextractor.h:86;
 * read by Frame::ComputeStereoMatches, R/lib_src/Frame.cc:834,918-932).  Copies the level to host memory. */
int rumi_orb_pyramid_level(rumi_orb* h, int level, uint8_t* dst, size_t dst_stride, int* w, int* h_px);
/* on != 0: every single-frame call (rumi_orb_extract / _begin) also brings the frame's pyramid to a pinned host block behind
 * the kernels, so that rumi_orb_pyramid_level is a host copy without a device round trip -- for callers that read
 * mvImagePyramid after every frame (the reference's Frame::ComputeStereoMatches, R/lib_src/Frame.cc:834, 918-932). */
int rumi_orb_set_pyramid_staging(rumi_orb* h, int on);
/* Blurred level of the last call (test hook for the 7x7 Gaussian, R/lib_src/ORBextractor.cc:1057-1058). */
int rumi_orb_blurred_level(rumi_orb* h, int level, uint8_t* dst, size_t dst_stride, int* w, int* h_px);
/* Stage-level test hooks: FAST candidates of `level` of the last single-frame call in the reference's insertion
 * order, as (x, y, response) int32 triples relative to (16,16); returns the count (<= cap written). */
int rumi_orb_debug_candidates(rumi_orb* h, int level, int32_t* xyr, int cap);
int rumi_orb_debug_selected(rumi_orb* h, int level, int32_t* xyr, int cap);
/* Test hook: arms (out == NULL) or reads back the shared-memory image + score tile of FAST cell `cell` of frame 0. */
int rumi_orb_debug_fast_tile(rumi_orb* h, int cell, uint8_t* out, int cap, int* dims5);
/* profiling hook: first call arms per-phase cycle counters of the quad-tree kernel (frame 0 of every chunk), later
   calls copy out[level*16 + phase] and reset them; returns the number of counters */
int rumi_orb_debug_octree_clocks(rumi_orb* h, long long* out, int cap);
/* timing experiments only (results become stale / wrong): bit s of mask = do not launch stage s (0 pyramid, 1 FAST,
   2 quad-tree, 4 blur, 5 describe) in later calls, to measure a stage's marginal cost inside the pipelined batch */
int rumi_orb_debug_skip_stages(rumi_orb* h, int mask);

/* ---- measurement (bench.py) ----
 * Device-side timing on the streams the kernels are launched on: start records an event on the handle's first
 * stream (the second waits for it), stop joins both streams, records, synchronises and returns the elapsed ms. */
int rumi_orb_timer_start(rumi_orb* h);
int rumi_orb_timer_stop(rumi_orb* h, float* ms);
/* Optional per-stage CUDA events around every launch group (0 pyramid, 1 FAST, 2 quad-tree, 3 slots, 4 blur,
 * 5 orientation+descriptors).  profile_read synchronises, accumulates and returns the number of stages. */
int rumi_orb_profile(rumi_orb* h, int enable);
/* Number of workspaces / CUDA streams consecutive chunks alternate between (1..4; 0 restores the default).  With 1 the
 * stages of a chunk run back to back on one stream, which makes the per-stage event times exclusive. */
int rumi_orb_set_streams(rumi_orb* h, int n);
int rumi_orb_profile_read(rumi_orb* h, double* stage_ms, long long* stage_launches, int reset);
/* Number of kernels this handle has launched (own kernels only; memsets / copies are not counted). */
long long rumi_orb_launch_count(rumi_orb* h, int reset);

/* ---- matching ---- */
int rumi_match_create(rumi_match** out, int device);
void rumi_match_destroy(rumi_match* m);

/* Brute-force top-2 Hamming of nq query vs nt train descriptors (32 B rows).  Semantics of the best/second-best
 * scan shared by ORBmatcher (R/lib_src/ORBmatcher.cc:253-261) and BFMatcher.knnMatch(k=2) (R/lib_src/Frame.cc:1139):
 * ascending train index, strict '<' -> ties keep the earliest index; d2 may equal d1; no candidate -> 256 / -1.
 * HOST pointers. */
int rumi_hamming_top2(rumi_match* m, const uint8_t* Q, int nq, const uint8_t* T, int nt, int32_t* idx1,
                      uint16_t* d1, uint16_t* d2);
/* Device-resident variant; train indices are reported as t_base + local index (train shard of a larger set). */
int rumi_hamming_top2_device(rumi_match* m, const uint8_t* dQ, int nq, const uint8_t* dT, int nt, int t_base,
                             int32_t* d_idx1, uint16_t* d_d1, uint16_t* d_d2, int sync);
/* The same raw top-2 for MANY independent (query set, train set) pairs in one launch: the descriptor association of
 * the matched key-frame pairs of a submap merge (the pairs R/lib_src/CloudMerging.cc:503-551 walks).
 * segs[s] = {q_start, q_count, t_start, t_count} (rows of Q / T).  Per query row: idx1 = best train row RELATIVE to
 * its segment's t_start (= feature index inside the second key frame), earliest index among ties, d2 may equal d1;
 * (-1, 256, 256) for an empty train set and for rows no segment covers.  Result == rumi_hamming_top2 pair by pair. */
int rumi_hamming_top2_pairs(rumi_match* m, const uint8_t* Q, int nq, const uint8_t* T, int nt, const int32_t* segs,
                            int nseg, int32_t* idx1, uint16_t* d1, uint16_t* d2);
/* Candidate-list matching -- the form every window / projection search of ORBmatcher takes (SearchByProjection,
 * R/lib_src/ORBmatcher.cc:70-111; SearchForInitialization, :606-630; ...): query q is compared ONLY with the train rows
 * cand_idx[cand_off[q] .. cand_off[q + 1]), e.g. what Frame::GetFeaturesInArea (R/lib_src/Frame.cc:695-750) returned for
 * it.  dist[k] = DescriptorDistance(Q[q], T[cand_idx[k]]) for every list entry k -- the adapters replay the call site's own
 * acceptance on them (several sites skip candidates depending on EARLIER acceptances: ORBmatcher.cc:86-88, :617) -- and,
 * when idx1 / d1 / idx2 / d2 are given, the raw best and second best of every list in list order (strict '<': the
 * earliest entry wins ties; idx2 lets the caller apply the level-aware ratio test of :101-104; -1 / 256 when a list has
 * fewer than one / two entries).  cand_off has nq + 1 entries, cand_off[0] == 0.  HOST pointers. */
int rumi_hamming_candidates(rumi_match* m, const uint8_t* Q, int nq, const uint8_t* T, int nt, const int32_t* cand_off,
                            const int32_t* cand_idx, uint16_t* dist, int32_t* idx1, uint16_t* d1, int32_t* idx2, uint16_t* d2);
int rumi_match_timer_start(rumi_match* m);
int rumi_match_timer_stop(rumi_match* m, float* ms);
/* Which top-2 kernel the last rumi_hamming_top2* call used: 1 = LOP3+POPC (small problems), 3 = tcgen05 / TMEM int8
 * kernel (>= 64 Mi pairs and >= 256 queries).
 * RUMI_MATCH=popc|umma forces one kernel; the results are identical. */
int rumi_match_last_path(const rumi_match* m);
long long rumi_match_launch_count(rumi_match* m, int reset);

/* Shard exchange step: pack local results into 8-byte candidates for the all-gather, and merge
 * `nshards` gathered candidate arrays ([shard][nq], ascending train ranges) with the same '<' rule. */
int rumi_top2_pack_device(rumi_match* m, const int32_t* d_idx1, const uint16_t* d_d1, const uint16_t* d_d2, int nq,
                          uint64_t* d_packed, int sync);
int rumi_top2_merge_device(rumi_match* m, const uint64_t* d_packed, int nshards, int nq, int32_t* d_idx1,
                           uint16_t* d_d1, uint16_t* d_d2, int sync);

int rumi_match_wait_stream(rumi_match* m, void* stream);
int rumi_match_signal_stream(rumi_match* m, void* stream);

/* ---- multi-GPU all-pairs matching (BASELINE config 5): one process per GPU ----
 * The train set is sharded into contiguous index ranges; every rank scans ALL queries against its shard (global train
 * indices = t_base + local row), the per-rank {d1,d2,idx} candidates (8 B per query) are exchanged with ONE
 * ncclAllGather and folded in rank order with the reference's strict '<' rule -- bit-identical to the single-GPU scan
 * because every index of rank r precedes every index of rank r+1.  Everything (scan, exchange, fold) is enqueued on
 * the matcher's stream; nothing synchronises the host unless `sync` != 0.  Every rank receives the full result.
 *
 * Communicator: rank 0 calls rumi_nccl_unique_id and hands the 128 bytes (== ncclUniqueId) to the other ranks by the
 * host application's own means (MPI, torch.distributed, a socket, a file); every rank then calls
 * rumi_match_comm_init (collective: ncclCommInitRank).  A host that already owns an ncclComm_t for these ranks passes
 * it with rumi_match_comm_adopt instead (not destroyed by the library).  NCCL is resolved at run time
 * (dlopen libnccl.so.2): single-GPU users need no NCCL. */
int rumi_nccl_unique_id(uint8_t* id128);
int rumi_match_comm_init(rumi_match* m, const uint8_t* id128, int rank, int nranks);
int rumi_match_comm_adopt(rumi_match* m, void* nccl_comm, int rank, int nranks);
int rumi_hamming_top2_sharded(rumi_match* m, const uint8_t* dQ, int nq, const uint8_t* dT_local, int nt_local,
                              int t_base, int32_t* d_idx1, uint16_t* d_d1, uint16_t* d_d2, int sync);

/* Stereo row-band best-1 search of Frame::ComputeStereoMatches (R/lib_src/Frame.cc:828-905): for every left
 * keypoint the right keypoint with the smallest Hamming distance among those whose row band [y-2s, y+2s] covers the
 * left row, octave within +-1 and uR in [uL-max_d, uL-min_d]; best_dist starts at TH_HIGH=100, ties keep the lowest
 * right index; best_r = -1 when nothing beat 100.  HOST pointers; keypoints are rumi_kp records (n < 2^20). */
int rumi_stereo_best1(rumi_match* m, const rumi_kp* Lk, const uint8_t* Ld, int nL, const rumi_kp* Rk,
                      const uint8_t* Rd, int nR, const float* scale_factors, int nlevels, int n_rows, float min_d,
                      float max_d, int32_t* best_r, uint16_t* best_dist);

/* Frame::ComputeStereoMatches, complete (R/lib_src/Frame.cc:828-985): rumi_stereo_best1, then the 11x11 SAD slide of
 * +-5 px on the pyramids of the two extractors (the pyramids of their LAST rumi_orb_extract call are used where they
 * are, in device memory), the parabola sub-pixel fit, disparity / depth, and the median outlier cut.
 * u_right / depth == mvuRight / mvDepth (-1 where no stereo match).  mbf = baseline * fx, mb = baseline.
 * HOST pointers; *n_matched = number of stereo matches kept. */
int rumi_stereo_match(rumi_match* m, rumi_orb* left, rumi_orb* right, const rumi_kp* Lk, const uint8_t* Ld, int nL,
                      const rumi_kp* Rk, const uint8_t* Rd, int nR, float mbf, float mb, float* u_right, float* depth,
                      int* n_matched);

/* ---- bag of words ----
 * rumi_vocab = the DBoW2 vocabulary tree (ORBVocabulary = TemplatedVocabulary<FORB::TDescriptor, FORB>,
 * R/include/cloud_edge_slam_lib/ORBVocabulary.h; R/Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h) resident on the
 * device.  Nodes are given in id order exactly as loadFromTextFile (:1338-1421) reads them: node 0 is the root,
 * parent[i] < i, is_leaf[i] is the file's leaf flag (word ids are handed out to flagged nodes in id order), desc is
 * [nnodes][32], weight the node weight (WordValue = double). */
typedef struct rumi_vocab rumi_vocab;
int rumi_vocab_create(rumi_vocab** out, int device, int k, int L, int nnodes, const int32_t* parent,
                      const uint8_t* is_leaf, const uint8_t* desc, const double* weight);
void rumi_vocab_destroy(rumi_vocab* v);
int rumi_vocab_words(const rumi_vocab* v);
int rumi_vocab_wait_stream(rumi_vocab* v, void* stream);
int rumi_vocab_signal_stream(rumi_vocab* v, void* stream);
/* Per-feature tree descent == TemplatedVocabulary::transform(feature, word_id, weight, &nid, levelsup) (:1218-1258):
 * word id and weight of the leaf reached, node id at depth L - levelsup (0 = root when that depth is <= 0).  Host
 * buffers; the BowVector / FeatureVector maps of Frame::ComputeBoW are assembled from these by the caller. */
int rumi_bow_transform(rumi_vocab* v, const uint8_t* desc, int n, int levelsup, int32_t* word_id, double* weight,
                       int32_t* node_id);
/* Same with device-resident descriptors and outputs (16-byte aligned descriptors). */
int rumi_bow_transform_device(rumi_vocab* v, const uint8_t* d_desc, int n, int levelsup, int32_t* d_word_id,
                              double* d_weight, int32_t* d_node_id, int sync);
long long rumi_vocab_launch_count(rumi_vocab* v, int reset);
/* ORBmatcher::SearchByBoW distance blocks (R/lib_src/ORBmatcher.cc:198-370).  segs[s] = {aStart, aCount, bStart,
 * bCount, outOff}: features a_idx[aStart .. aStart+aCount) of A and b_idx[bStart ..) of B share vocabulary node s;
 * dist[outOff + i * bCount + j] = DescriptorDistance(A[a_idx[aStart+i]], B[b_idx[bStart+j]]).  Host buffers. */
int rumi_bow_node_distances(rumi_match* m, const uint8_t* descA, int nA, const uint8_t* descB, int nB,
                            const int32_t* a_idx, int n_a_idx, const int32_t* b_idx, int n_b_idx,
                            const int32_t* segs, int nseg, uint16_t* dist, long long ndist);
/* MapPoint::ComputeDistinctiveDescriptors (R/lib_src/MapPoint.cc:355-426), batched over map points: the observed
 * descriptors of point p are desc[offsets[p] .. offsets[p+1]) (32 bytes each); best_idx[p] = index (within the point's
 * list) of the descriptor with the least median distance to the rest, first one on ties; best_median[p] that median
 * (-1 / -1 for a point without observations).  Host buffers. */
int rumi_distinctive_descriptors(rumi_match* m, const uint8_t* desc, const int32_t* offsets, int npoints,
                                 int32_t* best_idx, int32_t* best_median);

/* ---- sparse pyramidal Lucas-Kanade flow ----
 * Replaces cv::calcOpticalFlowPyrLK as KFDSample::Step calls it on every untracked frame
 * (R/lib_src/KFDSample.cc:131-132: winSize 31x31, maxLevel 2, TermCriteria(COUNT+EPS, 20, 0.03),
 * R/include/cloud_edge_slam_lib/KFDSample.h:47; flags 0, minEigThreshold 1e-4 = OpenCV defaults).
 * Points are (x, y) float pairs = cv::Point2f.  status[i] = 1 when the flow of point i was found; err[i] = mean
 * absolute patch difference / 32 (OpenCV's default error measure), 0 when status[i] = 0.  The pyramid stops at the
 * last level whose next size would not exceed the window, like cv::buildOpticalFlowPyramid.
 * The handle keeps the previous frame's pyramid and Scharr derivatives on the device: `advance` = 1 makes the frame
 * just tracked the new previous frame (imprvs = imnext.clone(), KFDSample.cc:169), so the steady state uploads one
 * image per call.  One handle is single-threaded like the reference's KFDSample. */
typedef struct rumi_flow rumi_flow;
int rumi_flow_create(rumi_flow** out, int device, int win, int max_level, int max_count, double epsilon,
                     float min_eig_threshold);
void rumi_flow_destroy(rumi_flow* f);
int rumi_flow_set_prev(rumi_flow* f, const uint8_t* img, int w, int h_px, size_t stride);
int rumi_flow_track_next(rumi_flow* f, const uint8_t* img, size_t stride, const float* prev_pts, int n,
                         float* next_pts, uint8_t* status, float* err /* may be NULL */, int advance);
/* one-shot form of cv::calcOpticalFlowPyrLK(prev, next, prevPts, nextPts, status, err, ...) */
int rumi_flow_track(rumi_flow* f, const uint8_t* prev, const uint8_t* next, int w, int h_px, size_t stride,
                    const float* prev_pts, int n, float* next_pts, uint8_t* status, float* err);
int rumi_flow_levels(const rumi_flow* f);            /* pyramid levels in use for the frames held */
long long rumi_flow_launches(const rumi_flow* f);    /* kernels launched so far (bench.py) */
/* parity hooks (tests): pyramid level of the previous (which = 0) / last tracked (1) frame, Scharr (dx, dy) pairs */
int rumi_flow_debug_level(rumi_flow* f, int which, int level, uint8_t* dst, int* w, int* h_px);
int rumi_flow_debug_deriv(rumi_flow* f, int level, int16_t* dst);
int rumi_flow_timer_start(rumi_flow* f);
int rumi_flow_timer_stop(rumi_flow* f, float* ms);

/* ORBmatcher::DescriptorDistance for one pair (host inline popcount; the API, not a fallback). */
int rumi_descriptor_distance(const uint8_t* a, const uint8_t* b);

#ifdef __cplusplus
}
#endif
#endif /* RUMI_ORB_H */
