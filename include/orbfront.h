/* include/orbfront.h — C ABI of the B200-native ORB front end (liborbfront_b200.so).
 *
 * Drop-in boundary for the data-parallel hot path of ttwang0303/Adaptive-RGBD-Localization-Mappig:
 *   ORB extraction -> brute-force Hamming kNN-2 matching -> RANSAC-Kabsch 3D-3D.
 * Plain pointers, sizes and POD structs only: no OpenCV / Eigen / PCL / torch types cross this line.
 * Every entry point returns an orbf_status (0 = OK), never throws, and is thread-safe per context.
 * There is NO CPU fallback: without a CUDA device orbf_create() fails with ORBF_ERR_CUDA.
 *
 * Reference interfaces replaced (paths relative to the reference checkout):
 *   orbf_create/orbf_get_tables      ORBextractor::ORBextractor + Get* getters   Features/orbextractor.cpp:346-404, orbextractor.h:44-55
 *   orbf_extract                      ORBextractor::operator()                    Features/orbextractor.cpp:756-815 (orbextractor.h:37)
 *                                     Extractor::Extract (ORB_SLAM2 route)        Features/extractor.cpp:39-42
 *   orbf_extract_batch*               the same, batched + Frame::ExtractFeatures' depth gather   Core/frame.cpp:135-170
 *   orbf_extract_batch_bgr            Frame::Frame (cvtColor BGR2GRAY) + ExtractFeatures   Core/frame.cpp:18-45,135-170
 *   orbf_pyramid_level                public member mvImagePyramid                Features/orbextractor.h:57
 *   orbf_knn2 / orbf_knn_match        cv::BFMatcher::knnMatch(k=2) + ratio test in Matcher::KnnMatch   Features/matcher.cpp:55-66 (23-35)
 *   orbf_descriptor_distance          Matcher::DescriptorDistance                 Features/matcher.cpp:355-358
 *   orbf_distinctive_descriptors      Landmark::ComputeDistinctiveDescriptors     Core/landmark.cpp:219-273
 *   orbf_undistort_points             Frame::UndistortKeyPoints / ComputeImageBounds (cv::undistortPoints)   Core/frame.cpp:286-343
 *   orbf_projection_match             Matcher::ProjectionMatch + Frame::GetFeaturesInArea   Features/matcher.cpp:90-143, Core/frame.cpp:258-274
 *   orbf_fuse_search                  Matcher::Fuse, projection + windowed search   Features/matcher.cpp:212-296
 *   orbf_bow_match                    Matcher::BoWMatch                           Features/matcher.cpp:145-209
 *   orbf_match_pairs                  Tracking::TrackFrame's matcher call, batched  System/tracking.cpp:197-199
 *   orbf_track_sequence*              Tracking::Track's per-frame loop (extract, match with the last frame, RANSAC)
 *                                     over a whole sequence, pipelined            System/tracking.cpp:38-46,193-208
 *   orbf_ransac_iterate               Ransac::Iterate(Frame*,Frame*,m12)          Odometry/ransac.cpp:155-267
 *   orbf_ransac_pairs                 Odometry::Compute -> Ransac::Iterate, batched   Odometry/odometry.cpp:48
 *   orbf_kabsch                       Kabsch::Compute                             Odometry/kabsch.cpp:14-57
 *   orbf_compose_trajectory           Odometry::Compute (RANSAC): composition rule + SetInlier   Odometry/odometry.cpp:78-90
 *   orbf_odometry_compute             Odometry::Compute (RANSAC) for one frame pair: Iterate + clouds + composition   Odometry/odometry.cpp:44-90
 *   orbf_kfdb_*                       keyframe descriptor storage of Core/keyframedatabase (config 5 many-to-many matching)
 *   orbf_adaptive_detect              Extractor(FAST, ., ADAPTIVE): VideoGridAdaptedFeatureDetector over VideoDynamicAdaptedFeatureDetector
 *                                     over DetectorAdjuster(FAST)  Features/extractor.cpp:52-77, videogridadaptedfeaturedetector.cpp:52-84,
 *                                     videodynamicadaptedfeaturedetector.cpp:24-44, detectoradjuster.cpp:22-59 (BASELINE config 4)
 */
#ifndef ORBFRONT_H
#define ORBFRONT_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORBF_MAX_LEVELS 16
#define ORBF_ABI_VERSION 6   /* 4: + projection_match, fuse_search, bow_match, compose_trajectory, undistort_points
                                5: + distortion coefficients in orbf_config (mvKeysUn feeds the unprojection), orbf_download_keys_un
                                6: + orbf_odometry_compute (additive) */

typedef enum {
    ORBF_OK = 0,
    ORBF_ERR_ARG = 1,        /* null pointer / out-of-range argument                          */
    ORBF_ERR_CAPACITY = 2,   /* caller buffer too small; required size is still reported       */
    ORBF_ERR_GEOMETRY = 3,   /* image too small for the pyramid / patch borders                */
    ORBF_ERR_CUDA = 4,       /* CUDA runtime error or no device (see orbf_last_error)          */
    ORBF_ERR_ALIGNMENT = 5,  /* device pointer / pitch not 16-byte aligned                     */
    ORBF_ERR_STATE = 6       /* call order violated (e.g. match before extract)                */
} orbf_status;

/* POD mirrors of cv::KeyPoint (28 B) and cv::DMatch (16 B): same field order and sizes, so a
 * std::vector<cv::KeyPoint> / std::vector<cv::DMatch> can be filled with one memcpy.          */
typedef struct { float x, y, size, angle, response; int32_t octave, class_id; } orbf_keypoint;
typedef struct { int32_t queryIdx, trainIdx, imgIdx; float distance; } orbf_dmatch;
/* FAST candidate before quadtree distribution: coordinates relative to minBorder (16), cv::FAST response */
typedef struct { int32_t x, y, score; } orbf_cand;

typedef struct {
    int32_t width, height;             /* frame size in pixels                                   */
    int32_t nfeatures, nlevels;        /* ORBextractor(nfeatures, scaleFactor, nlevels, iniTh, minTh) */
    float scale_factor;
    int32_t ini_th_fast, min_th_fast;
    int32_t max_frames;                /* frame slots kept device-resident (batch capacity)      */
    int32_t max_pairs;                 /* frame-pair slots for matching / RANSAC (0 = max_frames)*/
    int32_t device;                    /* CUDA device ordinal                                    */
    float fx, fy, cx, cy, mbf;         /* Calibration:: (Utils/common.h:35-38,71)               */
    float depth_factor;                /* Calibration::depthFactor = 1/5000 (Utils/common.h:67) */
    int32_t pipeline_chunk;            /* frames per pipeline chunk of the batched calls (0 = default: 64, or 256 with
                                          pipeline_overlap; < 0 = no chunking) */
    int32_t pipeline_streams;          /* internal worker streams, 1..4 (0 = default 4)           */
    int32_t depth_zero_copy;           /* host depth planes in pinned (page-locked) memory are not copied: the ~1000 depth
                                          samples a frame needs are read in place over PCIe by the kernel that unprojects
                                          the keypoints (0 = default on, -1 = always stage the whole plane in HBM)          */
    int32_t pipeline_overlap;          /* 1 = consecutive batched calls may overlap on the device (device inputs: see
                                          orbf_track_sequence_device_at).  Host inputs: a call's H2D copies and
                                          kernels are ordered only behind the copies / kernels of earlier calls on the same internal
                                          streams, not behind everything enqueued on the context stream.  The caller then alternates
                                          between disjoint frame / pair slot ranges (orbf_track_sequence_at) and does not reuse a range
                                          before it has read that range's results (0 = default: every call starts after the previous) */
    float k1, k2, p1, p2, k3;          /* Calibration::k1.. (Utils/common.h:40-44) as Frame::Frame loads them into mDistCoef
                                          (Core/frame.cpp:32-42).  k1 == 0 (the default, and what the synthetic benchmarks use)
                                          is the reference's own shortcut mvKeysUn = mvKeys (frame.cpp:288-291); otherwise
                                          every keypoint goes through cv::undistortPoints (frame.cpp:294-312) before mvuRight
                                          and mvKeys3Dc are formed from it (frame.cpp:157-162); the depth sample is still
                                          looked up at the distorted keypoint (frame.cpp:152-155, quirk Q11)              */
} orbf_config;

typedef struct {
    int32_t iterations;                /* Ransac(iters, minInlierTh, maxMahalanobisDist, sampleSize) */
    uint32_t min_inlier_th;
    float max_mahal;
    uint32_t sample_size;              /* <= 8 */
    int32_t check_depth;
    int32_t sort_mode;                 /* 0 = libstdc++ std::sort replay (reference), 1 = keep order, 2 = stable */
    double depth_cov;                  /* quirk Q7: process-wide depth covariance; < 0 => latch from the first
                                          scored pair of the call (batched: of pair 0) and return it            */
    uint32_t seed;                     /* sample tables: glibc srand(seed + pair_index) / rand() restated      */
} orbf_ransac_config;

typedef struct {
    int32_t ok;                        /* Ransac::Iterate's bool                                  */
    float rmse;
    float T12[16];                     /* row-major 4x4 (mT12)                                    */
    int32_t n_inliers, n_good, real_iters, valid_iters, used_identity;
    double depth_cov_used;
} orbf_ransac_result;

typedef struct {                       /* per-hypothesis trace (row k of the sample table), parity tests */
    int32_t n_refined, rounds;
    double refined_error;
    float T[16];
} orbf_hyp_trace;

typedef struct orbf_context orbf_context;

/* ---- lifecycle ------------------------------------------------------------------------------ */
int orbf_abi_version(void);
void orbf_default_config(orbf_config* cfg);                 /* 640x480, 1000 kp, 1.2, 8, 20, 7, FR1 intrinsics */
void orbf_default_ransac_config(orbf_ransac_config* cfg);   /* Ransac(200, 20, 3.0f, 4)                        */
int orbf_create(const orbf_config* cfg, orbf_context** out);
int orbf_destroy(orbf_context* ctx);
const char* orbf_status_string(int status);
const char* orbf_last_error(const orbf_context* ctx);       /* detail of the last ORBF_ERR_CUDA             */
int orbf_set_stream(orbf_context* ctx, void* cuda_stream);  /* run on a caller-owned cudaStream_t            */
int orbf_synchronize(orbf_context* ctx);
/* number of kernels this library launched on ctx since creation (bench.py's gpu_launches) */
int orbf_launch_count(const orbf_context* ctx, int64_t* n);

/* per-stage device timing with CUDA events on the context stream (off by default) */
int orbf_profile_enable(orbf_context* ctx, int32_t on);
int orbf_profile_collect(orbf_context* ctx);                 /* synchronises; call once per batch step */
int orbf_profile_read(const orbf_context* ctx, double* total_ms, int64_t* calls, int32_t cap /* >= 10 */);
const char* orbf_profile_stage_name(int32_t stage);

/* ---- extractor tables: GetScaleFactors / GetInverseScaleFactors / GetScaleSigmaSquares / ... ---- */
int orbf_get_tables(const orbf_context* ctx, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
    int32_t* nfeat_per_level, int32_t* level_w, int32_t* level_h);
int orbf_keypoint_capacity(const orbf_context* ctx, int32_t* cap);   /* max keypoints one frame can yield */

/* ---- extraction ------------------------------------------------------------------------------ */
/* Single frame, host in / host out (ORBextractor::operator()).  img may be pageable.  n_out always
 * receives the keypoint count; ORBF_ERR_CAPACITY if cap is too small.  Uses frame slot 0.          */
int orbf_extract(orbf_context* ctx, const uint8_t* img, int32_t width, int32_t height, int32_t stride,
    orbf_keypoint* kps, uint8_t* desc, int32_t cap, int32_t* n_out);
/* Batched, host input: copies n gray frames (+ optional u16 depth) into slots [slot0, slot0+n) on the
 * context stream and extracts them.  Asynchronous; results stay device-resident.                    */
int orbf_extract_batch(orbf_context* ctx, int32_t slot0, int32_t n, const uint8_t* gray, int64_t gray_stride,
    int64_t gray_frame_stride, const uint16_t* depth, int64_t depth_stride_elems, int64_t depth_frame_stride_elems);
/* Batched, device input (already in HBM): pointers must be 16-byte aligned, gray pitch a multiple of 16. */
int orbf_extract_batch_device(orbf_context* ctx, int32_t slot0, int32_t n, const uint8_t* d_gray, int64_t gray_pitch,
    int64_t gray_frame_stride, const uint16_t* d_depth, int64_t depth_pitch_elems, int64_t depth_frame_stride_elems);
/* Frame::Frame + ExtractFeatures from the colour image (Core/frame.cpp:18-45,135-170): n interleaved 8-bit BGR host frames
 * (row stride in bytes) -> gray on the device (cv::cvtColor CV_BGR2GRAY, OpenCV's fixed-point arithmetic) -> extraction.     */
int orbf_extract_batch_bgr(orbf_context* ctx, int32_t slot0, int32_t n, const uint8_t* bgr, int64_t bgr_stride,
    int64_t bgr_frame_stride, const uint16_t* depth, int64_t depth_stride_elems, int64_t depth_frame_stride_elems);
int orbf_download_gray(orbf_context* ctx, int32_t slot, uint8_t* out, int32_t out_stride);   /* mImGray of a slot (host-input paths) */
/* A whole sequence in one call: extraction of n frames plus, for each consecutive pair p = (slot0+p, slot0+p+1),
 * Matcher(ratio).KnnMatch and (ransac_cfg != NULL) Ransac::Iterate; results land in pair slots 0..n-2.  The call is
 * asynchronous and internally pipelined: frames are processed in chunks of orbf_config.pipeline_chunk on
 * orbf_config.pipeline_streams worker streams, so the H2D copy of one chunk overlaps the kernels of another (host
 * inputs only: with inputs already in HBM the stages run back to back on the context stream).                     */
int orbf_track_sequence(orbf_context* ctx, int32_t slot0, int32_t n, const uint8_t* gray, int64_t gray_stride,
    int64_t gray_frame_stride, const uint16_t* depth, int64_t depth_stride_elems, int64_t depth_frame_stride_elems, float ratio,
    int32_t cross_check, const orbf_ransac_config* ransac_cfg);
/* The same with the pair results in pair slots pair_slot0 .. pair_slot0 + n - 2 (orbf_track_sequence uses pair_slot0 = 0): lets a
 * caller double-buffer whole sequences in one context (frame slots [slot0, slot0 + n), pair slots [pair_slot0, ..)) so that, with
 * orbf_config.pipeline_overlap, the H2D copies of one call run under the kernels and result read-back of the previous one.        */
int orbf_track_sequence_at(orbf_context* ctx, int32_t slot0, int32_t pair_slot0, int32_t n, const uint8_t* gray, int64_t gray_stride,
    int64_t gray_frame_stride, const uint16_t* depth, int64_t depth_stride_elems, int64_t depth_frame_stride_elems, float ratio,
    int32_t cross_check, const orbf_ransac_config* ransac_cfg);
/* The device-input call with explicit pair slots.  With orbf_config.pipeline_overlap the RANSAC of such a call runs on an internal
 * high-priority stream and is NOT joined back into the context stream by the call itself: a following orbf_track_sequence_device_at on
 * DISJOINT frame / pair slots starts its pyramid / FAST kernels under it (RANSAC is a chain of latency-bound launches that leaves the
 * SMs almost idle).  Every other entry point — result downloads, orbf_synchronize, a call on overlapping slots — first orders the
 * context stream behind that RANSAC (no host wait); orbf_join does only that, e.g. before the caller records an event of its own.   */
int orbf_track_sequence_device_at(orbf_context* ctx, int32_t slot0, int32_t pair_slot0, int32_t n, const uint8_t* d_gray, int64_t gray_pitch,
    int64_t gray_frame_stride, const uint16_t* d_depth, int64_t depth_pitch_elems, int64_t depth_frame_stride_elems, float ratio,
    int32_t cross_check, const orbf_ransac_config* ransac_cfg);
int orbf_join(orbf_context* ctx);
/* Asynchronous result read-back for such double-buffered use: copies frame counts [slot0, +n), match counts and RANSAC results of pair
 * slots [pair_slot0, +n-1) into caller buffers (page-locked memory for a truly asynchronous copy; any may be NULL) on the context
 * stream and records marker `marker` (0..7) behind them; orbf_wait_marker blocks the host until that point has been reached.          */
int orbf_read_results_async(orbf_context* ctx, int32_t slot0, int32_t pair_slot0, int32_t n, int32_t* frame_counts, int32_t* match_counts,
    orbf_ransac_result* ransac, int32_t marker);
int orbf_wait_marker(orbf_context* ctx, int32_t marker);
/* The per-frame outputs the reference hands back to host vectors, for n frames at once and asynchronously: kps [n][K] in cv::KeyPoint
 * layout, desc [n][K][32], xyz [3][n][K] (mvKeys3Dc as x / y / z planes), matches [n - 1][K] (K = orbf_keypoint_capacity; rows past a
 * frame's / pair's count are unspecified; any pointer may be NULL), then marker `marker`.                                        */
int orbf_read_features_async(orbf_context* ctx, int32_t slot0, int32_t pair_slot0, int32_t n, orbf_keypoint* kps, uint8_t* desc, float* xyz,
    orbf_dmatch* matches, int32_t marker);
int orbf_track_sequence_device(orbf_context* ctx, int32_t slot0, int32_t n, const uint8_t* d_gray, int64_t gray_pitch,
    int64_t gray_frame_stride, const uint16_t* d_depth, int64_t depth_pitch_elems, int64_t depth_frame_stride_elems, float ratio,
    int32_t cross_check, const orbf_ransac_config* ransac_cfg);
/* Results of one slot -> host (synchronises the stream). xyz = mvKeys3Dc (N x 3 floats), may be NULL. */
int orbf_download_frame(orbf_context* ctx, int32_t slot, orbf_keypoint* kps, uint8_t* desc, float* xyz,
    int32_t cap, int32_t* n_out);
int orbf_frame_counts(orbf_context* ctx, int32_t slot0, int32_t n, int32_t* counts);
/* Frame::mvKeysUn of a slot: n interleaved (x, y) floats, the undistorted keypoint positions (= the keypoints when k1 == 0);
 * u_right (may be NULL) = Frame::mvuRight.                                                                              */
int orbf_download_keys_un(orbf_context* ctx, int32_t slot, float* xy, float* u_right, int32_t cap, int32_t* n_out);
/* Parity / mvImagePyramid access: level image, blurred level, FAST candidates (reference order). */
int orbf_pyramid_level(orbf_context* ctx, int32_t slot, int32_t level, int32_t blurred, uint8_t* out, int32_t out_stride);
int orbf_level_candidates(orbf_context* ctx, int32_t slot, int32_t level, orbf_cand* out, int32_t cap, int32_t* n_out);
int orbf_level_keypoint_counts(orbf_context* ctx, int32_t slot, int32_t* counts /* [nlevels] */);

/* ---- matching -------------------------------------------------------------------------------- */
/* Raw kNN-2 on host descriptor matrices (rows of 32 bytes): BFMatcher(NORM_HAMMING).knnMatch(k=2).
 * idx2/d2 = -1 when the train set has fewer than 2 rows.                                          */
int orbf_knn2(orbf_context* ctx, const uint8_t* q, int32_t nq, const uint8_t* t, int32_t nt, int32_t* idx1, int32_t* d1,
    int32_t* idx2, int32_t* d2);
/* kNN-2 + ratio test (m1.distance < ratio * m2.distance, float) + optional mutual-NN cross-check;
 * survivors in query order, as Matcher::KnnMatch builds them before its host-side landmark filter. */
int orbf_knn_match(orbf_context* ctx, const uint8_t* q, int32_t nq, const uint8_t* t, int32_t nt, float ratio,
    int32_t cross_check, orbf_dmatch* out, int32_t cap, int32_t* n_out);
int orbf_descriptor_distance(const uint8_t* a, const uint8_t* b, int32_t nbytes, int32_t* dist);  /* host helper */
/* Landmark::ComputeDistinctiveDescriptors (Core/landmark.cpp:219-273) for a batch of landmarks: desc = every landmark's observed
 * descriptors back to back (rows of 32 bytes), offsets [n_landmarks + 1]; best[l] = row inside landmark l with the least median
 * Hamming distance to the others (first wins ties, -1 without observations); median[l] optional (may be NULL).          */
int orbf_distinctive_descriptors(orbf_context* ctx, const uint8_t* desc, const int32_t* offsets, int32_t n_landmarks, int32_t* best,
    int32_t* median);
/* Frame::UndistortKeyPoints (Core/frame.cpp:286-313): cv::undistortPoints(pts, pts, K, dist, Mat(), K) with OpenCV's default 5 iterations;
 * xy / out = n interleaved (x, y) floats (may alias), dist = {k1, k2, p1, p2, k3}.  The same call on the four image corners gives
 * ComputeImageBounds (frame.cpp:320-343).  The extraction entry points apply the same arithmetic to their keypoints when
 * orbf_config.k1 != 0 (see there).                                                                                      */
int orbf_undistort_points(orbf_context* ctx, const float* xy, int32_t n, float fx, float fy, float cx, float cy, const float* dist, float* out);
/* The per-keypoint tail of Frame::ExtractFeatures (Core/frame.cpp:138-164) for keypoints that did not come out of the ORB extractor
 * (the adaptive-FAST route of Extractor): mvKeysUn (xy_un, n interleaved pairs; cv::undistortPoints when orbf_config.k1 != 0),
 * mvuRight (u_right, -1 without depth) and mvKeys3Dc (xyz, n interleaved triples, zeros without depth) from the keypoints and the
 * u16 depth plane (host; NULL = no depth; sampled at the truncated distorted position, scaled by orbf_config.depth_factor).            */
int orbf_unproject_keypoints(orbf_context* ctx, const orbf_keypoint* kps, int32_t n, const uint16_t* depth, int32_t width, int32_t height,
    int64_t depth_stride_elems, float* xyz, float* u_right, float* xy_un);
/* Matcher::ProjectionMatch (Features/matcher.cpp:90-143) for one frame: landmarks projected into the frame are matched, in order,
 * to the features inside the square window |dx| < radius && |dy| < radius; best <= th_high, and rejected when best and second best
 * share an octave and best > nn_ratio * second; a feature given to a landmark with Observations() > 0 is skipped by later landmarks.
 * Frame side: slot >= 0 takes keypoints / descriptors of that frame slot (output of orbf_extract_batch*; kp_* / desc ignored),
 * slot < 0 takes n_feat host rows kp_x / kp_y / kp_octave / desc.
 * lm_flags[i]: bit 0 = candidate (mbTrackInView && !isBad()), bit 1 = Observations() > 0.  feat_taken[j] (may be NULL) = feature j
 * already holds a landmark with Observations() > 0.  best_idx[i] = feature matched to landmark i or -1.                     */
int orbf_projection_match(orbf_context* ctx, int32_t slot, const float* kp_x, const float* kp_y, const int32_t* kp_octave, const uint8_t* desc,
    int32_t n_feat, const uint8_t* lm_desc, const float* proj_x, const float* proj_y, const uint8_t* lm_flags, int32_t n_landmarks,
    const uint8_t* feat_taken, float radius, float nn_ratio, int32_t th_high, int32_t* best_idx, int32_t* n_matches);
/* Matcher::Fuse (Features/matcher.cpp:212-296), the part that computes: landmark i (world position lm_pos[3i..], descriptor, lm_valid[i] =
 * pLM && !isBad() && !IsInKeyFrame(pKF)) is transformed by Rcw (row-major 3x3) / tcw, projected with camera = {fx, fy, cx, cy, mbf, mnMinX,
 * mnMaxX, mnMinY, mnMaxY}, and matched to the keyframe feature with the smallest Hamming distance inside the window that passes the
 * stereo (u_right[j] >= 0) / mono reprojection gate; best_idx[i] = that feature when the distance is <= th_low, else -1 (best_dist may
 * be NULL).  The Replace / AddObservation / AddLandmark that follows (matcher.cpp:297-311) edits the map graph and stays with the caller.
 * slot >= 0: keyframe features from that frame slot (kp_x / kp_y / u_right / desc ignored); slot < 0: n_feat host rows.            */
int orbf_fuse_search(orbf_context* ctx, int32_t slot, const float* Rcw, const float* tcw, const float* camera, const float* kp_x, const float* kp_y,
    const float* u_right, const uint8_t* desc, int32_t n_feat, const float* lm_pos, const uint8_t* lm_desc, const uint8_t* lm_valid,
    int32_t n_landmarks, float radius, int32_t th_low, int32_t* best_idx, int32_t* best_dist);
/* Matcher::BoWMatch (Features/matcher.cpp:145-209).  The two DBoW3 feature vectors arrive flattened: words[nw] ascending node ids,
 * off[nw + 1] bucket offsets, idx[off[nw]] feature indices in bucket order; desc1 / desc2 = the keyframes' descriptor matrices
 * (n1 / n2 rows of 32 bytes).  Survivors (best <= th_low, (float)best < nn_ratio * (float)second, train feature not used by an earlier
 * query) in the reference's order; imgIdx = -1 as a default-constructed cv::DMatch has it.                                     */
int orbf_bow_match(orbf_context* ctx, const int32_t* words1, const int32_t* off1, const int32_t* idx1, int32_t nw1, const uint8_t* desc1, int32_t n1,
    const int32_t* words2, const int32_t* off2, const int32_t* idx2, int32_t nw2, const uint8_t* desc2, int32_t n2, float nn_ratio, int32_t th_low,
    orbf_dmatch* out, int32_t cap, int32_t* n_out);
/* Device-resident: match frame slot pairs (query_slot, train_slot); results live in pair slots 0..npairs-1. */
int orbf_match_pairs(orbf_context* ctx, const int32_t* pairs /* 2*npairs */, int32_t npairs, float ratio,
    int32_t cross_check);
int orbf_download_matches(orbf_context* ctx, int32_t pair, orbf_dmatch* out, int32_t cap, int32_t* n_out);
int orbf_download_knn(orbf_context* ctx, int32_t pair, int32_t* idx1, int32_t* d1, int32_t* idx2, int32_t* d2, int32_t cap,
    int32_t* nq_out);
int orbf_match_counts(orbf_context* ctx, int32_t npairs, int32_t* counts);

/* ---- RANSAC / Kabsch ------------------------------------------------------------------------- */
/* Host in / host out Ransac::Iterate.  sample_table: iterations x sample_size ascending ids into the
 * sorted good-match list (-1 padded) or NULL => glibc rand() restated from cfg->seed.
 * Optional traces (may be NULL): per-hypothesis, the sorted good matches, the sample table used.   */
int orbf_ransac_iterate(orbf_context* ctx, const orbf_ransac_config* cfg, const float* src_xyz, int32_t nsrc,
    const float* dst_xyz, int32_t ndst, const orbf_dmatch* m12, int32_t nm, const int32_t* sample_table,
    orbf_dmatch* inliers_out, int32_t cap, orbf_ransac_result* out, orbf_hyp_trace* hyp_trace,
    orbf_dmatch* good_sorted_out, int32_t* sample_table_out);
/* Device-resident: RANSAC on the pairs last matched by orbf_match_pairs. */
int orbf_ransac_pairs(orbf_context* ctx, int32_t npairs, const orbf_ransac_config* cfg);
/* Quirk Q7 across shards: the covariance these pairs would latch on a fresh process (first pair, in order, that reaches scoring;
 * -1 if none does).  Neither reads nor writes the value latched on the context.                                              */
int orbf_ransac_probe_depth_cov(orbf_context* ctx, int32_t npairs, const orbf_ransac_config* cfg, double* cov);
int orbf_download_ransac(orbf_context* ctx, int32_t pair, orbf_ransac_result* out, orbf_dmatch* inliers, int32_t cap);
int orbf_download_ransac_summary(orbf_context* ctx, int32_t npairs, orbf_ransac_result* out /* [npairs] */);
/* Ransac::mpSourceCloud / mpTargetCloud (Odometry/ransac.h:67-68, filled at Odometry/ransac.cpp:163-189 for the GICP refinement that
 * follows): for every pair last solved (orbf_ransac_pairs / orbf_track_sequence* / orbf_ransac_iterate = pair 0), the 3D points of its
 * depth-valid matches in m12 order, as pcl::PointXYZ records (16 bytes: x, y, z, 1.0f), source frame and target frame; empty when the
 * pair had fewer than min_inlier_th matches (the early return at ransac.cpp:166-167).  orbf_ransac_clouds fills the device-resident
 * buffers for pairs [pair0, pair0 + npairs) and hands out their base pointers: cloud of pair p at base + p * points_per_pair * 4 floats,
 * its length at d_counts[p] — valid until the next RANSAC call on the context (asynchronous on the context stream).
 * orbf_download_ransac_clouds copies one pair's clouds to host arrays of cap points each.                                              */
int orbf_ransac_clouds(orbf_context* ctx, int32_t pair0, int32_t npairs, const float** d_src_xyzw, const float** d_tgt_xyzw,
    const int32_t** d_counts, int32_t* points_per_pair);
int orbf_download_ransac_clouds(orbf_context* ctx, int32_t pair, float* src_xyzw, float* tgt_xyzw, int32_t cap, int32_t* n_out);
int orbf_kabsch(orbf_context* ctx, const float* setA, const float* setB, int32_t n, float* T16);
/* Odometry::Compute, RANSAC strategy (Odometry/odometry.cpp:78-90) for the npairs consecutive pairs last solved by orbf_ransac_pairs /
 * orbf_track_sequence: poses[0] = pose0 (row-major 4x4, NULL = identity), poses[k + 1] = T12[k] * poses[k] with cv::Mat's float product;
 * outlier (may be NULL) = [npairs + 1][keypoint capacity] bytes, Frame::mvbOutlier after SetInlier(m.trainIdx) for the pair's inliers
 * (row 0, the first frame, stays all 1).                                                                              */
int orbf_compose_trajectory(orbf_context* ctx, int32_t npairs, const float* pose0, float* poses /* [(npairs + 1) * 16] */, uint8_t* outlier);
/* Odometry::Compute, RANSAC strategy (Odometry/odometry.cpp:44-90) for ONE frame pair, host in / host out, in one call with one
 * synchronisation: Ransac::Iterate(F1, F2, m12) as orbf_ransac_iterate runs it (glibc rand() from cfg->seed), the clouds it leaves for GICP
 * (cloud_*_xyzw: cloud_cap pcl::PointXYZ records each, may be NULL; *n_cloud = their length) and pose2 = T12 * pose1 with cv::Mat's float
 * product (pose1 NULL = identity; pose2 may be NULL).  The inlier flags are the caller's SetInlier(m.trainIdx) over inliers_out.      */
int orbf_odometry_compute(orbf_context* ctx, const orbf_ransac_config* cfg, const float* src_xyz, int32_t nsrc, const float* dst_xyz, int32_t ndst,
    const orbf_dmatch* m12, int32_t nm, orbf_dmatch* inliers_out, int32_t cap, orbf_ransac_result* out, float* cloud_src_xyzw, float* cloud_tgt_xyzw,
    int32_t cloud_cap, int32_t* n_cloud, const float* pose1, float* pose2);

/* ---- multi-GPU frame sharding (one process per GPU, no data-path collective) --------- */
/* Host arithmetic, no context: the contiguous chunk [start, stop) of `rank` among `world` ranks, the halo frame (the previous rank's
 * last frame, extracted again so that the pair straddling two chunks has an owner), the first frame the rank extracts (start - halo)
 * and the global pair indices [pair0, pair1) it owns; any output may be NULL.  A C++ host shards a sequence with this call,
 * orbf_extract_batch / orbf_match_pairs on its frames, orbf_ransac_probe_depth_cov + its own 8-byte all-gather (the first valid value
 * in rank order = quirk Q7's latch), orbf_ransac_pairs with that covariance and seed + pair0, and chains the absolute poses with
 * orbf_compose_trajectory(pose0 = the previous rank's last pose).                                                          */
int orbf_frame_shard(int32_t n_frames, int32_t world, int32_t rank, int32_t* start, int32_t* stop, int32_t* halo, int32_t* first,
    int32_t* pair0, int32_t* pair1);

/* ---- keyframe descriptor store (Core/keyframedatabase, BASELINE config 5) --------------------- */
/* Copies the descriptors of frame slot `slot` into keyframe entry `kf` of the device-resident store
 * (capacity max_keyframes, set on first use); many-to-many matching runs a-16 against every entry.  */
int orbf_kfdb_reserve(orbf_context* ctx, int32_t max_keyframes);
int orbf_kfdb_add_from_slot(orbf_context* ctx, int32_t kf, int32_t slot);
int orbf_kfdb_add_host(orbf_context* ctx, int32_t kf, const uint8_t* desc, int32_t n);
int orbf_kfdb_device_buffers(orbf_context* ctx, uint8_t** d_desc, int32_t** d_counts, int32_t* rows_per_kf,
    int32_t* n_kf);
/* Match against a caller-owned device store with the same layout ([n_kf][rows_per_kf][32] u8 + [n_kf] i32) — the
 * NCCL all-gather of every rank's shard (BASELINE config 5).  Not owned by the context; NULL detaches.           */
int orbf_kfdb_attach_device(orbf_context* ctx, const uint8_t* d_desc, const int32_t* d_counts, int32_t n_kf);
/* query (host, nq x 32) against keyframes [kf0, kf0+nkf): per keyframe top-2 per query and the number of
 * ratio survivors (out arrays sized nkf x nq, counts sized nkf).  With all four table pointers NULL only the survivor
 * counts are produced and downloaded (keyframe ranking).                                                */
int orbf_kfdb_match(orbf_context* ctx, const uint8_t* q, int32_t nq, int32_t kf0, int32_t nkf, float ratio,
    int32_t* idx1, int32_t* d1, int32_t* idx2, int32_t* d2, int32_t* survivors);

/* Device-resident query (BASELINE config 5 at scale): the descriptors of frame slot `slot` against keyframes [kf0, kf0 + nkf) of the
 * current store; asynchronous.  orbf_kfdb_results returns the tables / survivor counts of the last match (synchronises).          */
int orbf_kfdb_match_slot(orbf_context* ctx, int32_t slot, int32_t kf0, int32_t nkf, float ratio);
int orbf_kfdb_results(orbf_context* ctx, int32_t nkf, int32_t nq, int32_t* idx1, int32_t* d1, int32_t* idx2, int32_t* d2, int32_t* survivors);

/* ---- multi-GPU keyframe store: one process per GPU, every rank holds a shard of the keyframes -------------------
 * (a) NCCL: rank 0 draws a 128-byte id (orbf_comm_unique_id), the host application hands it to every rank by any means, every rank
 *     calls orbf_comm_init; orbf_kfdb_allgather then all-gathers the stores over NVLink (ncclAllGather, uint8) into a gathered copy
 *     that orbf_kfdb_match* read: keyframe g = rank * capacity + local index.  NCCL is loaded with dlopen on first use.
 * (b) no collective: every rank exports its store (orbf_kfdb_ipc_handles, 2 x 64 bytes), the application exchanges the handles, and
 *     orbf_kfdb_attach_peers makes the matcher read every keyframe's rows from the GPU that owns it over NVLink (CUDA IPC).       */
int orbf_comm_unique_id(uint8_t* id128);
int orbf_comm_init(orbf_context* ctx, const uint8_t* id128, int32_t nranks, int32_t rank);
int orbf_comm_destroy(orbf_context* ctx);
int orbf_kfdb_allgather(orbf_context* ctx, int32_t* n_kf_total);
int orbf_kfdb_ipc_handles(orbf_context* ctx, uint8_t* desc_handle64, uint8_t* count_handle64);
int orbf_kfdb_attach_peers(orbf_context* ctx, const uint8_t* desc_handles, const uint8_t* count_handles, int32_t nranks, int32_t rank,
    int32_t kf_per_rank);
int orbf_kfdb_detach_peers(orbf_context* ctx);

/* ---- adaptive-threshold FAST detector (Extractor mode ADAPTIVE with the FAST detector) ---------- */
typedef struct {
    int32_t min_features, max_features, max_iters, max_per_cell, grid, edge;   /* 67, 113, 5, 113, 3, 31 (extractor.cpp:65-77)   */
    double init_th, min_th, max_th, inc, dec;                                  /* 20, 2, 10000, 1.3, 0.7 (extractor.cpp:56)      */
    int32_t retain_best;                                                        /* Extract(): retainBest(nFeatures); 0 = off      */
} orbf_adaptive_config;
void orbf_default_adaptive_config(orbf_adaptive_config* cfg);
/* n host frames (gray, context width x height) through the grid of stateful detectors, in order: `thresh` [grid*grid] is
 * the per-cell threshold state (<= 0 => init_th), updated in place so the next call continues the video.  out [n][cap]
 * keypoints (cv::FAST KeyPoint fields: size 7, angle -1, octave 0), counts [n]; optional per-frame, per-cell integer
 * threshold of the final detection and number of keypoints it found (before keepStrongest).                         */
int orbf_adaptive_detect(orbf_context* ctx, const orbf_adaptive_config* cfg, const uint8_t* gray, int32_t n, int32_t stride,
    int64_t frame_stride, double* thresh, orbf_keypoint* out, int32_t* counts, int32_t cap, int32_t* cell_thresh, int32_t* cell_found);

/* BASELINE config 4, 8-level variant (a north-star extension — the reference's adaptive routes are single-scale):
 * ORB extraction of n host frames, in order, into frame slots [slot0, slot0 + n), with iniThFAST of every FAST cell replaced by the
 * state of its image region's controller: threshold = max(minThFAST, min(254, (int)state[region])), and after each frame
 * found[region] < cfg->min_features => state *= dec, > cfg->max_features => state *= inc (clamped to [min_th, max_th]); regions =
 * cfg->grid x cfg->grid partition of the image.  thresh [grid * grid] in / out (<= 0 => init_th) continues the video across calls;
 * region_th / region_found (optional, [n][grid * grid]): thresholds each frame was detected with, keypoints it returned per region.
 * Results are read like those of orbf_extract_batch (orbf_download_frame, matching, ...).                                        */
int orbf_extract_adapted(orbf_context* ctx, int32_t slot0, int32_t n, const uint8_t* gray, int64_t gray_stride, int64_t gray_frame_stride,
    const uint16_t* depth, int64_t depth_stride_elems, int64_t depth_frame_stride_elems, const orbf_adaptive_config* cfg, double* thresh,
    int32_t* region_th, int32_t* region_found);
/* The same for n_videos INDEPENDENT videos (cameras / sequences) of frames_per_video frames each: a video is sequential by definition
 * (frame t + 1's thresholds depend on frame t's keypoints), several videos advance together, V frames per step of the chain.
 * gray: frame t of video v at gray + v * gray_video_stride + t * gray_frame_stride; results: frame t of video v in slot
 * slot0 + t * n_videos + v.  thresh [n_videos][grid * grid] in / out; region_th / region_found [frames_per_video][n_videos][grid * grid]. */
int orbf_extract_adapted_videos(orbf_context* ctx, int32_t slot0, int32_t n_videos, int32_t frames_per_video, const uint8_t* gray, int64_t gray_stride,
    int64_t gray_frame_stride, int64_t gray_video_stride, const uint16_t* depth, int64_t depth_stride_elems, int64_t depth_frame_stride_elems,
    int64_t depth_video_stride_elems, const orbf_adaptive_config* cfg, double* thresh, int32_t* region_th, int32_t* region_found);

#ifdef __cplusplus
}
#endif
#endif /* ORBFRONT_H */
