/* oracle/oracle_api.h — C interface of the CPU oracle (TEST INFRASTRUCTURE ONLY).
 * Loaded with ctypes by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.  Never linked into, or called from, the product path.          */
#ifndef ORB_ORACLE_API_H
#define ORB_ORACLE_API_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_LEVELS 16
enum { ORC_OK = 0, ORC_ERR_ARG = 1, ORC_ERR_CAPACITY = 2, ORC_ERR_GEOMETRY = 3 };

/* POD mirror of cv::KeyPoint (28 B) and cv::DMatch (16 B) */
typedef struct { float x, y, size, angle, response; int octave, class_id; } orc_keypoint;
typedef struct { int queryIdx, trainIdx, imgIdx; float distance; } orc_dmatch;
/* FAST candidate: pixel coordinates relative to minBorder (16) and cv::FAST response */
typedef struct { int x, y, score; } orc_cand;

typedef struct {
    int width, height, nfeatures, nlevels;
    float scale_factor;
    int ini_th_fast, min_th_fast;
} orc_extract_cfg;

/* Optional per-stage dumps of orc_extract; any pointer may be NULL. */
typedef struct {
    uint8_t* pyramid;    /* all levels back to back, tight rows                         */
    uint8_t* blurred;    /* same layout; only levels with >=1 keypoint are written      */
    orc_cand* cands;     /* FAST candidates of all levels back to back (reference order) */
    int cand_cap;
    int* n_cands;        /* [nlevels] */
    int* n_kps;          /* [nlevels] keypoints kept by the quadtree                    */
    int* level_xy;       /* [2*n] integer level coordinates of the output keypoints     */
} orc_extract_debug;

int orc_tables(int nfeatures, float scaleFactor, int nlevels, float* scale, float* inv_scale, float* sigma2,
    float* inv_sigma2, int* nfeat_per_level, int* umax16);
int orc_level_sizes(int w, int h, float scaleFactor, int nlevels, int* ws, int* hs);
int orc_resize_linear(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride);
int orc_pyramid(const uint8_t* img, int w, int h, int stride, float scaleFactor, int nlevels, uint8_t* out);
int orc_fast_roi(const uint8_t* roi, int stride, int w, int h, int th, orc_cand* out, int cap, int* n);
int orc_fast_strength_map(const uint8_t* img, int w, int h, int stride, int16_t* out);
int orc_fast_cells(const uint8_t* img, int w, int h, int stride, int iniTh, int minTh, orc_cand* out, int cap, int* n);
int orc_distribute(const orc_cand* cands, int n, int minX, int maxX, int minY, int maxY, int N, int* out_idx, int cap,
    int* n_out);
float orc_fast_atan2(float y, float x);
int orc_ic_angle(const uint8_t* img, int w, int h, int stride, const int* xs, const int* ys, int n, float* angles);
int orc_gaussian_blur7(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride);
int orc_rbrief(const uint8_t* blurred, int w, int h, int stride, const int* xs, const int* ys, const float* angles, int n,
    uint8_t* desc);
const int8_t* orc_pattern(void);
int orc_extract(const orc_extract_cfg* cfg, const uint8_t* img, int stride, orc_keypoint* kps, uint8_t* desc, int cap,
    int* n_out, orc_extract_debug* dbg);
int orc_unproject(const orc_keypoint* kps, int n, const uint16_t* depth_u16, const float* depth_f32, int w, int h,
    int dstride_elems, float depth_factor, float fx, float fy, float cx, float cy, float mbf, float* xyz, float* uright);
int orc_unproject_un(const orc_keypoint* kps, const float* xy_un, int n, const uint16_t* depth_u16, const float* depth_f32, int w, int h,
    int dstride_elems, float depth_factor, float fx, float fy, float cx, float cy, float mbf, float* xyz, float* uright);

/* ---- adaptive FAST detector route (Features/video*adaptedfeaturedetector.cpp, detectoradjuster.cpp, extractor.cpp:52-77) ---- */
typedef struct {
    int min_features, max_features, max_iters, max_per_cell, grid, edge;
    double init_th, min_th, max_th, inc, dec;
} orc_adaptive_cfg;
int orc_adaptive_default(orc_adaptive_cfg* cfg);
/* thresh: [grid*grid] per-cell threshold state, in/out (persists from frame to frame like the stateful detectors).
 * cell_found / cell_thresh (optional): keypoints found by, and integer threshold of, each cell's last detection.  */
int orc_adaptive_detect(const orc_adaptive_cfg* cfg, const uint8_t* img, int w, int h, int stride, double* thresh, int retain_best,
    orc_keypoint* out, int cap, int* n_out, int* cell_found, int* cell_thresh);

/* BASELINE config 4, 8-level variant: ORB extraction with iniThFAST adapted per grid region by DetectorAdjuster-style controllers that
 * carry their state from frame to frame (defined in orb_oracle.cpp).  min_features / max_features of acfg are
 * the per-region band, grid the partition; thresh [grid * grid] in / out. */
int orc_extract_adapted(const orc_extract_cfg* cfg, const orc_adaptive_cfg* acfg, const uint8_t* img, int stride, double* thresh, orc_keypoint* kps,
    uint8_t* desc, int cap, int* n_out, int* region_th, int* region_found);

/* Landmark::ComputeDistinctiveDescriptors (Core/landmark.cpp:219-273) for a batch of landmarks */
int orc_distinctive_descriptors(const uint8_t* desc, const int* offsets, int n_landmarks, int* best, int* best_median);
/* Matcher::Fuse, search part (Features/matcher.cpp:212-296) and Matcher::BoWMatch (Features/matcher.cpp:145-209) */
int orc_fuse_search(const float* Rcw, const float* tcw, float fx, float fy, float cx, float cy, float mbf, float min_x, float max_x, float min_y, float max_y,
    const float* kp_x, const float* kp_y, const float* u_right, const uint8_t* desc, int n_feat, const float* lm_pos, const uint8_t* lm_desc,
    const uint8_t* lm_valid, int n_landmarks, float radius, double th_low, int* best_idx, int* best_dist);
int orc_bow_match(const int* words1, const int* off1, const int* idx1, int nw1, const uint8_t* desc1, const int* words2, const int* off2, const int* idx2,
    int nw2, const uint8_t* desc2, float nn_ratio, double th_low, orc_dmatch* out, int cap, int* n_out);
/* Odometry::Compute composition rule (Odometry/odometry.cpp:82-84) along a sequence */
int orc_compose_trajectory(const float* T12, int npairs, const float* pose0, float* poses);
/* Frame::UndistortKeyPoints (Core/frame.cpp:286-313): cv::undistortPoints with P = K, dist = {k1, k2, p1, p2, k3} */
int orc_undistort_points(const float* xy, int n, float fx, float fy, float cx, float cy, const float* dist, float* out);
/* Matcher::ProjectionMatch (Features/matcher.cpp:90-143) */
int orc_projection_match(const float* kp_x, const float* kp_y, const int* kp_octave, const uint8_t* desc, int n_feat, const uint8_t* lm_desc,
    const float* proj_x, const float* proj_y, const uint8_t* lm_flags, int n_landmarks, const uint8_t* feat_taken, float radius, float nn_ratio,
    double th_high, int* best_idx, int* n_matches);

/* ---- matching (Features/matcher.cpp:10-88,355-358) ---- */
int orc_hamming(const uint8_t* a, const uint8_t* b);
int orc_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, int* idx1, int* d1, int* idx2, int* d2);
int orc_knn_match(const uint8_t* q, int nq, const uint8_t* t, int nt, float ratio, int cross_check, orc_dmatch* out,
    int cap, int* n);

/* ---- RANSAC / Kabsch (Odometry/ransac.cpp, Odometry/kabsch.cpp) ---- */
typedef struct {
    int iterations;
    unsigned min_inlier_th;
    float max_mahal;
    unsigned sample_size;
    int check_depth;
    double depth_cov;   /* quirk Q7: explicit depth covariance; < 0 => latch from first scored pair */
} orc_ransac_cfg;

typedef struct {
    int ok;             /* Iterate()'s return value                      */
    float rmse;
    float T12[16];      /* row-major 4x4                                  */
    int n_inliers, n_good, real_iters, valid_iters, used_identity;
    double depth_cov_used;
} orc_ransac_out;

typedef struct {        /* per executed-or-not hypothesis k (row k of the sample table) */
    int n_refined;      /* vRefinedMatches.size() */
    int rounds;
    double refined_error;
    float T[16];
} orc_hyp_debug;

/* sort_mode: 0 = std::sort (libstdc++ introsort, what the reference runs), 1 = keep given order,
 * 2 = std::stable_sort.  sample_table: iterations x sample_size ascending ids (-1 padded) into the
 * sorted good-match list, or NULL => srand(seed) + libc rand() per reference rule a-12.            */
int orc_ransac_iterate(const orc_ransac_cfg* cfg, const float* src_xyz, int nsrc, const float* dst_xyz, int ndst,
    const orc_dmatch* m12, int nm, int sort_mode, const int* sample_table, unsigned seed, orc_dmatch* inliers_out,
    int cap, orc_ransac_out* out, orc_hyp_debug* per_hyp, orc_dmatch* good_sorted_out, int* sample_table_out);
int orc_sample_table_libc(unsigned seed, int M, int iterations, int sample_size, int* table);
int orc_libc_rand_sequence(unsigned seed, int n, int* out);
int orc_std_sort_dmatch(orc_dmatch* m, int n);
int orc_svd3(const float* A, float* U, float* S, float* V);
int orc_weighted_transform(const float* src_xyz, const float* dst_xyz, int n, float* T16);
double orc_mahalanobis2(const float* p1, const float* p2, const float* T16, double depth_cov);
/* the two third-party numerical routines of the path on their own (shared with the oracle/_ref stand-ins for PCL / Eigen) */
int orc_tfc_transform(const float* p_xyz, const float* q_xyz, const float* w, int n, float* T16);
int orc_llt3_solve(const double* S9, const double* b3, double* x3);
/* sensitivity probe: 1 = evaluate the 3- / 4-term inner sums of the PCL / Eigen restatements as Eigen's balanced tree; returns the old setting */
int orc_set_sum_order(int tree);
int orc_kabsch(const float* setA, const float* setB, int n, float* T16);

#ifdef __cplusplus
}
#endif
#endif
