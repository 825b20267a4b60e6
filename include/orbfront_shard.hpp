// include/orbfront_shard.hpp — C++ host logic of the multi-GPU sequence path (one process per GPU), header-only, on top of
// the C ABI in orbfront.h.  The same steps as adaptive-rgbd-localization-mappig_b200/sharding.py (which the tests and bench.py drive
// from Python), for a host application written in the reference's language:
//
//   * a sequence of n frames is cut into `world` contiguous chunks (orbf_frame_shard); rank r > 0 also extracts the last frame of rank
//     r - 1 (the halo frame) so that the pair straddling two chunks has an owner: extraction is independent per frame
//     (Features/orbextractor.cpp:756-815), matching + RANSAC per consecutive pair (System/tracking.cpp:193-208) — no data-path collective;
//   * quirk Q7: the reference latches its depth covariance in a function-local static on the first pair it ever scores
//     (Odometry/ransac.cpp:416-431).  Every rank probes the value its own pairs would latch (orbf_ransac_probe_depth_cov), the 8-byte
//     candidates are all-gathered, and the first valid one in rank order — the globally first pair that scores — is what every rank
//     scores with;
//   * quirk Q5: global pair p draws its samples from seed + p, exactly as one process running the whole sequence would;
//   * Odometry::Compute's composition rule pose[k + 1] = T12[k] * pose[k] (Odometry/odometry.cpp:82-84) is a chain of float products,
//     which are not associative: rank r starts from the pose of its first frame (= the last pose of rank r - 1), passed from rank to
//     rank, 64 bytes per hop.
//
// The two exchanges are callbacks (Exchange): the application already has a transport between its processes (MPI, NCCL, sockets) and
// this library does not impose one.  tests/cpp/shard_host_logic.cpp runs the class over a recording stand-in of the C ABI with the ranks
// as threads (CPU test: shards together == one process, for every world size and for shards whose first pairs do not score).
#pragma once
#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include "orbfront.h"

namespace orbf {

struct ShardError : std::runtime_error {
    int status;
    ShardError(int s, const std::string& what) : std::runtime_error(what + ": " + orbf_status_string(s)), status(s) {}
};

struct FrameShard {
    int32_t start = 0, stop = 0;       // the rank's chunk [start, stop) of the sequence
    int32_t halo = 0;                  // 1 when the rank also extracts frame start - 1
    int32_t first = 0;                 // first frame the rank extracts (start - halo)
    int32_t pair0 = 0, pair1 = 0;      // global pairs [pair0, pair1) it owns; pair p = frames (p, p + 1)
    int32_t frames() const { return stop - first; }
    int32_t pairs() const { return pair1 - pair0; }
};

inline FrameShard frame_shard(int32_t n_frames, int32_t world, int32_t rank)
{
    FrameShard s;
    const int st = orbf_frame_shard(n_frames, world, rank, &s.start, &s.stop, &s.halo, &s.first, &s.pair0, &s.pair1);
    if (st != ORBF_OK) throw ShardError(st, "orbf_frame_shard");
    return s;
}

// What crosses ranks, supplied by the host application.
struct Exchange {
    // every rank contributes one double and receives all of them in rank order (8 bytes per rank); may be empty when world == 1
    std::function<void(double local, double* all /* [world] */)> allgather_f64;
    // 16 floats to rank dst / from rank src (blocking); may be empty when world == 1
    std::function<void(const float* pose16, int32_t dst)> send_pose;
    std::function<void(float* pose16, int32_t src)> recv_pose;
};

class SequenceShard {
public:
    SequenceShard(orbf_context* ctx, int32_t n_frames_total, int32_t rank, int32_t world)
        : ctx_(ctx), rank_(rank), world_(world), shard_(frame_shard(n_frames_total, world, rank)) {}

    const FrameShard& shard() const { return shard_; }
    double depth_cov() const { return cov_; }

    // Extract + match + RANSAC of this rank's frames.  gray / depth hold exactly the frames [shard().first, shard().stop) in order (host
    // memory, layout as orbf_extract_batch takes it; depth may be NULL); they go into frame slots 0 .., the pairs into pair slots 0 ...
    // cfg.seed is the seed of the whole sequence, cfg.depth_cov >= 0 an explicit covariance (no exchange then).
    void run(const uint8_t* gray, int64_t gray_stride, int64_t gray_frame_stride, const uint16_t* depth, int64_t depth_stride_elems,
        int64_t depth_frame_stride_elems, float ratio, bool cross_check, const orbf_ransac_config& cfg, const Exchange& x)
    {
        const int32_t n = shard_.frames(), np = shard_.pairs();
        if (n > 0) chk(orbf_extract_batch(ctx_, 0, n, gray, gray_stride, gray_frame_stride, depth, depth_stride_elems, depth_frame_stride_elems), "orbf_extract_batch");
        if (np > 0) {
            std::vector<int32_t> pairs(2 * (size_t)np);
            for (int32_t k = 0; k < np; ++k) { pairs[2 * k] = k; pairs[2 * k + 1] = k + 1; }
            chk(orbf_match_pairs(ctx_, pairs.data(), np, ratio, cross_check ? 1 : 0), "orbf_match_pairs");
        }
        orbf_ransac_config c = cfg;
        c.seed = cfg.seed + (uint32_t)shard_.pair0;              // the library seeds pair slot k with seed + k: global pair p gets seed + p
        if (cfg.depth_cov >= 0.0) cov_ = cfg.depth_cov;
        else {
            double local = -1.0;
            if (np > 0) chk(orbf_ransac_probe_depth_cov(ctx_, np, &c, &local), "orbf_ransac_probe_depth_cov");
            cov_ = local;
            if (world_ > 1) {                                    // every rank takes part, also one without pairs
                if (!x.allgather_f64) throw ShardError(ORBF_ERR_ARG, "Exchange::allgather_f64 missing");
                std::vector<double> all((size_t)world_, -1.0);
                x.allgather_f64(local, all.data());
                cov_ = -1.0;
                for (double v : all) if (v >= 0.0) { cov_ = v; break; }
            }
        }
        c.depth_cov = cov_;                                      // < 0 (no pair of the sequence scores): nothing is scored anywhere
        if (np > 0) chk(orbf_ransac_pairs(ctx_, np, &c), "orbf_ransac_pairs");
        ran_ = true;
    }

    // Ransac results of the rank's pairs, in order (result k = global pair shard().pair0 + k).
    std::vector<orbf_ransac_result> results() const
    {
        need_run();
        std::vector<orbf_ransac_result> r((size_t)shard_.pairs());
        if (!r.empty()) chk(orbf_download_ransac_summary(ctx_, (int32_t)r.size(), r.data()), "orbf_download_ransac_summary");
        return r;
    }

    // Absolute poses of the frames this rank OWNS ([start, stop), row-major 4x4 each), composed down the ranks.  pose0 = pose of frame 0
    // of the sequence (NULL = identity), read by rank 0 only.  Blocking: rank r waits for rank r - 1.
    std::vector<float> compose(const float* pose0, const Exchange& x) const
    {
        need_run();
        float start[16];
        for (int i = 0; i < 16; ++i) start[i] = pose0 && rank_ == 0 ? pose0[i] : (i % 5 == 0 ? 1.f : 0.f);
        if (world_ > 1 && rank_ > 0) {
            if (!x.recv_pose) throw ShardError(ORBF_ERR_ARG, "Exchange::recv_pose missing");
            x.recv_pose(start, rank_ - 1);                       // pose of frame first + halo - 1 ... = the previous rank's last frame
        }
        const int32_t n = shard_.frames(), np = shard_.pairs();
        std::vector<float> poses;
        if (n > 0) {
            poses.assign((size_t)n * 16, 0.f);
            if (np > 0) chk(orbf_compose_trajectory(ctx_, np, start, poses.data(), nullptr), "orbf_compose_trajectory");
            else for (int i = 0; i < 16; ++i) poses[i] = start[i];          // a one-frame shard (rank 0 of a tiny sequence)
        }
        if (world_ > 1 && rank_ + 1 < world_) {
            if (!x.send_pose) throw ShardError(ORBF_ERR_ARG, "Exchange::send_pose missing");
            x.send_pose(n > 0 ? poses.data() + (size_t)(n - 1) * 16 : start, rank_ + 1);   // an empty shard passes the pose on
        }
        if (shard_.halo && n > 0) poses.erase(poses.begin(), poses.begin() + 16);   // the halo frame's pose belongs to the previous rank
        return poses;
    }

private:
    static void chk(int st, const char* what) { if (st != ORBF_OK) throw ShardError(st, what); }
    void need_run() const { if (!ran_) throw ShardError(ORBF_ERR_STATE, "SequenceShard::run has not been called"); }
    orbf_context* ctx_;
    int32_t rank_, world_;
    FrameShard shard_;
    double cov_ = -1.0;
    bool ran_ = false;
};

}  // namespace orbf
