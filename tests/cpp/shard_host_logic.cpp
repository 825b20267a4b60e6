// tests/cpp/shard_host_logic.cpp — include/orbfront_shard.hpp (the C++ host logic of the multi-GPU sequence path) run on the CPU.
//
// The device calls of the C ABI are replaced, in THIS executable only, by recording stand-ins with the semantics orbfront.h documents
// (pair slot k is seeded with cfg.seed + k; depth_cov < 0 latches from the first pair of the call that scores; the probe returns that
// value without touching anything; the composition rule is a chain of float products).  orbf_frame_shard and orbf_status_string are the
// REAL ones from liborbfront_b200.so (host arithmetic, no device).  The ranks of a job are threads; the two exchanges of
// orbf::Exchange go through shared memory.  Checked, for sequences of 0 .. 41 frames on 1 .. 8 ranks, with leading pairs that do not
// reach scoring (quirk Q7's latch then belongs to a later pair, possibly of a later rank) and with sequences where no pair scores:
//   every frame is extracted by its owner (+ the halo), every pair 0 .. n - 2 is solved exactly once, with seed + p and with the
//   covariance of the globally first scoring pair; results and absolute poses of the shards together are bit-identical to one process.
#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "orbfront_shard.hpp"

// ---------------------------------------------------------------- recording stand-in of the device side -----------------------
struct orbf_context {
    std::vector<int> slotFrame;          // global frame id held by each frame slot
    std::vector<int> pairFrame;          // global pair id (= id of its first frame) of each pair slot
    std::vector<orbf_ransac_result> res;
    int probes = 0;
};
static int g_deadUntil = 0;              // global pairs < g_deadUntil never reach scoring (too few matches)
static bool scores(int gp) { return gp >= g_deadUntil; }
static double cov_of(int gp) { return 100.0 + gp; }

static void make_T(uint32_t seed, double cov, float* T)
{   // a rigid motion that depends on everything a pair's result depends on; products of these do not commute
    const float a = 0.01f * (float)(seed % 97) + 0.001f * (float)cov, c = std::cos(a), s = std::sin(a);
    const float t[16] = { c, -s, 0, 0.1f * (float)(seed % 7), s, c, 0, 0.01f * (float)(seed % 13), 0, 0, 1, 0.5f, 0, 0, 0, 1 };
    std::memcpy(T, t, sizeof t);
}

extern "C" int orbf_extract_batch(orbf_context* c, int32_t slot0, int32_t n, const uint8_t* gray, int64_t, int64_t frame_stride, const uint16_t*,
    int64_t, int64_t)
{
    if ((int)c->slotFrame.size() < slot0 + n) c->slotFrame.resize(slot0 + n, -1);
    for (int i = 0; i < n; ++i) c->slotFrame[slot0 + i] = gray[(size_t)i * frame_stride];      // first byte of a frame = its global id
    return ORBF_OK;
}
extern "C" int orbf_match_pairs(orbf_context* c, const int32_t* pairs, int32_t np, float, int32_t)
{
    c->pairFrame.assign(np, -1);
    for (int k = 0; k < np; ++k) {
        const int q = c->slotFrame.at(pairs[2 * k]), t = c->slotFrame.at(pairs[2 * k + 1]);
        if (t != q + 1) return ORBF_ERR_ARG;                                                  // consecutive frames only
        c->pairFrame[k] = q;
    }
    return ORBF_OK;
}
extern "C" int orbf_ransac_probe_depth_cov(orbf_context* c, int32_t np, const orbf_ransac_config*, double* cov)
{
    if (np > (int)c->pairFrame.size()) return ORBF_ERR_STATE;
    ++c->probes;
    *cov = -1.0;
    for (int k = 0; k < np; ++k) if (scores(c->pairFrame[k])) { *cov = cov_of(c->pairFrame[k]); break; }
    return ORBF_OK;
}
extern "C" int orbf_ransac_pairs(orbf_context* c, int32_t np, const orbf_ransac_config* cfg)
{
    if (np > (int)c->pairFrame.size()) return ORBF_ERR_STATE;
    double cov = cfg->depth_cov;
    c->res.assign(np, orbf_ransac_result());
    for (int k = 0; k < np; ++k) {
        const int gp = c->pairFrame[k];
        orbf_ransac_result& r = c->res[k];
        std::memset(&r, 0, sizeof r);
        if (scores(gp) && cov < 0.0) cov = cov_of(gp);                                        // the latch of a call without a covariance
        r.ok = scores(gp);
        r.real_iters = (int32_t)(cfg->seed + (uint32_t)k);                                    // the seed the pair drew from
        r.n_good = gp;
        r.depth_cov_used = scores(gp) ? cov : -1.0;
        if (r.ok) make_T(cfg->seed + (uint32_t)k, cov, r.T12);
        else { r.used_identity = 1; for (int i = 0; i < 16; ++i) r.T12[i] = i % 5 == 0 ? 1.f : 0.f; }
    }
    return ORBF_OK;
}
extern "C" int orbf_download_ransac_summary(orbf_context* c, int32_t np, orbf_ransac_result* out)
{
    if (np > (int)c->res.size()) return ORBF_ERR_ARG;
    std::memcpy(out, c->res.data(), np * sizeof *out);
    return ORBF_OK;
}
extern "C" int orbf_compose_trajectory(orbf_context* c, int32_t np, const float* pose0, float* poses, uint8_t*)
{
    if (np > (int)c->res.size()) return ORBF_ERR_ARG;
    for (int i = 0; i < 16; ++i) poses[i] = pose0 ? pose0[i] : (i % 5 == 0 ? 1.f : 0.f);
    for (int k = 0; k < np; ++k) {
        const float* A = c->res[k].T12; const float* B = poses + 16 * k; float* C = poses + 16 * (k + 1);
        for (int r = 0; r < 4; ++r) for (int cc = 0; cc < 4; ++cc) {
            float t = A[4 * r] * B[cc];
            t = t + A[4 * r + 1] * B[4 + cc]; t = t + A[4 * r + 2] * B[8 + cc]; t = t + A[4 * r + 3] * B[12 + cc];
            C[4 * r + cc] = t;
        }
    }
    return ORBF_OK;
}

// ---------------------------------------------------------------- the ranks of a job as threads ------------------------------------
struct Job {
    int world;
    std::mutex m; std::condition_variable cv;
    std::vector<double> gathered; int arrived = 0, generation = 0;
    std::vector<std::vector<float>> mailbox; std::vector<int> full;
    explicit Job(int w) : world(w), gathered(w, -1.0), mailbox(w, std::vector<float>(16)), full(w, 0) {}
    orbf::Exchange exchange(int rank)
    {
        orbf::Exchange x;
        x.allgather_f64 = [this, rank](double local, double* all) {
            std::unique_lock<std::mutex> l(m);
            gathered[rank] = local;
            const int gen = generation;
            if (++arrived == world) { arrived = 0; ++generation; cv.notify_all(); }
            else cv.wait(l, [&] { return generation != gen; });
            for (int r = 0; r < world; ++r) all[r] = gathered[r];
        };
        x.send_pose = [this](const float* p, int32_t dst) {
            std::unique_lock<std::mutex> l(m);
            std::memcpy(mailbox[dst].data(), p, 64); full[dst] = 1; cv.notify_all();
        };
        x.recv_pose = [this, rank](float* p, int32_t src) {
            (void)src;
            std::unique_lock<std::mutex> l(m);
            cv.wait(l, [&] { return full[rank] != 0; });
            std::memcpy(p, mailbox[rank].data(), 64); full[rank] = 0;
        };
        return x;
    }
};

struct RankOut { orbf::FrameShard sh; std::vector<orbf_ransac_result> res; std::vector<float> poses; double cov; std::vector<int> extracted; int probes; };

static RankOut run_rank(Job& job, int n, int rank, uint32_t seed, const float* pose0, double explicitCov)
{
    orbf_context ctx;
    orbf::SequenceShard S(&ctx, n, rank, job.world);
    const orbf::FrameShard sh = S.shard();
    const int64_t fs = 64;                                       // a "frame" of 64 bytes whose first byte is its global id
    std::vector<uint8_t> gray((size_t)std::max(sh.frames(), 1) * fs, 0);
    for (int i = 0; i < sh.frames(); ++i) gray[(size_t)i * fs] = (uint8_t)(sh.first + i);
    orbf_ransac_config cfg; std::memset(&cfg, 0, sizeof cfg);
    cfg.iterations = 200; cfg.min_inlier_th = 20; cfg.max_mahal = 3.f; cfg.sample_size = 4; cfg.check_depth = 1; cfg.depth_cov = explicitCov; cfg.seed = seed;
    const orbf::Exchange x = job.exchange(rank);
    S.run(gray.data(), 8, fs, nullptr, 0, 0, 0.8f, true, cfg, x);
    RankOut o;
    o.sh = sh; o.res = S.results(); o.poses = S.compose(pose0, x); o.cov = S.depth_cov(); o.extracted = ctx.slotFrame; o.probes = ctx.probes;
    return o;
}

static std::vector<RankOut> run_job(int n, int world, uint32_t seed, const float* pose0, double explicitCov = -1.0)
{
    Job job(world);
    std::vector<RankOut> out(world);
    std::vector<std::thread> th;
    for (int r = 0; r < world; ++r) th.emplace_back([&, r] { out[r] = run_rank(job, n, r, seed, pose0, explicitCov); });
    for (auto& t : th) t.join();
    return out;
}

#define CHECK(cond, ...) do { if (!(cond)) { fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); return 1; } } while (0)

int main()
{
    float pose0[16]; make_T(5, 3.0, pose0);
    long cases = 0;
    const int ns[] = { 0, 1, 2, 3, 5, 8, 13, 40, 41 };
    for (int n : ns) for (int dead : { 0, 1, 3, 7, 1000 }) for (int useExplicit = 0; useExplicit < 2; ++useExplicit) {
        g_deadUntil = dead;
        const double explicitCov = useExplicit ? 0.25 : -1.0;
        const uint32_t seed = 42;
        const std::vector<RankOut> one = run_job(n, 1, seed, pose0, explicitCov);
        const int np = std::max(n - 1, 0);
        CHECK((int)one[0].res.size() == np && (int)one[0].poses.size() == 16 * n, "single process: n %d", n);
        const double wantCov = useExplicit ? 0.25 : (dead < np ? cov_of(dead) : -1.0);
        CHECK(one[0].cov == wantCov, "single process covariance %g, want %g (n %d dead %d)", one[0].cov, wantCov, n, dead);
        for (int p = 0; p < np; ++p) CHECK(one[0].res[p].real_iters == (int)seed + p && one[0].res[p].n_good == p, "single process pair %d", p);
        for (int world : { 2, 3, 4, 8 }) {
            const std::vector<RankOut> job = run_job(n, world, seed, pose0, explicitCov);
            std::vector<orbf_ransac_result> res; std::vector<float> poses; std::vector<int> owner(n, 0);
            for (int r = 0; r < world; ++r) {
                const RankOut& o = job[r];
                CHECK(o.cov == wantCov, "rank %d of %d scores with %g, want %g (n %d dead %d)", r, world, o.cov, wantCov, n, dead);
                CHECK((int)o.extracted.size() == o.sh.frames(), "rank %d extracted %zu frames, shard has %d", r, o.extracted.size(), o.sh.frames());
                for (int i = 0; i < o.sh.frames(); ++i) CHECK(o.extracted[i] == o.sh.first + i, "rank %d slot %d holds frame %d", r, i, o.extracted[i]);
                for (int f = o.sh.start; f < o.sh.stop; ++f) ++owner[f];
                CHECK(o.probes == ((useExplicit || o.sh.pairs() == 0) ? 0 : 1), "rank %d probed %d times", r, o.probes);
                CHECK((int)o.res.size() == o.sh.pairs(), "rank %d results", r);
                for (size_t k = 0; k < o.res.size(); ++k) CHECK(o.res[k].n_good == o.sh.pair0 + (int)k, "rank %d result %zu is pair %d", r, k, o.res[k].n_good);
                CHECK((int)o.poses.size() == 16 * (o.sh.stop - o.sh.start), "rank %d returns %zu floats of poses for %d frames", r, o.poses.size(), o.sh.stop - o.sh.start);
                res.insert(res.end(), o.res.begin(), o.res.end());
                poses.insert(poses.end(), o.poses.begin(), o.poses.end());
            }
            for (int f = 0; f < n; ++f) CHECK(owner[f] == 1, "frame %d has %d owners (world %d)", f, owner[f], world);
            CHECK(res.size() == one[0].res.size() && (res.empty() || !std::memcmp(res.data(), one[0].res.data(), res.size() * sizeof res[0])),
                "results of %d ranks differ from one process (n %d dead %d)", world, n, dead);
            CHECK(poses.size() == one[0].poses.size() && (poses.empty() || !std::memcmp(poses.data(), one[0].poses.data(), poses.size() * 4)),
                "poses of %d ranks differ from one process (n %d dead %d)", world, n, dead);
            ++cases;
        }
    }
    // error behaviour: calls before run(), a missing exchange
    {
        orbf_context ctx; orbf::SequenceShard S(&ctx, 10, 1, 2);
        bool threw = false;
        try { S.results(); } catch (const orbf::ShardError& e) { threw = e.status == ORBF_ERR_STATE; }
        CHECK(threw, "results() before run() must throw ORBF_ERR_STATE");
        threw = false;
        std::vector<uint8_t> gray(64 * 6, 0);
        for (int i = 0; i < 6; ++i) gray[64 * i] = (uint8_t)(4 + i);
        orbf_ransac_config cfg; std::memset(&cfg, 0, sizeof cfg); cfg.depth_cov = -1.0;
        try { S.run(gray.data(), 8, 64, nullptr, 0, 0, 0.8f, true, cfg, orbf::Exchange()); } catch (const orbf::ShardError& e) { threw = e.status == ORBF_ERR_ARG; }
        CHECK(threw, "run() on 2 ranks without an all-gather must throw ORBF_ERR_ARG");
        threw = false;
        try { orbf::frame_shard(10, 2, 2); } catch (const orbf::ShardError& e) { threw = e.status == ORBF_ERR_ARG; }
        CHECK(threw, "frame_shard with rank == world must throw");
    }
    printf("shard host logic: %ld sharded jobs identical to one process\n", cases);
    return 0;
}
