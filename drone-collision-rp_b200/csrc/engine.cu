// engine.cu -- the extern "C" launcher ABI of include/dcrp.h.
//
// Host logic only: argument checks, the HBM layout of a scenario, launches on
// CUDA streams, result gathering.  There is deliberately NO CPU implementation
// of the trial path in this library: without a usable sm_100 device
// dcrp_engine_create fails and nothing else can be called.
#include "../../include/dcrp.h"
#include "dcrp_device.cuh"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <climits>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>      // header-only NVTX v3: ranges show up in Nsight Systems / ncu --nvtx, cost nothing otherwise

#include "kernels.cu"
#include "screen3.cu"
#include "comm.cu"

using namespace dcrp;

static_assert(sizeof(dcrp_draw) == sizeof(Draw) && sizeof(Draw) == 32, "draw layout");
static_assert(sizeof(dcrp_record) == sizeof(Record) && sizeof(Record) == 64, "record layout");

// ------------------------------------------------------------------ errors
static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(DCRP_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),    \
                        __FILE__, __LINE__);                                                       \
    } while (0)

// ------------------------------------------------------------------ tracing (NVTX ranges in the launcher)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

#ifdef DCRP_TIMELINE
// instrumented build only: host-side time stamps (ns) of the last dcrp_simulate call, tools/e2e_breakdown.py
static double g_host_ns[16];
static std::chrono::steady_clock::time_point g_host_t0;
#define HOST_MARK(i) (g_host_ns[i] = std::chrono::duration<double, std::nano>(std::chrono::steady_clock::now() - g_host_t0).count())
extern "C" int dcrp_debug_host_times(double* out, int n)
{
    for (int i = 0; i < n && i < 16; ++i) out[i] = g_host_ns[i];
    return DCRP_OK;
}
#else
#define HOST_MARK(i) do {} while (0)
#endif

// ------------------------------------------------------------------ buffers
struct DevBuf {            // grow-only device buffer
    void*  p = nullptr;
    size_t cap = 0;
    cudaError_t need(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = (bytes + 255) & ~size_t(255);
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct PinnedBuf {         // grow-only pinned host buffer (async copies need page-locked memory)
    void*  p = nullptr;
    size_t cap = 0;
    cudaError_t need(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = (bytes + 4095) & ~size_t(4095);
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocMapped);      // kernels may write it (published runs)
        if (e == cudaSuccess) { cap = want; std::memset(p, 0, want); }     // (the flag word of a published run must not start on garbage)
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

static inline size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

// Screen kernels: k_screen3 (screen3.cu; 1-D slab level + queued 2-D re-scan, 256-trial warp
// items) for plain MIXED runs; k_screen2<R=2, BLOCK=384, NP=2> (kernels.cu; sorted 2-D scan, 3072-trial CTA
// items) for replay tables, reach culling, the sub-step extension, fp32 decisions and long tracks.
constexpr uint32_t DEFAULT_RECORD_CAP = 1u << 16;
constexpr uint32_t CAND_CAP = 1u << 22;
constexpr uint64_t MAX_ITEMS_PER_LAUNCH = 1ull << 30;

struct dcrp_engine {
    int device = 0;
    cudaDeviceProp prop{};
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_begin = nullptr, ev_k0 = nullptr, ev_trials = nullptr, ev_end = nullptr, ev_a = nullptr, ev_b = nullptr;
    uint64_t launches = 0;

    // scenario
    bool has_scenario = false;
    ScenarioDev sd{};
    std::vector<TrackDev> h_tracks;
    dcrp_model model{};
    float ms_prepare = 0.f;
    uint64_t h2d_scenario = 0;
    // what the resident scenario was built from (dcrp_simulate skips the upload when the caller passes it again)
    std::vector<dcrp_drone> h_drones;
    size_t scn_o_ax = 0, scn_o_ay = 0, scn_o_az = 0;
    // inputs live in ONE device arena filled by ONE H2D copy from a pinned mirror:
    //   [tracks | ax | ay | az | trk32 | drones], each segment 256-B aligned
    DevBuf d_scn;
    PinnedBuf h_scn;
    DevBuf d_p1x, d_p1y, d_s1fh, d_s1pm, d_kink, d_p1f, d_s1sub_pm, d_s1sub_fs, d_s1sub_ft;
    DevBuf d_reach, d_alive, d_reach_tot;
    DevBuf d_slab_dir;             // k_screen3: slab direction per pair
    bool slab_ready = false;       // k_slab_axis has run for the current scenario
    double sub_survivors = -1.0;   // slab survivors per trial of the last sub-step run through k_screen3 on this scenario (-1: none yet)
    uint32_t runs_on_scenario = 0; // dcrp_run_async calls since the last upload
    size_t s3_smem[4] = {0, 0, 0, 0};    // [0] k_screen3<2>, [1] k_screen3<4>, [2], [3] their sub-step builds
    int s3_occ[4] = {0, 0, 0, 0};
    bool prepare_timed = false;

    // run
    bool has_run = false;
    dcrp_run run{};
    cudaStream_t run_stream = nullptr, run_stream_prev = nullptr;
    bool external_counts = false;
    uint32_t rec_cap = 0;
    uint32_t launches_this_run = 0;
    uint64_t h2d_run = 0;
    // per-run outputs live in ONE device arena read back by ONE D2H copy into a pinned mirror:
    //   [rec_count, cand_count, work_counter, pad | stats[ST_COUNT] | counts[n_pairs] | first_conflict[n_pairs]]
    DevBuf d_out;
    PinnedBuf h_out;
    size_t out_counts_off = 0, out_first_off = 0, out_bytes = 0;     // out_bytes: the part before the records
    // Published runs: a plain MIXED run of one k_screen3 launch with engine-owned counts is read back by a kernel of
    // its own, k_publish, queued right behind the trial kernel: one CTA stores the arena and the first records into
    // h_out through its device mapping and raises a flag; dcrp_fetch polls the flag instead of queueing a D2H copy
    // behind the run and waiting for the stream (measured from the host: ~13 us from the end of the trial kernel to
    // knowing it, against ~3).  Without a communicator (whose all-reduce may still want the counts on the device)
    // k_publish also leaves the arena cleared, and the next such run needs no k_run_init in front of it: the trial
    // kernel is the first thing queued.  e2e per 1.0e7-trial call: 0.174 -> 0.159 ms.
    bool published = false;        // the last run publishes
    bool run_timed = true;         // the last run recorded its timing events (no DCRP_FLAG_NO_TIMING)
    uint32_t pub_seq = 0;          // flag value of the last published run
    bool arena_clean = false;      // the last run leaves (d_out, out_bytes) cleared
    void* clean_ptr = nullptr;
    size_t clean_bytes = 0;
    uint32_t run_status = 0;       // DCRP_STAT_* bits of the current run (reported in dcrp_stats.status)
    Comm* comm = nullptr;          // dcrp_engine_attach_comm
    bool run_event_valid = false;  // ev_end marks the end of the last run on its stream: later calls on OTHER streams wait for it
    bool reuse_device_replay = false;   // dcrp_fetch's fp64 fallback: d_replay already holds the run's table
    DevBuf d_cands, d_replay;      // (records live in the tail of d_out)
    DevBuf d_trial_hit, d_trial_first, d_trial_min;
    // launch configuration caches (per engine = per device: function attributes are per device)
    size_t screen_smem[4] = {0, 0, 0, 0};
    int screen_occ[4] = {0, 0, 0, 0};
    int ap_occ = 0;
    bool allpairs_run = false;
    int reserve_sms = 0;           // SMs kept free for the caller's communication kernels
    bool run_cull = false;         // the last run used DCRP_FLAG_REACH_CULL
    bool run_screen3 = false;      // the last run went through k_screen3
    bool has_confirm = false;      // the last run launched a separate confirm / resolve kernel (ev_trials recorded)
    bool k0_is_begin = false;      // ... launched programmatically behind k_run_init: no event between the two
    bool reach_ready = false;      // k_reach has run for the current scenario
    uint64_t ap_rays = 0;
    DevBuf d_ap_hist, d_ap_rank, d_ap_rayA, d_ap_rayB, d_ap_meta, d_ap_l2;
    DevBuf d_scratch;    // draws / trajectories
};

// Records kept in the first D2H copy of a fetch (one copy, one sync for the usual handful of hits).
constexpr uint32_t INLINE_RECORDS = 32;
// Largest output arena k_publish delivers (dcrp_engine::published); larger ones go through the copy engine.
constexpr size_t PUBLISH_MAX_BYTES = 64 * 1024;

// Sizes the per-run output arena: [rec_count, cand_count, work_counter, pad | stats | counts | first_conflict]
// (cleared by k_run_init) followed by rec_cap records.
static cudaError_t size_out_arena(dcrp_engine* e, size_t n_pairs)
{
    e->out_counts_off = align256(16 + sizeof(unsigned long long) * ST_COUNT);
    e->out_first_off = align256(e->out_counts_off + sizeof(uint64_t) * n_pairs);
    e->out_bytes = align256(e->out_first_off + sizeof(int32_t) * n_pairs);
    const size_t cap_before = e->d_out.cap;
    cudaError_t ce = e->d_out.need(e->out_bytes + sizeof(Record) * (size_t)e->rec_cap);
    if (e->d_out.cap != cap_before) e->arena_clean = false;     // a new allocation (even at the old address) is not a cleared arena
    if (ce != cudaSuccess) return ce;
    // the mirror: the arena, INLINE_RECORDS records, and the flag line of a published run
    return e->h_out.need(e->out_bytes + sizeof(Record) * (size_t)INLINE_RECORDS + 64);
}
static size_t pub_flag_off(const dcrp_engine* e) { return e->out_bytes + sizeof(Record) * (size_t)INLINE_RECORDS; }
static Record* out_records(dcrp_engine* e) { return reinterpret_cast<Record*>(e->d_out.as<unsigned char>() + e->out_bytes); }

static uint32_t* small_rec_count(dcrp_engine* e) { return e->d_out.as<uint32_t>(); }
static uint32_t* small_cand_count(dcrp_engine* e) { return e->d_out.as<uint32_t>() + 1; }
static unsigned long long* small_stats(dcrp_engine* e)
{
    return reinterpret_cast<unsigned long long*>(e->d_out.as<unsigned char>() + 16);
}

// ------------------------------------------------------------------ lifetime
extern "C" int dcrp_abi_version(void) { return DCRP_ABI_VERSION; }
extern "C" const char* dcrp_last_error(void) { return g_err.c_str(); }

extern "C" void dcrp_default_model(dcrp_model* m)
{
    if (!m) return;
    m->aircraft_radius = 3.5;   // Monte_Carlo.cpp:13
    m->drone_radius = 0.19;     // Monte_Carlo.cpp:18
    m->start_alt = 100.0;       // Drone.cpp:54
    m->v_straight = 19.0;       // Drone.cpp:55
    m->v_ascend = 8.0;          // Drone.cpp:56
}

extern "C" int dcrp_engine_create(int device, dcrp_engine** out)
{
    if (!out) return fail(DCRP_ERR_INVALID, "dcrp_engine_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t ce = cudaGetDeviceCount(&n);
    if (ce != cudaSuccess || n <= 0)
        return fail(DCRP_ERR_NO_DEVICE, "no CUDA device (%s); this engine has no CPU fallback",
                    ce == cudaSuccess ? "device count is 0" : cudaGetErrorString(ce));
    if (device < 0 || device >= n) return fail(DCRP_ERR_INVALID, "device %d out of range [0,%d)", device, n);
    CU(cudaSetDevice(device));
    dcrp_engine* e = new (std::nothrow) dcrp_engine();
    if (!e) return fail(DCRP_ERR_INVALID, "out of host memory");
    e->device = device;
    cudaError_t pe = cudaGetDeviceProperties(&e->prop, device);
    if (pe != cudaSuccess) { delete e; return fail(DCRP_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(pe)); }
    if (e->prop.major != 10) {
        int maj = e->prop.major, mn = e->prop.minor;
        delete e;
        return fail(DCRP_ERR_NO_DEVICE, "device %d is sm_%d%d; libdcrp is built for sm_100a only", device, maj, mn);
    }
    if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&e->ev_begin) != cudaSuccess || cudaEventCreate(&e->ev_k0) != cudaSuccess ||
        cudaEventCreate(&e->ev_trials) != cudaSuccess ||
        cudaEventCreate(&e->ev_end) != cudaSuccess || cudaEventCreate(&e->ev_a) != cudaSuccess ||
        cudaEventCreate(&e->ev_b) != cudaSuccess) {
        delete e;
        return fail(DCRP_ERR_CUDA, "stream/event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    *out = e;
    return DCRP_OK;
}

extern "C" int dcrp_engine_detach_comm(dcrp_engine* e);

extern "C" int dcrp_engine_destroy(dcrp_engine* e)
{
    if (!e) return DCRP_OK;
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    dcrp_engine_detach_comm(e);
    DevBuf* bufs[] = {&e->d_scn, &e->d_p1x, &e->d_p1y, &e->d_s1fh, &e->d_s1pm, &e->d_kink, &e->d_p1f, &e->d_s1sub_pm, &e->d_s1sub_fs,
                      &e->d_s1sub_ft, &e->d_out,
                      &e->d_cands, &e->d_replay, &e->d_trial_hit, &e->d_trial_first,
                      &e->d_trial_min, &e->d_scratch, &e->d_ap_hist, &e->d_ap_rank, &e->d_ap_rayA,
                      &e->d_ap_rayB, &e->d_ap_meta, &e->d_ap_l2, &e->d_reach, &e->d_alive, &e->d_reach_tot,
                      &e->d_slab_dir};
    for (DevBuf* b : bufs) b->release();
    e->h_scn.release();
    e->h_out.release();
    cudaEventDestroy(e->ev_begin); cudaEventDestroy(e->ev_k0); cudaEventDestroy(e->ev_trials); cudaEventDestroy(e->ev_end);
    cudaEventDestroy(e->ev_a); cudaEventDestroy(e->ev_b);
    cudaStreamDestroy(e->stream);
    delete e;
    return DCRP_OK;
}

// ------------------------------------------------------------------ host helper
extern "C" int dcrp_target_box(double ay0, double az0, double box[6])
{
    if (!box) return fail(DCRP_ERR_INVALID, "dcrp_target_box: box is NULL");
    const bool arrival = (az0 != 0);                       // Drone.cpp:47-52
    double min_lat, max_lat, min_long, max_long;
    if (ay0 <= 5706340.62 && ay0 >= 5705792.58) {          // 27R  Drone.cpp:135-137
        min_lat = 5705770.46; max_lat = min_lat + 500;
    } else if (ay0 <= 5705065.74 && ay0 >= 5704388.38) {   // 27L  Drone.cpp:148-150
        max_lat = 5704800.46; min_lat = max_lat - 500;
    } else {
        return fail(DCRP_ERR_RUNWAY, "first latitude %.3f is in neither the 27R nor the 27L band", ay0);
    }
    if (arrival) { min_long = 676345.60; max_long = min_long + 5000; }   // Drone.cpp:138-141
    else         { max_long = 676819.02; min_long = max_long - 5000; }   // Drone.cpp:142-145
    box[0] = min_long; box[1] = max_long; box[2] = min_lat; box[3] = max_lat;
    box[4] = 100; box[5] = 100 + 400;                                    // Drone.cpp:131-132
    return DCRP_OK;
}

// ------------------------------------------------------------------ scenario
static float ulp32_of(double v)
{
    float f = (float)std::fabs(v);
    return std::nextafterf(f, INFINITY) - f;
}

extern "C" int dcrp_scenario_upload(dcrp_engine* e, const dcrp_track* tracks, int32_t n_tracks,
                                    const dcrp_drone* drones, int32_t n_drones, const dcrp_model* model)
{
    NvtxRange nvtx_("dcrp_scenario_upload");
    if (!e) return fail(DCRP_ERR_INVALID, "engine is NULL");
    if (!tracks || n_tracks <= 0 || !drones || n_drones <= 0 || !model)
        return fail(DCRP_ERR_INVALID, "dcrp_scenario_upload: tracks/drones/model missing or empty");
    if ((uint64_t)n_tracks * (uint64_t)n_drones > (1ull << 31) - 1)
        return fail(DCRP_ERR_INVALID, "too many (track, drone) pairs");
    if (!(model->v_straight > 0) || !std::isfinite(model->v_straight) ||
        !(model->aircraft_radius >= 0) || !(model->drone_radius >= 0) ||
        !(model->aircraft_radius + model->drone_radius > 0) || !std::isfinite(model->aircraft_radius + model->drone_radius))
        return fail(DCRP_ERR_INVALID, "model: v_straight must be positive and finite, the radii non-negative with a positive finite sum");
    // pitch lies in [0, pi/2] (Drone.cpp:218), so the speed factor lies between v_straight and v_ascend: a negative
    // v_ascend would make the drone fly backwards and DESCEND, which the reach bound (k_reach) excludes by construction
    if (!(model->v_ascend >= 0) || !std::isfinite(model->v_ascend) || !std::isfinite(model->start_alt))
        return fail(DCRP_ERR_INVALID, "model: v_ascend must be non-negative and finite, start_alt finite");
    CU(cudaSetDevice(e->device));
    (void)cudaGetLastError();      // a stale error left in this thread by other CUDA users (torch, NCCL) is not ours to report
    e->has_scenario = false;
    e->has_run = false;

    // ---- validate, offsets, extents
    e->h_tracks.assign((size_t)n_tracks, TrackDev{});
    uint64_t total = 0, total_xy = 0;
    int t_max = 0, vl_max = 0;
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    auto grow = [&](double x, double y, double z) {
        lo[0] = std::min(lo[0], x); hi[0] = std::max(hi[0], x);
        lo[1] = std::min(lo[1], y); hi[1] = std::max(hi[1], y);
        lo[2] = std::min(lo[2], z); hi[2] = std::max(hi[2], z);
    };
    for (int a = 0; a < n_tracks; ++a) {
        const dcrp_track& t = tracks[a];
        if (!t.x || !t.y || !t.z || t.n_steps <= 0)
            return fail(DCRP_ERR_INVALID, "track %d: empty or NULL arrays", a);
        if (t.takeoff_t < 1 || t.takeoff_t > t.n_steps)
            return fail(DCRP_ERR_INVALID,
                        "track %d: takeoff_t=%d must be in [1, n_steps=%d] (the reference draws "
                        "uniform_int(1, TakeoffTime), Drone.cpp:108-114)", a, t.takeoff_t, t.n_steps);
        if (t.takeoff_t >= (1 << 20)) return fail(DCRP_ERR_INVALID, "track %d: takeoff_t too large", a);
        for (int k = 0; k < 3; ++k)
            if (!(t.box[2 * k] <= t.box[2 * k + 1]))
                return fail(DCRP_ERR_INVALID, "track %d: target box is inverted or NaN", a);
        TrackDev& d = e->h_tracks[(size_t)a];
        d.n_steps = t.n_steps; d.takeoff_t = t.takeoff_t; d.off = (uint32_t)total; d.off_xy = (uint32_t)total_xy;
        total_xy += (uint64_t)((t.n_steps + 1) & ~1);
        std::memcpy(d.box, t.box, sizeof d.box);
        d.wx32 = (t.box[1] - t.box[0]) * (1.0 / 4294967296.0);
        d.wy32 = (t.box[3] - t.box[2]) * (1.0 / 4294967296.0);
        d.wz32 = (float)((t.box[5] - t.box[4]) * (1.0 / 4294967296.0));
        d.z0rel = (float)(t.box[4] - model->start_alt);
        {   // default slab direction (screen3.cu): where the aircraft goes after take-off
            const int i0 = std::min(t.takeoff_t, t.n_steps) - 1;
            double ux = t.x[t.n_steps - 1] - t.x[i0], uy = t.y[t.n_steps - 1] - t.y[i0];
            const double um = std::hypot(ux, uy);
            if (!(um > 1e-6) || !std::isfinite(um)) { ux = 1.0; uy = 0.0; } else { ux /= um; uy /= um; }
            d.dir0x = (float)ux; d.dir0y = (float)uy;
        }
        total += (uint64_t)t.n_steps;
        if (total > (1ull << 31)) return fail(DCRP_ERR_INVALID, "tracks too long in total");
        t_max = std::max(t_max, t.takeoff_t);
        vl_max = std::max(vl_max, t.n_steps);
        for (int i = 0; i < t.n_steps; ++i) {
            if (!std::isfinite(t.x[i]) || !std::isfinite(t.y[i]) || !std::isfinite(t.z[i]))
                return fail(DCRP_ERR_INVALID, "track %d: non-finite sample at index %d", a, i);
            grow(t.x[i], t.y[i], t.z[i]);
        }
        grow(t.box[0], t.box[2], t.box[4]);
        grow(t.box[1], t.box[3], t.box[5]);
    }
    for (int j = 0; j < n_drones; ++j) {
        const dcrp_drone& d = drones[j];
        if (!std::isfinite(d.x0) || !std::isfinite(d.y0) || !std::isfinite(d.heading0))
            return fail(DCRP_ERR_INVALID, "drone %d: non-finite start row", j);
        const double reach = model->v_straight * (double)t_max;
        grow(d.x0 - reach, d.y0 - reach, model->start_alt);
        grow(d.x0 + reach, d.y0 + reach, model->start_alt);
    }

    ScenarioDev s{};
    s.n_tracks = n_tracks; s.n_drones = n_drones; s.t_max = t_max; s.n_steps_max = vl_max;
    s.n_sm = e->prop.multiProcessorCount;
    s.div_drones = make_fastdiv((uint32_t)n_drones);
    s.ox = std::floor(0.5 * (lo[0] + hi[0]));
    s.oy = std::floor(0.5 * (lo[1] + hi[1]));
    s.oz = std::floor(0.5 * (lo[2] + hi[2]));
    // + the distance flown inside one scan tile: the incremental scan's rounding is bounded by the
    // ulp of the largest coordinate met since the tile's exact re-base
    const double extent = std::max({hi[0] - s.ox, s.ox - lo[0], hi[1] - s.oy, s.oy - lo[1],
                                    hi[2] - s.oz, s.oz - lo[2]}) + 1.0 +
                          std::max(std::fabs(model->v_straight), std::fabs(model->v_ascend)) * (vl_max <= 128 ? 128.0 : 512.0);
    s.start_alt = model->start_alt; s.v_straight = model->v_straight; s.v_ascend = model->v_ascend;
    s.radius = model->drone_radius + model->aircraft_radius;          // Drone.cpp:260
    {   // sqrt_rn(d2) < R  <=>  d2 < c, c = smallest double whose rounded root reaches R
        double c = s.radius * s.radius;
        while (std::sqrt(c) >= s.radius) c = std::nextafter(c, 0.0);
        while (std::sqrt(c) < s.radius) c = std::nextafter(c, INFINITY);
        s.thr_d2 = c;
    }
    s.radius2 = s.radius * s.radius;
    // fp32 screen error budget per coordinate (DESIGN.md "screen margin"): the scan accumulates
    // P += S for at most tile_steps indices between exact re-bases (<= tile_steps/2 ulp), plus one
    // rounding each for the re-based start, the recentred track sample and the difference, plus the
    // set-up (target-kink rounded once from fp64: relative 2^-24 of the flown distance).  The margin
    // is that bound times sqrt(3) for the 3-D distance, plus a 0.02 m floor.
    const int screen_tile = vl_max <= 128 ? 128 : 512;
    const double ulp = (double)ulp32_of(extent);
    // after the kink the speed factor lies between v_straight and v_ascend (Drone.cpp:220, pitch in [0, pi/2])
    const double vmax = std::max(std::fabs(model->v_straight), std::fabs(model->v_ascend));
    const double margin = 0.02 + 1.7321 * (0.5 * screen_tile + 4.0) * ulp +
                          (double)vl_max * vmax * std::ldexp(1.0, -21);
    s.thr2_screen = std::nextafterf((float)((s.radius + margin) * (s.radius + margin)), INFINITY);
    // 1-D slab level (screen3.cu): the same error sources, on coordinates projected on a unit vector
    // (magnitudes up to sqrt(2) x extent, three more roundings for the projections), re-based every 128
    const double ulp_p = (double)ulp32_of(1.4143 * extent);
    const double margin_slab = 0.02 + (0.5 * 128.0 + 12.0) * ulp_p + (double)vl_max * vmax * std::ldexp(1.0, -20);
    s.thr_slab = std::nextafterf((float)(s.radius + margin_slab), INFINITY);
    for (int a = 0; a < n_tracks; ++a) {          // sub-step mode: per-track level-1 radius
        const dcrp_track& t = tracks[a];
        double amax = 0.0, amax3 = 0.0;
        for (int i = 0; i + 1 < t.n_steps; ++i) {
            const double h = std::hypot(t.x[i + 1] - t.x[i], t.y[i + 1] - t.y[i]);
            amax = std::max(amax, h);
            amax3 = std::max(amax3, std::hypot(h, t.z[i + 1] - t.z[i]));
        }
        // level 1 tests SAMPLES: a segment whose interpolated horizontal separation drops below R has an end
        // within R + half the horizontal relative step (<= aircraft step + vmax)
        const double rr = s.radius + margin + 0.5 * (amax + vmax) * (1.0 + 1e-6);
        TrackDev& d = e->h_tracks[(size_t)a];
        d.thr2_l1_sub = std::nextafterf((float)(rr * rr), INFINITY);
        // the slab level likewise: |n . D| at one END of a segment in conflict is below R + half the projected
        // relative step, and a projection is no longer than the step
        d.thr_slab_sub = std::nextafterf((float)(s.radius + margin_slab + 0.5 * (amax + vmax) * (1.0 + 1e-6)), INFINITY);
        // level 2 evaluates c + t(2b + ta) in fp32 with |c|, |b|, |a| up to (R + E)^2, E the 3-D relative step
        // (<= aircraft step + sqrt(2) vmax): ~10 roundings of 2^-24 each, bounded by 4e-6 (R + E)^2.  Nothing for
        // real tracks (E ~ 100 m: 0.05 m^2), decisive for sparse or teleporting ones.
        const double E = amax3 + 1.4143 * vmax;
        d.thr2_l2_sub = std::nextafterf((float)((s.radius + margin) * (s.radius + margin) +
                                                4e-6 * (s.radius + E) * (s.radius + E)), INFINITY);
        // |D(t)| >= min(|D_0|, |D_1|) - |E| / 2 on a segment: with both ends beyond this the closest approach is not
        // evaluated (1e-4 relative and 0.05 m for the fp32 rounding of the two squared end distances)
        const double near = (std::sqrt((double)d.thr2_l2_sub) + 0.5 * E) * (1.0 + 1e-4) + 0.05;
        d.thr2_near_sub = std::nextafterf((float)(near * near), INFINITY);
    }
    s.thr2_fp32 = (float)(s.radius * s.radius);
    s.v_straight_f = (float)model->v_straight;
    s.coef_f = (float)(2.0 * (model->v_straight - model->v_ascend) / DCRP_PI);
    s.bucket_shift = 0;
    while (((t_max - 1) >> s.bucket_shift) >= 64) ++s.bucket_shift;

    // ---- one pinned staging arena, one H2D copy
    const size_t n_pairs = (size_t)n_tracks * (size_t)n_drones;
    const size_t o_tracks = 0;
    const size_t o_ax = align256(o_tracks + sizeof(TrackDev) * (size_t)n_tracks);
    const size_t o_ay = align256(o_ax + sizeof(double) * total);
    const size_t o_az = align256(o_ay + sizeof(double) * total);
    const size_t o_t32 = align256(o_az + sizeof(double) * total);
    const size_t o_txy = align256(o_t32 + sizeof(float4) * total);
    const size_t o_dr = align256(o_txy + sizeof(float2) * total_xy);
    const size_t scn_bytes = align256(o_dr + sizeof(double4) * (size_t)n_drones);
    cudaStream_t st = e->stream;
    CU(cudaStreamSynchronize(st));          // the previous upload's copy has left the pinned mirror
    if (e->run_event_valid) CU(cudaStreamWaitEvent(st, e->ev_end, 0));   // a run on a caller's stream may still read the old scenario
    CU(e->h_scn.need(scn_bytes));
    CU(e->d_scn.need(scn_bytes));
    unsigned char* hb = e->h_scn.as<unsigned char>();
    std::memcpy(hb + o_tracks, e->h_tracks.data(), sizeof(TrackDev) * (size_t)n_tracks);
    double* hx = reinterpret_cast<double*>(hb + o_ax);
    double* hy = reinterpret_cast<double*>(hb + o_ay);
    double* hz = reinterpret_cast<double*>(hb + o_az);
    float4* h32 = reinterpret_cast<float4*>(hb + o_t32);
    float2* hxy = reinterpret_cast<float2*>(hb + o_txy);
    double4* hd = reinterpret_cast<double4*>(hb + o_dr);
    for (int a = 0; a < n_tracks; ++a) {
        const dcrp_track& t = tracks[a];
        const size_t off = e->h_tracks[(size_t)a].off;
        for (int i = 0; i < t.n_steps; ++i) {
            hx[off + i] = t.x[i]; hy[off + i] = t.y[i]; hz[off + i] = t.z[i];
            const float fx = -(float)(t.x[i] - s.ox), fy = -(float)(t.y[i] - s.oy),
                        fz = -(float)(t.z[i] - s.oz);
            h32[off + i] = make_float4(fx, fy, fz, (float)i);
            hxy[e->h_tracks[(size_t)a].off_xy + i] = make_float2(fx, fy);
        }
        if (t.n_steps & 1) hxy[e->h_tracks[(size_t)a].off_xy + t.n_steps] = make_float2(1e15f, 1e15f);   // pad
    }
    for (int j = 0; j < n_drones; ++j) {
        // Drone.cpp:121-122: sin(heading)*max_straight_speed, formed with the host libm
        // (the one the reference links) so stage 1 is bit-identical to the reference.
        const double sx = std::sin(drones[j].heading0) * model->v_straight;
        const double sy = std::cos(drones[j].heading0) * model->v_straight;
        hd[(size_t)j] = make_double4(drones[j].x0, drones[j].y0, sx, sy);
    }
    CU(e->d_p1x.need(sizeof(double) * (size_t)n_drones * (size_t)t_max));
    CU(e->d_p1y.need(sizeof(double) * (size_t)n_drones * (size_t)t_max));
    CU(e->d_s1fh.need(sizeof(int32_t) * n_pairs));
    CU(e->d_s1pm.need(sizeof(double) * n_pairs * (size_t)t_max));
    CU(e->d_kink.need(sizeof(double2) * n_pairs * (size_t)t_max));
    CU(e->d_s1sub_pm.need(sizeof(double) * n_pairs * (size_t)t_max));
    CU(e->d_s1sub_fs.need(sizeof(int32_t) * n_pairs));
    CU(e->d_s1sub_ft.need(sizeof(double) * n_pairs));
    CU(e->d_p1f.need(sizeof(float2) * (size_t)n_drones * (size_t)t_max));
    CU(e->d_reach.need(sizeof(int2) * n_pairs * (size_t)t_max));
    CU(e->d_alive.need(sizeof(uint32_t) * n_pairs));
    CU(e->d_reach_tot.need(sizeof(unsigned long long) * 2));
    CU(e->d_slab_dir.need(sizeof(float2) * n_pairs));
    CU(cudaEventRecord(e->ev_a, st));
    CU(cudaMemcpyAsync(e->d_scn.p, hb, scn_bytes, cudaMemcpyHostToDevice, st));
    e->h2d_scenario = scn_bytes;

    unsigned char* db = e->d_scn.as<unsigned char>();
    s.tracks = reinterpret_cast<const TrackDev*>(db + o_tracks);
    s.ax = reinterpret_cast<const double*>(db + o_ax);
    s.ay = reinterpret_cast<const double*>(db + o_ay);
    s.az = reinterpret_cast<const double*>(db + o_az);
    s.trk32 = reinterpret_cast<const float4*>(db + o_t32);
    s.trkxy = reinterpret_cast<const float2*>(db + o_txy);
    s.drones = reinterpret_cast<const double4*>(db + o_dr);
    s.p1x = e->d_p1x.as<double>(); s.p1y = e->d_p1y.as<double>();
    s.s1_first_hit = e->d_s1fh.as<int32_t>();
    s.s1_prefmin = e->d_s1pm.as<double>();
    s.kink = e->d_kink.as<double2>();
    s.p1f = e->d_p1f.as<float2>();
    s.s1sub_prefmin = e->d_s1sub_pm.as<double>();
    s.s1sub_first_seg = e->d_s1sub_fs.as<int32_t>();
    s.s1sub_first_time = e->d_s1sub_ft.as<double>();

    s.reach = e->d_reach.as<int2>();
    s.alive = e->d_alive.as<uint32_t>();
    s.reach_tot = e->d_reach_tot.as<unsigned long long>();
    s.slab_dir = e->d_slab_dir.as<float2>();

    k_stage1<<<(unsigned)((n_pairs + 3) / 4), 128, 0, st>>>(s);            // one warp per pair
    ++e->launches;
    CU(cudaGetLastError());
    e->reach_ready = false;        // the reach table is built by the first run that asks for it
    e->slab_ready = false;         // likewise the slab directions of k_screen3
    e->sub_survivors = -1.0;
    e->runs_on_scenario = 0;
    CU(cudaEventRecord(e->ev_b, st));   // no host sync here: runs queue behind it on the same stream
    e->prepare_timed = false;
    e->sd = s;
    e->model = *model;
    e->h_drones.assign(drones, drones + n_drones);
    e->scn_o_ax = o_ax; e->scn_o_ay = o_ay; e->scn_o_az = o_az;
    e->has_scenario = true;
    return DCRP_OK;
}

// Is the resident scenario bit for bit what the caller passes (again)?  Exact comparison against the
// pinned mirror of the last upload -- memory-speed, no hashing, no false positives.
static bool scenario_is_resident(const dcrp_engine* e, const dcrp_track* tracks, int32_t n_tracks,
                                 const dcrp_drone* drones, int32_t n_drones, const dcrp_model* model)
{
    if (!e->has_scenario || !tracks || !drones || !model) return false;
    if (n_tracks != e->sd.n_tracks || n_drones != e->sd.n_drones) return false;
    if (std::memcmp(model, &e->model, sizeof *model) != 0) return false;
    if (std::memcmp(drones, e->h_drones.data(), sizeof(dcrp_drone) * (size_t)n_drones) != 0) return false;
    const unsigned char* hb = e->h_scn.as<unsigned char>();
    const double* hx = reinterpret_cast<const double*>(hb + e->scn_o_ax);
    const double* hy = reinterpret_cast<const double*>(hb + e->scn_o_ay);
    const double* hz = reinterpret_cast<const double*>(hb + e->scn_o_az);
    for (int a = 0; a < n_tracks; ++a) {
        const dcrp_track& t = tracks[a];
        const TrackDev& d = e->h_tracks[(size_t)a];
        if (!t.x || !t.y || !t.z || t.n_steps != d.n_steps || t.takeoff_t != d.takeoff_t) return false;
        if (std::memcmp(t.box, d.box, sizeof d.box) != 0) return false;
        const size_t bytes = sizeof(double) * (size_t)t.n_steps;
        if (std::memcmp(t.x, hx + d.off, bytes) != 0 || std::memcmp(t.y, hy + d.off, bytes) != 0 ||
            std::memcmp(t.z, hz + d.off, bytes) != 0)
            return false;
    }
    return true;
}

// ------------------------------------------------------------------ run
// k_screen2's one configuration (profiles/r01_screen_cfg_sweep.txt has the sweep that chose it): R = 2 rounds,
// 384 threads, 2 CTAs per SM, NP = 2 packed pairs per lane -> 3072-trial items.
constexpr int S2_R = 2, S2_BLOCK = 384, S2_MINB = 2, S2_NP = 2;
static const void* const g_screen2[3] = {                   // plain, replay table, reach culling
    (const void*)k_screen2<S2_R, S2_BLOCK, false, S2_MINB, S2_NP, false, false>,
    (const void*)k_screen2<S2_R, S2_BLOCK, true, S2_MINB, S2_NP, false, false>,
    (const void*)k_screen2<S2_R, S2_BLOCK, false, S2_MINB, S2_NP, true, false>};
static uint64_t screen_chunk() { return (uint64_t)2 * S2_NP * S2_R * S2_BLOCK; }

// MaxDynamicSharedMemorySize belongs to the (function, device) pair, not to an engine: another engine on the same
// device may have set it for a smaller scenario.  So it is always RAISED to the most the device allows (opt-in
// maximum minus the kernel's static shared memory), never set to "what this launch needs".
static cudaError_t allow_max_dynamic_smem(const dcrp_engine* e, const void* kern)
{
    cudaFuncAttributes fa{};
    cudaError_t ce = cudaFuncGetAttributes(&fa, kern);
    if (ce != cudaSuccess) return ce;
    const int most = (int)e->prop.sharedMemPerBlockOptin - (int)fa.sharedSizeBytes;
    if (fa.maxDynamicSharedSizeBytes >= most) return cudaSuccess;
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, most);
}

static cudaError_t launch_screen2(dcrp_engine* e, const RunDev& r, uint32_t chunks_per_pair, uint32_t n_items,
                                  cudaStream_t st)
{
    const int v = r.replay ? 1 : (r.cull ? 2 : 0);     // r.cull is never set together with a replay table
    const void* kern = g_screen2[v];
    int tile_steps = e->sd.n_steps_max <= 128 ? 128 : 512;
    const size_t smem = (size_t)2 * tile_steps * 8 + (size_t)screen_chunk() * 16 + 80 * 4 + 16 +
                        (size_t)SCREEN_TAB * (sizeof(double2) + sizeof(float2) + sizeof(int2)) +
                        (2 * S2_NP * S2_R > 4 ? (size_t)screen_chunk() * 2 : 0);   // order[] of the K > 4 variants
    if (e->screen_smem[v] != smem) {
        cudaError_t ce = allow_max_dynamic_smem(e, kern);
        if (ce != cudaSuccess) return ce;
        ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&e->screen_occ[v], kern, S2_BLOCK, smem);
        if (ce != cudaSuccess) return ce;
        if (e->screen_occ[v] < 1) e->screen_occ[v] = 1;
        e->screen_smem[v] = smem;
    }
    const uint32_t grid = (uint32_t)std::min<uint64_t>(
        n_items, (uint64_t)(e->prop.multiProcessorCount - e->reserve_sms) * e->screen_occ[v]);
    uint32_t stagger_ns = 0, dbg = 0;      // kernel-side profiling switches of round 1: never set in the product
    ScenarioDev sd = e->sd;
    RunDev rr = r;
    void* args[] = {&sd, &rr, &chunks_per_pair, &n_items, &tile_steps, &stagger_ns, &dbg};
    cudaError_t ce = cudaLaunchKernel(kern, dim3(grid), dim3((unsigned)S2_BLOCK), args, smem, st);
    ++e->launches; ++e->launches_this_run;
    return ce != cudaSuccess ? ce : cudaGetLastError();
}

// The block schedule of a k_screen3 launch (screen3.cu S3Sched): `n_static` items are dealt out statically, the
// rest in blocks numbered by the work counter.  A block size 2^lg is used while at least twice a full round of it
// (2 * n_warps * 2^lg items) is left, then the size halves; the last 4 * n_warps items go one by one.
static S3Sched s3_schedule(uint32_t n_items, uint32_t n_warps, uint32_t n_static)
{
    S3Sched sc{};
    uint64_t pos = n_static, blk = 0, left = n_items > n_static ? n_items - n_static : 0;
    for (int lg = S3_LG_GMAX; lg >= 0 && left > 0; --lg) {
        const uint64_t size = 1ull << lg, keep = 2ull * n_warps * size;
        uint64_t nb;
        if (lg > 0) {
            if (left < keep) continue;
            nb = (left - keep) / size + 1;
        } else nb = left;
        sc.item0[sc.n_ph] = (uint32_t)pos; sc.blk0[sc.n_ph] = (uint32_t)blk; sc.lg[sc.n_ph] = (uint32_t)lg;
        ++sc.n_ph;
        pos += nb * size; blk += nb; left -= nb * size;
    }
    sc.item0[sc.n_ph] = n_items; sc.blk0[sc.n_ph] = (uint32_t)blk;
    return sc;
}

// The schedule as the kernel reads it (screen3.cu grab_take), block by block: for the CPU test tier.
extern "C" int64_t dcrp_debug_block_schedule(uint32_t n_items, uint32_t n_warps, uint32_t n_static,
                                             uint32_t* begin, uint32_t* end, int64_t cap)
{
    if (n_warps == 0 || (cap > 0 && (!begin || !end))) return fail(DCRP_ERR_INVALID, "dcrp_debug_block_schedule: bad arguments");
    const S3Sched sc = s3_schedule(n_items, n_warps, n_static);
    const int64_t n_blocks = sc.blk0[sc.n_ph];
    for (int64_t k = 0; k < n_blocks && k < cap; ++k) {
        uint32_t b = n_items, e = n_items;
        for (int p = 0; p < S3_MAX_PH; ++p)
            if ((uint32_t)p < sc.n_ph && (uint32_t)k >= sc.blk0[p] && (uint32_t)k < sc.blk0[p + 1]) {
                b = sc.item0[p] + (((uint32_t)k - sc.blk0[p]) << sc.lg[p]);
                e = std::min(b + (1u << sc.lg[p]), sc.item0[p + 1]);
            }
        begin[k] = b; end[k] = e;
    }
    return n_blocks;
}

static int launch_screen3(dcrp_engine* e, const RunDev& r, uint32_t chunks_per_pair, uint32_t n_items,
                          cudaStream_t st, bool behind_run_init)
{
    // Two builds.  k_screen3<2>: 128 registers, 2 CTAs = 4 warps per SM sub-partition -- the slab loop and Philox carry
    // 8 independent trials per lane, so four warps fill the dispatch port, the only spills left are the 232 B around
    // the out-of-line re-scan call, and fewer, faster warps leave a shorter ramp and tail: faster than 3 CTAs x 80 registers and 4 CTAs x 64 registers at
    // every launch size on both rings (1.0e7 trials on the 5 km ring 0.148 against 0.152 / 0.159 ms, 1.0e8 trials 1.232
    // against 1.267 / 1.272 ms; one CTA of 172 registers 0.163 / 1.310).  k_screen3<4>: 64 registers, 8 warps per
    // sub-partition, for long launches over MANY tracks (the 600-flight day on the 1 km ring: a tenth of the trials
    // go through the re-scan and its latency-bound rare path; 20.3 against 21.1 ms).
    const int wide0 = e->sd.n_tracks > 1 &&
                      (uint64_t)n_items >= (uint64_t)48 * e->prop.multiProcessorCount * 4 * (S3_BLOCK / 32) ? 1 : 0;
    // (the sub-step builds: same kernel, thresholds widened by half a relative step and the segment forms of levels 2 / 3)
    static const void* const kerns[4] = {(const void*)k_screen3<2, false>, (const void*)k_screen3<4, false>,
                                         (const void*)k_screen3<2, true>, (const void*)k_screen3<4, true>};
    const int wide = wide0 + (r.substep ? 2 : 0);
    const void* kern = kerns[wide];
    int vl_pad = (e->sd.n_steps_max + 3) & ~3;
    int tk_pad = (e->sd.t_max + 3) & ~3;
    const size_t smem = s3_warp_smem(vl_pad, tk_pad) * (S3_BLOCK / 32);
    if (e->s3_smem[wide] != smem) {
        CU(allow_max_dynamic_smem(e, kern));
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&e->s3_occ[wide], kern, S3_BLOCK, smem));
        if (e->s3_occ[wide] < 1) e->s3_occ[wide] = 1;
        e->s3_smem[wide] = smem;
    }
    const uint64_t ctas_max = (uint64_t)(e->prop.multiProcessorCount - e->reserve_sms) * e->s3_occ[wide];
    // Every warp starts on a static block of consecutive items.  A launch of up to four items per warp of a full grid
    // is latency-bound (its duration is the longest chain of items any warp runs): it is dealt out statically, as few
    // warps as give every warp the same number of items, one staging of the pair's tables per block.  Longer launches
    // start on two items per warp and take the rest from the work counter in guided blocks (screen3.cu).
    const uint64_t warps_max = ctas_max * (S3_BLOCK / 32);
    uint32_t first_block = (uint64_t)n_items <= 4 * warps_max
        ? (uint32_t)std::max<uint64_t>(1, ((uint64_t)n_items + warps_max - 1) / warps_max) : 2u;
    const uint64_t per_cta = (uint64_t)first_block * (S3_BLOCK / 32);
    const uint64_t ctas_wanted = ((uint64_t)n_items + per_cta - 1) / per_cta;
    const uint32_t grid = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(ctas_wanted, ctas_max));
    // the first launch of a run finds the counter cleared by k_run_init (word 2 of the output arena)
    if (r.slice_first != 0) CU(cudaMemsetAsync(e->d_out.as<uint32_t>() + 2, 0, 4, st));
    ScenarioDev sd = e->sd;
    RunDev rr = r;
    rr.work_counter = e->d_out.as<uint32_t>() + 2;
    S3Sched sch = s3_schedule(n_items, grid * (S3_BLOCK / 32), grid * (S3_BLOCK / 32) * first_block);
    void* args[] = {&sd, &rr, &chunks_per_pair, &n_items, &vl_pad, &tk_pad, &first_block, &sch};
    // Directly behind k_run_init the grid is launched programmatically dependent: its first items overlap the clearing
    // launch and the launch latency (the kernel waits before it first touches the arena, screen3.cu).
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(S3_BLOCK); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr; cfg.numAttrs = behind_run_init ? 1 : 0;
    const cudaError_t ce = behind_run_init ? cudaLaunchKernelExC(&cfg, kern, args)
                                           : cudaLaunchKernel(kern, dim3(grid), dim3(S3_BLOCK), args, smem, st);
    ++e->launches; ++e->launches_this_run;
    if (ce != cudaSuccess)
        return fail(DCRP_ERR_CUDA, "k_screen3 launch (grid %u, %zu B shared memory, device %d): %s", grid, smem, e->device,
                    cudaGetErrorString(ce));
    CU(cudaGetLastError());
    return DCRP_OK;
}

extern "C" int dcrp_run_async(dcrp_engine* e, const dcrp_run* run, void* stream, uint64_t* d_counts)
{
    NvtxRange nvtx_("dcrp_run_async");
    if (!e || !run) return fail(DCRP_ERR_INVALID, "dcrp_run_async: NULL argument");
    if (!e->has_scenario) return fail(DCRP_ERR_STATE, "dcrp_run_async before dcrp_scenario_upload");
    if (run->precision < DCRP_PRECISION_FP64 || run->precision > DCRP_PRECISION_FP32)
        return fail(DCRP_ERR_INVALID, "unknown precision %d", run->precision);
    if ((run->flags & DCRP_FLAG_PER_TRIAL) && run->precision != DCRP_PRECISION_FP64)
        return fail(DCRP_ERR_INVALID, "DCRP_FLAG_PER_TRIAL needs DCRP_PRECISION_FP64");
    if ((run->flags & DCRP_FLAG_SUBSTEP) && run->precision == DCRP_PRECISION_FP32)
        return fail(DCRP_ERR_INVALID, "DCRP_FLAG_SUBSTEP needs DCRP_PRECISION_FP64 or DCRP_PRECISION_MIXED");
    if ((run->flags & DCRP_FLAG_SUBSTEP) && e->sd.n_steps_max < 2)
        return fail(DCRP_ERR_INVALID, "DCRP_FLAG_SUBSTEP needs tracks of at least 2 samples");
    if (run->trial_begin + run->trial_count < run->trial_begin)
        return fail(DCRP_ERR_INVALID, "trial range overflows 64 bits");
    CU(cudaSetDevice(e->device));
    (void)cudaGetLastError();      // a stale error left in this thread by other CUDA users (torch, NCCL) is not ours to report
    cudaStream_t st = stream ? (cudaStream_t)stream : e->stream;
    const ScenarioDev& s = e->sd;
    const size_t n_pairs = (size_t)s.n_tracks * (size_t)s.n_drones;

    e->has_run = false;            // until every launch below is queued: a run that failed half-way cannot be fetched
    e->run = *run;
    e->allpairs_run = false;
    e->run_stream_prev = e->run_stream;
    e->run_stream = st;
    e->launches_this_run = 0;
    e->h2d_run = 0;
    e->rec_cap = run->record_capacity ? run->record_capacity : DEFAULT_RECORD_CAP;
    e->external_counts = (d_counts != nullptr);

    CU(size_out_arena(e, n_pairs));
    CU(e->d_cands.need(sizeof(Candidate) * (size_t)CAND_CAP));
    if (st != e->stream) CU(cudaStreamWaitEvent(st, e->ev_b, 0));   // after the scenario upload + stage 1
    if (e->run_event_valid && st != e->run_stream_prev) CU(cudaStreamWaitEvent(st, e->ev_end, 0));   // the arenas are shared between runs

    RunDev r{};
    r.seed = run->seed; r.trial_begin = run->trial_begin; r.trial_count = run->trial_count;
    r.track_base = run->track_id_base; r.drone_base = run->drone_id_base;
    r.replay = nullptr;
    if (run->replay && run->trial_count) {
        const size_t bytes = sizeof(Draw) * n_pairs * (size_t)run->trial_count;
        if (!e->reuse_device_replay) {
            // The kernels index per-pair tables with t1st: a row from another track's table (or garbage) must be
            // DCRP_ERR_INVALID here, not an out-of-bounds access on the device.
            for (size_t p = 0; p < n_pairs; ++p) {
                const int tk = e->h_tracks[p / (size_t)s.n_drones].takeoff_t;
                const dcrp_draw* row = run->replay + p * (size_t)run->trial_count;
                for (uint64_t k = 0; k < run->trial_count; ++k) {
                    const dcrp_draw& d = row[k];
                    if (d.t1st < 1 || d.t1st > tk || !std::isfinite(d.tx) || !std::isfinite(d.ty) || !std::isfinite(d.tz))
                        return fail(DCRP_ERR_INVALID,
                                    "replay table: row %llu of pair %zu has t1st=%d (must be in [1, takeoff_t=%d]) or a "
                                    "non-finite target", (unsigned long long)k, p, d.t1st, tk);
                }
            }
            CU(e->d_replay.need(bytes));
            CU(cudaMemcpyAsync(e->d_replay.p, run->replay, bytes, cudaMemcpyHostToDevice, st));
            e->h2d_run += bytes;
        }
        e->reuse_device_replay = false;
        r.replay = e->d_replay.as<Draw>();
    }
    unsigned long long* own_counts =
        reinterpret_cast<unsigned long long*>(e->d_out.as<unsigned char>() + e->out_counts_off);
    r.counts = d_counts ? reinterpret_cast<unsigned long long*>(d_counts) : own_counts;
    r.first_conflict = reinterpret_cast<int32_t*>(e->d_out.as<unsigned char>() + e->out_first_off);
    r.records = out_records(e);
    r.rec_count = small_rec_count(e);
    r.rec_cap = e->rec_cap;
    r.cands = e->d_cands.as<Candidate>();
    r.cand_count = small_cand_count(e);
    r.cand_cap = CAND_CAP;
    r.stats = small_stats(e);
    r.decide_fp32 = (run->precision == DCRP_PRECISION_FP32) ? 1 : 0;
    r.substep = (run->flags & DCRP_FLAG_SUBSTEP) ? 1 : 0;
    philox_round_keys(run->seed, r.rk);
    // exact reach culling: screen kernel only, sample mode only, no per-trial outputs
    r.cull = ((run->flags & DCRP_FLAG_REACH_CULL) && run->precision == DCRP_PRECISION_MIXED && !r.substep &&
              !(run->flags & DCRP_FLAG_PER_TRIAL) && !run->replay) ? 1 : 0;
    e->run_cull = r.cull != 0;
    e->run_status &= DCRP_STAT_FP64_FALLBACK;      // (set by dcrp_fetch just before it re-runs in fp64)
    if ((run->flags & DCRP_FLAG_REACH_CULL) && !r.cull) e->run_status |= DCRP_STAT_CULL_IGNORED;
    if (r.cull && !e->reach_ready) {
        CU(cudaMemsetAsync(e->d_reach_tot.p, 0, sizeof(unsigned long long) * 2, st));
        k_reach<<<(unsigned)((n_pairs + 3) / 4), 128, 0, st>>>(s);      // one warp per pair, once per scenario
        ++e->launches; ++e->launches_this_run;
        CU(cudaGetLastError());
        e->reach_ready = true;
    }
    if (run->flags & DCRP_FLAG_PER_TRIAL) {
        const size_t n = n_pairs * (size_t)run->trial_count;
        CU(e->d_trial_hit.need(n ? n : 1));
        CU(e->d_trial_first.need(sizeof(int32_t) * (n ? n : 1)));
        CU(e->d_trial_min.need(sizeof(double) * (n ? n : 1)));
        r.trial_hit = e->d_trial_hit.as<uint8_t>();
        r.trial_first = e->d_trial_first.as<int32_t>();
        r.trial_min = e->d_trial_min.as<double>();
    }

    // slices keep the item count of one launch below 2^30
    const bool exact = (run->precision == DCRP_PRECISION_FP64);
    // the slab-first screen serves MIXED runs (sample and sub-step mode) on tracks its per-warp tables hold; everything
    // else (replay tables, reach culling, fp32 decisions, long tracks) goes through k_screen2
    // Sub-step mode widens the slab by half a relative step (tens of metres): around a 5 km ring it still clears 90 % of
    // the trials (0.32 against 0.42 ms per 1.0e7 trials), around a 1 km ring 60 % -- there the 2-D screen alone is the
    // faster one (0.74 against 0.81 ms).  Same answers either way, so the choice follows what the previous sub-step run
    // on this scenario saw (dcrp_fetch keeps the survivor fraction).
    const bool sub_prefers_2d = r.substep && e->sub_survivors > 0.25;
    const bool use3 = run->precision == DCRP_PRECISION_MIXED && !r.replay && !r.cull && !sub_prefers_2d &&
                      !(run->flags & DCRP_FLAG_SCREEN_2D) && s.n_steps_max <= S3_MAX_VL && s.t_max <= S3_MAX_T;
    e->run_screen3 = use3;
    const uint64_t per_item = exact ? (uint64_t)EXB : (use3 ? (uint64_t)S3_CHW : screen_chunk());
    // DCRP_MAX_ITEMS: test aid, forces the slicing below at small sizes (tests/test_gpu_parity.py)
    static const uint64_t max_items = std::getenv("DCRP_MAX_ITEMS")
        ? std::max<uint64_t>(1, std::strtoull(std::getenv("DCRP_MAX_ITEMS"), nullptr, 10)) : MAX_ITEMS_PER_LAUNCH;
    const uint64_t max_chunks = std::max<uint64_t>(1, max_items / n_pairs);
    const uint64_t slice_max = max_chunks * per_item;

    // Published run (dcrp_engine::published): one k_screen3 launch with engine-owned counts.  Without a communicator
    // (whose all-reduce may still want the counts on the device) the kernel also leaves the arena cleared, and a run
    // that finds it so needs no clearing launch in front of it.
    // (one CTA storing over PCIe: for the scenarios Drone::Simulation and the benchmark run -- up to a few thousand
    // pairs; the 713 KB arena of the 600-flight day takes it 0.24 ms, the copy engine 0.03)
    e->published = use3 && run->trial_count > 0 && run->trial_count <= slice_max && !d_counts && e->out_bytes <= PUBLISH_MAX_BYTES;
    const bool self_clean = e->published && !e->comm;
    const bool arena_is_clean = e->arena_clean && e->clean_ptr == e->d_out.p && e->clean_bytes == e->out_bytes;
    const bool need_init = !(e->published && arena_is_clean);
    if (e->published && ++e->pub_seq == 0) ++e->pub_seq;

    // DCRP_FLAG_NO_TIMING: no timing events around the kernels (each is a microsecond of host time in front of the
    // trial kernel, and one between k_screen3 and k_publish keeps the two from being queued back to back)
    const bool timed = !(run->flags & DCRP_FLAG_NO_TIMING);
    e->run_timed = timed;
    HOST_MARK(2);
    if (timed) CU(cudaEventRecord(e->ev_begin, st));
    HOST_MARK(3);
    // The per-pair choice of the slab direction costs ~50 us (1024 sample trials per pair) and saves ~10 % of a kernel:
    // it pays off for one big run (>= 3e5 trials per pair) or once a scenario is run AGAIN with >= 2e4 trials per pair
    // (a bench loop, a study streaming trial ranges).  A one-off dcrp_simulate on a new scenario, Drone::Simulation's
    // 10 000-trial calls and the 600-flight day keep the per-track default.
    if (use3 && !e->slab_ready && run->trial_count >= 20000 && (run->trial_count >= 300000 || e->runs_on_scenario >= 1)) {
        k_slab_axis<<<(unsigned)n_pairs, S3_AXIS_BLOCK, 0, st>>>(s, 1.5f * s.thr_slab);   // once per scenario
        ++e->launches; ++e->launches_this_run;
        CU(cudaGetLastError());
        e->slab_ready = true;
    }
    // one launch clears the whole output arena (counters, stats, counts) and sets first_conflict = INT32_MAX
    if (need_init) {
        const bool zero_ext = d_counts && (run->flags & DCRP_FLAG_ZERO_COUNTS);
        const size_t words = std::max(e->out_bytes / 4, zero_ext ? n_pairs : (size_t)0);
        k_run_init<<<(unsigned)((words + 255) / 256), 256, 0, st>>>(
            e->d_out.as<uint32_t>(), (uint32_t)(e->out_first_off / 4), (uint32_t)(e->out_bytes / 4),
            zero_ext ? reinterpret_cast<unsigned long long*>(d_counts) : nullptr, zero_ext ? (uint32_t)n_pairs : 0u);
        ++e->launches; ++e->launches_this_run;
        CU(cudaGetLastError());
    }
    // whatever runs next finds the arena as this run leaves it
    e->arena_clean = self_clean;
    e->clean_ptr = e->d_out.p;
    e->clean_bytes = e->out_bytes;
    HOST_MARK(4);
    // ms_trials starts here: the trial kernel(s) alone.  (Not for k_screen3: launched programmatically behind
    // k_run_init an event between the two would serialise them, and ms_trials includes the 2 us it overlaps.)
    const bool pdl = use3 && need_init;
    if (!use3 && timed) CU(cudaEventRecord(e->ev_k0, st));
    e->k0_is_begin = use3;
    for (uint64_t first = 0; first < run->trial_count; first += slice_max) {
        r.slice_first = first;
        r.slice_count = std::min<uint64_t>(slice_max, run->trial_count - first);
        const uint32_t chunks = (uint32_t)((r.slice_count + per_item - 1) / per_item);
        const uint32_t n_items = (uint32_t)(chunks * n_pairs);
        if (exact) {
            const size_t smem = sizeof(double) * 3 * (size_t)s.n_steps_max;
            const int use_smem = smem <= 48 * 1024 ? 1 : 0;
            const uint32_t grid = (uint32_t)std::min<uint64_t>(n_items, (uint64_t)e->prop.multiProcessorCount * 8);
            k_exact<<<grid, EXB, use_smem ? smem : 0, st>>>(s, r, chunks, n_items, use_smem);
            ++e->launches; ++e->launches_this_run;
            CU(cudaGetLastError());
        } else {
            r.div_chunks = make_fastdiv(chunks);
            if (use3) {
                const int rc3 = launch_screen3(e, r, chunks, n_items, st, pdl && first == 0);
                if (rc3 != DCRP_OK) return rc3;
            } else {
                const cudaError_t ce = launch_screen2(e, r, chunks, n_items, st);
                if (ce != cudaSuccess) return fail(DCRP_ERR_CUDA, "k_screen2 launch (device %d): %s", e->device, cudaGetErrorString(ce));
            }
        }
    }
    e->has_confirm = run->precision == DCRP_PRECISION_MIXED && run->trial_count && !use3;    // k_screen3 confirms in-kernel
    if (e->has_confirm) {
        if (timed) CU(cudaEventRecord(e->ev_trials, st));
        // grid-stride over the candidate list whose length lives on the device
        k_confirm_all<<<e->prop.multiProcessorCount * 4, 128, 0, st>>>(s, r);
        ++e->launches; ++e->launches_this_run;
        CU(cudaGetLastError());
    }
    if (e->published) {
        if (timed) CU(cudaEventRecord(e->ev_trials, st));       // the end of the trial kernel
        PublishArgs pa{};
        CU(cudaHostGetDevicePointer(reinterpret_cast<void**>(&pa.host), e->h_out.p, 0));
        pa.arena = e->d_out.as<uint32_t>();
        pa.records = reinterpret_cast<const uint32_t*>(out_records(e));
        pa.words = (uint32_t)(e->out_bytes / 4);
        pa.first_word = (uint32_t)(e->out_first_off / 4);
        pa.rec_cap = e->rec_cap;
        pa.inline_recs = std::min(e->rec_cap, INLINE_RECORDS);
        pa.flag_word = (uint32_t)(pub_flag_off(e) / 4);
        // The flag's place in the mirror moves with the scenario's size: whatever an earlier run left there (a count,
        // a first-conflict index -- small integers, like sequence numbers) must not read as "delivered".
        *reinterpret_cast<volatile uint32_t*>(e->h_out.as<unsigned char>() + pub_flag_off(e)) = 0u;
        pa.seq = e->pub_seq;
        pa.clean = self_clean ? 1 : 0;
        k_publish<<<1, 256, 0, st>>>(pa);
        ++e->launches; ++e->launches_this_run;
        CU(cudaGetLastError());
    }
    HOST_MARK(5);
    CU(cudaEventRecord(e->ev_end, st));      // (without a confirm kernel this is also the end of the trial kernels)
    HOST_MARK(6);
    e->run_event_valid = true;
    e->has_run = true;
    ++e->runs_on_scenario;
    return DCRP_OK;
}

#ifdef DCRP_TIMELINE
// instrumented build only (screen3.cu): the per-warp time stamps of the last k_screen3 launch
extern "C" int dcrp_debug_timeline(unsigned long long* out, int n_words)
{
    if (!out || n_words <= 0) return fail(DCRP_ERR_INVALID, "dcrp_debug_timeline: bad arguments");
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpyFromSymbol(out, dcrp::g_timeline, sizeof(unsigned long long) * (size_t)std::min(n_words, TL_WARPS * TL_SLOTS)));
    return DCRP_OK;
}
#endif

extern "C" int dcrp_last_run_ms(dcrp_engine* e, float* ms)
{
    if (!e || !ms) return fail(DCRP_ERR_INVALID, "dcrp_last_run_ms: NULL argument");
    if (!e->has_run) return fail(DCRP_ERR_STATE, "no run to time");
    if (!e->run_timed) return fail(DCRP_ERR_STATE, "the last run was not timed (DCRP_FLAG_NO_TIMING)");
    CU(cudaSetDevice(e->device));
    (void)cudaGetLastError();      // a stale error left in this thread by other CUDA users (torch, NCCL) is not ours to report
    CU(cudaEventSynchronize(e->ev_end));
    CU(cudaEventElapsedTime(ms, e->ev_begin, e->ev_end));
    return DCRP_OK;
}

extern "C" int dcrp_fetch(dcrp_engine* e, dcrp_result* res)
{
    NvtxRange nvtx_("dcrp_fetch");
    if (!e || !res) return fail(DCRP_ERR_INVALID, "dcrp_fetch: NULL argument");
    if (!e->has_run) return fail(DCRP_ERR_STATE, "dcrp_fetch before dcrp_run_async");
    CU(cudaSetDevice(e->device));
    (void)cudaGetLastError();      // a stale error left in this thread by other CUDA users (torch, NCCL) is not ours to report
    cudaStream_t st = e->run_stream;
    const ScenarioDev& s = e->sd;
    const size_t n_pairs = (size_t)s.n_tracks * (size_t)s.n_drones;
    // one D2H copy of the output arena and the first few records into the pinned mirror
    const uint32_t inl = res->records ? std::min(e->rec_cap, INLINE_RECORDS) : 0u;
    HOST_MARK(7);
    bool have = false;              // the kernel has delivered the results itself
    if (e->published) {
        // poll the flag behind the mirror (short runs), then fall back to waiting for the stream (long runs; a run
        // that failed never raises the flag and the error surfaces here)
        volatile uint32_t* flag = reinterpret_cast<volatile uint32_t*>(e->h_out.as<unsigned char>() + pub_flag_off(e));
        const auto t0 = std::chrono::steady_clock::now();
        while (!(have = (*flag == e->pub_seq)) && std::chrono::steady_clock::now() - t0 < std::chrono::milliseconds(2)) {}
        if (!have) {
            const cudaError_t q = cudaStreamSynchronize(st);
            if (q != cudaSuccess) return fail(DCRP_ERR_CUDA, "waiting for the run failed: %s", cudaGetErrorString(q));
            have = (*flag == e->pub_seq);
        }
        std::atomic_thread_fence(std::memory_order_acquire);
    }
    if (!have) {
    CU(cudaMemcpyAsync(e->h_out.p, e->d_out.p, e->out_bytes + sizeof(Record) * (size_t)inl, cudaMemcpyDeviceToHost, st));
    HOST_MARK(8);
    {   // short runs: poll (a blocking sync wakes up 10-20 us late, a tenth of a 1e7-trial run); long runs: block
        const auto t0 = std::chrono::steady_clock::now();
        cudaError_t q;
        while ((q = cudaStreamQuery(st)) == cudaErrorNotReady &&
               std::chrono::steady_clock::now() - t0 < std::chrono::milliseconds(2)) {}
        if (q == cudaErrorNotReady) q = cudaStreamSynchronize(st);
        if (q != cudaSuccess) return fail(DCRP_ERR_CUDA, "waiting for the run failed: %s", cudaGetErrorString(q));
    }
    }
    HOST_MARK(9);
    if (e->comm && e->comm->h_err && *e->comm->h_err)
        return fail(DCRP_ERR_COMM, "count exchange: a peer rank did not deliver its counts within 10 s");
    const unsigned char* small = e->h_out.as<unsigned char>();
    uint32_t rec_count, cand_count;
    unsigned long long stats[ST_COUNT];
    std::memcpy(&rec_count, small, 4);
    std::memcpy(&cand_count, small + 4, 4);
    std::memcpy(stats, small + 16, sizeof stats);
    uint64_t d2h = e->out_bytes + sizeof(Record) * (size_t)inl;
    if (have)       // what k_publish stored into the mirror: the arena, the records there are, the flag
        d2h = e->out_bytes + sizeof(Record) * (size_t)std::min(std::min(rec_count, e->rec_cap), INLINE_RECORDS) + 4;

    if (e->run.precision == DCRP_PRECISION_MIXED && cand_count > CAND_CAP) {
        // More candidates than the list holds (absurd hit rates, e.g. a huge radius):
        // redo the whole run with the fp64 kernel -- still on the GPU, same answers.
        if (e->external_counts)
            return fail(DCRP_ERR_STATE, "candidate list overflow with caller-owned counts: rerun with DCRP_PRECISION_FP64");
        dcrp_run again = e->run;
        again.precision = DCRP_PRECISION_FP64;
        e->reuse_device_replay = again.replay != nullptr;   // the table is already in d_replay: the caller's copy may be gone
        e->run_status = DCRP_STAT_FP64_FALLBACK;
        int rc = dcrp_run_async(e, &again, (void*)st, nullptr);
        if (rc != DCRP_OK) { e->run_status = 0; return rc; }
        e->run.precision = DCRP_PRECISION_MIXED;   // report what was asked for
        rc = dcrp_fetch(e, res);
        e->run_status = 0;
        return rc;
    }

    if (res->counts && !e->external_counts) {
        const uint64_t* h = reinterpret_cast<const uint64_t*>(small + e->out_counts_off);
        for (size_t i = 0; i < n_pairs; ++i) res->counts[i] += h[i];     // accumulate, Drone.cpp:276
    }
    if (res->first_conflict)
        std::memcpy(res->first_conflict, small + e->out_first_off, sizeof(int32_t) * n_pairs);
    res->n_records = 0;
    res->overflow = rec_count > e->rec_cap ? 1 : 0;
    if (res->records) {
        const uint32_t n = std::min(rec_count, e->rec_cap);
        if (n) {
            std::memcpy(res->records, small + e->out_bytes, sizeof(Record) * (size_t)std::min(n, inl));
            if (n > inl) {      // more hits than travel with the first copy
                CU(cudaMemcpy(res->records + inl, out_records(e) + inl, sizeof(Record) * (size_t)(n - inl), cudaMemcpyDeviceToHost));
                d2h += sizeof(Record) * (size_t)(n - inl);
            }
            std::sort(res->records, res->records + n, [](const dcrp_record& a, const dcrp_record& b) {
                if (a.track != b.track) return a.track < b.track;
                if (a.drone != b.drone) return a.drone < b.drone;
                return a.trial < b.trial;
            });
        }
        res->n_records = n;
    }
    if (e->run.flags & DCRP_FLAG_PER_TRIAL) {
        const size_t n = n_pairs * (size_t)e->run.trial_count;
        if (res->trial_hit && n) { CU(cudaMemcpy(res->trial_hit, e->d_trial_hit.p, n, cudaMemcpyDeviceToHost)); d2h += n; }
        if (res->trial_first_hit && n) { CU(cudaMemcpy(res->trial_first_hit, e->d_trial_first.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost)); d2h += 4 * n; }
        if (res->trial_min_dist && n) { CU(cudaMemcpy(res->trial_min_dist, e->d_trial_min.p, sizeof(double) * n, cudaMemcpyDeviceToHost)); d2h += 8 * n; }
    }

    dcrp_stats& o = res->stats;
    std::memset(&o, 0, sizeof o);
    o.trials = e->allpairs_run ? e->ap_rays : (uint64_t)n_pairs * e->run.trial_count;
    // the kernels accumulate sum(t1st) (= stage-1 tests shared per pair) and the issued lane-steps;
    // evaluated stage-2 tests and drone-steps follow from the track lengths
    uint64_t all_tests = 0, all_steps = 0;
    for (int a = 0; a < s.n_tracks; ++a) {
        all_tests += (uint64_t)e->h_tracks[(size_t)a].n_steps * e->run.trial_count * (uint64_t)s.n_drones;
        all_steps += (uint64_t)(e->h_tracks[(size_t)a].n_steps - 1) * e->run.trial_count * (uint64_t)s.n_drones;
    }
    o.tests_shared = stats[ST_SHARED];
    o.tests_evaluated = all_tests - stats[ST_SHARED];
    o.tests_issued = stats[ST_ISSUED];
    o.drone_steps = all_steps;
    if (e->allpairs_run) {            // every ray against every track: the scan kernel counted them
        o.tests_evaluated = stats[ST_EVALUATED];
        o.tests_shared = 0;
        o.drone_steps = 0;
    }
    if (!e->allpairs_run && e->run_cull) {
        // evaluated and culled tests are reported separately.  Processed pairs: every stage-2
        // test is either evaluated or culled; cleared pairs: the guaranteed minimum, all culled.
        o.trials_culled = stats[ST_CULLED_TRIALS];
        o.tests_culled = stats[ST_CULLED_TESTS] + stats[ST_DEAD_TESTS];
        o.tests_evaluated = stats[ST_PROCESSED_TESTS] - stats[ST_SHARED] - stats[ST_CULLED_TESTS];
        o.drone_steps = 0;
    }
    o.candidates = stats[ST_CANDIDATES];
    o.level2 = stats[ST_LEVEL2];
    o.slab_survivors = stats[ST_SLAB_SURVIVORS];
    if (!e->allpairs_run && e->run_screen3 && (e->run.flags & DCRP_FLAG_SUBSTEP) && o.trials >= 100000)
        e->sub_survivors = (double)o.slab_survivors / (double)o.trials;
    o.hits = stats[ST_HITS];
    o.kernel_launches = e->launches_this_run;
    o.status = e->run_status;
    if (!e->prepare_timed) {
        CU(cudaEventElapsedTime(&e->ms_prepare, e->ev_a, e->ev_b));
        e->prepare_timed = true;
    }
    o.ms_prepare = e->ms_prepare;
    float ms = 0.f;
    if (e->run.flags & DCRP_FLAG_NO_TIMING) {
        // the caller does not read them (the end-of-run event of a published run completes a few microseconds after
        // the flag: not waited for)
    } else if (have) {
        cudaError_t q;
        while ((q = cudaEventQuery(e->ev_end)) == cudaErrorNotReady) {}
        if (q != cudaSuccess) return fail(DCRP_ERR_CUDA, "waiting for the run failed: %s", cudaGetErrorString(q));
        CU(cudaEventElapsedTime(&ms, e->ev_begin, e->ev_trials));    o.ms_trials = ms;
        CU(cudaEventElapsedTime(&ms, e->ev_begin, e->ev_end));       o.ms_device_total = ms;
    } else {
    // as few event queries as the run needs (each is a driver call; a 1e7-trial run is 150 us)
    cudaEvent_t trials_end = e->has_confirm ? e->ev_trials : e->ev_end;
    CU(cudaEventElapsedTime(&ms, e->ev_begin, e->ev_end));    o.ms_device_total = ms;
    if (e->k0_is_begin && !e->has_confirm) o.ms_trials = ms;
    else { CU(cudaEventElapsedTime(&ms, e->k0_is_begin ? e->ev_begin : e->ev_k0, trials_end)); o.ms_trials = ms; }
    if (e->has_confirm) { CU(cudaEventElapsedTime(&ms, e->ev_trials, e->ev_end)); o.ms_confirm = ms; }
    }
    o.h2d_bytes = e->h2d_run;
    o.d2h_bytes = d2h;
    return DCRP_OK;
}

extern "C" int dcrp_simulate(dcrp_engine* e, const dcrp_track* tracks, int32_t n_tracks,
                             const dcrp_drone* drones, int32_t n_drones, const dcrp_model* model,
                             const dcrp_run* run, dcrp_result* result)
{
    NvtxRange nvtx_("dcrp_simulate");
    if (!result || !result->counts) return fail(DCRP_ERR_INVALID, "dcrp_simulate: result->counts is required");
    if (!e) return fail(DCRP_ERR_INVALID, "engine is NULL");
    // Drone::Simulation calls this once per (aircraft, drone) with the same arrays, a bench once per step:
    // the scenario already in HBM (tables of stage 1 included) is reused when the inputs are unchanged.
#ifdef DCRP_TIMELINE
    g_host_t0 = std::chrono::steady_clock::now();
#endif
    HOST_MARK(0);
    const bool cached = scenario_is_resident(e, tracks, n_tracks, drones, n_drones, model);
    HOST_MARK(1);
    int rc = cached ? DCRP_OK : dcrp_scenario_upload(e, tracks, n_tracks, drones, n_drones, model);
    if (rc != DCRP_OK) return rc;
    rc = dcrp_run_async(e, run, nullptr, nullptr);
    if (rc != DCRP_OK) return rc;
    rc = dcrp_fetch(e, result);
    HOST_MARK(10);
    if (rc != DCRP_OK) return rc;
    if (cached) result->stats.status |= DCRP_STAT_SCENARIO_CACHED;
    else        result->stats.h2d_bytes += e->h2d_scenario;
    return DCRP_OK;
}

// ------------------------------------------------------------------ all-pairs stress
extern "C" int dcrp_allpairs(dcrp_engine* e, const dcrp_allpairs_run* run, dcrp_result* result)
{
    NvtxRange nvtx_("dcrp_allpairs");
    if (!e || !run || !result) return fail(DCRP_ERR_INVALID, "dcrp_allpairs: NULL argument");
    if (!result->counts) return fail(DCRP_ERR_INVALID, "dcrp_allpairs: result->counts is required");
    if (!e->has_scenario) return fail(DCRP_ERR_STATE, "dcrp_allpairs before dcrp_scenario_upload");
    const ScenarioDev& s = e->sd;
    if (run->ref_takeoff_t < 1 || run->ref_takeoff_t > s.t_max)
        return fail(DCRP_ERR_INVALID, "ref_takeoff_t=%d must be in [1, %d] (largest takeoff_t of the scenario)",
                    run->ref_takeoff_t, s.t_max);
    for (int k = 0; k < 3; ++k)
        if (!(run->ref_box[2 * k] <= run->ref_box[2 * k + 1]))
            return fail(DCRP_ERR_INVALID, "ref_box is inverted or NaN");
    const uint64_t n_rays = (uint64_t)s.n_drones * run->ray_count;
    if (n_rays >= (1ull << 32)) return fail(DCRP_ERR_INVALID, "at most 2^32 rays per call");
    CU(cudaSetDevice(e->device));
    (void)cudaGetLastError();      // a stale error left in this thread by other CUDA users (torch, NCCL) is not ours to report
    cudaStream_t st = e->stream;
    const size_t n_pairs = (size_t)s.n_tracks * (size_t)s.n_drones;

    dcrp_run as_run{};                      // what dcrp_fetch reports against
    as_run.seed = run->seed; as_run.trial_begin = run->ray_begin; as_run.trial_count = run->ray_count;
    as_run.precision = DCRP_PRECISION_MIXED; as_run.record_capacity = run->record_capacity;
    e->has_run = false;
    e->run = as_run;
    e->allpairs_run = true;
    e->published = false;
    e->arena_clean = false;
    e->run_timed = true;
    e->run_cull = false;
    e->run_screen3 = false;
    e->run_status = 0;
    e->ap_rays = n_rays;
    e->run_stream_prev = e->run_stream;
    e->run_stream = st;
    e->launches_this_run = 0;
    e->h2d_run = 0;
    e->external_counts = false;
    e->rec_cap = run->record_capacity ? run->record_capacity : DEFAULT_RECORD_CAP;

    CU(size_out_arena(e, n_pairs));
    if (e->run_event_valid && st != e->run_stream_prev) CU(cudaStreamWaitEvent(st, e->ev_end, 0));   // the arenas are shared between runs
    constexpr uint32_t L2_CAP = 1u << 22;
    const size_t hist_bytes = sizeof(uint32_t) * ((size_t)s.n_drones * run->ref_takeoff_t + 1);
    CU(e->d_ap_hist.need(hist_bytes));
    CU(e->d_ap_rank.need(sizeof(uint32_t) * std::max<uint64_t>(1, n_rays)));
    const uint32_t groups = (uint32_t)((run->ray_count + 63) / 64);
    const size_t ray_bytes = sizeof(float) * 4 * 32 * std::max<size_t>(1, (size_t)s.n_drones * groups);
    CU(e->d_ap_rayA.need(ray_bytes));
    CU(e->d_ap_rayB.need(ray_bytes));
    CU(e->d_ap_meta.need(sizeof(uint2) * std::max<uint64_t>(1, n_rays)));
    CU(e->d_ap_l2.need(sizeof(uint2) * (size_t)L2_CAP));

    RunDev r{};
    r.seed = run->seed; r.trial_begin = run->ray_begin; r.trial_count = run->ray_count;
    r.drone_base = run->drone_id_base;
    r.counts = reinterpret_cast<unsigned long long*>(e->d_out.as<unsigned char>() + e->out_counts_off);
    r.first_conflict = reinterpret_cast<int32_t*>(e->d_out.as<unsigned char>() + e->out_first_off);
    r.records = out_records(e);
    r.rec_count = small_rec_count(e);
    r.rec_cap = e->rec_cap;
    r.cand_count = small_cand_count(e);
    r.stats = small_stats(e);

    AllPairsDev ap{};
    ap.seed = run->seed; ap.ray_begin = run->ray_begin; ap.ray_count = run->ray_count;
    ap.drone_base = run->drone_id_base;
    ap.ref_takeoff_t = run->ref_takeoff_t;
    std::memcpy(ap.ref_box, run->ref_box, sizeof ap.ref_box);
    ap.wx32 = (run->ref_box[1] - run->ref_box[0]) * (1.0 / 4294967296.0);
    ap.wy32 = (run->ref_box[3] - run->ref_box[2]) * (1.0 / 4294967296.0);
    ap.wz32 = (float)((run->ref_box[5] - run->ref_box[4]) * (1.0 / 4294967296.0));
    ap.z0rel = (float)(run->ref_box[4] - s.start_alt);
    ap.hist = e->d_ap_hist.as<uint32_t>();
    ap.rank = e->d_ap_rank.as<uint32_t>();
    ap.rayA = e->d_ap_rayA.as<float>();
    ap.rayB = e->d_ap_rayB.as<float>();
    ap.groups = groups;
    ap.meta = e->d_ap_meta.as<uint2>();
    ap.l2 = e->d_ap_l2.as<uint2>();
    ap.l2_count = ap.hist + (size_t)s.n_drones * run->ref_takeoff_t;      // last word of the hist buffer
    ap.l2_cap = L2_CAP;
    constexpr int ap_np = 2;               // packed pairs per lane of k_ap_scan_tracks
    const uint32_t ap_rays = (uint32_t)(2 * ap_np * AP_BLOCK);         // rays per CTA
    ap.blocks_per_drone = (uint32_t)((run->ray_count + ap_rays - 1) / ap_rays);

    CU(cudaEventRecord(e->ev_begin, st));
    k_run_init<<<(unsigned)((e->out_bytes / 4 + 255) / 256), 256, 0, st>>>(
        e->d_out.as<uint32_t>(), (uint32_t)(e->out_first_off / 4), (uint32_t)(e->out_bytes / 4), nullptr, 0u);
    ++e->launches; ++e->launches_this_run;
    CU(cudaGetLastError());
    CU(cudaEventRecord(e->ev_k0, st));
    CU(cudaMemsetAsync(ap.hist, 0, hist_bytes, st));
    if (n_rays) {
        const unsigned rb = (unsigned)((n_rays + 255) / 256);
        k_ap_count<<<rb, 256, 0, st>>>(s, ap);
        k_ap_scan<<<(unsigned)s.n_drones, 32, 0, st>>>(s, ap);
        k_ap_scatter<<<(unsigned)(((uint64_t)s.n_drones * groups * 64 + 255) / 256), 256, 0, st>>>(s, ap);
        e->launches += 3; e->launches_this_run += 3;
        CU(cudaGetLastError());

        const size_t smem = (size_t)AP_STAGES * AP_TILE * 8 + 2 * AP_STAGES * sizeof(uint64_t);
        const void* kern = (const void*)k_ap_scan_tracks<ap_np>;
        int& occ = e->ap_occ;
        if (occ == 0) {
            CU(allow_max_dynamic_smem(e, kern));
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, AP_BLOCK, smem));
            if (occ < 1) occ = 1;
        }
        const uint64_t n_rblocks = (uint64_t)s.n_drones * ap.blocks_per_drone;
        const uint64_t resident = (uint64_t)e->prop.multiProcessorCount * occ;
        // enough items to balance the persistent CTAs: split the tracks when there are few ray blocks
        uint64_t tchunks = std::max<uint64_t>(1, (4 * resident + n_rblocks - 1) / n_rblocks);
        tchunks = std::min<uint64_t>(tchunks, (uint64_t)s.n_tracks);
        ap.tracks_per_chunk = (uint32_t)(((uint64_t)s.n_tracks + tchunks - 1) / tchunks);
        ap.n_tchunks = (uint32_t)(((uint64_t)s.n_tracks + ap.tracks_per_chunk - 1) / ap.tracks_per_chunk);
        const uint64_t n_items = n_rblocks * ap.n_tchunks;
        if (n_items >= (1ull << 32)) return fail(DCRP_ERR_INVALID, "too many all-pairs work items");
        const uint32_t grid = (uint32_t)std::min<uint64_t>(n_items, resident);
        uint32_t n_items32 = (uint32_t)n_items;
        ScenarioDev sd = s;
        void* args[] = {&sd, &r, &ap, &n_items32};
        CU(cudaLaunchKernel(kern, dim3(grid), dim3(AP_BLOCK), args, smem, st));
        ++e->launches; ++e->launches_this_run;
        CU(cudaGetLastError());
    }
    CU(cudaEventRecord(e->ev_trials, st));
    e->has_confirm = true;
    e->k0_is_begin = false;
    if (n_rays) {
        k_ap_resolve<<<e->prop.multiProcessorCount * 4, 128, 0, st>>>(s, r, ap);
        ++e->launches; ++e->launches_this_run;
        CU(cudaGetLastError());
    }
    CU(cudaEventRecord(e->ev_end, st));
    e->run_event_valid = true;
    e->has_run = true;
    uint32_t l2n = 0;
    CU(cudaMemcpyAsync(&l2n, ap.l2_count, sizeof l2n, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (l2n > L2_CAP)       // before anything is merged into the caller's buffers: a retry must not double-count
        return fail(DCRP_ERR_STATE, "level-2 list overflow (%u > %u entries): lower ray_count per call", l2n, L2_CAP);
    return dcrp_fetch(e, result);
}

// ------------------------------------------------------------------ replay / records
extern "C" int dcrp_dump_draws(dcrp_engine* e, int32_t track, int32_t drone, const dcrp_run* run,
                               dcrp_draw* out)
{
    if (!e || !run || !out) return fail(DCRP_ERR_INVALID, "dcrp_dump_draws: NULL argument");
    if (!e->has_scenario) return fail(DCRP_ERR_STATE, "dcrp_dump_draws before dcrp_scenario_upload");
    if (track < 0 || track >= e->sd.n_tracks || drone < 0 || drone >= e->sd.n_drones)
        return fail(DCRP_ERR_INVALID, "track/drone index out of range");
    const uint64_t n = run->trial_count;
    if (n == 0) return DCRP_OK;
    if (n > (1ull << 31)) return fail(DCRP_ERR_INVALID, "at most 2^31 draws per call");
    CU(cudaSetDevice(e->device));
    (void)cudaGetLastError();      // a stale error left in this thread by other CUDA users (torch, NCCL) is not ours to report
    CU(e->d_scratch.need(sizeof(Draw) * (size_t)n));
    k_draws<<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(
        e->sd, track, run->track_id_base + (uint32_t)track, run->drone_id_base + (uint32_t)drone, run->seed,
        run->trial_begin, n, e->d_scratch.as<Draw>());
    ++e->launches;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, e->d_scratch.p, sizeof(Draw) * (size_t)n, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return DCRP_OK;
}

extern "C" int dcrp_trajectories(dcrp_engine* e, const dcrp_record* records, uint32_t n_records, double* xyz,
                                 int32_t* stride_out)
{
    if (!e) return fail(DCRP_ERR_INVALID, "engine is NULL");
    if (!e->has_scenario) return fail(DCRP_ERR_STATE, "dcrp_trajectories before dcrp_scenario_upload");
    if (stride_out) *stride_out = e->sd.n_steps_max;
    if (n_records == 0) return DCRP_OK;
    if (!records || !xyz) return fail(DCRP_ERR_INVALID, "dcrp_trajectories: NULL argument");
    for (uint32_t i = 0; i < n_records; ++i) {
        const dcrp_record& r = records[i];
        if (r.track >= (uint32_t)e->sd.n_tracks || r.drone >= (uint32_t)e->sd.n_drones)
            return fail(DCRP_ERR_INVALID, "record %u: track/drone out of range", i);
        const TrackDev& t = e->h_tracks[r.track];
        if (r.t1st < 1 || r.t1st > t.takeoff_t)
            return fail(DCRP_ERR_INVALID, "record %u: t1st=%d outside [1,%d]", i, r.t1st, t.takeoff_t);
    }
    CU(cudaSetDevice(e->device));
    (void)cudaGetLastError();      // a stale error left in this thread by other CUDA users (torch, NCCL) is not ours to report
    const size_t stride = (size_t)e->sd.n_steps_max;
    const size_t rec_bytes = (sizeof(Record) * (size_t)n_records + 255) & ~size_t(255);
    const size_t xyz_bytes = sizeof(double) * 3 * stride * (size_t)n_records;
    CU(e->d_scratch.need(rec_bytes + xyz_bytes));
    Record* d_rec = e->d_scratch.as<Record>();
    double* d_xyz = reinterpret_cast<double*>(e->d_scratch.as<unsigned char>() + rec_bytes);
    CU(cudaMemcpyAsync(d_rec, records, sizeof(Record) * (size_t)n_records, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemsetAsync(d_xyz, 0, xyz_bytes, e->stream));
    k_trajectory<<<(n_records + 63) / 64, 64, 0, e->stream>>>(e->sd, d_rec, n_records, d_xyz, (int)stride);
    ++e->launches;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(xyz, d_xyz, xyz_bytes, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return DCRP_OK;
}

// ------------------------------------------------------------------ multi-process exchange (comm.cu)
#define NC(call)                                                                                      \
    do {                                                                                              \
        ncclResult_t r__ = (call);                                                                    \
        if (r__ != ncclSuccess)                                                                       \
            return fail(DCRP_ERR_COMM, "%s failed: %s (%s:%d)", #call, nccl_api().GetErrorString(r__), \
                        __FILE__, __LINE__);                                                          \
    } while (0)

extern "C" int dcrp_comm_unique_id(void* out128)
{
    if (!out128) return fail(DCRP_ERR_INVALID, "dcrp_comm_unique_id: out is NULL");
    static_assert(sizeof(ncclUniqueId) == DCRP_UNIQUE_ID_BYTES, "unique id size");
    if (!nccl_api().load()) return fail(DCRP_ERR_COMM, "NCCL is not available: %s", nccl_api().error.c_str());
    ncclUniqueId id;
    NC(nccl_api().GetUniqueId(&id));
    std::memcpy(out128, &id, sizeof id);
    return DCRP_OK;
}

extern "C" int dcrp_engine_detach_comm(dcrp_engine* e)
{
    if (!e) return fail(DCRP_ERR_INVALID, "engine is NULL");
    Comm* c = e->comm;
    if (!c) return DCRP_OK;
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();        // our last sum has seen every peer's last push: nobody writes here any more
    for (int r = 0; r < c->nranks; ++r)
        if (r != c->rank && c->peer[r]) cudaIpcCloseMemHandle(c->peer[r]);
    if (c->mailbox) cudaFree(c->mailbox);
    if (c->h_err) cudaFreeHost(c->h_err);
    if (c->comm) nccl_api().CommDestroy(c->comm);
    delete c;
    e->comm = nullptr;
    return DCRP_OK;
}

extern "C" int dcrp_engine_attach_comm(dcrp_engine* e, const void* unique_id128, int rank, int nranks, uint32_t flags)
{
    NvtxRange nvtx_("dcrp_engine_attach_comm");
    if (!e || !unique_id128) return fail(DCRP_ERR_INVALID, "dcrp_engine_attach_comm: NULL argument");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(DCRP_ERR_INVALID, "rank %d of %d ranks", rank, nranks);
    if (e->comm) return fail(DCRP_ERR_STATE, "a communicator is already attached (dcrp_engine_detach_comm first)");
    if (!nccl_api().load()) return fail(DCRP_ERR_COMM, "NCCL is not available: %s", nccl_api().error.c_str());
    CU(cudaSetDevice(e->device));
    (void)cudaGetLastError();      // a stale error left in this thread by other CUDA users (torch, NCCL) is not ours to report
    Comm* c = new (std::nothrow) Comm();
    if (!c) return fail(DCRP_ERR_INVALID, "out of host memory");
    c->rank = rank; c->nranks = nranks;
    if (nccl_api().GetVersion) nccl_api().GetVersion(&c->nccl_version);
    ncclUniqueId id;
    std::memcpy(&id, unique_id128, sizeof id);
    ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
    cfg.blocking = 1;
    cfg.minCTAs = 1; cfg.maxCTAs = 1;       // a few KB of counters: one CTA, so the kernel can find room beside a trial grid
    ncclResult_t nr = nccl_api().CommInitRankConfig(&c->comm, nranks, id, rank, &cfg);
    if (nr != ncclSuccess) {
        delete c;
        return fail(DCRP_ERR_COMM, "ncclCommInitRankConfig failed: %s", nccl_api().GetErrorString(nr));
    }
    e->comm = c;
    if (cudaHostAlloc(reinterpret_cast<void**>(&c->h_err), sizeof(uint32_t), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer(reinterpret_cast<void**>(&c->d_err), c->h_err, 0) != cudaSuccess) {
        dcrp_engine_detach_comm(e);
        return fail(DCRP_ERR_CUDA, "mapped host word: %s", cudaGetErrorString(cudaGetLastError()));
    }
    *c->h_err = 0;
    c->peer[rank] = nullptr;
    if (nranks == 1 || (flags & DCRP_COMM_NCCL_ONLY) || nranks > COMM_MAX_RANKS) return DCRP_OK;

    // ---- peer-memory mailboxes: allocate, exchange the IPC handles through NCCL, map every peer's
    cudaStream_t st = e->stream;
    int mine_ok = 1;
    unsigned char* d_handles = nullptr;
    std::vector<cudaIpcMemHandle_t> handles((size_t)nranks);
    if (cudaMalloc(reinterpret_cast<void**>(&c->mailbox), comm_mailbox_bytes(nranks)) != cudaSuccess) { c->mailbox = nullptr; mine_ok = 0; }
    if (mine_ok && cudaMemsetAsync(c->mailbox, 0, comm_mailbox_bytes(nranks), st) != cudaSuccess) mine_ok = 0;
    cudaIpcMemHandle_t h{};
    if (mine_ok && cudaIpcGetMemHandle(&h, c->mailbox) != cudaSuccess) mine_ok = 0;
    cudaGetLastError();
    CU(cudaMalloc(reinterpret_cast<void**>(&d_handles), sizeof(cudaIpcMemHandle_t) * (size_t)(nranks + 1) + 16));
    unsigned char* d_mine = d_handles + sizeof(cudaIpcMemHandle_t) * (size_t)nranks;
    CU(cudaMemcpyAsync(d_mine, &h, sizeof h, cudaMemcpyHostToDevice, st));
    NC(nccl_api().AllGather(d_mine, d_handles, sizeof h, ncclChar, c->comm, st));
    CU(cudaMemcpyAsync(handles.data(), d_handles, sizeof h * (size_t)nranks, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    for (int r = 0; r < nranks && mine_ok; ++r) {
        if (r == rank) { c->peer[r] = c->mailbox; continue; }
        void* p = nullptr;
        if (cudaIpcOpenMemHandle(&p, handles[(size_t)r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { mine_ok = 0; cudaGetLastError(); break; }
        c->peer[r] = static_cast<unsigned long long*>(p);
    }
    // every rank must take the same route: all-reduce (min) of "my mailboxes are mapped"; doubles as the
    // barrier after which all mailboxes are zeroed and mapped everywhere
    int* d_ok = reinterpret_cast<int*>(d_handles);
    CU(cudaMemcpyAsync(d_ok, &mine_ok, sizeof(int), cudaMemcpyHostToDevice, st));
    NC(nccl_api().AllReduce(d_ok, d_ok, 1, ncclInt32, ncclMin, c->comm, st));
    int all_ok = 0;
    CU(cudaMemcpyAsync(&all_ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    cudaFree(d_handles);
    c->p2p = all_ok != 0;
    if (!c->p2p) {                              // fall back to ncclAllReduce on every rank
        for (int r = 0; r < nranks; ++r)
            if (r != rank && c->peer[r]) { cudaIpcCloseMemHandle(c->peer[r]); c->peer[r] = nullptr; }
    }
    return DCRP_OK;
}

extern "C" int dcrp_allreduce_counts(dcrp_engine* e, void* stream, uint64_t* d_counts)
{
    NvtxRange nvtx_("dcrp_allreduce_counts");
    if (!e) return fail(DCRP_ERR_INVALID, "engine is NULL");
    Comm* c = e->comm;
    if (!c) return fail(DCRP_ERR_STATE, "dcrp_allreduce_counts before dcrp_engine_attach_comm");
    if (!e->has_scenario) return fail(DCRP_ERR_STATE, "dcrp_allreduce_counts before dcrp_scenario_upload");
    if (!d_counts && !e->has_run) return fail(DCRP_ERR_STATE, "no run whose counts could be reduced");
    CU(cudaSetDevice(e->device));
    (void)cudaGetLastError();      // a stale error left in this thread by other CUDA users (torch, NCCL) is not ours to report
    cudaStream_t st = stream ? (cudaStream_t)stream : (e->has_run ? e->run_stream : e->stream);
    const size_t n_pairs = (size_t)e->sd.n_tracks * (size_t)e->sd.n_drones;
    unsigned long long* counts = d_counts ? reinterpret_cast<unsigned long long*>(d_counts)
        : reinterpret_cast<unsigned long long*>(e->d_out.as<unsigned char>() + e->out_counts_off);
    if (c->nranks == 1) return DCRP_OK;
    if (!d_counts) e->published = false;     // the arena's counts change behind the run: dcrp_fetch copies them from the device
    if (c->p2p && n_pairs <= COMM_MAX_PAIRS) {
        ++c->seq;
        const int parity = (int)(c->seq & 1ull);
        CommPeers peers;
        for (int r = 0; r < COMM_MAX_RANKS; ++r) peers.p[r] = r < c->nranks ? c->peer[r] : nullptr;
        k_comm_push<<<c->nranks - 1, 256, 0, st>>>(peers, c->rank, c->nranks, parity, c->seq, counts, (uint32_t)n_pairs);
        k_comm_sum<<<1, 256, 0, st>>>(c->mailbox, c->rank, c->nranks, parity, c->seq, counts, (uint32_t)n_pairs,
                                       10ull * 1000000000ull, c->d_err);
        e->launches += 2;
        CU(cudaGetLastError());
        ++c->n_p2p;
    } else {
        NC(nccl_api().AllReduce(counts, counts, n_pairs, ncclUint64, ncclSum, c->comm, st));
        ++c->n_nccl;
    }
    return DCRP_OK;
}

extern "C" int dcrp_comm_info(dcrp_engine* e, int32_t info[8])
{
    if (!e || !info) return fail(DCRP_ERR_INVALID, "dcrp_comm_info: NULL argument");
    std::memset(info, 0, sizeof(int32_t) * 8);
    Comm* c = e->comm;
    if (!c) return DCRP_OK;
    info[0] = 1; info[1] = c->rank; info[2] = c->nranks; info[3] = c->p2p ? 1 : 0; info[4] = c->nccl_version;
    info[5] = (int32_t)std::min<uint64_t>(c->n_p2p, 0x7fffffff); info[6] = (int32_t)std::min<uint64_t>(c->n_nccl, 0x7fffffff);
    info[7] = c->h_err ? (int32_t)*c->h_err : 0;
    return DCRP_OK;
}

// ------------------------------------------------------------------ device queries
extern "C" int dcrp_device_count(int* n)
{
    if (!n) return fail(DCRP_ERR_INVALID, "dcrp_device_count: NULL argument");
    *n = 0;
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess || c <= 0) {
        cudaGetLastError();
        return fail(DCRP_ERR_NO_DEVICE, "no CUDA device visible");
    }
    *n = c;
    return DCRP_OK;
}

extern "C" int dcrp_engine_reserve_sms(dcrp_engine* e, int n_sms)
{
    if (!e) return fail(DCRP_ERR_INVALID, "engine is NULL");
    if (n_sms < 0 || n_sms >= e->prop.multiProcessorCount)
        return fail(DCRP_ERR_INVALID, "reserve_sms must be in [0, %d)", e->prop.multiProcessorCount);
    e->reserve_sms = n_sms;
    return DCRP_OK;
}

extern "C" int dcrp_launch_count(dcrp_engine* e, uint64_t* n)
{
    if (!e || !n) return fail(DCRP_ERR_INVALID, "dcrp_launch_count: NULL argument");
    *n = e->launches;
    return DCRP_OK;
}

extern "C" int dcrp_device_info(dcrp_engine* e, char* name, int name_cap, int* sm_count, int* cc_major,
                                int* cc_minor, int* clock_khz)
{
    if (!e) return fail(DCRP_ERR_INVALID, "engine is NULL");
    if (name && name_cap > 0) { std::strncpy(name, e->prop.name, (size_t)name_cap - 1); name[name_cap - 1] = 0; }
    if (sm_count) *sm_count = e->prop.multiProcessorCount;
    if (cc_major) *cc_major = e->prop.major;
    if (cc_minor) *cc_minor = e->prop.minor;
    if (clock_khz) {
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, e->device);
        *clock_khz = khz;
    }
    return DCRP_OK;
}
