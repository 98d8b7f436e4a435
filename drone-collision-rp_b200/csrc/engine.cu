// engine.cu -- the extern "C" launcher ABI of include/dcrp.h.
//
// Host logic only: argument checks, the HBM layout of a scenario, launches on
// CUDA streams, result gathering.  There is deliberately NO CPU implementation
// of the trial path in this library: without a usable sm_100 device
// dcrp_engine_create fails and nothing else can be called.
#include "../../include/dcrp.h"
#include "dcrp_device.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <climits>
#include <string>
#include <vector>

#include "kernels.cu"

using namespace dcrp;

static_assert(sizeof(dcrp_draw) == sizeof(Draw) && sizeof(Draw) == 32, "draw layout");
static_assert(sizeof(dcrp_record) == sizeof(Record) && sizeof(Record) == 56, "record layout");

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

// screen kernel configuration (512 threads x 4 trials, 2 CTAs / SM)
constexpr int SCREEN_K = 4;
constexpr int SCREEN_BLOCK = 512;
constexpr int SCREEN_CH = SCREEN_K * SCREEN_BLOCK;
constexpr int SCREEN_TILE_STEPS = 512;       // 16 KB per tile buffer
constexpr uint32_t DEFAULT_RECORD_CAP = 1u << 16;
constexpr uint32_t CAND_CAP = 1u << 22;
constexpr uint64_t MAX_ITEMS_PER_LAUNCH = 1ull << 30;

struct dcrp_engine {
    int device = 0;
    cudaDeviceProp prop{};
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_begin = nullptr, ev_trials = nullptr, ev_end = nullptr, ev_a = nullptr, ev_b = nullptr;
    uint64_t launches = 0;

    // scenario
    bool has_scenario = false;
    ScenarioDev sd{};
    std::vector<TrackDev> h_tracks;
    dcrp_model model{};
    float ms_prepare = 0.f;
    uint64_t h2d_scenario = 0;
    DevBuf d_tracks, d_ax, d_ay, d_az, d_trk32, d_drones, d_p1x, d_p1y, d_s1fh, d_s1pm;

    // run
    bool has_run = false;
    dcrp_run run{};
    cudaStream_t run_stream = nullptr;
    bool external_counts = false;
    uint32_t rec_cap = 0;
    uint32_t launches_this_run = 0;
    uint64_t h2d_run = 0;
    DevBuf d_counts, d_first, d_records, d_small /* rec_count, cand_count, stats */, d_cands, d_replay;
    DevBuf d_trial_hit, d_trial_first, d_trial_min;
    DevBuf d_scratch;    // peak kernels / draws / trajectories
    DevBuf d_flush;      // 256 MB written by dcrp_flush_l2 (> 126 MB of L2)
};

static uint32_t* small_rec_count(dcrp_engine* e) { return e->d_small.as<uint32_t>(); }
static uint32_t* small_cand_count(dcrp_engine* e) { return e->d_small.as<uint32_t>() + 1; }
static unsigned long long* small_stats(dcrp_engine* e)
{
    return reinterpret_cast<unsigned long long*>(e->d_small.as<unsigned char>() + 16);
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
        cudaEventCreate(&e->ev_begin) != cudaSuccess || cudaEventCreate(&e->ev_trials) != cudaSuccess ||
        cudaEventCreate(&e->ev_end) != cudaSuccess || cudaEventCreate(&e->ev_a) != cudaSuccess ||
        cudaEventCreate(&e->ev_b) != cudaSuccess) {
        delete e;
        return fail(DCRP_ERR_CUDA, "stream/event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    *out = e;
    return DCRP_OK;
}

extern "C" int dcrp_engine_destroy(dcrp_engine* e)
{
    if (!e) return DCRP_OK;
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    DevBuf* bufs[] = {&e->d_tracks, &e->d_ax, &e->d_ay, &e->d_az, &e->d_trk32, &e->d_drones, &e->d_p1x,
                      &e->d_p1y, &e->d_s1fh, &e->d_s1pm, &e->d_counts, &e->d_first, &e->d_records,
                      &e->d_small, &e->d_cands, &e->d_replay, &e->d_trial_hit, &e->d_trial_first,
                      &e->d_trial_min, &e->d_scratch, &e->d_flush};
    for (DevBuf* b : bufs) b->release();
    cudaEventDestroy(e->ev_begin); cudaEventDestroy(e->ev_trials); cudaEventDestroy(e->ev_end);
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
    if (!e) return fail(DCRP_ERR_INVALID, "engine is NULL");
    if (!tracks || n_tracks <= 0 || !drones || n_drones <= 0 || !model)
        return fail(DCRP_ERR_INVALID, "dcrp_scenario_upload: tracks/drones/model missing or empty");
    if ((uint64_t)n_tracks * (uint64_t)n_drones > (1ull << 31) - 1)
        return fail(DCRP_ERR_INVALID, "too many (track, drone) pairs");
    if (!(model->v_straight > 0) || !(model->aircraft_radius + model->drone_radius > 0))
        return fail(DCRP_ERR_INVALID, "model: v_straight and the radii must be positive");
    CU(cudaSetDevice(e->device));
    e->has_scenario = false;
    e->has_run = false;

    // ---- validate, offsets, extents
    e->h_tracks.assign((size_t)n_tracks, TrackDev{});
    uint64_t total = 0;
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
        d.n_steps = t.n_steps; d.takeoff_t = t.takeoff_t; d.off = (uint32_t)total; d.pad = 0;
        std::memcpy(d.box, t.box, sizeof d.box);
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
    s.ox = std::floor(0.5 * (lo[0] + hi[0]));
    s.oy = std::floor(0.5 * (lo[1] + hi[1]));
    s.oz = std::floor(0.5 * (lo[2] + hi[2]));
    const double extent = std::max({hi[0] - s.ox, s.ox - lo[0], hi[1] - s.oy, s.oy - lo[1],
                                    hi[2] - s.oz, s.oz - lo[2]}) + 1.0;
    s.start_alt = model->start_alt; s.v_straight = model->v_straight; s.v_ascend = model->v_ascend;
    s.radius = model->drone_radius + model->aircraft_radius;          // Drone.cpp:260
    {   // sqrt_rn(d2) < R  <=>  d2 < c, c = smallest double whose rounded root reaches R
        double c = s.radius * s.radius;
        while (std::sqrt(c) >= s.radius) c = std::nextafter(c, 0.0);
        while (std::sqrt(c) < s.radius) c = std::nextafter(c, INFINITY);
        s.thr_d2 = c;
    }
    // fp32 screen: per-coordinate error <= ~2 ulp32(extent) + i*|s|*2^-24; the
    // margin below is >= 8x that bound.
    const double margin = 0.02 + 32.0 * (double)ulp32_of(extent) +
                          (double)vl_max * model->v_straight * std::ldexp(1.0, -22);
    s.thr2_screen = std::nextafterf((float)((s.radius + margin) * (s.radius + margin)), INFINITY);
    s.thr2_fp32 = (float)(s.radius * s.radius);
    s.v_straight_f = (float)model->v_straight;
    s.coef_f = (float)(2.0 * (model->v_straight - model->v_ascend) / DCRP_PI);
    s.bucket_shift = 0;
    while (((t_max - 1) >> s.bucket_shift) >= 64) ++s.bucket_shift;

    // ---- host staging
    std::vector<double> hx(total), hy(total), hz(total);
    std::vector<float4> h32(2 * total);
    for (int a = 0; a < n_tracks; ++a) {
        const dcrp_track& t = tracks[a];
        const size_t off = e->h_tracks[(size_t)a].off;
        for (int i = 0; i < t.n_steps; ++i) {
            hx[off + i] = t.x[i]; hy[off + i] = t.y[i]; hz[off + i] = t.z[i];
            const float fx = -(float)(t.x[i] - s.ox), fy = -(float)(t.y[i] - s.oy),
                        fz = -(float)(t.z[i] - s.oz);
            h32[2 * (off + i)] = make_float4(fx, fx, fy, fy);
            h32[2 * (off + i) + 1] = make_float4(fz, fz, (float)i, (float)i);
        }
    }
    std::vector<double4> hd((size_t)n_drones);
    for (int j = 0; j < n_drones; ++j) {
        // Drone.cpp:121-122: sin(heading)*max_straight_speed, formed with the host libm
        // (the one the reference links) so stage 1 is bit-identical to the reference.
        const double sx = std::sin(drones[j].heading0) * model->v_straight;
        const double sy = std::cos(drones[j].heading0) * model->v_straight;
        hd[(size_t)j] = make_double4(drones[j].x0, drones[j].y0, sx, sy);
    }

    // ---- device buffers + upload
    const size_t n_pairs = (size_t)n_tracks * (size_t)n_drones;
    CU(e->d_tracks.need(sizeof(TrackDev) * (size_t)n_tracks));
    CU(e->d_ax.need(sizeof(double) * total));
    CU(e->d_ay.need(sizeof(double) * total));
    CU(e->d_az.need(sizeof(double) * total));
    CU(e->d_trk32.need(sizeof(float4) * 2 * total));
    CU(e->d_drones.need(sizeof(double4) * (size_t)n_drones));
    CU(e->d_p1x.need(sizeof(double) * (size_t)n_drones * (size_t)t_max));
    CU(e->d_p1y.need(sizeof(double) * (size_t)n_drones * (size_t)t_max));
    CU(e->d_s1fh.need(sizeof(int32_t) * n_pairs));
    CU(e->d_s1pm.need(sizeof(double) * n_pairs * (size_t)t_max));
    cudaStream_t st = e->stream;
    CU(cudaEventRecord(e->ev_a, st));
    CU(cudaMemcpyAsync(e->d_tracks.p, e->h_tracks.data(), sizeof(TrackDev) * (size_t)n_tracks, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(e->d_ax.p, hx.data(), sizeof(double) * total, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(e->d_ay.p, hy.data(), sizeof(double) * total, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(e->d_az.p, hz.data(), sizeof(double) * total, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(e->d_trk32.p, h32.data(), sizeof(float4) * 2 * total, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(e->d_drones.p, hd.data(), sizeof(double4) * (size_t)n_drones, cudaMemcpyHostToDevice, st));
    e->h2d_scenario = sizeof(TrackDev) * (size_t)n_tracks + sizeof(double) * 3 * total +
                      sizeof(float4) * 2 * total + sizeof(double4) * (size_t)n_drones;

    s.tracks = e->d_tracks.as<TrackDev>();
    s.ax = e->d_ax.as<double>(); s.ay = e->d_ay.as<double>(); s.az = e->d_az.as<double>();
    s.trk32 = e->d_trk32.as<float4>();
    s.drones = e->d_drones.as<double4>();
    s.p1x = e->d_p1x.as<double>(); s.p1y = e->d_p1y.as<double>();
    s.s1_first_hit = e->d_s1fh.as<int32_t>();
    s.s1_prefmin = e->d_s1pm.as<double>();

    k_stage1<<<(unsigned)((n_pairs + 127) / 128), 128, 0, st>>>(s);
    ++e->launches;
    CU(cudaGetLastError());
    CU(cudaEventRecord(e->ev_b, st));
    CU(cudaStreamSynchronize(st));      // staging vectors go out of scope
    CU(cudaEventElapsedTime(&e->ms_prepare, e->ev_a, e->ev_b));
    e->sd = s;
    e->model = *model;
    e->has_scenario = true;
    return DCRP_OK;
}

// ------------------------------------------------------------------ run
template <bool REPLAY>
static cudaError_t launch_screen(dcrp_engine* e, const RunDev& r, uint32_t chunks_per_pair, uint32_t n_items,
                                 cudaStream_t st)
{
    auto kern = k_screen<SCREEN_K, SCREEN_BLOCK, REPLAY>;
    const size_t smem = (size_t)2 * SCREEN_TILE_STEPS * 32 + (size_t)SCREEN_CH * 16 + 80 * 4 + 16;
    static bool configured[2] = {false, false};
    if (!configured[REPLAY ? 1 : 0]) {
        cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (ce != cudaSuccess) return ce;
        configured[REPLAY ? 1 : 0] = true;
    }
    int occ = 0;
    cudaError_t ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, SCREEN_BLOCK, smem);
    if (ce != cudaSuccess) return ce;
    if (occ < 1) occ = 1;
    const uint32_t grid = (uint32_t)std::min<uint64_t>(n_items, (uint64_t)e->prop.multiProcessorCount * occ);
    kern<<<grid, SCREEN_BLOCK, smem, st>>>(e->sd, r, chunks_per_pair, n_items, SCREEN_TILE_STEPS);
    ++e->launches; ++e->launches_this_run;
    return cudaGetLastError();
}

extern "C" int dcrp_run_async(dcrp_engine* e, const dcrp_run* run, void* stream, uint64_t* d_counts)
{
    if (!e || !run) return fail(DCRP_ERR_INVALID, "dcrp_run_async: NULL argument");
    if (!e->has_scenario) return fail(DCRP_ERR_STATE, "dcrp_run_async before dcrp_scenario_upload");
    if (run->precision < DCRP_PRECISION_FP64 || run->precision > DCRP_PRECISION_FP32)
        return fail(DCRP_ERR_INVALID, "unknown precision %d", run->precision);
    if ((run->flags & DCRP_FLAG_PER_TRIAL) && run->precision != DCRP_PRECISION_FP64)
        return fail(DCRP_ERR_INVALID, "DCRP_FLAG_PER_TRIAL needs DCRP_PRECISION_FP64");
    if (run->trial_begin + run->trial_count < run->trial_begin)
        return fail(DCRP_ERR_INVALID, "trial range overflows 64 bits");
    CU(cudaSetDevice(e->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : e->stream;
    const ScenarioDev& s = e->sd;
    const size_t n_pairs = (size_t)s.n_tracks * (size_t)s.n_drones;

    e->run = *run;
    e->run_stream = st;
    e->launches_this_run = 0;
    e->h2d_run = 0;
    e->rec_cap = run->record_capacity ? run->record_capacity : DEFAULT_RECORD_CAP;
    e->external_counts = (d_counts != nullptr);

    CU(e->d_counts.need(sizeof(uint64_t) * n_pairs));
    CU(e->d_first.need(sizeof(int32_t) * n_pairs));
    CU(e->d_records.need(sizeof(Record) * (size_t)e->rec_cap));
    CU(e->d_small.need(16 + sizeof(unsigned long long) * ST_COUNT));
    CU(e->d_cands.need(sizeof(Candidate) * (size_t)CAND_CAP));

    RunDev r{};
    r.seed = run->seed; r.trial_begin = run->trial_begin; r.trial_count = run->trial_count;
    r.track_base = run->track_id_base; r.drone_base = run->drone_id_base;
    r.replay = nullptr;
    if (run->replay && run->trial_count) {
        const size_t bytes = sizeof(Draw) * n_pairs * (size_t)run->trial_count;
        CU(e->d_replay.need(bytes));
        CU(cudaMemcpyAsync(e->d_replay.p, run->replay, bytes, cudaMemcpyHostToDevice, st));
        e->h2d_run += bytes;
        r.replay = e->d_replay.as<Draw>();
    }
    r.counts = d_counts ? reinterpret_cast<unsigned long long*>(d_counts) : e->d_counts.as<unsigned long long>();
    r.first_conflict = e->d_first.as<int32_t>();
    r.records = e->d_records.as<Record>();
    r.rec_count = small_rec_count(e);
    r.rec_cap = e->rec_cap;
    r.cands = e->d_cands.as<Candidate>();
    r.cand_count = small_cand_count(e);
    r.cand_cap = CAND_CAP;
    r.stats = small_stats(e);
    r.decide_fp32 = (run->precision == DCRP_PRECISION_FP32) ? 1 : 0;
    if (run->flags & DCRP_FLAG_PER_TRIAL) {
        const size_t n = n_pairs * (size_t)run->trial_count;
        CU(e->d_trial_hit.need(n ? n : 1));
        CU(e->d_trial_first.need(sizeof(int32_t) * (n ? n : 1)));
        CU(e->d_trial_min.need(sizeof(double) * (n ? n : 1)));
        r.trial_hit = e->d_trial_hit.as<uint8_t>();
        r.trial_first = e->d_trial_first.as<int32_t>();
        r.trial_min = e->d_trial_min.as<double>();
    }

    CU(cudaEventRecord(e->ev_begin, st));
    if (!d_counts) CU(cudaMemsetAsync(e->d_counts.p, 0, sizeof(uint64_t) * n_pairs, st));
    CU(cudaMemsetAsync(e->d_first.p, 0x7f, sizeof(int32_t) * n_pairs, st));   // 0x7f7f7f7f: "none" marker
    CU(cudaMemsetAsync(e->d_small.p, 0, 16 + sizeof(unsigned long long) * ST_COUNT, st));

    // slices keep the item count of one launch below 2^30
    const bool exact = (run->precision == DCRP_PRECISION_FP64);
    const uint64_t per_item = exact ? (uint64_t)EXB : (uint64_t)SCREEN_CH;
    const uint64_t max_chunks = std::max<uint64_t>(1, MAX_ITEMS_PER_LAUNCH / n_pairs);
    const uint64_t slice_max = max_chunks * per_item;
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
            cudaError_t ce = r.replay ? launch_screen<true>(e, r, chunks, n_items, st)
                                      : launch_screen<false>(e, r, chunks, n_items, st);
            if (ce != cudaSuccess) return fail(DCRP_ERR_CUDA, "k_screen launch: %s", cudaGetErrorString(ce));
        }
    }
    CU(cudaEventRecord(e->ev_trials, st));
    if (run->precision == DCRP_PRECISION_MIXED && run->trial_count) {
        // grid-stride over the candidate list whose length lives on the device
        k_confirm_all<<<e->prop.multiProcessorCount * 4, 128, 0, st>>>(s, r);
        ++e->launches; ++e->launches_this_run;
        CU(cudaGetLastError());
    }
    CU(cudaEventRecord(e->ev_end, st));
    e->has_run = true;
    return DCRP_OK;
}

extern "C" int dcrp_last_run_ms(dcrp_engine* e, float* ms)
{
    if (!e || !ms) return fail(DCRP_ERR_INVALID, "dcrp_last_run_ms: NULL argument");
    if (!e->has_run) return fail(DCRP_ERR_STATE, "no run to time");
    CU(cudaSetDevice(e->device));
    CU(cudaEventSynchronize(e->ev_end));
    CU(cudaEventElapsedTime(ms, e->ev_begin, e->ev_end));
    return DCRP_OK;
}

extern "C" int dcrp_fetch(dcrp_engine* e, dcrp_result* res)
{
    if (!e || !res) return fail(DCRP_ERR_INVALID, "dcrp_fetch: NULL argument");
    if (!e->has_run) return fail(DCRP_ERR_STATE, "dcrp_fetch before dcrp_run_async");
    CU(cudaSetDevice(e->device));
    cudaStream_t st = e->run_stream;
    const ScenarioDev& s = e->sd;
    const size_t n_pairs = (size_t)s.n_tracks * (size_t)s.n_drones;
    CU(cudaStreamSynchronize(st));

    unsigned char small[16 + sizeof(unsigned long long) * ST_COUNT];
    CU(cudaMemcpy(small, e->d_small.p, sizeof small, cudaMemcpyDeviceToHost));
    uint32_t rec_count, cand_count;
    unsigned long long stats[ST_COUNT];
    std::memcpy(&rec_count, small, 4);
    std::memcpy(&cand_count, small + 4, 4);
    std::memcpy(stats, small + 16, sizeof stats);
    uint64_t d2h = sizeof small;

    if (e->run.precision == DCRP_PRECISION_MIXED && cand_count > CAND_CAP) {
        // More candidates than the list holds (absurd hit rates, e.g. a huge radius):
        // redo the whole run with the fp64 kernel -- still on the GPU, same answers.
        if (e->external_counts)
            return fail(DCRP_ERR_STATE, "candidate list overflow with caller-owned counts: rerun with DCRP_PRECISION_FP64");
        dcrp_run again = e->run;
        again.precision = DCRP_PRECISION_FP64;
        int rc = dcrp_run_async(e, &again, (void*)st, nullptr);
        if (rc != DCRP_OK) return rc;
        e->run.precision = DCRP_PRECISION_MIXED;   // report what was asked for
        return dcrp_fetch(e, res);
    }

    if (res->counts && !e->external_counts) {
        std::vector<uint64_t> h(n_pairs);
        CU(cudaMemcpy(h.data(), e->d_counts.p, sizeof(uint64_t) * n_pairs, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < n_pairs; ++i) res->counts[i] += h[i];     // accumulate, Drone.cpp:276
        d2h += sizeof(uint64_t) * n_pairs;
    }
    if (res->first_conflict) {
        CU(cudaMemcpy(res->first_conflict, e->d_first.p, sizeof(int32_t) * n_pairs, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < n_pairs; ++i)
            if (res->first_conflict[i] == 0x7f7f7f7f) res->first_conflict[i] = INT32_MAX;
        d2h += sizeof(int32_t) * n_pairs;
    }
    res->n_records = 0;
    res->overflow = rec_count > e->rec_cap ? 1 : 0;
    if (res->records) {
        const uint32_t n = std::min(rec_count, e->rec_cap);
        if (n) {
            CU(cudaMemcpy(res->records, e->d_records.p, sizeof(Record) * (size_t)n, cudaMemcpyDeviceToHost));
            d2h += sizeof(Record) * (size_t)n;
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
    o.trials = (uint64_t)n_pairs * e->run.trial_count;
    o.tests_evaluated = stats[ST_EVALUATED];
    o.tests_shared = stats[ST_SHARED];
    o.tests_issued = stats[ST_ISSUED];
    o.drone_steps = stats[ST_STEPS];
    o.candidates = stats[ST_CANDIDATES];
    o.hits = stats[ST_HITS];
    o.kernel_launches = e->launches_this_run;
    o.ms_prepare = e->ms_prepare;
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, e->ev_begin, e->ev_trials)); o.ms_trials = ms;
    CU(cudaEventElapsedTime(&ms, e->ev_trials, e->ev_end));   o.ms_confirm = ms;
    CU(cudaEventElapsedTime(&ms, e->ev_begin, e->ev_end));    o.ms_device_total = ms;
    o.h2d_bytes = e->h2d_run;
    o.d2h_bytes = d2h;
    return DCRP_OK;
}

extern "C" int dcrp_simulate(dcrp_engine* e, const dcrp_track* tracks, int32_t n_tracks,
                             const dcrp_drone* drones, int32_t n_drones, const dcrp_model* model,
                             const dcrp_run* run, dcrp_result* result)
{
    if (!result || !result->counts) return fail(DCRP_ERR_INVALID, "dcrp_simulate: result->counts is required");
    int rc = dcrp_scenario_upload(e, tracks, n_tracks, drones, n_drones, model);
    if (rc != DCRP_OK) return rc;
    rc = dcrp_run_async(e, run, nullptr, nullptr);
    if (rc != DCRP_OK) return rc;
    rc = dcrp_fetch(e, result);
    if (rc != DCRP_OK) return rc;
    result->stats.h2d_bytes += e->h2d_scenario;
    return DCRP_OK;
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

// ------------------------------------------------------------------ measurement
extern "C" int dcrp_measure_peak(dcrp_engine* e, int which, double* tflops, float* ms_out)
{
    if (!e || !tflops) return fail(DCRP_ERR_INVALID, "dcrp_measure_peak: NULL argument");
    if (which < 0 || which > 3) return fail(DCRP_ERR_INVALID, "which must be 0..3");
    CU(cudaSetDevice(e->device));
    const int threads = 256, blocks = e->prop.multiProcessorCount * 8;
    const int iters = 2048;
    CU(e->d_scratch.need(sizeof(double) * (size_t)threads * blocks));
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CU(cudaEventRecord(e->ev_a, e->stream));
        switch (which) {
            case 0: k_peak_ffma<<<blocks, threads, 0, e->stream>>>(e->d_scratch.as<float>(), iters, 0.999f, 0.001f); break;
            case 1: k_peak_ffma2<<<blocks, threads, 0, e->stream>>>(e->d_scratch.as<float>(), iters, 0.999f, 0.001f); break;
            case 2: k_peak_dfma<<<blocks, threads, 0, e->stream>>>(e->d_scratch.as<double>(), iters, 0.999, 0.001); break;
            default: k_peak_scanmix<<<blocks, threads, 0, e->stream>>>(e->d_scratch.as<float>(), iters, 0.5f, -3.0f); break;
        }
        ++e->launches;
        CU(cudaGetLastError());
        CU(cudaEventRecord(e->ev_b, e->stream));
        CU(cudaEventSynchronize(e->ev_b));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, e->ev_a, e->ev_b));
        if (rep > 0) best = std::min(best, ms);
    }
    const double lanes = (double)threads * blocks;
    double flops;
    if (which == 0)      flops = lanes * iters * 16.0 * 8.0 * 2.0;
    else if (which == 1) flops = lanes * iters * 16.0 * 8.0 * 4.0;
    else if (which == 2) flops = lanes * iters * 16.0 * 8.0 * 2.0;
    else                 flops = lanes * iters * 8.0 * 4.0 * 2.0 * 11.0;   // 2 tests per packed pair, 11 flop each
    *tflops = flops / ((double)best * 1e-3) * 1e-12;
    if (ms_out) *ms_out = best;
    return DCRP_OK;
}

extern "C" int dcrp_flush_l2(dcrp_engine* e, void* stream)
{
    if (!e) return fail(DCRP_ERR_INVALID, "engine is NULL");
    CU(cudaSetDevice(e->device));
    DevBuf& flush = e->d_flush;
    const size_t bytes = (size_t)256 << 20;
    CU(flush.need(bytes));
    cudaStream_t st = stream ? (cudaStream_t)stream : e->stream;
    k_fill<<<e->prop.multiProcessorCount * 8, 256, 0, st>>>(flush.as<float4>(), bytes / sizeof(float4), 1.0f);
    ++e->launches;
    CU(cudaGetLastError());
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
