// kernels.cu -- hand-written sm_100a kernels of the Monte-Carlo trial path.
//
//   k_stage1      once per scenario: stage-1 paths + per-(track,drone) stage-1 tables (one warp per pair)
//   k_reach       once per scenario, on demand: feasible index ranges for DCRP_FLAG_REACH_CULL
//   k_screen2     every trial, fp32 packed (FFMA2/FADD2/FMUL2/FMNMX3): Philox draws,
//                 CTA-level binning by t1st, fp64->fp32 ray set-up, scan of the track
//                 staged in shared memory by TMA bulk copies, warp-aggregated append
//                 of the (rare) candidates
//   k_confirm_all candidates only, fp64 in the reference's order of operations (one warp per candidate)
//   k_ap_*        all-pairs stress path (BASELINE config 5): rays sorted once, tracks streamed by a TMA ring
//   k_exact       every trial in fp64 (DCRP_PRECISION_FP64; the parity kernel)
//   k_draws       Philox replay-table dump
//   k_trajectory  full fp64 trajectories of recorded collisions (Drone::Output rows)
//
// Reference semantics: /root/reference/Drone.cpp:66-80,107-127,164-231,253-281
// (restated in dcrp_math.cuh).  Nothing here is translated from the reference:
// it has no device code at all.
#include "dcrp_device.cuh"

#include <limits.h>

namespace dcrp {

// =========================================================================
// small device helpers
// =========================================================================
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// Warp-aggregated atomic slot claim; call from inside the (divergent) branch of
// the lanes that need a slot.  One atomicAdd per group of converged lanes.
__device__ __forceinline__ uint32_t warp_claim_slot(uint32_t* counter)
{
    const unsigned active = __activemask();
    const int leader = __ffs(active) - 1;
    const unsigned lt = (1u << lane_id()) - 1u;
    uint32_t base = 0;
    if ((int)lane_id() == leader) base = atomicAdd(counter, (uint32_t)__popc(active));
    base = __shfl_sync(active, base, leader);
    return base + (uint32_t)__popc(active & lt);
}

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- packed fp32x2 arithmetic (sm_100: FFMA2 / FADD2 / FMUL2) and 3-input min (FMNMX3)
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c)
{
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a);
    unsigned long long rb = *reinterpret_cast<unsigned long long*>(&b);
    unsigned long long rc = *reinterpret_cast<unsigned long long*>(&c);
    unsigned long long rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b)
{
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a);
    unsigned long long rb = *reinterpret_cast<unsigned long long*>(&b);
    unsigned long long rd;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b)
{
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a);
    unsigned long long rb = *reinterpret_cast<unsigned long long*>(&b);
    unsigned long long rd;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2*>(&rd);
}
// The same on opaque 64-bit handles: a value that lives in ONE aligned register pair for its whole
// life.  (Assembling a float2 from two unrelated scalars lets ptxas re-pack it with MOVs inside the
// hot loop; a mov.b64 pack done once outside does not.)
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t pk2(float lo, float hi)
{
    f2_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float2 up2(f2_t v)
{
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ f2_t ffma2(f2_t a, f2_t b, f2_t c)
{
    f2_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f2_t fadd2(f2_t a, f2_t b)
{
    f2_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f2_t fmul2(f2_t a, f2_t b)
{
    f2_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float fmin3(float a, float b, float c)
{
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// ---- mbarrier + 1-D TMA bulk copy (global -> shared), SASS: UBLKCP / SYNCS
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}

// Level-1 scan of one tile for NP packed pairs.  `txy` points at the tile's first sample (an EVEN track
// index t0), [lo, hi) are absolute indices inside the tile.  Incremental form: D = P + (-A_i), P += S,
// so every packed op reads at most two distinct register pairs and FFMA2/FADD2/FMUL2 issue back to back
// every 2 cycles (the fused P = i*S + B form reads three and measured 12 % slower).  One broadcast
// LDS.128 feeds two indices; the {v, v} pairs cost nothing (ptxas folds them into the packed ops'
// scalar-broadcast operand form).
template <int NP>
__device__ __forceinline__ void scan_tile_xy(const float2* __restrict__ txy, int t0, int lo, int hi,
                                             f2_t (&PX)[NP], f2_t (&PY)[NP], const f2_t (&SX)[NP],
                                             const f2_t (&SY)[NP], float (&mn)[2 * NP])
{
#define DCRP_STEP(vx, vy)                                                     \
    {                                                                         \
        const f2_t cx_ = pk2((vx), (vx)), cy_ = pk2((vy), (vy));              \
        _Pragma("unroll") for (int p_ = 0; p_ < NP; ++p_) {                   \
            const f2_t dx_ = fadd2(PX[p_], cx_);                              \
            const f2_t dy_ = fadd2(PY[p_], cy_);                              \
            const float2 d2_ = up2(ffma2(dy_, dy_, fmul2(dx_, dx_)));         \
            mn[2 * p_] = fminf(mn[2 * p_], d2_.x);                            \
            mn[2 * p_ + 1] = fminf(mn[2 * p_ + 1], d2_.y);                    \
            PX[p_] = fadd2(PX[p_], SX[p_]);                                   \
            PY[p_] = fadd2(PY[p_], SY[p_]);                                   \
        }                                                                     \
    }
    int i = lo;
    if (i & 1) {                                   // odd start: one single step to reach an even index
        const float2 c = txy[i - t0];
        DCRP_STEP(c.x, c.y)
        ++i;
    }
    const int npair = (hi - i) >> 1;
    const float4* __restrict__ q4 = reinterpret_cast<const float4*>(txy + (i - t0));
#pragma unroll 4
    for (int k = 0; k < npair; ++k) {
        const float4 c = q4[k];
        DCRP_STEP(c.x, c.y)
        DCRP_STEP(c.z, c.w)
    }
    i += 2 * npair;
    if (i < hi) {
        const float2 c = txy[i - t0];
        DCRP_STEP(c.x, c.y)
    }
#undef DCRP_STEP
}

// =========================================================================
// stage 1: once per scenario
// =========================================================================
// One WARP per (track, drone), one lane per stage-1 index (32 indices per pass).  The drone's stage-1
// path is the reference's running sum (Drone.cpp:120-126: x[i] = x[i-1] + sin(h0)*v, two roundings per
// step, z = start altitude), so lane l re-runs the l additions from the pass's first position -- the same
// additions in the same order, hence the same bits -- instead of waiting for its neighbour.  Each index
// is tested against the track sample of the same index (Drone.cpp:253-262); first hit and running
// minimum come from a ballot and a warp prefix-min (min is order-independent).
__device__ __forceinline__ double warp_prefix_min(double v, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v = fmin(v, u);
    }
    return v;
}

__global__ void __launch_bounds__(128) k_stage1(const ScenarioDev s)
{
    const int pair = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (pair >= s.n_tracks * s.n_drones) return;
    const int a = pair / s.n_drones, j = pair - a * s.n_drones;
    const TrackDev t = s.tracks[a];
    const double4 d = s.drones[j];
    double bx = d.x, by = d.y;                   // position at the first index of the pass
    double mn_carry = INFINITY, smn_carry = INFINITY;
    int first = INT_MAX, sfirst = INT_MAX;
    double stime = -1.0;
    double qx = 0, qy = 0, qz = 0;               // relative position at the sample before the pass
    for (int c0 = 0; c0 < s.t_max; c0 += 32) {
        const int i = c0 + lane;
        double x = bx, y = by;
        for (int k = 0; k < lane; ++k) { x = DADD(x, d.z); y = DADD(y, d.w); }
        bx = __shfl_sync(0xffffffffu, DADD(x, d.z), 31);
        by = __shfl_sync(0xffffffffu, DADD(y, d.w), 31);
        const bool in_t = i < s.t_max;
        const bool on = in_t && i < t.n_steps;   // all i < t_max: all-pairs rays may kink later than this track's takeoff_t
        double cx = 0, cy = 0, cz = 0, d2 = INFINITY;
        if (on) {
            const double ax = s.ax[t.off + i], ay = s.ay[t.off + i], az = s.az[t.off + i];
            cx = DSUB(x, ax); cy = DSUB(y, ay); cz = DSUB(s.start_alt, az);
            d2 = sep2(x, y, s.start_alt, ax, ay, az);
        }
        const uint32_t hm = __ballot_sync(0xffffffffu, on && d2 < s.thr_d2);
        if (first == INT_MAX && hm) first = c0 + __ffs((int)hm) - 1;
        const double pm = fmin(warp_prefix_min(d2, lane), mn_carry);
        mn_carry = __shfl_sync(0xffffffffu, pm, 31);

        // sub-step extension: segment i-1 -> i belongs to lane i; s1sub_prefmin[i] covers every segment
        // that ends at a sample <= i, i.e. all the segments a trial kinking at sample i flies in stage 1
        double px = __shfl_up_sync(0xffffffffu, cx, 1), py = __shfl_up_sync(0xffffffffu, cy, 1),
               pz = __shfl_up_sync(0xffffffffu, cz, 1);
        if (lane == 0) { px = qx; py = qy; pz = qz; }
        double sd2 = INFINITY, tc = -1.0;
        if (on && i > 0) {
            const SegOut sg = seg_closest(px, py, pz, cx, cy, cz, s.radius2);
            sd2 = sg.d2; tc = sg.tc;
        }
        const uint32_t sm = __ballot_sync(0xffffffffu, tc >= 0.0);
        if (sfirst == INT_MAX && sm) {
            const int src = __ffs((int)sm) - 1;
            sfirst = c0 + src - 1;
            stime = DADD((double)sfirst, __shfl_sync(0xffffffffu, tc, src));
        }
        const double spm = fmin(warp_prefix_min(sd2, lane), smn_carry);
        smn_carry = __shfl_sync(0xffffffffu, spm, 31);
        qx = __shfl_sync(0xffffffffu, cx, 31); qy = __shfl_sync(0xffffffffu, cy, 31); qz = __shfl_sync(0xffffffffu, cz, 31);

        if (in_t) {
            if (a == 0) {
                s.p1x[(size_t)j * s.t_max + i] = x; s.p1y[(size_t)j * s.t_max + i] = y;
                s.p1f[(size_t)j * s.t_max + i] = make_float2((float)DSUB(x, s.ox), (float)DSUB(y, s.oy));
            }
            s.kink[(size_t)pair * s.t_max + i] = make_double2(DSUB(t.box[0], x), DSUB(t.box[2], y));
            s.s1_prefmin[(size_t)pair * s.t_max + i] = pm;
            s.s1sub_prefmin[(size_t)pair * s.t_max + i] = spm;
        }
    }
    if (lane == 0) {
        s.s1_first_hit[pair] = first;
        s.s1sub_first_seg[pair] = sfirst;
        s.s1sub_first_time[pair] = stime;
        s.slab_dir[pair] = make_float2(t.dir0x, t.dir0y);   // refined by k_slab_axis when a run is large enough to pay for it
    }
}

// Reachability table (DCRP_FLAG_REACH_CULL), once per scenario, one warp per (track, drone), one lane
// per kink index.  After the kink the drone's step is (sin b * vf, cos b * vf, sin p * vf) with
// vf = v_straight - coef * pitch and pitch in [0, pi/2] (Drone.cpp:218-227), so vf lies between
// v_straight and v_ascend whatever the target: at index i > L the drone is horizontally within
// (i - L) * vmax of the kink and between start_alt and start_alt + (i - L) * vmax in altitude,
// vmax = max(|v_straight|, |v_ascend|).  Sample i can only be a conflict if the aircraft sample i is
// inside that set grown by R (+ a 1 cm margin that dwarfs the fp64 rounding of the accumulated path).
// [lo, hi) is the hull of the feasible indices of (pair, t1st); pairs with a stage-1 conflict keep
// the full range (their trials must be drawn to know which of them kink after the conflict).
__global__ void __launch_bounds__(128) k_reach(const ScenarioDev s)
{
    const int pair = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (pair >= s.n_tracks * s.n_drones) return;
    const int a = pair / s.n_drones, j = pair - a * s.n_drones;
    const TrackDev t = s.tracks[a];
    const bool keep_all = s.s1_first_hit[pair] != INT_MAX;
    const double vmax = fmax(fabs(s.v_straight), fabs(s.v_ascend)) * (1.0 + 1e-9);
    const double rm = s.radius + 0.01;
    bool any = false;
    for (int t1 = lane + 1; t1 <= s.t_max; t1 += 32) {
        int lo = INT_MAX, hi = 0;
        if (t1 <= t.takeoff_t) {
            if (keep_all) {
                lo = 0; hi = INT_MAX;        // never empty: a trial with t1st == n_steps has no stage-2 index at all
            } else {
                const int last = t1 - 1;
                const double kx = s.p1x[(size_t)j * s.t_max + last], ky = s.p1y[(size_t)j * s.t_max + last];
                for (int i = t1; i < t.n_steps; ++i) {
                    const double dx = s.ax[t.off + i] - kx, dy = s.ay[t.off + i] - ky;
                    const double az = s.az[t.off + i];
                    const double reach = (double)(i - last) * vmax + rm;
                    const bool feas = (dx * dx + dy * dy < reach * reach) && (az > s.start_alt - rm) &&
                                      (az < s.start_alt + reach);
                    if (feas) { lo = min(lo, i); hi = i + 1; }
                }
            }
            if (hi > lo) any = true;
        }
        s.reach[(size_t)pair * s.t_max + (t1 - 1)] = make_int2(lo, hi);
    }
    any = __any_sync(0xffffffffu, any);
    if (lane == 0) {
        if (any) s.alive[atomicAdd(&s.reach_tot[0], 1ull)] = (uint32_t)pair;
        else     atomicAdd(&s.reach_tot[1], (unsigned long long)(t.n_steps - t.takeoff_t));
    }
}

// =========================================================================
// the exact trial (fp64, reference order of operations)
// =========================================================================
struct TrialOut {
    int32_t first;     // first index (sub-step mode: first segment) in collision, -1 if none
    double  min_d2;    // min squared separation over all indices (sub-step: over all segments)
    double  first_time;// first-conflict time in index units (= first unless sub-step mode)
};

// ax/ay/az point at sample 0 of the track (global or shared).
__device__ __forceinline__ TrialOut exact_trial(const ScenarioDev& s, int pair, int j, int n_steps,
                                                const double* __restrict__ ax,
                                                const double* __restrict__ ay,
                                                const double* __restrict__ az, const Draw& d)
{
    const int last = d.t1st - 1;
    const double xl = s.p1x[(size_t)j * s.t_max + last];
    const double yl = s.p1y[(size_t)j * s.t_max + last];
    const Ray64 r = stage2_ray(xl, yl, s.start_alt, d.tx, d.ty, d.tz, s.v_straight, s.v_ascend);
    TrialOut o;
    const int s1 = s.s1_first_hit[pair];
    o.first = (s1 <= last) ? s1 : -1;
    o.min_d2 = s.s1_prefmin[(size_t)pair * s.t_max + last];
    double x = r.x, y = r.y, z = r.z;
    DCRP_ASSERT(d.t1st >= 1 && last < s.t_max);
    for (int i = d.t1st; i < n_steps; ++i) {                // Drone.cpp:224-230
        x = DADD(x, r.sx); y = DADD(y, r.sy); z = DADD(z, r.sz);
        const double d2 = sep2(x, y, z, ax[i], ay[i], az[i]);
        if (d2 < o.min_d2) o.min_d2 = d2;
        if (d2 < s.thr_d2 && o.first < 0) o.first = i;      // Drone.cpp:260 (sqrt folded into thr_d2)
    }
    o.first_time = (double)o.first;
    return o;
}

// The same trial with the separation interpolated between samples (extension; see seg_closest).
// Segments 0..t1st-2 lie in stage 1 and come from the per-pair tables; segment t1st-1 starts at the kink.
__device__ __forceinline__ TrialOut exact_trial_sub(const ScenarioDev& s, int pair, int j, int n_steps,
                                                    const double* __restrict__ ax,
                                                    const double* __restrict__ ay,
                                                    const double* __restrict__ az, const Draw& d)
{
    const int last = d.t1st - 1;
    const double xl = s.p1x[(size_t)j * s.t_max + last];
    const double yl = s.p1y[(size_t)j * s.t_max + last];
    const Ray64 r = stage2_ray(xl, yl, s.start_alt, d.tx, d.ty, d.tz, s.v_straight, s.v_ascend);
    TrialOut o;
    const int s1 = s.s1sub_first_seg[pair];
    o.first = (s1 < last) ? s1 : -1;
    o.first_time = (s1 < last) ? s.s1sub_first_time[pair] : -1.0;
    o.min_d2 = s.s1sub_prefmin[(size_t)pair * s.t_max + last];
    double x = r.x, y = r.y, z = r.z;
    double px = DSUB(x, ax[last]), py = DSUB(y, ay[last]), pz = DSUB(z, az[last]);
    for (int i = d.t1st; i < n_steps; ++i) {
        x = DADD(x, r.sx); y = DADD(y, r.sy); z = DADD(z, r.sz);
        const double cx = DSUB(x, ax[i]), cy = DSUB(y, ay[i]), cz = DSUB(z, az[i]);
        const SegOut sg = seg_closest(px, py, pz, cx, cy, cz, s.radius2);
        if (sg.d2 < o.min_d2) o.min_d2 = sg.d2;
        if (sg.tc >= 0.0 && o.first < 0) { o.first = i - 1; o.first_time = DADD((double)(i - 1), sg.tc); }
        px = cx; py = cy; pz = cz;
    }
    return o;
}

__device__ __forceinline__ TrialOut exact_any(const ScenarioDev& s, const RunDev& r, int pair, int j, int n_steps,
                                              const double* __restrict__ ax, const double* __restrict__ ay,
                                              const double* __restrict__ az, const Draw& d)
{
    return r.substep ? exact_trial_sub(s, pair, j, n_steps, ax, ay, az, d)
                     : exact_trial(s, pair, j, n_steps, ax, ay, az, d);
}

__device__ __forceinline__ Draw fetch_draw(const ScenarioDev& s, const RunDev& r, const TrackDev& t,
                                           int pair, int a, int j, uint64_t local)
{
    if (r.replay) return r.replay[(size_t)pair * r.trial_count + local];
    return make_draw(r.seed, r.trial_begin + local, r.drone_base + (uint32_t)j, r.track_base + (uint32_t)a,
                     t.takeoff_t, t.box);
}

__device__ __forceinline__ void publish_hit(const ScenarioDev& s, const RunDev& r, int pair, int a, int j,
                                            uint64_t trial, const Draw& d, const TrialOut& o)
{
    atomicAdd(&r.counts[pair], 1ull);
    atomicMin(&r.first_conflict[pair], o.first);
    const uint32_t slot = warp_claim_slot(r.rec_count);
    DCRP_ASSERT(pair >= 0 && pair < s.n_tracks * s.n_drones && o.first >= 0);
    if (slot < r.rec_cap) {
        Record rec;
        rec.track = (uint32_t)a; rec.drone = (uint32_t)j; rec.trial = trial;
        rec.t1st = d.t1st; rec.first_hit = o.first;
        rec.tx = d.tx; rec.ty = d.ty; rec.tz = d.tz;
        rec.min_dist = DSQRT(o.min_d2);
        rec.first_time = o.first_time;
        r.records[slot] = rec;
    }
}

__device__ __forceinline__ void flush_stats(const RunDev& r, unsigned long long ev, unsigned long long sh,
                                            unsigned long long is, unsigned long long st,
                                            unsigned long long hits)
{
    ev = warp_sum_u64(ev); sh = warp_sum_u64(sh); is = warp_sum_u64(is);
    st = warp_sum_u64(st); hits = warp_sum_u64(hits);
    if (lane_id() == 0) {
        if (ev) atomicAdd(&r.stats[ST_EVALUATED], ev);
        if (sh) atomicAdd(&r.stats[ST_SHARED], sh);
        if (is) atomicAdd(&r.stats[ST_ISSUED], is);
        if (st) atomicAdd(&r.stats[ST_STEPS], st);
        if (hits) atomicAdd(&r.stats[ST_HITS], hits);
    }
}

// Every trial in fp64.  Persistent CTAs walk (pair, chunk-of-EXB-trials) items.
constexpr int EXB = 256;

__global__ void __launch_bounds__(EXB) k_exact(const ScenarioDev s, const RunDev r,
                                               uint32_t chunks_per_pair, uint32_t n_items, int use_smem)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sm = reinterpret_cast<double*>(smem_raw);
    unsigned long long ev = 0, sh = 0, st = 0, hits = 0;
    int loaded = -1;
    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int pair = item / chunks_per_pair;
        const uint32_t c = item - pair * chunks_per_pair;
        const int a = pair / s.n_drones, j = pair - a * s.n_drones;
        const TrackDev t = s.tracks[a];
        const double *ax = s.ax + t.off, *ay = s.ay + t.off, *az = s.az + t.off;
        if (use_smem) {
            if (loaded != a) {
                __syncthreads();
                for (int i = threadIdx.x; i < t.n_steps; i += EXB) {
                    sm[i] = ax[i];
                    sm[s.n_steps_max + i] = ay[i];
                    sm[2 * s.n_steps_max + i] = az[i];
                }
                __syncthreads();
                loaded = a;
            }
            ax = sm; ay = sm + s.n_steps_max; az = sm + 2 * s.n_steps_max;
        }
        const uint64_t in_slice = (uint64_t)c * EXB + threadIdx.x;
        const uint64_t local = r.slice_first + in_slice;       // index within the run
        if (in_slice < r.slice_count) {
            const Draw d = fetch_draw(s, r, t, pair, a, j, local);
            const TrialOut o = exact_any(s, r, pair, j, t.n_steps, ax, ay, az, d);
            ev += (unsigned long long)(t.n_steps - d.t1st);
            sh += (unsigned long long)d.t1st;
            st += (unsigned long long)(t.n_steps - 1);
            if (r.trial_hit) {
                const size_t idx = (size_t)pair * r.trial_count + local;
                r.trial_hit[idx] = o.first >= 0 ? 1 : 0;
                r.trial_first[idx] = o.first;
                r.trial_min[idx] = DSQRT(o.min_d2);
            }
            if (o.first >= 0) {
                ++hits;
                publish_hit(s, r, pair, a, j, r.trial_begin + local, d, o);
            }
        }
    }
    flush_stats(r, ev, sh, ev, st, hits);
}

// The exact trial, one WARP per trial: lane l takes the stage-2 indices c0 + l of every 32-index pass.
// The positions are the reference's running sums (Drone.cpp:224-230), so each lane re-runs the l + 1
// additions from the pass's base position -- the same additions in the same order as exact_trial, hence
// the same bits -- while the 32 track samples of the pass arrive in one coalesced load instead of 32
// dependent ones.  First conflict = lowest index (ballot), minimum = warp min: both order-independent.
__device__ __forceinline__ TrialOut exact_any_warp(const ScenarioDev& s, const RunDev& r, int pair, int j,
                                                   int n_steps, const double* __restrict__ ax,
                                                   const double* __restrict__ ay, const double* __restrict__ az,
                                                   const Draw& d, int lane)
{
    const int last = d.t1st - 1;
    const double xl = s.p1x[(size_t)j * s.t_max + last];
    const double yl = s.p1y[(size_t)j * s.t_max + last];
    const Ray64 ray = stage2_ray(xl, yl, s.start_alt, d.tx, d.ty, d.tz, s.v_straight, s.v_ascend);
    int first = -1;
    double ftime = -1.0, mn;
    if (r.substep) {
        const int s1 = s.s1sub_first_seg[pair];
        if (s1 < last) { first = s1; ftime = s.s1sub_first_time[pair]; }
        mn = s.s1sub_prefmin[(size_t)pair * s.t_max + last];
    } else {
        const int s1 = s.s1_first_hit[pair];
        if (s1 <= last) { first = s1; ftime = (double)s1; }
        mn = s.s1_prefmin[(size_t)pair * s.t_max + last];
    }
    double bx = ray.x, by = ray.y, bz = ray.z;                         // position at index c0 - 1
    double qx = DSUB(bx, ax[last]), qy = DSUB(by, ay[last]), qz = DSUB(bz, az[last]);   // sub-step: previous sample
    for (int c0 = d.t1st; c0 < n_steps; c0 += 32) {
        const int i = c0 + lane;
        double x = bx, y = by, z = bz;
        for (int k = 0; k <= lane; ++k) { x = DADD(x, ray.sx); y = DADD(y, ray.sy); z = DADD(z, ray.sz); }
        bx = __shfl_sync(0xffffffffu, x, 31); by = __shfl_sync(0xffffffffu, y, 31); bz = __shfl_sync(0xffffffffu, z, 31);
        const bool on = i < n_steps;
        double cx = 0, cy = 0, cz = 0, d2 = INFINITY, tc = -1.0;
        if (on) { cx = DSUB(x, ax[i]); cy = DSUB(y, ay[i]); cz = DSUB(z, az[i]); }
        if (r.substep) {
            double px = __shfl_up_sync(0xffffffffu, cx, 1), py = __shfl_up_sync(0xffffffffu, cy, 1),
                   pz = __shfl_up_sync(0xffffffffu, cz, 1);
            if (lane == 0) { px = qx; py = qy; pz = qz; }
            if (on) {
                const SegOut sg = seg_closest(px, py, pz, cx, cy, cz, s.radius2);
                d2 = sg.d2; tc = sg.tc;
            }
            qx = __shfl_sync(0xffffffffu, cx, 31); qy = __shfl_sync(0xffffffffu, cy, 31); qz = __shfl_sync(0xffffffffu, cz, 31);
            const uint32_t hm = __ballot_sync(0xffffffffu, tc >= 0.0);
            if (first < 0 && hm) {
                const int src = __ffs((int)hm) - 1;
                first = c0 + src - 1;                                           // the segment's first sample
                ftime = DADD((double)first, __shfl_sync(0xffffffffu, tc, src));
            }
        } else {
            if (on) d2 = sep2(x, y, z, ax[i], ay[i], az[i]);
            const uint32_t hm = __ballot_sync(0xffffffffu, d2 < s.thr_d2);     // Drone.cpp:260, sqrt folded into thr_d2
            if (first < 0 && hm) { first = c0 + __ffs((int)hm) - 1; ftime = (double)first; }
        }
        mn = fmin(mn, d2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    TrialOut out;
    out.first = first; out.min_d2 = mn; out.first_time = r.substep ? ftime : (double)first;
    return out;
}

// fp64 confirmation of the screen's candidates, one warp per candidate.  The list length lives on the
// device (no host round trip between screen and confirm): warp-stride over it.
__global__ void __launch_bounds__(128) k_confirm_all(const ScenarioDev s, const RunDev r)
{
    const uint32_t n_cand = min(*r.cand_count, r.cand_cap);
    const int lane = threadIdx.x & 31;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    unsigned long long hits = 0;
    for (uint32_t idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; idx < n_cand; idx += n_warps) {
        const Candidate cd = r.cands[idx];
        const int pair = (int)cd.pair;
        const int a = pair / s.n_drones, j = pair - a * s.n_drones;
        const TrackDev t = s.tracks[a];
        const uint64_t local = cd.trial - r.trial_begin;
        const Draw d = fetch_draw(s, r, t, pair, a, j, local);
        const TrialOut o = exact_any_warp(s, r, pair, j, t.n_steps, s.ax + t.off, s.ay + t.off, s.az + t.off, d, lane);
        if (o.first >= 0 && lane == 0) {
            ++hits;
            publish_hit(s, r, pair, a, j, cd.trial, d, o);
        }
    }
    flush_stats(r, 0, 0, 0, 0, hits);
}

// =========================================================================
// the fp32 packed screen: every trial
// =========================================================================
struct Setup32 {
    Ray32   ray;
    int32_t t1st;
    bool    force;       // must be confirmed regardless of the scan (degenerate / stage-1 hit)
    double  tx, ty, tz;
};

// q = staged trial: (w1, w2, w3, local | t1st << 12)
__device__ __forceinline__ Setup32 setup_trial32(const ScenarioDev& s, const RunDev& r, const TrackDev& t,
                                                 int pair, int j, int s1_first, uint64_t chunk_first,
                                                 const uint4 q)
{
    Setup32 o;
    o.t1st = (int32_t)(q.w >> 12);
    const uint32_t local = q.w & 4095u;
    if (r.replay) {
        const Draw d = r.replay[(size_t)pair * r.trial_count + chunk_first + local];
        o.tx = d.tx; o.ty = d.ty; o.tz = d.tz;
    } else {
        o.tx = draw_real(q.x, t.box[0], t.box[1]);
        o.ty = draw_real(q.y, t.box[2], t.box[3]);
        o.tz = draw_real(q.z, t.box[4], t.box[5]);
    }
    const int last = o.t1st - 1;
    const double xl = __ldg(&s.p1x[(size_t)j * s.t_max + last]);
    const double yl = __ldg(&s.p1y[(size_t)j * s.t_max + last]);
    // differences and the recentring are formed in fp64 and rounded ONCE to fp32
    const float dls = (float)DSUB(o.tx, xl);
    const float dbs = (float)DSUB(o.ty, yl);
    const float da  = (float)DSUB(o.tz, s.start_alt);
    const float xlf = (float)DSUB(xl, s.ox);
    const float ylf = (float)DSUB(yl, s.oy);
    const float zlf = (float)DSUB(s.start_alt, s.oz);
    o.ray = stage2_ray32(dls, dbs, da, xlf, ylf, zlf, (float)last, s.v_straight_f, s.coef_f);
    o.force = o.ray.degenerate || (s1_first <= last);
    return o;
}

// Scalar fp32 re-scan of ONE trial with exact index bookkeeping (rare path of
// DCRP_PRECISION_FP32): starts at the trial's own t1st, keeps first index + min.
__device__ __forceinline__ void rescan_fp32(const ScenarioDev& s, const RunDev& r, const TrackDev& t,
                                            int pair, int a, int j, int s1_first, uint64_t trial,
                                            const Setup32& u)
{
    const float4* __restrict__ g = s.trk32 + (size_t)t.off;
    const int last = u.t1st - 1;
    int first = (s1_first <= last) ? s1_first : -1;
    float mn = (float)__ldg(&s.s1_prefmin[(size_t)pair * s.t_max + last]);
    for (int i = u.t1st; i < t.n_steps; ++i) {
        const float4 p = __ldg(&g[i]);                  // (-ax, -ay, -az, i)
        const float dx = fmaf(p.w, u.ray.sx, u.ray.bx) + p.x;
        const float dy = fmaf(p.w, u.ray.sy, u.ray.by) + p.y;
        const float dz = fmaf(p.w, u.ray.sz, u.ray.bz) + p.z;
        const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        mn = fminf(mn, d2);
        if (d2 < s.thr2_fp32 && first < 0) first = i;
    }
    if (first >= 0) {
        Draw d; d.t1st = u.t1st; d.reserved = 0; d.tx = u.tx; d.ty = u.ty; d.tz = u.tz;
        TrialOut o; o.first = first; o.min_d2 = (double)mn; o.first_time = (double)first;
        atomicAdd(&r.stats[ST_HITS], 1ull);
        publish_hit(s, r, pair, a, j, trial, d, o);
    }
}

// Level 2 of the screen: min over the stage-2 indices of the full 3-D fp32 squared separation of ONE
// trial (fused form P = i*S + B: no accumulation), track read through L1 from global memory.
__device__ __forceinline__ float min3d_fp32(const ScenarioDev& s, const TrackDev& t, const Setup32& u)
{
    const float4* __restrict__ g = s.trk32 + (size_t)t.off;
    float mn = 3.0e38f;
    for (int i = u.t1st; i < t.n_steps; ++i) {
        const float4 p = __ldg(&g[i]);                  // (-ax, -ay, -az, i)
        const float dx = fmaf(p.w, u.ray.sx, u.ray.bx) + p.x;
        const float dy = fmaf(p.w, u.ray.sy, u.ray.by) + p.y;
        const float dz = fmaf(p.w, u.ray.sz, u.ray.bz) + p.z;
        mn = fminf(mn, fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
    }
    return mn;
}

// The same minimum by a whole warp on ONE trial (u is warp-uniform): lane l takes the indices t1st + l + 32 k.
// Every index goes through the same fused expression as in min3d_fp32 and a minimum does not depend on the
// order, so the value is the same bits; the latency is that of 3 indices instead of 96.
__device__ __forceinline__ float min3d_fp32_warp(const ScenarioDev& s, const TrackDev& t, const Setup32& u, int lane)
{
    const float4* __restrict__ g = s.trk32 + (size_t)t.off;
    float mn = 3.0e38f;
    for (int i = u.t1st + lane; i < t.n_steps; i += 32) {
        const float4 p = __ldg(&g[i]);                  // (-ax, -ay, -az, i)
        const float dx = fmaf(p.w, u.ray.sx, u.ray.bx) + p.x;
        const float dy = fmaf(p.w, u.ray.sy, u.ray.by) + p.y;
        const float dz = fmaf(p.w, u.ray.sz, u.ray.bz) + p.z;
        mn = fminf(mn, fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    return mn;
}

// Level 2 in sub-step mode: min over the per-trial SEGMENTS [i, i+1], i >= t1st-1, of the fp32 closest
// approach of the interpolated relative position.
__device__ __forceinline__ float min3d_seg_fp32(const ScenarioDev& s, const TrackDev& t, const Setup32& u)
{
    const float4* __restrict__ g = s.trk32 + (size_t)t.off;
    float mn = 3.0e38f;
    const int last = u.t1st - 1;
    float4 p = __ldg(&g[last]);
    float x0 = fmaf(p.w, u.ray.sx, u.ray.bx) + p.x, y0 = fmaf(p.w, u.ray.sy, u.ray.by) + p.y,
          z0 = fmaf(p.w, u.ray.sz, u.ray.bz) + p.z;
    float c = fmaf(z0, z0, fmaf(y0, y0, x0 * x0));
    const float near = t.thr2_near_sub;
    for (int i = u.t1st; i < t.n_steps; ++i) {
        p = __ldg(&g[i]);
        const float x1 = fmaf(p.w, u.ray.sx, u.ray.bx) + p.x, y1 = fmaf(p.w, u.ray.sy, u.ray.by) + p.y,
                    z1 = fmaf(p.w, u.ray.sz, u.ray.bz) + p.z;
        const float c1 = fmaf(z1, z1, fmaf(y1, y1, x1 * x1));
        // most of a trial's ~100 segments are kilometres away: with both ends beyond thr2_near_sub the closest approach
        // cannot reach the level-2 radius (TrackDev) and the value is only ever compared with that radius
        if (fminf(c, c1) < near) {
            const float ex = x1 - x0, ey = y1 - y0, ez = z1 - z0;
            const float a = fmaf(ez, ez, fmaf(ey, ey, ex * ex));
            const float b = fmaf(z0, ez, fmaf(y0, ey, x0 * ex));
            const float tt = a > 0.f ? fminf(fmaxf(__fdividef(-b, a), 0.f), 1.f) : 0.f;
            // the interior minimum c - b^2/a cancels: bound its fp32 error by comparing against both ends too
            const float dmin = fmaf(tt, fmaf(tt, a, 2.f * b), c);
            mn = fminf(mn, fminf(dmin, c));
        }
        x0 = x1; y0 = y1; z0 = z1; c = c1;
    }
    return mn;
}

// The same minimum by a whole warp on ONE trial (u is warp-uniform): lane l takes the segments that END at the indices
// t1st + l + 32 k.  Both ends of a segment go through the expressions of min3d_seg_fp32 (there the first end is carried
// over from the previous iteration: the same fused expression at the same index, hence the same bits), and a minimum
// does not depend on the order.
__device__ __forceinline__ float min3d_seg_fp32_warp(const ScenarioDev& s, const TrackDev& t, const Setup32& u, int lane)
{
    const float4* __restrict__ g = s.trk32 + (size_t)t.off;
    float mn = 3.0e38f;
    for (int i = u.t1st + lane; i < t.n_steps; i += 32) {
        const float4 q = __ldg(&g[i - 1]), p = __ldg(&g[i]);
        const float x0 = fmaf(q.w, u.ray.sx, u.ray.bx) + q.x, y0 = fmaf(q.w, u.ray.sy, u.ray.by) + q.y,
                    z0 = fmaf(q.w, u.ray.sz, u.ray.bz) + q.z;
        const float x1 = fmaf(p.w, u.ray.sx, u.ray.bx) + p.x, y1 = fmaf(p.w, u.ray.sy, u.ray.by) + p.y,
                    z1 = fmaf(p.w, u.ray.sz, u.ray.bz) + p.z;
        const float c = fmaf(z0, z0, fmaf(y0, y0, x0 * x0));
        const float c1 = fmaf(z1, z1, fmaf(y1, y1, x1 * x1));
        if (fminf(c, c1) < t.thr2_near_sub) {            // (as in min3d_seg_fp32)
            const float ex = x1 - x0, ey = y1 - y0, ez = z1 - z0;
            const float a = fmaf(ez, ez, fmaf(ey, ey, ex * ex));
            const float b = fmaf(z0, ez, fmaf(y0, ey, x0 * ex));
            const float tt = a > 0.f ? fminf(fmaxf(__fdividef(-b, a), 0.f), 1.f) : 0.f;
            const float dmin = fmaf(tt, fmaf(tt, a, 2.f * b), c);
            mn = fminf(mn, fminf(dmin, c));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    return mn;
}

// =========================================================================
// k_screen2: the fp32 packed screen, balanced rounds
// =========================================================================
// Every trial is screened conservatively in fp32 and the rare candidates are appended for the fp64
// confirm.  Per work item a CTA draws CH = BLOCK * 2*NP * R trials (Philox), bins them by t1st (counting
// sort in shared memory), then every warp runs R rounds of (set up 2*NP trials per lane -> scan them as NP
// packed pairs).  Round pairs take complementary bins (short scan + long scan), so all warps of a CTA do
// the same amount of scan work per item and do not wait for each other at the next barrier.  Production:
// R = 2, BLOCK = 384, NP = 2 (3072 trials per item, 80 registers, 2 CTAs per SM); DESIGN.md 5.1 has the
// measurements behind each choice.  Template flags: REPLAY (draws from a table), CULL (reach culling,
// DESIGN.md 5.3), DBG (phase-ablation switches for profiling; never set in a production variant).
struct Fast32 {          // horizontal ray of one trial for level 1: position(i) = b + i*s
    float bx, by, sx, sy;
    int32_t t1st;
    bool force;
};

// kk / pl point at this pair's kink tables (shared memory when takeoff_t <= SCREEN_TAB, else global).
struct TrackScale {      // per-track constants of the fast set-up, fetched once per item
    double wx32, wy32;
    float  wz32, z0rel;
};

__device__ __forceinline__ Fast32 setup_fast32(const ScenarioDev& s, const TrackScale& ts,
                                               const double2* __restrict__ kk_tab,
                                               const float2* __restrict__ pl_tab, int s1_first, const uint4 q)
{
    const double wx32 = ts.wx32, wy32 = ts.wy32;
    const float wz32 = ts.wz32, z0rel = ts.z0rel;
    Fast32 o;
    o.t1st = (int32_t)(q.w >> 12);
    const int last = o.t1st - 1;
    DCRP_ASSERT(last >= 0 && last < s.t_max);
    const double2 kk = kk_tab[last];
    const float2 pl = pl_tab[last];
    // target - kink: one fp64 FMA per coordinate, rounded once to fp32
    const float dls = (float)fma((double)q.x, wx32, kk.x);
    const float dbs = (float)fma((double)q.y, wy32, kk.y);
    const float ada = fabsf(fmaf((float)q.z, wz32, z0rel));
    const float m2 = fmaf(dls, dls, dbs * dbs);
    float inv_m;                                                // one MUFU.RSQ: m2 > 1e-12 or the trial is forced
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(inv_m) : "f"(m2));
    const float pitch = atan_nonneg(ada * inv_m);               // Drone.cpp:218
    const float k = inv_m * fmaf(-s.coef_f, pitch, s.v_straight_f);   // vf / m, Drone.cpp:220
    const float lastf = (float)last;
    o.sx = dls * k;
    o.sy = dbs * k;
    o.bx = fmaf(-lastf, o.sx, pl.x);
    o.by = fmaf(-lastf, o.sy, pl.y);
    o.force = !(m2 > 1e-12f) || (s1_first <= last);
    return o;
}

constexpr int SCREEN_TAB = 128;     // kink-table entries kept in shared memory per CTA

template <int R, int BLOCK, bool REPLAY, int MINB, int NP, bool CULL, bool DBG>
__global__ void __launch_bounds__(BLOCK, MINB)
k_screen2(const ScenarioDev s, const RunDev r, uint32_t chunks_per_pair, uint32_t n_items, int tile_steps,
          uint32_t stagger_ns, uint32_t dbg)
{
    constexpr int TR = 2 * NP;               // trials per thread per round (NP packed pairs)
    constexpr int K = TR * R;                // trials per thread per item
    constexpr int CH = K * BLOCK;
    constexpr int NW = BLOCK / 32;
    static_assert(R % 2 == 0, "rounds come in complementary (long, short) pairs");
    static_assert(CH <= 4096, "local trial index is packed in 12 bits");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2*   tiles = reinterpret_cast<float2*>(smem_raw);                          // 2 x tile_steps (x, y)
    uint4*    stage = reinterpret_cast<uint4*>(smem_raw + (size_t)2 * tile_steps * 8);
    uint32_t* hist  = reinterpret_cast<uint32_t*>(stage + CH);
    uint64_t* bars  = reinterpret_cast<uint64_t*>(hist + 80);
    double2*  tabk  = reinterpret_cast<double2*>(bars + 2);            // this pair's kink tables
    float2*   tabp  = reinterpret_cast<float2*>(tabk + SCREEN_TAB);
    int2*     tabr  = reinterpret_cast<int2*>(tabp + SCREEN_TAB);          // this pair's reach ranges
    // K > 4: the draws stay where phase A wrote them (stage[local]) and only a 2-byte permutation is
    // sorted (order[sorted slot] = local), so phase A never holds more than one trial's words in registers
    constexpr bool IND = (K > 4);
    uint16_t* order = reinterpret_cast<uint16_t*>(tabr + SCREEN_TAB);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile_shift = tile_steps == 128 ? 7 : 9;      // the host passes 128 or 512

    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_fence_init(); }
    __syncthreads();

    uint32_t phase0 = 0, phase1 = 0;
    int loaded_track = -1;
    uint32_t acc_t1 = 0, acc_is = 0, acc_cull = 0, ncand = 0, npass = 0;   // partial sums (flushed before overflow)

    // CTAs co-resident on an SM all start at once and, doing identical work per item, would stay in
    // phase (all in the integer/set-up phases together, then all fighting for the FMA pipe).  Offset
    // the k-th CTA of an SM by k/occupancy of an item so that scans overlap the other phases.
    if (stagger_ns) __nanosleep(stagger_ns * (blockIdx.x / s.n_sm));

    // Reach culling: only the pairs of the scenario's alive list get items; the others are accounted
    // for here, once per launch, with the tests every one of their trials is guaranteed to hold.
    if (CULL) {
        const unsigned long long n_alive = s.reach_tot[0];
        n_items = (uint32_t)(n_alive * chunks_per_pair);
        if (blockIdx.x == 0 && tid == 0) {
            const unsigned long long dead = (unsigned long long)(s.n_tracks * s.n_drones) - n_alive;
            atomicAdd(&r.stats[ST_CULLED_TRIALS], dead * r.slice_count);
            atomicAdd(&r.stats[ST_DEAD_TESTS], s.reach_tot[1] * r.slice_count);
        }
    }

    // (pair, chunk) of the item, carried along the grid-stride loop: divisions only when a pair ends
    int po = (int)(blockIdx.x / chunks_per_pair);          // ordinal in the list of pairs that get work
    uint32_t c = blockIdx.x - (uint32_t)po * chunks_per_pair;
    int pair = (CULL && blockIdx.x < n_items) ? (int)s.alive[po] : po;
    int a = pair / s.n_drones, j = pair - a * s.n_drones;
    int tab_pair = -1;
    const bool tab_smem = s.t_max <= SCREEN_TAB;

    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const TrackDev* __restrict__ tk = s.tracks + a;
        const int VL = __ldg(&tk->n_steps);
        const int takeoff_t = __ldg(&tk->takeoff_t);
        const int n_tiles = (VL + tile_steps - 1) >> tile_shift;
        const uint64_t chunk_first64 = r.slice_first + (uint64_t)c * CH;
        const uint32_t n_valid = (uint32_t)min((uint64_t)CH, r.slice_count - (uint64_t)c * CH);
        // first stage-1 conflict, as "every trial whose kink index is >= this must be confirmed":
        // a sample index in sample mode, (segment index + 1) in sub-step mode
        int s1_first = __ldg(&s.s1_first_hit[pair]);
        if (r.substep) {
            const int sg = __ldg(&s.s1sub_first_seg[pair]);
            s1_first = sg == INT_MAX ? INT_MAX : sg + 1;
        }
        const float2* __restrict__ gtrk = s.trkxy + (size_t)__ldg(&tk->off_xy);

        if (acc_t1 > 0x40000000u || acc_is > 0x40000000u || acc_cull > 0x40000000u) {   // rare: keep the 32-bit partial sums exact
            atomicAdd(&r.stats[ST_SHARED], (unsigned long long)acc_t1);
            atomicAdd(&r.stats[ST_ISSUED], (unsigned long long)acc_is);
            atomicAdd(&r.stats[ST_CULLED_TESTS], (unsigned long long)acc_cull);
            acc_t1 = acc_is = acc_cull = 0;
        }
        if (CULL && tid == 0) atomicAdd(&r.stats[ST_PROCESSED_TESTS], (unsigned long long)n_valid * (unsigned long long)VL);

        // ---------------- phase A: draws + histogram over t1st buckets
        for (int h = tid; h < 80; h += BLOCK) hist[h] = 0;
        __syncthreads();      // every thread is done with the previous item's stage[] and tiles

        if (tab_smem && tab_pair != pair) {       // after the barrier above: nobody reads the old tables
            if (tid < takeoff_t) {
                tabk[tid] = __ldg(&s.kink[(size_t)pair * s.t_max + tid]);
                tabp[tid] = __ldg(&s.p1f[(size_t)j * s.t_max + tid]);
                if (CULL) tabr[tid] = __ldg(&s.reach[(size_t)pair * s.t_max + tid]);
            }
            tab_pair = pair;
        }
        TrackScale ts;
        ts.wx32 = __ldg(&tk->wx32); ts.wy32 = __ldg(&tk->wy32);
        ts.wz32 = __ldg(&tk->wz32); ts.z0rel = __ldg(&tk->z0rel);

        const bool single = (n_tiles == 1);
        const bool need_load = !(single && loaded_track == a);
        if (single && need_load && tid == 0) {
            fence_proxy_async();
            const uint32_t b0 = (uint32_t)((VL + 1) & ~1) * 8u;       // whole 16-B units (tracks are padded)
            mbar_arrive_expect_tx(&bars[0], b0);
            tma_bulk_g2s(tiles, gtrk, b0, &bars[0]);
        }

        {
            uint32_t w1[IND ? 1 : K], w2[IND ? 1 : K], w3[IND ? 1 : K], meta[IND ? 1 : K], rank[K];
#pragma unroll
            for (int m = 0; m < K; ++m) {
                const int mm = IND ? 0 : m;
                const uint32_t local = (uint32_t)(m * BLOCK + tid);
                uint32_t bucket = 64u;
                meta[mm] = local;
                w1[mm] = w2[mm] = w3[mm] = 0u;
                if (local < n_valid) {
                    int32_t t1st;
                    if (REPLAY) {
                        t1st = r.replay[(size_t)pair * r.trial_count + chunk_first64 + local].t1st;
                    } else if (DBG && (dbg & 4u)) {      // profiling aid: no Philox (results meaningless)
                        const uint32_t h = (uint32_t)(chunk_first64 + local) * 2654435761u;
                        t1st = draw_t1st(h, takeoff_t);
                        w1[mm] = h ^ 0x9e3779b9u; w2[mm] = h * 40503u; w3[mm] = h + 0x7f4a7c15u;
                    } else {
                        const uint64_t trial = r.trial_begin + chunk_first64 + local;
                        const U4 w = philox4x32_10_rk((uint32_t)trial, (uint32_t)(trial >> 32),
                                                      r.drone_base + (uint32_t)j, r.track_base + (uint32_t)a, r.rk);
                        t1st = draw_t1st(w.x, takeoff_t);
                        w1[mm] = w.y; w2[mm] = w.z; w3[mm] = w.w;
                    }
                    bucket = (uint32_t)(t1st - 1) >> s.bucket_shift;
                    DCRP_ASSERT(t1st >= 1 && t1st <= takeoff_t && bucket < 64u);
                    meta[mm] = local | ((uint32_t)t1st << 12);
                }
                if (IND) stage[local] = make_uint4(w1[0], w2[0], w3[0], meta[0]);
                rank[m] = atomicAdd(&hist[bucket], 1u) | (bucket << 16);
            }
            __syncthreads();
            if (warp == 0) {
                const uint32_t v0 = hist[lane], v1 = hist[lane + 32];
                uint32_t i0 = v0, i1 = v1;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t u0 = __shfl_up_sync(0xffffffffu, i0, o);
                    const uint32_t u1 = __shfl_up_sync(0xffffffffu, i1, o);
                    if (lane >= o) { i0 += u0; i1 += u1; }
                }
                const uint32_t tot0 = __shfl_sync(0xffffffffu, i0, 31);
                const uint32_t tot1 = __shfl_sync(0xffffffffu, i1, 31);
                hist[lane] = i0 - v0;
                hist[lane + 32] = tot0 + i1 - v1;
                if (lane == 0) hist[64] = tot0 + tot1;
            }
            __syncthreads();
#pragma unroll
            for (int m = 0; m < K; ++m) {
                const uint32_t slot = hist[rank[m] >> 16] + (rank[m] & 0xffffu);
                DCRP_ASSERT(slot < (uint32_t)CH);
                if (IND) order[slot] = (uint16_t)(m * BLOCK + tid);
                else     stage[slot] = make_uint4(w1[IND ? 0 : m], w2[IND ? 0 : m], w3[IND ? 0 : m], meta[IND ? 0 : m]);
            }
        }
        __syncthreads();

        if (single && need_load) { mbar_wait(&bars[0], phase0); phase0 ^= 1u; }
        loaded_track = single ? a : -1;

        // sub-step mode: a segment can only come within R of the aircraft if one of its end SAMPLES is
        // within R + half the largest relative step, so level 1 keeps testing samples against a wider radius
        const float thr = r.decide_fp32 ? s.thr2_fp32 : (r.substep ? __ldg(&tk->thr2_l1_sub) : s.thr2_screen);

        // ---------------- R rounds: set up one packed pair, scan it
#pragma unroll 1
        for (int rd = 0; rd < R; ++rd) {
            // bins of 32*TR trials in ascending t1st order; rounds (2q, 2q+1) take bin b and its mirror
            const int q2 = rd >> 1;
            const int bin = (rd & 1) ? (q2 * 2 * NW + 2 * NW - 1 - warp) : (q2 * 2 * NW + warp);
            const uint32_t slot0 = (uint32_t)(bin * (32 * TR) + lane);
            float bxs[TR], bys[TR], sxs[TR], sys[TR];  // level 1 needs the horizontal ray only
            int tmin = INT_MAX;
            uint32_t fmask = 0;                   // bit m: trial m must go to level 2 regardless of the scan
            uint32_t sidx[TR];                    // where this lane's trials sit in stage[]
            // Reach culling: the union [clo, chi) of the feasible index ranges of this warp's trials (the
            // bins are sorted by t1st, so it is tight); an empty union means none of them can collide
            // whatever their targets, and the round needs neither ray set-up nor scan.
            int clo = 0, chi = INT_MAX;
            bool skip_round = false;
            if (CULL) {
                int lo_l = INT_MAX, hi_l = 0;
                uint32_t t1s = 0, nv = 0;
#pragma unroll
                for (int m = 0; m < TR; ++m) {
                    const uint32_t si = IND ? (uint32_t)order[slot0 + m * 32] : slot0 + m * 32;
                    const uint32_t t1 = stage[si].w >> 12;
                    if (t1 != 0u) {
                        const int2 rg = tab_smem ? tabr[t1 - 1] : __ldg(&s.reach[(size_t)pair * s.t_max + (t1 - 1)]);
                        lo_l = min(lo_l, rg.x); hi_l = max(hi_l, rg.y);
                        t1s += t1; ++nv;
                    }
                }
                clo = __reduce_min_sync(0xffffffffu, lo_l);
                chi = __reduce_max_sync(0xffffffffu, hi_l);
                if (chi <= clo) {
                    acc_t1 += t1s;
                    acc_cull += nv * (uint32_t)VL - t1s;
                    if (single) continue;
                    skip_round = true;      // multi-tile tracks: the tile hand-over below is CTA-wide, walk it idle
                }
            }
#pragma unroll
            for (int m = 0; m < TR; ++m) {
                sidx[m] = IND ? (uint32_t)order[slot0 + m * 32] : slot0 + m * 32;
                const uint4 q = stage[sidx[m]];
                float bx = 1e15f, by = 1e15f, sx = 0.f, sy = 0.f;
                if ((q.w >> 12) != 0u && !skip_round) {
                    int32_t t1;
                    if (REPLAY) {
                        const Setup32 u = setup_trial32(s, r, *tk, pair, j, s1_first, chunk_first64, q);
                        bx = u.ray.bx; by = u.ray.by; sx = u.ray.sx; sy = u.ray.sy;
                        fmask |= u.force ? (1u << m) : 0u; t1 = u.t1st;
                    } else if (DBG && (dbg & 2u)) {      // profiling aid: no ray set-up (results meaningless)
                        bx = __uint_as_float(q.x >> 9); by = __uint_as_float(q.y >> 9);
                        sx = 1.f; sy = 2.f; t1 = (int32_t)(q.w >> 12);
                    } else {
                        // two calls so that each sees a typed (shared / global) table pointer, not a generic one
                        const Fast32 u = tab_smem
                            ? setup_fast32(s, ts, tabk, tabp, s1_first, q)
                            : setup_fast32(s, ts, s.kink + (size_t)pair * s.t_max, s.p1f + (size_t)j * s.t_max,
                                           s1_first, q);
                        bx = u.bx; by = u.by; sx = u.sx; sy = u.sy;
                        fmask |= u.force ? (1u << m) : 0u; t1 = u.t1st;
                    }
                    tmin = min(tmin, t1);
                    acc_t1 += (uint32_t)t1;
                    if (CULL) acc_cull += (uint32_t)((VL - t1) - max(0, min(VL, chi) - max(clo, t1)));
                }
                bxs[m] = bx; bys[m] = by; sxs[m] = sx; sys[m] = sy;
            }
            f2_t BX[NP], BY[NP], SX[NP], SY[NP];
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                BX[p] = pk2(bxs[2 * p], bxs[2 * p + 1]); BY[p] = pk2(bys[2 * p], bys[2 * p + 1]);
                SX[p] = pk2(sxs[2 * p], sxs[2 * p + 1]); SY[p] = pk2(sys[2 * p], sys[2 * p + 1]);
            }
            // sub-step mode also needs the kink sample (index t1st-1), the first end of the first segment
            const int imin = max(__reduce_min_sync(0xffffffffu, tmin) - (r.substep ? 1 : 0), clo);
            const int iend = min(VL, chi);
            if (imin < iend) acc_is += (uint32_t)(iend - imin) * (uint32_t)TR;

            float mn[TR];                           // running min of the horizontal d2 of each trial
#pragma unroll
            for (int m = 0; m < TR; ++m) mn[m] = 3.0e38f;
            for (int tl = 0; tl < n_tiles; ++tl) {
                const int buf = single ? 0 : (tl & 1);
                if (!single) {
                    // long tracks: stream the tiles through both buffers, one CTA-wide hand-over per tile
                    __syncthreads();
                    if (tid == 0) {
                        fence_proxy_async();
                        const int tn = tl * tile_steps;
                        const uint32_t bn = (uint32_t)((min(VL - tn, tile_steps) + 1) & ~1) * 8u;
                        mbar_arrive_expect_tx(&bars[buf], bn);
                        tma_bulk_g2s(tiles + (size_t)buf * tile_steps, gtrk + (size_t)tn, bn, &bars[buf]);
                    }
                    if (buf == 0) { mbar_wait(&bars[0], phase0); phase0 ^= 1u; }
                    else          { mbar_wait(&bars[1], phase1); phase1 ^= 1u; }
                }
                const int t0 = tl * tile_steps;
                const int lo = max(imin, t0), hi = (DBG && (dbg & 1u)) ? lo : min(iend, t0 + tile_steps);   // dbg 1: no scan
                // LEVEL 1, horizontal screen: d2xy(i) = |Pxy(i) - Axy(i)|^2 <= d2(i), so a trial whose
                // horizontal separation never drops below R + margin cannot collide at any index, whatever
                // the altitudes.  6 packed ops per pair-step instead of 9.  P(lo) = B + lo*S is re-based
                // exactly at every tile start; the rounding accumulated inside a tile (<= tile_steps/2 ulp
                // per coordinate) is inside the screen margin.
                if (lo < hi) {
                    const float lof = (float)lo;
                    const f2_t lo2 = pk2(lof, lof);
                    f2_t PX[NP], PY[NP];
#pragma unroll
                    for (int p = 0; p < NP; ++p) { PX[p] = ffma2(lo2, SX[p], BX[p]); PY[p] = ffma2(lo2, SY[p], BY[p]); }
                    scan_tile_xy<NP>(tiles + (size_t)buf * tile_steps, t0, lo, hi, PX, PY, SX, SY, mn);
                }
            }

            // ---------------- rare path (~1e-3 of trials): LEVEL 2, the full 3-D fp32 distance of the
            // flagged trial alone; what survives goes to the fp64 confirm (LEVEL 3)
#pragma unroll
            for (int m = 0; m < TR; ++m) if (mn[m] < thr) fmask |= 1u << m;
            if (fmask) {
#pragma unroll 1
                for (int m = 0; m < TR; ++m) {
                    if (!((fmask >> m) & 1u)) continue;
                    uint32_t si = sidx[0];
#pragma unroll
                    for (int mm = 1; mm < TR; ++mm) si = (m == mm) ? sidx[mm] : si;
                    const uint4 q = stage[si];
                    if ((q.w >> 12) == 0u) continue;
                    ++npass;
                    const uint64_t trial = r.trial_begin + chunk_first64 + (q.w & 4095u);
                    const Setup32 u = setup_trial32(s, r, *tk, pair, j, s1_first, chunk_first64, q);
                    if (r.decide_fp32) {
                        rescan_fp32(s, r, *tk, pair, a, j, s1_first, trial, u);
                    } else if (u.force ||
                               (r.substep ? min3d_seg_fp32(s, *tk, u) < __ldg(&tk->thr2_l2_sub)
                                          : min3d_fp32(s, *tk, u) < s.thr2_screen)) {
                        ++ncand;
                        const uint32_t slot = warp_claim_slot(r.cand_count);
                        if (slot < r.cand_cap) {
                            Candidate cd; cd.pair = (uint32_t)pair; cd.pad = 0; cd.trial = trial;
                            r.cands[slot] = cd;
                        }
                    }
                }
            }
        }

        c += gridDim.x;                           // next item of this CTA
        if (c >= chunks_per_pair) {
            const uint32_t adv = r.div_chunks.div(c);
            c -= adv * chunks_per_pair;
            po += (int)adv;
            pair = (CULL && (uint32_t)po * chunks_per_pair + c < n_items) ? (int)s.alive[po] : po;
            a = (int)s.div_drones.div((uint32_t)pair);
            j = pair - a * s.n_drones;
        }
    }
    flush_stats(r, 0, acc_t1, acc_is, 0, 0);
    const unsigned long long nc = warp_sum_u64(ncand), np = warp_sum_u64(npass), ncl = warp_sum_u64(acc_cull);
    if (lane == 0 && ncl) atomicAdd(&r.stats[ST_CULLED_TESTS], ncl);
    if (lane == 0 && nc) atomicAdd(&r.stats[ST_CANDIDATES], nc);
    if (lane == 0 && np) atomicAdd(&r.stats[ST_LEVEL2], np);
}

// =========================================================================
// all-pairs stress path (BASELINE config 5): n_drones x ray_count rays, drawn ONCE against a
// reference box / takeoff time, each tested against EVERY track of the scenario.  Not in the
// reference (it only ever pairs one aircraft with one drone's trials); same kinematics, same
// three-level decision (horizontal fp32 -> 3-D fp32 -> fp64 reference arithmetic).
// =========================================================================
constexpr uint32_t AP_TRACK_ID = 0xFFFFFFFFu;   // Philox track id of all-pairs rays (no MC trial uses it)
constexpr int AP_BLOCK = 256;
constexpr int AP_TILE = 512;                    // track samples per shared-memory tile (8 KB)
constexpr int AP_STAGES = 3;

__device__ __forceinline__ Draw ap_draw(const AllPairsDev& ap, uint32_t j, uint32_t ray_local)
{
    return make_draw(ap.seed, ap.ray_begin + ray_local, ap.drone_base + j, AP_TRACK_ID, ap.ref_takeoff_t,
                     ap.ref_box);
}

// pass 1: t1st of every ray, rank within its (drone, t1st) class
__global__ void __launch_bounds__(256) k_ap_count(const ScenarioDev s, const AllPairsDev ap)
{
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)s.n_drones * ap.ray_count) return;
    const uint32_t j = (uint32_t)(idx / ap.ray_count), rl = (uint32_t)(idx - (uint64_t)j * ap.ray_count);
    const U4 w = trial_words(ap.seed, ap.ray_begin + rl, ap.drone_base + j, AP_TRACK_ID);
    const int t1st = draw_t1st(w.x, ap.ref_takeoff_t);
    ap.rank[idx] = atomicAdd(&ap.hist[(size_t)j * ap.ref_takeoff_t + (t1st - 1)], 1u);
}

// pass 2: per drone, exclusive prefix over t1st classes (one warp per drone)
__global__ void __launch_bounds__(32) k_ap_scan(const ScenarioDev s, const AllPairsDev ap)
{
    const int j = blockIdx.x;
    uint32_t carry = 0;
    for (int base = 0; base < ap.ref_takeoff_t; base += 32) {
        const int i = base + (int)threadIdx.x;
        const uint32_t v = i < ap.ref_takeoff_t ? ap.hist[(size_t)j * ap.ref_takeoff_t + i] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
            if ((int)threadIdx.x >= o) inc += u;
        }
        if (i < ap.ref_takeoff_t) ap.hist[(size_t)j * ap.ref_takeoff_t + i] = carry + inc - v;
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
}

// pass 3: ray set-up (the level-1 horizontal ray), written at its sorted slot; threads past
// ray_count fill the dead slots of the last 64-ray group
__global__ void __launch_bounds__(256) k_ap_scatter(const ScenarioDev s, const AllPairsDev ap)
{
    const uint64_t padded = (uint64_t)ap.groups * 64;
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)s.n_drones * padded) return;
    const uint32_t j = (uint32_t)(idx / padded), rl = (uint32_t)(idx - (uint64_t)j * padded);
    uint32_t slot = rl;
    float bx = 1e15f, by = 1e15f, sx = 0.f, sy = 0.f;
    if (rl < ap.ray_count) {
        const U4 w = trial_words(ap.seed, ap.ray_begin + rl, ap.drone_base + j, AP_TRACK_ID);
        const int t1st = draw_t1st(w.x, ap.ref_takeoff_t);
        const int last = t1st - 1;
        const double xl = s.p1x[(size_t)j * s.t_max + last], yl = s.p1y[(size_t)j * s.t_max + last];
        const float2 pl = s.p1f[(size_t)j * s.t_max + last];
        const float dls = (float)fma((double)w.y, ap.wx32, ap.ref_box[0] - xl);
        const float dbs = (float)fma((double)w.z, ap.wy32, ap.ref_box[2] - yl);
        const float ada = fabsf(fmaf((float)w.w, ap.wz32, ap.z0rel));
        const float m2 = fmaf(dls, dls, dbs * dbs);
        const float inv_m = rsqrtf(m2);
        const float k = inv_m * fmaf(-s.coef_f, atanf(ada * inv_m), s.v_straight_f);
        const float lastf = (float)last;
        sx = dls * k; sy = dbs * k;
        bx = fmaf(-lastf, sx, pl.x); by = fmaf(-lastf, sy, pl.y);
        slot = ap.hist[(size_t)j * ap.ref_takeoff_t + last] + ap.rank[(size_t)j * ap.ray_count + rl];
        ap.meta[(size_t)j * ap.ray_count + slot] =
            make_uint2((uint32_t)t1st | (!(m2 > 1e-12f) ? 0x80000000u : 0u), rl);
    }
    const uint32_t g = slot >> 6, m = (slot >> 5) & 1u, ln = slot & 31u;
    const size_t rec = (((size_t)j * ap.groups + g) * 32 + ln) * 4;
    ap.rayA[rec + m] = bx; ap.rayA[rec + 2 + m] = by;
    ap.rayB[rec + m] = sx; ap.rayB[rec + 2 + m] = sy;
}

// The scan: a CTA keeps NP*512 sorted rays of one drone in registers (NP packed pairs per lane) and
// streams the samples of a chunk of tracks through a 3-stage TMA/mbarrier ring.  Per track the
// running horizontal minima are reset, and a (ray, track) pair that comes within R + margin
// horizontally is appended (warp-aggregated) to the level-2 list.
template <int NP>
__global__ void __launch_bounds__(AP_BLOCK, (NP == 1 ? 3 : 2))
k_ap_scan_tracks(const ScenarioDev s, const RunDev r, const AllPairsDev ap, uint32_t n_items)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2*   tiles = reinterpret_cast<float2*>(smem_raw);                        // AP_STAGES x AP_TILE (x, y)
    uint64_t* full  = reinterpret_cast<uint64_t*>(smem_raw + (size_t)AP_STAGES * AP_TILE * 8);
    uint64_t* empty = full + AP_STAGES;
    constexpr int NW = AP_BLOCK / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int b = 0; b < AP_STAGES; ++b) { mbar_init(&full[b], 1); mbar_init(&empty[b], NW); }
        mbar_fence_init();
    }
    __syncthreads();

    uint32_t G = 0;                       // tiles consumed so far by this CTA (ring position, all items)
    unsigned long long ev = 0, is = 0, nl2 = 0;

    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const uint32_t rb = item / ap.n_tchunks, tc = item - rb * ap.n_tchunks;
        const uint32_t j = rb / ap.blocks_per_drone, blk = rb - j * ap.blocks_per_drone;
        const int a_begin = (int)(tc * ap.tracks_per_chunk);
        const int a_end = min(s.n_tracks, a_begin + (int)ap.tracks_per_chunk);

        // ---- this lane's NP packed pairs of rays: 16-byte loads deliver aligned f32x2 pairs
        f2_t BX[NP], BY[NP], SX[NP], SY[NP];
        int t1[2 * NP];
        uint32_t rid[2 * NP];
        uint32_t degen = 0;
        int tmin = INT_MAX;
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const uint32_t g = (blk * NW + (uint32_t)warp) * NP + p;      // 64-ray group
            BX[p] = pk2(1e15f, 1e15f); BY[p] = BX[p]; SX[p] = pk2(0.f, 0.f); SY[p] = SX[p];
            t1[2 * p] = t1[2 * p + 1] = INT_MAX;
            rid[2 * p] = rid[2 * p + 1] = 0xFFFFFFFFu;
            if (g < ap.groups) {
                const size_t rec = ((size_t)j * ap.groups + g) * 32 + lane;
                const ulonglong2 qa = __ldg(reinterpret_cast<const ulonglong2*>(ap.rayA) + rec);
                const ulonglong2 qb = __ldg(reinterpret_cast<const ulonglong2*>(ap.rayB) + rec);
                BX[p] = qa.x; BY[p] = qa.y; SX[p] = qb.x; SY[p] = qb.y;
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    const uint32_t slot = g * 64 + (uint32_t)(m * 32 + lane);
                    if (slot < ap.ray_count) {
                        const size_t o = (size_t)j * ap.ray_count + slot;
                        const uint2 mt = __ldg(&ap.meta[o]);
                        t1[2 * p + m] = (int)(mt.x & 0x7fffffffu);
                        degen |= (mt.x >> 31) << (2 * p + m);
                        rid[2 * p + m] = (uint32_t)o;
                        tmin = min(tmin, t1[2 * p + m]);
                    }
                }
            }
        }
        const int imin = __reduce_min_sync(0xffffffffu, tmin);

        // ---- tile sequence of the item: tracks a_begin..a_end-1, each ceil(n_steps / AP_TILE) tiles.
        // Producer cursor (thread 0) runs AP_STAGES-1 tiles ahead of the consumer cursor (all threads).
        int pa = a_begin, pt = 0;                     // producer: next tile to load
        uint32_t PG = G;                              // its ring position
        auto produce = [&]() {
            if (pa >= a_end) return;
            const int VLp = __ldg(&s.tracks[pa].n_steps);
            const uint32_t off = __ldg(&s.tracks[pa].off_xy);
            const int b = (int)(PG % AP_STAGES);
            if (PG >= (uint32_t)AP_STAGES) mbar_wait(&empty[b], ((PG / AP_STAGES) - 1u) & 1u);
            fence_proxy_async();
            const uint32_t bytes = (uint32_t)((min(AP_TILE, VLp - pt * AP_TILE) + 1) & ~1) * 8u;
            mbar_arrive_expect_tx(&full[b], bytes);
            tma_bulk_g2s(tiles + (size_t)b * AP_TILE, s.trkxy + off + (size_t)pt * AP_TILE, bytes, &full[b]);
            ++PG;
            if ((pt + 1) * AP_TILE >= VLp) { ++pa; pt = 0; } else { ++pt; }
        };
        if (tid == 0) {
#pragma unroll 1
            for (int k = 0; k < AP_STAGES - 1; ++k) produce();
        }

        float mn[2 * NP];
#pragma unroll
        for (int q = 0; q < 2 * NP; ++q) mn[q] = 3.0e38f;
        for (int a = a_begin; a < a_end; ++a) {
            const int VL = __ldg(&s.tracks[a].n_steps);
            const int n_tiles = (VL + AP_TILE - 1) / AP_TILE;
            const int s1f = __ldg(&s.s1_first_hit[(size_t)a * s.n_drones + j]);
            for (int tl = 0; tl < n_tiles; ++tl) {
                if (tid == 0) produce();
                const int b = (int)(G % AP_STAGES);
                mbar_wait(&full[b], (G / AP_STAGES) & 1u);
                const int t0 = tl * AP_TILE;
                const int lo = max(imin, t0), hi = min(VL, t0 + AP_TILE);
                if (lo < hi) {
                    const float lof = (float)lo;
                    const f2_t lo2 = pk2(lof, lof);
                    f2_t PX[NP], PY[NP];
#pragma unroll
                    for (int p = 0; p < NP; ++p) {                 // exact re-base per tile
                        PX[p] = ffma2(lo2, SX[p], BX[p]); PY[p] = ffma2(lo2, SY[p], BY[p]);
                    }
                    scan_tile_xy<NP>(tiles + (size_t)b * AP_TILE, t0, lo, hi, PX, PY, SX, SY, mn);
                    is += (unsigned long long)(hi - lo) * (2ull * NP);
                }
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[b])) : "memory");
                ++G;
            }
            // ---- end of track a: bookkeeping + rare level-2 append, then reset the minima
#pragma unroll
            for (int q = 0; q < 2 * NP; ++q) {
                if (t1[q] != INT_MAX) {
                    ev += (unsigned long long)max(0, VL - t1[q]);
                    if (mn[q] < s.thr2_screen || ((degen >> q) & 1u) || s1f <= t1[q] - 1) {
                        ++nl2;
                        const uint32_t slot = warp_claim_slot(ap.l2_count);
                        if (slot < ap.l2_cap) ap.l2[slot] = make_uint2((uint32_t)a, rid[q]);
                    }
                }
                mn[q] = 3.0e38f;
            }
        }
    }
    ev = warp_sum_u64(ev); is = warp_sum_u64(is); nl2 = warp_sum_u64(nl2);
    if (lane == 0) {
        if (ev) atomicAdd(&r.stats[ST_EVALUATED], ev);
        if (is) atomicAdd(&r.stats[ST_ISSUED], is);
        if (nl2) atomicAdd(&r.stats[ST_LEVEL2], nl2);
    }
}

// Levels 2 and 3 of the all-pairs path: one thread per (track, ray) entry of the level-2 list.
__global__ void __launch_bounds__(128) k_ap_resolve(const ScenarioDev s, const RunDev r, const AllPairsDev ap)
{
    const uint32_t n = min(*ap.l2_count, ap.l2_cap);
    unsigned long long hits = 0, ncand = 0;
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
        const uint2 e = ap.l2[idx];
        const int a = (int)e.x;
        const uint32_t j = (uint32_t)(e.y / ap.ray_count);
        const uint2 mt = ap.meta[e.y];
        const TrackDev t = s.tracks[a];
        const int pair = a * s.n_drones + (int)j;
        const Draw d = ap_draw(ap, j, mt.y);
        const int last = d.t1st - 1;
        const int s1f = s.s1_first_hit[pair];
        bool go = (mt.x >> 31) != 0u || s1f <= last;
        if (!go) {                                    // level 2: 3-D fp32, fused form
            const double xl = s.p1x[(size_t)j * s.t_max + last], yl = s.p1y[(size_t)j * s.t_max + last];
            Setup32 u;
            u.t1st = d.t1st;
            u.ray = stage2_ray32((float)DSUB(d.tx, xl), (float)DSUB(d.ty, yl), (float)DSUB(d.tz, s.start_alt),
                                 (float)DSUB(xl, s.ox), (float)DSUB(yl, s.oy), (float)DSUB(s.start_alt, s.oz),
                                 (float)last, s.v_straight_f, s.coef_f);
            go = u.ray.degenerate || min3d_fp32(s, t, u) < s.thr2_screen;
        }
        if (!go) continue;
        ++ncand;                                      // level 3: fp64, reference order of operations
        const TrialOut o = exact_trial(s, pair, (int)j, t.n_steps, s.ax + t.off, s.ay + t.off, s.az + t.off, d);
        if (o.first >= 0) {
            ++hits;
            publish_hit(s, r, pair, a, (int)j, ap.ray_begin + mt.y, d, o);
        }
    }
    hits = warp_sum_u64(hits); ncand = warp_sum_u64(ncand);
    if (lane_id() == 0) {
        if (hits) atomicAdd(&r.stats[ST_HITS], hits);
        if (ncand) atomicAdd(&r.stats[ST_CANDIDATES], ncand);
    }
}

// =========================================================================
// replay table dump, trajectories
// =========================================================================
__global__ void __launch_bounds__(256) k_draws(const ScenarioDev s, int a, uint32_t track_id,
                                               uint32_t drone_id, uint64_t seed, uint64_t trial_begin,
                                               uint64_t n, Draw* out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const TrackDev t = s.tracks[a];
    out[i] = make_draw(seed, trial_begin + i, drone_id, track_id, t.takeoff_t, t.box);
}

// One thread per record; xyz layout: record r -> x[0..n) y[0..n) z[0..n) at r*3*stride.
__global__ void __launch_bounds__(64) k_trajectory(const ScenarioDev s, const Record* recs, uint32_t n,
                                                   double* xyz, int stride)
{
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const Record rc = recs[idx];
    const TrackDev t = s.tracks[rc.track];
    double* X = xyz + (size_t)idx * 3 * stride;
    double* Y = X + stride;
    double* Z = Y + stride;
    const int j = (int)rc.drone;
    for (int i = 0; i < rc.t1st && i < t.n_steps; ++i) {       // stage 1, Drone.cpp:120-126
        X[i] = s.p1x[(size_t)j * s.t_max + i];
        Y[i] = s.p1y[(size_t)j * s.t_max + i];
        Z[i] = s.start_alt;
    }
    const int last = rc.t1st - 1;
    const Ray64 r = stage2_ray(X[last], Y[last], s.start_alt, rc.tx, rc.ty, rc.tz, s.v_straight, s.v_ascend);
    double x = r.x, y = r.y, z = r.z;
    for (int i = rc.t1st; i < t.n_steps; ++i) {                // stage 2, Drone.cpp:224-230
        x = DADD(x, r.sx); y = DADD(y, r.sy); z = DADD(z, r.sz);
        X[i] = x; Y[i] = y; Z[i] = z;
    }
}

// Clears a run's output arena: words [0, first_off) = 0 (counters, stats, counts),
// words [first_off, n) = INT32_MAX (first-conflict index "none").
// ext / n_ext: a caller-owned counts buffer to clear as well (DCRP_FLAG_ZERO_COUNTS), or nullptr / 0.
__global__ void __launch_bounds__(256) k_run_init(uint32_t* out, uint32_t first_off, uint32_t n,
                                                  unsigned long long* ext, uint32_t n_ext)
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // a dependent k_screen3 may start its first item now
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (i < first_off) ? 0u : 0x7fffffffu;
    if (i < n_ext) ext[i] = 0ull;
}

}  // namespace dcrp
