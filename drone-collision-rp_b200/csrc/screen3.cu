// screen3.cu -- k_screen3: the Monte-Carlo screen with a 1-D slab level in front of the 2-D scan.
//
// Every (trial, index) separation test is first evaluated along ONE horizontal direction n chosen per
// (track, drone) pair: |D(i) . n| <= |D_xy(i)| <= |D(i)|, so a trial whose projected separation never
// drops below R + margin cannot collide at any index.  That costs two packed adds and half an FMNMX3
// per two tests instead of six packed ops, and with a well-chosen n it clears 90-99 % of the trials
// (5 km ring: 99 %); the survivors queue up per warp and are re-scanned 128 at a time by the 2-D
// level-1 scan of kernels.cu (scan_tile_xy), then level 2 (3-D fp32) and level 3 (the fp64 decision in the
// reference's order of operations), one flagged trial at a time by the whole warp -- all inside this one kernel.
// Decisions are those of the all-fp64 kernel by construction: every level is conservative and the last
// one is exact (tests/test_gpu_parity.py, test_gpu_fuzz.py).
//
// No sort: the slab test of an index before the trial's kink (i < t1st) evaluates the stage-2 ray
// extrapolated backwards, which can only ADD survivors, so all lanes scan the whole track from index 0
// with warp-uniform bounds and nothing is staged in shared memory per trial.  Warps are independent
// (no CTA barrier in steady state): each runs 256-trial items (8 trials per lane), taken in blocks of
// consecutive items -- a static first block, then blocks numbered by a device-side counter whose size tapers
// from 32 items to 1 towards the end of the launch (S3Sched) -- keeps the track (one 1-D TMA bulk copy per
// track change, its own mbarrier), its pair's projected track and kink tables in its own shared-memory slice
// (re-staged only when the pair changes) and keeps its own survivor queue.  The grid is launched
// programmatically dependent on the launch that clears the output arena (griddep_wait below), or alone when
// the previous run left the arena cleared (k_publish, at the end of this file).
//
// Reference semantics: /root/reference/Drone.cpp:107-127,164-231,253-265 (restated in dcrp_math.cuh).
#include "dcrp_device.cuh"

namespace dcrp {

constexpr int S3_NPL = 4;                 // packed pairs per lane in the slab phase
constexpr int S3_TL = 2 * S3_NPL;         // trials per lane per round
constexpr int S3_RND = 32 * S3_TL;        // trials per warp round
constexpr int S3_R = 1;                   // rounds per item
constexpr int S3_CHW = S3_RND * S3_R;     // trials per warp item
constexpr int S3_LG_GMAX = 5;             // largest block of the work counter: 32 consecutive items (same pair, same track)
constexpr int S3_RESCAN = 128;            // survivors per 2-D re-scan round (4 per lane = 2 packed pairs)
constexpr int S3_LIST = S3_RESCAN + S3_RND;
constexpr int S3_BLOCK = 256;
constexpr int S3_MAX_VL = 512;            // longest track the per-warp projected table holds
constexpr int S3_MAX_T = 128;             // largest takeoff_t the per-warp kink tables hold
constexpr int S3_REBASE = 128;            // exact re-base of the running position every so many indices
constexpr int S3_DIRS = 16;               // candidate slab directions k*pi/16
constexpr int S3_L2_SERIAL = 6;           // sub-step mode: up to this many flagged trials of a re-scan round are taken one at a time by the warp

static_assert(S3_CHW <= 4096, "local trial index is packed in 12 bits");

// How the work counter deals out the items behind the static first blocks (host-built, engine.cu s3_schedule):
// block k of the launch is items [item0[p] + ((k - blk0[p]) << lg[p]), + 2^lg[p]) of the phase p that holds k.
// Block sizes taper 32, 16, ... 1 items ("factoring": a size is used while twice a full round of it is left), as a
// function of the block NUMBER, so a warp that has not looked at the counter for a long block cannot overdraw.
constexpr int S3_MAX_PH = 6;
struct S3Sched {
    uint32_t n_ph;
    uint32_t item0[S3_MAX_PH + 1];          // first item of phase p; [n_ph] = n_items
    uint32_t blk0[S3_MAX_PH + 1];           // first block number of phase p; [n_ph] = number of blocks
    uint32_t lg[S3_MAX_PH];                 // log2 of the phase's block size
};

__host__ __device__ inline size_t s3_warp_smem(int vl_pad, int tk_pad)
{
    // [tile: vl_pad float2 (the track's (x, y) samples, one TMA bulk copy per track change)][tabk: tk_pad double2]
    // [tabT: vl_pad float][list: S3_LIST uint2][tabpn: tk_pad float][mbarrier + its phase: 16 B], 16-B multiple
    size_t b = (size_t)vl_pad * 8 + (size_t)tk_pad * 16 + (size_t)vl_pad * 4 + (size_t)S3_LIST * 8 + (size_t)tk_pad * 4 + 16;
    return (b + 15) & ~(size_t)15;
}

// Slab direction of every pair, once per scenario (k_stage1 has already filled in the per-track default:
// the aircraft's main direction of travel): of S3_DIRS candidate directions k*pi/S3_DIRS the one whose slab
// passes the fewest of S3_AXIS_SAMPLES deterministic sample trials (at 1.5x the real threshold: enough
// passes to rank the directions, still close to the quantity that matters).  A heuristic only: any unit
// vector gives the same decisions, just more or fewer survivors for the 2-D re-scan.
// One CTA of S3_AXIS_BLOCK threads per pair, S3_AXIS_TL sample trials per thread.
constexpr int S3_AXIS_BLOCK = 512;
constexpr int S3_AXIS_TL = 2;
constexpr int S3_AXIS_SAMPLES = S3_AXIS_BLOCK * S3_AXIS_TL;

__global__ void __launch_bounds__(S3_AXIS_BLOCK) k_slab_axis(const ScenarioDev s, float thr_rank)
{
    __shared__ uint32_t cnt[S3_DIRS];
    __shared__ float2 dirs[S3_DIRS];
    __shared__ float2 trk[S3_MAX_VL];
    const int pair = (int)blockIdx.x;
    const int lane = threadIdx.x & 31;
    const int a = (int)s.div_drones.div((uint32_t)pair), j = pair - a * s.n_drones;
    const TrackDev* __restrict__ tk = s.tracks + a;
    const int VL = tk->n_steps, takeoff_t = tk->takeoff_t;
    if (threadIdx.x < S3_DIRS) {
        cnt[threadIdx.x] = 0;
        float sn, cs;
        sincospif((float)threadIdx.x / (float)S3_DIRS, &sn, &cs);
        dirs[threadIdx.x] = make_float2(cs, sn);
    }
    for (int i = threadIdx.x; i < VL; i += S3_AXIS_BLOCK) trk[i] = __ldg(&s.trkxy[(size_t)tk->off_xy + i]);
    __syncthreads();
    TrackScale ts;
    ts.wx32 = tk->wx32; ts.wy32 = tk->wy32; ts.wz32 = tk->wz32; ts.z0rel = tk->z0rel;
    uint32_t bits[S3_AXIS_TL];
#pragma unroll 1
    for (int m = 0; m < S3_AXIS_TL; ++m) {
        const U4 w = philox4x32_10((uint32_t)(m * S3_AXIS_BLOCK + threadIdx.x), 0u, (uint32_t)j, (uint32_t)a, 0xA511E9B3u,
                                   0x5AB5A5E1u);
        const int t1st = draw_t1st(w.x, takeoff_t);
        const uint4 q = make_uint4(w.y, w.z, w.w, (uint32_t)t1st << 12);
        const Fast32 u = setup_fast32(s, ts, s.kink + (size_t)pair * s.t_max, s.p1f + (size_t)j * s.t_max, INT_MAX, q);
        uint32_t b = 0;
        if (!u.force) {
            for (int i = 1; i < VL; ++i) {
                const float2 p = trk[i];
                const float fi = (float)i;
                const float dx = fmaf(fi, u.sx, u.bx) + p.x, dy = fmaf(fi, u.sy, u.by) + p.y;
#pragma unroll
                for (int k = 0; k < S3_DIRS; ++k)
                    if (fabsf(fmaf(dirs[k].x, dx, dirs[k].y * dy)) < thr_rank) b |= 1u << k;
            }
        }
        bits[m] = b;
    }
#pragma unroll
    for (int k = 0; k < S3_DIRS; ++k) {
        uint32_t c = 0;
#pragma unroll
        for (int m = 0; m < S3_AXIS_TL; ++m) c += (uint32_t)__popc(__ballot_sync(0xffffffffu, (bits[m] >> k) & 1u));
        if (lane == 0 && c) atomicAdd(&cnt[k], c);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int best = 0;
        for (int k = 1; k < S3_DIRS; ++k) if (cnt[k] < cnt[best]) best = k;
        s.slab_dir[pair] = dirs[best];
    }
}

// First stage-1 conflict of a pair as "every trial whose kink index is >= this must be confirmed": a sample index in
// sample mode, (segment index + 1) in sub-step mode (as in k_screen2).
template <bool SUB>
__device__ __forceinline__ int s3_s1_first(const ScenarioDev& s, int pair)
{
    if (!SUB) return __ldg(&s.s1_first_hit[pair]);
    const int sg = __ldg(&s.s1sub_first_seg[pair]);
    return sg == INT_MAX ? INT_MAX : sg + 1;
}

// 2-D level-1 re-scan of n <= S3_RESCAN queued survivors of track a (4 per lane as two packed pairs);
// the draws are regenerated from the trial id (nothing but (item, local) was kept), the track is the warp's
// own shared-memory tile (the queue is drained before the tile is reloaded for another track).
// NP packed pairs per lane: 2 for a full round of S3_RESCAN survivors, 1 for the up to 64 left over when a warp
// changes track or ends (on the 5 km ring a warp queues a dozen survivors per launch: that one call is its whole
// re-scan work, and it stands at the very end of the launch).
// SUB (DCRP_FLAG_SUBSTEP, the interpolated-separation extension): the scan also takes the kink sample (the first end
// of the first stage-2 segment), level 1b and level 2 use the per-track thresholds widened by half a relative step
// (engine.cu), level 2 is the fp32 closest approach per segment, level 3 (exact_any_warp) follows r.substep.
template <int NP, bool SUB>
__device__ __noinline__ void s3_rescan_np(const ScenarioDev& s, const RunDev& r, const uint2* list, int n, int a,
                                          uint32_t chunks_per_pair, const float2* __restrict__ tile)
{
    const int lane = (int)lane_id();
    uint32_t acc_is = 0, npass = 0, ncand = 0;
    const TrackDev* __restrict__ tk = s.tracks + a;
    const int VL = __ldg(&tk->n_steps), takeoff_t = __ldg(&tk->takeoff_t);
    TrackScale ts;
    ts.wx32 = __ldg(&tk->wx32); ts.wy32 = __ldg(&tk->wy32); ts.wz32 = __ldg(&tk->wz32); ts.z0rel = __ldg(&tk->z0rel);
    constexpr int TR = 2 * NP;
    float bxs[TR], bys[TR], sxs[TR], sys[TR];
    uint32_t fmask = 0;
    int tmin = INT_MAX;
#pragma unroll
    for (int m = 0; m < TR; ++m) {
        const int idx = m * 32 + lane;
        float bx = 1e15f, by = 1e15f, sx = 0.f, sy = 0.f;
        if (idx < n) {
            DCRP_ASSERT(n <= 32 * TR);
            const uint2 e = list[idx];
            const int pair = (int)r.div_chunks.div(e.x);
            DCRP_ASSERT(pair >= a * s.n_drones && pair < (a + 1) * s.n_drones && e.y < (uint32_t)S3_CHW);
            const uint32_t c = e.x - (uint32_t)pair * chunks_per_pair;
            const int j = pair - a * s.n_drones;
            const uint64_t trial = r.trial_begin + r.slice_first + (uint64_t)c * S3_CHW + e.y;
            const U4 w = philox4x32_10_rk((uint32_t)trial, (uint32_t)(trial >> 32), r.drone_base + (uint32_t)j,
                                          r.track_base + (uint32_t)a, r.rk);
            const int t1st = draw_t1st(w.x, takeoff_t);
            const uint4 q = make_uint4(w.y, w.z, w.w, e.y | ((uint32_t)t1st << 12));
            const Fast32 u = setup_fast32(s, ts, s.kink + (size_t)pair * s.t_max, s.p1f + (size_t)j * s.t_max,
                                          s3_s1_first<SUB>(s, pair), q);
            bx = u.bx; by = u.by; sx = u.sx; sy = u.sy;
            fmask |= u.force ? (1u << m) : 0u;
            tmin = min(tmin, t1st);
        }
        bxs[m] = bx; bys[m] = by; sxs[m] = sx; sys[m] = sy;
    }
    f2_t BX[NP], BY[NP], SX[NP], SY[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        BX[p] = pk2(bxs[2 * p], bxs[2 * p + 1]); BY[p] = pk2(bys[2 * p], bys[2 * p + 1]);
        SX[p] = pk2(sxs[2 * p], sxs[2 * p + 1]); SY[p] = pk2(sys[2 * p], sys[2 * p + 1]);
    }
    const int imin = __reduce_min_sync(0xffffffffu, tmin) - (SUB ? 1 : 0);
    if (imin < VL) acc_is += (uint32_t)(VL - imin) * (uint32_t)TR;
    float mn[TR];
#pragma unroll
    for (int m = 0; m < TR; ++m) mn[m] = 3.0e38f;
    for (int t0 = 0; t0 < VL; t0 += S3_REBASE) {
        const int lo = max(imin, t0), hi = min(VL, t0 + S3_REBASE);
        if (lo < hi) {
            const float lof = (float)lo;
            const f2_t lo2 = pk2(lof, lof);
            f2_t PX[NP], PY[NP];
#pragma unroll
            for (int p = 0; p < NP; ++p) { PX[p] = ffma2(lo2, SX[p], BX[p]); PY[p] = ffma2(lo2, SY[p], BY[p]); }
            scan_tile_xy<NP>(tile + t0, t0, lo, hi, PX, PY, SX, SY, mn);
        }
    }
    const float thr1 = SUB ? __ldg(&tk->thr2_l1_sub) : s.thr2_screen;
#pragma unroll
    for (int m = 0; m < TR; ++m) if (mn[m] < thr1) fmask |= 1u << m;
    // ---- rare: LEVEL 2 (3-D fp32, fused form) of the flagged trials, then LEVEL 3 (fp64) of what survives -- one
    // trial at a time by the WHOLE warp (every lane re-derives the trial from its queue entry, the indices are dealt
    // out over the lanes: min3d_fp32_warp, exact_any_warp).  Same decisions and bits as one lane per trial; what
    // changes is the latency: this code also runs when a warp empties its queue at the very end of a launch, where
    // a lane alone on a 96-index fp32 scan (2.5 us) or an fp64 trial (20 us) was the tail of the whole kernel.
    // Candidates are ~3e-7 of the trials: deciding them here saves the candidate list, a kernel launch and its
    // ~10 us of latency per run.
    uint32_t hits = 0;
    if (__any_sync(0xffffffffu, fmask != 0u)) {
        const TrackDev t = *tk;
        // Sub-step mode: the level-1 radius is wider by half a relative step (tens of metres), so next to the flight
        // path a round of survivors can hold dozens of flagged trials, and taking those one at a time -- ~1500 cycles of
        // latency each -- made single 256-trial items last tens of microseconds (the tail of the launch).  A round with
        // more than S3_L2_SERIAL flagged trials therefore runs level 2 with one LANE per trial (min3d_seg_fp32, as
        // k_screen2 does), and only what level 2 passes goes through the warp-wide level 3 below.
        bool l2_done = false;
        if (SUB) {
            int nflag = 0;
#pragma unroll
            for (int m = 0; m < TR; ++m) nflag += __popc(__ballot_sync(0xffffffffu, ((fmask >> m) & 1u) && m * 32 + lane < n));
            if (nflag > S3_L2_SERIAL) {
#pragma unroll 1
                for (int m = 0; m < TR; ++m) {
                    if (!((fmask >> m) & 1u) || m * 32 + lane >= n) continue;
                    const uint2 e = list[m * 32 + lane];
                    const int pair = (int)r.div_chunks.div(e.x);
                    const uint32_t c = e.x - (uint32_t)pair * chunks_per_pair;
                    const int j = pair - a * s.n_drones;
                    const uint64_t trial = r.trial_begin + r.slice_first + (uint64_t)c * S3_CHW + e.y;
                    const U4 w = philox4x32_10_rk((uint32_t)trial, (uint32_t)(trial >> 32), r.drone_base + (uint32_t)j,
                                                  r.track_base + (uint32_t)a, r.rk);
                    const int t1st = draw_t1st(w.x, takeoff_t);
                    const uint4 q = make_uint4(w.y, w.z, w.w, e.y | ((uint32_t)t1st << 12));
                    const Setup32 u = setup_trial32(s, r, t, pair, j, s3_s1_first<SUB>(s, pair), 0, q);
                    const bool cand = u.force || min3d_seg_fp32(s, t, u) < t.thr2_l2_sub;
                    ++npass;
                    if (cand) ++ncand; else fmask &= ~(1u << m);
                }
                __syncwarp();
                l2_done = true;
            }
        }
#pragma unroll 1
        for (int m = 0; m < TR; ++m) {
            uint32_t todo = __ballot_sync(0xffffffffu, ((fmask >> m) & 1u) && m * 32 + lane < n);
            while (todo) {
                const int src = __ffs((int)todo) - 1;
                todo &= todo - 1u;
                const uint2 e = list[m * 32 + src];                       // the same entry for every lane
                const int pair = (int)r.div_chunks.div(e.x);
                const uint32_t c = e.x - (uint32_t)pair * chunks_per_pair;
                const int j = pair - a * s.n_drones;
                const uint64_t trial = r.trial_begin + r.slice_first + (uint64_t)c * S3_CHW + e.y;
                const U4 w = philox4x32_10_rk((uint32_t)trial, (uint32_t)(trial >> 32), r.drone_base + (uint32_t)j,
                                              r.track_base + (uint32_t)a, r.rk);
                const int t1st = draw_t1st(w.x, takeoff_t);
                const uint4 q = make_uint4(w.y, w.z, w.w, e.y | ((uint32_t)t1st << 12));
                bool cand = true;                                         // (level 2 already done lane by lane)
                if (!(SUB && l2_done)) {
                    const Setup32 u = setup_trial32(s, r, t, pair, j, s3_s1_first<SUB>(s, pair), 0, q);
                    cand = u.force || (SUB ? min3d_seg_fp32_warp(s, t, u, lane) < t.thr2_l2_sub
                                           : min3d_fp32_warp(s, t, u, lane) < s.thr2_screen);
                    if (lane == 0) { ++npass; ncand += cand ? 1u : 0u; }
                }
                if (cand) {
                    const Draw d = make_draw(r.seed, trial, r.drone_base + (uint32_t)j, r.track_base + (uint32_t)a,
                                             t.takeoff_t, t.box);
                    const TrialOut o = exact_any_warp(s, r, pair, j, t.n_steps, s.ax + t.off, s.ay + t.off, s.az + t.off, d, lane);
                    if (o.first >= 0 && lane == 0) {
                        ++hits;
                        publish_hit(s, r, pair, a, j, trial, d, o);
                    }
                }
            }
        }
    }
    // a rare path (one call per 128 survivors): its own warp-reduced statistics
    const unsigned long long is = warp_sum_u64(acc_is), np = warp_sum_u64(npass), nc = warp_sum_u64(ncand),
                             nh = warp_sum_u64(hits);
    if (lane == 0) {
        atomicAdd(&r.stats[ST_ISSUED], is);
        if (np) atomicAdd(&r.stats[ST_LEVEL2], np);
        if (nc) atomicAdd(&r.stats[ST_CANDIDATES], nc);
        if (nh) atomicAdd(&r.stats[ST_HITS], nh);
    }
}

template <bool SUB>
__device__ __forceinline__ void s3_rescan(const ScenarioDev& s, const RunDev& r, const uint2* list, int n, int a,
                                          uint32_t chunks_per_pair, const float2* __restrict__ tile)
{
    if (n <= 64) s3_rescan_np<1, SUB>(s, r, list, n, a, chunks_per_pair, tile);
    else         s3_rescan_np<2, SUB>(s, r, list, n, a, chunks_per_pair, tile);
}

// griddepcontrol.wait: returns once the grid this one was programmatically launched behind has completed and its
// writes are visible (immediately when the launch carried no such dependency).
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

#ifdef DCRP_TIMELINE
// Opt-in instrumentation (make EXTRA=-DDCRP_TIMELINE, tools/warp_timeline.py): %globaltimer of every warp at entry,
// after its first staging, after its first item, when it leaves the item loop and at its end, plus its item count,
// its SM and the start of its last item.
constexpr int TL_WARPS = 8192, TL_SLOTS = 8;
__device__ unsigned long long g_timeline[TL_WARPS * TL_SLOTS];
__device__ __forceinline__ unsigned long long tl_now()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define TL_MARK(slot)                                                                             \
    do {                                                                                          \
        const uint32_t gw_ = blockIdx.x * (S3_BLOCK / 32) + (threadIdx.x >> 5);                   \
        if ((threadIdx.x & 31) == 0 && gw_ < (uint32_t)TL_WARPS) g_timeline[gw_ * TL_SLOTS + (slot)] = tl_now(); \
    } while (0)
#define TL_COUNT(n)                                                                               \
    do {                                                                                          \
        const uint32_t gw_ = blockIdx.x * (S3_BLOCK / 32) + (threadIdx.x >> 5);                   \
        if ((threadIdx.x & 31) == 0 && gw_ < (uint32_t)TL_WARPS) g_timeline[gw_ * TL_SLOTS + 5] = (n); \
    } while (0)
#else
#define TL_MARK(slot) do {} while (0)
#define TL_COUNT(n) do {} while (0)
#endif

template <int MINB, bool SUB>
__global__ void __launch_bounds__(S3_BLOCK, MINB)
k_screen3(const __grid_constant__ ScenarioDev s, const __grid_constant__ RunDev r, uint32_t chunks_per_pair, uint32_t n_items, int vl_pad, int tk_pad,
          uint32_t first_block, const __grid_constant__ S3Sched sch)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* wb = smem_raw + (size_t)warp * s3_warp_smem(vl_pad, tk_pad);
    float2*  tile  = reinterpret_cast<float2*>(wb);                                   // the warp's current track, (-x', -y')
    double2* tabk  = reinterpret_cast<double2*>(wb + (size_t)vl_pad * 8);
    float*   tabT  = reinterpret_cast<float*>(wb + (size_t)vl_pad * 8 + (size_t)tk_pad * 16);
    uint2*   list  = reinterpret_cast<uint2*>(wb + (size_t)vl_pad * 12 + (size_t)tk_pad * 16);
    float*   tabpn = reinterpret_cast<float*>(wb + (size_t)vl_pad * 12 + (size_t)tk_pad * 16 + (size_t)S3_LIST * 8);
    uint64_t* bar  = reinterpret_cast<uint64_t*>(wb + (size_t)vl_pad * 12 + (size_t)tk_pad * 20 + (size_t)S3_LIST * 8);
    TL_MARK(0);
    if (lane == 0) { mbar_init(bar, 1); mbar_fence_init(); bar[1] = 0ull; }     // bar[1]: the barrier's phase bit
    __syncwarp();

    uint32_t acc_t1 = 0, acc_is = 0, nsurv = 0;       // per-lane partial sums
    int list_n = 0, list_track = -1;       // warp-uniform

    // Work items are S3_CHW consecutive trials of one pair, numbered pair-major.  A warp takes them in BLOCKS of
    // consecutive items (same pair, same track: tables and track are staged only when they change, and the survivor
    // queue fills up across the items of a block before it is drained).  The first block of every warp is static --
    // items [w * first_block, (w + 1) * first_block), no round trip to the one L2 line all warps would queue on at
    // launch -- the later ones are numbered by the device counter and sized by the schedule `sch` (big blocks while
    // there is plenty, single items at the end of the launch, where the finish times of the warps decide the tail).
    uint32_t pend_v = 0;                   // the grab in flight: lane 0 holds the counter's answer until the block is needed
    auto grab_issue = [&]() {
        if (sch.n_ph == 0u) return;        // a launch dealt out statically: grab_take answers "nothing left"
        if (lane == 0) pend_v = atomicAdd(r.work_counter, 1u);
    };
    auto grab_take = [&](uint32_t& b, uint32_t& e) {
        const uint32_t k = __shfl_sync(0xffffffffu, pend_v, 0);
        b = e = n_items;                   // (the counter has run out)
#pragma unroll
        for (int p = 0; p < S3_MAX_PH; ++p)
            if ((uint32_t)p < sch.n_ph && k >= sch.blk0[p] && k < sch.blk0[p + 1]) {
                b = sch.item0[p] + ((k - sch.blk0[p]) << sch.lg[p]);
                e = min(b + (1u << sch.lg[p]), sch.item0[p + 1]);
            }
    };

    // Programmatic dependent launch: this grid may start while the launch that clears the output arena (k_run_init) is
    // still running.  Tables, draws and the slab scan of the first item touch nothing it writes; griddep_wait() stands
    // before the first access to the arena (the work counter, and whatever s3_rescan publishes).
    uint32_t item = (blockIdx.x * (S3_BLOCK / 32) + (uint32_t)warp) * first_block;
    uint32_t blk_end = min(item + first_block, n_items);
    bool waited = false, prefetch = true;
    int cur_pair = -1;
#ifdef DCRP_TIMELINE
    unsigned long long tl_items = 0;
#endif
    while (item < n_items) {
        const bool last_of_block = item + 1 == blk_end;
        // the next block is asked for while the last item of a BIG block is processed; towards the end of the launch
        // (blocks of one or two items) only when the warp is ready for it, so that no warp sits on unstarted work
        // that a warp with nothing left to do could have taken
        if (last_of_block && waited && prefetch) grab_issue();
        const int pair = (int)r.div_chunks.div(item);
        const uint32_t c = item - (uint32_t)pair * chunks_per_pair;
        const int a = (int)s.div_drones.div((uint32_t)pair), j = pair - a * s.n_drones;
        const TrackDev* __restrict__ tk = s.tracks + a;
        const int VL = __ldg(&tk->n_steps), takeoff_t = __ldg(&tk->takeoff_t);
        bool tile_wait = false;
        uint32_t phase = 0;
        if (a != list_track) {                    // (once per run for a one-track scenario)
            // queued survivors share one track: drain them before it changes
            if (list_n) { griddep_wait(); s3_rescan<SUB>(s, r, list, list_n, list_track, chunks_per_pair, tile); list_n = 0; }
            list_track = a;
            // the track's samples come in by ONE 1-D TMA bulk copy (waited for below, behind the loads that do not
            // need them); the per-pair projection on the pair's direction then reads shared memory, not L2
            __syncwarp();
            phase = (uint32_t)bar[1];
            if (lane == 0) {
                const uint32_t bytes = (uint32_t)((VL + 1) & ~1) * 8u;              // tracks are padded to an even length
                DCRP_ASSERT(bytes <= (uint32_t)vl_pad * 8u);
                fence_proxy_async();
                mbar_arrive_expect_tx(bar, bytes);
                tma_bulk_g2s(tile, s.trkxy + (size_t)__ldg(&tk->off_xy), bytes, bar);
            }
            tile_wait = true;
        }
        const uint32_t n_valid = (uint32_t)min((uint64_t)S3_CHW, r.slice_count - (uint64_t)c * S3_CHW);
        const int s1_first = s3_s1_first<SUB>(s, pair);
        const float thr_slab = SUB ? __ldg(&tk->thr_slab_sub) : s.thr_slab;
        const float2 nd = __ldg(&s.slab_dir[pair]);
        TrackScale ts;
        ts.wx32 = __ldg(&tk->wx32); ts.wy32 = __ldg(&tk->wy32); ts.wz32 = __ldg(&tk->wz32); ts.z0rel = __ldg(&tk->z0rel);

        // ---- this pair's tables, shared by the items of a block: kink tables (target-box corner minus kink in fp64;
        // kink projected on n) first -- they do not need the track -- then the projected track T_i = n . (-A_i)
        // (index 0 and the padding never pass)
        const bool new_pair = pair != cur_pair;
        if (new_pair) {
            __syncwarp();
            for (int t = lane; t < takeoff_t; t += 32) {
                DCRP_ASSERT(t < tk_pad && pair < s.n_tracks * s.n_drones && takeoff_t <= s.t_max);
                tabk[t] = __ldg(&s.kink[(size_t)pair * s.t_max + t]);
                const float2 pl = __ldg(&s.p1f[(size_t)j * s.t_max + t]);
                tabpn[t] = fmaf(nd.x, pl.x, nd.y * pl.y);
            }
        }
        if (tile_wait) {
            mbar_wait(bar, phase);
            __syncwarp();
            if (lane == 0) bar[1] = phase ^ 1u;
        }
        if (new_pair) {
            cur_pair = pair;
            const int vlr = (VL + 3) & ~3;
            for (int i = lane; i < vlr; i += 32) {
                float v = 1e15f;
                // (index 0 is never a stage-2 sample; in sub-step mode it can be the first end of the first segment)
                if (i >= (SUB ? 0 : 1) && i < VL) { const float2 p = tile[i]; v = fmaf(nd.x, p.x, nd.y * p.y); }
                DCRP_ASSERT(i < vl_pad);
                tabT[i] = v;
            }
            __syncwarp();
        }

#ifdef DCRP_TIMELINE
        if (!waited) TL_MARK(1);
        TL_MARK(7);
        ++tl_items;
#endif
        uint32_t t1s = 0;                          // item-local: the run-long accumulators are touched once per item
        const uint64_t trial0 = r.trial_begin + r.slice_first + (uint64_t)c * S3_CHW + (uint32_t)lane;
        const uint32_t t0lo = (uint32_t)trial0, t0hi = (uint32_t)(trial0 >> 32);
        const uint32_t pj = r.drone_base + (uint32_t)j, pa = r.track_base + (uint32_t)a;
#pragma unroll 1
        for (uint32_t rbase = 0; rbase < n_valid; rbase += S3_RND) {
        // ---- draws + slab ray of this lane's S3_TL trials of the round.  Branch-free: the slots past a ragged last
        // item compute a (harmless) trial beyond the range and are masked afterwards.
        float bn[S3_TL], sn[S3_TL];
        uint32_t fmask = 0;
#pragma unroll
        for (int m = 0; m < S3_TL; ++m) {
            const bool valid = rbase + (uint32_t)(m * 32 + lane) < n_valid;
            const uint32_t c0 = t0lo + rbase + (uint32_t)(m * 32);
            const U4 w = philox4x32_10_rk(c0, t0hi + (c0 < t0lo ? 1u : 0u), pj, pa, r.rk);
            const int t1st = draw_t1st(w.x, takeoff_t);
            const int last = t1st - 1;
            DCRP_ASSERT(last >= 0 && last < takeoff_t && last < tk_pad);
            const double2 kk = tabk[last];
            const float dls = (float)fma((double)w.y, ts.wx32, kk.x);     // target - kink, rounded once
            const float dbs = (float)fma((double)w.z, ts.wy32, kk.y);
            const float ada = fabsf(fmaf((float)w.w, ts.wz32, ts.z0rel));
            const Slab32 u = slab_ray32(dls, dbs, ada, tabpn[last], (float)last, nd.x, nd.y, s.v_straight_f, s.coef_f);
            const float sv = u.s, b = u.b;
            const bool force = u.degenerate || s1_first <= last;          // degenerate ray / stage-1 conflict
            fmask |= (valid && force) ? (1u << m) : 0u;
            t1s += valid ? (uint32_t)t1st : 0u;
            bn[m] = valid ? b : 1e15f;
            sn[m] = valid ? sv : 0.f;
        }

        // ---- LEVEL 1a, the slab: min over the indices of |n . D(i)|, D(i) = P(i) - A(i), P(i) = B + i*S
        f2_t BN[S3_NPL], SN[S3_NPL], PN[S3_NPL];
#pragma unroll
        for (int p = 0; p < S3_NPL; ++p) { BN[p] = pk2(bn[2 * p], bn[2 * p + 1]); SN[p] = pk2(sn[2 * p], sn[2 * p + 1]); }
        float mn[S3_TL];
#pragma unroll
        for (int m = 0; m < S3_TL; ++m) mn[m] = 3.0e38f;
        const float4* __restrict__ t4 = reinterpret_cast<const float4*>(tabT);
        for (int t0 = 0; t0 < VL; t0 += S3_REBASE) {
            const float t0f = (float)t0;
            const f2_t t02 = pk2(t0f, t0f);
#pragma unroll
            for (int p = 0; p < S3_NPL; ++p) PN[p] = ffma2(t02, SN[p], BN[p]);      // exact re-base
            const int nq = (min(VL, t0 + S3_REBASE) - t0 + 3) >> 2;
#pragma unroll 4
            for (int k = 0; k < nq; ++k) {
                DCRP_ASSERT(((t0 >> 2) + k) * 4 + 3 < vl_pad);
                const float4 cq = t4[(t0 >> 2) + k];
#define S3_STEP2(va, vb)                                                                      \
                {                                                                             \
                    const f2_t ca_ = pk2((va), (va)), cb_ = pk2((vb), (vb));                  \
                    _Pragma("unroll") for (int p_ = 0; p_ < S3_NPL; ++p_) {                   \
                        const float2 d0_ = up2(fadd2(PN[p_], ca_));                           \
                        PN[p_] = fadd2(PN[p_], SN[p_]);                                       \
                        const float2 d1_ = up2(fadd2(PN[p_], cb_));                           \
                        PN[p_] = fadd2(PN[p_], SN[p_]);                                       \
                        mn[2 * p_] = fminf(fminf(mn[2 * p_], fabsf(d0_.x)), fabsf(d1_.x));    \
                        mn[2 * p_ + 1] = fminf(fminf(mn[2 * p_ + 1], fabsf(d0_.y)), fabsf(d1_.y)); \
                    }                                                                         \
                }
                S3_STEP2(cq.x, cq.y)
                S3_STEP2(cq.z, cq.w)
#undef S3_STEP2
            }
        }

        // ---- survivors -> this warp's queue (item, local); a full round of them goes through the 2-D scan
        // (one compare of the lane's smallest minimum first: 99 % of the lanes have nothing to queue)
        float mall = mn[0];
#pragma unroll
        for (int m = 1; m + 1 < S3_TL; m += 2) mall = fminf(fminf(mall, mn[m]), mn[m + 1]);
        mall = fminf(mall, mn[S3_TL - 1]);
        uint32_t pass = fmask;
        if (mall < thr_slab) {
#pragma unroll
            for (int m = 0; m < S3_TL; ++m) if (mn[m] < thr_slab) pass |= 1u << m;
        }
        if (__any_sync(0xffffffffu, pass != 0u)) {
#pragma unroll
            for (int m = 0; m < S3_TL; ++m) {
                const bool on = (pass >> m) & 1u;
                const uint32_t bal = __ballot_sync(0xffffffffu, on);
                DCRP_ASSERT(list_n + __popc(bal) <= S3_LIST);
                if (on) list[list_n + __popc(bal & ((1u << lane) - 1u))] = make_uint2(item, rbase + (uint32_t)(m * 32 + lane));
                list_n += __popc(bal);
            }
            nsurv += (uint32_t)__popc(pass);
            __syncwarp();
            while (list_n >= S3_RESCAN) {
                griddep_wait();
                list_n -= S3_RESCAN;
                s3_rescan<SUB>(s, r, list + list_n, S3_RESCAN, list_track, chunks_per_pair, tile);
                __syncwarp();
            }
        }
        }   // rounds
        if (!waited) { TL_MARK(2); griddep_wait(); waited = true; if (last_of_block) grab_issue(); }
        else if (last_of_block && !prefetch) grab_issue();
        const uint32_t nv = n_valid > (uint32_t)lane ? (n_valid - (uint32_t)lane + 31u) >> 5 : 0u;   // this lane's valid slots
        acc_t1 += t1s;
        acc_is += nv * (uint32_t)VL;                 // slab lane-steps of this lane's trials
        if (acc_t1 > 0x40000000u || acc_is > 0x40000000u) {      // rare: keep the 32-bit partial sums exact
            atomicAdd(&r.stats[ST_SHARED], (unsigned long long)acc_t1);
            atomicAdd(&r.stats[ST_ISSUED], (unsigned long long)acc_is);
            acc_t1 = acc_is = 0;
        }
        if (++item == blk_end) {                                // (every block issues exactly one grab, at its last item)
            grab_take(item, blk_end);
            prefetch = blk_end - item >= 4u;
        }
    }
    TL_MARK(3);
    griddep_wait();
    if (list_n) s3_rescan<SUB>(s, r, list, list_n, list_track, chunks_per_pair, tile);
    TL_MARK(4);
    TL_COUNT(tl_items);
#ifdef DCRP_TIMELINE
    {
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        const uint32_t gw_ = blockIdx.x * (S3_BLOCK / 32) + (threadIdx.x >> 5);
        if (lane == 0 && gw_ < (uint32_t)TL_WARPS) g_timeline[gw_ * TL_SLOTS + 6] = smid;
    }
#endif

    // statistics: one set of global atomics per CTA, not per warp (they all land on the same three L2 lines, and
    // the warps of a launch finish within microseconds of each other)
    __shared__ unsigned long long cta_acc[3];
    if (threadIdx.x < 3) cta_acc[threadIdx.x] = 0ull;
    __syncthreads();
    const unsigned long long w1 = warp_sum_u64(acc_t1), w2 = warp_sum_u64(acc_is), w3 = warp_sum_u64(nsurv);
    if (lane == 0) { atomicAdd(&cta_acc[0], w1); atomicAdd(&cta_acc[1], w2); atomicAdd(&cta_acc[2], w3); }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (cta_acc[0]) atomicAdd(&r.stats[ST_SHARED], cta_acc[0]);
        if (cta_acc[1]) atomicAdd(&r.stats[ST_ISSUED], cta_acc[1]);
        if (cta_acc[2]) atomicAdd(&r.stats[ST_SLAB_SURVIVORS], cta_acc[2]);
    }

}

// Published run (engine.cu): launched right behind k_screen3 on the same stream, one CTA.  Copies the output arena
// (counters, statistics, counts, first-conflict indices) and the first records into the host's pinned mirror through
// its device mapping, fences at system scope and raises the flag the host polls; then -- clean -- clears the arena,
// which is what k_run_init would do in front of the next run.  A kernel of its own rather than the last CTA of
// k_screen3: code added to that kernel, even a time stamp at its entry, cost 1-3 % of it (A/B on one box).
struct PublishArgs {
    uint32_t* host;                  // the mirror as the device sees it
    uint32_t* arena;                 // the output arena (word 0 = record count)
    const uint32_t* records;
    uint32_t  words;                 // arena words in front of the records
    uint32_t  first_word;            // first word of first_conflict[] (cleared to INT32_MAX, everything before it to 0)
    uint32_t  rec_cap, inline_recs;  // records that travel with the mirror (right behind the arena words)
    uint32_t  flag_word;             // mirror word of the flag
    uint32_t  seq;                   // what the flag reads once the mirror is complete
    int32_t   clean;
};

__global__ void __launch_bounds__(256) k_publish(const PublishArgs a)
{
    for (uint32_t i = threadIdx.x; i < a.words; i += 256) a.host[i] = a.arena[i];
    const uint32_t nrec = min(min(a.arena[0], a.rec_cap), a.inline_recs);
    for (uint32_t i = threadIdx.x; i < nrec * (uint32_t)(sizeof(Record) / 4); i += 256) a.host[a.words + i] = a.records[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) *reinterpret_cast<volatile uint32_t*>(a.host + a.flag_word) = a.seq;
    if (a.clean)
        for (uint32_t i = threadIdx.x; i < a.words; i += 256) a.arena[i] = i < a.first_word ? 0u : 0x7fffffffu;
}

}  // namespace dcrp
