// dcrp_device.cuh -- device-side data layout of a scenario and a run.
//
// HBM layout (all engine-owned, built by dcrp_scenario_upload):
//   tracks[n_tracks]            TrackDev: n_steps, takeoff_t, sample offset, target box
//   ax, ay, az [sum n_steps]    fp64 SoA, tracks back to back  (exact / confirm kernels)
//   trk32 [sum n_steps]         fp32 screen copy, 16 B per sample, relative to the scenario
//                               origin and negated: (-ax, -ay, -az, i)  (i as float); one
//                               broadcast LDS.128 per index feeds the packed f32x2 ops through
//                               their scalar-broadcast operand form
//   trkxy [sum even(n_steps)]   the horizontal part alone, 8 B per sample: what level 1 streams
//                               (one broadcast LDS.128 feeds TWO indices)
//   drones[n_drones]            x0, y0, sin(h0)*v, cos(h0)*v   (fp64; trig by the host's libm,
//                               the same one the reference links, Drone.cpp:121-122)
//   p1x, p1y [n_drones*t_max]   stage-1 path of every drone, sequentially accumulated fp64
//   s1_first_hit [n_pairs]      first stage-1 index in collision (INT32_MAX if none)
//   s1_prefmin  [n_pairs*t_max] prefix minimum of the stage-1 squared separation
//   kink [n_pairs*t_max], p1f [n_drones*t_max]  pre-subtracted kink tables of the fp32 screen
// Stage 1 does not depend on the draws except through t1st, so it is evaluated
// once per (track, drone) and shared by all of that pair's trials.
#pragma once

#include <stdint.h>
#include <cuda_runtime.h>
#include "dcrp_math.cuh"

// Bounds checks of our own (compute-sanitizer is closed on the GPU pool): `make EXTRA=-DDCRP_BOUNDS` turns every
// DCRP_ASSERT on an index into a printf + trap, and the GPU test tier is run once against that build
// (profiles/r02_bounds_build_pytest.log).  Nothing in the release build.
#if defined(DCRP_BOUNDS) && defined(__CUDA_ARCH__)
#include <stdio.h>
#define DCRP_ASSERT(cond)                                                                   \
    do {                                                                                    \
        if (!(cond)) {                                                                      \
            printf("DCRP_BOUNDS violated: %s  (%s:%d)\n", #cond, __FILE__, __LINE__);       \
            __trap();                                                                       \
        }                                                                                   \
    } while (0)
#else
#define DCRP_ASSERT(cond) ((void)0)
#endif

namespace dcrp {

struct TrackDev {
    int32_t  n_steps;
    int32_t  takeoff_t;
    uint32_t off;        // first sample of this track in ax/ay/az and trk32
    uint32_t off_xy;     // first sample of this track in trkxy (even: every track starts 16-B aligned)
    double   box[6];
    // fast screen set-up: target = box_min + w * (extent * 2^-32), w = Philox word
    double   wx32, wy32; // (box[1]-box[0]) * 2^-32, (box[3]-box[2]) * 2^-32
    float    wz32;       // (box[5]-box[4]) * 2^-32
    float    z0rel;      // box[4] - start_alt
    float    thr2_l1_sub;// sub-step mode, level-1 threshold: (R + margin + (max |dA_xy| + v_straight)/2)^2
    float    thr2_l2_sub;// sub-step mode, level-2 threshold: (R + margin)^2 + the fp32 error of the segment closest approach
    float    dir0x, dir0y;// default slab direction of this track's pairs: the aircraft's main direction of travel (unit)
    float    thr_slab_sub;// sub-step mode, slab threshold (NOT squared): R + slab margin + (max |dA_xy| + vmax)/2
    float    thr2_near_sub;// sub-step mode, level 2: a segment whose two ends are both farther than this (squared) cannot
                          // come within the level-2 radius (ends closer than sqrt(thr2_l2_sub) + half the longest relative step)
};

struct ScenarioDev {
    const TrackDev* tracks;
    const double*   ax;
    const double*   ay;
    const double*   az;
    const float4*   trk32;
    const float2*   trkxy;           // level-1 copy: (-ax', -ay') only, 8 B per sample, tracks padded to even
    const double4*  drones;
    double*         p1x;
    double*         p1y;
    int32_t*        s1_first_hit;
    double*         s1_prefmin;
    double*         s1sub_prefmin;   // [n_pairs*t_max] sub-step: min d2 over stage-1 SEGMENTS with index < L
    int32_t*        s1sub_first_seg; // [n_pairs] first stage-1 segment in conflict (INT32_MAX if none)
    double*         s1sub_first_time;// [n_pairs] its conflict time (segment index + t)
    double2*        kink;            // [n_pairs*t_max] (box.xmin - p1x[L], box.ymin - p1y[L]): fast screen set-up
    float2*         p1f;             // [n_drones*t_max] stage-1 path relative to the origin, rounded once to fp32
    // reachability (DCRP_FLAG_REACH_CULL): feasible stage-2 index range [lo, hi) per (pair, t1st), the pairs
    // with any feasible index (or a stage-1 conflict), and the guaranteed tests of the others
    int2*           reach;           // [n_pairs*t_max] entry t1st-1; empty = (INT32_MAX, 0)
    uint32_t*       alive;           // [n_pairs] pair ids, n_alive of them valid (unordered)
    unsigned long long* reach_tot;   // [0] n_alive  [1] sum over dead pairs of (n_steps - takeoff_t)
    float2*         slab_dir;        // [n_pairs] unit vector n of the 1-D slab level (k_slab_axis; screen3.cu)
    int32_t n_tracks, n_drones, t_max, n_steps_max;
    int32_t n_sm;                    // SMs of the device (CTA stagger)
    FastDiv div_drones;              // pair / n_drones
    double ox, oy, oz;               // fp32 origin
    double start_alt, v_straight, v_ascend;
    double radius;                   // R_drone + R_aircraft, summed in fp64 (Drone.cpp:260)
    double thr_d2;                   // sqrt_rn(d2) < radius  <=>  d2 < thr_d2   (exact, host-searched)
    double radius2;                  // radius * radius: the sub-step mode's threshold on |D(t)|^2
    float  thr2_screen;              // (radius + margin)^2 : fp32 screen, conservative
    float  thr2_fp32;                // radius^2 in fp32    : DCRP_PRECISION_FP32 decisions
    float  thr_slab;                 // radius + slab margin (NOT squared): 1-D slab level, conservative
    float  v_straight_f, coef_f;     // 19, 2*(19-8)/pi
    int32_t bucket_shift;            // t1st bucket = (t1st-1) >> bucket_shift   (<= 64 buckets)
};

struct Record {          // == dcrp_record (64 bytes)
    uint32_t track, drone;
    uint64_t trial;
    int32_t  t1st, first_hit;
    double   tx, ty, tz;
    double   min_dist;
    double   first_time;
};

struct Candidate {       // a trial the fp32 screen could not rule out
    uint32_t pair;
    uint32_t pad;
    uint64_t trial;
};

// All-pairs stress run (BASELINE config 5): every ray against every track.
struct AllPairsDev {
    uint64_t seed, ray_begin, ray_count;       // rays [begin, begin+count) of every drone
    uint32_t drone_base;
    int32_t  ref_takeoff_t;
    double   ref_box[6];
    double   wx32, wy32;                       // ref box extents * 2^-32 (fast set-up)
    float    wz32, z0rel;
    uint32_t* hist;                            // [n_drones * ref_takeoff_t] rays per (drone, t1st); then bases
    uint32_t* rank;                            // [n_drones * ray_count]
    // rays per drone sorted by t1st and stored pair-interleaved, so that a lane's two rays arrive as
    // ready-made aligned f32x2 register pairs: group g = slot/64 holds 32 lane records,
    //   rayA[(j*groups + g)*32 + lane] = (bx0, bx1, by0, by1),  rayB[...] = (sx0, sx1, sy0, sy1)
    // with ray m of the lane = sorted slot g*64 + m*32 + lane (slots >= ray_count are dead: b = 1e15)
    float*    rayA;
    float*    rayB;
    uint32_t  groups;                          // ceil(ray_count / 64)
    uint2*    meta;                            // [n_drones * ray_count] (t1st | degenerate << 31, ray_local)
    uint2*    l2;                              // level-2 list: (track, drone * ray_count + slot)
    uint32_t* l2_count;
    uint32_t  l2_cap;
    uint32_t  blocks_per_drone;                // ray blocks of AP_RAYS rays
    uint32_t  tracks_per_chunk, n_tchunks;
};

enum StatSlot { ST_EVALUATED = 0, ST_SHARED, ST_ISSUED, ST_STEPS, ST_CANDIDATES, ST_HITS, ST_LEVEL2,
                ST_CULLED_TRIALS, ST_CULLED_TESTS, ST_DEAD_TESTS, ST_PROCESSED_TESTS, ST_SLAB_SURVIVORS, ST_COUNT };

struct RunDev {
    uint64_t seed, trial_begin, trial_count;   // the whole run: trials [begin, begin+count) per pair
    uint64_t slice_first, slice_count;         // this launch covers local trials [first, first+count)
    uint32_t track_base, drone_base;           // Philox ids = base + scenario index
    const Draw* replay;                  // device copy of the replay table or nullptr
    unsigned long long* counts;          // [n_pairs]
    int32_t*  first_conflict;            // [n_pairs]
    Record*   records;
    uint32_t* rec_count;
    uint32_t  rec_cap;
    Candidate* cands;
    uint32_t* cand_count;
    uint32_t  cand_cap;
    unsigned long long* stats;           // [ST_COUNT]
    uint8_t*  trial_hit;                 // per-trial outputs or nullptr
    int32_t*  trial_first;
    double*   trial_min;
    int32_t   decide_fp32;               // screen kernel: 1 = DCRP_PRECISION_FP32
    int32_t   substep;                   // 1 = DCRP_FLAG_SUBSTEP: interpolated separation between samples
    uint32_t  rk[20];                    // Philox round keys of `seed` (philox_round_keys)
    FastDiv   div_chunks;                // item / chunks_per_pair of the launch
    int32_t   cull;                      // 1 = DCRP_FLAG_REACH_CULL
    uint32_t* work_counter;              // k_screen3: next unclaimed block of the launch
};

}  // namespace dcrp
