/*
 * dcrp.h -- C ABI of the B200 Monte-Carlo collision-sampling engine (libdcrp.so).
 *
 * This is the drop-in boundary for the reference's trial loop.  The reference
 * (mohittank01/drone-collision-RP) has no FFI of its own; its seam is the class
 * API called from main() (Monte_Carlo.cpp:26-35).  The new host classes
 * (drone-collision-rp_b200/host/{Aircraft,Drone}.{h,cpp}) keep that API and
 * forward the per-trial work here.  Each entry point cites what it replaces.
 *
 * Conventions
 *   - plain C types, caller-owned host buffers borrowed for the call,
 *     engine-owned device memory; no torch / C++ types cross this ABI;
 *   - every function returns DCRP_OK (0) or a negative DCRP_ERR_*;
 *     dcrp_last_error() gives the message for the calling thread;
 *   - never exit(), never throws, and there is NO CPU fallback: without a
 *     usable CUDA device dcrp_engine_create fails with DCRP_ERR_NO_DEVICE;
 *   - thread-compatible: one engine per host thread.
 */
#ifndef DCRP_H
#define DCRP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCRP_ABI_VERSION 5

/* status codes */
#define DCRP_OK               0
#define DCRP_ERR_INVALID     -1   /* bad argument                                  */
#define DCRP_ERR_CUDA        -2   /* CUDA runtime error (message in last_error)    */
#define DCRP_ERR_NO_DEVICE   -3   /* no usable sm_100 device: no fallback          */
#define DCRP_ERR_RUNWAY      -4   /* latitude[0] in neither runway band            */
#define DCRP_ERR_STATE       -5   /* call sequence error (e.g. run before upload)  */
#define DCRP_ERR_IO          -6   /* file could not be opened / parsed             */
#define DCRP_ERR_COMM        -7   /* NCCL missing / failed, or a peer rank did not arrive */

/* arithmetic modes of the trial kernels */
#define DCRP_PRECISION_FP64   0   /* every trial in IEEE double, reference order of operations      */
#define DCRP_PRECISION_MIXED  1   /* fp32 packed screen of every trial + fp64 confirm of candidates:
                                     same counts / ids / records as FP64 (default, fast)            */
#define DCRP_PRECISION_FP32   2   /* fp32 decisions, no fp64 confirm (for fp32-vs-fp64 studies)     */

/* flags */
#define DCRP_FLAG_PER_TRIAL   1u  /* fill per-trial outputs (FP64 mode only; test aid)              */
#define DCRP_FLAG_SUBSTEP     2u  /* EXTENSION (not in the reference, which compares sample i with
                                     sample i only, Drone.cpp:254-259): drone and aircraft are
                                     interpolated linearly between samples and a collision is
                                     min over t of the separation on a segment < R.  first_hit is then
                                     the first SEGMENT in conflict, first_time the conflict time in
                                     index units, min_dist the continuous minimum.  FP64 / MIXED only. */
#define DCRP_FLAG_REACH_CULL  4u  /* OPT-IN, MIXED mode, sample mode only.  Exact culling by reachability:
                                     after the kink the drone moves at most max(v_straight, v_ascend) per
                                     index horizontally and only upwards, so sample i of a trial with kink
                                     index L can only be a conflict if the aircraft sample i lies within
                                     (i - L) * vmax + R of the kink and in the altitude band the drone can
                                     have reached.  Indices outside the feasible range [lo, hi) of a
                                     (track, drone, t1st), computed once per scenario, are decided by that
                                     bound and not scanned; a (track, drone) pair with no feasible index
                                     for any t1st and no stage-1 conflict gets no trial work at all (its
                                     count is 0 by the bound, no draws are made).  Counts, colliding trial
                                     ids and records are identical to the un-culled run; stats report what
                                     was computed (tests_issued) and what the bound decided (tests_culled). */

#define DCRP_FLAG_SCREEN_2D   8u  /* MIXED mode: skip the 1-D slab level and put every trial through the sorted
                                     2-D level-1 scan (the round-1 kernel, k_screen2).  Same results; for A/B
                                     measurements and as a cross-check of the two screens.                    */

#define DCRP_FLAG_ZERO_COUNTS 16u  /* dcrp_run_async with a caller-owned d_counts: clear it first (in the launch that
                                     clears the engine's own accumulators), instead of adding to what is there   */
#define DCRP_FLAG_NO_TIMING   32u  /* leave dcrp_stats.ms_* at 0 and record no timing events (dcrp_last_run_ms then
                                      returns DCRP_ERR_STATE).  For callers that do not read the times: a plain MIXED
                                      run of one launch with engine-owned counts is read back by a small kernel queued
                                      behind the trial kernel (it stores the results into the host's pinned mirror and
                                      raises a flag: no D2H copy, no wait for the stream), and the host learns of a
                                      completed CUDA event ~13 us later than of that flag                            */

typedef struct dcrp_engine dcrp_engine;

/* One aircraft track = the arrays Aircraft::Vector_Allocation produces
 * (Aircraft.cpp:113-128: longitude_vector, latitude_vector, altitude_vector,
 * Vector_length, takeoff_t) plus the target box Drone::CubedVolume derives from
 * it (Drone.cpp:129-161). */
typedef struct dcrp_track {
    const double* x;        /* UTM easting  per sample ("longitude")             */
    const double* y;        /* UTM northing per sample ("latitude")              */
    const double* z;        /* altitude per sample                               */
    int32_t n_steps;        /* Vector_length                                     */
    int32_t takeoff_t;      /* TakeoffTime: t1st ~ U{1..takeoff_t}               */
    double  box[6];         /* min_long,max_long,min_lat,max_lat,min_alt,max_alt */
} dcrp_track;

/* One drone start row: drone_inital_positions_*.csv columns 1..3 of row
 * DroneIndex+1 (Drone.cpp:33-35). */
typedef struct dcrp_drone {
    double x0, y0, heading0;   /* heading: radians clockwise from north */
} dcrp_drone;

/* One trial's four draws (Drone.cpp:113-116, 175-187): the replay-table row. */
typedef struct dcrp_draw {
    int32_t t1st;
    int32_t reserved;
    double  tx, ty, tz;
} dcrp_draw;

/* Model constants (Monte_Carlo.cpp:13,18; Drone.cpp:54-56). */
typedef struct dcrp_model {
    double aircraft_radius;    /* 3.5   */
    double drone_radius;       /* 0.19  */
    double start_alt;          /* 100.0 */
    double v_straight;         /* 19.0  */
    double v_ascend;           /* 8.0   */
} dcrp_model;

/* What to run.  Trials are identified by a global 64-bit id; the draws of
 * trial t for (track a, drone j) are Philox4x32-10(key = seed,
 * counter = (t_lo, t_hi, drone_id_base + j, track_id_base + a)), so results do
 * not depend on how [trial_begin, trial_begin+trial_count) -- or the drones, or
 * the tracks -- are split across calls or GPUs.  Records and counts are indexed
 * by the scenario-local (a, j). */
typedef struct dcrp_run {
    uint64_t seed;
    uint64_t trial_begin;
    uint64_t trial_count;      /* per (track, drone); the reference's number_runs (Drone.h:74) widened */
    int32_t  precision;        /* DCRP_PRECISION_*                                                    */
    uint32_t flags;            /* DCRP_FLAG_*                                                         */
    uint32_t record_capacity;  /* max collision records kept (0 = engine default)                     */
    uint32_t track_id_base;    /* Philox ids: scenario track a draws as track (track_id_base + a),    */
    uint32_t drone_id_base;    /* drone j as drone (drone_id_base + j) -- lets one pair, or a shard   */
    uint32_t reserved;         /* of the drones, reproduce exactly what the full batch draws          */
    const dcrp_draw* replay;   /* optional HOST table [n_tracks][n_drones][trial_count]: draws are
                                  read from it instead of Philox (replayed-draw parity runs)         */
} dcrp_run;

/* Compact collision record: what Drone::Output (Drone.cpp:233-243) needs to
 * regenerate a block, plus the first-conflict index the reference never kept. */
typedef struct dcrp_record {
    uint32_t track;
    uint32_t drone;
    uint64_t trial;            /* global trial id                   */
    int32_t  t1st;
    int32_t  first_hit;        /* first index with distance < R     */
    double   tx, ty, tz;       /* stage-2 target                    */
    double   min_dist;         /* min separation over the track (m) */
    double   first_time;       /* first-conflict time in index units: == first_hit in sample mode,
                                  segment + t in [0,1) with DCRP_FLAG_SUBSTEP                         */
} dcrp_record;

/* dcrp_stats.status bits */
#define DCRP_STAT_SCENARIO_CACHED 1u  /* dcrp_simulate: tracks/drones/model equal the resident scenario bit for bit;
                                         nothing was validated, packed or uploaded again (h2d_bytes counts 0 for it) */
#define DCRP_STAT_FP64_FALLBACK   2u  /* MIXED run: the candidate list overflowed (absurd hit rates) and the whole
                                         run was redone by the fp64 kernel -- same answers, ~4x the time             */
#define DCRP_STAT_CULL_IGNORED    4u  /* DCRP_FLAG_REACH_CULL was set but does not apply to this run (FP64 / FP32
                                         precision, per-trial outputs, replay table, sub-step mode): un-culled run   */

typedef struct dcrp_stats {
    uint64_t trials;           /* trials simulated                                                    */
    uint64_t tests_evaluated;  /* stage-2 (trial, index) separation tests the kernels computed:
                                  sum over trials of (n_steps - t1st)                                 */
    uint64_t tests_shared;     /* stage-1 tests: identical for every trial of a (track, drone), so
                                  evaluated once per pair and shared: sum over trials of t1st        */
    uint64_t tests_issued;     /* lane-steps actually issued by the scan loops (>= evaluated)         */
    uint64_t drone_steps;      /* sum over trials of (n_steps - 1)                                    */
    uint64_t level2;           /* trials the horizontal fp32 screen (level 1) passed to the 3-D fp32
                                  re-scan (level 2)                                                   */
    uint64_t candidates;       /* trials level 2 passed on to the fp64 confirm (level 3)              */
    uint64_t hits;             /* total collisions                                                    */
    uint32_t kernel_launches;  /* engine kernels launched by the call                                 */
    uint32_t status;           /* DCRP_STAT_* bits: what the engine did differently from what was asked,
                                  or did not have to do (never an error; results are the same)           */
    float    ms_prepare;       /* stage-1 / per-pair tables                                           */
    float    ms_trials;        /* trial kernel(s): screen or exact                                    */
    float    ms_confirm;       /* fp64 confirm of candidates                                          */
    float    ms_device_total;  /* first to last kernel, CUDA events on the run's stream (0 with
                                  DCRP_FLAG_NO_TIMING)                                                    */
    uint64_t h2d_bytes;
    uint64_t d2h_bytes;        /* bytes brought to the host: by the copy engine, or by the kernel's stores    */
    /* DCRP_FLAG_REACH_CULL only (0 otherwise) */
    uint64_t trials_culled;    /* trials of (track, drone) pairs the reach bound cleared as a whole: no
                                  draws made, nothing in tests_shared                                     */
    uint64_t tests_culled;     /* stage-2 tests decided by the reach bound instead of being evaluated:
                                  exact for the pairs that ran (there tests_evaluated + tests_culled =
                                  sum of n_steps - t1st), the GUARANTEED minimum n_steps - takeoff_t per
                                  trial for the cleared pairs; never part of tests_evaluated             */
    uint64_t slab_survivors;   /* MIXED runs through the slab-first screen: trials the 1-D slab level could not
                                  clear and queued for the 2-D level-1 re-scan (0 when every trial goes through
                                  the 2-D scan: DCRP_FLAG_SCREEN_2D, replay, culling, long tracks, and sub-step
                                  runs on a scenario whose previous sub-step run kept > 25 % survivors)      */
} dcrp_stats;

/* Result buffers (host, caller-owned).  counts is required; the rest optional. */
typedef struct dcrp_result {
    uint64_t*    counts;          /* [n_tracks*n_drones] hits, ADDED to what is there (the reference
                                     accumulates into *total_collisions, Drone.cpp:276)              */
    int32_t*     first_conflict;  /* [n_tracks*n_drones] min first_hit over hits, INT32_MAX if none   */
    dcrp_record* records;         /* [record_capacity]                                                */
    uint32_t     n_records;       /* out: records written (sorted by track, drone, trial)             */
    int32_t      overflow;        /* out: 1 when more hits than record_capacity                       */
    uint8_t*     trial_hit;       /* DCRP_FLAG_PER_TRIAL: [n_tracks*n_drones*trial_count]             */
    int32_t*     trial_first_hit; /* idem                                                             */
    double*      trial_min_dist;  /* idem                                                             */
    dcrp_stats   stats;           /* out                                                              */
} dcrp_result;

/* ---- lifetime --------------------------------------------------------- */
int  dcrp_abi_version(void);
const char* dcrp_last_error(void);
void dcrp_default_model(dcrp_model* m);

/* Owns a CUDA stream, device buffers and scratch on `device`. */
int  dcrp_engine_create(int device, dcrp_engine** out);
int  dcrp_engine_destroy(dcrp_engine* e);

/* ---- host-side helper -------------------------------------------------- */
/* Replaces Drone::CubedVolume (Drone.cpp:129-161) + depart_or_arrive
 * (Drone.cpp:47-52).  Unlike the reference (stale values, no else branch) an
 * out-of-band first latitude is an error: DCRP_ERR_RUNWAY. */
int  dcrp_target_box(double ay0, double az0, double box[6]);

/* ---- the trial loop ----------------------------------------------------- */
/* Replaces the body of Drone::Simulation (Drone.cpp:269-281) for every
 * (track, drone) pair at once: SetInitialConditions, FirstStage, SecondStage,
 * Collision, the hit counter and the data Output needs.  Blocking; host
 * buffers in, host buffers out (uploads, kernels, downloads inside). */
int  dcrp_simulate(dcrp_engine* e,
                   const dcrp_track* tracks, int32_t n_tracks,
                   const dcrp_drone* drones, int32_t n_drones,
                   const dcrp_model* model, const dcrp_run* run,
                   dcrp_result* result);

/* The same in three steps, for callers that keep a scenario resident in HBM
 * (bench, multi-GPU ranks): upload once, run many trial ranges, fetch. */
int  dcrp_scenario_upload(dcrp_engine* e,
                          const dcrp_track* tracks, int32_t n_tracks,
                          const dcrp_drone* drones, int32_t n_drones,
                          const dcrp_model* model);
/* Asynchronous on `stream` (a cudaStream_t; NULL = the engine's own stream).
 * d_counts: optional DEVICE buffer [n_tracks*n_drones] of uint64 that hits are
 * atomically added to (e.g. a torch tensor that is all-reduced afterwards);
 * NULL = engine-owned.  Clears the engine's per-run accumulators first. */
int  dcrp_run_async(dcrp_engine* e, const dcrp_run* run, void* stream, uint64_t* d_counts);
/* Waits for the run and copies counts / records / stats to the host. */
int  dcrp_fetch(dcrp_engine* e, dcrp_result* result);
/* Milliseconds of the last run (device time between the engine's events). */
int  dcrp_last_run_ms(dcrp_engine* e, float* ms);

/* ---- all-pairs stress (BASELINE config 5; an extension, not in the reference) ---------- */
/* Every ray against every track.  A ray is one trial's flight -- drone j, t1st ~ U{1..ref_takeoff_t},
 * target uniform in ref_box -- drawn ONCE from Philox(key = seed, counter = (ray_lo, ray_hi,
 * drone_id_base + j, 0xFFFFFFFF)) and then tested, with the kinematics and the separation test of
 * Drone.cpp:107-127,164-231,253-265, against EVERY track of the uploaded scenario (the reference
 * only ever pairs one aircraft with one drone's own trials).  ref_takeoff_t must not exceed the
 * largest takeoff_t of the scenario.  result->counts[a * n_drones + j] += rays of drone j that
 * collide with track a; records carry trial = ray id.  Blocking. */
typedef struct dcrp_allpairs_run {
    uint64_t seed;
    uint64_t ray_begin;        /* rays [ray_begin, ray_begin + ray_count) of every drone */
    uint64_t ray_count;
    int32_t  ref_takeoff_t;
    uint32_t record_capacity;
    uint32_t drone_id_base;
    uint32_t reserved;
    double   ref_box[6];
} dcrp_allpairs_run;
int  dcrp_allpairs(dcrp_engine* e, const dcrp_allpairs_run* run, dcrp_result* result);

/* ---- replay / records --------------------------------------------------- */
/* The Philox draws of trials [run->trial_begin, +run->trial_count) of scenario
 * pair (track, drone) -- seed and id bases from `run` -- generated ON DEVICE by
 * the same code the trial kernels use: the table the reference is replayed on.
 * out[run->trial_count]. */
int  dcrp_dump_draws(dcrp_engine* e, int32_t track, int32_t drone, const dcrp_run* run,
                     dcrp_draw* out);
/* Full trajectories of recorded collisions, computed ON DEVICE in fp64 with the
 * exact-kernel arithmetic: the rows Drone::Output writes (Drone.cpp:237-241).
 * xyz receives, per record r, x[0..n) y[0..n) z[0..n) at xyz + r*3*stride,
 * n = that track's n_steps, stride = max n_steps over the scenario. */
int  dcrp_trajectories(dcrp_engine* e, const dcrp_record* records, uint32_t n_records,
                       double* xyz, int32_t* stride_out);

/* ---- multi-process: one process per GPU, per-rank hit counts combined ---------------------------------- */
/* rank g runs a disjoint Philox trial range of every (track, drone) -- the loop of
 * Monte_Carlo.cpp:28-34 sharded -- and the per-pair counters (Drone.cpp:276) are summed over the ranks.
 * Rank 0 calls dcrp_comm_unique_id and ships the 128 bytes to the other ranks by any means (a file, a
 * socket, torch.distributed); every rank then attaches.  The engine owns an NCCL communicator (libnccl.so.2
 * is dlopen'ed on first use: no link-time dependency) and, unless DCRP_COMM_NCCL_ONLY is set, maps every
 * peer's mailbox through CUDA IPC so that the all-reduce is two small kernels of its own over NVLink peer
 * memory (csrc/comm.cu says why); ncclAllReduce(sum, uint64) is the fallback and the cross-check. */
#define DCRP_UNIQUE_ID_BYTES 128
#define DCRP_COMM_NCCL_ONLY  1u
int  dcrp_comm_unique_id(void* out128);
int  dcrp_engine_attach_comm(dcrp_engine* e, const void* unique_id128, int rank, int nranks, uint32_t flags);
int  dcrp_engine_detach_comm(dcrp_engine* e);      /* collective-free; also done by dcrp_engine_destroy */
/* In-place sum over the ranks of the n_tracks*n_drones uint64 counters in d_counts (a device buffer, e.g. the
 * one passed to dcrp_run_async; NULL = the engine's own counts of the last run, so that the next dcrp_fetch
 * returns the global counts).  Asynchronous on `stream` (NULL = the stream of the last run).  Every rank must
 * call it the same number of times in the same order.  A peer that does not arrive within 10 s makes the next
 * dcrp_fetch return DCRP_ERR_COMM instead of hanging. */
int  dcrp_allreduce_counts(dcrp_engine* e, void* stream, uint64_t* d_counts);
/* info[0] attached, [1] rank, [2] nranks, [3] peer-memory route in use, [4] NCCL version code,
 * [5] all-reduces done over peer memory, [6] by ncclAllReduce, [7] peer-timeout flag */
int  dcrp_comm_info(dcrp_engine* e, int32_t info[8]);

/* ---- device queries ------------------------------------------------------ */
/* (The roofline micro-benchmarks and the L2 flush of bench.py live in libdcrp_bench.so, csrc/bench_kernels.cu:
 * measurement code is not part of this library.) */
/* Number of CUDA devices visible to the process (0 and DCRP_ERR_NO_DEVICE when there is none). */
int  dcrp_device_count(int* n);
/* Keeps `n_sms` SMs out of the persistent trial grids (0 = use them all, the default).  A multi-GPU caller
 * that overlaps a collective with the next run reserves one: the trial kernels fill every SM's register file,
 * so a communication kernel launched beside them may otherwise have to wait for their last wave to drain
 * (measured at N=8: the per-step all-reduce was hidden in some runs and fully exposed, +70 us, in others). */
int  dcrp_engine_reserve_sms(dcrp_engine* e, int n_sms);
/* Kernels this engine has launched since creation. */
int  dcrp_launch_count(dcrp_engine* e, uint64_t* n);
/* Device properties the bench reports. */
int  dcrp_device_info(dcrp_engine* e, char* name, int name_cap, int* sm_count,
                      int* cc_major, int* cc_minor, int* clock_khz);


/* ---- diagnostics ----------------------------------------------------------- */
/* How a slab-screen launch of `n_items` 256-trial work items deals them out to `n_warps` warps after the first
 * `n_static` items (which go to the warps statically): block k of the device work counter covers the items
 * [begin[k], end[k]).  Host arithmetic only (no device needed); fills at most `cap` blocks and returns the number
 * of blocks of the schedule (negative = error).  The CPU test tier checks with it that every item is dealt out
 * exactly once for launch sizes the GPU tests do not reach. */
int64_t dcrp_debug_block_schedule(uint32_t n_items, uint32_t n_warps, uint32_t n_static,
                                  uint32_t* begin, uint32_t* end, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* DCRP_H */
