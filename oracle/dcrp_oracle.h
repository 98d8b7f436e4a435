/*
 * dcrp_oracle.h -- CPU restatement of the reference's Monte-Carlo trial path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call it, and only as the checker.
 * The product (libdcrp.so) never links or calls this file and has no CPU path.
 *
 * Parity status: PINNED.  (1) kinematics + record format against the 52
 * collision blocks the reference ships (Drone_Collisions/Drone_coords_*.csv,
 * compacted into tests/golden/ref_collision_blocks.json); (2) flag / first-hit /
 * min-distance / trajectory bit-for-bit against the UNMODIFIED reference
 * Drone.cpp rebuilt with the replay shim (oracle/_ref/libref_replay.so) on the
 * same replayed draws.  See tests/test_oracle_pinning.py.
 *
 * Every function cites the reference lines it restates (/root/reference/...).
 */
#ifndef DCRP_ORACLE_H
#define DCRP_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* One trial's four random draws (Drone.cpp:113-116 int draw, :175-187 three
 * real draws).  Same 32-byte layout as dcrp_draw in include/dcrp.h. */
typedef struct orc_draw {
    int32_t t1st;      /* random_t_1st, in [1, TakeoffTime]            */
    int32_t reserved;
    double  tx, ty, tz;/* random_long, random_lat, random_alt          */
} orc_draw;

typedef struct orc_model {
    double aircraft_radius;   /* Monte_Carlo.cpp:13  (3.5)  */
    double drone_radius;      /* Monte_Carlo.cpp:18  (0.19) */
    double start_alt;         /* Drone.cpp:54  (100.0) */
    double v_straight;        /* Drone.cpp:55  (19.0)  */
    double v_ascend;          /* Drone.cpp:56  (8.0)   */
} orc_model;

typedef struct orc_trial_result {
    int32_t hit;        /* Drone::Collision() return value                        */
    int32_t first_hit;  /* first index with distance < R (extension; -1 if none) */
    double  min_dist;   /* min over indices of the separation distance            */
} orc_trial_result;

void orc_default_model(orc_model* m);

/* Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy
 * as 1, 2, 3", SC'11).  Not in the reference (it uses random_device, which is
 * not replayable); this is the build's counter-based stream. */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* Draw mapping shared with the GPU engine: counter = (trial_lo, trial_hi,
 * drone, track), key = (seed_lo, seed_hi); word0 -> t1st in [1,T] (uniform int,
 * Drone.cpp:108-116), words 1..3 -> u = w * 2^-32, value = u*(b-a)+a
 * (libstdc++ uniform_real_distribution form, Drone.cpp:175-187). */
void orc_make_draw(uint64_t seed, uint64_t trial, uint32_t drone, uint32_t track,
                   int32_t takeoff_t, const double box[6], orc_draw* out);
void orc_make_draws(uint64_t seed, uint64_t trial_begin, int64_t n, uint32_t drone, uint32_t track,
                    int32_t takeoff_t, const double box[6], orc_draw* out);

/* Drone::CubedVolume, Drone.cpp:129-161 (+ depart_or_arrive, Drone.cpp:47-52).
 * box = {min_long,max_long,min_lat,max_lat,min_alt,max_alt}.  Returns 0, or -1
 * when latitude[0] is in neither runway band (reference leaves stale values). */
int orc_target_box(double ay0, double az0, double box[6]);

/* One trial: SetInitialConditions + FirstStage + SecondStage + Collision
 * (Drone.cpp:66-80, 107-127, 164-231, 253-265) with the draws supplied.
 * traj_x/y/z (each n_steps doubles) may be NULL. */
void orc_trial(const double* ax, const double* ay, const double* az, int32_t n_steps,
               double x0, double y0, double heading0,
               const orc_draw* draw, const orc_model* model,
               double* traj_x, double* traj_y, double* traj_z,
               orc_trial_result* out);

/* n trials of one (track, drone) with draws[n]; any output pointer may be NULL.
 * Returns the number of hits. */
int64_t orc_run_trials(const double* ax, const double* ay, const double* az, int32_t n_steps,
                       double x0, double y0, double heading0,
                       const orc_draw* draws, int64_t n, const orc_model* model,
                       uint8_t* hit, int32_t* first_hit, double* min_dist);

/* Free-running: Philox draws generated on the fly for trials
 * [trial_begin, trial_begin+n).  Returns hits; hit_ids (capacity cap) receives
 * the colliding trial ids in ascending order, *n_ids how many were written. */
int64_t orc_run_philox(const double* ax, const double* ay, const double* az, int32_t n_steps,
                       int32_t takeoff_t, const double box[6],
                       double x0, double y0, double heading0,
                       uint64_t seed, uint32_t drone, uint32_t track,
                       uint64_t trial_begin, int64_t n, const orc_model* model,
                       uint64_t* hit_ids, int64_t cap, int64_t* n_ids);

/* EXTENSION -- NOT in the reference (it compares sample i with sample i only,
 * Drone.cpp:254-259), parity UNPINNED for this function: it is the build's own
 * definition of the interpolated ("sub-step") separation and only checks that
 * the GPU implements the same definition.  Drone and aircraft move linearly
 * between consecutive samples; on segment i the relative position is
 * D(t) = D_i + t (D_{i+1} - D_i); a conflict is min_t |D(t)|^2 < R^2.
 * n trials of one (track, drone); trajectories come from orc_trial.
 * What IS anchored (tests/test_oracle_pinning.py::test_substep_extension_anchored_on_reference_trajectories):
 * the trajectories it interpolates are the reference's bit for bit; at the samples it reduces to the
 * reference (every reference hit is a sub-step hit in the adjacent segment, min distance <= the reference's);
 * and an independent numpy evaluation (closed form in long double + brute force on 4097 points per
 * segment) of trajectories written by the unmodified Drone.cpp gives the same minimum and conflict time. */
int64_t orc_run_trials_substep(const double* ax, const double* ay, const double* az, int32_t n_steps,
                               double x0, double y0, double heading0,
                               const orc_draw* draws, int64_t n, const orc_model* model,
                               uint8_t* hit, int32_t* first_seg, double* first_time, double* min_dist);

/* Drone::Output row format, Drone.cpp:237-241: ostream precision(10), default
 * floatfield == printf "%.10g".  Writes n_steps lines "x,y,z,run\n" plus one
 * blank line into buf (capacity cap); returns bytes written or -1. */
int64_t orc_format_block(const double* x, const double* y, const double* z, int32_t n_steps,
                         int32_t run_no, char* buf, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif
