/*
 * ref_replay_harness.cpp -- drives the UNMODIFIED reference Drone.cpp on
 * replayed draws.  TEST INFRASTRUCTURE ONLY.  Compiled by oracle/Makefile as
 *
 *   g++ -O3 -ffp-contract=off -I/root/reference -Ioracle/ref_shim \
 *       -include replay_shim.h -shared -fPIC ref_replay_harness.cpp \
 *       -o oracle/_ref/libref_replay.so
 *
 * The reference translation unit is #included (not copied) from where it lies,
 * with `private` opened up so the harness can call the per-trial stages
 * (Drone.cpp:66-80, 107-127, 164-231, 253-265) and read the trajectory arrays
 * before Drone::Deallocate (Drone.cpp:245-251) frees them.
 */
#include "replay_shim.h"   /* also force-included; guarded */

#define private public
#include "Drone.cpp"       /* resolved through -I/root/reference */
#undef private

#include <cstring>
#include <limits>

namespace dcrp_replay {
thread_local const Draw* g_next = nullptr;
thread_local const Draw* g_cur = nullptr;
thread_local int         g_real_pos = 0;
thread_local long        g_range_violations = 0;
}

extern "C" {

/* Stage-by-stage replay of n trials of one (aircraft, drone).
 * hit[r] is the reference's own Drone::Collision() result; first_hit/min_dist
 * are harness-side reductions over the reference's trajectory arrays using the
 * reference's distance expression.  traj (optional) receives n*3*VL doubles:
 * for trial r, x at [r*3*VL .. +VL), then y, then z. */
long ref_replay_run(const char* drone_csv, int total_cols, int drone_index, int aircraft_index,
                    int VL, int takeoff_t,
                    const double* ax, const double* ay, const double* az,
                    double r_ac, double r_dr,
                    const dcrp_replay::Draw* draws, long n,
                    unsigned char* hit, int* first_hit, double* min_dist,
                    double* traj, long* range_violations)
{
    Drone d;
    d.SetInitialParameters(std::string(drone_csv), VL, total_cols, drone_index, aircraft_index,
                           takeoff_t, const_cast<double*>(ax), const_cast<double*>(ay),
                           const_cast<double*>(az), r_ac, r_dr);
    dcrp_replay::g_range_violations = 0;
    long hits = 0;
    for (long r = 0; r < n; ++r) {
        dcrp_replay::g_next = &draws[r];
        d.SetInitialConditions();
        d.FirstStage();
        d.SecondStage();
        bool h = d.Collision();
        hits += h ? 1 : 0;
        if (hit) hit[r] = h ? 1 : 0;
        if (first_hit || min_dist) {
            int first = -1;
            double mind = std::numeric_limits<double>::infinity();
            const double R = r_dr + r_ac;
            for (int i = 0; i < VL; ++i) {
                double dist = sqrt(
                    (d.longitude_vector[i] - ax[i]) * (d.longitude_vector[i] - ax[i]) +
                    (d.latitude_vector[i] - ay[i]) * (d.latitude_vector[i] - ay[i]) +
                    (d.altitude_vector[i] - az[i]) * (d.altitude_vector[i] - az[i]));
                if (dist < mind) mind = dist;
                if (dist < R && first < 0) first = i;
            }
            if (first_hit) first_hit[r] = first;
            if (min_dist) min_dist[r] = mind;
        }
        if (traj) {
            double* t = traj + (size_t)r * 3 * (size_t)VL;
            std::memcpy(t, d.longitude_vector, sizeof(double) * (size_t)VL);
            std::memcpy(t + VL, d.latitude_vector, sizeof(double) * (size_t)VL);
            std::memcpy(t + 2 * (size_t)VL, d.altitude_vector, sizeof(double) * (size_t)VL);
        }
        d.Deallocate();
    }
    if (range_violations) *range_violations = dcrp_replay::g_range_violations;
    return hits;
}

/* The reference's own trial loop, Drone::Simulation (Drone.cpp:269-281), on
 * replayed draws: exercises Drone::Output / ClearOutput too (files land in
 * ./Drone_Collisions relative to the caller's cwd, Drone.cpp:62,236). */
double ref_replay_simulation(const char* drone_csv, int total_cols, int drone_index,
                             int aircraft_index, int VL, int takeoff_t,
                             const double* ax, const double* ay, const double* az,
                             double r_ac, double r_dr,
                             const dcrp_replay::Draw* draws, int n, int clear_output)
{
    Drone d;
    double total = 0;
    if (clear_output) d.ClearOutput(aircraft_index);
    d.SetInitialParameters(std::string(drone_csv), VL, total_cols, drone_index, aircraft_index,
                           takeoff_t, const_cast<double*>(ax), const_cast<double*>(ay),
                           const_cast<double*>(az), r_ac, r_dr);
    dcrp_replay::g_next = draws;
    d.Simulation(n, &total);
    return total;
}

/* Drone::CubedVolume (Drone.cpp:129-161) as the reference evaluates it; the six
 * bounds are pre-filled with NaN so an out-of-band latitude is visible. */
void ref_cubed_volume(double ay0, double az0, double box[6])
{
    Drone d;
    double x = 0, y = ay0, z = az0;
    d.aircraft_longitude = &x; d.aircraft_latitude = &y; d.aircraft_altitude = &z;
    d.depart_or_arrive = (z != 0) ? 1 : 0;   /* Drone.cpp:47-52 */
    const double nan = std::numeric_limits<double>::quiet_NaN();
    d.min_cube_long = d.max_cube_long = d.min_cube_lat = d.max_cube_lat = nan;
    d.min_cube_alt = d.max_cube_alt = nan;
    d.CubedVolume();
    box[0] = d.min_cube_long; box[1] = d.max_cube_long;
    box[2] = d.min_cube_lat;  box[3] = d.max_cube_lat;
    box[4] = d.min_cube_alt;  box[5] = d.max_cube_alt;
}

}  // extern "C"
