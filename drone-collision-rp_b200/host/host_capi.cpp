// host_capi.cpp -- extern "C" wrappers declared in include/dcrp_host.h.
#include "dcrp_host.h"

#include <fstream>
#include <string>

#include "Aircraft.h"
#include "Batch.h"
#include "Drone.h"
#include "synth_day.h"

static thread_local std::string g_host_err;

extern "C" const char* dcrph_last_error(void) { return g_host_err.c_str(); }
extern "C" const char* dcrph_reference_aircraft_csv(void) { return dcrp_host::kReferenceAircraftCsv; }

extern "C" int dcrph_synth_day_write(const char* path, int n_flights, uint64_t seed, int trailing_dummy)
{
    if (!path || n_flights <= 0) { g_host_err = "bad arguments"; return -1; }
    if (dcrp_host::write_synthetic_day(path, n_flights, seed, trailing_dummy != 0) != 0) {
        g_host_err = std::string("cannot write ") + path;
        return -6;
    }
    return 0;
}

extern "C" int dcrph_synth_day_write_mixed(const char* path, int n_flights, uint64_t seed, int trailing_dummy,
                                          int arrival_every)
{
    if (!path || n_flights <= 0 || arrival_every < 0) { g_host_err = "bad arguments"; return -1; }
    if (dcrp_host::write_synthetic_day_mixed(path, n_flights, seed, trailing_dummy != 0, arrival_every) != 0) {
        g_host_err = std::string("cannot write ") + path;
        return -6;
    }
    return 0;
}

extern "C" int dcrph_ring_write(const char* path, double radius_m, int n_rows)
{
    if (!path) { g_host_err = "path is NULL"; return -1; }
    if (dcrp_host::write_ring_csv(path, radius_m, n_rows) != 0) {
        g_host_err = std::string("cannot write ring ") + path;
        return -6;
    }
    return 0;
}

extern "C" int dcrph_aircraft_load(const char* path, int n_cols, int index, int* vector_length, int* takeoff_t,
                                   int* index_size, double* x, double* y, double* z, double* gs, int cap)
{
    if (!path) { g_host_err = "path is NULL"; return -1; }
    Aircraft a;
    a.Set_Parameters_and_Data(path, n_cols);
    if (!a.ok()) { g_host_err = a.last_error(); return -6; }
    if (index_size) *index_size = a.Aircraft_Index_size;
    if (index < 0) return 0;
    try {
        a.Vector_Allocation(index);
    } catch (const std::exception& ex) {
        g_host_err = std::string("unparsable cell: ") + ex.what();
        return -6;
    }
    if (a.Vector_length <= 0) { g_host_err = "flight index out of range"; return -1; }
    if (vector_length) *vector_length = a.Vector_length;
    if (takeoff_t) *takeoff_t = a.takeoff_t;
    const int n = a.Vector_length < cap ? a.Vector_length : cap;
    for (int i = 0; i < n; ++i) {
        if (x) x[i] = a.longitude_vector[i];
        if (y) y[i] = a.latitude_vector[i];
        if (z) z[i] = a.altitude_vector[i];
        if (gs) gs[i] = a.groundspeed_vector[i];
    }
    a.Deallocation();
    return n;
}

extern "C" int dcrph_ring_row(const char* path, int total_cols, int drone_index, double out[3])
{
    if (!path || !out) { g_host_err = "NULL argument"; return -1; }
    Drone d;
    double one = 0.0;
    d.SetInitialParameters(path, 1, total_cols, drone_index, 0, 1, &one, &one, &one, 3.5, 0.19);
    if (d.LastStatus() != 0) { g_host_err = d.LastError(); return d.LastStatus(); }
    d.InitialRow(out);
    return 0;
}

// Process-wide DCRP_FLAG_* applied by dcrph_monte_carlo / dcrph_monte_carlo_batch (their signatures mirror
// the reference's main() and carry no flag word).
static uint32_t g_host_run_flags = 0;
extern "C" void dcrph_set_run_flags(uint32_t flags) { g_host_run_flags = flags; }
// Devices dcrph_monte_carlo_batch spreads the trial range over (1 = default, 0 = all visible).
static int g_host_gpus = 1;
extern "C" void dcrph_set_gpus(int n_gpus) { g_host_gpus = n_gpus; }

extern "C" int dcrph_monte_carlo(const char* aircraft_csv, int aircraft_cols, int aircraft_begin,
                                 int aircraft_end, const char* drone_csv, int drone_cols, int drone_total,
                                 int number_runs, double aircraft_radius, double drone_radius, uint64_t seed,
                                 int precision, const char* out_dir, double* total_collisions,
                                 uint64_t* exact_total)
{
    if (!aircraft_csv || !drone_csv || !total_collisions) { g_host_err = "NULL argument"; return -1; }
    Aircraft aircraft;
    Drone drone;
    drone.SetSeed(seed);
    drone.SetPrecision(precision);
    drone.SetRunFlags(g_host_run_flags);
    if (out_dir && *out_dir) drone.SetOutputDirectory(out_dir);
    aircraft.Set_Parameters_and_Data(aircraft_csv, aircraft_cols);            // Monte_Carlo.cpp:26
    if (!aircraft.ok()) { g_host_err = aircraft.last_error(); return -6; }
    for (int i = aircraft_begin; i < aircraft_end; ++i) {                     // Monte_Carlo.cpp:28
        aircraft.Vector_Allocation(i);                                        // :29
        if (aircraft.Vector_length <= 0) { g_host_err = "flight index out of range"; return -1; }
        drone.ClearOutput(i);                                                 // :30
        for (int j = 0; j < drone_total; ++j) {                               // :31
            drone.SetInitialParameters(drone_csv, aircraft.Vector_length, drone_cols, j, i, aircraft.takeoff_t,
                                       aircraft.longitude_vector, aircraft.latitude_vector,
                                       aircraft.altitude_vector, aircraft_radius, drone_radius);   // :32
            drone.Simulation(number_runs, total_collisions);                  // :33
            if (drone.LastStatus() != 0) {
                g_host_err = drone.LastError();
                aircraft.Deallocation();
                return drone.LastStatus();
            }
        }
        aircraft.Deallocation();                                              // :35
    }
    if (exact_total) *exact_total = drone.ExactCollisions();
    return 0;
}

extern "C" int dcrph_monte_carlo_batch(const char* aircraft_csv, int aircraft_cols, int aircraft_begin,
                                       int aircraft_end, const char* drone_csv, int drone_cols, int drone_total,
                                       uint64_t number_runs, double aircraft_radius, double drone_radius,
                                       uint64_t seed, int precision, const char* out_dir,
                                       uint64_t* total_collisions)
{
    if (!aircraft_csv || !drone_csv || !total_collisions) { g_host_err = "NULL argument"; return -1; }
    const dcrp_host::BatchResult r = dcrp_host::monte_carlo_batch(
        aircraft_csv, aircraft_cols, aircraft_begin, aircraft_end, drone_csv, drone_cols, drone_total, number_runs,
        aircraft_radius, drone_radius, seed, precision, (out_dir && *out_dir) ? out_dir : "Drone_Collisions",
        g_host_run_flags, g_host_gpus);
    if (r.status != 0) { g_host_err = r.error; return r.status; }
    *total_collisions = r.collisions;
    return 0;
}

extern "C" int dcrph_write_block(const char* path, const double* x, const double* y, const double* z, int n,
                                 long run_no)
{
    if (!path || !x || !y || !z) { g_host_err = "NULL argument"; return -1; }
    std::ofstream out(path, std::ofstream::out | std::ofstream::app);
    if (!out) { g_host_err = std::string("cannot open ") + path; return -6; }
    dcrp_host::write_block(out, x, y, z, n, run_no);
    return 0;
}
