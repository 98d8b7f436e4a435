/*
 * dcrp_host.h -- C entry points into the C++ host mirror of the reference API
 * (libdcrp_host.so: drone-collision-rp_b200/host/).  They exist so that tests
 * and tools written in other languages (the Python tests bind them with ctypes)
 * can drive the very classes a C++ user of the reference links against:
 *   Aircraft  (reference Aircraft.h:36-48, Aircraft.cpp:10-136)
 *   Drone     (reference Drone.h:71-81,   Drone.cpp:11-64,269-281)
 *   main loop (reference Monte_Carlo.cpp:26-38)
 * Every function returns 0 / a non-negative count on success, negative on error
 * (dcrph_last_error() has the text).
 */
#ifndef DCRP_HOST_H
#define DCRP_HOST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* dcrph_last_error(void);
const char* dcrph_reference_aircraft_csv(void);      /* Monte_Carlo.cpp:11 file name */

/* Seeded synthetic ADS-B day in the reference's 22-column schema (synth_day.h). */
int dcrph_synth_day_write(const char* path, int n_flights, uint64_t seed, int trailing_dummy);
/* The same with every arrival_every-th flight an ARRIVAL (first altitude != 0: the depart_or_arrive = 1
 * branch of Drone.cpp:47-52,138-141,152-155); arrival_every = 0 writes departures only. */
int dcrph_synth_day_write_mixed(const char* path, int n_flights, uint64_t seed, int trailing_dummy,
                                int arrival_every);

/* Drone start ring of any radius in the format of drone_inital_positions_{1km,5km}.csv (synth_day.h). */
int dcrph_ring_write(const char* path, double radius_m, int n_rows);

/* Aircraft::Set_Parameters_and_Data + Vector_Allocation(index) + Deallocation.
 * Writes Vector_length, takeoff_t, Aircraft_Index_size and up to `cap` samples
 * of the four column arrays; index < 0 only reports Aircraft_Index_size.
 * Returns the number of samples copied. */
int dcrph_aircraft_load(const char* path, int n_cols, int index, int* vector_length, int* takeoff_t,
                        int* index_size, double* x, double* y, double* z, double* gs, int cap);

/* The drone row Drone::SetInitialParameters binds (Drone.cpp:33-35):
 * out = {longitude, latitude, heading} of row drone_index+1. */
int dcrph_ring_row(const char* path, int total_cols, int drone_index, double out[3]);

/* The loop of Monte_Carlo.cpp:26-38 over aircraft [aircraft_begin, aircraft_end)
 * and drones [0, drone_total), through the Aircraft and Drone classes
 * (one Drone::Simulation call per pair, output under <out_dir>/Drone_coords_<i>.csv).
 * total_collisions is the reference's double counter; exact_total the integer one. */
int dcrph_monte_carlo(const char* aircraft_csv, int aircraft_cols, int aircraft_begin, int aircraft_end,
                      const char* drone_csv, int drone_cols, int drone_total, int number_runs,
                      double aircraft_radius, double drone_radius, uint64_t seed, int precision,
                      const char* out_dir, double* total_collisions, uint64_t* exact_total);

/* The same workload in ONE dcrp_simulate call (all flights x all drones; extension).  Identical counts
 * and files to dcrph_monte_carlo with the same seed. */
int dcrph_monte_carlo_batch(const char* aircraft_csv, int aircraft_cols, int aircraft_begin, int aircraft_end,
                            const char* drone_csv, int drone_cols, int drone_total, uint64_t number_runs,
                            double aircraft_radius, double drone_radius, uint64_t seed, int precision,
                            const char* out_dir, uint64_t* total_collisions);

/* DCRP_FLAG_* (dcrp.h) applied to every later dcrph_monte_carlo / dcrph_monte_carlo_batch call of the
 * process, e.g. DCRP_FLAG_REACH_CULL.  Extension: the reference's main() has no such knob. */
void dcrph_set_run_flags(uint32_t flags);
/* Devices later dcrph_monte_carlo_batch calls split their trial range over, inside this one process (1 = default,
 * 0 = all visible).  Same counts and files as one device: the Philox counters carry the global trial id. */
void dcrph_set_gpus(int n_gpus);

/* Drone::Output's row format (Drone.cpp:237-241) applied to one trajectory;
 * appends to `path`. */
int dcrph_write_block(const char* path, const double* x, const double* y, const double* z, int n,
                      long run_no);

#ifdef __cplusplus
}
#endif
#endif
