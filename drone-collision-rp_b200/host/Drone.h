// Drone.h -- the reference's Drone public surface (Drone.h:71-81) with the trial
// loop moved to the GPU engine behind include/dcrp.h.
//
//   SetInitialParameters(...)  Drone.cpp:11-57   binds one (aircraft, drone) pair
//   Simulation(n, &total)      Drone.cpp:269-281 n trials -> *total += hits, collision
//                                                blocks appended to Drone_Collisions/
//   ClearOutput(i)             Drone.cpp:59-64   truncates Drone_Collisions/Drone_coords_i.csv
//
// The per-trial members of the reference (SetInitialConditions, FirstStage,
// CubedVolume, SecondStage, Collision, Deallocate) have no host counterpart: they
// run as CUDA kernels (csrc/kernels.cu).  There is no CPU trial path here; if no
// sm_100 GPU is usable Simulation reports the engine's error and counts nothing.
#ifndef DCRP_HOST_DRONE_H
#define DCRP_HOST_DRONE_H

#include <cstdint>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "dcrp.h"

class Drone {
  public:
    Drone();
    ~Drone();
    Drone(const Drone&) = delete;
    Drone& operator=(const Drone&) = delete;

    void SetInitialParameters(std::string FilePath_input, int Vector_length_input, int TotalCols_input,
                              int drone_index_input, int aircraft_index_input, int takeoff_time_input,
                              double* air_long, double* air_lat, double* air_alt,
                              double aircraft_radius_input, double drone_radius_input);
    void Simulation(int number_runs, double* total_collisions);
    void ClearOutput(int Aircraft_Index);

    // In the reference these hold the trajectory of the trial in flight and are
    // freed at the end of every trial (Drone.cpp:278).  Here they point at the
    // trajectory of the LAST collision recorded by the latest Simulation call
    // (Vector_length entries each), or are null when that call recorded none.
    // speed_vector / heading_vector are never consumed by the reference
    // (Drone.cpp:126,228-229 only write them) and stay null.
    double* longitude_vector = nullptr;
    double* latitude_vector = nullptr;
    double* altitude_vector = nullptr;
    double* speed_vector = nullptr;
    double* heading_vector = nullptr;

    // ---- extensions (not in the reference)
    void SetSeed(uint64_t seed) { seed_ = seed; }
    void SetPrecision(int precision) { precision_ = precision; }     // DCRP_PRECISION_*
    void SetRunFlags(uint32_t flags) { run_flags_ = flags; }         // DCRP_FLAG_* (e.g. DCRP_FLAG_REACH_CULL)
    void SetOutputDirectory(const std::string& dir) { out_dir_ = dir; }
    void Simulation64(uint64_t number_runs, double* total_collisions);   // number_runs is int in Drone.h:74
    uint64_t ExactCollisions() const { return exact_total_; }        // the double counter loses integers past 2^53
    const dcrp_stats& LastStats() const { return last_stats_; }
    void InitialRow(double out[3]) const { out[0] = x0_; out[1] = y0_; out[2] = heading0_; }
    int LastStatus() const { return last_status_; }
    const std::string& LastError() const { return last_error_; }

  private:
    // bound pair (Drone.cpp:12-56)
    std::string file_path_;
    int vector_length_ = 0, total_cols_ = 0, drone_index_ = 0, aircraft_index_ = 0, takeoff_time_ = 0;
    double aircraft_radius_ = 0, drone_radius_ = 0;
    double x0_ = 0, y0_ = 0, heading0_ = 0;
    std::vector<double> air_x_, air_y_, air_z_;       // deep copies, Drone.cpp:37-45
    bool bound_ = false;

    // drone CSV cache: the reference re-reads the file on every call (Drone.cpp:27)
    std::string cached_path_;
    std::vector<std::string> cells_;

    uint64_t seed_ = 20170630ull;
    int precision_ = DCRP_PRECISION_MIXED;
    uint32_t run_flags_ = 0;
    std::string out_dir_ = "Drone_Collisions";
    std::map<std::pair<int, int>, uint64_t> next_trial_;   // trials already consumed per (aircraft, drone)
    uint64_t exact_total_ = 0;
    dcrp_stats last_stats_{};
    int last_status_ = 0;
    std::string last_error_;
    std::vector<double> last_traj_;
    std::vector<dcrp_record> records_;                 // record buffer of Simulation, allocated once

    bool load_cells(const std::string& path);
    void fail(int status, const std::string& what);
};

namespace dcrp_host {
// Process-wide engine on device $DCRP_DEVICE (default 0), created on first use.
void set_shared_device(int device);            // before the first shared_engine(): which CUDA device it is created on
dcrp_engine* shared_engine(std::string* error);
void release_shared_engine();
// Drone::Output row format (Drone.cpp:237-241): precision(10), "x,y,z,run\n" per
// index and one blank line per block.
void write_block(std::ostream& out, const double* x, const double* y, const double* z, int n, long run_no);
}  // namespace dcrp_host

#endif
