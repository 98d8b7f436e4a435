// Drone.cpp -- see Drone.h.  Host side of the drop-in: argument marshalling and
// the collision-block writer.  Every per-trial computation happens in libdcrp.so.
#include "Drone.h"

#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <mutex>

namespace dcrp_host {

static std::mutex g_mu;
static dcrp_engine* g_engine = nullptr;
static int g_device = -1;          // set_shared_device; -1 = DCRP_DEVICE or 0

void set_shared_device(int device)
{
    std::lock_guard<std::mutex> lock(g_mu);
    g_device = device;
}

dcrp_engine* shared_engine(std::string* error)
{
    std::lock_guard<std::mutex> lock(g_mu);
    if (g_engine) return g_engine;
    int device = g_device >= 0 ? g_device : 0;
    if (g_device < 0) if (const char* d = std::getenv("DCRP_DEVICE")) device = std::atoi(d);
    dcrp_engine* e = nullptr;
    const int rc = dcrp_engine_create(device, &e);
    if (rc != DCRP_OK) {
        if (error) *error = dcrp_last_error();
        return nullptr;
    }
    g_engine = e;
    return g_engine;
}

void release_shared_engine()
{
    std::lock_guard<std::mutex> lock(g_mu);
    if (g_engine) dcrp_engine_destroy(g_engine);
    g_engine = nullptr;
}

void write_block(std::ostream& out, const double* x, const double* y, const double* z, int n, long run_no)
{
    const std::streamsize old = out.precision(10);               // Drone.cpp:237
    for (int i = 0; i < n; ++i) out << x[i] << "," << y[i] << "," << z[i] << "," << run_no << '\n';
    out << '\n';                                                  // Drone.cpp:241
    out.precision(old);
}

}  // namespace dcrp_host

Drone::Drone() = default;
Drone::~Drone() = default;

void Drone::fail(int status, const std::string& what)
{
    last_status_ = status;
    last_error_ = what;
    std::cout << "ERROR - " << what << std::endl;
}

// Drone.cpp:83-105: every line split on ',' into one flat list (header included).
bool Drone::load_cells(const std::string& path)
{
    if (path == cached_path_ && !cells_.empty()) return true;
    cells_.clear();
    cached_path_.clear();
    std::ifstream in(path);
    if (!in) {
        std::cout << "ERROR - Failed to open " << path << std::endl;      // Drone.cpp:88-90
        return false;
    }
    std::string line;
    while (std::getline(in, line)) {
        size_t pos = 0;
        while (pos < line.size()) {
            size_t comma = line.find(',', pos);
            if (comma == std::string::npos) comma = line.size();
            cells_.emplace_back(line, pos, comma - pos);
            pos = comma + 1;
        }
    }
    cached_path_ = path;
    return true;
}

void Drone::SetInitialParameters(std::string FilePath_input, int Vector_length_input, int TotalCols_input,
                                 int drone_index_input, int aircraft_index_input, int takeoff_time_input,
                                 double* air_long, double* air_lat, double* air_alt,
                                 double aircraft_radius_input, double drone_radius_input)
{
    bound_ = false;
    last_status_ = DCRP_OK;
    last_error_.clear();
    file_path_ = FilePath_input;
    vector_length_ = Vector_length_input;
    total_cols_ = TotalCols_input;
    drone_index_ = drone_index_input;
    aircraft_index_ = aircraft_index_input;
    takeoff_time_ = takeoff_time_input;
    aircraft_radius_ = aircraft_radius_input;
    drone_radius_ = drone_radius_input;

    if (!load_cells(file_path_)) { last_status_ = DCRP_ERR_IO; last_error_ = "failed to open " + file_path_; return; }
    // row DroneIndex+1 (row 0 is the header), columns 1..3 (Drone.cpp:19-21,33-35)
    const size_t base = (size_t)(drone_index_ + 1) * (size_t)total_cols_;
    if (drone_index_ < 0 || total_cols_ < 4 || base + 3 >= cells_.size()) {
        fail(DCRP_ERR_INVALID, "drone index " + std::to_string(drone_index_) + " not in " + file_path_);
        return;
    }
    try {
        x0_ = std::stod(cells_[base + 1]);
        y0_ = std::stod(cells_[base + 2]);
        heading0_ = std::stod(cells_[base + 3]);
    } catch (const std::exception&) {
        fail(DCRP_ERR_IO, "unparsable drone row " + std::to_string(drone_index_) + " in " + file_path_);
        return;
    }
    if (vector_length_ <= 0 || !air_long || !air_lat || !air_alt) {
        fail(DCRP_ERR_INVALID, "aircraft arrays missing or empty");
        return;
    }
    air_x_.assign(air_long, air_long + vector_length_);           // Drone.cpp:37-45
    air_y_.assign(air_lat, air_lat + vector_length_);
    air_z_.assign(air_alt, air_alt + vector_length_);
    bound_ = true;
}

void Drone::ClearOutput(int Aircraft_Index)                        // Drone.cpp:59-64
{
    std::ofstream out(out_dir_ + "/Drone_coords_" + std::to_string(Aircraft_Index) + ".csv",
                      std::ofstream::out | std::ofstream::trunc);
}

void Drone::Simulation(int number_runs, double* total_collisions)
{
    Simulation64(number_runs > 0 ? (uint64_t)number_runs : 0, total_collisions);
}

void Drone::Simulation64(uint64_t number_runs, double* total_collisions)
{
    longitude_vector = latitude_vector = altitude_vector = speed_vector = heading_vector = nullptr;
    if (!bound_) { if (last_status_ == DCRP_OK) fail(DCRP_ERR_STATE, "Simulation before a successful SetInitialParameters"); return; }
    if (number_runs == 0) return;

    std::string err;
    dcrp_engine* eng = dcrp_host::shared_engine(&err);
    if (!eng) { fail(DCRP_ERR_NO_DEVICE, err); return; }

    dcrp_track track{};
    track.x = air_x_.data(); track.y = air_y_.data(); track.z = air_z_.data();
    track.n_steps = vector_length_;
    track.takeoff_t = takeoff_time_;
    int rc = dcrp_target_box(air_y_[0], air_z_[0], track.box);    // Drone.cpp:129-161, :47-52
    if (rc != DCRP_OK) { fail(rc, dcrp_last_error()); return; }
    dcrp_drone drone{x0_, y0_, heading0_};
    dcrp_model model;
    dcrp_default_model(&model);                                   // Drone.cpp:54-56
    model.aircraft_radius = aircraft_radius_;
    model.drone_radius = drone_radius_;

    uint64_t& consumed = next_trial_[std::make_pair(aircraft_index_, drone_index_)];
    dcrp_run run{};
    run.seed = seed_;
    run.trial_begin = consumed;
    run.trial_count = number_runs;
    run.precision = precision_;
    run.flags = run_flags_ | DCRP_FLAG_NO_TIMING;                 // (the reference's interface has no timings to fill)
    run.record_capacity = 1u << 16;
    // one-pair scenario: the Philox counter still carries the real indices, so this
    // call draws exactly what a whole-day batch draws for (aircraft, drone)
    run.track_id_base = (uint32_t)aircraft_index_;
    run.drone_id_base = (uint32_t)drone_index_;

    uint64_t count = 0;
    // (a member: a fresh vector would zero 4 MB on every call, more host time than a 10 000-trial call takes on the GPU)
    if (records_.size() < run.record_capacity) records_.resize(run.record_capacity);
    std::vector<dcrp_record>& records = records_;
    dcrp_result res{};
    res.counts = &count;
    res.records = records.data();
    rc = dcrp_simulate(eng, &track, 1, &drone, 1, &model, &run, &res);
    if (rc != DCRP_OK) { fail(rc, dcrp_last_error()); return; }
    consumed += number_runs;
    last_stats_ = res.stats;
    exact_total_ += count;
    if (total_collisions) *total_collisions += (double)count;     // Drone.cpp:276

    if (res.n_records) {
        int stride = 0;
        last_traj_.assign((size_t)res.n_records * 3 * (size_t)vector_length_, 0.0);
        rc = dcrp_trajectories(eng, records.data(), res.n_records, last_traj_.data(), &stride);
        if (rc != DCRP_OK) { fail(rc, dcrp_last_error()); return; }
        std::ofstream out(out_dir_ + "/Drone_coords_" + std::to_string(aircraft_index_) + ".csv",
                          std::ofstream::out | std::ofstream::app);          // Drone.cpp:236
        for (uint32_t k = 0; k < res.n_records; ++k) {
            const double* x = last_traj_.data() + (size_t)k * 3 * (size_t)stride;
            const long run_no = (long)(records[k].trial - run.trial_begin);  // the reference's loop index i
            if (out) dcrp_host::write_block(out, x, x + stride, x + 2 * (size_t)stride, vector_length_, run_no);
        }
        const double* x = last_traj_.data() + (size_t)(res.n_records - 1) * 3 * (size_t)stride;
        longitude_vector = const_cast<double*>(x);
        latitude_vector = const_cast<double*>(x + stride);
        altitude_vector = const_cast<double*>(x + 2 * (size_t)stride);
    }
    if (res.overflow)
        std::cout << "WARNING - more than " << run.record_capacity
                  << " collisions in one Simulation call; extra blocks not written" << std::endl;
}
