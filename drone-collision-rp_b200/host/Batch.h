// Batch.h -- whole-day driver on the batch ABI (extension; the reference loops one (aircraft, drone)
// pair at a time, Monte_Carlo.cpp:28-34).  Every flight [aircraft_begin, aircraft_end) of the CSV and
// every drone [0, drone_total) go to the GPU in ONE dcrp_simulate call; results are identical to the
// per-pair loop (same Philox ids) and the collision files are written in the reference's format and
// order (Drone.cpp:233-243: blocks by drone, then run).
#ifndef DCRP_HOST_BATCH_H
#define DCRP_HOST_BATCH_H

#include <cstdint>
#include <string>

#include "dcrp.h"

namespace dcrp_host {

struct BatchResult {
    uint64_t collisions = 0;
    uint64_t trials = 0;
    int status = 0;
    std::string error;
    dcrp_stats stats{};
};

// One PROCESS per GPU (mpirun / a shell loop / any launcher): rank r of n_ranks runs the contiguous trial range
// [r n / R, (r+1) n / R) of every (track, drone) on device `rank % visible devices` (or DCRP_DEVICE), the per-pair
// counts are summed over the ranks by the engine's communicator (dcrp_engine_attach_comm / dcrp_allreduce_counts;
// no torch, no MPI), so every rank reports the GLOBAL count; the compact collision records travel to rank 0
// through files next to the rendezvous file and rank 0 writes the Drone_coords_<i>.csv files.  `rendezvous` is
// a path all ranks can write (rank 0 publishes the 128-byte communicator id there).  n_ranks <= 1: ignored.
struct RankSpec {
    int rank = 0, n_ranks = 1;
    std::string rendezvous;
};

BatchResult monte_carlo_batch(const std::string& aircraft_csv, int aircraft_cols, int aircraft_begin,
                              int aircraft_end, const std::string& drone_csv, int drone_cols, int drone_total,
                              uint64_t number_runs, double aircraft_radius, double drone_radius, uint64_t seed,
                              int precision, const std::string& out_dir, uint32_t run_flags = 0, int n_gpus = 1,
                              const RankSpec& ranks = RankSpec());

}  // namespace dcrp_host

#endif
