// Monte_Carlo.cpp -- the reference driver (Monte_Carlo.cpp:8-39) on the GPU engine.
// With no arguments it does what the reference's main() does, with the same
// constants: first departure of the LHR day x 99 drones of the 1 km ring x
// 10 000 trials, "Number of collisions: N" on stdout, collision blocks in
// Drone_Collisions/Drone_coords_0.csv.  The constants the reference makes you
// recompile for (Monte_Carlo.cpp:11-18,28,33) are flags here.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <sys/stat.h>

#include "Aircraft.h"
#include "Batch.h"
#include "Drone.h"
#include "synth_day.h"

using namespace std;

static void usage(const char* argv0)
{
    cout << "usage: " << argv0 << " [--aircraft-csv F] [--drone-csv F] [--aircraft N] [--drones N]\n"
            "       [--runs N] [--seed S] [--precision fp64|mixed|fp32] [--synthetic N_FLIGHTS [--arrival-every K]] [--batch [--gpus N]] [--cull]\n"
            "       [--batch --ranks R --rank r --rendezvous FILE]   one process per GPU, counts all-reduced by the engine\n"
            "defaults are the reference's constants (Monte_Carlo.cpp:11-18,28,33)\n";
}

int main(int argc, char** argv)
{
    // Monte_Carlo.cpp:11-18
    string FilePath_aircraft = dcrp_host::kReferenceAircraftCsv;
    int no_col_aircraft = 22;
    double aircraft_radius = 3.5;
    string FilePath_drone = "drone_inital_positions_1km.csv";
    int no_col_drone = 4;
    int drone_total = 99;
    double drone_radius = 0.19;
    int aircraft_total = 1;          // Monte_Carlo.cpp:28
    long long number_runs = 10000;   // Monte_Carlo.cpp:33 (an int there; widened: BASELINE config 3 is 1e10 trials)
    uint64_t seed = 20170630ull;
    int precision = DCRP_PRECISION_MIXED;
    int synthetic = 0, arrival_every = 0;
    bool batch = false;          // one GPU launch for all (aircraft, drone) pairs instead of one per pair
    int n_gpus = 1;              // --gpus N (0 = all): --batch only, trial ranges split over the devices of this process
    dcrp_host::RankSpec ranks;   // --ranks R --rank r --rendezvous FILE: --batch only, one PROCESS per GPU
    uint32_t run_flags = 0;      // --cull: DCRP_FLAG_REACH_CULL (exact culling by reachability, same results)

    for (int i = 1; i < argc; ++i) {
        const string a = argv[i];
        auto next = [&](const char* name) -> const char* {
            if (i + 1 >= argc) { cout << "ERROR - " << name << " needs a value" << endl; exit(2); }
            return argv[++i];
        };
        if (a == "--aircraft-csv") FilePath_aircraft = next("--aircraft-csv");
        else if (a == "--drone-csv") FilePath_drone = next("--drone-csv");
        else if (a == "--aircraft") aircraft_total = atoi(next("--aircraft"));
        else if (a == "--drones") drone_total = atoi(next("--drones"));
        else if (a == "--runs") number_runs = atoll(next("--runs"));
        else if (a == "--seed") seed = strtoull(next("--seed"), nullptr, 10);
        else if (a == "--synthetic") synthetic = atoi(next("--synthetic"));
        else if (a == "--arrival-every") arrival_every = atoi(next("--arrival-every"));
        else if (a == "--batch") batch = true;
        else if (a == "--cull") run_flags |= DCRP_FLAG_REACH_CULL;
        else if (a == "--gpus") n_gpus = atoi(next("--gpus"));
        else if (a == "--ranks") ranks.n_ranks = atoi(next("--ranks"));
        else if (a == "--rank") ranks.rank = atoi(next("--rank"));
        else if (a == "--rendezvous") ranks.rendezvous = next("--rendezvous");
        else if (a == "--precision") {
            const string p = next("--precision");
            if (p == "fp64") precision = DCRP_PRECISION_FP64;
            else if (p == "mixed") precision = DCRP_PRECISION_MIXED;
            else if (p == "fp32") precision = DCRP_PRECISION_FP32;
            else { cout << "ERROR - unknown precision " << p << endl; return 2; }
        } else { usage(argv[0]); return a == "--help" ? 0 : 2; }
    }

    if (synthetic > 0) {
        ifstream probe(FilePath_aircraft);
        if (!probe) {
            if (dcrp_host::write_synthetic_day_mixed(FilePath_aircraft, synthetic, seed, true, arrival_every) != 0) {
                cout << "ERROR - cannot write " << FilePath_aircraft << endl;
                return 1;
            }
            cout << "wrote synthetic day (" << synthetic << " flights) to " << FilePath_aircraft << endl;
        }
    }
    mkdir("Drone_Collisions", 0777);   // the reference needs it to pre-exist (Drone.cpp:62,236)

    if (batch) {
        const dcrp_host::BatchResult r = dcrp_host::monte_carlo_batch(
            FilePath_aircraft, no_col_aircraft, 0, aircraft_total, FilePath_drone, no_col_drone, drone_total,
            (uint64_t)number_runs, aircraft_radius, drone_radius, seed, precision, "Drone_Collisions", run_flags, n_gpus, ranks);
        if (r.status != 0) { cout << "ERROR - " << r.error << endl; return 1; }
        if (ranks.n_ranks > 1) cout << "rank " << ranks.rank << " of " << ranks.n_ranks << ": ";
        cout << "Number of collisions: " << (double)r.collisions << endl;
        dcrp_host::release_shared_engine();
        return 0;
    }

    double* total_collisions = new double[1];
    *total_collisions = 0;

    Aircraft Aircraft;
    Drone Drone;
    Drone.SetSeed(seed);
    Drone.SetPrecision(precision);
    Drone.SetRunFlags(run_flags);

    Aircraft.Set_Parameters_and_Data(FilePath_aircraft, no_col_aircraft);
    if (!Aircraft.ok()) {
        cout << "hint: the reference's aircraft CSV is not distributed; pass --synthetic 16 to generate a "
                "schema-compatible stand-in" << endl;
        return 1;
    }

    for (int i = 0; i < aircraft_total; ++i) {
        Aircraft.Vector_Allocation(i);
        if (Aircraft.Vector_length <= 0) { cout << "ERROR - no flight " << i << " in " << FilePath_aircraft << endl; return 1; }
        Drone.ClearOutput(i);
        for (int j = 0; j < drone_total; ++j) {
            Drone.SetInitialParameters(FilePath_drone, Aircraft.Vector_length, no_col_drone, j, i, Aircraft.takeoff_t,
                                       Aircraft.longitude_vector, Aircraft.latitude_vector, Aircraft.altitude_vector,
                                       aircraft_radius, drone_radius);
            Drone.Simulation64(number_runs > 0 ? (uint64_t)number_runs : 0, total_collisions);
            if (Drone.LastStatus() != 0) return 1;
        }
        Aircraft.Deallocation();
    }
    cout << "Number of collisions: " << *total_collisions << endl;
    delete[] total_collisions;
    dcrp_host::release_shared_engine();
    return 0;
}
