/*
 * ref_asshipped_main.cpp -- times the reference's own CPU path, as shipped:
 * Aircraft.cpp and Drone.cpp compiled UNMODIFIED (no shim: every draw seeds a
 * fresh engine from std::random_device, Drone.cpp:113,170-172) and driven the
 * way Monte_Carlo.cpp:26-38 drives them.  TEST / BASELINE INFRASTRUCTURE ONLY.
 *
 * The reference is single-threaded and writes to a cwd-relative path
 * (Drone.cpp:62,236), so "all host cores" = nproc forked workers, each in its
 * own scratch cwd, drones dealt round-robin.
 *
 * usage: ref_asshipped <aircraft.csv> <ncols> <aircraft_index> <drone.csv>
 *                      <drone_total> <number_runs> <nproc> <scratch_dir>
 * prints one JSON line: trials, hits, vector_length, seconds, nproc.
 */
#include "Aircraft.h"
#include "Drone.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>
#include <limits.h>

static std::string abspath(const char* p)
{
    char buf[PATH_MAX];
    if (realpath(p, buf)) return std::string(buf);
    return std::string(p);
}

int main(int argc, char** argv)
{
    if (argc < 9) {
        fprintf(stderr, "usage: %s aircraft.csv ncols aircraft_index drone.csv drone_total "
                        "number_runs nproc scratch_dir\n", argv[0]);
        return 2;
    }
    std::string air_csv = abspath(argv[1]);
    int ncols = atoi(argv[2]);
    int aircraft_index = atoi(argv[3]);
    std::string drone_csv = abspath(argv[4]);
    int drone_total = atoi(argv[5]);
    int number_runs = atoi(argv[6]);
    int nproc = atoi(argv[7]);
    std::string scratch = abspath(argv[8]);
    if (nproc < 1) nproc = 1;

    Aircraft aircraft;
    aircraft.Set_Parameters_and_Data(air_csv, ncols);          /* Monte_Carlo.cpp:26 */
    aircraft.Vector_Allocation(aircraft_index);                /* Monte_Carlo.cpp:29 */

    auto t0 = std::chrono::steady_clock::now();
    int fds[256][2];
    pid_t pids[256];
    if (nproc > 256) nproc = 256;
    for (int w = 0; w < nproc; ++w) {
        if (pipe(fds[w]) != 0) return 3;
        pid_t pid = fork();
        if (pid == 0) {
            close(fds[w][0]);
            std::string dir = scratch + "/w" + std::to_string(w);
            mkdir(scratch.c_str(), 0777);
            mkdir(dir.c_str(), 0777);
            mkdir((dir + "/Drone_Collisions").c_str(), 0777);
            if (chdir(dir.c_str()) != 0) _exit(4);
            double* total = new double[1];
            *total = 0;
            Drone drone;
            drone.ClearOutput(aircraft_index);                 /* Monte_Carlo.cpp:30 */
            for (int j = w; j < drone_total; j += nproc) {     /* Monte_Carlo.cpp:31 */
                drone.SetInitialParameters(drone_csv, aircraft.Vector_length, 4, j, aircraft_index,
                                           aircraft.takeoff_t, aircraft.longitude_vector,
                                           aircraft.latitude_vector, aircraft.altitude_vector,
                                           3.5, 0.19);         /* Monte_Carlo.cpp:32 */
                drone.Simulation(number_runs, total);          /* Monte_Carlo.cpp:33 */
            }
            double v = *total;
            if (write(fds[w][1], &v, sizeof v) != (ssize_t)sizeof v) _exit(5);
            _exit(0);
        }
        pids[w] = pid;
        close(fds[w][1]);
    }
    double hits = 0;
    for (int w = 0; w < nproc; ++w) {
        double v = 0;
        if (read(fds[w][0], &v, sizeof v) == (ssize_t)sizeof v) hits += v;
        int st = 0;
        waitpid(pids[w], &st, 0);
    }
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("{\"trials\": %.0f, \"hits\": %.0f, \"vector_length\": %d, \"takeoff_t\": %d, "
           "\"seconds\": %.6f, \"nproc\": %d}\n",
           (double)drone_total * (double)number_runs, hits, aircraft.Vector_length,
           aircraft.takeoff_t, secs, nproc);
    return 0;
}
