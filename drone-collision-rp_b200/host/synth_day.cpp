// synth_day.cpp -- see synth_day.h.
#include "synth_day.h"

#include <cmath>
#include <cstdio>

namespace dcrp_host {

const char* const kReferenceAircraftCsv =
    "TAKEOFF_UTM_DEPART_A320 LHR time - 2017-06-30 04-00 to 2017-06-30 23-45.csv";

namespace {
struct SplitMix64 {
    uint64_t s;
    explicit SplitMix64(uint64_t seed) : s(seed) {}
    uint64_t next()
    {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double unit() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    int range(int lo, int hi) { return lo + (int)(next() % (uint64_t)(hi - lo + 1)); }
};
}  // namespace

int write_synthetic_day(const std::string& path, int n_flights, uint64_t seed, bool trailing_dummy)
{
    return write_synthetic_day_mixed(path, n_flights, seed, trailing_dummy, 0);
}

// One arrival (Drone.cpp:47-52: first altitude != 0 selects the arrival box of Drone.cpp:138-141,152-155):
// a 3-degree approach to runway 27L / 27R from the east at constant speed, touchdown near the eastern
// threshold, roll-out with the on-ground flag set.  Aircraft::Takeoff_Time (Aircraft.cpp:102-110) then
// yields the number of airborne rows minus one, i.e. the drone kinks while the aircraft is on approach.
static void write_arrival(FILE* f, SplitMix64& rng, int k, long& row)
{
    const int n = rng.range(92, 105);
    const int air = rng.range(40, 60);                    // airborne rows before touchdown
    const bool north = (k & 1) != 0;
    const double centre = north ? 5705983.0 : 5704545.0;
    const double northing0 = centre + (rng.unit() * 16.0 - 8.0);
    const double vapp = 66.0 + 8.0 * rng.unit();          // approach speed, m/s
    const double sink = vapp * 0.0524;                    // 3 degree glide path
    const double decel = 1.8 + 0.6 * rng.unit();
    const double drift = rng.unit() - 0.5;
    const double touchdown = 678300.0 + (rng.unit() * 60.0 - 30.0);
    double easting = touchdown + vapp * (double)air;
    double v = vapp;
    char callsign[16];
    std::snprintf(callsign, sizeof callsign, "ARR%04d", k);
    for (int i = 0; i < n; ++i, ++row) {
        const bool on_ground = i >= air;
        const double alt = on_ground ? 0.0 : sink * (double)(air - i);
        const double northing = northing0 + (on_ground ? 0.0 : drift * (double)(air - i));
        std::fprintf(f, "%ld,2017-06-30 %02d:%02d:%02d,4ca%03x,%s,XXXX,EGLL,A320,", row, 4 + (k / 60) % 20, k % 60,
                     i % 60, k & 0xfff, callsign);
        if (on_ground) std::fprintf(f, "0,");
        else std::fprintf(f, "%.3f,", alt);
        std::fprintf(f, "%.3f,270.0,%.1f,1000,False,%.3f,%.3f,%s,%ld,%d,%.3f,False,%d,%d\n", v * 1.943844,
                     on_ground ? 0.0 : -sink * 196.85, northing, easting, on_ground ? "True" : "False",
                     1498795200L + row, 4 + (k / 60) % 20, alt, k, k + 1);
        if (on_ground) { v -= decel; if (v < 8.0) v = 8.0; }
        easting -= v;
    }
}

int write_synthetic_day_mixed(const std::string& path, int n_flights, uint64_t seed, bool trailing_dummy, int arrival_every)
{
    FILE* f = std::fopen(path.c_str(), "w");
    if (!f) return -1;
    // 22 header cells; only columns 3, 7, 8, 13, 14, 15 are ever read (Aircraft.cpp:14-24)
    std::fprintf(f, "index,timestamp,icao24,callsign,origin,destination,typecode,altitude,groundspeed,"
                    "track,vertical_rate,squawk,spi,latitude,longitude,onground,last_position,hour,"
                    "geoaltitude,alert,firstseen,lastseen\n");
    const int total = n_flights + (trailing_dummy ? 1 : 0);
    long row = 0;
    for (int k = 0; k < total; ++k) {
        SplitMix64 rng(seed * 0x100000001B3ull + (uint64_t)k * 0x9E3779B97F4A7C15ull + 1);
        if (arrival_every > 0 && k < n_flights && (k % arrival_every) == arrival_every - 1) {
            write_arrival(f, rng, k, row);
            continue;
        }
        const int n = rng.range(92, 105);
        const int ground = rng.range(24, 34);
        const bool north = (k & 1) != 0;                      // 27R when odd, 27L when even
        const double centre = north ? 5705983.0 : 5704545.0;  // inside the bands of Drone.cpp:135,148
        const double northing0 = centre + (rng.unit() * 16.0 - 8.0);
        const double accel = 2.4 + 0.4 * rng.unit();
        const double vmax = 80.0 + 10.0 * rng.unit();
        const double climb = 8.0 + 3.0 * rng.unit();
        const double drift = rng.unit() - 0.5;
        double easting = 678500.0 + (rng.unit() * 60.0 - 30.0);
        double v = 0.0;
        char callsign[16];
        std::snprintf(callsign, sizeof callsign, "SYN%04d", k);
        for (int i = 0; i < n; ++i, ++row) {
            const bool on_ground = i < ground;
            const double alt = on_ground ? 0.0 : climb * (double)(i - ground + 1);
            const double northing = northing0 + (on_ground ? 0.0 : drift * (double)(i - ground + 1));
            std::fprintf(f, "%ld,2017-06-30 %02d:%02d:%02d,4ca%03x,%s,EGLL,XXXX,A320,", row, 4 + (k / 60) % 20,
                         k % 60, i % 60, k & 0xfff, callsign);
            if (on_ground) std::fprintf(f, "0,");            // exactly 0: departures, Drone.cpp:47-52
            else std::fprintf(f, "%.3f,", alt);
            std::fprintf(f, "%.3f,270.0,%.1f,1000,False,%.3f,%.3f,%s,%ld,%d,%.3f,False,%d,%d\n", v * 1.943844,
                         on_ground ? 0.0 : climb * 196.85, northing, easting, on_ground ? "True" : "False",
                         1498795200L + row, 4 + (k / 60) % 20, alt, k, k + 1);
            v += accel;
            if (v > vmax) v = vmax;
            easting -= v;
        }
    }
    std::fclose(f);
    return 0;
}

int write_ring_csv(const std::string& path, double radius_m, int n_rows)
{
    if (n_rows < 2 || !(radius_m > 0)) return -1;
    FILE* f = std::fopen(path.c_str(), "w");
    if (!f) return -1;
    const double pi = 3.14159265358979323846;
    const double cx = 676346.3657, cy = 5705295.6395;      // centre of both shipped rings
    // calibrated on the shipped rings: grid convergence, and slightly different north / east scales
    const double convergence = 0.0347, scale_n = 1.000535, scale_e = 1.003159;
    std::fprintf(f, ",longitude,latitude,heading\r\n");
    for (int k = 0; k < n_rows; ++k) {
        const double bearing = 2.0 * pi * (double)(k % (n_rows - 1)) / (double)(n_rows - 1);
        const double grid = bearing - convergence;
        std::fprintf(f, "%d,%.10f,%.10f,%.15f\r\n", k, cx + scale_e * radius_m * std::sin(grid),
                     cy + scale_n * radius_m * std::cos(grid), std::fmod(bearing + pi, 2.0 * pi));
    }
    std::fclose(f);
    return 0;
}

}  // namespace dcrp_host
