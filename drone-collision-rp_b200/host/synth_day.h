// synth_day.h -- seeded synthetic ADS-B day in the reference's 22-column schema.
//
// The real input, "TAKEOFF_UTM_DEPART_A320 LHR time - 2017-06-30 04-00 to
// 2017-06-30 23-45.csv" (Monte_Carlo.cpp:11), is not distributed with the
// reference.  This generator writes a schema-compatible stand-in (SURVEY.md
// App. B: columns 3 call-sign, 7 altitude, 8 ground speed, 13 northing,
// 14 easting, 15 on-ground; header row; one contiguous block per flight) so the
// loader and the whole pipeline run unchanged; a user's real file drops in.
#ifndef DCRP_HOST_SYNTH_DAY_H
#define DCRP_HOST_SYNTH_DAY_H

#include <cstdint>
#include <string>

namespace dcrp_host {

// The reference's default aircraft file name (Monte_Carlo.cpp:11).
extern const char* const kReferenceAircraftCsv;

// Writes n_flights departures (alternating 27L / 27R, 92..105 rows at 1 Hz,
// 24..34 on-ground rows, altitude exactly 0 on the ground).  trailing_dummy
// appends one extra flight so that every real one is reachable by its natural
// index in the reference loader too (Aircraft.cpp:73-82 quirk).
// Returns 0, or -1 if the file cannot be written.
int write_synthetic_day(const std::string& path, int n_flights, uint64_t seed, bool trailing_dummy);

// The same day with every `arrival_every`-th flight (k % arrival_every == arrival_every - 1) replaced by an
// ARRIVAL: airborne first row (altitude != 0, so Drone.cpp:47-52 selects the arrival box of
// Drone.cpp:138-141,152-155), 3-degree approach from the east to 27L / 27R, 40..60 airborne rows, then the
// roll-out with the on-ground flag set.  arrival_every = 0 writes departures only (identical to the above).
int write_synthetic_day_mixed(const std::string& path, int n_flights, uint64_t seed, bool trailing_dummy,
                              int arrival_every);

// Drone start ring in the format of drone_inital_positions_{1km,5km}.csv (header
// ",longitude,latitude,heading", rows "k,easting,northing,heading", CRLF): n_rows points on a circle of
// `radius_m` metres about the Heathrow reference point, row k at true bearing 2*pi*k/(n_rows-1) from it
// (row n_rows-1 repeats row 0, as in the shipped files), heading = bearing + pi (towards the centre).
// (mod 2*pi).  Calibrated on the two shipped rings: grid convergence 0.0347 rad and
// north / east scales 1.000535 / 1.003159 reproduce them to < 0.1 % of the radius; headings to 1e-12.
// Returns 0, or -1 on I/O failure.
int write_ring_csv(const std::string& path, double radius_m, int n_rows);

}  // namespace dcrp_host

#endif
