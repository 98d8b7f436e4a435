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

}  // namespace dcrp_host

#endif
