// Aircraft.h -- host-side ADS-B day loader with the reference's public surface
// (reference: Aircraft.h:36-48, Aircraft.cpp:10-136).  Arithmetic-free: it stays
// on the host; the GPU engine consumes the arrays it produces.
//
// Same members and call order as the reference:
//   Set_Parameters_and_Data(path, n_cols) -> Vector_Allocation(i) -> ... -> Deallocation()
// Differences, all outside the reference's defined behaviour:
//   * index N-1 (the last flight) is served instead of reading past the index
//     table (Aircraft.cpp:73-82 only reaches it through index == N, which still works);
//   * a missing file prints the reference's message and leaves an empty, safe
//     object (ok() == false) instead of running into undefined behaviour;
//   * takeoff_t never reads past the last row (Aircraft.cpp:102-110 does when the
//     on-ground flag never flips).
#ifndef DCRP_HOST_AIRCRAFT_H
#define DCRP_HOST_AIRCRAFT_H

#include <string>
#include <vector>

class Aircraft {
  public:
    void Set_Parameters_and_Data(std::string File_Path, int no_cols);
    void Vector_Allocation(int index_input);
    void Deallocation();

    int Vector_length = 0;
    int takeoff_t = 0;
    int Aircraft_Index_size = 0;

    double* longitude_vector = nullptr;
    double* latitude_vector = nullptr;
    double* altitude_vector = nullptr;
    double* groundspeed_vector = nullptr;

    // ---- extensions
    bool ok() const { return ok_; }
    const std::string& last_error() const { return error_; }
    int flight_count() const { return Aircraft_Index_size; }   // flights addressable as 0..N-1

  private:
    // 0-based CSV columns (Aircraft.cpp:14-24)
    int callsign_ = 3, longitude_ = 14, latitude_ = 13, altitude_ = 7, groundspeed_ = 8, onground_ = 15;
    std::string path_;
    int cols_ = 0;
    std::vector<std::string> cells_;      // every line split on ',' into one flat list, header included
    std::vector<size_t> flight_start_;    // first cell of each flight
    size_t sel_begin_ = 0, sel_end_ = 0;  // cell range of the selected flight
    bool ok_ = false;
    std::string error_;

    void read_cells();
    void index_flights();
    void extract_column(int column, double* out) const;
};

#endif
