// Aircraft.cpp -- see Aircraft.h.  Semantics follow the reference loader
// (Aircraft.cpp:33-128) token for token: one flat list of comma-separated cells
// with the header row included, a new flight wherever the call-sign cell of two
// consecutive rows differs, takeoff_t = number of leading rows whose on-ground
// cell equals the next row's.
#include "Aircraft.h"

#include <fstream>
#include <iostream>

void Aircraft::Set_Parameters_and_Data(std::string File_Path, int no_cols)
{
    path_ = File_Path;
    cols_ = no_cols;
    cells_.clear();
    flight_start_.clear();
    Aircraft_Index_size = 0;
    ok_ = false;
    error_.clear();
    if (cols_ <= 0) { error_ = "column count must be positive"; return; }
    read_cells();
    if (!ok_) return;
    index_flights();
}

// Aircraft.cpp:33-55.  std::getline(stream, cell, ',') semantics: a trailing
// empty field and an empty line contribute no cell.
void Aircraft::read_cells()
{
    std::ifstream in(path_);
    if (!in) {
        std::cout << "ERROR - Failed to open " << path_ << std::endl;   // Aircraft.cpp:38-40
        error_ = "failed to open " + path_;
        return;
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
    ok_ = true;
}

// Aircraft.cpp:58-69: a boundary after row i when call-sign(i) != call-sign(i+1).
void Aircraft::index_flights()
{
    const size_t rows = cells_.size() / (size_t)cols_;
    for (size_t i = 0; i + 1 < rows; ++i) {
        const size_t next = (i + 1) * (size_t)cols_ + (size_t)callsign_;
        if (next >= cells_.size()) break;
        if (cells_[i * (size_t)cols_ + (size_t)callsign_] != cells_[next]) flight_start_.push_back((i + 1) * (size_t)cols_);
    }
    Aircraft_Index_size = (int)flight_start_.size();
}

// Aircraft.cpp:113-128 (SingleAircraft :72-84, Takeoff_Time :102-110, ColumnSelect :86-99)
void Aircraft::Vector_Allocation(int index_input)
{
    Vector_length = 0;
    takeoff_t = 0;
    longitude_vector = latitude_vector = altitude_vector = groundspeed_vector = nullptr;
    const int n = Aircraft_Index_size;
    if (!ok_ || n == 0 || index_input < 0 || index_input > n) {
        if (ok_) error_ = "flight index out of range";
        sel_begin_ = sel_end_ = 0;
        return;
    }
    if (index_input >= n - 1) {            // last flight: natural index N-1 or the reference's N
        sel_begin_ = flight_start_[(size_t)n - 1];
        sel_end_ = cells_.size();
    } else {
        sel_begin_ = flight_start_[(size_t)index_input];
        sel_end_ = flight_start_[(size_t)index_input + 1];
    }
    const size_t rows = (sel_end_ - sel_begin_) / (size_t)cols_;
    Vector_length = (int)rows;

    for (size_t i = 0; i + 1 < rows; ++i) {
        const std::string& a = cells_[sel_begin_ + i * (size_t)cols_ + (size_t)onground_];
        const std::string& b = cells_[sel_begin_ + (i + 1) * (size_t)cols_ + (size_t)onground_];
        if (a != b) break;
        takeoff_t += 1;
    }

    longitude_vector = new double[rows];
    latitude_vector = new double[rows];
    altitude_vector = new double[rows];
    groundspeed_vector = new double[rows];
    extract_column(longitude_, longitude_vector);
    extract_column(latitude_, latitude_vector);
    extract_column(altitude_, altitude_vector);
    extract_column(groundspeed_, groundspeed_vector);
}

void Aircraft::extract_column(int column, double* out) const
{
    const size_t rows = (sel_end_ - sel_begin_) / (size_t)cols_;
    for (size_t i = 0; i < rows; ++i) out[i] = std::stod(cells_[sel_begin_ + i * (size_t)cols_ + (size_t)column]);
}

void Aircraft::Deallocation()               // Aircraft.cpp:130-136
{
    delete[] longitude_vector;
    delete[] latitude_vector;
    delete[] altitude_vector;
    delete[] groundspeed_vector;
    longitude_vector = latitude_vector = altitude_vector = groundspeed_vector = nullptr;
    sel_begin_ = sel_end_ = 0;
}
