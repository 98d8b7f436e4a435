/*
 * ref_aircraft_harness.cpp -- exposes the UNMODIFIED reference Aircraft loader
 * (Aircraft.cpp:10-136) through a C entry so tests can compare the new host
 * loader against it on the same CSV.  TEST INFRASTRUCTURE ONLY.
 */
#include "Aircraft.cpp"    /* resolved through -I/root/reference */

extern "C" int ref_aircraft_load(const char* path, int ncols, int index,
                                 int* vector_length, int* takeoff_t, int* index_size,
                                 double* x, double* y, double* z, double* gs, int cap)
{
    Aircraft a;
    a.Set_Parameters_and_Data(std::string(path), ncols);
    *index_size = a.Aircraft_Index_size;
    if (index < 0) return 0;
    a.Vector_Allocation(index);
    *vector_length = a.Vector_length;
    *takeoff_t = a.takeoff_t;
    int n = a.Vector_length < cap ? a.Vector_length : cap;
    for (int i = 0; i < n; ++i) {
        x[i] = a.longitude_vector[i];
        y[i] = a.latitude_vector[i];
        z[i] = a.altitude_vector[i];
        gs[i] = a.groundspeed_vector[i];
    }
    a.Deallocation();
    return n;
}
