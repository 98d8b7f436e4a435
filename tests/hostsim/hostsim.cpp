// hostsim.cpp -- TEST AID, never shipped: compiles the __host__ __device__ trial
// arithmetic of drone-collision-rp_b200/csrc/dcrp_math.cuh with g++ so that the
// CPU-only test tier can check the exact statements the kernels execute
// (Philox, draw mapping, fp64 ray, fp32 screen set-up) against the oracle.
// It mirrors the kernels' control flow around that header; it is not a product
// code path and libdcrp.so contains none of it.
#include <cmath>
#include <cstdint>
#include <climits>
#include <vector>

#include "dcrp_math.cuh"

using namespace dcrp;

extern "C" {

void hs_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    U4 o = philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
    out[0] = o.x; out[1] = o.y; out[2] = o.z; out[3] = o.w;
}

void hs_make_draw(uint64_t seed, uint64_t trial, uint32_t drone, uint32_t track, int32_t takeoff_t,
                  const double box[6], Draw* out)
{
    *out = make_draw(seed, trial, drone, track, takeoff_t, box);
}

// atan_nonneg over n inputs (tests/test_hostsim_math.py bounds its error against libm's fp64 atan)
void hs_atan_nonneg(const float* x, long n, float* out)
{
    for (long i = 0; i < n; ++i) out[i] = atan_nonneg(x[i]);
}

// FastDiv against the machine's own division; returns the number of mismatches over a strided sweep of n
long hs_fastdiv_mismatches(uint32_t d, uint32_t stride)
{
    const FastDiv f = make_fastdiv(d);
    long bad = 0;
    for (uint64_t n = 0; n <= 0xffffffffull; n += stride) bad += f.div((uint32_t)n) != (uint32_t)n / d;
    const uint32_t edge[] = {0u, 1u, d - 1u, d, d + 1u, 2u * d - 1u, 2u * d, 0x7fffffffu, 0x80000000u, 0xfffffffeu, 0xffffffffu};
    for (uint32_t n : edge) bad += f.div(n) != n / d;
    return bad;
}

double hs_thr_d2(double radius)
{
    double c = radius * radius;
    while (std::sqrt(c) >= radius) c = std::nextafter(c, 0.0);
    while (std::sqrt(c) < radius) c = std::nextafter(c, INFINITY);
    return c;
}

// k_stage1 + exact_trial for n draws of one (track, drone).
void hs_exact(const double* ax, const double* ay, const double* az, int n_steps, int takeoff_t,
              double x0, double y0, double heading0, double radius, double start_alt, double v_straight,
              double v_ascend, const Draw* draws, long n, int32_t* first_out, double* min_dist_out)
{
    const double thr = hs_thr_d2(radius);
    const double s1x = std::sin(heading0) * v_straight, s1y = std::cos(heading0) * v_straight;
    std::vector<double> p1x(takeoff_t), p1y(takeoff_t), pm(takeoff_t);
    int s1_first = INT_MAX;
    double x = x0, y = y0, mn = INFINITY;
    for (int i = 0; i < takeoff_t; ++i) {
        if (i > 0) { x = DADD(x, s1x); y = DADD(y, s1y); }
        p1x[i] = x; p1y[i] = y;
        if (i < n_steps) {
            const double d2 = sep2(x, y, start_alt, ax[i], ay[i], az[i]);
            if (d2 < mn) mn = d2;
            if (d2 < thr && s1_first == INT_MAX) s1_first = i;
        }
        pm[i] = mn;
    }
    for (long r = 0; r < n; ++r) {
        const Draw& d = draws[r];
        const int last = d.t1st - 1;
        const Ray64 ray = stage2_ray(p1x[last], p1y[last], start_alt, d.tx, d.ty, d.tz, v_straight, v_ascend);
        int first = (s1_first <= last) ? s1_first : -1;
        double m = pm[last];
        double px = ray.x, py = ray.y, pz = ray.z;
        for (int i = d.t1st; i < n_steps; ++i) {
            px = DADD(px, ray.sx); py = DADD(py, ray.sy); pz = DADD(pz, ray.sz);
            const double d2 = sep2(px, py, pz, ax[i], ay[i], az[i]);
            if (d2 < m) m = d2;
            if (d2 < thr && first < 0) first = i;
        }
        first_out[r] = first;
        min_dist_out[r] = DSQRT(m);
    }
}

// The fp32 screen of k_screen2 for n draws of one (track, drone), scalar form of the packed
// sequence, relative to origin (ox, oy, oz):
//   level 1  min over stage-2 indices of the HORIZONTAL squared separation, incremental scan
//            (exact re-base P = lo*S + B at every tile start, then D = P + (-A), P += S)
//   level 2  min of the full 3-D squared separation, fused form P = i*S + B
void hs_screen(const double* ax, const double* ay, const double* az, int n_steps, int takeoff_t,
               double x0, double y0, double heading0, double start_alt, double v_straight, double v_ascend,
               double ox, double oy, double oz, const Draw* draws, long n, float* min_d2xy_out,
               float* min_d2_out)
{
    const double s1x = std::sin(heading0) * v_straight, s1y = std::cos(heading0) * v_straight;
    std::vector<double> p1x(takeoff_t), p1y(takeoff_t);
    double x = x0, y = y0;
    for (int i = 0; i < takeoff_t; ++i) {
        if (i > 0) { x = DADD(x, s1x); y = DADD(y, s1y); }
        p1x[i] = x; p1y[i] = y;
    }
    std::vector<float> nx(n_steps), ny(n_steps), nz(n_steps);
    for (int i = 0; i < n_steps; ++i) {
        nx[i] = -(float)(ax[i] - ox); ny[i] = -(float)(ay[i] - oy); nz[i] = -(float)(az[i] - oz);
    }
    const float coef = (float)(2.0 * (v_straight - v_ascend) / DCRP_PI);
    const int tile = n_steps <= 128 ? 128 : 512;
    for (long r = 0; r < n; ++r) {
        const Draw& d = draws[r];
        const int last = d.t1st - 1;
        const double xl = p1x[last], yl = p1y[last];
        const Ray32 ray = stage2_ray32((float)DSUB(d.tx, xl), (float)DSUB(d.ty, yl), (float)DSUB(d.tz, start_alt),
                                       (float)DSUB(xl, ox), (float)DSUB(yl, oy), (float)DSUB(start_alt, oz),
                                       (float)last, (float)v_straight, coef);
        float mn2 = 3.0e38f, mn3 = 3.0e38f;
        for (int t0 = 0; t0 < n_steps; t0 += tile) {
            const int lo = d.t1st > t0 ? d.t1st : t0, hi = (t0 + tile < n_steps) ? t0 + tile : n_steps;
            if (lo >= hi) continue;
            float px = fmaf((float)lo, ray.sx, ray.bx), py = fmaf((float)lo, ray.sy, ray.by);
            for (int i = lo; i < hi; ++i) {
                const float dx = px + nx[i], dy = py + ny[i];
                mn2 = fminf(mn2, fmaf(dy, dy, dx * dx));
                px += ray.sx; py += ray.sy;
            }
        }
        for (int i = d.t1st; i < n_steps; ++i) {
            const float fi = (float)i;
            const float dx = fmaf(fi, ray.sx, ray.bx) + nx[i];
            const float dy = fmaf(fi, ray.sy, ray.by) + ny[i];
            const float dz = fmaf(fi, ray.sz, ray.bz) + nz[i];
            mn3 = fminf(mn3, fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
        }
        min_d2xy_out[r] = ray.degenerate ? 0.f : mn2;
        min_d2_out[r] = ray.degenerate ? 0.f : mn3;
    }
}

// Level 1a of k_screen3 for n draws of one (track, drone) along the unit vector (nx, ny): the scalar form of the packed
// sequence -- projected track table T_i = n . (-A_i) (index 0 never passes), projected kink table, slab_ray32, scan of
// ALL indices from 0 (the ray extrapolated backwards before the kink), D = P + T_i, P += S, exact re-base every 128.
// Returns min over i >= 1 of |n . D(i)| per draw (0 for a degenerate ray).
void hs_slab(const double* ax, const double* ay, int n_steps, int takeoff_t, double x0, double y0, double heading0,
             double start_alt, double v_straight, double v_ascend, double ox, double oy, double nx, double ny,
             const Draw* draws, long n, float* min_abs_out)
{
    const double s1x = std::sin(heading0) * v_straight, s1y = std::cos(heading0) * v_straight;
    std::vector<double> p1x(takeoff_t), p1y(takeoff_t);
    std::vector<float> pln(takeoff_t), T(n_steps);
    const float fnx = (float)nx, fny = (float)ny;
    double x = x0, y = y0;
    for (int i = 0; i < takeoff_t; ++i) {
        if (i > 0) { x = DADD(x, s1x); y = DADD(y, s1y); }
        p1x[i] = x; p1y[i] = y;
        pln[i] = fmaf(fnx, (float)DSUB(x, ox), fny * (float)DSUB(y, oy));
    }
    for (int i = 0; i < n_steps; ++i)
        T[i] = i == 0 ? 1e15f : fmaf(fnx, -(float)(ax[i] - ox), fny * -(float)(ay[i] - oy));
    const float coef = (float)(2.0 * (v_straight - v_ascend) / DCRP_PI);
    for (long r = 0; r < n; ++r) {
        const Draw& d = draws[r];
        const int last = d.t1st - 1;
        const Slab32 u = slab_ray32((float)DSUB(d.tx, p1x[last]), (float)DSUB(d.ty, p1y[last]),
                                    fabsf((float)DSUB(d.tz, start_alt)), pln[last], (float)last, fnx, fny,
                                    (float)v_straight, coef);
        float mn = 3.0e38f;
        for (int t0 = 0; t0 < n_steps; t0 += 128) {
            float p = fmaf((float)t0, u.s, u.b);
            const int hi = t0 + 128 < n_steps ? t0 + 128 : n_steps;
            for (int i = t0; i < hi; ++i) { mn = fminf(mn, fabsf(p + T[i])); p += u.s; }
        }
        min_abs_out[r] = u.degenerate ? 0.f : mn;
    }
}

}  // extern "C"
