// dcrp_math.cuh -- the per-trial arithmetic shared by every trial kernel.
//
// Everything here is __host__ __device__ so that tests/hostsim can compile the
// very same statements with g++ and compare them with the oracle on a machine
// without a GPU.  That host build is a TEST aid: libdcrp.so contains no CPU
// trial path (see engine.cu).
//
// Reference lines restated here (all /root/reference/Drone.cpp):
//   draws            :113-116 (int, [1,TakeoffTime]), :175-187 (three reals)
//   stage 1          :120-126   x += sin(h)*19, y += cos(h)*19, z = 100
//   stage 2 set-up   :193-220   |dL|,|dB|,|dA|, quadrant bearing, pitch, vf
//   stage 2 steps    :224-230   x += sin(b)*vf, y += cos(b)*vf, z += sin(p)*vf
//   separation test  :253-265   sqrt(dx^2+dy^2+dz^2) < R_drone + R_aircraft
#pragma once

#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define DCRP_HD __host__ __device__ __forceinline__
#else
#define DCRP_HD inline
#endif

// One rounding per source-level operation, never contracted into an FMA: the
// reference object code is scalar SSE2 (mulsd/addsd).  On the device the _rn
// intrinsics are immune to -fmad; the host test build uses -ffp-contract=off.
#if defined(__CUDA_ARCH__)
#define DADD(a, b) __dadd_rn((a), (b))
#define DSUB(a, b) __dsub_rn((a), (b))
#define DMUL(a, b) __dmul_rn((a), (b))
#define DDIV(a, b) __ddiv_rn((a), (b))
#define DSQRT(a)   __dsqrt_rn((a))
#else
#define DADD(a, b) ((a) + (b))
#define DSUB(a, b) ((a) - (b))
#define DMUL(a, b) ((a) * (b))
#define DDIV(a, b) ((a) / (b))
#define DSQRT(a)   sqrt((a))
#endif

#define DCRP_PI 3.14159265358979323846

namespace dcrp {

// ---------------------------------------------------------------- Philox4x32-10
struct U4 { uint32_t x, y, z, w; };

DCRP_HD uint32_t mulhi32(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

DCRP_HD U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                         uint32_t k0, uint32_t k1)
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = mulhi32(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = mulhi32(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    U4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

// Same function with the ten round keys (k0 + r*W0, k1 + r*W1) precomputed by the caller: in a kernel
// they sit in the parameter constant bank and feed the XORs directly, so the per-round key bump
// (two integer adds per round per trial) disappears.  rk[2r] = k0_r, rk[2r+1] = k1_r.
DCRP_HD U4 philox4x32_10_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const uint32_t (&rk)[20])
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = mulhi32(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = mulhi32(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ rk[2 * r];
        uint32_t n2 = hi0 ^ c3 ^ rk[2 * r + 1];
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    }
    U4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

inline void philox_round_keys(uint64_t seed, uint32_t (&rk)[20])
{
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) { rk[2 * r] = k0; rk[2 * r + 1] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
}

// counter = (trial_lo, trial_hi, drone, track), key = seed
DCRP_HD U4 trial_words(uint64_t seed, uint64_t trial, uint32_t drone, uint32_t track)
{
    return philox4x32_10((uint32_t)trial, (uint32_t)(trial >> 32), drone, track,
                         (uint32_t)seed, (uint32_t)(seed >> 32));
}

// word0 -> t1st in [1, T]   (Drone.cpp:108-116)
DCRP_HD int32_t draw_t1st(uint32_t w0, int32_t takeoff_t)
{
    return 1 + (int32_t)mulhi32(w0, (uint32_t)takeoff_t);
}

// word -> uniform real on [a, b): u*(b-a)+a with u = w/2^32 (Drone.cpp:175-187,
// libstdc++ uniform_real_distribution form); two roundings, no FMA.
DCRP_HD double draw_real(uint32_t w, double a, double b)
{
    double u = (double)w * (1.0 / 4294967296.0);   // exact
    return DADD(DMUL(u, DSUB(b, a)), a);
}

struct Draw {            // == dcrp_draw (32 bytes)
    int32_t t1st;
    int32_t reserved;
    double  tx, ty, tz;
};

DCRP_HD Draw make_draw(uint64_t seed, uint64_t trial, uint32_t drone, uint32_t track,
                       int32_t takeoff_t, const double* box)
{
    U4 w = trial_words(seed, trial, drone, track);
    Draw d;
    d.t1st = draw_t1st(w.x, takeoff_t);
    d.reserved = 0;
    d.tx = draw_real(w.y, box[0], box[1]);
    d.ty = draw_real(w.z, box[2], box[3]);
    d.tz = draw_real(w.w, box[4], box[5]);
    return d;
}

// ---------------------------------------------------------------- division by a launch constant
// n / d for a divisor fixed per launch (Granlund-Montgomery): 4 instructions instead of the ~20 of a 32-bit
// integer division.  q = (t + ((n - t) >> s1)) >> s2 with t = mulhi(m, n); exact for every 32-bit n and d >= 1
// (tests/test_hostsim_math.py).
struct FastDiv {
    uint32_t m, s1, s2;
    DCRP_HD uint32_t div(uint32_t n) const
    {
        const uint32_t t = mulhi32(m, n);
        return (t + ((n - t) >> s1)) >> s2;
    }
};
inline FastDiv make_fastdiv(uint32_t d)
{
    FastDiv f;
    uint32_t l = 0;
    while (l < 32 && (1ull << l) < d) ++l;
    f.m = (uint32_t)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
    f.s1 = l < 1 ? l : 1;
    f.s2 = l > 0 ? l - 1 : 0;
    return f;
}

// ---------------------------------------------------------------- stage 2 set-up
struct Ray64 {           // the trial's second-stage ray, fp64
    double x, y, z;      // position at index t1st-1 (the kink)
    double sx, sy, sz;   // per-index step for indices >= t1st
};

// Drone.cpp:193-220 followed by the loop-invariant products of :225-227.
// (xl, yl, zl) = drone position at last_index = t1st-1.
DCRP_HD Ray64 stage2_ray(double xl, double yl, double zl,
                         double tx, double ty, double tz,
                         double v_straight, double v_ascend)
{
    double long_diff = fabs(DSUB(tx, xl));
    double lat_diff  = fabs(DSUB(ty, yl));
    double alt_diff  = fabs(DSUB(tz, zl));
    double heading_angle = 0.0;
    if (tx <= xl && ty >= yl)                                    // TOP LEFT     :200
        heading_angle = DSUB(DMUL(2.0, DCRP_PI), atan(DDIV(long_diff, lat_diff)));
    else if (tx <= xl && ty <= yl)                               // BOTTOM LEFT  :204
        heading_angle = DSUB(DMUL(1.5, DCRP_PI), atan(DDIV(lat_diff, long_diff)));
    else if (tx >= xl && ty <= yl)                               // BOTTOM RIGHT :208
        heading_angle = DSUB(DCRP_PI, atan(DDIV(long_diff, lat_diff)));
    else if (tx >= xl && ty >= yl)                               // TOP RIGHT    :212
        heading_angle = atan(DDIV(long_diff, lat_diff));
    double modulus = DSQRT(DADD(DMUL(long_diff, long_diff), DMUL(lat_diff, lat_diff)));   // :216
    double pitch = atan(DDIV(alt_diff, modulus));                                          // :218
    // :220  (max_straight - (2*(max_straight - max_ascend)/M_PI)*pitch)
    double coef = DDIV(DMUL(2.0, DSUB(v_straight, v_ascend)), DCRP_PI);
    double vf = DSUB(v_straight, DMUL(coef, pitch));
    Ray64 r;
    r.x = xl; r.y = yl; r.z = zl;
    r.sx = DMUL(sin(heading_angle), vf);                         // :225
    r.sy = DMUL(cos(heading_angle), vf);                         // :226
    r.sz = DMUL(sin(pitch), vf);                                 // :227
    return r;
}

// squared separation in the reference's summation order (Drone.cpp:255-259)
DCRP_HD double sep2(double x, double y, double z, double ax, double ay, double az)
{
    double dx = DSUB(x, ax), dy = DSUB(y, ay), dz = DSUB(z, az);
    return DADD(DADD(DMUL(dx, dx), DMUL(dy, dy)), DMUL(dz, dz));
}

// ---------------------------------------------------------------- sub-step (interpolated) separation
// EXTENSION, no counterpart in the reference (which compares sample i with sample i only,
// Drone.cpp:254-259): drone and aircraft move linearly between consecutive samples; the relative
// position on segment [i, i+1] is D(t) = D0 + t*(D1 - D0), t in [0, 1].
struct SegOut {
    double d2;    // min over t of |D(t)|^2
    double tc;    // first t in [0, 1] with |D(t)|^2 < r2, or -1 when the segment stays outside
};

DCRP_HD SegOut seg_closest(double x0, double y0, double z0, double x1, double y1, double z1, double r2)
{
    const double ex = DSUB(x1, x0), ey = DSUB(y1, y0), ez = DSUB(z1, z0);
    const double a = DADD(DADD(DMUL(ex, ex), DMUL(ey, ey)), DMUL(ez, ez));
    const double b = DADD(DADD(DMUL(x0, ex), DMUL(y0, ey)), DMUL(z0, ez));
    const double c = DADD(DADD(DMUL(x0, x0), DMUL(y0, y0)), DMUL(z0, z0));
    double t = 0.0;
    if (a > 0.0) {
        t = DDIV(-b, a);
        if (t < 0.0) t = 0.0;
        if (t > 1.0) t = 1.0;
    }
    SegOut o;
    o.d2 = DADD(c, DMUL(t, DADD(DMUL(2.0, b), DMUL(t, a))));
    if (o.d2 < 0.0) o.d2 = 0.0;
    o.tc = -1.0;
    if (c < r2) {
        o.tc = 0.0;
    } else if (o.d2 < r2) {              // enters the sphere inside the segment: smaller root of |D(t)|^2 = r2
        const double disc = DSUB(DMUL(b, b), DMUL(a, DSUB(c, r2)));
        double tc = DDIV(DSUB(-b, DSQRT(disc > 0.0 ? disc : 0.0)), a);
        if (tc < 0.0) tc = 0.0;
        if (tc > 1.0) tc = 1.0;
        o.tc = tc;
    }
    return o;
}

// ---------------------------------------------------------------- fp32 screen set-up
struct Ray32 {           // second-stage ray relative to the scenario origin, fp32:
    float bx, by, bz;    // position(i) = b + i * s  for absolute index i >= t1st
    float sx, sy, sz;
    bool  degenerate;    // target (numerically) on top of the kink: must be confirmed
};

// atan(x) for x >= 0 in fp32 (the pitch of the screen's ray, Drone.cpp:218): degree-7 minimax polynomial of
// atan(sqrt(u))/sqrt(u) on [0, 1] (fit and measured in fp32 by tests/test_hostsim_math.py: |error| <= 2e-7 rad,
// the same as the 2 ulp of the library atanf at pi/4) and atan(x) = pi/2 - atan(1/x) above 1.  14 instructions
// instead of the library's 25; the error is inside the screen margin (engine.cu: n_steps * vmax * 2^-21 term).
DCRP_HD float atan_nonneg(float x)
{
    const bool big = x > 1.0f;
#if defined(__CUDA_ARCH__)
    float rc;                                   // one MUFU.RCP (2 ulp); __fdividef's range fix-ups cost 12 instructions
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(x));
    const float r = big ? rc : x;
#else
    const float r = big ? 1.0f / x : x;
#endif
    const float u = r * r;
    float p = -0.004668773151934147f;
    p = fmaf(p, u, 0.02416618913412094f);
    p = fmaf(p, u, -0.0593671016395092f);
    p = fmaf(p, u, 0.09906096756458282f);
    p = fmaf(p, u, -0.14016585052013397f);
    p = fmaf(p, u, 0.19969235360622406f);
    p = fmaf(p, u, -0.33331960439682007f);
    p = fmaf(p, u, 0.9999998807907104f);
    p *= r;
    return big ? 1.57079632679489662f - p : p;
}

// (dls, dbs, da) = target - kink, formed in fp64 and rounded once to fp32;
// (xlf, ylf, zlf) = kink - origin rounded once to fp32; lastf = float(t1st-1).
// Direction without trig: sin(atan(dL/dB)) = dL/m etc.; pitch needs one atanf.
DCRP_HD Ray32 stage2_ray32(float dls, float dbs, float da,
                           float xlf, float ylf, float zlf, float lastf,
                           float v_straight, float coef)
{
    Ray32 r;
    float m2 = fmaf(dls, dls, dbs * dbs);
    r.degenerate = !(m2 > 1e-12f);
#if defined(__CUDA_ARCH__)
    float inv_m = rsqrtf(m2);
    inv_m = inv_m * fmaf(-0.5f * m2 * inv_m, inv_m, 1.5f);        // one Newton step
    float n2 = fmaf(da, da, m2);
    float inv_n = rsqrtf(n2);
    inv_n = inv_n * fmaf(-0.5f * n2 * inv_n, inv_n, 1.5f);
#else
    float inv_m = 1.0f / sqrtf(m2);
    float n2 = fmaf(da, da, m2);
    float inv_n = 1.0f / sqrtf(n2);
#endif
    float ada = fabsf(da);
    float pitch = atan_nonneg(ada * inv_m);
    float vf = fmaf(-coef, pitch, v_straight);
    r.sx = dls * inv_m * vf;
    r.sy = dbs * inv_m * vf;
    r.sz = ada * inv_n * vf;
    r.bx = fmaf(-lastf, r.sx, xlf);
    r.by = fmaf(-lastf, r.sy, ylf);
    r.bz = fmaf(-lastf, r.sz, zlf);
    return r;
}

// ---------------------------------------------------------------- 1-D slab level (screen3.cu)
// The stage-2 ray projected on a unit vector n, relative to the scenario origin: n . position(i) = b + i * s for the
// absolute index i.  (dls, dbs) = target - kink and ada = |target altitude - start altitude|, each formed in fp64
// and rounded once; pln = n . (kink - origin); lastf = float(t1st - 1).  Same speed law as stage2_ray32
// (Drone.cpp:216-220), one MUFU.RSQ on the device.
struct Slab32 {
    float b, s;
    bool  degenerate;    // target (numerically) on top of the kink: the trial must be confirmed
};

DCRP_HD Slab32 slab_ray32(float dls, float dbs, float ada, float pln, float lastf, float nx, float ny,
                          float v_straight, float coef)
{
    Slab32 o;
    const float m2 = fmaf(dls, dls, dbs * dbs);
    o.degenerate = !(m2 > 1e-12f);
#if defined(__CUDA_ARCH__)
    float inv_m;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(inv_m) : "f"(m2));
#else
    const float inv_m = 1.0f / sqrtf(m2);
#endif
    const float pitch = atan_nonneg(ada * inv_m);                 // Drone.cpp:218
    const float k = inv_m * fmaf(-coef, pitch, v_straight);       // vf / m, Drone.cpp:220
    o.s = fmaf(dls, nx, dbs * ny) * k;                            // n . step
    o.b = fmaf(-lastf, o.s, pln);                                 // n . position(0)
    return o;
}

}  // namespace dcrp
