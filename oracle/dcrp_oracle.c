/*
 * dcrp_oracle.c -- plain-C CPU restatement of the reference trial path.
 * TEST INFRASTRUCTURE ONLY (see dcrp_oracle.h).  Build: oracle/Makefile,
 * gcc -O2 -ffp-contract=off (the reference Makefile:3 uses -O3 without -march,
 * so its object has no FMA; we forbid contraction explicitly).
 *
 * All arithmetic is IEEE double with one rounding per source-level operation,
 * in the reference's order.
 */
#include "dcrp_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

void orc_default_model(orc_model* m)
{
    m->aircraft_radius = 3.5;   /* Monte_Carlo.cpp:13 */
    m->drone_radius    = 0.19;  /* Monte_Carlo.cpp:18 */
    m->start_alt       = 100.0; /* Drone.cpp:54 */
    m->v_straight      = 19.0;  /* Drone.cpp:55 */
    m->v_ascend        = 8.0;   /* Drone.cpp:56 */
}

/* ------------------------------------------------------------------ Philox */
static inline void mulhilo32(uint32_t a, uint32_t b, uint32_t* hi, uint32_t* lo)
{
    uint64_t p = (uint64_t)a * (uint64_t)b;
    *hi = (uint32_t)(p >> 32);
    *lo = (uint32_t)p;
}

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo32(M0, c0, &hi0, &lo0);
        mulhilo32(M1, c2, &hi1, &lo1);
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n1 = lo1;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        uint32_t n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void orc_make_draw(uint64_t seed, uint64_t trial, uint32_t drone, uint32_t track,
                   int32_t takeoff_t, const double box[6], orc_draw* out)
{
    uint32_t ctr[4] = { (uint32_t)trial, (uint32_t)(trial >> 32), drone, track };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t w[4];
    orc_philox4x32_10(ctr, key, w);
    /* uniform int on [1, T] (Drone.cpp:108-116): multiply-shift map of one word */
    out->t1st = 1 + (int32_t)(((uint64_t)w[0] * (uint64_t)(uint32_t)takeoff_t) >> 32);
    out->reserved = 0;
    /* uniform real on [a, b) (Drone.cpp:175-187): u*(b-a)+a, u = w/2^32 exact */
    const double s = 1.0 / 4294967296.0;
    double u1 = (double)w[1] * s, u2 = (double)w[2] * s, u3 = (double)w[3] * s;
    out->tx = u1 * (box[1] - box[0]) + box[0];
    out->ty = u2 * (box[3] - box[2]) + box[2];
    out->tz = u3 * (box[5] - box[4]) + box[4];
}

/* n consecutive trials' draws (the replay table of one (track, drone) pair) */
void orc_make_draws(uint64_t seed, uint64_t trial_begin, int64_t n, uint32_t drone, uint32_t track,
                    int32_t takeoff_t, const double box[6], orc_draw* out)
{
    for (int64_t r = 0; r < n; ++r) orc_make_draw(seed, trial_begin + (uint64_t)r, drone, track, takeoff_t, box, out + r);
}

/* --------------------------------------------------------------- target box */
int orc_target_box(double ay0, double az0, double box[6])
{
    int depart_or_arrive = (az0 != 0) ? 1 : 0;       /* Drone.cpp:47-52 */
    double min_alt = 100;                            /* Drone.cpp:131 */
    double max_alt = min_alt + 400;                  /* Drone.cpp:132 */
    double min_lat, max_lat, min_long, max_long;
    if (ay0 <= 5706340.62 && ay0 >= 5705792.58) {    /* 27R, Drone.cpp:135 */
        min_lat = 5705770.46;
        max_lat = min_lat + 500;
    } else if (ay0 <= 5705065.74 && ay0 >= 5704388.38) { /* 27L, Drone.cpp:148 */
        max_lat = 5704800.46;
        min_lat = max_lat - 500;
    } else {
        return -1;
    }
    if (depart_or_arrive) {                          /* Drone.cpp:138-141,152-155 */
        min_long = 676345.60;
        max_long = min_long + 5000;
    } else {                                         /* Drone.cpp:142-145,156-159 */
        max_long = 676819.02;
        min_long = max_long - 5000;
    }
    box[0] = min_long; box[1] = max_long;
    box[2] = min_lat;  box[3] = max_lat;
    box[4] = min_alt;  box[5] = max_alt;
    return 0;
}

/* -------------------------------------------------------------------- trial */
void orc_trial(const double* ax, const double* ay, const double* az, int32_t n_steps,
               double x0, double y0, double heading0,
               const orc_draw* draw, const orc_model* model,
               double* traj_x, double* traj_y, double* traj_z,
               orc_trial_result* out)
{
    const int VL = n_steps;
    double* X = traj_x ? traj_x : (double*)malloc(sizeof(double) * (size_t)VL);
    double* Y = traj_y ? traj_y : (double*)malloc(sizeof(double) * (size_t)VL);
    double* Z = traj_z ? traj_z : (double*)malloc(sizeof(double) * (size_t)VL);
    double* H = (double*)malloc(sizeof(double) * (size_t)VL);

    /* SetInitialConditions, Drone.cpp:66-80 */
    X[0] = x0; Y[0] = y0; H[0] = heading0; Z[0] = model->start_alt;

    /* FirstStage, Drone.cpp:107-127 (draw supplied) */
    const int t1st = draw->t1st;
    for (int i = 1; i < t1st && i < VL; ++i) {
        X[i] = X[i - 1] + sin(H[i - 1]) * model->v_straight;
        Y[i] = Y[i - 1] + cos(H[i - 1]) * model->v_straight;
        H[i] = H[i - 1];
        Z[i] = model->start_alt;
    }

    /* SecondStage, Drone.cpp:164-231 */
    const double random_long = draw->tx, random_lat = draw->ty, random_alt = draw->tz;
    const int last = t1st - 1;
    double long_diff = fabs(random_long - X[last]);
    double lat_diff  = fabs(random_lat  - Y[last]);
    double alt_diff  = fabs(random_alt  - Z[last]);
    double heading_angle = 0.0;
    if (random_long <= X[last] && random_lat >= Y[last])        /* TOP LEFT  :200 */
        heading_angle = 2 * M_PI - atan(long_diff / lat_diff);
    else if (random_long <= X[last] && random_lat <= Y[last])   /* BOTTOM LEFT :204 */
        heading_angle = 1.5 * M_PI - atan(lat_diff / long_diff);
    else if (random_long >= X[last] && random_lat <= Y[last])   /* BOTTOM RIGHT :208 */
        heading_angle = M_PI - atan(long_diff / lat_diff);
    else if (random_long >= X[last] && random_lat >= Y[last])   /* TOP RIGHT :212 */
        heading_angle = atan(long_diff / lat_diff);

    double modulus_long_lat = sqrt(long_diff * long_diff + lat_diff * lat_diff);   /* :216 */
    double pitch_angle = atan(alt_diff / modulus_long_lat);                         /* :218 */
    double velocity_factor = (model->v_straight -
        (2 * (model->v_straight - model->v_ascend) / M_PI) * pitch_angle);         /* :220 */

    for (int i = t1st; i < VL; ++i) {                                               /* :224-230 */
        X[i] = X[i - 1] + sin(heading_angle) * velocity_factor;
        Y[i] = Y[i - 1] + cos(heading_angle) * velocity_factor;
        Z[i] = Z[i - 1] + sin(pitch_angle) * velocity_factor;
        H[i] = heading_angle;
    }

    /* Collision, Drone.cpp:253-265 -- first-hit index and min distance are
     * extensions; `hit` is exactly the reference's bool. */
    const double R = model->drone_radius + model->aircraft_radius;
    int hit = 0, first = -1;
    double mind = INFINITY;
    for (int i = 0; i < VL; ++i) {
        double distance = sqrt(
            (X[i] - ax[i]) * (X[i] - ax[i]) +
            (Y[i] - ay[i]) * (Y[i] - ay[i]) +
            (Z[i] - az[i]) * (Z[i] - az[i]));
        if (distance < mind) mind = distance;
        if (distance < R && !hit) { hit = 1; first = i; }
    }
    out->hit = hit; out->first_hit = first; out->min_dist = mind;

    if (!traj_x) free(X);
    if (!traj_y) free(Y);
    if (!traj_z) free(Z);
    free(H);
}

int64_t orc_run_trials(const double* ax, const double* ay, const double* az, int32_t n_steps,
                       double x0, double y0, double heading0,
                       const orc_draw* draws, int64_t n, const orc_model* model,
                       uint8_t* hit, int32_t* first_hit, double* min_dist)
{
    int64_t hits = 0;
    double* X = (double*)malloc(sizeof(double) * (size_t)n_steps * 3);
    for (int64_t r = 0; r < n; ++r) {                      /* Drone.cpp:270 */
        orc_trial_result res;
        orc_trial(ax, ay, az, n_steps, x0, y0, heading0, &draws[r], model,
                  X, X + n_steps, X + 2 * (size_t)n_steps, &res);
        hits += res.hit;                                   /* Drone.cpp:276 */
        if (hit) hit[r] = (uint8_t)res.hit;
        if (first_hit) first_hit[r] = res.first_hit;
        if (min_dist) min_dist[r] = res.min_dist;
    }
    free(X);
    return hits;
}

int64_t orc_run_philox(const double* ax, const double* ay, const double* az, int32_t n_steps,
                       int32_t takeoff_t, const double box[6],
                       double x0, double y0, double heading0,
                       uint64_t seed, uint32_t drone, uint32_t track,
                       uint64_t trial_begin, int64_t n, const orc_model* model,
                       uint64_t* hit_ids, int64_t cap, int64_t* n_ids)
{
    int64_t hits = 0, nid = 0;
    double* X = (double*)malloc(sizeof(double) * (size_t)n_steps * 3);
    for (int64_t r = 0; r < n; ++r) {
        orc_draw d;
        orc_trial_result res;
        orc_make_draw(seed, trial_begin + (uint64_t)r, drone, track, takeoff_t, box, &d);
        orc_trial(ax, ay, az, n_steps, x0, y0, heading0, &d, model,
                  X, X + n_steps, X + 2 * (size_t)n_steps, &res);
        if (res.hit) {
            ++hits;
            if (hit_ids && nid < cap) hit_ids[nid++] = trial_begin + (uint64_t)r;
        }
    }
    if (n_ids) *n_ids = nid;
    free(X);
    return hits;
}

/* ------------------------------------------------------- sub-step extension */
static void seg_closest(double x0, double y0, double z0, double x1, double y1, double z1, double r2,
                        double* d2_out, double* tc_out)
{
    double ex = x1 - x0, ey = y1 - y0, ez = z1 - z0;
    double a = (ex * ex + ey * ey) + ez * ez;
    double b = (x0 * ex + y0 * ey) + z0 * ez;
    double c = (x0 * x0 + y0 * y0) + z0 * z0;
    double t = 0.0;
    if (a > 0.0) {
        t = -b / a;
        if (t < 0.0) t = 0.0;
        if (t > 1.0) t = 1.0;
    }
    double d2 = c + t * (2.0 * b + t * a);
    if (d2 < 0.0) d2 = 0.0;
    double tc = -1.0;
    if (c < r2) {
        tc = 0.0;
    } else if (d2 < r2) {
        double disc = b * b - a * (c - r2);
        tc = (-b - sqrt(disc > 0.0 ? disc : 0.0)) / a;
        if (tc < 0.0) tc = 0.0;
        if (tc > 1.0) tc = 1.0;
    }
    *d2_out = d2;
    *tc_out = tc;
}

int64_t orc_run_trials_substep(const double* ax, const double* ay, const double* az, int32_t n_steps,
                               double x0, double y0, double heading0,
                               const orc_draw* draws, int64_t n, const orc_model* model,
                               uint8_t* hit, int32_t* first_seg, double* first_time, double* min_dist)
{
    int64_t hits = 0;
    const double R = model->drone_radius + model->aircraft_radius;
    const double r2 = R * R;
    double* X = (double*)malloc(sizeof(double) * (size_t)n_steps * 3);
    double* Y = X + n_steps;
    double* Z = X + 2 * (size_t)n_steps;
    for (int64_t r = 0; r < n; ++r) {
        orc_trial_result res;
        orc_trial(ax, ay, az, n_steps, x0, y0, heading0, &draws[r], model, X, Y, Z, &res);
        int first = -1;
        double ftime = -1.0, mn = INFINITY;
        for (int i = 0; i + 1 < n_steps; ++i) {
            double d2, tc;
            seg_closest(X[i] - ax[i], Y[i] - ay[i], Z[i] - az[i],
                        X[i + 1] - ax[i + 1], Y[i + 1] - ay[i + 1], Z[i + 1] - az[i + 1], r2, &d2, &tc);
            if (d2 < mn) mn = d2;
            if (tc >= 0.0 && first < 0) { first = i; ftime = (double)i + tc; }
        }
        if (first >= 0) ++hits;
        if (hit) hit[r] = first >= 0;
        if (first_seg) first_seg[r] = first;
        if (first_time) first_time[r] = ftime;
        if (min_dist) min_dist[r] = sqrt(mn);
    }
    free(X);
    return hits;
}

int64_t orc_format_block(const double* x, const double* y, const double* z, int32_t n_steps,
                         int32_t run_no, char* buf, int64_t cap)
{
    int64_t pos = 0;
    for (int i = 0; i < n_steps; ++i) {                    /* Drone.cpp:238-240 */
        int w = snprintf(buf + pos, (size_t)(cap - pos), "%.10g,%.10g,%.10g,%d\n",
                         x[i], y[i], z[i], run_no);
        if (w < 0 || pos + w >= cap) return -1;
        pos += w;
    }
    if (pos + 2 > cap) return -1;
    buf[pos++] = '\n';                                     /* Drone.cpp:241 */
    buf[pos] = 0;
    return pos;
}
