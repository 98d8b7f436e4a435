/*
 * replay_shim.h -- force-included (-include) in front of the UNMODIFIED
 * /root/reference/Drone.cpp so that its four per-trial random draws
 * (Drone.cpp:113-116 and :170-187) come from a replay table instead of
 * std::random_device.  TEST INFRASTRUCTURE ONLY; the reference source is
 * compiled where it lies, never copied.
 *
 * Mechanism: pull in every standard header Drone.cpp uses FIRST (so the macros
 * below cannot disturb libstdc++), then rename the four <random> identifiers
 * the reference spells out to replay types.
 */
#ifndef DCRP_REPLAY_SHIM_H
#define DCRP_REPLAY_SHIM_H

#include <random>
#include <iostream>
#include <iomanip>
#include <fstream>
#include <sstream>
#include <vector>
#include <string>
#include <cmath>
#include <cstdint>
#include <cstdlib>

namespace dcrp_replay {

struct Draw {            /* == orc_draw == dcrp_draw (32 bytes) */
    int32_t t1st;
    int32_t reserved;
    double  tx, ty, tz;
};

/* g_next: the table entry the NEXT trial will consume (advanced by the int
 * draw, which is the first draw of every trial, Drone.cpp:116); g_cur: the
 * entry of the trial in flight; g_real_pos: which of the three reals is next */
extern thread_local const Draw* g_next;
extern thread_local const Draw* g_cur;
extern thread_local int         g_real_pos;
extern thread_local long        g_range_violations;

struct device {
    unsigned int operator()() { return 0u; }
};

struct engine {
    explicit engine(unsigned int) {}
};

template <class T>
struct int_dist {
    T a, b;
    int_dist(T lo, T hi) : a(lo), b(hi) {}
    T operator()(engine&) {
        g_cur = g_next++;
        T v = static_cast<T>(g_cur->t1st);
        if (v < a || v > b) ++g_range_violations;
        g_real_pos = 0;
        return v;
    }
};

template <class T>
struct real_dist {
    T a, b;
    real_dist(T lo, T hi) : a(lo), b(hi) {}
    /* Drone.cpp:183,185,187 draw long, lat, alt in that order */
    T operator()(engine&) {
        T v = (g_real_pos == 0) ? g_cur->tx : (g_real_pos == 1) ? g_cur->ty : g_cur->tz;
        if (v < a || v > b) ++g_range_violations;
        g_real_pos = (g_real_pos + 1) % 3;
        return v;
    }
};

}  // namespace dcrp_replay

#define random_device             dcrp_replay::device
#define default_random_engine     dcrp_replay::engine
#define uniform_int_distribution  dcrp_replay::int_dist
#define uniform_real_distribution dcrp_replay::real_dist

#endif
