// Batch.cpp -- see Batch.h.
#include "Batch.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <memory>
#include <thread>
#include <vector>

#include "Aircraft.h"
#include "Drone.h"

namespace dcrp_host {

namespace {
// small files as the only transport between ranks: written under a temporary name, then renamed (atomic)
bool publish_file(const std::string& path, const void* data, size_t bytes)
{
    const std::string tmp = path + ".tmp";
    FILE* f = std::fopen(tmp.c_str(), "wb");
    if (!f) return false;
    const bool ok = bytes == 0 || std::fwrite(data, 1, bytes, f) == bytes;
    std::fclose(f);
    return ok && std::rename(tmp.c_str(), path.c_str()) == 0;
}
bool await_file(const std::string& path, std::vector<unsigned char>& out, double timeout_s)
{
    const auto t0 = std::chrono::steady_clock::now();
    for (;;) {
        if (FILE* f = std::fopen(path.c_str(), "rb")) {
            out.clear();
            unsigned char buf[4096];
            size_t n;
            while ((n = std::fread(buf, 1, sizeof buf, f)) > 0) out.insert(out.end(), buf, buf + n);
            std::fclose(f);
            return true;
        }
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > timeout_s) return false;
        std::this_thread::sleep_for(std::chrono::milliseconds(20));
    }
}
}  // namespace

BatchResult monte_carlo_batch(const std::string& aircraft_csv, int aircraft_cols, int aircraft_begin,
                              int aircraft_end, const std::string& drone_csv, int drone_cols, int drone_total,
                              uint64_t number_runs, double aircraft_radius, double drone_radius, uint64_t seed,
                              int precision, const std::string& out_dir, uint32_t run_flags, int n_gpus,
                              const RankSpec& ranks)
{
    BatchResult out;
    auto fail = [&](int status, const std::string& what) { out.status = status; out.error = what; return out; };
    if (aircraft_end <= aircraft_begin || drone_total <= 0) return fail(DCRP_ERR_INVALID, "empty aircraft or drone range");

    // ---- inputs through the same loaders the per-pair path uses
    Aircraft aircraft;
    aircraft.Set_Parameters_and_Data(aircraft_csv, aircraft_cols);
    if (!aircraft.ok()) return fail(DCRP_ERR_IO, aircraft.last_error());
    const int n_tracks = aircraft_end - aircraft_begin;
    std::vector<std::vector<double>> xs(n_tracks), ys(n_tracks), zs(n_tracks);
    std::vector<dcrp_track> tracks(n_tracks);
    for (int k = 0; k < n_tracks; ++k) {
        aircraft.Vector_Allocation(aircraft_begin + k);
        if (aircraft.Vector_length <= 0) return fail(DCRP_ERR_INVALID, "no flight " + std::to_string(aircraft_begin + k));
        xs[k].assign(aircraft.longitude_vector, aircraft.longitude_vector + aircraft.Vector_length);
        ys[k].assign(aircraft.latitude_vector, aircraft.latitude_vector + aircraft.Vector_length);
        zs[k].assign(aircraft.altitude_vector, aircraft.altitude_vector + aircraft.Vector_length);
        tracks[k].x = xs[k].data(); tracks[k].y = ys[k].data(); tracks[k].z = zs[k].data();
        tracks[k].n_steps = aircraft.Vector_length;
        tracks[k].takeoff_t = aircraft.takeoff_t;
        const int rc = dcrp_target_box(ys[k][0], zs[k][0], tracks[k].box);
        aircraft.Deallocation();
        if (rc != DCRP_OK) return fail(rc, "flight " + std::to_string(aircraft_begin + k) + ": " + dcrp_last_error());
    }
    std::vector<dcrp_drone> drones(drone_total);
    {
        Drone probe;
        double one = 0.0, row[3];
        for (int j = 0; j < drone_total; ++j) {
            probe.SetInitialParameters(drone_csv, 1, drone_cols, j, 0, 1, &one, &one, &one, aircraft_radius, drone_radius);
            if (probe.LastStatus() != 0) return fail(probe.LastStatus(), probe.LastError());
            probe.InitialRow(row);
            drones[j] = dcrp_drone{row[0], row[1], row[2]};
        }
    }

    // ---- one launch for everything
    const bool multi = ranks.n_ranks > 1;
    if (multi && (ranks.rank < 0 || ranks.rank >= ranks.n_ranks || ranks.rendezvous.empty()))
        return fail(DCRP_ERR_INVALID, "multi-process run: rank must be in [0, n_ranks) and a rendezvous path given");
    if (multi && !std::getenv("DCRP_DEVICE")) {
        int visible = 1;
        if (dcrp_device_count(&visible) != DCRP_OK) return fail(DCRP_ERR_NO_DEVICE, dcrp_last_error());
        set_shared_device(ranks.rank % visible);
    }
    std::string err;
    dcrp_engine* eng = shared_engine(&err);
    if (!eng) return fail(DCRP_ERR_NO_DEVICE, err);
    dcrp_model model;
    dcrp_default_model(&model);
    model.aircraft_radius = aircraft_radius;
    model.drone_radius = drone_radius;
    dcrp_run run{};
    run.seed = seed; run.trial_begin = 0; run.trial_count = number_runs; run.precision = precision;
    run.flags = run_flags;
    run.record_capacity = 1u << 20;
    run.track_id_base = (uint32_t)aircraft_begin;       // same Philox ids as Drone::Simulation(i, j)
    std::vector<uint64_t> counts((size_t)n_tracks * drone_total, 0);
    std::vector<dcrp_record> records(run.record_capacity);
    dcrp_result res{};
    res.counts = counts.data();
    res.records = records.data();
    int rc = DCRP_OK;
    int visible = 1;
    if (n_gpus != 1 && dcrp_device_count(&visible) != DCRP_OK) return fail(DCRP_ERR_NO_DEVICE, dcrp_last_error());
    const int G = multi ? 1 : (n_gpus <= 0 ? visible : std::min(n_gpus, visible));
    if (multi) {
        // ---- one process per GPU: communicator id through the rendezvous file, trial range of this rank,
        // counts all-reduced on the device, records to rank 0 through files
        unsigned char uid[DCRP_UNIQUE_ID_BYTES];
        if (ranks.rank == 0) {
            if ((rc = dcrp_comm_unique_id(uid)) != DCRP_OK) return fail(rc, dcrp_last_error());
            if (!publish_file(ranks.rendezvous, uid, sizeof uid)) return fail(DCRP_ERR_IO, "cannot write " + ranks.rendezvous);
        } else {
            std::vector<unsigned char> got;
            if (!await_file(ranks.rendezvous, got, 120.0) || got.size() != sizeof uid)
                return fail(DCRP_ERR_COMM, "no communicator id at " + ranks.rendezvous + " after 120 s");
            std::copy(got.begin(), got.end(), uid);
        }
        if ((rc = dcrp_engine_attach_comm(eng, uid, ranks.rank, ranks.n_ranks, 0)) != DCRP_OK) return fail(rc, dcrp_last_error());
        if ((rc = dcrp_scenario_upload(eng, tracks.data(), n_tracks, drones.data(), drone_total, &model)) != DCRP_OK)
            return fail(rc, dcrp_last_error());
        dcrp_run part = run;
        part.trial_begin = number_runs * (uint64_t)ranks.rank / (uint64_t)ranks.n_ranks;
        part.trial_count = number_runs * (uint64_t)(ranks.rank + 1) / (uint64_t)ranks.n_ranks - part.trial_begin;
        if ((rc = dcrp_run_async(eng, &part, nullptr, nullptr)) != DCRP_OK) return fail(rc, dcrp_last_error());
        if ((rc = dcrp_allreduce_counts(eng, nullptr, nullptr)) != DCRP_OK) return fail(rc, dcrp_last_error());
        if ((rc = dcrp_fetch(eng, &res)) != DCRP_OK) return fail(rc, dcrp_last_error());      // counts: global; records: this rank's
        const std::string rec_base = ranks.rendezvous + ".records.";
        if (ranks.rank != 0) {
            if (!publish_file(rec_base + std::to_string(ranks.rank), records.data(), sizeof(dcrp_record) * (size_t)res.n_records))
                return fail(DCRP_ERR_IO, "cannot write the records of rank " + std::to_string(ranks.rank));
        } else {
            uint32_t n_rec = res.n_records;
            for (int r = 1; r < ranks.n_ranks; ++r) {
                std::vector<unsigned char> got;
                if (!await_file(rec_base + std::to_string(r), got, 600.0) || got.size() % sizeof(dcrp_record) != 0)
                    return fail(DCRP_ERR_COMM, "no records from rank " + std::to_string(r));
                const size_t k = got.size() / sizeof(dcrp_record);
                if ((size_t)n_rec + k > records.size()) records.resize((size_t)n_rec + k);
                std::copy(got.begin(), got.end(), reinterpret_cast<unsigned char*>(records.data() + n_rec));
                n_rec += (uint32_t)k;
                std::remove((rec_base + std::to_string(r)).c_str());
            }
            std::sort(records.begin(), records.begin() + n_rec, [](const dcrp_record& a, const dcrp_record& b) {
                if (a.track != b.track) return a.track < b.track;
                if (a.drone != b.drone) return a.drone < b.drone;
                return a.trial < b.trial;
            });
            res.n_records = n_rec;
            std::remove(ranks.rendezvous.c_str());
        }
    } else if (G <= 1) {
        rc = dcrp_simulate(eng, tracks.data(), n_tracks, drones.data(), drone_total, &model, &run, &res);
        if (rc != DCRP_OK) return fail(rc, dcrp_last_error());
    } else {
        // One process, G devices: the same scenario on every device, device g runs the contiguous trial range
        // [g n / G, (g+1) n / G) of every (track, drone) -- the Philox counter carries the global trial id, so the
        // union is the single-GPU result bit for bit -- all launches queued before the first fetch; counts are
        // summed and records merged here (one process needs no collective).
        std::vector<dcrp_engine*> engs(G, nullptr);
        engs[0] = eng;
        auto cleanup = [&]() { for (int g = 1; g < G; ++g) if (engs[g]) dcrp_engine_destroy(engs[g]); };
        for (int g = 1; g < G; ++g)
            if ((rc = dcrp_engine_create(g, &engs[g])) != DCRP_OK) { cleanup(); return fail(rc, dcrp_last_error()); }
        for (int g = 0; g < G && rc == DCRP_OK; ++g) {
            rc = dcrp_scenario_upload(engs[g], tracks.data(), n_tracks, drones.data(), drone_total, &model);
            if (rc != DCRP_OK) break;
            dcrp_run part = run;
            part.trial_begin = number_runs * (uint64_t)g / (uint64_t)G;
            part.trial_count = number_runs * (uint64_t)(g + 1) / (uint64_t)G - part.trial_begin;
            rc = dcrp_run_async(engs[g], &part, nullptr, nullptr);
        }
        if (rc != DCRP_OK) { const std::string why = dcrp_last_error(); cleanup(); return fail(rc, why); }
        uint32_t n_rec = 0;
        int overflow = 0;
        std::vector<uint64_t> part_counts(counts.size());
        dcrp_stats total{};
        for (int g = 0; g < G; ++g) {
            std::fill(part_counts.begin(), part_counts.end(), 0);
            dcrp_result pr{};
            pr.counts = part_counts.data();
            pr.records = records.data() + n_rec;
            // dcrp_fetch fills at most the run's record_capacity; keep the merged buffer within bounds
            rc = dcrp_fetch(engs[g], &pr);
            if (rc != DCRP_OK) { const std::string why = dcrp_last_error(); cleanup(); return fail(rc, why); }
            for (size_t k = 0; k < counts.size(); ++k) counts[k] += part_counts[k];
            overflow |= pr.overflow;
            n_rec += pr.n_records;
            if ((size_t)n_rec + run.record_capacity > records.size()) records.resize((size_t)n_rec + run.record_capacity);
            total.trials += pr.stats.trials; total.tests_evaluated += pr.stats.tests_evaluated;
            total.tests_shared += pr.stats.tests_shared; total.tests_issued += pr.stats.tests_issued;
            total.tests_culled += pr.stats.tests_culled; total.trials_culled += pr.stats.trials_culled;
            total.candidates += pr.stats.candidates; total.level2 += pr.stats.level2; total.hits += pr.stats.hits;
            total.kernel_launches += pr.stats.kernel_launches;
            total.ms_trials = std::max(total.ms_trials, pr.stats.ms_trials);
            total.ms_device_total = std::max(total.ms_device_total, pr.stats.ms_device_total);
        }
        cleanup();
        std::sort(records.begin(), records.begin() + n_rec, [](const dcrp_record& a, const dcrp_record& b) {
            if (a.track != b.track) return a.track < b.track;
            if (a.drone != b.drone) return a.drone < b.drone;
            return a.trial < b.trial;
        });
        res.n_records = n_rec;
        res.overflow = overflow;
        res.stats = total;
    }
    for (uint64_t c : counts) out.collisions += c;
    out.trials = (uint64_t)n_tracks * drone_total * number_runs;
    out.stats = res.stats;
    if (res.overflow) return fail(DCRP_ERR_STATE, "more collisions than the record buffer holds; lower number_runs per call");
    if (multi && ranks.rank != 0) return out;          // rank 0 writes the files

    // ---- collision blocks, one file per aircraft (records arrive sorted by track, drone, trial)
    int stride = 0;
    std::vector<double> xyz;
    if (res.n_records) {
        rc = dcrp_trajectories(eng, nullptr, 0, nullptr, &stride);
        if (rc != DCRP_OK) return fail(rc, dcrp_last_error());
        xyz.assign((size_t)res.n_records * 3 * (size_t)stride, 0.0);
        rc = dcrp_trajectories(eng, records.data(), res.n_records, xyz.data(), &stride);
        if (rc != DCRP_OK) return fail(rc, dcrp_last_error());
    }
    uint32_t k = 0;
    for (int a = 0; a < n_tracks; ++a) {
        std::ofstream f(out_dir + "/Drone_coords_" + std::to_string(aircraft_begin + a) + ".csv",
                        std::ofstream::out | std::ofstream::trunc);          // ClearOutput, Drone.cpp:59-64
        for (; k < res.n_records && (int)records[k].track == a; ++k) {
            const double* x = xyz.data() + (size_t)k * 3 * (size_t)stride;
            if (f) write_block(f, x, x + stride, x + 2 * (size_t)stride, tracks[a].n_steps, (long)records[k].trial);
        }
    }
    return out;
}

}  // namespace dcrp_host
