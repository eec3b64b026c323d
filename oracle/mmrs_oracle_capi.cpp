// =============================================================================
// mmrs_oracle_capi.cpp — C entry points of the CPU ORACLE (test infrastructure).
// See mmrs_oracle.hpp for scope, parity status and reference citations.
// Loaded with ctypes by tests/ (checker), __graft_entry__.smoke() (checker) and
// bench.py (cpu_baseline / --impl reference legs). Never by the product.
// =============================================================================
#include "mmrs_oracle.hpp"

using namespace ora;

static thread_local std::string g_err;

template <class F>
static int guard(F&& f) {
    try {
        f();
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

static std::vector<ContourPoint> pts_from_xy(const double* xy, long n) {
    std::vector<ContourPoint> v((size_t)std::max<long>(n, 0));
    for (long i = 0; i < n; ++i) {
        v[i].x = xy[2 * i];
        v[i].y = xy[2 * i + 1];
        v[i].point_index = (uint32_t)i;
    }
    return v;
}
static double* dup_vec(const std::vector<double>& v) {
    double* p = (double*)std::malloc(std::max<size_t>(v.size(), 1) * sizeof(double));
    std::memcpy(p, v.data(), v.size() * sizeof(double));
    return p;
}
static std::vector<double> encode_logs(const std::vector<AlignLog>& logs) {
    std::vector<double> o;
    for (auto& l : logs) o.insert(o.end(), {(double)l.contour_id, (double)l.matched_to, l.rot_deg, l.tx, l.ty, l.cx, l.cy});
    return o;
}

extern "C" {

const char* ora_last_error() { return g_err.c_str(); }
void ora_free(void* p) { std::free(p); }

// process_utils.rs:78-121
double ora_hausdorff(const double* a_xy, long na, const double* b_xy, long nb) {
    return hausdorff_distance(pts_from_xy(a_xy, na), pts_from_xy(b_xy, nb));
}
double ora_directed_hausdorff(const double* a_xy, long na, const double* b_xy, long nb) {
    auto a = pts_from_xy(a_xy, na), b = pts_from_xy(b_xy, nb);
    return directed_hausdorff(a.data(), a.size(), b.data(), b.size());
}

// Grid of process_utils.rs:43-67. Returns the candidate count (angles written
// up to `cap`), or -1 when search_range would return early with *fallback.
long ora_search_grid(double step_deg, double range_deg, int has_center, double center, double limes_deg, double* out,
                     long cap, double* fallback) {
    Grid g = search_grid(step_deg, range_deg, has_center ? std::optional<double>(center) : std::nullopt, limes_deg);
    if (fallback) *fallback = g.fallback;
    if (g.degenerate) return -1;
    for (long i = 0; i < (long)g.angles.size() && i < cap; ++i) out[i] = g.angles[i];
    return (long)g.angles.size();
}

// search_range on the analytic costs of process_utils.rs:130-212.
// kind 0: (angle-param)^2, 1: sin(angle), 2: constant 1.0
double ora_search_range_analytic(int kind, double param, double step_deg, double range_deg, int has_center,
                                 double center, double limes_deg, int threads) {
    std::function<double(double)> f;
    if (kind == 0)
        f = [param](double a) { return (a - param) * (a - param); };  // powi(2)
    else if (kind == 1)
        f = [](double a) { return std::sin(a); };
    else
        f = [](double) { return 1.0; };
    return search_range(f, step_deg, range_deg, has_center ? std::optional<double>(center) : std::nullopt, limes_deg, threads).angle;
}

// One search_range call with the Hausdorff cost closure. mode 0 = intrapullback
// closure (align_within.rs:99-105, zero-angle shortcut), 1 = inter-pullback
// closure (align_between.rs:189-216). costs_out (optional) receives every
// candidate's cost.
int ora_sweep(const double* test_xy, long n, const double* ref_xy, long m, double cx, double cy, int mode,
              double step_deg, double range_deg, int has_center, double center, double limes_deg, int threads,
              double* best_angle, long* best_idx, double* best_cost, double* costs_out, long costs_cap) {
    return guard([&] {
        auto test = pts_from_xy(test_xy, n), ref = pts_from_xy(ref_xy, m);
        auto cost = mode == 0 ? within_cost(ref, test, cx, cy) : between_cost(ref, test, cx, cy);
        std::vector<double> costs;
        auto r = search_range(cost, step_deg, range_deg, has_center ? std::optional<double>(center) : std::nullopt,
                              limes_deg, threads, &costs);
        if (best_angle) *best_angle = r.angle;
        if (best_idx) *best_idx = r.index;
        if (best_cost) *best_cost = r.cost;
        if (costs_out)
            for (long i = 0; i < (long)costs.size() && i < costs_cap; ++i) costs_out[i] = costs[i];
    });
}

// Costs of an explicit list of angles (same closures as ora_sweep).
int ora_costs(const double* test_xy, long n, const double* ref_xy, long m, double cx, double cy, int mode,
              const double* angles, long k, double* out) {
    return guard([&] {
        auto test = pts_from_xy(test_xy, n), ref = pts_from_xy(ref_xy, m);
        auto cost = mode == 0 ? within_cost(ref, test, cx, cy) : between_cost(ref, test, cx, cy);
        for (long i = 0; i < k; ++i) out[i] = cost(angles[i]);
    });
}

// find_best_rotation (align_within.rs:193-247, mode 0) /
// find_best_rotation_between (align_between.rs:180-258, mode 1; the rotation
// centre is the reference cloud's mean, cx/cy ignored).
int ora_find_best_rotation(const double* test_xy, long n, const double* ref_xy, long m, double cx, double cy, int mode,
                           double step_deg, double range_deg, int threads, double* angle) {
    return guard([&] {
        auto test = pts_from_xy(test_xy, n), ref = pts_from_xy(ref_xy, m);
        *angle = mode == 0 ? find_best_rotation(ref, test, step_deg, range_deg, cx, cy, threads)
                           : find_best_rotation_between(ref, test, step_deg, range_deg, threads);
    });
}

long ora_downsample_indices(long len, long n, long* out) {  // contour.rs:47-58
    auto idx = downsample_indices((size_t)len, (size_t)n);
    for (size_t i = 0; i < idx.size(); ++i) out[i] = (long)idx[i];
    return (long)idx.size();
}

// io/build.rs:9-205 via io/input.rs:62-147
int ora_build_geometry_from_dir(const char* path, const char* label, int diastole, double icx, double icy,
                                double radius, unsigned n_points, double** blob, long* len) {
    return guard([&] {
        InputData in = process_directory(path, diastole != 0, label);
        Geometry g = build_geometry_from_inputdata(in, label, diastole != 0, icx, icy, radius, n_points);
        auto v = encode_geometry(g);
        *blob = dup_vec(v);
        *len = (long)v.size();
    });
}

// from (N,4) [frame, x, y, z] arrays (what PyInputData carries, py_input_data.rs:103-172).
// records: (R,4) [frame, phase(1=D,0=S), m1, m2] with NaN for missing measurements.
int ora_build_geometry_from_arrays(const double* lumen, long n_lumen, const double* eem, long n_eem,
                                   const double* calc, long n_calc, const double* side, long n_side,
                                   const double* records, long n_rec, const double* ref_point, int diastole,
                                   const char* label, double icx, double icy, double radius, unsigned n_points,
                                   double** blob, long* len) {
    return guard([&] {
        auto conv = [](const double* a, long n) {
            std::vector<ContourPoint> v((size_t)n);
            for (long i = 0; i < n; ++i) {
                v[i].frame_index = (uint32_t)a[4 * i];
                v[i].x = a[4 * i + 1];
                v[i].y = a[4 * i + 2];
                v[i].z = a[4 * i + 3];
            }
            return v;
        };
        InputData in;
        in.lumen = conv(lumen, n_lumen);
        if (eem) in.eem = conv(eem, n_eem);
        if (calc) in.calcification = conv(calc, n_calc);
        if (side) in.sidebranch = conv(side, n_side);
        if (records) {
            std::vector<Record> rs;
            for (long i = 0; i < n_rec; ++i) {
                Record r;
                r.frame = (uint32_t)records[4 * i];
                r.phase = records[4 * i + 1] != 0.0 ? "D" : "S";
                if (records[4 * i + 2] == records[4 * i + 2]) r.measurement_1 = records[4 * i + 2];
                if (records[4 * i + 3] == records[4 * i + 3]) r.measurement_2 = records[4 * i + 3];
                rs.push_back(r);
            }
            in.record = rs;
        }
        in.ref_point = conv(ref_point, 1)[0];
        in.diastole = diastole != 0;
        in.label = label;
        Geometry g = build_geometry_from_inputdata(in, label, diastole != 0, icx, icy, radius, n_points);
        auto v = encode_geometry(g);
        *blob = dup_vec(v);
        *len = (long)v.size();
    });
}

// align_within.rs:24-171. post_steps=0 stops after the frame chain (:72-134).
int ora_align_within(const double* blob, long len, double step_deg, double range_deg, int smooth, int bruteforce,
                     long sample_size, int threads, int post_steps, double** out_blob, long* out_len, double** logs,
                     long* n_logs, int* anomalous) {
    return guard([&] {
        Geometry g = decode_geometry(blob, (size_t)len, "geom");
        auto r = align_frames_in_geometry(g, step_deg, range_deg, smooth != 0, bruteforce != 0, (size_t)sample_size,
                                          threads, nullptr, post_steps != 0);
        auto v = encode_geometry(r.geometry);
        *out_blob = dup_vec(v);
        *out_len = (long)v.size();
        auto l = encode_logs(r.logs);
        *logs = dup_vec(l);
        *n_logs = (long)r.logs.size();
        *anomalous = r.anomalous ? 1 : 0;
    });
}

// The frame chain of align_within.rs:72-134 with a tap: for frame pair `pair` (0-based) returns the chain-state
// sample points the reference's cost closure sees (testing_points, reference_points), the rotation centre
// (current.centroid after the translation) and the chosen angle. Used to measure how far the "decoupled" costs
// the product sweeps are from the chain's own.
int ora_within_chain_tap(const double* blob, long len, double step_deg, double range_deg, int bruteforce,
                         long sample_size, int threads, long pair, double** test_xy, long* n_test, double** ref_xy,
                         long* n_ref, double centre[2], double* best) {
    return guard([&] {
        Geometry g = decode_geometry(blob, (size_t)len, "geom");
        ChainTap tap;
        align_frames_in_geometry(g, step_deg, range_deg, false, bruteforce != 0, (size_t)sample_size, threads, &tap, false);
        if (pair < 0 || pair >= (long)tap.test.size()) throw Error("pair index out of range");
        auto flat = [](const std::vector<ContourPoint>& v) {
            std::vector<double> o;
            for (auto& p : v) o.insert(o.end(), {p.x, p.y});
            return o;
        };
        auto t = flat(tap.test[pair]), r = flat(tap.ref[pair]);
        *test_xy = dup_vec(t);
        *n_test = (long)tap.test[pair].size();
        *ref_xy = dup_vec(r);
        *n_ref = (long)tap.ref[pair].size();
        centre[0] = tap.centre[pair].first;
        centre[1] = tap.centre[pair].second;
        *best = tap.best[pair];
    });
}

// align_between.rs:11-92. Returns the (unchanged) A and the moved B, plus the chosen rotation.
int ora_align_between(const double* blob_a, long len_a, const double* blob_b, long len_b, double rot_deg,
                      double step_deg, long sample_size, int threads, double** out_b, long* out_b_len,
                      double* best_angle) {
    return guard([&] {
        Geometry a = decode_geometry(blob_a, (size_t)len_a, "a"), b = decode_geometry(blob_b, (size_t)len_b, "b");
        BetweenTap tap;
        align_between_geometries(a, b, rot_deg, step_deg, (size_t)sample_size, threads, &tap);
        auto v = encode_geometry(b);
        *out_b = dup_vec(v);
        *out_b_len = (long)v.size();
        if (best_angle) *best_angle = tap.best;
    });
}

// Orchestration of binding/entry.rs. mode: 4 = full (:71-361), 3 = double pair
// (:363-570), 2 = single pair (:572-689), 1 = single (:691-780).
// in: `n_geoms` blobs (1, 2 or 4). out: for mode 4 eight blobs (ab.a, ab.b,
// cd.a, cd.b, ac.a, ac.b, bd.a, bd.b), mode 3 four, mode 2 two, mode 1 one;
// logs: one (n,7) array per input geometry. postprocessing != 0 applies postprocess_geom_pair
// (processing/postprocessing.rs:12-87) to every pair (modes 2-4). OBJ export is out of scope.
int ora_process(int mode, const double* const* blobs, const long* lens, double step_deg, double range_deg, int smooth,
                int bruteforce, long sample_size, int threads, int postprocessing, double** out_blobs, long* out_lens,
                double** out_logs, long* out_nlogs) {
    return guard([&] {
        ProcessParams p;
        p.step_deg = step_deg;
        p.range_deg = range_deg;
        p.smooth = smooth != 0;
        p.bruteforce = bruteforce != 0;
        p.sample_size = (size_t)sample_size;
        p.threads = threads;
        p.postprocessing = postprocessing != 0;
        int n_in = mode >= 3 ? 4 : mode;
        std::vector<Geometry> geoms;
        for (int i = 0; i < n_in; ++i) geoms.push_back(decode_geometry(blobs[i], (size_t)lens[i], "g" + std::to_string(i)));
        std::vector<const Geometry*> outs;
        std::vector<std::vector<AlignLog>> logs;
        FullResult fr;
        PairResult pr;
        WithinResult wr;
        if (mode >= 3) {
            fr = full_processing(geoms, p, mode == 3);
            outs = {&fr.ab.geom_a, &fr.ab.geom_b, &fr.cd.geom_a, &fr.cd.geom_b};
            if (mode == 4) outs.insert(outs.end(), {&fr.ac.geom_a, &fr.ac.geom_b, &fr.bd.geom_a, &fr.bd.geom_b});
            for (int i = 0; i < 4; ++i) logs.push_back(fr.logs[i]);
        } else if (mode == 2) {
            pr = pair_processing(geoms, p);
            outs = {&pr.pair.geom_a, &pr.pair.geom_b};
            logs = {pr.logs[0], pr.logs[1]};
        } else {
            wr = align_frames_in_geometry(geoms[0], step_deg, range_deg, p.smooth, p.bruteforce, p.sample_size, threads);
            outs = {&wr.geometry};
            logs = {wr.logs};
        }
        for (size_t i = 0; i < outs.size(); ++i) {
            auto v = encode_geometry(*outs[i]);
            out_blobs[i] = dup_vec(v);
            out_lens[i] = (long)v.size();
        }
        for (size_t i = 0; i < logs.size(); ++i) {
            out_logs[i] = dup_vec(encode_logs(logs[i]));
            out_nlogs[i] = (long)logs[i].size();
        }
    });
}

// postprocess_geom_pair (processing/postprocessing.rs:12-87) on one pair.
int ora_postprocess_pair(const double* blob_a, long len_a, const double* blob_b, long len_b, double tol, int anomalous,
                         double** out_a, long* out_a_len, double** out_b, long* out_b_len) {
    return guard([&] {
        GeometryPair p;
        p.geom_a = decode_geometry(blob_a, (size_t)len_a, "a");
        p.geom_b = decode_geometry(blob_b, (size_t)len_b, "b");
        p.label = "a - b";
        GeometryPair r = postprocess_geom_pair(p, tol, anomalous != 0);
        auto va = encode_geometry(r.geom_a), vb = encode_geometry(r.geom_b);
        *out_a = dup_vec(va);
        *out_a_len = (long)va.size();
        *out_b = dup_vec(vb);
        *out_b_len = (long)vb.size();
    });
}
// predict_z_positions (:142-195); returns the count.
long ora_predict_z_positions(double ref_z, double start_z, double stop_z, double z_diff, double* out, long cap) {
    auto z = predict_z_positions(ref_z, start_z, stop_z, z_diff);
    for (long i = 0; i < (long)z.size() && i < cap; ++i) out[i] = z[i];
    return (long)z.size();
}

// CPU baseline kernel for bench.py: the reference's loop nest (threads over
// candidate angles, serial N x M inside: process_utils.rs:69-118) on a batch of
// units that share one grid. Returns best index per unit; costs are discarded.
int ora_sweep_batch(const double* test_xy, const long* test_off, const double* ref_xy, const long* ref_off,
                    const double* centre_xy, long n_units, int mode, double step_deg, double range_deg,
                    double limes_deg, int threads, long* best_idx, double* best_cost) {
    return guard([&] {
        for (long u = 0; u < n_units; ++u) {
            auto test = pts_from_xy(test_xy + 2 * test_off[u], test_off[u + 1] - test_off[u]);
            auto ref = pts_from_xy(ref_xy + 2 * ref_off[u], ref_off[u + 1] - ref_off[u]);
            double cx = centre_xy[2 * u], cy = centre_xy[2 * u + 1];
            auto cost = mode == 0 ? within_cost(ref, test, cx, cy) : between_cost(ref, test, cx, cy);
            auto r = search_range(cost, step_deg, range_deg, std::nullopt, limes_deg, threads);
            best_idx[u] = r.index;
            best_cost[u] = r.cost;
        }
    });
}

}  // extern "C"
