// =============================================================================
// mmrs_oracle.hpp — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
//
// A plain C++17 / f64 restatement of the reference's (yungselm/multimoda-rs
// v0.7.0) brute-force / coarse-to-fine rotation sweep scored by symmetric
// Hausdorff distance, together with the host logic either side of it
// (ingest, frame chain, post steps, inter-pullback alignment, orchestration).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may compile, link, import or execute anything in this
// directory. The product (multimoda-rs_b200/) never does.
//
// PARITY STATUS: pinned against the reference's own known-answer tests (the
// in-file Rust #[test]s that fix this path: process_utils.rs:130-547,
// align_within.rs:791-1001, align_between.rs:281-373, contour.rs:547-604,
// geometry.rs:450-503) — see tests/test_oracle_kat.py. The reference itself
// cannot be built or imported here (no cargo/rustc/maturin; `import
// multimodars` fails), so there is no oracle/_ref and no reference-generated
// golden output; the golden fixtures under tests/golden/ are produced by THIS
// oracle (script committed beside them) and say so.
//
// Arithmetic rules that make this bit-faithful to the Rust code on the same
// glibc: f64 everywhere, compiled with -ffp-contract=off (rustc never fuses),
// sin/cos/atan2/sqrt/fmod from glibc libm (what Rust's f64 methods call on
// x86_64-unknown-linux-gnu), sequential left folds where Rust folds, stable
// sorts where Rust uses sort_by, "last maximum" where Rust uses max_by.
//
// Every function cites the reference file:line it follows (paths relative to
// the reference checkout root).
// =============================================================================
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <limits>
#include <map>
#include <optional>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

namespace ora {

constexpr double PI = 3.14159265358979323846264338327950288;  // std::f64::consts::PI
// f64::to_radians / to_degrees (Rust core: self * (PI/180), self * (180/PI)).
inline double to_radians(double d) { return d * (PI / 180.0); }
inline double to_degrees(double r) { return r * (180.0 / PI); }
// f64::rem_euclid (Rust core): r = self % rhs; if r < 0 { r + |rhs| } else { r }
inline double rem_euclid(double a, double b) {
    double r = std::fmod(a, b);
    return (r < 0.0) ? r + std::fabs(b) : r;
}
// Rust `as usize` on f64: saturating, NaN -> 0.
inline size_t f64_as_usize(double v) {
    if (!(v == v)) return 0;
    if (v <= 0.0) return 0;
    if (v >= 18446744073709551615.0) return std::numeric_limits<size_t>::max();
    return (size_t)v;
}

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// ---- types (src/types/native/contour_point.rs:55-67, contour.rs:8-44,
//      frame.rs:7-15, geometry.rs:8-12, geometry_pair.rs:4-20) --------------
struct ContourPoint {
    uint32_t frame_index = 0;
    uint32_t point_index = 0;
    double x = 0, y = 0, z = 0;
    bool aortic = false;
};

enum ContourType : int { Lumen = 0, Eem = 1, Calcification = 2, Sidebranch = 3, Catheter = 4, Wall = 5 };

using Vec3 = std::tuple<double, double, double>;

struct Contour {
    uint32_t id = 0;
    uint32_t original_frame = 0;
    std::vector<ContourPoint> points;
    std::optional<Vec3> centroid;
    std::optional<double> aortic_thickness;
    std::optional<double> pulmonary_thickness;
    ContourType kind = Lumen;
};

struct Frame {
    uint32_t id = 0;
    double cx = 0, cy = 0, cz = 0;  // centroid
    Contour lumen;
    // Rust uses HashMap<ContourType, Contour>; iteration order there is random
    // but every loop over it on this path is order-independent.
    std::map<ContourType, Contour> extras;
    std::optional<ContourPoint> reference_point;
};

struct Geometry {
    std::vector<Frame> frames;
    std::string label;
};

struct GeometryPair {
    Geometry geom_a, geom_b;
    std::string label;
};

struct Record {  // src/types/native/record.rs:3-11
    uint32_t frame = 0;
    std::string phase;
    std::optional<double> measurement_1, measurement_2;
};

struct InputData {  // src/intravascular/io/input.rs:27-37
    std::vector<ContourPoint> lumen;
    std::optional<std::vector<ContourPoint>> eem, calcification, sidebranch;
    std::optional<std::vector<Record>> record;
    ContourPoint ref_point;
    bool diastole = true;
    std::string label;
};

struct AlignLog {  // align_within.rs:14-22
    uint32_t contour_id, matched_to;
    double rot_deg, tx, ty, cx, cy;
};

// ---- point / contour / frame transforms ------------------------------------
// contour_point.rs:29-36
inline ContourPoint pt_translate(ContourPoint p, double dx, double dy, double dz) {
    p.x = p.x + dx;
    p.y = p.y + dy;
    p.z = p.z + dz;
    return p;
}
// contour_point.rs:38-52 (identity iff angle == 0.0; x*cos - y*sin + cx)
inline ContourPoint pt_rotate(ContourPoint p, double angle, double cx, double cy) {
    if (angle == 0.0) return p;
    double x = p.x - cx;
    double y = p.y - cy;
    double cos_a = std::cos(angle);
    double sin_a = std::sin(angle);
    p.x = x * cos_a - y * sin_a + cx;
    p.y = x * sin_a + y * cos_a + cy;
    return p;
}
// contour.rs:213-224
inline void compute_centroid(Contour& c) {
    if (c.points.empty()) {
        c.centroid.reset();
        return;
    }
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (const auto& p : c.points) {
        sx = sx + p.x;
        sy = sy + p.y;
        sz = sz + p.z;
    }
    double n = (double)c.points.size();
    c.centroid = Vec3(sx / n, sy / n, sz / n);
}
// contour.rs:61-66
inline void contour_translate(Contour& c, double dx, double dy, double dz) {
    for (auto& p : c.points) p = pt_translate(p, dx, dy, dz);
}
// contour.rs:68-76
inline void contour_rotate(Contour& c, double angle, double cx, double cy) {
    if (angle == 0.0) return;
    for (auto& p : c.points) p = pt_rotate(p, angle, cx, cy);
}
// frame.rs:18-38
inline void frame_translate(Frame& f, double dx, double dy, double dz) {
    contour_translate(f.lumen, dx, dy, dz);
    compute_centroid(f.lumen);
    for (auto& kv : f.extras) {
        contour_translate(kv.second, dx, dy, dz);
        compute_centroid(kv.second);
    }
    if (f.reference_point) f.reference_point = pt_translate(*f.reference_point, dx, dy, dz);
    f.cx += dx;
    f.cy += dy;
    f.cz += dz;
}
// frame.rs:40-63
inline void frame_rotate(Frame& f, double angle, double cx, double cy) {
    if (angle == 0.0) return;
    contour_rotate(f.lumen, angle, cx, cy);
    for (auto& kv : f.extras) contour_rotate(kv.second, angle, cx, cy);
    if (f.reference_point) f.reference_point = pt_rotate(*f.reference_point, angle, cx, cy);
    double x = f.cx - cx;
    double y = f.cy - cy;
    double cos_a = std::cos(angle);
    double sin_a = std::sin(angle);
    f.cx = x * cos_a - y * sin_a + cx;
    f.cy = x * sin_a + y * cos_a + cy;
}

// contour.rs:47-58
inline std::vector<size_t> downsample_indices(size_t len, size_t n) {
    std::vector<size_t> idx;
    if (len <= n) {
        idx.resize(len);
        for (size_t i = 0; i < len; ++i) idx[i] = i;
        return idx;
    }
    double step = (double)len / (double)n;
    idx.resize(n);
    for (size_t i = 0; i < n; ++i) idx[i] = f64_as_usize((double)i * step);
    return idx;
}
inline std::vector<ContourPoint> downsample_contour_points(const std::vector<ContourPoint>& pts, size_t n) {
    std::vector<ContourPoint> out;
    for (size_t i : downsample_indices(pts.size(), n)) out.push_back(pts[i]);
    return out;
}

// ---- the metric -------------------------------------------------------------
// process_utils.rs:84-121. The reference splits A into rayon chunks and takes
// f64::max over chunk maxima starting from 0.0 — order-independent, so a
// serial loop gives the identical value.
inline double directed_hausdorff(const ContourPoint* a, size_t na, const ContourPoint* b, size_t nb) {
    if (na == 0 || nb == 0) return 0.0;
    double max_sq = 0.0;
    for (size_t i = 0; i < na; ++i) {
        double min_sq = std::numeric_limits<double>::infinity();
        for (size_t j = 0; j < nb; ++j) {
            double dx = a[i].x - b[j].x;
            double dy = a[i].y - b[j].y;
            double d2 = dx * dx + dy * dy;
            if (d2 < min_sq) min_sq = d2;
        }
        if (std::isfinite(min_sq) && min_sq > max_sq) max_sq = min_sq;
    }
    return std::sqrt(max_sq);
}
// process_utils.rs:78-82
inline double hausdorff_distance(const std::vector<ContourPoint>& s1, const std::vector<ContourPoint>& s2) {
    double fwd = directed_hausdorff(s1.data(), s1.size(), s2.data(), s2.size());
    double bwd = directed_hausdorff(s2.data(), s2.size(), s1.data(), s1.size());
    return std::fmax(fwd, bwd);  // f64::max
}

// ---- the sweep ----------------------------------------------------------------
struct Grid {
    bool degenerate = false;   // early return: search_range returns `fallback`
    double fallback = 0.0;
    std::vector<double> angles;  // wrapped to [-pi, pi), candidate order
};
// process_utils.rs:43-67 (grid construction part of search_range)
inline Grid search_grid(double step_deg, double range_deg, std::optional<double> center_angle, double limes_deg) {
    Grid g;
    double range_rad = to_radians(range_deg);
    double step_rad = to_radians(step_deg);
    if (step_rad <= 0.0) {
        g.degenerate = true;
        g.fallback = center_angle.value_or(0.0);
        return g;
    }
    double center = center_angle.value_or(0.0);
    double limes = to_radians(limes_deg);
    double start_angle = std::fmax(center - range_rad, -limes);
    double stop_angle = std::fmin(center + range_rad, limes);
    if (stop_angle <= start_angle) {
        g.degenerate = true;
        g.fallback = center;
        return g;
    }
    size_t steps = std::max<size_t>(f64_as_usize(std::ceil((stop_angle - start_angle) / step_rad)), 1);
    for (size_t i = 0; i <= steps; ++i) {
        double a = start_angle + (double)i * step_rad;
        if (!(a <= stop_angle)) break;  // take_while
        g.angles.push_back(rem_euclid(a + PI, 2.0 * PI) - PI);
    }
    g.fallback = center;  // .unwrap_or(center) on an empty candidate list
    return g;
}

struct SweepResult {
    double angle = 0.0;
    long index = -1;  // -1: degenerate grid (no candidate evaluated)
    double cost = 0.0;
};
// process_utils.rs:69-74: evaluate every candidate, leftmost arg-min (strict <).
// `threads` > 1 splits the candidate list into contiguous blocks (what rayon's
// indexed par_iter does); the combine keeps the left operand on ties, so the
// result is the same for any thread count.
inline SweepResult search_range(const std::function<double(double)>& cost_fn, double step_deg, double range_deg,
                                std::optional<double> center_angle, double limes_deg, int threads = 1,
                                std::vector<double>* costs_out = nullptr) {
    Grid g = search_grid(step_deg, range_deg, center_angle, limes_deg);
    SweepResult r;
    if (g.degenerate || g.angles.empty()) {
        r.angle = g.fallback;
        return r;
    }
    const size_t n = g.angles.size();
    std::vector<double> costs(n);
    if (threads <= 1) {
        for (size_t i = 0; i < n; ++i) costs[i] = cost_fn(g.angles[i]);
    } else {
        std::vector<std::thread> pool;
        size_t per = (n + threads - 1) / threads;
        for (int t = 0; t < threads; ++t) {
            size_t lo = t * per, hi = std::min(n, lo + per);
            if (lo >= hi) break;
            pool.emplace_back([&, lo, hi] {
                for (size_t i = lo; i < hi; ++i) costs[i] = cost_fn(g.angles[i]);
            });
        }
        for (auto& th : pool) th.join();
    }
    size_t best = 0;
    for (size_t i = 1; i < n; ++i)
        if (costs[i] < costs[best]) best = i;
    r.angle = g.angles[best];
    r.index = (long)best;
    r.cost = costs[best];
    if (costs_out) *costs_out = std::move(costs);
    return r;
}

// The 1..4-stage coarse-to-fine driver shared by align_within.rs:208-246 and
// align_between.rs:219-257 (identical match arms: [1,inf], [0.1,1), [0.01,0.1), else).
inline double coarse_to_fine(const std::function<double(double)>& cost_fn, double step_deg, double range_deg,
                             int threads = 1) {
    auto sr = [&](double step, double range, std::optional<double> c) {
        return search_range(cost_fn, step, range, c, range_deg, threads).angle;
    };
    if (step_deg >= 1.0) {  // 1.0..=INFINITY (NaN falls to the last arm)
        return sr(step_deg, range_deg, std::nullopt);
    } else if (step_deg >= 0.1 && step_deg < 1.0) {
        double coarse = sr(1.0, range_deg, std::nullopt);
        double range = (range_deg > 5.0) ? 5.0 : range_deg;
        return sr(step_deg, range, coarse);
    } else if (step_deg >= 0.01 && step_deg < 0.1) {
        double coarse = sr(1.0, range_deg, std::nullopt);
        double range = (range_deg > 5.0) ? 5.0 : range_deg;
        double medium = sr(0.1, range, coarse);
        double range_small = (range_deg > 10.0 * step_deg) ? 10.0 * step_deg : range_deg;
        return sr(step_deg, range_small, medium);
    } else {
        double coarse = sr(1.0, range_deg, std::nullopt);
        double range = (range_deg > 5.0) ? 5.0 : range_deg;
        double medium = sr(0.1, range, coarse);
        double range_small = (range_deg > 0.1) ? 0.1 : range_deg;
        double fine = sr(0.01, range_small, medium);
        double range_fine = (range_deg > 10.0 * step_deg) ? 10.0 * step_deg : range_deg;
        return sr(step_deg, range_fine, fine);
    }
}

// Cost closure of the intrapullback path: align_within.rs:99-105 / :200-206.
inline std::function<double(double)> within_cost(const std::vector<ContourPoint>& reference,
                                                 const std::vector<ContourPoint>& target, double cx, double cy) {
    return [&reference, &target, cx, cy](double angle) {
        std::vector<ContourPoint> rotated(target.size());
        for (size_t i = 0; i < target.size(); ++i) rotated[i] = pt_rotate(target[i], angle, cx, cy);
        return hausdorff_distance(reference, rotated);
    };
}
// align_within.rs:193-247
inline double find_best_rotation(const std::vector<ContourPoint>& reference, const std::vector<ContourPoint>& target,
                                 double step_deg, double range_deg, double cx, double cy, int threads = 1) {
    return coarse_to_fine(within_cost(reference, target, cx, cy), step_deg, range_deg, threads);
}

// Cost closure of the inter-pullback path: align_between.rs:189-216 (no
// zero-angle shortcut; sin/cos recomputed per point but identical per angle).
inline std::function<double(double)> between_cost(const std::vector<ContourPoint>& reference,
                                                  const std::vector<ContourPoint>& target, double rcx, double rcy) {
    return [&reference, &target, rcx, rcy](double angle) {
        std::vector<ContourPoint> rotated(target.size());
        double cos_angle = std::cos(angle);
        double sin_angle = std::sin(angle);
        for (size_t i = 0; i < target.size(); ++i) {
            double tx = target[i].x - rcx;
            double ty = target[i].y - rcy;
            double rx = tx * cos_angle - ty * sin_angle;
            double ry = tx * sin_angle + ty * cos_angle;
            rotated[i] = target[i];
            rotated[i].x = rx + rcx;
            rotated[i].y = ry + rcy;
        }
        return hausdorff_distance(reference, rotated);
    };
}
// align_between.rs:260-271
inline Vec3 global_centroid(const std::vector<ContourPoint>& pts) {
    if (pts.empty()) return Vec3(0.0, 0.0, 0.0);
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (auto& p : pts) sx += p.x;
    for (auto& p : pts) sy += p.y;
    for (auto& p : pts) sz += p.z;
    double n = (double)pts.size();
    return Vec3(sx / n, sy / n, sz / n);
}
// align_between.rs:180-258
inline double find_best_rotation_between(const std::vector<ContourPoint>& reference,
                                         const std::vector<ContourPoint>& target, double step_deg, double range_deg,
                                         int threads = 1) {
    auto [rcx, rcy, rcz] = global_centroid(reference);
    (void)rcz;
    return coarse_to_fine(between_cost(reference, target, rcx, rcy), step_deg, range_deg, threads);
}

// ---- Contour geometry helpers (contour.rs) ---------------------------------
inline double dist3(const ContourPoint& a, const ContourPoint& b) {  // native.rs:27-32
    double dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    return std::sqrt(dx * dx + dy * dy + dz * dz);
}
// contour.rs:227-243
inline std::tuple<size_t, size_t, double> find_farthest_points(const Contour& c) {
    double max_dist = 0.0;
    size_t bi = 0, bj = 0;
    for (size_t i = 0; i < c.points.size(); ++i)
        for (size_t j = i + 1; j < c.points.size(); ++j) {
            double d = dist3(c.points[i], c.points[j]);
            if (d > max_dist) {
                max_dist = d;
                bi = i;
                bj = j;
            }
        }
    return {bi, bj, max_dist};
}
// contour.rs:313-333
inline double find_closest_opposite_3d(const Contour& c) {
    size_t n = c.points.size();
    if (n <= 2) throw Error("Need at least 3 points");
    size_t half = n / 2;
    double min_dist = std::numeric_limits<double>::max();
    for (size_t i = 0; i < n; ++i) {
        size_t j = (i + half) % n;
        double d = dist3(c.points[i], c.points[j]);
        if (d < min_dist) min_dist = d;
    }
    return min_dist;
}
// contour.rs:335-343
inline double elliptic_ratio(const Contour& c) {
    double major = std::get<2>(find_farthest_points(c));
    double minor = find_closest_opposite_3d(c);
    return (major < minor) ? minor / major : major / minor;
}
// contour.rs:368-405 (stable sort by atan2 ascending; LAST max-y to front; re-index)
inline void sort_contour_points(Contour& c) {
    double n = (double)c.points.size();
    if (n == 0.0) return;
    double sx = 0.0, sy = 0.0;
    for (auto& p : c.points) {
        sx = sx + p.x;
        sy = sy + p.y;
    }
    double cx = sx / n, cy = sy / n;
    std::stable_sort(c.points.begin(), c.points.end(), [cx, cy](const ContourPoint& a, const ContourPoint& b) {
        return std::atan2(a.y - cy, a.x - cx) < std::atan2(b.y - cy, b.x - cx);
    });
    size_t start = 0;
    for (size_t i = 1; i < c.points.size(); ++i)
        if (!(c.points[i].y < c.points[start].y)) start = i;  // Iterator::max_by keeps the last maximum
    std::rotate(c.points.begin(), c.points.begin() + start, c.points.end());
    for (size_t i = 0; i < c.points.size(); ++i) c.points[i].point_index = (uint32_t)i;
}
inline void sort_frame_points(Frame& f) {  // frame.rs:120-126
    sort_contour_points(f.lumen);
    for (auto& kv : f.extras) sort_contour_points(kv.second);
}

// ---- Geometry helpers (geometry.rs) ------------------------------------------
inline size_t find_proximal_end_idx(const Geometry& g) {  // :42-60
    size_t n = g.frames.size();
    if (n == 0) return 0;
    if (n == 1) return g.frames[0].lumen.id;
    return (g.frames[0].lumen.original_frame > g.frames[n - 1].lumen.original_frame) ? g.frames[0].lumen.id
                                                                                     : g.frames[n - 1].lumen.id;
}
inline std::optional<size_t> find_ref_frame_idx(const Geometry& g) {  // :62-69
    for (auto& f : g.frames)
        if (f.reference_point) return (size_t)f.id;
    return std::nullopt;
}
inline size_t ref_or_proximal(const Geometry& g) {
    auto r = find_ref_frame_idx(g);
    return r ? *r : find_proximal_end_idx(g);
}
inline void rotate_geometry(Geometry& g, double angle) {  // :241-250
    if (angle == 0.0) return;
    for (auto& f : g.frames) {
        frame_rotate(f, angle, f.cx, f.cy);
        sort_frame_points(f);
    }
}
inline void translate_geometry(Geometry& g, double dx, double dy, double dz) {  // :278-283
    for (auto& f : g.frames) frame_translate(f, dx, dy, dz);
}
inline void insert_frame(Geometry& g, Frame fr, std::optional<size_t> idx) {  // :285-323
    double z = fr.cz;
    size_t pos;
    if (idx)
        pos = *idx;
    else {
        pos = g.frames.size();
        for (size_t i = 0; i < g.frames.size(); ++i)
            if (g.frames[i].cz > z) {
                pos = i;
                break;
            }
    }
    g.frames.insert(g.frames.begin() + pos, std::move(fr));
    for (size_t i = 0; i < g.frames.size(); ++i) {
        Frame& f = g.frames[i];
        uint32_t nid = (uint32_t)i;
        f.id = nid;
        f.lumen.id = nid;
        for (auto& p : f.lumen.points) p.frame_index = nid;
        for (auto& kv : f.extras) {
            kv.second.id = nid;
            for (auto& p : kv.second.points) p.frame_index = nid;
        }
        if (f.reference_point) f.reference_point->frame_index = nid;
    }
}
// geometry.rs:165-239
inline Geometry smooth_frames(const Geometry& g) {
    Geometry out;
    out.label = g.label;
    const size_t nf = g.frames.size();
    for (size_t i = 0; i < nf; ++i) {
        Frame cur = g.frames[i];
        size_t point_count = cur.lumen.points.size();
        auto smooth_contour = [point_count](const Contour& c, const Contour& prev, const Contour& next) {
            Contour o;
            o.id = c.id;
            o.original_frame = c.original_frame;
            o.aortic_thickness = c.aortic_thickness;
            o.pulmonary_thickness = c.pulmonary_thickness;
            o.kind = c.kind;
            for (size_t j = 0; j < point_count; ++j) {
                const ContourPoint& cp = c.points.at(j);
                const ContourPoint& pp = prev.points.at(j);
                const ContourPoint& np = next.points.at(j);
                ContourPoint q = cp;
                q.x = (pp.x + cp.x + np.x) / 3.0;
                q.y = (pp.y + cp.y + np.y) / 3.0;
                o.points.push_back(q);
            }
            compute_centroid(o);
            return o;
        };
        const Frame& prev = (i == 0) ? g.frames[i] : g.frames[i - 1];
        const Frame& next = (i == nf - 1) ? g.frames[i] : g.frames[i + 1];
        cur.lumen = smooth_contour(cur.lumen, prev.lumen, next.lumen);
        for (ContourType kind : {Eem, Wall}) {
            auto it = cur.extras.find(kind);
            if (it != cur.extras.end()) {
                auto ip = prev.extras.find(kind), in = next.extras.find(kind);
                if (ip != prev.extras.end() && in != next.extras.end())
                    it->second = smooth_contour(it->second, ip->second, in->second);
            }
        }
        out.frames.push_back(std::move(cur));
    }
    return out;
}

// ---- wall.rs -------------------------------------------------------------------
// wall.rs:52-103
inline Contour offset_contour(const Contour& cin, double distance, std::optional<std::pair<uint32_t, uint32_t>> range) {
    Contour c = cin;
    compute_centroid(c);
    auto [cx, cy, cz] = *c.centroid;
    Contour o;
    o.id = c.id;
    o.original_frame = c.original_frame;
    o.centroid = c.centroid;
    o.aortic_thickness = c.aortic_thickness;
    o.pulmonary_thickness = c.pulmonary_thickness;
    o.kind = Wall;
    for (const auto& pt : c.points) {
        ContourPoint p = pt;
        bool do_offset = range ? (pt.point_index >= range->first && pt.point_index <= range->second) : true;
        if (do_offset) {
            double dx = pt.x - cx, dy = pt.y - cy, dz = pt.z - cz;
            double len = std::sqrt(dx * dx + dy * dy + dz * dz);
            if (len > std::numeric_limits<double>::epsilon()) {
                double ux = dx / len, uy = dy / len, uz = dz / len;
                p.x += ux * distance;
                p.y += uy * distance;
                p.z += uz * distance;
            }
        }
        o.points.push_back(p);
    }
    return o;
}
// wall.rs:112-210
inline Contour create_aortic_wall(const Contour& c) {
    size_t n = c.points.size();
    size_t first_quarter = n / 4, half = n / 2, third_quarter = first_quarter * 3;
    const ContourPoint& ref_pt = c.points.at(third_quarter);
    double thickness = c.aortic_thickness.value();
    double outer_x = ref_pt.x + thickness;
    double z = ref_pt.z;
    std::pair<double, double> up_mid{c.points[0].x, c.points[0].y + 1.0};
    std::pair<double, double> up_right{outer_x, up_mid.second};
    std::pair<double, double> low_mid{c.points[half].x, c.points[half].y - 1.0};
    std::pair<double, double> low_right{outer_x, low_mid.second};
    double dist_up = std::fabs(up_right.first - up_mid.first);
    double dist_right = std::fabs(up_right.second - low_right.second);
    double dist_low = std::fabs(low_right.first - low_mid.first);
    double total_dist = dist_up + dist_right + dist_low;
    size_t n_up = f64_as_usize(std::round(dist_up / total_dist * (double)half));
    size_t n_mid = f64_as_usize(std::round(dist_right / total_dist * (double)half));
    if (n_up + n_mid > half) throw Error("attempt to subtract with overflow (create_aortic_wall)");
    size_t n_low = half - n_up - n_mid;
    std::vector<std::pair<double, double>> right_points;
    for (size_t i = 0; i < n_low; ++i) {
        double t = (double)i / (double)(n_low - 1);
        right_points.push_back({low_mid.first + t * (low_right.first - low_mid.first), low_mid.second});
    }
    for (size_t i = 0; i < n_mid; ++i) {
        double t = (double)i / (double)(n_mid - 1);
        right_points.push_back({low_right.first, low_right.second + t * (up_right.second - low_right.second)});
    }
    for (size_t i = 0; i < n_up; ++i) {
        double t = (double)i / (double)(std::max<size_t>(n_up, 1) - 1);
        right_points.push_back({up_right.first - t * (up_right.first - up_mid.first), up_right.second});
    }
    std::vector<ContourPoint> left_wall = offset_contour(c, 1.0, std::make_pair(0u, (uint32_t)half)).points;
    if (left_wall.size() % 2 != 0)
        left_wall.resize(std::min(left_wall.size(), half + 1));
    else
        left_wall.resize(std::min(left_wall.size(), half));
    size_t left_len = left_wall.size();
    std::vector<ContourPoint> pts = left_wall;
    for (size_t i = 0; i < right_points.size(); ++i) {
        size_t src_index = left_len + i;
        if (src_index >= c.points.size()) throw Error("Index out of bounds (create_aortic_wall)");
        ContourPoint q = c.points[src_index];
        q.x = right_points[i].first;
        q.y = right_points[i].second;
        q.z = z;
        pts.push_back(q);
    }
    Contour o;
    o.id = c.id;
    o.original_frame = c.original_frame;
    o.points = std::move(pts);
    o.centroid = c.centroid;
    o.aortic_thickness = c.aortic_thickness;
    o.pulmonary_thickness = c.pulmonary_thickness;
    o.kind = Wall;
    return o;
}
// wall.rs:7-43
inline std::vector<Frame> create_wall_frames(const std::vector<Frame>& frames, bool anomalous) {
    std::vector<Frame> out;
    for (const Frame& f : frames) {
        auto aortic_only = [](const Contour& c) {
            return c.aortic_thickness ? create_aortic_wall(c) : offset_contour(c, 1.0, std::nullopt);
        };
        auto eem = f.extras.find(Eem);
        Contour w = (anomalous || eem == f.extras.end()) ? aortic_only(f.lumen) : aortic_only(eem->second);
        Frame nf = f;
        nf.extras[Wall] = std::move(w);
        out.push_back(std::move(nf));
    }
    return out;
}

// ---- align_within.rs post steps ---------------------------------------------
inline double median(std::vector<double> v) {  // :333-344
    if (v.empty()) return 0.0;
    std::sort(v.begin(), v.end());
    size_t n = v.size();
    return (n % 2 == 1) ? v[n / 2] : (v[n / 2 - 1] + v[n / 2]) / 2.0;
}
inline std::pair<bool, double> detect_holes(const Geometry& g) {  // :348-370
    std::vector<double> zd;
    for (size_t i = 1; i < g.frames.size(); ++i) zd.push_back(std::fabs(g.frames[i].cz - g.frames[i - 1].cz));
    if (zd.empty()) return {false, 0.0};
    double baseline = median(zd);
    if (baseline <= std::numeric_limits<double>::epsilon()) return {false, baseline};
    bool hole = false;
    for (double d : zd)
        if (d >= 1.5 * baseline) hole = true;
    return {hole, baseline};
}
inline std::optional<double> interp_opt(std::optional<double> a, std::optional<double> b, double t) {
    if (a && b) return *a + (*b - *a) * t;
    if (a) return a;
    if (b) return b;
    return std::nullopt;
}
inline std::optional<double> avg_opt(std::optional<double> a, std::optional<double> b) {
    if (a && b) return (*a + *b) / 2.0;
    if (a) return a;
    if (b) return b;
    return std::nullopt;
}
inline ContourPoint interp_point(const ContourPoint& p1, const ContourPoint& p2, double t, uint32_t fi, uint32_t pi_) {
    ContourPoint q;
    q.frame_index = fi;
    q.point_index = pi_;
    q.x = p1.x + (p2.x - p1.x) * t;
    q.y = p1.y + (p2.y - p1.y) * t;
    q.z = p1.z + (p2.z - p1.z) * t;
    q.aortic = p1.aortic || p2.aortic;
    return q;
}
inline Contour fill_frame_gap(const Contour& c1, const Contour& c2, double t, uint32_t id, uint32_t of) {  // :573-598
    Contour o;
    size_t len = std::min(c1.points.size(), c2.points.size());
    for (size_t i = 0; i < len; ++i) o.points.push_back(interp_point(c1.points[i], c2.points[i], t, of, (uint32_t)i));
    o.id = id;
    o.original_frame = of;
    if (c1.centroid && c2.centroid) {
        auto [ax, ay, az] = *c1.centroid;
        auto [bx, by, bz] = *c2.centroid;
        o.centroid = Vec3(ax + (bx - ax) * t, ay + (by - ay) * t, az + (bz - az) * t);
    } else if (c1.centroid)
        o.centroid = c1.centroid;
    else if (c2.centroid)
        o.centroid = c2.centroid;
    o.aortic_thickness = interp_opt(c1.aortic_thickness, c2.aortic_thickness, t);
    o.pulmonary_thickness = interp_opt(c1.pulmonary_thickness, c2.pulmonary_thickness, t);
    o.kind = c1.kind;
    return o;
}
inline Contour avg_contour(const Contour& c1, const Contour& c2, uint32_t id, uint32_t of) {  // :476-497
    Contour o;
    size_t len = std::min(c1.points.size(), c2.points.size());
    for (size_t i = 0; i < len; ++i) {
        ContourPoint q;
        q.frame_index = of;
        q.point_index = (uint32_t)i;
        q.x = (c1.points[i].x + c2.points[i].x) / 2.0;
        q.y = (c1.points[i].y + c2.points[i].y) / 2.0;
        q.z = (c1.points[i].z + c2.points[i].z) / 2.0;
        q.aortic = c1.points[i].aortic || c2.points[i].aortic;
        o.points.push_back(q);
    }
    o.id = id;
    o.original_frame = of;
    if (c1.centroid && c2.centroid) {
        auto [ax, ay, az] = *c1.centroid;
        auto [bx, by, bz] = *c2.centroid;
        o.centroid = Vec3((ax + bx) / 2.0, (ay + by) / 2.0, (az + bz) / 2.0);
    } else if (c1.centroid)
        o.centroid = c1.centroid;
    else if (c2.centroid)
        o.centroid = c2.centroid;
    o.aortic_thickness = avg_opt(c1.aortic_thickness, c2.aortic_thickness);
    o.pulmonary_thickness = avg_opt(c1.pulmonary_thickness, c2.pulmonary_thickness);
    o.kind = c1.kind;
    return o;
}
template <class F>
inline std::map<ContourType, Contour> merge_extras(const Frame& f1, const Frame& f2, F&& both) {
    std::map<ContourType, Contour> ex;
    for (auto* src : {&f1.extras, &f2.extras})
        for (auto& kv : *src) {
            if (ex.count(kv.first)) continue;
            auto i1 = f1.extras.find(kv.first), i2 = f2.extras.find(kv.first);
            if (i1 != f1.extras.end() && i2 != f2.extras.end())
                ex[kv.first] = both(i1->second, i2->second);
            else if (i1 != f1.extras.end())
                ex[kv.first] = i1->second;
            else
                ex[kv.first] = i2->second;
        }
    return ex;
}
inline Frame fix_one_frame_hole(const Frame& f1, const Frame& f2) {  // :499-543
    Frame o;
    o.cx = (f1.cx + f2.cx) / 2.0;
    o.cy = (f1.cy + f2.cy) / 2.0;
    o.cz = (f1.cz + f2.cz) / 2.0;
    o.lumen = avg_contour(f1.lumen, f2.lumen, f2.lumen.id, f2.lumen.original_frame);
    o.extras = merge_extras(f1, f2, [](const Contour& a, const Contour& b) {
        return avg_contour(a, b, b.id, b.original_frame);
    });
    o.id = f2.id;
    return o;
}
inline Frame create_interpolated_frame(const Frame& f1, const Frame& f2, double t) {  // :600-651
    Frame o;
    o.cx = f1.cx + (f2.cx - f1.cx) * t;
    o.cy = f1.cy + (f2.cy - f1.cy) * t;
    o.cz = f1.cz + (f2.cz - f1.cz) * t;
    o.lumen = fill_frame_gap(f1.lumen, f2.lumen, t, f2.lumen.id, f2.lumen.original_frame);
    o.extras = merge_extras(f1, f2, [t](const Contour& a, const Contour& b) {
        return fill_frame_gap(a, b, t, b.id, b.original_frame);
    });
    if (f1.reference_point && f2.reference_point)
        o.reference_point = interp_point(*f1.reference_point, *f2.reference_point, t, f2.id, 0);
    else if (f1.reference_point)
        o.reference_point = f1.reference_point;
    else if (f2.reference_point)
        o.reference_point = f2.reference_point;
    o.id = f2.id;
    return o;
}
// align_within.rs:378-449
inline Geometry fill_holes(Geometry& g) {
    auto [hole, baseline] = detect_holes(g);
    if (!hole) return g;
    if (baseline <= std::numeric_limits<double>::epsilon()) throw Error("Baseline spacing is zero or too small to decide.");
    size_t i = 1;
    while (i < g.frames.size()) {
        Frame prev = g.frames[i - 1];
        Frame curr = g.frames[i];
        double diff = std::fabs(curr.cz - prev.cz);
        double ratio = diff / baseline;
        if (ratio < 1.5) {
            i += 1;
        } else if (ratio >= 1.5 && ratio < 2.5) {
            insert_frame(g, fix_one_frame_hole(prev, curr), i);
            i += 2;
        } else if (ratio >= 2.5 && ratio < 3.5) {
            insert_frame(g, create_interpolated_frame(prev, curr, 1.0 / 3.0), i);
            insert_frame(g, create_interpolated_frame(prev, curr, 2.0 / 3.0), i + 1);
            i += 3;
        } else {
            size_t missing = f64_as_usize(std::fmax(std::floor(ratio - 1.0), 1.0));
            for (size_t k = 1; k <= missing; ++k) {
                double t = (double)k / (double)(missing + 1);
                insert_frame(g, create_interpolated_frame(prev, curr, t), i + k - 1);
            }
            i += missing + 1;
        }
    }
    return g;
}
// align_within.rs:249-254
inline bool is_anomalous_coronary(const Frame& ref) {
    return elliptic_ratio(ref.lumen) > 2.0 || ref.lumen.aortic_thickness.has_value() ||
           ref.lumen.pulmonary_thickness.has_value();
}
// align_within.rs:256-317
inline double angle_ref_point_to_right(const Frame& ref, bool anomalous) {
    if (!ref.reference_point) throw Error("No reference point found in frame");
    ContourPoint rp = *ref.reference_point;
    double p1x, p1y, p2x, p2y;
    if (anomalous) {
        auto [i, j, d] = find_farthest_points(ref.lumen);
        (void)d;
        p1x = ref.lumen.points[i].x;
        p1y = ref.lumen.points[i].y;
        p2x = ref.lumen.points[j].x;
        p2y = ref.lumen.points[j].y;
    } else {
        p1x = ref.cx;
        p1y = ref.cy;
        p2x = rp.x;
        p2y = rp.y;
    }
    double dx = p2x - p1x, dy = p2y - p1y;
    double line_angle = std::atan2(dy, dx);
    double desired = anomalous ? (PI / 2.0) : 0.0;  // FRAC_PI_2
    double rotation = rem_euclid(desired - line_angle, 2.0 * PI);
    auto rotate2 = [](double px, double py, double cx, double cy, double angle) {
        double ddx = px - cx, ddy = py - cy;
        double c = std::cos(angle), s = std::sin(angle);
        double xr = ddx * c - ddy * s, yr = ddx * s + ddy * c;
        return std::make_pair(xr + cx, yr + cy);
    };
    auto rotated_ref = rotate2(rp.x, rp.y, p1x, p1y, rotation);
    bool all_good = true;
    const double others[2][2] = {{p1x, p1y}, {p2x, p2y}};
    const double eps = std::numeric_limits<double>::epsilon();
    for (auto& op : others) {
        if (std::fabs(op[0] - rp.x) <= eps && std::fabs(op[1] - rp.y) <= eps) continue;  // approx::abs_diff_eq!
        auto r_op = rotate2(op[0], op[1], p1x, p1y, rotation);
        if (rotated_ref.first <= r_op.first) {
            all_good = false;
            break;
        }
    }
    if (!all_good) rotation = rem_euclid(rotation + PI, 2.0 * PI);
    return rotation;
}
inline void assign_aortic(Geometry& g) {  // :319-331
    for (auto& f : g.frames) {
        size_t len = f.lumen.points.size();
        if (len == 0) continue;
        size_t half = len / 2;
        for (size_t i = 0; i < len; ++i) f.lumen.points[i].aortic = i >= half;
    }
}

// align_within.rs:173-191
inline std::vector<ContourPoint> catheter_lumen_vec_from_frames(const Frame& f, size_t sample_lumen,
                                                               std::optional<size_t> sample_cath) {
    std::vector<ContourPoint> pts = downsample_contour_points(f.lumen.points, sample_lumen);
    if (sample_cath) {
        auto it = f.extras.find(Catheter);
        if (it != f.extras.end()) {
            auto c = downsample_contour_points(it->second.points, *sample_cath);
            pts.insert(pts.end(), c.begin(), c.end());
        }
    }
    return pts;
}

struct WithinResult {
    Geometry geometry;
    std::vector<AlignLog> logs;
    bool anomalous = false;
};
// Optional tap: per frame-pair candidate statistics, for building sweep test vectors.
struct ChainTap {
    // chain-state sample points, rotation centre and chosen angle per frame pair
    std::vector<std::vector<ContourPoint>> test, ref;
    std::vector<std::pair<double, double>> centre;
    std::vector<double> best;
};
// align_within.rs:24-171
inline WithinResult align_frames_in_geometry(Geometry& geometry, double step_deg, double range_deg, bool smooth,
                                             bool bruteforce, size_t sample_size, int threads = 1,
                                             ChainTap* tap = nullptr, bool post_steps = true) {
    if (geometry.frames.empty()) throw Error("Geometry contains no frames");
    if (geometry.frames[0].lumen.points.empty()) throw Error("Lumen contours have no points");
    if (sample_size == 0) throw Error("sample_size must be > 0");

    size_t ref_idx = ref_or_proximal(geometry);
    double sample_ratio = (double)sample_size / (double)geometry.frames[0].lumen.points.size();
    std::optional<size_t> sample_cath;
    {
        auto it = geometry.frames[0].extras.find(Catheter);
        if (it != geometry.frames[0].extras.end())
            sample_cath = f64_as_usize(std::ceil((double)it->second.points.size() * sample_ratio));
    }
    std::vector<AlignLog> logs;
    double cumulative_rotation = 0.0;
    for (size_t i = 1; i < geometry.frames.size(); ++i) {
        Frame prev = geometry.frames[i - 1];
        Frame& cur = geometry.frames[i];
        if (cumulative_rotation != 0.0) frame_rotate(cur, cumulative_rotation, cur.cx, cur.cy);
        double tx = prev.cx - cur.cx, ty = prev.cy - cur.cy;
        frame_translate(cur, tx, ty, 0.0);
        auto testing = catheter_lumen_vec_from_frames(cur, sample_size, sample_cath);
        auto reference = catheter_lumen_vec_from_frames(prev, sample_size, sample_cath);
        double best;
        if (bruteforce)
            best = search_range(within_cost(reference, testing, cur.cx, cur.cy), step_deg, range_deg, std::nullopt,
                                range_deg, threads)
                       .angle;
        else
            best = find_best_rotation(reference, testing, step_deg, range_deg, cur.cx, cur.cy, threads);
        if (tap) {
            tap->test.push_back(testing);
            tap->ref.push_back(reference);
            tap->centre.push_back({cur.cx, cur.cy});
            tap->best.push_back(best);
        }
        frame_rotate(cur, best, cur.cx, cur.cy);
        cumulative_rotation += best;
        logs.push_back({cur.id, prev.id, to_degrees(best), tx, ty, cur.cx, cur.cy});
    }
    WithinResult res;
    res.logs = std::move(logs);
    if (!post_steps) {
        res.geometry = geometry;
        return res;
    }
    Geometry g = fill_holes(geometry);  // fix_spacing is a clone (:653-656)
    if (ref_idx >= g.frames.size()) throw Error("index out of bounds: reference frame");
    bool anomalous = is_anomalous_coronary(g.frames[ref_idx]);
    double additional = angle_ref_point_to_right(g.frames[ref_idx], anomalous);
    rotate_geometry(g, additional);
    if (anomalous) assign_aortic(g);
    g.frames = create_wall_frames(g.frames, anomalous);
    if (smooth) g = smooth_frames(g);
    res.geometry = std::move(g);
    res.anomalous = anomalous;
    return res;
}

// ---- align_between.rs ----------------------------------------------------------
// :154-178
inline std::vector<ContourPoint> extract_geometry_points(const Geometry& g, size_t sample_size) {
    size_t total = 0;
    for (auto& f : g.frames) total += f.lumen.points.size();
    double ratio = (double)sample_size / (double)total;
    std::vector<ContourPoint> all;
    for (auto& f : g.frames) {
        size_t fs = f64_as_usize(std::ceil((double)f.lumen.points.size() * ratio));
        auto s = downsample_contour_points(f.lumen.points, std::max<size_t>(fs, 1));
        all.insert(all.end(), s.begin(), s.end());
    }
    return all;
}
// :95-145
inline void rotate_geometry_around_point(Geometry& g, double angle, double cx, double cy) {
    double c = std::cos(angle), s = std::sin(angle);
    auto rot = [&](double x, double y) {
        double tx = x - cx, ty = y - cy;
        double rx = tx * c - ty * s, ry = tx * s + ty * c;
        return std::make_pair(rx + cx, ry + cy);
    };
    for (auto& f : g.frames) {
        for (auto& p : f.lumen.points) std::tie(p.x, p.y) = rot(p.x, p.y);
        std::tie(f.cx, f.cy) = rot(f.cx, f.cy);
        for (auto& kv : f.extras) {
            for (auto& p : kv.second.points) std::tie(p.x, p.y) = rot(p.x, p.y);
            if (kv.second.centroid) {
                auto [ox, oy, oz] = *kv.second.centroid;
                auto r = rot(ox, oy);
                kv.second.centroid = Vec3(r.first, r.second, oz);
            }
        }
        if (f.reference_point) std::tie(f.reference_point->x, f.reference_point->y) = rot(f.reference_point->x, f.reference_point->y);
    }
}
struct BetweenTap {
    std::vector<ContourPoint> ref, target;
    double best = 0.0;
};
// :11-92 — mutates geom_b only; returns clones.
inline GeometryPair align_between_geometries(Geometry& a, Geometry& b, double rot_deg, double step_deg,
                                             size_t sample_size, int threads = 1, BetweenTap* tap = nullptr) {
    if (a.frames.empty() || b.frames.empty()) throw Error("index out of bounds: empty geometry");
    size_t ia = ref_or_proximal(a), ib = ref_or_proximal(b);
    if (ia >= a.frames.size() || ib >= b.frames.size()) throw Error("index out of bounds: reference frame");
    double acx = a.frames[ia].cx, acy = a.frames[ia].cy, acz = a.frames[ia].cz;
    double bcx = b.frames[ib].cx, bcy = b.frames[ib].cy, bcz = b.frames[ib].cz;
    translate_geometry(b, acx - bcx, acy - bcy, acz - bcz);
    auto ta = extract_geometry_points(a, std::max<size_t>(sample_size, 500));
    auto tb = extract_geometry_points(b, std::max<size_t>(sample_size, 500));
    double best = find_best_rotation_between(ta, tb, step_deg, rot_deg, threads);
    if (tap) {
        tap->ref = ta;
        tap->target = tb;
        tap->best = best;
    }
    rotate_geometry_around_point(b, best, acx, acy);
    ia = ref_or_proximal(a);
    ib = ref_or_proximal(b);
    const Frame& fb = b.frames.at(ib);
    const Frame& fa = a.frames.at(ia);
    translate_geometry(b, fa.cx - fb.cx, fa.cy - fb.cy, fa.cz - fb.cz);
    GeometryPair p;
    p.geom_a = a;
    p.geom_b = b;
    p.label = a.label + " - " + b.label;
    return p;
}

// ---- processing/postprocessing.rs ---------------------------------------------------
// :100-114
inline double get_avg_z_diff(const Geometry& g) {
    if (g.frames.size() < 2) return 0.0;
    double sum = 0.0;
    for (size_t i = 1; i < g.frames.size(); ++i) sum += g.frames[i].cz - g.frames[i - 1].cz;
    return sum / (double)(g.frames.size() - 1);
}
// Frame::set_value(None, None, None, Some(z)), frame.rs:95-118
inline void frame_set_z(Frame& f, double z) {
    for (auto& p : f.lumen.points) p.z = z;
    if (f.lumen.centroid) std::get<2>(*f.lumen.centroid) = z;
    for (auto& kv : f.extras) {
        for (auto& p : kv.second.points) p.z = z;
        if (kv.second.centroid) std::get<2>(*kv.second.centroid) = z;
    }
    if (f.reference_point) f.reference_point->z = z;
    f.cz = z;
}
// :116-140
inline Geometry resample_by_diff(const Geometry& gin, double diff) {
    Geometry g = gin;
    if (!g.frames.empty()) {
        size_t min_idx = 0;  // Iterator::min_by keeps the FIRST minimum
        for (size_t i = 1; i < g.frames.size(); ++i)
            if (g.frames[i].cz < g.frames[min_idx].cz) min_idx = i;
        if (min_idx != 0) std::rotate(g.frames.begin(), g.frames.begin() + min_idx, g.frames.end());
    }
    if (g.frames.empty()) throw Error("index out of bounds: resample_by_diff on an empty geometry");
    double start_z = g.frames[0].cz;
    for (size_t i = 1; i < g.frames.size(); ++i) frame_set_z(g.frames[i], start_z + (double)i * diff);
    return g;
}
// :142-195
inline std::vector<double> predict_z_positions(double ref_z, double start_z, double stop_z, double z_diff) {
    std::vector<double> z;
    if (!std::isfinite(z_diff) || z_diff == 0.0) return z;
    const double eps = 1e-9;
    if (std::fabs(ref_z - start_z) > eps && std::fabs(ref_z - stop_z) > eps) {
        double cur = ref_z;
        while (cur >= start_z - eps) {
            z.push_back(cur);
            cur -= z_diff;
            if (!std::isfinite(cur)) break;
        }
        std::stable_sort(z.begin(), z.end());
        cur = ref_z + z_diff;
        while (cur <= stop_z + eps) {
            z.push_back(cur);
            cur += z_diff;
            if (!std::isfinite(cur)) break;
        }
    } else {
        double cur = start_z;
        if (stop_z >= start_z && z_diff > 0.0) {
            while (cur <= stop_z + eps) {
                z.push_back(cur);
                cur += z_diff;
                if (!std::isfinite(cur)) break;
            }
        } else if (stop_z <= start_z && z_diff < 0.0) {
            while (cur >= stop_z - eps) {
                z.push_back(cur);
                cur += z_diff;
                if (!std::isfinite(cur)) break;
            }
        }
    }
    return z;
}
// :302-340
inline Contour blend_contour(const Contour& c1, const Contour& c2, double t) {
    Contour o;
    size_t n = std::min(c1.points.size(), c2.points.size());
    for (size_t i = 0; i < n; ++i) {
        ContourPoint p = c1.points[i];
        p.x = c1.points[i].x + t * (c2.points[i].x - c1.points[i].x);
        p.y = c1.points[i].y + t * (c2.points[i].y - c1.points[i].y);
        o.points.push_back(p);
    }
    if (c1.centroid && c2.centroid) {
        auto [ax, ay, az] = *c1.centroid;
        auto [bx, by, bz] = *c2.centroid;
        o.centroid = Vec3(ax + t * (bx - ax), ay + t * (by - ay), az + t * (bz - az));
    }
    if (c1.aortic_thickness && c2.aortic_thickness)
        o.aortic_thickness = *c1.aortic_thickness + t * (*c2.aortic_thickness - *c1.aortic_thickness);
    if (c1.pulmonary_thickness && c2.pulmonary_thickness)
        o.pulmonary_thickness = *c1.pulmonary_thickness + t * (*c2.pulmonary_thickness - *c1.pulmonary_thickness);
    o.id = c1.id;
    o.original_frame = c1.original_frame;
    o.kind = c1.kind;
    return o;
}
// :197-300
inline Geometry new_frames_by_sample_rate(const Geometry& g, std::vector<double> z_coords) {
    std::vector<Frame> nf;
    std::stable_sort(z_coords.begin(), z_coords.end());
    if (g.frames.empty()) throw Error("index out of bounds: new_frames_by_sample_rate on an empty geometry");
    double max_z = g.frames.back().cz;
    for (double z : z_coords) {
        if (z > max_z) break;
        const Frame* exact = nullptr;
        for (auto& f : g.frames)
            if (std::fabs(f.cz - z) < 1e-9) {
                exact = &f;
                break;
            }
        if (exact) {
            nf.push_back(*exact);
            continue;
        }
        const Frame *lo = nullptr, *up = nullptr;
        for (size_t i = 0; i + 1 < g.frames.size(); ++i)
            if (g.frames[i].cz <= z && g.frames[i + 1].cz >= z) {
                lo = &g.frames[i];
                up = &g.frames[i + 1];
                break;
            }
        if (!lo) throw Error("Cannot find frames to interpolate between");
        double t = (z - lo->cz) / (up->cz - lo->cz);
        Frame f;
        f.lumen = blend_contour(lo->lumen, up->lumen, t);
        for (ContourType k : {Eem, Calcification, Sidebranch, Catheter, Wall}) {
            auto a = lo->extras.find(k), b = up->extras.find(k);
            if (a != lo->extras.end() && b != up->extras.end()) f.extras[k] = blend_contour(a->second, b->second, t);
        }
        f.id = lo->id;
        f.cx = lo->cx + t * (up->cx - lo->cx);
        f.cy = lo->cy + t * (up->cy - lo->cy);
        f.cz = z;
        nf.push_back(std::move(f));
    }
    std::stable_sort(nf.begin(), nf.end(), [](const Frame& a, const Frame& b) { return a.cz < b.cz; });
    for (size_t i = 0; i < nf.size(); ++i) {
        Frame& f = nf[i];
        f.id = (uint32_t)i;
        f.lumen.id = (uint32_t)i;
        for (auto& p : f.lumen.points) p.z = f.cz;
        if (f.lumen.centroid) std::get<2>(*f.lumen.centroid) = f.cz;
        for (auto& kv : f.extras) {
            kv.second.id = (uint32_t)i;
            for (auto& p : kv.second.points) p.z = f.cz;
        }
        if (f.reference_point) f.reference_point->z = f.cz;
    }
    Geometry o;
    o.frames = std::move(nf);
    o.label = g.label;
    return o;
}
// :342-409
inline GeometryPair trim_geom_pair(const GeometryPair& p) {
    auto trim = [](const Geometry& g, size_t ref, size_t before, size_t after) {
        size_t start = ref - before, end = ref + after;
        std::vector<Frame> fr;
        if (start < end && end <= g.frames.size())
            fr.assign(g.frames.begin() + start, g.frames.begin() + end);
        else
            fr = g.frames;
        for (size_t i = 0; i < fr.size(); ++i) {
            fr[i].id = (uint32_t)i;
            fr[i].lumen.id = (uint32_t)i;
            for (auto& kv : fr[i].extras) kv.second.id = (uint32_t)i;
        }
        Geometry o;
        o.frames = std::move(fr);
        o.label = g.label;
        return o;
    };
    size_t ra = find_ref_frame_idx(p.geom_a).value_or(0), rb = find_ref_frame_idx(p.geom_b).value_or(0);
    if (ra > p.geom_a.frames.size() || rb > p.geom_b.frames.size()) throw Error("attempt to subtract with overflow (trim_geom_pair)");
    size_t before = std::min(ra, rb);
    size_t after = std::min(p.geom_a.frames.size() - ra, p.geom_b.frames.size() - rb);
    GeometryPair o;
    o.geom_a = trim(p.geom_a, ra, before, after);
    o.geom_b = trim(p.geom_b, rb, before, after);
    o.label = p.label;
    return o;
}
// :411-470
inline GeometryPair adjust_walls_anomalous_geom_pair(const GeometryPair& p) {
    std::vector<Frame> fa, fb;
    size_t n = std::min(p.geom_a.frames.size(), p.geom_b.frames.size());
    for (size_t i = 0; i < n; ++i) {
        Frame a = p.geom_a.frames[i], b = p.geom_b.frames[i];
        auto ta = a.lumen.aortic_thickness, tb = b.lumen.aortic_thickness;
        if (ta || tb) {
            double adj = (ta && tb) ? (*ta + *tb) / 2.0 : (ta ? *ta : *tb);
            a.lumen.aortic_thickness = adj;
            b.lumen.aortic_thickness = adj;
        }
        fa.push_back(std::move(a));
        fb.push_back(std::move(b));
    }
    GeometryPair o;
    o.geom_a.frames = create_wall_frames(fa, true);
    o.geom_a.label = p.geom_a.label;
    o.geom_b.frames = create_wall_frames(fb, true);
    o.geom_b.label = p.geom_b.label;
    o.label = p.label;
    return o;
}
// :12-87
inline GeometryPair postprocess_geom_pair(const GeometryPair& p, double tol, bool anomalous) {
    double da = get_avg_z_diff(p.geom_a), db = get_avg_z_diff(p.geom_b);
    bool same = (da - db) < tol;  // sic: signed difference (:93)
    auto ria = find_ref_frame_idx(p.geom_a), rib = find_ref_frame_idx(p.geom_b);
    if (!ria || !rib) throw Error("No reference point found in any frame");
    double ref_z_a = p.geom_a.frames[*ria].cz, ref_z_b = p.geom_b.frames[*rib].cz;
    GeometryPair r;
    r.label = p.label;
    auto span = [](const Geometry& g) {
        double z0 = g.frames.front().cz, zn = g.frames.back().cz;
        return (z0 < zn) ? std::make_pair(z0, zn) : std::make_pair(zn, z0);
    };
    if (same) {
        double mean = (da + db) / 2.0;
        r.geom_a = resample_by_diff(p.geom_a, mean);
        r.geom_b = resample_by_diff(p.geom_b, mean);
    } else if (da < db) {
        auto [start, stop] = span(p.geom_b);
        r.geom_b = new_frames_by_sample_rate(p.geom_b, predict_z_positions(ref_z_b, start, stop, da));
        r.geom_a = resample_by_diff(p.geom_a, da);
    } else {
        auto [start, stop] = span(p.geom_a);
        r.geom_a = new_frames_by_sample_rate(p.geom_a, predict_z_positions(ref_z_a, start, stop, db));
        r.geom_b = resample_by_diff(p.geom_b, db);
    }
    auto ra2 = find_ref_frame_idx(r.geom_a), rb2 = find_ref_frame_idx(r.geom_b);
    if (!ra2 || !rb2) throw Error("No reference point found in any frame");
    if (*ra2 >= p.geom_a.frames.size() || *rb2 >= p.geom_b.frames.size()) throw Error("index out of bounds: postprocess_geom_pair");
    double translation = p.geom_a.frames[*ra2].cz - p.geom_b.frames[*rb2].cz;  // sic: indexes the ORIGINAL pair (:76-77)
    translate_geometry(r.geom_a, 0.0, 0.0, translation);
    GeometryPair t = trim_geom_pair(r);
    return anomalous ? adjust_walls_anomalous_geom_pair(t) : t;
}

// ---- ingest: io/input.rs, io/build.rs, geometry.rs reorder/proximal ------------
inline std::vector<std::string> split_line(const std::string& line, char delim) {
    std::vector<std::string> out;
    std::string cur;
    for (char ch : line) {
        if (ch == delim) {
            out.push_back(cur);
            cur.clear();
        } else if (ch != '\r')
            cur.push_back(ch);
    }
    out.push_back(cur);
    return out;
}
inline std::string trim(const std::string& s) {
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return (a == std::string::npos) ? "" : s.substr(a, b - a + 1);
}
inline char detect_delimiter(const std::string& path) {  // input.rs:149-170
    std::ifstream f(path);
    if (!f) throw Error("failed to open file for delimiter sniffing: \"" + path + "\"");
    std::string first;
    std::getline(f, first);
    size_t tabs = std::count(first.begin(), first.end(), '\t');
    size_t commas = std::count(first.begin(), first.end(), ',');
    return (tabs > commas) ? '\t' : ',';
}
inline bool parse_u32(const std::string& s, uint32_t& v) {
    if (s.empty()) return false;
    char* end = nullptr;
    errno = 0;
    unsigned long long t = std::strtoull(s.c_str(), &end, 10);
    if (errno || *end != '\0' || s[0] == '-' || s[0] == '+' || t > 0xffffffffull) return false;
    v = (uint32_t)t;
    return true;
}
inline bool parse_f64(const std::string& s, double& v) {
    if (s.empty()) return false;
    char* end = nullptr;
    v = std::strtod(s.c_str(), &end);  // correctly rounded, like Rust's str::parse::<f64>
    return *end == '\0';
}
inline bool parse_point(const std::vector<std::string>& f, ContourPoint& p) {
    if (f.size() < 4 || f.size() > 5) return false;
    if (!parse_u32(trim(f[0]), p.frame_index)) return false;
    if (!parse_f64(trim(f[1]), p.x) || !parse_f64(trim(f[2]), p.y) || !parse_f64(trim(f[3]), p.z)) return false;
    p.point_index = 0;
    p.aortic = false;
    if (f.size() == 5) {
        std::string b = trim(f[4]);
        if (b == "true")
            p.aortic = true;
        else if (b == "false")
            p.aortic = false;
        else
            return false;
    }
    return true;
}
// input.rs:172-194 (headerless; invalid rows are skipped; the csv crate also
// rejects rows whose field count differs from the first row's)
inline std::vector<ContourPoint> read_contour_data(const std::string& path) {
    char delim = detect_delimiter(path);
    std::ifstream f(path);
    std::vector<ContourPoint> pts;
    std::string line;
    size_t nfields = 0;
    while (std::getline(f, line)) {
        if (trim(line).empty()) continue;
        auto fields = split_line(line, delim);
        if (nfields == 0) nfields = fields.size();
        if (fields.size() != nfields) continue;
        ContourPoint p;
        if (parse_point(fields, p)) pts.push_back(p);
    }
    return pts;
}
inline ContourPoint read_reference_point(const std::string& path) {  // input.rs:214-235
    char delim = detect_delimiter(path);
    std::ifstream f(path);
    std::string line;
    while (std::getline(f, line)) {
        if (trim(line).empty()) continue;
        ContourPoint p;
        if (!parse_point(split_line(line, delim), p)) throw Error("failed to deserialize first reference-point record");
        return p;
    }
    throw Error("reference-point file \"" + path + "\" was empty — this data is required");
}
inline std::vector<Record> read_records(const std::string& path) {  // input.rs:237-251
    char delim = detect_delimiter(path);
    std::ifstream f(path);
    std::string line;
    if (!std::getline(f, line)) return {};
    auto hdr = split_line(line, delim);
    int c_frame = -1, c_phase = -1, c_m1 = -1, c_m2 = -1;
    for (size_t i = 0; i < hdr.size(); ++i) {
        std::string h = trim(hdr[i]);
        if (h == "frame") c_frame = (int)i;
        if (h == "phase") c_phase = (int)i;
        if (h == "measurement_1") c_m1 = (int)i;
        if (h == "measurement_2") c_m2 = (int)i;
    }
    if (c_frame < 0 || c_phase < 0 || c_m1 < 0 || c_m2 < 0) throw Error("reading " + path + ": missing field");
    std::vector<Record> recs;
    while (std::getline(f, line)) {
        if (trim(line).empty()) continue;
        auto fl = split_line(line, delim);
        if (fl.size() != hdr.size()) throw Error("reading " + path + ": unequal record length");
        Record r;
        if (!parse_u32(trim(fl[c_frame]), r.frame)) throw Error("reading " + path + ": invalid frame");
        r.phase = fl[c_phase];
        double v;
        if (parse_f64(trim(fl[c_m1]), v)) r.measurement_1 = v;  // csv::invalid_option
        if (parse_f64(trim(fl[c_m2]), v)) r.measurement_2 = v;
        recs.push_back(r);
    }
    return recs;
}
inline bool file_exists(const std::string& p) {
    std::ifstream f(p);
    return (bool)f;
}
// input.rs:62-147 with the default names map of build.rs:20-27
inline InputData process_directory(const std::string& dir, bool diastole, const std::string& label) {
    InputData in;
    in.diastole = diastole;
    in.label = label;
    std::string phase = diastole ? "diastolic" : "systolic";
    std::string cp = dir + "/" + phase + "_contours.csv";
    if (!file_exists(cp)) throw Error("required contours file missing: \"" + cp + "\"");
    in.lumen = read_contour_data(cp);
    std::string rp = dir + "/" + phase + "_reference_points.csv";
    if (!file_exists(rp)) throw Error("required reference-point file missing: \"" + rp + "\"");
    in.ref_point = read_reference_point(rp);
    auto opt = [&](const std::string& prefix) -> std::optional<std::vector<ContourPoint>> {
        std::string p = dir + "/" + prefix + "_" + phase + "_contours.csv";
        if (!file_exists(p)) return std::nullopt;
        return read_contour_data(p);
    };
    in.sidebranch = opt("branch");
    in.calcification = opt("calcium");
    in.eem = opt("eem");
    std::string rec = dir + "/combined_sorted_manual.csv";
    if (!file_exists(rec)) rec = dir + "/diastolic_systolic_records.csv";
    if (file_exists(rec)) in.record = read_records(rec);
    return in;
}

// contour.rs:158-211
inline std::vector<Contour> build_contour_with_mapping(const std::vector<ContourPoint>& points,
                                                       const std::optional<std::vector<Record>>& records,
                                                       ContourType kind, const std::map<uint32_t, uint32_t>& mapping) {
    std::map<uint32_t, std::vector<ContourPoint>> groups;  // sorted by frame idx == sort_by_key
    for (auto& p : points) groups[p.frame_index].push_back(p);
    std::map<uint32_t, std::pair<std::optional<double>, std::optional<double>>> meas;
    if (kind == Lumen && records)
        for (auto& r : *records) meas[r.frame] = {r.measurement_1, r.measurement_2};
    std::vector<Contour> out;
    for (auto& kv : groups) {
        auto m = mapping.find(kv.first);
        if (m == mapping.end()) throw Error("No mapping found for original frame " + std::to_string(kv.first));
        Contour c;
        c.id = m->second;
        c.original_frame = kv.first;
        c.points = kv.second;
        c.kind = kind;
        if (kind == Lumen) {
            auto q = meas.find(kv.first);
            if (q != meas.end()) {
                c.aortic_thickness = q->second.first;
                c.pulmonary_thickness = q->second.second;
            }
        }
        out.push_back(std::move(c));
    }
    return out;
}
// frame.rs:163-204
inline std::vector<ContourPoint> create_catheter_points(const std::vector<ContourPoint>& points, double icx, double icy,
                                                        double radius, uint32_t n_points) {
    std::map<uint32_t, double> frame_z;
    for (auto& p : points) frame_z.emplace(p.frame_index, p.z);  // first z seen per frame
    std::vector<ContourPoint> out;
    for (auto& kv : frame_z)
        for (uint32_t i = 0; i < n_points; ++i) {
            double angle = 2.0 * PI * (double)i / (double)n_points;
            ContourPoint p;
            p.frame_index = kv.first;
            p.point_index = i;
            p.x = icx + radius * std::cos(angle);
            p.y = icy + radius * std::sin(angle);
            p.z = kv.second;
            out.push_back(p);
        }
    return out;
}
// geometry.rs:72-155
inline void reorder_frames(Geometry& g, const std::vector<Record>& records, bool diastole) {
    std::string phase = diastole ? "D" : "S";
    std::vector<uint32_t> filtered;
    for (auto& r : records)
        if (r.phase == phase) filtered.push_back(r.frame);
    std::map<uint32_t, double> orig_z;
    for (auto& f : g.frames)
        if (!f.lumen.points.empty()) orig_z.emplace(f.lumen.original_frame, f.lumen.points.front().z);
    std::map<uint32_t, Frame> fmap;
    for (auto& f : g.frames) fmap[f.lumen.original_frame] = f;
    std::vector<Frame> nf;
    for (uint32_t id : filtered) {
        auto it = fmap.find(id);
        if (it != fmap.end()) {
            nf.push_back(std::move(it->second));
            fmap.erase(it);
        }
    }
    for (auto& kv : fmap) nf.push_back(std::move(kv.second));  // sorted by original_frame
    for (size_t i = 0; i < nf.size(); ++i) {
        Frame& f = nf[i];
        uint32_t nid = (uint32_t)i;
        auto zi = orig_z.find(f.lumen.original_frame);
        double z = (zi != orig_z.end()) ? zi->second : (double)nid;
        f.id = nid;
        f.lumen.id = nid;
        for (auto& p : f.lumen.points) {
            p.frame_index = nid;
            p.z = z;
        }
        if (f.lumen.centroid) std::get<2>(*f.lumen.centroid) = z;
        for (auto& kv : f.extras) {
            kv.second.id = nid;
            for (auto& p : kv.second.points) {
                p.frame_index = nid;
                p.z = z;
            }
            if (kv.second.centroid) std::get<2>(*kv.second.centroid) = z;
        }
        if (f.reference_point) f.reference_point->z = z;
        f.cz = z;
    }
    g.frames = std::move(nf);
}
// geometry.rs:325-381
inline void ensure_proximal_at_position_zero(Geometry& g) {
    size_t n = g.frames.size();
    if (n == 0) return;
    size_t prox = std::min(find_proximal_end_idx(g), n - 1);
    if (prox != 0) std::reverse(g.frames.begin(), g.frames.end());
    std::vector<double> zs;
    for (auto& f : g.frames) zs.push_back(f.cz);
    std::stable_sort(zs.begin(), zs.end());
    uint32_t next_id = 0;
    for (size_t i = 0; i < n; ++i) {
        Frame& f = g.frames[i];
        f.id = (uint32_t)i;
        double z = zs[i];
        f.cz = z;
        f.lumen.id = next_id++;
        for (auto& p : f.lumen.points) p.z = z;
        if (f.lumen.centroid) std::get<2>(*f.lumen.centroid) = z;
        for (auto& kv : f.extras) {
            kv.second.id = next_id++;
            for (auto& p : kv.second.points) p.z = z;
            if (kv.second.centroid) std::get<2>(*kv.second.centroid) = z;
        }
        if (f.reference_point) f.reference_point->z = z;
    }
}
// io/integrity_check.rs:8-247
inline void check_geometry_integrity(const Geometry& g) {
    if (g.frames.empty()) throw Error("Geometry has no frames");
    for (size_t i = 0; i < g.frames.size(); ++i)
        if (g.frames[i].id != i)
            throw Error("Frame IDs are not consecutive. Expected ID " + std::to_string(i) + ", found ID " +
                        std::to_string(g.frames[i].id));
    auto approx = [](double a, double b) { return std::fabs(a - b) < 1e-6; };
    for (auto& f : g.frames) {
        double lx, ly, lz;
        if (f.lumen.centroid)
            std::tie(lx, ly, lz) = *f.lumen.centroid;
        else {
            Contour t = f.lumen;
            compute_centroid(t);
            std::tie(lx, ly, lz) = t.centroid.value_or(Vec3(0, 0, 0));
        }
        if (!(approx(f.cx, lx) && approx(f.cy, ly) && approx(f.cz, lz)))
            throw Error("Frame centroid does not match lumen centroid in frame " + std::to_string(f.id));
    }
    for (auto& f : g.frames) {
        if (f.lumen.points.empty()) throw Error("Lumen contour has no points in frame " + std::to_string(f.id));
        if (f.lumen.kind != Lumen) throw Error("Lumen contour has incorrect type in frame " + std::to_string(f.id));
    }
    size_t nref = 0;
    for (auto& f : g.frames) nref += f.reference_point ? 1 : 0;
    if (nref != 1) throw Error("Expected exactly one reference point, found " + std::to_string(nref));
    std::map<ContourType, size_t> expected;
    for (auto& f : g.frames) {
        auto chk = [&](ContourType k, size_t cnt) {
            auto it = expected.find(k);
            if (it == expected.end())
                expected[k] = cnt;
            else if (it->second != cnt)
                throw Error("contour point count mismatch in frame " + std::to_string(f.id) + ". Expected " +
                            std::to_string(it->second) + ", found " + std::to_string(cnt));
        };
        chk(Lumen, f.lumen.points.size());
        for (auto& kv : f.extras) chk(kv.second.kind, kv.second.points.size());
    }
    for (auto& f : g.frames) {
        for (auto& kv : f.extras)
            if (kv.second.original_frame != f.lumen.original_frame)
                throw Error("Original frame mismatch in frame " + std::to_string(f.id));
        if (f.reference_point && f.reference_point->frame_index != f.lumen.original_frame)
            throw Error("Reference point original frame mismatch in frame " + std::to_string(f.id));
    }
    size_t prox = find_proximal_end_idx(g), min_idx = 0;
    double min_z = std::numeric_limits<double>::infinity();
    for (size_t i = 0; i < g.frames.size(); ++i)
        if (g.frames[i].cz < min_z) {
            min_z = g.frames[i].cz;
            min_idx = i;
        }
    if (prox != min_idx)
        throw Error("Proximal end index is " + std::to_string(prox) + ", but frame with minimum z is " +
                    std::to_string(min_idx));
    if (g.frames.front().cz > g.frames.back().cz) throw Error("First frame has higher z-coords than last frame");
}
// io/build.rs:9-205
inline Geometry build_geometry_from_inputdata(const InputData& in, const std::string& label, bool diastole, double icx,
                                              double icy, double radius, uint32_t n_points) {
    std::set<uint32_t> all_frames;
    for (auto& p : in.lumen) all_frames.insert(p.frame_index);
    for (auto* o : {&in.eem, &in.calcification, &in.sidebranch})
        if (*o)
            for (auto& p : **o) all_frames.insert(p.frame_index);
    all_frames.insert(in.ref_point.frame_index);
    std::map<uint32_t, uint32_t> mapping;
    uint32_t k = 0;
    for (uint32_t f : all_frames) mapping[f] = k++;

    auto lumen = build_contour_with_mapping(in.lumen, in.record, Lumen, mapping);
    std::map<uint32_t, Frame> fmap;
    for (auto& c : lumen) {
        compute_centroid(c);
        Frame f;
        f.id = c.id;
        std::tie(f.cx, f.cy, f.cz) = c.centroid.value_or(Vec3(0, 0, 0));
        f.lumen = c;
        auto m = mapping.find(in.ref_point.frame_index);
        if (m != mapping.end() && m->second == f.id) f.reference_point = in.ref_point;
        fmap[f.id] = std::move(f);
    }
    auto attach = [&](const std::optional<std::vector<ContourPoint>>& pts, ContourType kind) {
        if (!pts) return;
        for (auto& c : build_contour_with_mapping(*pts, std::nullopt, kind, mapping)) {
            Contour cc = c;
            compute_centroid(cc);
            auto it = fmap.find(cc.id);
            if (it != fmap.end()) it->second.extras[kind] = std::move(cc);
        }
    };
    attach(in.eem, Eem);
    attach(in.calcification, Calcification);
    attach(in.sidebranch, Sidebranch);
    if (n_points > 0) {
        std::vector<ContourPoint> all;
        for (auto& kv : fmap) all.insert(all.end(), kv.second.lumen.points.begin(), kv.second.lumen.points.end());
        attach(create_catheter_points(all, icx, icy, radius, n_points), Catheter);
    }
    Geometry g;
    g.label = label;
    for (auto& kv : fmap) g.frames.push_back(std::move(kv.second));  // sorted by id
    if (in.record) reorder_frames(g, *in.record, diastole);
    for (auto& f : g.frames) sort_frame_points(f);
    ensure_proximal_at_position_zero(g);
    for (auto& f : g.frames) {  // Frame::set_value(Some(id), ..) — frame.rs:76-82
        f.lumen.id = f.id;
        for (auto& kv : f.extras) kv.second.id = f.id;
    }
    check_geometry_integrity(g);
    return g;
}

// ---- orchestration: binding/entry.rs ---------------------------------------------
struct ProcessParams {
    double step_deg = 0.5, range_deg = 90.0;
    size_t sample_size = 500;
    bool smooth = true, bruteforce = false, postprocessing = false;
    int threads = 1;
};
constexpr double TOLERANCE = 0.03;  // entry.rs:21
struct FullResult {
    GeometryPair ab, cd, ac, bd;
    std::vector<AlignLog> logs[4];
};
// entry.rs:122-277 (full) / :412-530 (double pair: stops after AB, CD)
inline FullResult full_processing(std::vector<Geometry> geoms, const ProcessParams& p, bool double_pair) {
    if (geoms.size() != 4) throw Error("Full processing requires exactly 4 geometries, got " + std::to_string(geoms.size()));
    FullResult r;
    Geometry g[4];
    bool anomalous = false;
    for (int i = 0; i < 4; ++i) {
        auto w = align_frames_in_geometry(geoms[i], p.step_deg, p.range_deg, p.smooth, p.bruteforce, p.sample_size, p.threads);
        g[i] = std::move(w.geometry);
        r.logs[i] = std::move(w.logs);
        anomalous = anomalous || w.anomalous;
    }
    r.ab = align_between_geometries(g[0], g[1], p.range_deg, p.step_deg, p.sample_size, p.threads);
    r.cd = align_between_geometries(g[2], g[3], p.range_deg, p.step_deg, p.sample_size, p.threads);
    if (!double_pair) {
        r.ac = align_between_geometries(g[0], g[2], p.range_deg, p.step_deg, p.sample_size, p.threads);
        r.bd = align_between_geometries(g[1], g[3], p.range_deg, p.step_deg, p.sample_size, p.threads);
    }
    if (p.postprocessing) {  // maybe_postprocess, entry.rs:56-69, :279-289
        r.ab = postprocess_geom_pair(r.ab, TOLERANCE, anomalous);
        r.cd = postprocess_geom_pair(r.cd, TOLERANCE, anomalous);
        if (!double_pair) {
            r.ac = postprocess_geom_pair(r.ac, TOLERANCE, anomalous);
            r.bd = postprocess_geom_pair(r.bd, TOLERANCE, anomalous);
        }
    }
    return r;
}
struct PairResult {
    GeometryPair pair;
    std::vector<AlignLog> logs[2];
};
// entry.rs:617-666
inline PairResult pair_processing(std::vector<Geometry> geoms, const ProcessParams& p) {
    if (geoms.size() != 2) throw Error("Single Pair processing requires exactly 2 geometries, got " + std::to_string(geoms.size()));
    PairResult r;
    Geometry g[2];
    bool anomalous = false;
    for (int i = 0; i < 2; ++i) {
        auto w = align_frames_in_geometry(geoms[i], p.step_deg, p.range_deg, p.smooth, p.bruteforce, p.sample_size, p.threads);
        g[i] = std::move(w.geometry);
        r.logs[i] = std::move(w.logs);
        anomalous = anomalous || w.anomalous;
    }
    r.pair = align_between_geometries(g[0], g[1], p.range_deg, p.step_deg, p.sample_size, p.threads);
    if (p.postprocessing) r.pair = postprocess_geom_pair(r.pair, TOLERANCE, anomalous);  // entry.rs:668-671
    return r;
}

// ---- flat "geometry blob" codec (f64 stream; layout documented in include/mmrs_b200.h) ----
inline void put_point(std::vector<double>& o, const ContourPoint& p) {
    o.insert(o.end(), {(double)p.frame_index, (double)p.point_index, p.x, p.y, p.z, p.aortic ? 1.0 : 0.0});
}
inline void put_contour(std::vector<double>& o, const Contour& c) {
    double cx = 0, cy = 0, cz = 0;
    if (c.centroid) std::tie(cx, cy, cz) = *c.centroid;
    o.insert(o.end(), {(double)c.kind, (double)c.id, (double)c.original_frame, c.centroid ? 1.0 : 0.0, cx, cy, cz,
                       c.aortic_thickness ? 1.0 : 0.0, c.aortic_thickness.value_or(0.0),
                       c.pulmonary_thickness ? 1.0 : 0.0, c.pulmonary_thickness.value_or(0.0),
                       (double)c.points.size()});
    for (auto& p : c.points) put_point(o, p);
}
inline std::vector<double> encode_geometry(const Geometry& g) {
    std::vector<double> o;
    o.push_back((double)g.frames.size());
    for (auto& f : g.frames) {
        o.insert(o.end(), {(double)f.id, f.cx, f.cy, f.cz, f.reference_point ? 1.0 : 0.0});
        put_point(o, f.reference_point.value_or(ContourPoint{}));
        o.push_back((double)(1 + f.extras.size()));
        put_contour(o, f.lumen);
        for (auto& kv : f.extras) put_contour(o, kv.second);
    }
    return o;
}
struct Cursor {
    const double* p;
    const double* end;
    double next() {
        if (p >= end) throw Error("geometry blob truncated");
        return *p++;
    }
};
inline ContourPoint get_point(Cursor& c) {
    ContourPoint p;
    p.frame_index = (uint32_t)c.next();
    p.point_index = (uint32_t)c.next();
    p.x = c.next();
    p.y = c.next();
    p.z = c.next();
    p.aortic = c.next() != 0.0;
    return p;
}
inline Contour get_contour(Cursor& c) {
    Contour o;
    o.kind = (ContourType)(int)c.next();
    o.id = (uint32_t)c.next();
    o.original_frame = (uint32_t)c.next();
    bool hc = c.next() != 0.0;
    double x = c.next(), y = c.next(), z = c.next();
    if (hc) o.centroid = Vec3(x, y, z);
    bool ha = c.next() != 0.0;
    double a = c.next();
    if (ha) o.aortic_thickness = a;
    bool hp = c.next() != 0.0;
    double pv = c.next();
    if (hp) o.pulmonary_thickness = pv;
    size_t n = (size_t)c.next();
    o.points.reserve(n);
    for (size_t i = 0; i < n; ++i) o.points.push_back(get_point(c));
    return o;
}
inline Geometry decode_geometry(const double* data, size_t len, const std::string& label) {
    Cursor c{data, data + len};
    Geometry g;
    g.label = label;
    size_t nf = (size_t)c.next();
    for (size_t i = 0; i < nf; ++i) {
        Frame f;
        f.id = (uint32_t)c.next();
        f.cx = c.next();
        f.cy = c.next();
        f.cz = c.next();
        bool hr = c.next() != 0.0;
        ContourPoint rp = get_point(c);
        if (hr) f.reference_point = rp;
        size_t nc = (size_t)c.next();
        for (size_t k = 0; k < nc; ++k) {
            Contour ct = get_contour(c);
            if (k == 0)
                f.lumen = std::move(ct);
            else
                f.extras[ct.kind] = std::move(ct);
        }
        g.frames.push_back(std::move(f));
    }
    return g;
}

}  // namespace ora
