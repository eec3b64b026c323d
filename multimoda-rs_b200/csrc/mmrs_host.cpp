// =============================================================================
// mmrs_host.cpp — host side of the drop-in: geometry model, ingest, the frame
// chain, post steps, inter-pullback alignment and the mode orchestration of the
// reference, re-designed around ONE idea: every rotation search of every
// pullback of every case in a call is a unit of a BATCHED GPU sweep, one launch
// sequence per search stage (include/mmrs_b200.h: mmrs_sweep_batched).
//
// The reference walks frames serially (align_within.rs:72-134): frame i is
// matched against the already aligned frame i-1. Hausdorff distance is invariant
// under a common rigid motion and all rotations share one centre, so the cost
// curve of frame pair (i-1, i) does not depend on the chain: it is
// H(P[i-1] - c[i-1], R(theta) (P[i] - c[i])). The sweep therefore runs on
// "decoupled" units built from the ORIGINAL frames, all at once. Rounding makes
// the decoupled f64 costs differ from the chain's by ~1e-14 relative, so a unit's
// arg-min is accepted only when no other rechecked candidate lies within
// kTieMargin of it ("certified"); otherwise the search of that frame is redone
// on the chain's own points, in order, with the exact (tie_margin = 0) sweep —
// which is the reference's arithmetic literally. The chain itself (rigid motions
// of every contour, logs, cumulative angle) is replayed on the host in f64.
//
// Data model: structure-of-arrays contours (x[], y[], z[], ...) so that sample
// extraction into sweep batches is a strided gather; this is a different layout
// from the reference's Vec<ContourPoint> (src/types/native/contour_point.rs:55-67)
// but carries the same fields, and the blob codec maps one to the other.
// =============================================================================
#include <algorithm>
#include <array>
#include <atomic>
#include <cctype>
#include <charconv>
#include <exception>
#include <condition_variable>
#include <deque>
#include <type_traits>
#include <future>
#include <mutex>
#include <thread>
#include <chrono>
#include <cstdio>
#include <cerrno>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
#include <map>
#include <numeric>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/types.h>

#include "mmrs_internal.hpp"
#include "mmrs_pool.hpp"

namespace {

using std::size_t;
constexpr double kPi = 3.14159265358979323846264338327950288;
// Certification margin (multiplied by max(1, Rmax) on the device): 5 orders of
// magnitude above the worst-case decoupled-vs-chain rounding gap.
constexpr double kTieMargin = 1e-9;

struct InputErr : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// MMRS_TRACE=1: phase timings of mmrs_process_cases on stderr.
struct Trace {
    bool on = std::getenv("MMRS_TRACE") != nullptr;
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void lap(const char* what) {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[mmrs] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t).count());
        t = now;
    }
};

using mmrs::host_thread_budget;
using mmrs::parallel_for;

inline double rad2deg(double r) { return r * (180.0 / kPi); }
inline double rem_euclid(double a, double b) {
    double r = std::fmod(a, b);
    return (r < 0.0) ? r + std::fabs(b) : r;
}
inline size_t as_usize(double v) {
    if (!(v == v) || v <= 0.0) return 0;
    if (v >= 18446744073709551615.0) return std::numeric_limits<size_t>::max();
    return (size_t)v;
}

enum Kind : int { kLumen = 0, kEem = 1, kCalc = 2, kSide = 3, kCatheter = 4, kWall = 5 };

// ---- SoA contour ------------------------------------------------------------------
struct Contour {
    int kind = kLumen;
    uint32_t id = 0, original_frame = 0;
    std::vector<uint32_t> fi, pi;  // frame_index, point_index
    std::vector<double> x, y, z;
    std::vector<uint8_t> ao;  // aortic
    bool has_c = false;
    double c[3] = {0, 0, 0};
    bool has_at = false, has_pt = false;
    double at = 0, pt = 0;  // aortic / pulmonary thickness

    size_t size() const { return x.size(); }
    void push(uint32_t f, uint32_t p, double px, double py, double pz, bool a) {
        fi.push_back(f);
        pi.push_back(p);
        x.push_back(px);
        y.push_back(py);
        z.push_back(pz);
        ao.push_back(a ? 1 : 0);
    }
    void resize(size_t n) {
        fi.resize(n);
        pi.resize(n);
        x.resize(n);
        y.resize(n);
        z.resize(n);
        ao.resize(n);
    }
    // Contour::compute_centroid, contour.rs:213-224 (sequential left fold)
    void centroid() {
        if (x.empty()) {
            has_c = false;
            return;
        }
        double sx = 0.0, sy = 0.0, sz = 0.0;
        for (size_t i = 0; i < x.size(); ++i) {
            sx = sx + x[i];
            sy = sy + y[i];
            sz = sz + z[i];
        }
        const double n = (double)x.size();
        c[0] = sx / n, c[1] = sy / n, c[2] = sz / n;
        has_c = true;
    }
    void shift(double dx, double dy, double dz) {  // contour_point.rs:29-36
        for (size_t i = 0; i < x.size(); ++i) {
            x[i] = x[i] + dx;
            y[i] = y[i] + dy;
            z[i] = z[i] + dz;
        }
    }
    void spin(double angle, double cx, double cy) {  // contour_point.rs:38-52, cos/sin hoisted (same values)
        if (angle == 0.0) return;
        const double ca = std::cos(angle), sa = std::sin(angle);
        for (size_t i = 0; i < x.size(); ++i) {
            const double px = x[i] - cx, py = y[i] - cy;
            x[i] = px * ca - py * sa + cx;
            y[i] = px * sa + py * ca + cy;
        }
    }
    void permute(const std::vector<size_t>& order) {
        auto apply = [&](auto& v) {
            auto t = v;
            for (size_t i = 0; i < order.size(); ++i) v[i] = t[order[i]];
        };
        apply(fi), apply(pi), apply(x), apply(y), apply(z), apply(ao);
    }
    // Contour::sort_contour_points, contour.rs:368-405: stable ascending atan2 about the
    // mean, the LAST highest-y point rotated to the front, point_index = position.
    void sort_ccw() {
        const size_t n = x.size();
        if (n == 0) return;
        double sx = 0.0, sy = 0.0;
        for (size_t i = 0; i < n; ++i) {
            sx = sx + x[i];
            sy = sy + y[i];
        }
        const double mx = sx / (double)n, my = sy / (double)n;
        std::vector<std::pair<double, size_t>> keyed(n);
        for (size_t i = 0; i < n; ++i) keyed[i] = {std::atan2(y[i] - my, x[i] - mx), i};
        std::stable_sort(keyed.begin(), keyed.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
        std::vector<size_t> order(n);
        for (size_t i = 0; i < n; ++i) order[i] = keyed[i].second;
        size_t start = 0;
        for (size_t i = 1; i < n; ++i)
            if (!(y[order[i]] < y[order[start]])) start = i;
        std::rotate(order.begin(), order.begin() + start, order.end());
        permute(order);
        for (size_t i = 0; i < n; ++i) pi[i] = (uint32_t)i;
    }
};

struct RefPoint {
    uint32_t fi = 0, pi = 0;
    double x = 0, y = 0, z = 0;
    bool ao = false;
};

struct Frame {
    uint32_t id = 0;
    double c[3] = {0, 0, 0};
    Contour lumen;
    std::map<int, Contour> extras;
    bool has_ref = false;
    RefPoint ref;

    Contour* extra(int k) {
        auto it = extras.find(k);
        return it == extras.end() ? nullptr : &it->second;
    }
    const Contour* extra(int k) const {
        auto it = extras.find(k);
        return it == extras.end() ? nullptr : &it->second;
    }
    void shift(double dx, double dy, double dz) {  // Frame::translate, frame.rs:18-38
        lumen.shift(dx, dy, dz);
        lumen.centroid();
        for (auto& kv : extras) {
            kv.second.shift(dx, dy, dz);
            kv.second.centroid();
        }
        if (has_ref) {
            ref.x = ref.x + dx;
            ref.y = ref.y + dy;
            ref.z = ref.z + dz;
        }
        c[0] += dx, c[1] += dy, c[2] += dz;
    }
    void spin(double angle, double cx, double cy) {  // Frame::rotate, frame.rs:40-63
        if (angle == 0.0) return;
        lumen.spin(angle, cx, cy);
        for (auto& kv : extras) kv.second.spin(angle, cx, cy);
        const double ca = std::cos(angle), sa = std::sin(angle);
        if (has_ref) {
            const double px = ref.x - cx, py = ref.y - cy;
            ref.x = px * ca - py * sa + cx;
            ref.y = px * sa + py * ca + cy;
        }
        const double px = c[0] - cx, py = c[1] - cy;
        c[0] = px * ca - py * sa + cx;
        c[1] = px * sa + py * ca + cy;
    }
    void sort_points() {
        lumen.sort_ccw();
        for (auto& kv : extras) kv.second.sort_ccw();
    }
};

struct Geometry {
    std::vector<Frame> frames;
    std::string label;
    size_t proximal_idx() const {  // geometry.rs:42-60
        const size_t n = frames.size();
        if (n == 0) return 0;
        if (n == 1) return frames[0].lumen.id;
        return frames[0].lumen.original_frame > frames[n - 1].lumen.original_frame ? frames[0].lumen.id
                                                                                  : frames[n - 1].lumen.id;
    }
    size_t ref_or_proximal() const {  // find_ref_frame_idx().unwrap_or(find_proximal_end_idx())
        for (auto& f : frames)
            if (f.has_ref) return f.id;
        return proximal_idx();
    }
    void shift_all(double dx, double dy, double dz) {  // translate_geometry, geometry.rs:278-283
        for (auto& f : frames) f.shift(dx, dy, dz);
    }
    void renumber() {  // tail of insert_frame, geometry.rs:299-322
        for (size_t i = 0; i < frames.size(); ++i) {
            Frame& f = frames[i];
            const uint32_t id = (uint32_t)i;
            f.id = id;
            f.lumen.id = id;
            std::fill(f.lumen.fi.begin(), f.lumen.fi.end(), id);
            for (auto& kv : f.extras) {
                kv.second.id = id;
                std::fill(kv.second.fi.begin(), kv.second.fi.end(), id);
            }
            if (f.has_ref) f.ref.fi = id;
        }
    }
};

// ---- blob codec (layout: include/mmrs_b200.h "geometry blob") -----------------------
struct Reader {
    const double* p;
    const double* e;
    double get() {
        if (p >= e) throw InputErr("geometry blob truncated");
        return *p++;
    }
};
Contour read_contour(Reader& r) {
    Contour c;
    c.kind = (int)r.get();
    c.id = (uint32_t)r.get();
    c.original_frame = (uint32_t)r.get();
    c.has_c = r.get() != 0.0;
    c.c[0] = r.get(), c.c[1] = r.get(), c.c[2] = r.get();
    c.has_at = r.get() != 0.0;
    c.at = r.get();
    c.has_pt = r.get() != 0.0;
    c.pt = r.get();
    const size_t n = as_usize(r.get());   // counts inside a blob are not trusted: NaN / negative / huge values are errors
    if (n > (size_t)(r.e - r.p) / 6) throw InputErr("geometry blob truncated");
    c.resize(n);
    for (size_t i = 0; i < n; ++i) {
        c.fi[i] = (uint32_t)r.p[0], c.pi[i] = (uint32_t)r.p[1];
        c.x[i] = r.p[2], c.y[i] = r.p[3], c.z[i] = r.p[4];
        c.ao[i] = r.p[5] != 0.0;
        r.p += 6;
    }
    return c;
}
Frame read_frame(Reader& r) {
    Frame f;
    f.id = (uint32_t)r.get();
    f.c[0] = r.get(), f.c[1] = r.get(), f.c[2] = r.get();
    f.has_ref = r.get() != 0.0;
    f.ref.fi = (uint32_t)r.get(), f.ref.pi = (uint32_t)r.get();
    f.ref.x = r.get(), f.ref.y = r.get(), f.ref.z = r.get();
    f.ref.ao = r.get() != 0.0;
    const size_t nc = as_usize(r.get());
    if (nc > (size_t)(r.e - r.p) / 12) throw InputErr("geometry blob truncated");   // every contour has a 12-double header
    for (size_t q = 0; q < nc; ++q) {
        Contour c = read_contour(r);
        if (q == 0)
            f.lumen = std::move(c);
        else
            f.extras[c.kind] = std::move(c);
    }
    return f;
}
// Big geometries (OCT-resolution pullbacks are ~100 MB of blob) are decoded / encoded frame-parallel: the work is
// first-touch page faults on fresh memory, which spread over threads. The frame boundaries come from the headers.
constexpr size_t kParallelBlobDoubles = 1u << 17;
constexpr size_t kBlobThreads = 8;
Geometry decode(const double* data, int64_t len) {
    if (!data || len < 1) throw InputErr("geometry blob is empty");
    Reader r{data, data + len};
    Geometry g;
    const size_t nf = as_usize(r.get());
    if (nf > (size_t)len / 24) throw InputErr("geometry blob truncated");  // a frame is at least two 12-double headers
    if ((size_t)len < kParallelBlobDoubles || nf < 2 * kBlobThreads) {
        g.frames.reserve(nf);
        for (size_t k = 0; k < nf; ++k) g.frames.push_back(read_frame(r));
        return g;
    }
    std::vector<const double*> at(nf);
    for (size_t k = 0; k < nf; ++k) {  // skip over frame k using the counts in its headers
        at[k] = r.p;
        if (r.e - r.p < 12) throw InputErr("geometry blob truncated");
        const size_t nc = as_usize(r.p[11]);
        r.p += 12;
        if (nc > (size_t)(r.e - r.p) / 12) throw InputErr("geometry blob truncated");
        for (size_t q = 0; q < nc; ++q) {
            if (r.e - r.p < 12) throw InputErr("geometry blob truncated");
            const size_t n = as_usize(r.p[11]);
            if (n > (size_t)(r.e - r.p - 12) / 6) throw InputErr("geometry blob truncated");
            r.p += 12 + 6 * n;
        }
    }
    g.frames.resize(nf);
    parallel_for(nf, [&](size_t k) {
        Reader rk{at[k], r.e};
        g.frames[k] = read_frame(rk);
    }, kBlobThreads);
    return g;
}
size_t contour_doubles(const Contour& c) { return 12 + 6 * c.size(); }
double* write_contour(double* w, const Contour& c) {
    const double h[12] = {(double)c.kind, (double)c.id,        (double)c.original_frame, c.has_c ? 1.0 : 0.0,
                          c.has_c ? c.c[0] : 0.0, c.has_c ? c.c[1] : 0.0, c.has_c ? c.c[2] : 0.0, c.has_at ? 1.0 : 0.0,
                          c.has_at ? c.at : 0.0, c.has_pt ? 1.0 : 0.0, c.has_pt ? c.pt : 0.0, (double)c.size()};
    std::memcpy(w, h, sizeof h);
    w += 12;
    const size_t n = c.size();
    for (size_t i = 0; i < n; ++i, w += 6) {
        w[0] = (double)c.fi[i];
        w[1] = (double)c.pi[i];
        w[2] = c.x[i];
        w[3] = c.y[i];
        w[4] = c.z[i];
        w[5] = c.ao[i] ? 1.0 : 0.0;
    }
    return w;
}
// Encodes straight into a malloc'ed buffer (what the C ABI hands out; release with mmrs_free).
double* write_frame(double* w, const Frame& f) {
    const double h[12] = {(double)f.id, f.c[0], f.c[1], f.c[2], f.has_ref ? 1.0 : 0.0,
                          f.has_ref ? (double)f.ref.fi : 0.0, f.has_ref ? (double)f.ref.pi : 0.0,
                          f.has_ref ? f.ref.x : 0.0, f.has_ref ? f.ref.y : 0.0, f.has_ref ? f.ref.z : 0.0,
                          (f.has_ref && f.ref.ao) ? 1.0 : 0.0, (double)(1 + f.extras.size())};
    std::memcpy(w, h, sizeof h);
    w += 12;
    w = write_contour(w, f.lumen);
    for (auto& kv : f.extras) w = write_contour(w, kv.second);
    return w;
}
// Result blobs are fresh memory, so writing them is first-touch bound (a 40 MB geometry = 10 000 page faults). Large
// blobs are therefore 2 MB-aligned and marked for transparent huge pages where the kernel offers them (madvise is a hint:
// without THP this is a plain allocation); the caller still releases them with free() / mmrs_free.
void* blob_alloc(size_t bytes) {
    constexpr size_t kHuge = 2u << 20;
    if (bytes >= 4 * kHuge) {
        void* p = nullptr;
        if (posix_memalign(&p, kHuge, (bytes + kHuge - 1) / kHuge * kHuge) == 0 && p) {
            madvise(p, (bytes + kHuge - 1) / kHuge * kHuge, MADV_HUGEPAGE);
            return p;
        }
    }
    return std::malloc(bytes);
}
double* encode_malloc(const Geometry& g, int64_t* len_out) {
    const size_t nf = g.frames.size();
    std::vector<size_t> at(nf + 1);
    at[0] = 1;
    for (size_t k = 0; k < nf; ++k) {
        size_t n = 12 + contour_doubles(g.frames[k].lumen);
        for (auto& kv : g.frames[k].extras) n += contour_doubles(kv.second);
        at[k + 1] = at[k] + n;
    }
    const size_t n = at[nf];
    double* base = (double*)blob_alloc(n * sizeof(double));
    if (!base) throw std::bad_alloc();
    base[0] = (double)nf;
    if (n < kParallelBlobDoubles || nf < 2 * kBlobThreads) {
        for (size_t k = 0; k < nf; ++k) write_frame(base + at[k], g.frames[k]);
    } else {
        parallel_for(nf, [&](size_t k) { write_frame(base + at[k], g.frames[k]); }, kBlobThreads);
    }
    *len_out = (int64_t)n;
    return base;
}
double* to_malloc(const std::vector<double>& v) {
    double* p = (double*)std::malloc(std::max<size_t>(v.size(), 1) * sizeof(double));
    if (!v.empty()) std::memcpy(p, v.data(), v.size() * sizeof(double));
    return p;
}

// =============================================================================
// Ingest (io/input.rs, io/build.rs, geometry.rs reorder / proximal, integrity_check.rs)
// =============================================================================
struct RawPoint {
    uint32_t frame;
    double x, y, z;
    bool aortic;
};
struct Rec {
    uint32_t frame;
    bool diastole_phase, systole_phase;
    bool has1, has2;
    double m1, m2;
};
// Read-only rows of one layer: rows parsed from a file, or the caller's (n, 4) [frame, x, y, z] f64 array as it is
// (mmrs_geometry_from_arrays borrows it: no intermediate copy of a 2-million-point pullback).
struct Points {
    const RawPoint* raw = nullptr;
    const double* rows = nullptr;
    size_t n = 0;
    size_t size() const { return n; }
    uint32_t frame(size_t i) const { return raw ? raw[i].frame : (uint32_t)rows[4 * i]; }
    double x(size_t i) const { return raw ? raw[i].x : rows[4 * i + 1]; }
    double y(size_t i) const { return raw ? raw[i].y : rows[4 * i + 2]; }
    double z(size_t i) const { return raw ? raw[i].z : rows[4 * i + 3]; }
    bool aortic(size_t i) const { return raw ? raw[i].aortic : false; }
};
struct Input {
    std::vector<RawPoint> lumen;
    bool has_eem = false, has_calc = false, has_side = false, has_rec = false;
    std::vector<RawPoint> eem, calc, side;
    Points a_lumen, a_eem, a_calc, a_side;  // set instead of the vectors by the array entry point
    std::vector<Rec> rec;
    RawPoint ref{};
    static Points of(const std::vector<RawPoint>& v, const Points& a) {
        return a.rows ? a : Points{v.data(), nullptr, v.size()};
    }
    Points lumen_pts() const { return of(lumen, a_lumen); }
    Points eem_pts() const { return of(eem, a_eem); }
    Points calc_pts() const { return of(calc, a_calc); }
    Points side_pts() const { return of(side, a_side); }
};

std::string strip(const std::string& s) {
    const size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}
std::vector<std::string> fields_of(const std::string& line, char delim) {
    std::vector<std::string> out(1);
    for (char ch : line) {
        if (ch == delim)
            out.emplace_back();
        else if (ch != '\r')
            out.back().push_back(ch);
    }
    return out;
}
bool to_u32(const std::string& s, uint32_t& v) {
    if (s.empty() || s[0] == '-' || s[0] == '+') return false;
    char* end = nullptr;
    errno = 0;
    const unsigned long long t = std::strtoull(s.c_str(), &end, 10);
    if (errno || *end || t > 0xffffffffull) return false;
    v = (uint32_t)t;
    return true;
}
bool to_f64(const std::string& s, double& v) {
    if (s.empty()) return false;
    char* end = nullptr;
    v = std::strtod(s.c_str(), &end);
    return *end == '\0';
}
bool exists(const std::string& p) { return (bool)std::ifstream(p); }
char sniff(const std::string& path) {  // input.rs:149-170
    std::ifstream f(path);
    if (!f) throw InputErr("failed to open file for delimiter sniffing: \"" + path + "\"");
    std::string first;
    std::getline(f, first);
    return std::count(first.begin(), first.end(), '\t') > std::count(first.begin(), first.end(), ',') ? '\t' : ',';
}
bool row_to_point(const std::vector<std::string>& f, RawPoint& p) {
    if (f.size() < 4 || f.size() > 5) return false;
    if (!to_u32(strip(f[0]), p.frame) || !to_f64(strip(f[1]), p.x) || !to_f64(strip(f[2]), p.y) ||
        !to_f64(strip(f[3]), p.z))
        return false;
    p.aortic = false;
    if (f.size() == 5) {
        const std::string b = strip(f[4]);
        if (b == "true")
            p.aortic = true;
        else if (b != "false")
            return false;
    }
    return true;
}
// Contour files are the bulk of from_file_* wall time (10 000 rows of 16-digit decimals per pullback phase), so they are
// parsed in place: the file is read once, rows and fields are [begin, end) views, numbers go through std::from_chars
// (correctly rounded, like strtod). Anything a view cannot decide exactly like the general path above — a '\r' inside
// a row, a number that is not plain decimal — is handed to that path, so both accept and produce the same rows.
struct View {
    const char *b, *e;
    size_t size() const { return (size_t)(e - b); }
};
inline View strip_view(View v) {
    auto ws = [](char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n'; };
    while (v.b < v.e && ws(*v.b)) ++v.b;
    while (v.e > v.b && ws(v.e[-1])) --v.e;
    return v;
}
inline bool view_to_u32(View v, uint32_t& out) {
    if (v.size() == 0 || v.size() > 10) return v.size() != 0 && to_u32(std::string(v.b, v.e), out);
    uint64_t t = 0;
    for (const char* p = v.b; p < v.e; ++p) {
        if (*p < '0' || *p > '9') return false;
        t = t * 10 + (uint64_t)(*p - '0');
    }
    if (t > 0xffffffffull) return false;
    out = (uint32_t)t;
    return true;
}
inline bool view_to_f64(View v, double& out) {
    if (v.size() == 0) return false;
    bool plain = true;
    for (const char* p = v.b; p < v.e; ++p)
        plain &= (*p >= '0' && *p <= '9') || *p == '.' || *p == '-' || *p == 'e' || *p == 'E';
    if (plain) {
        const auto r = std::from_chars(v.b, v.e, out);
        if (r.ec == std::errc() && r.ptr == v.e) return true;
    }
    return to_f64(std::string(v.b, v.e), out);  // "+1.5", "1e+3", "inf", out-of-range ...: strtod decides, as before
}
std::vector<RawPoint> read_points(const std::string& path) {  // input.rs:172-194
    const char d = sniff(path);
    std::string buf;
    {
        std::ifstream f(path, std::ios::binary);
        f.seekg(0, std::ios::end);
        const std::streamoff n = f.tellg();
        f.seekg(0, std::ios::beg);
        if (n > 0) {
            buf.resize((size_t)n);
            f.read(&buf[0], n);
            buf.resize((size_t)f.gcount());
        }
    }
    std::vector<RawPoint> out;
    out.reserve(buf.size() / 40 + 16);
    size_t width = 0;
    const char* p = buf.data();
    const char* const end = p + buf.size();
    while (p < end) {
        const char* nl = (const char*)std::memchr(p, '\n', (size_t)(end - p));
        View line{p, nl ? nl : end};
        p = nl ? nl + 1 : end;
        if (strip_view(line).size() == 0) continue;
        if (line.e > line.b && line.e[-1] == '\r') --line.e;
        RawPoint pt;
        size_t nf = 0;
        bool ok;
        if (std::memchr(line.b, '\r', line.size())) {  // fields_of drops a '\r' wherever it stands: general path
            const auto fl = fields_of(std::string(line.b, line.e), d);
            nf = fl.size();
            ok = row_to_point(fl, pt);
        } else {
            View f[5];
            const char* q = line.b;
            for (;;) {
                const char* sep = (const char*)std::memchr(q, d, (size_t)(line.e - q));
                if (nf < 5) f[nf] = strip_view(View{q, sep ? sep : line.e});
                ++nf;
                if (!sep) break;
                q = sep + 1;
            }
            ok = (nf == 4 || nf == 5) && view_to_u32(f[0], pt.frame) && view_to_f64(f[1], pt.x) &&
                 view_to_f64(f[2], pt.y) && view_to_f64(f[3], pt.z);
            pt.aortic = false;
            if (ok && nf == 5) {
                const size_t n = f[4].size();
                if (n == 4 && std::memcmp(f[4].b, "true", 4) == 0)
                    pt.aortic = true;
                else if (!(n == 5 && std::memcmp(f[4].b, "false", 5) == 0))
                    ok = false;
            }
        }
        if (!width) width = nf;
        if (nf == width && ok) out.push_back(pt);
    }
    return out;
}
Input read_directory(const std::string& dir, bool diastole) {  // input.rs:62-147 + build.rs:20-27
    Input in;
    const std::string phase = diastole ? "diastolic" : "systolic";
    const std::string cp = dir + "/" + phase + "_contours.csv";
    if (!exists(cp)) throw InputErr("required contours file missing: \"" + cp + "\"");
    in.lumen = read_points(cp);
    const std::string rp = dir + "/" + phase + "_reference_points.csv";
    if (!exists(rp)) throw InputErr("required reference-point file missing: \"" + rp + "\"");
    {
        const char d = sniff(rp);
        std::ifstream f(rp);
        std::string line;
        bool got = false;
        while (std::getline(f, line)) {
            if (strip(line).empty()) continue;
            if (!row_to_point(fields_of(line, d), in.ref))
                throw InputErr("reading " + rp + ": failed to deserialize first reference-point record");
            got = true;
            break;
        }
        if (!got) throw InputErr("reference-point file \"" + rp + "\" was empty — this data is required");
    }
    auto optional = [&](const char* prefix, std::vector<RawPoint>& dst, bool& has) {
        const std::string p = dir + "/" + prefix + "_" + phase + "_contours.csv";
        if (exists(p)) {
            dst = read_points(p);
            has = true;
        }
    };
    optional("branch", in.side, in.has_side);
    optional("calcium", in.calc, in.has_calc);
    optional("eem", in.eem, in.has_eem);
    std::string rec = dir + "/combined_sorted_manual.csv";
    if (!exists(rec)) rec = dir + "/diastolic_systolic_records.csv";
    if (exists(rec)) {  // input.rs:237-251, Record (record.rs:3-11) by header name
        const char d = sniff(rec);
        std::ifstream f(rec);
        std::string line;
        if (std::getline(f, line)) {
            auto hdr = fields_of(line, d);
            int cf = -1, cph = -1, c1 = -1, c2 = -1;
            for (size_t i = 0; i < hdr.size(); ++i) {
                const std::string h = strip(hdr[i]);
                if (h == "frame") cf = (int)i;
                if (h == "phase") cph = (int)i;
                if (h == "measurement_1") c1 = (int)i;
                if (h == "measurement_2") c2 = (int)i;
            }
            if (cf < 0 || cph < 0 || c1 < 0 || c2 < 0) throw InputErr("reading " + rec + ": missing field");
            while (std::getline(f, line)) {
                if (strip(line).empty()) continue;
                auto fl = fields_of(line, d);
                if (fl.size() != hdr.size()) throw InputErr("reading " + rec + ": unequal record length");
                Rec r{};
                if (!to_u32(strip(fl[cf]), r.frame)) throw InputErr("reading " + rec + ": invalid frame");
                r.diastole_phase = fl[cph] == "D";
                r.systole_phase = fl[cph] == "S";
                r.has1 = to_f64(strip(fl[c1]), r.m1);
                r.has2 = to_f64(strip(fl[c2]), r.m2);
                in.rec.push_back(r);
            }
        }
        in.has_rec = true;
    }
    return in;
}

void integrity(const Geometry& g) {  // integrity_check.rs:8-247
    if (g.frames.empty()) throw InputErr("Geometry has no frames");
    for (size_t i = 0; i < g.frames.size(); ++i)
        if (g.frames[i].id != i)
            throw InputErr("Frame IDs are not consecutive. Expected ID " + std::to_string(i) + ", found ID " +
                           std::to_string(g.frames[i].id));
    auto near = [](double a, double b) { return std::fabs(a - b) < 1e-6; };
    for (auto& f : g.frames) {
        double l[3] = {0, 0, 0};
        if (f.lumen.has_c)
            std::copy(f.lumen.c, f.lumen.c + 3, l);
        else {
            Contour t = f.lumen;
            t.centroid();
            if (t.has_c) std::copy(t.c, t.c + 3, l);
        }
        if (!(near(f.c[0], l[0]) && near(f.c[1], l[1]) && near(f.c[2], l[2])))
            throw InputErr("Frame centroid does not match lumen centroid in frame " + std::to_string(f.id));
    }
    for (auto& f : g.frames) {
        if (f.lumen.size() == 0) throw InputErr("Lumen contour has no points in frame " + std::to_string(f.id));
        if (f.lumen.kind != kLumen) throw InputErr("Lumen contour has incorrect type in frame " + std::to_string(f.id));
    }
    size_t nref = 0;
    for (auto& f : g.frames) nref += f.has_ref;
    if (nref != 1) throw InputErr("Expected exactly one reference point, found " + std::to_string(nref));
    std::map<int, size_t> want;
    for (auto& f : g.frames) {
        auto chk = [&](int k, size_t n) {
            auto it = want.find(k);
            if (it == want.end())
                want[k] = n;
            else if (it->second != n)
                throw InputErr("contour point count mismatch in frame " + std::to_string(f.id) + ". Expected " +
                               std::to_string(it->second) + ", found " + std::to_string(n));
        };
        chk(kLumen, f.lumen.size());
        for (auto& kv : f.extras) chk(kv.second.kind, kv.second.size());
    }
    for (auto& f : g.frames) {
        for (auto& kv : f.extras)
            if (kv.second.original_frame != f.lumen.original_frame)
                throw InputErr("Original frame mismatch in frame " + std::to_string(f.id));
        if (f.has_ref && f.ref.fi != f.lumen.original_frame)
            throw InputErr("Reference point original frame mismatch in frame " + std::to_string(f.id));
    }
    size_t min_idx = 0;
    double min_z = std::numeric_limits<double>::infinity();
    for (size_t i = 0; i < g.frames.size(); ++i)
        if (g.frames[i].c[2] < min_z) min_z = g.frames[i].c[2], min_idx = i;
    if (g.proximal_idx() != min_idx)
        throw InputErr("Proximal end index is " + std::to_string(g.proximal_idx()) +
                       ", but frame with minimum z is " + std::to_string(min_idx));
    if (g.frames.front().c[2] > g.frames.back().c[2]) throw InputErr("First frame has higher z-coords than last frame");
}

Geometry build_geometry(const Input& in, const std::string& label, bool diastole, double icx, double icy, double radius,
                        uint32_t n_points) {  // build.rs:9-205
    Trace tr;
    // Rows of one frame are consecutive in every input we have seen, so the per-point map / set work below is done
    // once per RUN of equal frame ids (same result for any row order).
    std::set<uint32_t> seen;
    auto note = [&](const Points& pts) {
        for (size_t i = 0; i < pts.size(); ++i)
            if (i == 0 || pts.frame(i) != pts.frame(i - 1)) seen.insert(pts.frame(i));
    };
    note(in.lumen_pts());
    if (in.has_eem) note(in.eem_pts());
    if (in.has_calc) note(in.calc_pts());
    if (in.has_side) note(in.side_pts());
    seen.insert(in.ref.frame);
    std::map<uint32_t, uint32_t> slot;  // original frame -> sequential id
    for (uint32_t f : seen) slot.emplace(f, (uint32_t)slot.size());

    // group by original frame (Contour::build_contour_with_mapping, contour.rs:158-211)
    auto group = [&](const Points& pts, int kind) {
        std::map<uint32_t, Contour> by;
        struct Run {
            size_t i, j, at;
        };
        std::vector<std::pair<Contour*, std::vector<Run>>> jobs;  // map nodes do not move
        std::map<uint32_t, size_t> job_of;
        std::vector<size_t> total;
        for (size_t i = 0; i < pts.size();) {  // one append per run of equal frame ids
            size_t j = i + 1;
            const uint32_t fr = pts.frame(i);
            while (j < pts.size() && pts.frame(j) == fr) ++j;
            auto it = job_of.find(fr);
            if (it == job_of.end()) {
                it = job_of.emplace(fr, jobs.size()).first;
                jobs.push_back({&by[fr], {}});
                total.push_back(0);
            }
            jobs[it->second].second.push_back(Run{i, j, total[it->second]});
            total[it->second] += j - i;
            i = j;
        }
        auto fill = [&](size_t q) {
            Contour& c = *jobs[q].first;
            c.resize(total[q]);
            for (const Run& r : jobs[q].second)
                for (size_t k = r.i; k < r.j; ++k) {
                    const size_t w = r.at + (k - r.i);
                    c.fi[w] = pts.frame(k), c.pi[w] = 0, c.x[w] = pts.x(k), c.y[w] = pts.y(k), c.z[w] = pts.z(k);
                    c.ao[w] = pts.aortic(k) ? 1 : 0;
                }
        };
        if (pts.size() >= (1u << 15))
            parallel_for(jobs.size(), fill, kBlobThreads);  // first-touch of fresh memory: spreads over threads
        else
            for (size_t q = 0; q < jobs.size(); ++q) fill(q);
        for (auto& kv : by) {
            auto s = slot.find(kv.first);
            if (s == slot.end()) throw InputErr("No mapping found for original frame " + std::to_string(kv.first));
            kv.second.kind = kind;
            kv.second.id = s->second;
            kv.second.original_frame = kv.first;
        }
        return by;
    };
    tr.lap("ingest: frame ids");
    std::map<uint32_t, Frame> by_id;
    {
        auto lum = group(in.lumen_pts(), kLumen);
        tr.lap("ingest: group lumen");
        std::map<uint32_t, const Rec*> meas;
        if (in.has_rec)
            for (auto& r : in.rec) meas[r.frame] = &r;  // later rows overwrite earlier ones
        for (auto& kv : lum) {
            Contour& c = kv.second;
            auto m = meas.find(kv.first);
            if (m != meas.end()) {
                c.has_at = m->second->has1, c.at = m->second->m1;
                c.has_pt = m->second->has2, c.pt = m->second->m2;
            }
            c.centroid();
            Frame f;
            f.id = c.id;
            if (c.has_c) std::copy(c.c, c.c + 3, f.c);
            auto rs = slot.find(in.ref.frame);
            if (rs != slot.end() && rs->second == f.id) {
                f.has_ref = true;
                f.ref = RefPoint{in.ref.frame, 0, in.ref.x, in.ref.y, in.ref.z, in.ref.aortic};
            }
            f.lumen = std::move(c);
            by_id[f.id] = std::move(f);
        }
    }
    auto attach = [&](const Points& pts, int kind) {
        for (auto& kv : group(pts, kind)) {
            kv.second.centroid();
            auto it = by_id.find(kv.second.id);
            if (it != by_id.end()) it->second.extras[kind] = std::move(kv.second);
        }
    };
    if (in.has_eem) attach(in.eem_pts(), kEem);
    if (in.has_calc) attach(in.calc_pts(), kCalc);
    if (in.has_side) attach(in.side_pts(), kSide);
    tr.lap("ingest: frames + extras");
    if (n_points > 0) {  // Frame::create_catheter_points, frame.rs:163-204
        std::map<uint32_t, double> z_of;
        for (auto& kv : by_id) {
            const Contour& l = kv.second.lumen;
            for (size_t i = 0; i < l.size(); ++i)
                if (i == 0 || l.fi[i] != l.fi[i - 1]) z_of.emplace(l.fi[i], l.z[i]);  // emplace keeps the first z per frame
        }
        std::vector<RawPoint> cath;
        for (auto& kv : z_of)
            for (uint32_t i = 0; i < n_points; ++i) {
                const double a = 2.0 * kPi * (double)i / (double)n_points;
                cath.push_back(RawPoint{kv.first, icx + radius * std::cos(a), icy + radius * std::sin(a), kv.second, false});
            }
        auto by = group(Points{cath.data(), nullptr, cath.size()}, kCatheter);
        for (auto& kv : by) {
            for (size_t i = 0; i < kv.second.size(); ++i) kv.second.pi[i] = (uint32_t)i;
            kv.second.centroid();
            auto it = by_id.find(kv.second.id);
            if (it != by_id.end()) it->second.extras[kCatheter] = std::move(kv.second);
        }
    }
    tr.lap("ingest: catheter");
    Geometry g;
    g.label = label;
    for (auto& kv : by_id) g.frames.push_back(std::move(kv.second));

    if (in.has_rec) {  // Geometry::reorder_frames, geometry.rs:72-155
        std::map<uint32_t, double> z_orig;
        for (auto& f : g.frames)
            if (f.lumen.size()) z_orig.emplace(f.lumen.original_frame, f.lumen.z[0]);
        std::map<uint32_t, Frame> pool;
        for (auto& f : g.frames) pool[f.lumen.original_frame] = std::move(f);
        std::vector<Frame> ordered;
        for (auto& r : in.rec) {
            if (!(diastole ? r.diastole_phase : r.systole_phase)) continue;
            auto it = pool.find(r.frame);
            if (it != pool.end()) {
                ordered.push_back(std::move(it->second));
                pool.erase(it);
            }
        }
        for (auto& kv : pool) ordered.push_back(std::move(kv.second));
        for (size_t i = 0; i < ordered.size(); ++i) {
            Frame& f = ordered[i];
            const uint32_t id = (uint32_t)i;
            auto zi = z_orig.find(f.lumen.original_frame);
            const double z = zi != z_orig.end() ? zi->second : (double)id;
            f.id = id;
            auto fix = [&](Contour& c) {
                c.id = id;
                std::fill(c.fi.begin(), c.fi.end(), id);
                std::fill(c.z.begin(), c.z.end(), z);
                if (c.has_c) c.c[2] = z;
            };
            fix(f.lumen);
            for (auto& kv : f.extras) fix(kv.second);
            if (f.has_ref) f.ref.z = z;
            f.c[2] = z;
        }
        g.frames = std::move(ordered);
    }
    tr.lap("ingest: reorder");
    parallel_for(g.frames.size(), [&](size_t i) { g.frames[i].sort_points(); }, 8);  // frames are independent
    tr.lap("ingest: sort points");
    {  // ensure_proximal_at_position_zero, geometry.rs:325-381
        const size_t n = g.frames.size();
        if (n) {
            if (std::min(g.proximal_idx(), n - 1) != 0) std::reverse(g.frames.begin(), g.frames.end());
            std::vector<double> zs;
            for (auto& f : g.frames) zs.push_back(f.c[2]);
            std::stable_sort(zs.begin(), zs.end());
            for (size_t i = 0; i < n; ++i) {
                Frame& f = g.frames[i];
                f.id = (uint32_t)i;
                f.c[2] = zs[i];
                auto fix = [&](Contour& c) {
                    c.id = (uint32_t)i;  // final value after Frame::set_value(Some(id)), build.rs:190-193
                    std::fill(c.z.begin(), c.z.end(), zs[i]);
                    if (c.has_c) c.c[2] = zs[i];
                };
                fix(f.lumen);
                for (auto& kv : f.extras) fix(kv.second);
                if (f.has_ref) f.ref.z = zs[i];
            }
        }
    }
    tr.lap("ingest: proximal first");
    integrity(g);
    tr.lap("ingest: integrity");
    return g;
}

// =============================================================================
// Sweep plumbing
// =============================================================================
struct SweepUnits {
    std::vector<double> test, ref, centre;
    std::vector<int64_t> toff{0}, roff{0};
    size_t count() const { return toff.size() - 1; }
    void close_unit(double cx, double cy) {
        toff.push_back((int64_t)test.size() / 2);
        roff.push_back((int64_t)ref.size() / 2);
        centre.push_back(cx);
        centre.push_back(cy);
    }
};

struct Searcher {
    mmrs_ctx* ctx;
    int64_t* stats;
    std::mutex gpu;  // one context = one stream: sweeps issued from the chain-replay threads are serialised
    void check(int rc) {
        if (rc != MMRS_OK) throw std::runtime_error(mmrs_last_error(ctx));
    }
    // Grids of one stage: `centres` empty => centre None for every unit (one shared grid);
    // otherwise one grid per unit. A unit with skip[i] != 0 gets a degenerate grid (not swept).
    void make_grids(size_t U, double step_deg, double window_deg, double limes_deg, const std::vector<double>& centres,
                    const std::vector<char>* skip, std::vector<mmrs_grid>& grids, std::vector<int32_t>& which) {
        grids.clear();
        which.clear();
        if (centres.empty() && !skip) {
            grids.resize(1);
            check(mmrs_grid_from_reference_params(step_deg, window_deg, 0, 0.0, limes_deg, &grids[0]));
            return;
        }
        grids.resize(U);
        which.resize(U);
        for (size_t i = 0; i < U; ++i) {
            which[i] = (int32_t)i;
            if (skip && (*skip)[i]) {
                grids[i] = mmrs_grid{};
                grids[i].degenerate = 1;
                continue;
            }
            if (centres.empty())
                check(mmrs_grid_from_reference_params(step_deg, window_deg, 0, 0.0, limes_deg, &grids[i]));
            else
                check(mmrs_grid_from_reference_params(step_deg, window_deg, 1, centres[i], limes_deg, &grids[i]));
        }
    }
    // Multi-GPU: the sweep layer partitions the batches across the ranks of the context and merges the results
    // (mmrs_ctx_comm_init / mmrs_ctx_set_shard); every rank must therefore issue the same batched sweeps in the same
    // order (the chain's re-search rounds included: they are a function of the data only).
    std::vector<mmrs_unit_result> finish(size_t U, const std::vector<mmrs_grid>& grids,
                                         const std::vector<int32_t>& which) {
        std::vector<mmrs_unit_result> out(U);
        check(mmrs_sweep_run(ctx));
        check(mmrs_sweep_download(ctx, out.data()));
        for (size_t i = 0; i < U; ++i) {
            const mmrs_grid& g = grids[which.empty() ? 0 : i];
            if (g.degenerate) continue;
            stats[0] += 1;
            stats[1] += g.n_cand;
            stats[2] += out[i].n_shortlist > 0 ? out[i].n_shortlist : 0;
        }
        stats[4] += ctx->launches + ctx->upload_launches;
        return out;
    }
    // Uploads the point sets of `u` and runs the first stage of their search.
    std::vector<mmrs_unit_result> first(const SweepUnits& u, int mode, double step_deg, double window_deg,
                                        double limes_deg, const std::vector<double>& centres, double tie_margin) {
        const size_t U = u.count();
        if (U == 0) return {};
        std::vector<mmrs_grid> grids;
        std::vector<int32_t> which;
        make_grids(U, step_deg, window_deg, limes_deg, centres, nullptr, grids, which);
        mmrs_sweep_batch b{};
        b.n_units = (int64_t)U;
        b.test_xy = u.test.data();
        b.test_off = u.toff.data();
        b.ref_xy = u.ref.data();
        b.ref_off = u.roff.data();
        b.centre_xy = u.centre.data();
        b.grids = grids.data();
        b.n_grids = (int64_t)grids.size();
        b.grid_of_unit = which.empty() ? nullptr : which.data();
        b.mode = mode;
        mmrs_sweep_opts o{};
        o.tie_margin = tie_margin;
        o.partition = 0;   // the context's policy (mmrs_b200.h): units, candidate sub-ranges, or none for a tiny batch
        check(mmrs_sweep_upload(ctx, &b, &o));
        return finish(U, grids, which);
    }
    // Explicit per-unit grids (the chain's own re-search rounds): upload + run when `upload`, else regrid + run.
    // single_frame: a handful of single-frame sweeps — with a communicator bound they are split along the CANDIDATE axis.
    std::vector<mmrs_unit_result> run_grids(const SweepUnits& u, bool upload, int mode, const std::vector<mmrs_grid>& grids,
                                            const std::vector<int32_t>& which, double tie_margin, bool single_frame) {
        const size_t U = u.count();
        if (U == 0) return {};
        if (upload) {
            mmrs_sweep_batch b{};
            b.n_units = (int64_t)U;
            b.test_xy = u.test.data();
            b.test_off = u.toff.data();
            b.ref_xy = u.ref.data();
            b.ref_off = u.roff.data();
            b.centre_xy = u.centre.data();
            b.grids = grids.data();
            b.n_grids = (int64_t)grids.size();
            b.grid_of_unit = which.data();
            b.mode = mode;
            mmrs_sweep_opts o{};
            o.tie_margin = tie_margin;
            int32_t info[4] = {0, 1, 0, 0};
            mmrs_ctx_comm_info(ctx, info);
            // info[2] = the context's axis: 0 = the caller switched the partition off (ranks working on DIFFERENT cases, e.g.
            // a cohort dealt patient-wise): then no rank may enter a collective here — the other ranks are not in this call
            o.partition = info[2] == 0 ? -1 : 0;   // else the context's policy
            check(mmrs_sweep_upload(ctx, &b, &o));
        } else {
            check(mmrs_sweep_regrid(ctx, grids.data(), (int64_t)grids.size(), which.data(), tie_margin));
        }
        return finish(U, grids, which);
    }
    // Next window of the same search: the points stay on the device, only the grids change.
    std::vector<mmrs_unit_result> next(size_t U, double step_deg, double window_deg, double limes_deg,
                                       const std::vector<double>& centres, const std::vector<char>* skip,
                                       double tie_margin) {
        if (U == 0) return {};
        std::vector<mmrs_grid> grids;
        std::vector<int32_t> which;
        make_grids(U, step_deg, window_deg, limes_deg, centres, skip, grids, which);
        check(mmrs_sweep_regrid(ctx, grids.data(), (int64_t)grids.size(), which.empty() ? nullptr : which.data(), tie_margin));
        return finish(U, grids, which);
    }
};

// The stage list of a search: brute force = one stage (step, range, None); otherwise
// find_best_rotation's plan (align_within.rs:208-246 == align_between.rs:219-257).
struct Plan {
    int n = 0;
    double step[4], window[4];
};
Plan make_plan(double step_deg, double range_deg, bool bruteforce) {
    Plan p;
    if (bruteforce) {
        p.n = 1;
        p.step[0] = step_deg;
        p.window[0] = range_deg;
    } else {
        p.n = mmrs_stage_plan(step_deg, range_deg, p.step, p.window);
    }
    return p;
}

// downsample_contour_points, contour.rs:47-58
std::vector<size_t> stride_pick(size_t len, size_t n) {
    std::vector<size_t> idx;
    if (len <= n) {
        idx.resize(len);
        std::iota(idx.begin(), idx.end(), 0);
        return idx;
    }
    const double step = (double)len / (double)n;
    idx.resize(n);
    for (size_t i = 0; i < n; ++i) idx[i] = as_usize((double)i * step);
    return idx;
}

// catheter_lumen_vec_from_frames, align_within.rs:173-191 — appended as (x - ox, y - oy)
void gather_frame_sample(const Frame& f, size_t n_lumen, bool use_cath, size_t n_cath, double ox, double oy,
                         std::vector<double>& dst) {
    for (size_t i : stride_pick(f.lumen.size(), n_lumen)) {
        dst.push_back(f.lumen.x[i] - ox);
        dst.push_back(f.lumen.y[i] - oy);
    }
    if (use_cath)
        if (const Contour* c = f.extra(kCatheter))
            for (size_t i : stride_pick(c->size(), n_cath)) {
                dst.push_back(c->x[i] - ox);
                dst.push_back(c->y[i] - oy);
            }
}

// =============================================================================
// Post steps of align_frames_in_geometry (align_within.rs:136-158)
// =============================================================================
double dist3(const Contour& c, size_t i, size_t j) {
    const double dx = c.x[i] - c.x[j], dy = c.y[i] - c.y[j], dz = c.z[i] - c.z[j];
    return std::sqrt(dx * dx + dy * dy + dz * dz);
}
void farthest_pair(const Contour& c, size_t& bi, size_t& bj, double& best) {  // contour.rs:227-243
    best = 0.0, bi = 0, bj = 0;
    for (size_t i = 0; i < c.size(); ++i)
        for (size_t j = i + 1; j < c.size(); ++j) {
            const double d = dist3(c, i, j);
            if (d > best) best = d, bi = i, bj = j;
        }
}
double elliptic_ratio(const Contour& c) {  // contour.rs:313-343
    size_t i, j;
    double major;
    farthest_pair(c, i, j, major);
    const size_t n = c.size();
    if (n <= 2) throw InputErr("Need at least 3 points");
    double minor = std::numeric_limits<double>::max();
    for (size_t a = 0; a < n; ++a) minor = std::min(minor, dist3(c, a, (a + n / 2) % n));
    return major < minor ? minor / major : major / minor;
}

Contour blend(const Contour& a, const Contour& b, double t, bool average, uint32_t id, uint32_t of) {
    // average: avg_contour (align_within.rs:476-497); else fill_frame_gap (:573-598)
    Contour o;
    o.kind = a.kind;
    o.id = id;
    o.original_frame = of;
    const size_t n = std::min(a.size(), b.size());
    for (size_t i = 0; i < n; ++i) {
        if (average)
            o.push(of, (uint32_t)i, (a.x[i] + b.x[i]) / 2.0, (a.y[i] + b.y[i]) / 2.0, (a.z[i] + b.z[i]) / 2.0,
                   a.ao[i] || b.ao[i]);
        else
            o.push(of, (uint32_t)i, a.x[i] + (b.x[i] - a.x[i]) * t, a.y[i] + (b.y[i] - a.y[i]) * t,
                   a.z[i] + (b.z[i] - a.z[i]) * t, a.ao[i] || b.ao[i]);
    }
    if (a.has_c && b.has_c) {
        o.has_c = true;
        for (int k = 0; k < 3; ++k) o.c[k] = average ? (a.c[k] + b.c[k]) / 2.0 : a.c[k] + (b.c[k] - a.c[k]) * t;
    } else if (a.has_c || b.has_c) {
        o.has_c = true;
        std::copy(a.has_c ? a.c : b.c, (a.has_c ? a.c : b.c) + 3, o.c);
    }
    auto mix = [&](bool ha, double va, bool hb, double vb, bool& ho, double& vo) {
        if (ha && hb)
            ho = true, vo = average ? (va + vb) / 2.0 : va + (vb - va) * t;
        else if (ha)
            ho = true, vo = va;
        else if (hb)
            ho = true, vo = vb;
    };
    mix(a.has_at, a.at, b.has_at, b.at, o.has_at, o.at);
    mix(a.has_pt, a.pt, b.has_pt, b.pt, o.has_pt, o.pt);
    return o;
}
Frame bridge(const Frame& a, const Frame& b, double t, bool average) {
    // average: fix_one_frame_hole (:499-543); else create_interpolated_frame (:600-651)
    Frame o;
    for (int k = 0; k < 3; ++k) o.c[k] = average ? (a.c[k] + b.c[k]) / 2.0 : a.c[k] + (b.c[k] - a.c[k]) * t;
    o.lumen = blend(a.lumen, b.lumen, t, average, b.lumen.id, b.lumen.original_frame);
    std::set<int> kinds;
    for (auto& kv : a.extras) kinds.insert(kv.first);
    for (auto& kv : b.extras) kinds.insert(kv.first);
    for (int k : kinds) {
        const Contour *ca = a.extra(k), *cb = b.extra(k);
        o.extras[k] = (ca && cb) ? blend(*ca, *cb, t, average, cb->id, cb->original_frame) : (ca ? *ca : *cb);
    }
    if (!average) {
        if (a.has_ref && b.has_ref) {
            o.has_ref = true;
            o.ref = RefPoint{b.id, 0, a.ref.x + (b.ref.x - a.ref.x) * t, a.ref.y + (b.ref.y - a.ref.y) * t,
                             a.ref.z + (b.ref.z - a.ref.z) * t, a.ref.ao || b.ref.ao};
        } else if (a.has_ref || b.has_ref) {
            o.has_ref = true;
            o.ref = a.has_ref ? a.ref : b.ref;
        }
    }
    o.id = b.id;
    return o;
}
void fill_holes(Geometry& g) {  // align_within.rs:348-449
    std::vector<double> gaps;
    for (size_t i = 1; i < g.frames.size(); ++i) gaps.push_back(std::fabs(g.frames[i].c[2] - g.frames[i - 1].c[2]));
    if (gaps.empty()) return;
    std::vector<double> s = gaps;
    std::sort(s.begin(), s.end());
    const size_t n = s.size();
    const double base = (n % 2) ? s[n / 2] : (s[n / 2 - 1] + s[n / 2]) / 2.0;
    if (base <= std::numeric_limits<double>::epsilon()) return;
    if (!std::any_of(gaps.begin(), gaps.end(), [&](double d) { return d >= 1.5 * base; })) return;
    auto insert = [&](Frame f, size_t pos) {
        g.frames.insert(g.frames.begin() + pos, std::move(f));
        g.renumber();
    };
    size_t i = 1;
    while (i < g.frames.size()) {
        const Frame prev = g.frames[i - 1], cur = g.frames[i];
        const double ratio = std::fabs(cur.c[2] - prev.c[2]) / base;
        if (ratio < 1.5) {
            i += 1;
        } else if (ratio < 2.5) {
            insert(bridge(prev, cur, 0.5, true), i);
            i += 2;
        } else if (ratio < 3.5) {
            insert(bridge(prev, cur, 1.0 / 3.0, false), i);
            insert(bridge(prev, cur, 2.0 / 3.0, false), i + 1);
            i += 3;
        } else {
            const size_t missing = as_usize(std::fmax(std::floor(ratio - 1.0), 1.0));
            for (size_t k = 1; k <= missing; ++k)
                insert(bridge(prev, cur, (double)k / (double)(missing + 1), false), i + k - 1);
            i += missing + 1;
        }
    }
}
double ref_point_to_right(const Frame& f, bool anomalous) {  // align_within.rs:256-317
    if (!f.has_ref) throw InputErr("No reference point found in frame");
    double p1[2], p2[2];
    if (anomalous) {
        size_t i, j;
        double d;
        farthest_pair(f.lumen, i, j, d);
        p1[0] = f.lumen.x[i], p1[1] = f.lumen.y[i], p2[0] = f.lumen.x[j], p2[1] = f.lumen.y[j];
    } else {
        p1[0] = f.c[0], p1[1] = f.c[1], p2[0] = f.ref.x, p2[1] = f.ref.y;
    }
    const double line = std::atan2(p2[1] - p1[1], p2[0] - p1[0]);
    double rot = rem_euclid((anomalous ? kPi / 2.0 : 0.0) - line, 2.0 * kPi);
    const double ca = std::cos(rot), sa = std::sin(rot);
    auto rx = [&](double px, double py) {  // x of rotate2(pt, center = p1, rot)
        const double dx = px - p1[0], dy = py - p1[1];
        return (dx * ca - dy * sa) + p1[0];
    };
    const double ref_x = rx(f.ref.x, f.ref.y);
    const double eps = std::numeric_limits<double>::epsilon();
    bool ok = true;
    for (const double* op : {p1, p2}) {
        if (std::fabs(op[0] - f.ref.x) <= eps && std::fabs(op[1] - f.ref.y) <= eps) continue;
        if (ref_x <= rx(op[0], op[1])) {
            ok = false;
            break;
        }
    }
    if (!ok) rot = rem_euclid(rot + kPi, 2.0 * kPi);
    return rot;
}
Contour pushed_out(const Contour& src, double dist, bool ranged, uint32_t lo, uint32_t hi) {  // wall.rs:52-103
    Contour c = src;
    c.centroid();
    Contour o = c;
    o.kind = kWall;
    for (size_t i = 0; i < c.size(); ++i) {
        if (ranged && !(c.pi[i] >= lo && c.pi[i] <= hi)) continue;
        const double dx = c.x[i] - c.c[0], dy = c.y[i] - c.c[1], dz = c.z[i] - c.c[2];
        const double len = std::sqrt(dx * dx + dy * dy + dz * dz);
        if (len > std::numeric_limits<double>::epsilon()) {
            o.x[i] += dx / len * dist;
            o.y[i] += dy / len * dist;
            o.z[i] += dz / len * dist;
        }
    }
    return o;
}
Contour aortic_wall(const Contour& c) {  // wall.rs:112-210
    const size_t n = c.size(), q1 = n / 4, half = n / 2, q3 = q1 * 3;
    if (q3 >= n || half >= n) throw InputErr("index out of bounds (create_aortic_wall)");
    const double outer_x = c.x[q3] + c.at, z = c.z[q3];
    const double up_mid[2] = {c.x[0], c.y[0] + 1.0}, up_right[2] = {outer_x, up_mid[1]};
    const double low_mid[2] = {c.x[half], c.y[half] - 1.0}, low_right[2] = {outer_x, low_mid[1]};
    const double d_up = std::fabs(up_right[0] - up_mid[0]), d_right = std::fabs(up_right[1] - low_right[1]),
                 d_low = std::fabs(low_right[0] - low_mid[0]);
    const double total = d_up + d_right + d_low;
    const size_t n_up = as_usize(std::round(d_up / total * (double)half));
    const size_t n_mid = as_usize(std::round(d_right / total * (double)half));
    if (n_up + n_mid > half) throw InputErr("attempt to subtract with overflow (create_aortic_wall)");
    const size_t n_low = half - n_up - n_mid;
    std::vector<double> rxs, rys;
    for (size_t i = 0; i < n_low; ++i) {
        const double t = (double)i / (double)(n_low - 1);
        rxs.push_back(low_mid[0] + t * (low_right[0] - low_mid[0]));
        rys.push_back(low_mid[1]);
    }
    for (size_t i = 0; i < n_mid; ++i) {
        const double t = (double)i / (double)(n_mid - 1);
        rxs.push_back(low_right[0]);
        rys.push_back(low_right[1] + t * (up_right[1] - low_right[1]));
    }
    for (size_t i = 0; i < n_up; ++i) {
        const double t = (double)i / (double)(std::max<size_t>(n_up, 1) - 1);
        rxs.push_back(up_right[0] - t * (up_right[0] - up_mid[0]));
        rys.push_back(up_right[1]);
    }
    Contour o = pushed_out(c, 1.0, true, 0, (uint32_t)half);
    o.has_c = c.has_c;
    std::copy(c.c, c.c + 3, o.c);
    const size_t left = std::min(o.size(), (o.size() % 2) ? half + 1 : half);
    o.resize(left);
    for (size_t i = 0; i < rxs.size(); ++i) {
        const size_t s = left + i;
        if (s >= c.size()) throw InputErr("Index out of bounds (create_aortic_wall)");
        o.push(c.fi[s], c.pi[s], rxs[i], rys[i], z, c.ao[s]);
    }
    return o;
}
// Frames are independent in the post steps below; a pullback of a few hundred frames is walked frame-parallel on the host
// pool (nested under the per-pullback job), a short one serially.
constexpr size_t kFrameParallelMin = 32;
template <class F>
void for_frames(size_t n, F&& f) {
    if (n >= kFrameParallelMin) parallel_for(n, f, 8);
    else for (size_t i = 0; i < n; ++i) f(i);
}
void add_walls(Geometry& g, bool anomalous) {  // wall.rs:7-43
    for_frames(g.frames.size(), [&](size_t i) {
        Frame& f = g.frames[i];
        const Contour* eem = f.extra(kEem);
        const Contour& base = (anomalous || !eem) ? f.lumen : *eem;
        if (base.has_at) {
            Contour w = aortic_wall(base);   // built before the map is touched: `base` may live in f.extras
            f.extras[kWall] = std::move(w);
        } else {
            Contour w = pushed_out(base, 1.0, false, 0, 0);
            f.extras[kWall] = std::move(w);
        }
    });
}
void smooth(Geometry& g) {  // Geometry::smooth_frames, geometry.rs:165-239
    // Every smoothed contour is the OLD contour truncated to the frame's lumen count with x, y replaced by the 3-frame mean
    // of the OLD values, so only the old x / y of the three kinds involved are saved (not a deep copy of the geometry) and
    // the contours are rewritten in place, frame-parallel.
    constexpr int kKinds[3] = {-1, (int)kEem, (int)kWall};   // -1 = the lumen
    struct Saved {
        std::vector<double> x, y;
        bool present = false;
    };
    const size_t nf = g.frames.size();
    std::vector<std::array<Saved, 3>> old(nf);
    auto contour_of = [](Frame& f, int kind) -> Contour* { return kind < 0 ? &f.lumen : f.extra(kind); };
    for_frames(nf, [&](size_t i) {
        for (int q = 0; q < 3; ++q)
            if (Contour* c = contour_of(g.frames[i], kKinds[q])) old[i][q].x = c->x, old[i][q].y = c->y, old[i][q].present = true;
    });
    for_frames(nf, [&](size_t i) {
        const size_t ip = i == 0 ? i : i - 1, in = i == nf - 1 ? i : i + 1;
        const size_t count = old[i][0].x.size();
        for (int q = 0; q < 3; ++q) {
            const Saved &cur = old[i][q], &p = old[ip][q], &n = old[in][q];
            if (!(cur.present && p.present && n.present)) continue;   // the lumen is always present
            if (cur.x.size() < count || p.x.size() < count || n.x.size() < count)
                throw InputErr("index out of bounds (smooth_frames)");
            Contour& o = *contour_of(g.frames[i], kKinds[q]);
            o.resize(count);
            for (size_t j = 0; j < count; ++j) {
                o.x[j] = (p.x[j] + cur.x[j] + n.x[j]) / 3.0;
                o.y[j] = (p.y[j] + cur.y[j] + n.y[j]) / 3.0;
            }
            o.centroid();
        }
    });
}

// =============================================================================
// Intrapullback alignment of MANY pullbacks at once (align_within.rs:24-171)
// =============================================================================
struct WithinOut {
    std::vector<double> logs;  // n x 7
    bool anomalous = false;
};

void align_within_many(Searcher& S, std::vector<Geometry*>& geoms, const mmrs_align_params& P,
                       std::vector<WithinOut>& outs) {
    Trace tr;
    const size_t G = geoms.size();
    outs.assign(G, WithinOut{});
    struct Meta {
        size_t first_unit = 0, ref_idx = 0;
        bool use_cath = false;
        size_t n_cath = 0;
    };
    std::vector<Meta> meta(G);
    const size_t sample = (size_t)std::max<int64_t>(P.sample_size, 0);
    // guards (align_within.rs:32-40) and sampling parameters (:42-59)
    for (size_t g = 0; g < G; ++g) {
        Geometry& geo = *geoms[g];
        if (geo.frames.empty()) throw InputErr("Geometry contains no frames");
        if (geo.frames[0].lumen.size() == 0) throw InputErr("Lumen contours have no points");
        if (P.sample_size <= 0) throw InputErr("sample_size must be > 0");
        meta[g].ref_idx = geo.ref_or_proximal();
        const double ratio = (double)sample / (double)geo.frames[0].lumen.size();
        if (const Contour* c = geo.frames[0].extra(kCatheter)) {
            meta[g].use_cath = true;
            meta[g].n_cath = as_usize(std::ceil((double)c->size() * ratio));
        }
    }
    tr.lap("within: guards");
    // 1. decoupled units from the ORIGINAL frames, each centred on its own frame centroid. Frame i of a pullback is the
    //    test set of unit i and the reference set of unit i + 1 with the SAME centring, so every frame is sampled once
    //    (frame-parallel over all pullbacks) and the two point arrays are assembled from those samples by block copies.
    std::vector<size_t> frame0(G + 1, 0);   // index of a pullback's first frame in the flat list of frames
    for (size_t g = 0; g < G; ++g) frame0[g + 1] = frame0[g] + geoms[g]->frames.size();
    std::vector<std::vector<double>> fs(frame0[G]);
    {
        std::vector<uint32_t> owner(frame0[G]);
        for (size_t g = 0; g < G; ++g)
            for (size_t k = frame0[g]; k < frame0[g + 1]; ++k) owner[k] = (uint32_t)g;
        auto sample_frame = [&](size_t k) {
            const size_t g = owner[k];
            const Frame& f = geoms[g]->frames[k - frame0[g]];
            fs[k].reserve(2 * (std::min(sample, f.lumen.size()) + meta[g].n_cath));
            gather_frame_sample(f, sample, meta[g].use_cath, meta[g].n_cath, f.c[0], f.c[1], fs[k]);
        };
        if (fs.size() >= 64) parallel_for(fs.size(), sample_frame);
        else for (size_t k = 0; k < fs.size(); ++k) sample_frame(k);
    }
    tr.lap("within: sample frames");
    SweepUnits units;
    // the two big point arrays live in the context between calls (warm pages); handed back when this function leaves
    struct Lend {
        SweepUnits& u;
        mmrs_ctx* ctx;
        Lend(SweepUnits& u_, mmrs_ctx* c) : u(u_), ctx(c) {
            u.test.swap(ctx->host_units_test);
            u.ref.swap(ctx->host_units_ref);
            u.test.clear();
            u.ref.clear();
        }
        ~Lend() {
            u.test.swap(ctx->host_units_test);
            u.ref.swap(ctx->host_units_ref);
        }
    } lend(units, S.ctx);
    std::vector<size_t> unit_frame;   // flat frame index of each unit's test frame (its reference frame is the one before)
    for (size_t g = 0; g < G; ++g) {
        meta[g].first_unit = units.count();
        for (size_t k = frame0[g] + 1; k < frame0[g + 1]; ++k) {
            unit_frame.push_back(k);
            units.toff.push_back(units.toff.back() + (int64_t)fs[k].size() / 2);
            units.roff.push_back(units.roff.back() + (int64_t)fs[k - 1].size() / 2);
        }
    }
    units.centre.assign(2 * unit_frame.size(), 0.0);
    units.test.resize(2 * (size_t)units.toff.back());
    units.ref.resize(2 * (size_t)units.roff.back());
    {
        auto place = [&](size_t u) {
            const size_t k = unit_frame[u];
            if (!fs[k].empty()) std::memcpy(units.test.data() + 2 * units.toff[u], fs[k].data(), fs[k].size() * sizeof(double));
            if (!fs[k - 1].empty())
                std::memcpy(units.ref.data() + 2 * units.roff[u], fs[k - 1].data(), fs[k - 1].size() * sizeof(double));
        };
        if (unit_frame.size() >= 64) parallel_for(unit_frame.size(), place, 8);
        else for (size_t u = 0; u < unit_frame.size(); ++u) place(u);
    }
    fs.clear();
    const size_t U = units.count();
    tr.lap("within: build units");
    const Plan plan = make_plan(P.step_deg, P.range_deg, P.bruteforce != 0);
    // 2. stage-by-stage batched search; a unit stays "certified" while every stage has a unique
    //    winner by more than kTieMargin
    std::vector<double> angle(U, 0.0);
    std::vector<int> good_stages(U, 0);  // number of leading stages certified
    {
        std::vector<char> dropped(U, 0);
        size_t live = U;
        for (int s = 0; s < plan.n && live > 0; ++s) {
            auto res = (s == 0) ? S.first(units, 0, plan.step[0], plan.window[0], P.range_deg, {}, kTieMargin)
                                : S.next(U, plan.step[s], plan.window[s], P.range_deg, angle, &dropped, kTieMargin);
            for (size_t u = 0; u < U; ++u) {
                if (dropped[u]) continue;
                const bool unique = (res[u].flags & MMRS_FLAG_DEGENERATE) || res[u].n_ties == 1;
                if (unique) {
                    angle[u] = res[u].best_angle;
                    good_stages[u] = s + 1;
                } else {
                    dropped[u] = 1;  // resolved on the chain below
                    --live;
                }
            }
        }
    }
    tr.lap("within: batched sweeps");
    // 3. replay the chain on the host (one thread per pullback). An uncertified frame is searched again on the chain's own
    //    points — the reference's computation verbatim. The chains run in ROUNDS: every pullback advances to its next
    //    uncertified frame (or its end), the pending frames of all pullbacks are searched in ONE batched sweep, and the
    //    chains continue. The order of the sweeps is therefore a function of the data only: every rank of a multi-GPU run
    //    issues the same batches, so they can be partitioned like any other (along the CANDIDATE axis when a communicator
    //    is bound: a round is a handful of single-frame sweeps, mmrs_b200.h "axis 2").
    struct Chain {
        size_t i = 1;
        double cumulative = 0.0, tx = 0.0, ty = 0.0, result = 0.0;
        bool placed = false, have_result = false, done = false;
    };
    std::vector<Chain> chain(G);
    std::vector<char> pending(G, 0);
    auto post_steps = [&](size_t g) {
        Geometry& geo = *geoms[g];
        // 4. post steps (align_within.rs:136-158)
        fill_holes(geo);
        if (meta[g].ref_idx >= geo.frames.size()) throw InputErr("index out of bounds: reference frame");
        const Frame& rf = geo.frames[meta[g].ref_idx];
        const bool anomalous = elliptic_ratio(rf.lumen) > 2.0 || rf.lumen.has_at || rf.lumen.has_pt;
        const double extra = ref_point_to_right(rf, anomalous);
        if (extra != 0.0)  // Geometry::rotate_geometry, geometry.rs:241-250
            for_frames(geo.frames.size(), [&](size_t i) {
                Frame& f = geo.frames[i];
                f.spin(extra, f.c[0], f.c[1]);
                f.sort_points();
            });
        if (anomalous)  // assign_aortic, :319-331
            for (auto& f : geo.frames) {
                const size_t len = f.lumen.size();
                for (size_t k = 0; k < len; ++k) f.lumen.ao[k] = k >= len / 2;
            }
        add_walls(geo, anomalous);
        if (P.smooth) smooth(geo);
        outs[g].anomalous = anomalous;
        };
    for (;;) {
        parallel_for(G, [&](size_t g) {
            Geometry& geo = *geoms[g];
            Chain& ch = chain[g];
            if (ch.done) return;
            for (;;) {
                if (ch.i >= geo.frames.size()) {
                    ch.done = true;
                    return;
                }
                const size_t u = meta[g].first_unit + (ch.i - 1);
                const Frame& prev = geo.frames[ch.i - 1];
                Frame& cur = geo.frames[ch.i];
                if (!ch.placed) {
                    if (ch.cumulative != 0.0) cur.spin(ch.cumulative, cur.c[0], cur.c[1]);
                    ch.tx = prev.c[0] - cur.c[0], ch.ty = prev.c[1] - cur.c[1];
                    cur.shift(ch.tx, ch.ty, 0.0);
                    ch.placed = true;
                }
                double best = angle[u];
                if (good_stages[u] < plan.n) {
                    if (!ch.have_result) {
                        pending[g] = 1;   // searched by the batched sweep below, then the chain goes on
                        return;
                    }
                    best = ch.result;
                    ch.have_result = false;
                }
                cur.spin(best, cur.c[0], cur.c[1]);
                ch.cumulative += best;
                const double row[7] = {(double)cur.id, (double)prev.id, rad2deg(best), ch.tx, ch.ty, cur.c[0], cur.c[1]};
                outs[g].logs.insert(outs[g].logs.end(), row, row + 7);
                ch.i += 1;
                ch.placed = false;
            }
        });
        std::vector<size_t> todo;
        for (size_t g = 0; g < G; ++g)
            if (pending[g]) todo.push_back(g);
        if (todo.empty()) break;
        // the pending frames on the chain's own points (centre = the frame centroid), remaining stages of each search
        SweepUnits batch;
        std::vector<int> stage0(todo.size());
        int rounds = 0;
        for (size_t k = 0; k < todo.size(); ++k) {
            const size_t g = todo[k];
            const Geometry& geo = *geoms[g];
            const Frame &cur = geo.frames[chain[g].i], &prev = geo.frames[chain[g].i - 1];
            gather_frame_sample(cur, sample, meta[g].use_cath, meta[g].n_cath, 0.0, 0.0, batch.test);
            gather_frame_sample(prev, sample, meta[g].use_cath, meta[g].n_cath, 0.0, 0.0, batch.ref);
            batch.close_unit(cur.c[0], cur.c[1]);
            stage0[k] = good_stages[meta[g].first_unit + (chain[g].i - 1)];
            rounds = std::max(rounds, plan.n - stage0[k]);
        }
        std::vector<double> best(todo.size(), 0.0);
        for (size_t k = 0; k < todo.size(); ++k) best[k] = angle[meta[todo[k]].first_unit + (chain[todo[k]].i - 1)];
        for (int t = 0; t < rounds; ++t) {
            // unit k runs stage stage0[k] + t of the plan (its own step / window / centre); finished searches sit out
            std::vector<mmrs_grid> grids(todo.size());
            std::vector<int32_t> which(todo.size());
            for (size_t k = 0; k < todo.size(); ++k) {
                which[k] = (int32_t)k;
                const int sidx = stage0[k] + t;
                grids[k] = mmrs_grid{};
                grids[k].degenerate = 1;
                if (sidx >= plan.n) continue;
                S.check(mmrs_grid_from_reference_params(plan.step[sidx], plan.window[sidx], sidx > 0 ? 1 : 0, best[k],
                                                        P.range_deg, &grids[k]));
            }
            auto r = S.run_grids(batch, t == 0, 0, grids, which, 0.0, /*single_frame=*/true);
            for (size_t k = 0; k < todo.size(); ++k)
                if (stage0[k] + t < plan.n) best[k] = r[k].best_angle;
        }
        for (size_t k = 0; k < todo.size(); ++k) {
            chain[todo[k]].result = best[k];
            chain[todo[k]].have_result = true;
            pending[todo[k]] = 0;
            S.stats[3] += 1;
        }
    }
    tr.lap("within: chain replay");
    parallel_for(G, [&](size_t g) { post_steps(g); });
    tr.lap("within: post steps");
}

// =============================================================================
// Inter-pullback alignment of MANY pairs at once (align_between.rs:11-92)
// =============================================================================
void sample_cloud(const Geometry& g, size_t sample, std::vector<double>& dst) {  // :154-178
    size_t total = 0;
    for (auto& f : g.frames) total += f.lumen.size();
    const double ratio = (double)sample / (double)total;
    for (auto& f : g.frames) {
        const size_t k = std::max<size_t>(as_usize(std::ceil((double)f.lumen.size() * ratio)), 1);
        for (size_t i : stride_pick(f.lumen.size(), k)) {
            dst.push_back(f.lumen.x[i]);
            dst.push_back(f.lumen.y[i]);
        }
    }
}

void align_between_many(Searcher& S, std::vector<std::pair<Geometry*, Geometry*>>& pairs, const mmrs_align_params& P) {
    const size_t N = pairs.size();
    if (!N) return;
    const size_t sample = std::max<size_t>((size_t)std::max<int64_t>(P.sample_size, 0), 500);
    std::vector<std::array<double, 2>> pivot(N);
    // The pairs of one level move disjoint geometries (AB | CD, then AC | BD: only the second of a pair is written), so
    // they are prepared side by side, each walking its frames on the host pool; the units are appended in pair order.
    auto shift_frames = [](Geometry& g, double dx, double dy, double dz) {
        for_frames(g.frames.size(), [&](size_t i) { g.frames[i].shift(dx, dy, dz); });
    };
    struct Prep {
        std::vector<double> ref, test;
        double cx = 0.0, cy = 0.0;
    };
    std::vector<Prep> prep(N);
    parallel_for(N, [&](size_t k) {
        Geometry &a = *pairs[k].first, &b = *pairs[k].second;
        if (a.frames.empty() || b.frames.empty()) throw InputErr("index out of bounds: empty geometry");
        const size_t ia = a.ref_or_proximal(), ib = b.ref_or_proximal();
        if (ia >= a.frames.size() || ib >= b.frames.size()) throw InputErr("index out of bounds: reference frame");
        const Frame &fa = a.frames[ia], &fb = b.frames[ib];
        pivot[k] = {fa.c[0], fa.c[1]};
        shift_frames(b, fa.c[0] - fb.c[0], fa.c[1] - fb.c[1], fa.c[2] - fb.c[2]);
        sample_cloud(a, sample, prep[k].ref);
        sample_cloud(b, sample, prep[k].test);
        // calculate_global_centroid of the reference cloud, :260-271 (x then y, sequential sums)
        double sx = 0.0, sy = 0.0;
        const size_t cnt = prep[k].ref.size() / 2;
        for (size_t i = 0; i < cnt; ++i) sx += prep[k].ref[2 * i];
        for (size_t i = 0; i < cnt; ++i) sy += prep[k].ref[2 * i + 1];
        prep[k].cx = cnt ? sx / (double)cnt : 0.0;
        prep[k].cy = cnt ? sy / (double)cnt : 0.0;
    });
    SweepUnits units;
    for (size_t k = 0; k < N; ++k) {
        units.ref.insert(units.ref.end(), prep[k].ref.begin(), prep[k].ref.end());
        units.test.insert(units.test.end(), prep[k].test.begin(), prep[k].test.end());
        units.close_unit(prep[k].cx, prep[k].cy);
    }
    prep.clear();
    // there is no brute-force switch on this path (align_between.rs:219-257)
    const Plan plan = make_plan(P.step_deg, P.range_deg, false);
    std::vector<double> angle(N, 0.0);
    for (int s = 0; s < plan.n; ++s) {
        auto res = (s == 0) ? S.first(units, 1, plan.step[0], plan.window[0], P.range_deg, {}, 0.0)
                            : S.next(N, plan.step[s], plan.window[s], P.range_deg, angle, nullptr, 0.0);
        for (size_t k = 0; k < N; ++k) angle[k] = res[k].best_angle;
    }
    parallel_for(N, [&](size_t k) {
        Geometry &a = *pairs[k].first, &b = *pairs[k].second;
        // rotate_geometry_around_point, :95-145 (contour centroids rotate too, unlike Frame::rotate)
        const double ca = std::cos(angle[k]), sa = std::sin(angle[k]), cx = pivot[k][0], cy = pivot[k][1];
        auto rot = [&](double& x, double& y) {
            const double tx = x - cx, ty = y - cy;
            const double rx = tx * ca - ty * sa, ry = tx * sa + ty * ca;
            x = rx + cx;
            y = ry + cy;
        };
        for_frames(b.frames.size(), [&](size_t fi) {
            Frame& f = b.frames[fi];
            for (size_t i = 0; i < f.lumen.size(); ++i) rot(f.lumen.x[i], f.lumen.y[i]);
            rot(f.c[0], f.c[1]);
            for (auto& kv : f.extras) {
                for (size_t i = 0; i < kv.second.size(); ++i) rot(kv.second.x[i], kv.second.y[i]);
                if (kv.second.has_c) rot(kv.second.c[0], kv.second.c[1]);
            }
            if (f.has_ref) rot(f.ref.x, f.ref.y);
        });
        const Frame &fa = a.frames.at(a.ref_or_proximal()), &fb = b.frames.at(b.ref_or_proximal());
        const double dx = fa.c[0] - fb.c[0], dy = fa.c[1] - fb.c[1], dz = fa.c[2] - fb.c[2];   // by value: fb is shifted too
        shift_frames(b, dx, dy, dz);
    });
}

// =============================================================================
// Pair post-processing (processing/postprocessing.rs:12-87): equalise the z-sampling of the two
// pullbacks of a pair, put the reference frames at the same z, trim to the common span around the
// reference frame, optionally rebuild the aortic walls with a shared thickness. Host-only f64.
// =============================================================================
double mean_z_step(const Geometry& g) {  // get_avg_z_diff, :100-114
    if (g.frames.size() < 2) return 0.0;
    double sum = 0.0;
    for (size_t i = 1; i < g.frames.size(); ++i) sum += g.frames[i].c[2] - g.frames[i - 1].c[2];
    return sum / (double)(g.frames.size() - 1);
}
void set_frame_z(Frame& f, double z) {  // Frame::set_value(None, None, None, Some(z)), frame.rs:95-118
    auto fix = [z](Contour& c) {
        std::fill(c.z.begin(), c.z.end(), z);
        if (c.has_c) c.c[2] = z;
    };
    fix(f.lumen);
    for (auto& kv : f.extras) fix(kv.second);
    if (f.has_ref) f.ref.z = z;
    f.c[2] = z;
}
bool find_ref(const Geometry& g, size_t& idx) {  // find_ref_frame_idx, geometry.rs:62-69
    for (auto& f : g.frames)
        if (f.has_ref) {
            idx = f.id;
            return true;
        }
    return false;
}
Geometry respace(const Geometry& src, double step) {  // resample_by_diff, :116-140
    Geometry g = src;
    if (g.frames.empty()) throw InputErr("index out of bounds: resample_by_diff on an empty geometry");
    size_t lo = 0;
    for (size_t i = 1; i < g.frames.size(); ++i)
        if (g.frames[i].c[2] < g.frames[lo].c[2]) lo = i;
    if (lo) std::rotate(g.frames.begin(), g.frames.begin() + lo, g.frames.end());
    const double z0 = g.frames[0].c[2];
    for (size_t i = 1; i < g.frames.size(); ++i) set_frame_z(g.frames[i], z0 + (double)i * step);
    return g;
}
std::vector<double> z_ladder(double ref_z, double start, double stop, double step) {  // predict_z_positions, :142-195
    std::vector<double> z;
    if (!std::isfinite(step) || step == 0.0) return z;
    const double eps = 1e-9;
    auto walk = [&](double cur, double inc, bool up, double limit) {
        while (up ? cur <= limit + eps : cur >= limit - eps) {
            z.push_back(cur);
            cur += inc;
            if (!std::isfinite(cur)) break;
        }
    };
    if (std::fabs(ref_z - start) > eps && std::fabs(ref_z - stop) > eps) {
        walk(ref_z, -step, false, start);
        std::stable_sort(z.begin(), z.end());
        walk(ref_z + step, step, true, stop);
    } else if (stop >= start && step > 0.0) {
        walk(start, step, true, stop);
    } else if (stop <= start && step < 0.0) {
        walk(start, step, false, stop);
    }
    return z;
}
Contour lerp_contour(const Contour& a, const Contour& b, double t) {  // blend_contour, :302-340
    Contour o = a;
    const size_t n = std::min(a.size(), b.size());
    o.resize(n);
    for (size_t i = 0; i < n; ++i) {
        o.x[i] = a.x[i] + t * (b.x[i] - a.x[i]);
        o.y[i] = a.y[i] + t * (b.y[i] - a.y[i]);
    }
    o.has_c = a.has_c && b.has_c;
    if (o.has_c)
        for (int k = 0; k < 3; ++k) o.c[k] = a.c[k] + t * (b.c[k] - a.c[k]);
    o.has_at = a.has_at && b.has_at;
    o.at = o.has_at ? a.at + t * (b.at - a.at) : 0.0;
    o.has_pt = a.has_pt && b.has_pt;
    o.pt = o.has_pt ? a.pt + t * (b.pt - a.pt) : 0.0;
    return o;
}
Geometry resample_at(const Geometry& g, std::vector<double> zs) {  // new_frames_by_sample_rate, :197-300
    if (g.frames.empty()) throw InputErr("index out of bounds: new_frames_by_sample_rate on an empty geometry");
    std::stable_sort(zs.begin(), zs.end());
    const double top = g.frames.back().c[2];
    Geometry o;
    o.label = g.label;
    for (double z : zs) {
        if (z > top) break;
        auto hit = std::find_if(g.frames.begin(), g.frames.end(), [z](const Frame& f) { return std::fabs(f.c[2] - z) < 1e-9; });
        if (hit != g.frames.end()) {
            o.frames.push_back(*hit);
            continue;
        }
        size_t k = 0;
        while (k + 1 < g.frames.size() && !(g.frames[k].c[2] <= z && g.frames[k + 1].c[2] >= z)) ++k;
        if (k + 1 >= g.frames.size()) throw InputErr("Cannot find frames to interpolate between");
        const Frame &lo = g.frames[k], &up = g.frames[k + 1];
        const double t = (z - lo.c[2]) / (up.c[2] - lo.c[2]);
        Frame f;
        f.id = lo.id;
        f.lumen = lerp_contour(lo.lumen, up.lumen, t);
        for (int kind : {(int)kEem, (int)kCalc, (int)kSide, (int)kCatheter, (int)kWall}) {
            const Contour *a = lo.extra(kind), *b = up.extra(kind);
            if (a && b) f.extras[kind] = lerp_contour(*a, *b, t);
        }
        f.c[0] = lo.c[0] + t * (up.c[0] - lo.c[0]);
        f.c[1] = lo.c[1] + t * (up.c[1] - lo.c[1]);
        f.c[2] = z;
        o.frames.push_back(std::move(f));
    }
    std::stable_sort(o.frames.begin(), o.frames.end(), [](const Frame& a, const Frame& b) { return a.c[2] < b.c[2]; });
    for (size_t i = 0; i < o.frames.size(); ++i) {
        Frame& f = o.frames[i];
        f.id = (uint32_t)i;
        f.lumen.id = (uint32_t)i;
        std::fill(f.lumen.z.begin(), f.lumen.z.end(), f.c[2]);
        if (f.lumen.has_c) f.lumen.c[2] = f.c[2];
        for (auto& kv : f.extras) {
            kv.second.id = (uint32_t)i;
            std::fill(kv.second.z.begin(), kv.second.z.end(), f.c[2]);
        }
        if (f.has_ref) f.ref.z = f.c[2];
    }
    return o;
}
void postprocess_pair(Geometry& ga, Geometry& gb, double tol, bool anomalous) {  // postprocess_geom_pair, :12-87
    const double da = mean_z_step(ga), db = mean_z_step(gb);
    size_t ra, rb;
    if (!find_ref(ga, ra) || !find_ref(gb, rb)) throw InputErr("No reference point found in any frame");
    const double ref_za = ga.frames.at(ra).c[2], ref_zb = gb.frames.at(rb).c[2];
    auto span = [](const Geometry& g) {
        const double a = g.frames.front().c[2], b = g.frames.back().c[2];
        return a < b ? std::make_pair(a, b) : std::make_pair(b, a);
    };
    Geometry na, nb;
    if ((da - db) < tol) {  // sic: signed difference (:93)
        const double mean = (da + db) / 2.0;
        na = respace(ga, mean);
        nb = respace(gb, mean);
    } else if (da < db) {
        const auto s = span(gb);
        nb = resample_at(gb, z_ladder(ref_zb, s.first, s.second, da));
        na = respace(ga, da);
    } else {
        const auto s = span(ga);
        na = resample_at(ga, z_ladder(ref_za, s.first, s.second, db));
        nb = respace(gb, db);
    }
    size_t ra2, rb2;
    if (!find_ref(na, ra2) || !find_ref(nb, rb2)) throw InputErr("No reference point found in any frame");
    if (ra2 >= ga.frames.size() || rb2 >= gb.frames.size()) throw InputErr("index out of bounds: postprocess_geom_pair");
    na.shift_all(0.0, 0.0, ga.frames[ra2].c[2] - gb.frames[rb2].c[2]);  // sic: indexes the ORIGINAL pair (:76-77)
    // trim_geom_pair, :342-409
    size_t ta = 0, tb = 0;
    find_ref(na, ta);
    find_ref(nb, tb);
    if (ta > na.frames.size() || tb > nb.frames.size()) throw InputErr("attempt to subtract with overflow (trim_geom_pair)");
    const size_t before = std::min(ta, tb), after = std::min(na.frames.size() - ta, nb.frames.size() - tb);
    auto cut = [&](Geometry& g, size_t ref) {
        const size_t lo = ref - before, hi = ref + after;
        if (lo < hi && hi <= g.frames.size()) g.frames = std::vector<Frame>(g.frames.begin() + lo, g.frames.begin() + hi);
        for (size_t i = 0; i < g.frames.size(); ++i) {
            g.frames[i].id = (uint32_t)i;
            g.frames[i].lumen.id = (uint32_t)i;
            for (auto& kv : g.frames[i].extras) kv.second.id = (uint32_t)i;
        }
    };
    cut(na, ta);
    cut(nb, tb);
    if (anomalous) {  // adjust_walls_anomalous_geom_pair, :411-470 (zip => the shorter length)
        const size_t n = std::min(na.frames.size(), nb.frames.size());
        na.frames.resize(n);
        nb.frames.resize(n);
        for (size_t i = 0; i < n; ++i) {
            Contour &la = na.frames[i].lumen, &lb = nb.frames[i].lumen;
            if (la.has_at || lb.has_at) {
                const double t = (la.has_at && lb.has_at) ? (la.at + lb.at) / 2.0 : (la.has_at ? la.at : lb.at);
                la.has_at = lb.has_at = true;
                la.at = lb.at = t;
            }
        }
        add_walls(na, true);
        add_walls(nb, true);
    }
    na.label = ga.label;
    nb.label = gb.label;
    ga = std::move(na);
    gb = std::move(nb);
}

#include "mmrs_export_align.inc"

template <class F>
int guarded(mmrs_ctx* ctx, F&& f) {
    try {
        f();
        return MMRS_OK;
    } catch (const InputErr& e) {
        return mmrs::set_err(ctx, MMRS_ERR_INPUT, e.what());
    } catch (const std::exception& e) {
        return mmrs::set_err(ctx, MMRS_ERR_CUDA, e.what());
    }
}

}  // namespace

// =============================================================================
// C ABI
// =============================================================================
extern "C" void mmrs_free(void* p) { std::free(p); }

extern "C" int mmrs_geometry_from_dir(mmrs_ctx* ctx, const char* path, const char* label, int diastole, double icx,
                                      double icy, double radius, uint32_t n_points, double** blob_out,
                                      int64_t* len_out) {
    if (!path || !label || !blob_out || !len_out) return mmrs::set_err(ctx, MMRS_ERR_ARG, "mmrs_geometry_from_dir: NULL argument");
    return guarded(ctx, [&] {
        const Input in = read_directory(path, diastole != 0);
        *blob_out = encode_malloc(build_geometry(in, label, diastole != 0, icx, icy, radius, n_points), len_out);
    });
}

extern "C" int mmrs_geometry_from_arrays(mmrs_ctx* ctx, const double* lumen, int64_t n_lumen, const double* eem,
                                         int64_t n_eem, const double* calc, int64_t n_calc, const double* side,
                                         int64_t n_side, const double* records, int64_t n_rec, const double* ref_point,
                                         int diastole, const char* label, double icx, double icy, double radius,
                                         uint32_t n_points, double** blob_out, int64_t* len_out) {
    if (!lumen || !ref_point || !label || !blob_out || !len_out)
        return mmrs::set_err(ctx, MMRS_ERR_ARG, "mmrs_geometry_from_arrays: NULL argument");
    return guarded(ctx, [&] {
        auto borrow = [](const double* a, int64_t n) { return Points{nullptr, a, (size_t)std::max<int64_t>(n, 0)}; };
        static const double kNoRows[4] = {0, 0, 0, 0};  // an empty layer still needs a non-null `rows` to be "set"
        Input in;
        in.a_lumen = borrow(n_lumen > 0 ? lumen : kNoRows, n_lumen);
        if (eem) in.a_eem = borrow(n_eem > 0 ? eem : kNoRows, n_eem), in.has_eem = true;
        if (calc) in.a_calc = borrow(n_calc > 0 ? calc : kNoRows, n_calc), in.has_calc = true;
        if (side) in.a_side = borrow(n_side > 0 ? side : kNoRows, n_side), in.has_side = true;
        if (records) {
            in.has_rec = true;
            for (int64_t i = 0; i < n_rec; ++i) {
                Rec r{};
                r.frame = (uint32_t)records[4 * i];
                r.diastole_phase = records[4 * i + 1] != 0.0;
                r.systole_phase = !r.diastole_phase;
                r.has1 = records[4 * i + 2] == records[4 * i + 2];
                r.m1 = records[4 * i + 2];
                r.has2 = records[4 * i + 3] == records[4 * i + 3];
                r.m2 = records[4 * i + 3];
                in.rec.push_back(r);
            }
        }
        in.ref = RawPoint{(uint32_t)ref_point[0], ref_point[1], ref_point[2], ref_point[3], false};
        Trace tr;
        const Geometry g = build_geometry(in, label, diastole != 0, icx, icy, radius, n_points);
        tr.lap("ingest: build_geometry");
        *blob_out = encode_malloc(g, len_out);
        tr.lap("ingest: encode");
    });
}

extern "C" int mmrs_process_cases(mmrs_ctx* ctx, int32_t mode, int64_t n_cases, const double* const* blobs,
                                  const int64_t* blob_lens, const mmrs_align_params* params, double** out_blobs,
                                  int64_t* out_lens, double** out_logs, int64_t* out_nlogs, int32_t* out_anomalous) {
    if (!ctx) return mmrs::set_err(nullptr, MMRS_ERR_ARG, "mmrs_process_cases: ctx is NULL (a CUDA context is required)");
    if (mode < 1 || mode > 4 || n_cases < 0 || !params || (n_cases > 0 && (!blobs || !blob_lens || !out_blobs || !out_lens || !out_logs || !out_nlogs)))
        return mmrs::set_err(ctx, MMRS_ERR_ARG, "mmrs_process_cases: bad arguments");
    const int n_in = mode >= 3 ? 4 : mode;
    const int n_out = mode == 4 ? 8 : mode == 3 ? 4 : mode;
    for (int i = 0; i < 5; ++i) ctx->stats[i] = 0;
    // Output contract: every slot is NULL / 0 on entry; on failure everything written so far is released and the slots are
    // NULL again, so the caller never owns anything after a non-zero return.
    for (int64_t k = 0; k < n_cases * n_out; ++k) out_blobs[k] = nullptr, out_lens[k] = 0;
    for (int64_t k = 0; k < n_cases * n_in; ++k) out_logs[k] = nullptr, out_nlogs[k] = 0;
    const int rc = guarded(ctx, [&] {
        Searcher S{ctx, ctx->stats};
        Trace tr;
        std::vector<Geometry> geo((size_t)n_cases * n_in);
        parallel_for(geo.size(), [&](size_t k) { geo[k] = decode(blobs[k], blob_lens[k]); });
        tr.lap("decode blobs");
        // intrapullback: every pullback of every case in one batch per stage (entry.rs:140-203)
        std::vector<Geometry*> all;
        for (auto& g : geo) all.push_back(&g);
        std::vector<WithinOut> w;
        align_within_many(S, all, *params, w);
        tr.lap("align_within_many (total)");
        for (size_t k = 0; k < geo.size(); ++k) {
            out_logs[k] = to_malloc(w[k].logs);
            out_nlogs[k] = (int64_t)w[k].logs.size() / 7;
            if (out_anomalous) out_anomalous[k] = w[k].anomalous ? 1 : 0;
        }
        // Outputs are clones taken at the moment the reference takes them (align_between.rs:91) — except at the LAST
        // dependency level of a mode, where nothing moves the geometries any more: those are post-processed and encoded in
        // place (a 400-frame x 1 000-point geometry is 40 MB; cloning four of them cost as much as encoding them). They are
        // post-processed (maybe_postprocess, entry.rs:56-69, with the case-wide anomalous flag, :279-289) and encoded in
        // parallel at the end of each dependency level.
        struct Out {
            int64_t slot;
            Geometry a, b;                          // the clones (levels that are followed by another)
            Geometry *pa = nullptr, *pb = nullptr;  // or the geometries themselves (last level)
            bool pair, anomalous;
        };
        std::vector<Out> pending;
        auto emit = [&](int64_t c, int slot, Geometry& g) {   // mode 1: its only level
            pending.push_back(Out{c * n_out + slot, Geometry{}, Geometry{}, &g, nullptr, false, false});
        };
        auto emit_pair = [&](int64_t c, int slot, Geometry& a, Geometry& b, bool last_level) {
            bool anomalous = false;
            for (int k = 0; k < n_in; ++k) anomalous = anomalous || w[c * n_in + k].anomalous;
            if (last_level) pending.push_back(Out{c * n_out + slot, Geometry{}, Geometry{}, &a, &b, true, anomalous});
            else pending.push_back(Out{c * n_out + slot, a, b, nullptr, nullptr, true, anomalous});
        };
        auto finish_outputs = [&](std::vector<Out>& batch) {
            parallel_for(batch.size(), [&](size_t i) {
                Out& o = batch[i];
                Geometry& A = o.pa ? *o.pa : o.a;
                Geometry& B = o.pb ? *o.pb : o.b;
                if (o.pair && params->postprocessing) postprocess_pair(A, B, 0.03, o.anomalous);
                out_blobs[o.slot] = encode_malloc(A, &out_lens[o.slot]);
                if (o.pair) out_blobs[o.slot + 1] = encode_malloc(B, &out_lens[o.slot + 1]);
            });
            batch.clear();
        };
        auto flush = [&] { finish_outputs(pending); };
        if (mode == 1) {
            for (int64_t c = 0; c < n_cases; ++c) emit(c, 0, geo[c]);
            flush();
            return;
        }
        // inter-pullback level 1: A<-B (and C<-D), every case at once (entry.rs:206-240, :617-666)
        std::vector<std::pair<Geometry*, Geometry*>> level;
        for (int64_t c = 0; c < n_cases; ++c) {
            Geometry* g = &geo[c * n_in];
            level.push_back({g + 0, g + 1});
            if (mode >= 3) level.push_back({g + 2, g + 3});
        }
        align_between_many(S, level, *params);
        tr.lap("align_between_many level 1");
        for (int64_t c = 0; c < n_cases; ++c) {
            Geometry* g = &geo[c * n_in];
            emit_pair(c, 0, g[0], g[1], mode != 4);
            if (mode >= 3) emit_pair(c, 2, g[2], g[3], mode != 4);  // pair CD holds C and D as they were BEFORE level 2 moves them
        }
        if (mode != 4) {
            flush();
            tr.lap("postprocess + encode level 1");
            return;
        }
        // The level-1 outputs are clones: they are post-processed and encoded on a side thread while level 2 searches
        // (std::async's future joins in its destructor, so an exception in level 2 still waits for the side thread).
        std::vector<Out> level1;
        level1.swap(pending);
        auto side = std::async(std::launch::async, [&] { finish_outputs(level1); });
        // level 2: A<-C and B<-D on the already moved B and D (entry.rs:243-277)
        level.clear();
        for (int64_t c = 0; c < n_cases; ++c) {
            Geometry* g = &geo[c * n_in];
            level.push_back({g + 0, g + 2});
            level.push_back({g + 1, g + 3});
        }
        align_between_many(S, level, *params);
        tr.lap("align_between_many level 2 (level-1 outputs encoded beside it)");
        for (int64_t c = 0; c < n_cases; ++c) {
            Geometry* g = &geo[c * n_in];
            emit_pair(c, 4, g[0], g[2], true);
            emit_pair(c, 6, g[1], g[3], true);
        }
        flush();
        side.get();
        tr.lap("level 2 encode");
    });
    if (rc != MMRS_OK) {
        for (int64_t k = 0; k < n_cases * n_out; ++k) std::free(out_blobs[k]), out_blobs[k] = nullptr, out_lens[k] = 0;
        for (int64_t k = 0; k < n_cases * n_in; ++k) std::free(out_logs[k]), out_logs[k] = nullptr, out_nlogs[k] = 0;
    }
    return rc;
}

extern "C" int mmrs_ctx_set_shard(mmrs_ctx* ctx, int32_t rank, int32_t world, mmrs_exchange_fn fn, void* user) {
    if (!ctx) return mmrs::set_err(nullptr, MMRS_ERR_ARG, "mmrs_ctx_set_shard: ctx is NULL");
    if (world > 1 && fn && (rank < 0 || rank >= world)) return mmrs::set_err(ctx, MMRS_ERR_ARG, "mmrs_ctx_set_shard: bad rank");
    ctx->shard_rank = (world > 1 && fn) ? rank : 0;
    ctx->shard_world = (world > 1 && fn) ? world : 1;
    ctx->exchange = (world > 1) ? fn : nullptr;
    ctx->exchange_user = user;
    return MMRS_OK;
}

extern "C" int mmrs_process_stats(mmrs_ctx* ctx, int64_t s[5]) {
    if (!ctx || !s) return mmrs::set_err(ctx, MMRS_ERR_ARG, "mmrs_process_stats: NULL argument");
    for (int i = 0; i < 5; ++i) s[i] = ctx->stats[i];
    return MMRS_OK;
}

// ---- OBJ / MTL / PNG export ---------------------------------------------------------------
static std::vector<int> kinds_vec(const int32_t* kinds, int32_t n) {
    std::vector<int> v;
    for (int32_t i = 0; i < n; ++i) {
        if (kinds[i] < 0 || kinds[i] > 5) throw InputErr("contour kind out of range");
        v.push_back(kinds[i]);
    }
    return v;
}

extern "C" int mmrs_export_pair(mmrs_ctx* ctx, const double* blob_a, int64_t len_a, const double* blob_b, int64_t len_b,
                                const char* label_a, const char* case_name, const char* output_dir,
                                int64_t interpolation_steps, int32_t watertight, const int32_t* kinds, int32_t n_kinds) {
    if (!blob_a || !blob_b || !case_name || !output_dir || (n_kinds > 0 && !kinds) || interpolation_steps < 0)
        return mmrs::set_err(ctx, MMRS_ERR_ARG, "mmrs_export_pair: bad arguments");
    return guarded(ctx, [&] {
        Geometry a = decode(blob_a, len_a), b = decode(blob_b, len_b);
        a.label = label_a ? label_a : "";
        export_pair(a, b, case_name, output_dir, (size_t)interpolation_steps, watertight != 0, kinds_vec(kinds, n_kinds));
    });
}

// Contour::{area, find_farthest_points, find_closest_opposite, find_closest_opposite_3d} of the reference's value types
// (types/native/contour.rs:227-363), the loops as they stand there, on a packed (n, 3) f64 array: what the Python
// value types call instead of n x n numpy temporaries.
extern "C" int mmrs_contour_metrics(const double* xyz, int64_t n, int32_t has_centroid, const double* centroid,
                                    double out[8]) {
    if (!xyz || !out || n < 0 || (has_centroid && !centroid))
        return mmrs::set_err(nullptr, MMRS_ERR_ARG, "mmrs_contour_metrics: bad arguments");
    const double nan = std::numeric_limits<double>::quiet_NaN();
    for (int k = 0; k < 8; ++k) out[k] = nan;
    auto X = [&](int64_t i) { return xyz[3 * i]; };
    auto Y = [&](int64_t i) { return xyz[3 * i + 1]; };
    auto Z = [&](int64_t i) { return xyz[3 * i + 2]; };
    // area (contour.rs:345-363): half the norm of the summed cross products of consecutive points
    if (n < 3) {
        out[0] = 0.0;
    } else {
        double cx = 0.0, cy = 0.0, cz = 0.0;
        for (int64_t i = 0; i < n; ++i) {
            const int64_t j = (i + 1) % n;
            cx = cx + (Y(i) * Z(j) - Z(i) * Y(j));
            cy = cy + (Z(i) * X(j) - X(i) * Z(j));
            cz = cz + (X(i) * Y(j) - Y(i) * X(j));
        }
        out[0] = 0.5 * std::sqrt(cx * cx + cy * cy + cz * cz);
    }
    if (n == 0) return 0;
    // farthest pair (contour.rs:227-243): strict `>` from ((p0, p0), 0.0)
    {
        int64_t bi = 0, bj = 0;
        double best = 0.0;
        for (int64_t i = 0; i < n; ++i)
            for (int64_t j = i + 1; j < n; ++j) {
                const double dx = X(i) - X(j), dy = Y(i) - Y(j), dz = Z(i) - Z(j);
                const double d = std::sqrt(dx * dx + dy * dy + dz * dz);
                if (d > best) best = d, bi = i, bj = j;
            }
        out[1] = (double)bi, out[2] = (double)bj, out[3] = best;
    }
    if (n <= 2) return 0;  // the two "closest opposite" searches assert n > 2
    {  // find_closest_opposite (contour.rs:247-296)
        double cx, cy;
        if (has_centroid) {
            cx = centroid[0], cy = centroid[1];
        } else {
            double sx = 0.0, sy = 0.0;
            for (int64_t i = 0; i < n; ++i) sx = sx + X(i), sy = sy + Y(i);
            cx = sx / (double)n, cy = sy / (double)n;
        }
        std::vector<double> th((size_t)n);
        for (int64_t i = 0; i < n; ++i) {
            double t = std::atan2(Y(i) - cy, X(i) - cx);
            if (t < 0.0) t += 2.0 * kPi;
            th[(size_t)i] = t;
        }
        double min_dist = std::numeric_limits<double>::max();
        int64_t pi = 0, pj = 1;
        for (int64_t i = 0; i < n; ++i) {
            double best_diff = std::numeric_limits<double>::max();
            int64_t bj = i;
            for (int64_t j = 0; j < n; ++j) {
                if (j == i) continue;
                double delta = std::fabs(th[(size_t)j] - th[(size_t)i]);
                if (delta > kPi) delta = 2.0 * kPi - delta;
                const double diff = std::fabs(delta - kPi);
                if (diff < best_diff) best_diff = diff, bj = j;
            }
            const double dx = X(i) - X(bj), dy = Y(i) - Y(bj);
            const double d = std::sqrt(dx * dx + dy * dy);
            if (d < min_dist) min_dist = d, pi = i, pj = bj;
        }
        out[4] = (double)pi, out[5] = (double)pj, out[6] = min_dist;
    }
    {  // find_closest_opposite_3d (contour.rs:298-320): the minor axis of elliptic_ratio
        const int64_t half = n / 2;
        double min_dist = std::numeric_limits<double>::max();
        for (int64_t i = 0; i < n; ++i) {
            const int64_t j = (i + half) % n;
            const double dx = X(i) - X(j), dy = Y(i) - Y(j), dz = Z(i) - Z(j);
            const double d = std::sqrt(dx * dx + dy * dy + dz * dz);
            if (d < min_dist) min_dist = d;
        }
        out[7] = min_dist;
    }
    return 0;
}

extern "C" int mmrs_export_single(mmrs_ctx* ctx, const double* blob, int64_t len, const char* name,
                                  const char* output_dir, int32_t watertight, const int32_t* kinds, int32_t n_kinds,
                                  int32_t naming) {
    if (!blob || !name || !output_dir || (n_kinds > 0 && !kinds) || (naming < 0 || naming > 2))
        return mmrs::set_err(ctx, MMRS_ERR_ARG, "mmrs_export_single: bad arguments");
    return guarded(ctx, [&] {
        export_single(decode(blob, len), name, output_dir, watertight != 0, kinds_vec(kinds, n_kinds), naming);
    });
}

// ---- centerline alignment ----------------------------------------------------------------------
extern "C" int mmrs_align_centerline(mmrs_ctx* ctx, int32_t method, const double* centerline, int64_t n_cl,
                                     int32_t n_geoms, const double* const* blobs, const int64_t* blob_lens,
                                     const mmrs_centerline_params* params, double** out_blobs, int64_t* out_lens,
                                     double* spacing_out, double* rotation_rad_out, double* refine_out) {
    if (method < 0 || method > 2 || n_cl < 0 || (n_cl > 0 && !centerline) || (n_geoms != 1 && n_geoms != 2) || !blobs ||
        !blob_lens || !params || !out_blobs || !out_lens || (params->n_points > 0 && !params->points))
        return mmrs::set_err(ctx, MMRS_ERR_ARG, "mmrs_align_centerline: bad arguments");
    if (method == 2 && !ctx)
        return mmrs::set_err(nullptr, MMRS_ERR_ARG,
                             "mmrs_align_centerline: align_combined needs a CUDA context (no CPU fallback for the "
                             "Hausdorff refinement)");
    return guarded(ctx, [&] {
        Centerline cl((size_t)n_cl);
        for (int64_t i = 0; i < n_cl; ++i) {
            const double* r = centerline + 8 * i;
            cl[i].fi = cl[i].pi = (uint32_t)i;
            cl[i].p = {r[0], r[1], r[2]};
            cl[i].t = {r[3], r[4], r[5]};
            cl[i].branch = (uint32_t)r[6];
            cl[i].radius = r[7];
        }
        std::vector<Geometry> target;
        for (int32_t g = 0; g < n_geoms; ++g) target.push_back(decode(blobs[g], blob_lens[g]));
        CenterlineAlignOut o;
        if (ctx) {
            for (int i = 0; i < 5; ++i) ctx->stats[i] = 0;
            Searcher S{ctx, ctx->stats};
            o = align_to_centerline(&S, method, cl, target, *params);
        } else {
            o = align_to_centerline(nullptr, method, cl, target, *params);
        }
        for (int32_t g = 0; g < n_geoms; ++g) out_blobs[g] = encode_malloc(target[g], &out_lens[g]);
        if (spacing_out) *spacing_out = o.spacing;
        if (rotation_rad_out) *rotation_rad_out = o.rotation;
        if (refine_out) refine_out[0] = o.refine_hausdorff, refine_out[1] = (double)o.refine_candidates;
    });
}
