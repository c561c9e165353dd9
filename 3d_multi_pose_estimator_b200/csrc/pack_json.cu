// Host-side frame packer: the reference's frame JSON -> the packed skeleton batch of include/b200pose.h.
//
// Replaces, for whole files of frames, the per-frame `json.loads` + per-joint Python loops of the reference
// (skeleton_matching/graph_generator.py:573-605, test/metrics_from_model.py:182-191) and of pack.py, which at
// ~1 ms per frame would feed the GPU path 300x slower than it computes. Frame schema (SURVEY.md App. A,
// panoptic_conversor/get_joints_from_panoptic_model_multi.py:236,281,287):
//     file  = [frame, ...]            (a single frame object is accepted too)
//     frame = {camera: [skeletons, timestamp, 'no_image', bodies_3D], ...}      only element 0 is read
//     skeletons = a JSON *string* holding  [ {"<joint>": [joint, x, y, valid, prob], ..., "ID": ...}, ... ]
//                 (or that list inline)
// Head order = frame-dict camera order restricted to the configured cameras, then skeleton order; skeletons
// without joint keys are skipped (graph_generator.py:583-601); "ID" keys are ignored (:483). Numbers are
// converted exactly like Python's float(): a Clinger fast path for short decimals, strtod otherwise.
// Pure host code (no CUDA calls): frames are parsed in parallel by std::thread workers.
#include "common.cuh"
#include <locale.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <thread>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <vector>

namespace b200pose {

struct Skel {
    double xy[B200POSE_N_JOINTS][2];
    float vp[B200POSE_N_JOINTS][2];
    uint32_t mask;
    int32_t cam;
    int32_t index;      // position in its camera's skeleton list
};

struct FrameOut {
    std::vector<Skel> sk;
    int64_t enodes = 0;
    std::string error;
};

struct PackCfg {
    std::vector<std::string> names;
    std::vector<int32_t> cam_index;
};

// ---- a small JSON scanner over [p, end) ----------------------------------------------------------
struct Cur {
    const char* p;
    const char* end;
    std::string* err;
    bool fail(const char* what) {
        if (err && err->empty()) *err = what;
        return false;
    }
    void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p; }
    bool eat(char c) { ws(); if (p < end && *p == c) { ++p; return true; } return false; }
    bool peek(char c) { ws(); return p < end && *p == c; }
};

// end of the string whose opening quote is at p[-1]: pointer to the closing quote (or end). memchr over the
// bulk, then the parity of the backslashes in front of the quote decides whether it is escaped.
static const char* string_end(const char* p, const char* end) {
    while (p < end) {
        const char* q = static_cast<const char*>(memchr(p, '"', (size_t)(end - p)));
        if (!q) return end;
        const char* b = q;
        while (b > p && b[-1] == '\\') --b;
        if (((q - b) & 1) == 0) return q;
        p = q + 1;
    }
    return end;
}

// raw string token: on entry p is at the opening quote; returns [b, e) of the body (escapes not decoded)
static bool scan_string(Cur& c, const char*& b, const char*& e, bool& has_escape) {
    c.ws();
    if (c.p >= c.end || *c.p != '"') return c.fail("expected a string");
    ++c.p;
    b = c.p;
    e = string_end(b, c.end);
    if (e >= c.end) return c.fail("unterminated string");
    has_escape = memchr(b, '\\', (size_t)(e - b)) != nullptr;
    c.p = e + 1;
    return true;
}

static void append_utf8(std::string& out, unsigned cp) {
    if (cp < 0x80) out.push_back((char)cp);
    else if (cp < 0x800) { out.push_back((char)(0xC0 | (cp >> 6))); out.push_back((char)(0x80 | (cp & 0x3F))); }
    else { out.push_back((char)(0xE0 | (cp >> 12))); out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out.push_back((char)(0x80 | (cp & 0x3F))); }
}

static void unescape(const char* b, const char* e, std::string& out) {
    out.clear();
    out.reserve((size_t)(e - b));
    for (const char* p = b; p < e; ++p) {
        if (*p != '\\' || p + 1 >= e) { out.push_back(*p); continue; }
        ++p;
        switch (*p) {
            case 'n': out.push_back('\n'); break;
            case 't': out.push_back('\t'); break;
            case 'r': out.push_back('\r'); break;
            case 'b': out.push_back('\b'); break;
            case 'f': out.push_back('\f'); break;
            case 'u': {
                unsigned cp = 0;
                for (int i = 0; i < 4 && p + 1 < e; ++i) {
                    ++p;
                    const char h = *p;
                    cp = cp * 16 + (unsigned)(h >= '0' && h <= '9' ? h - '0' : (h | 32) >= 'a' && (h | 32) <= 'f' ? (h | 32) - 'a' + 10 : 0);
                }
                append_utf8(out, cp);
                break;
            }
            default: out.push_back(*p);      // \" \\ \/
        }
    }
}

// Skips one JSON value without building anything: depth counting over a 256-entry class table (most of a test
// file is ground-truth arrays and detections of cameras the configuration does not use).
static bool skip_value(Cur& c) {
    static const struct Table {
        unsigned char t[256];
        Table() {
            memset(t, 0, sizeof(t));
            t[(unsigned char)'"'] = 1; t[(unsigned char)'{'] = 2; t[(unsigned char)'['] = 2;
            t[(unsigned char)'}'] = 3; t[(unsigned char)']'] = 3;
        }
    } T;
    c.ws();
    if (c.p >= c.end) return c.fail("unexpected end of input");
    const char ch = *c.p;
    if (ch == '"') {
        const char* q = string_end(c.p + 1, c.end);
        if (q >= c.end) return c.fail("unterminated string");
        c.p = q + 1;
        return true;
    }
    if (ch != '{' && ch != '[') {             // number / true / false / null
        while (c.p < c.end && *c.p != ',' && *c.p != ']' && *c.p != '}' && *c.p != ' ' && *c.p != '\n' && *c.p != '\t' && *c.p != '\r') ++c.p;
        return true;
    }
    int depth = 0;
    const char* p = c.p;
    while (p < c.end) {
        while (p < c.end && !T.t[(unsigned char)*p]) ++p;
        if (p >= c.end) break;
        const unsigned char k = T.t[(unsigned char)*p];
        ++p;
        if (k == 1) {
            p = string_end(p, c.end);
            if (p >= c.end) break;
            ++p;
        } else if (k == 2) {
            ++depth;
        } else if (--depth == 0) {
            c.p = p;
            return true;
        }
    }
    c.p = c.end;
    return c.fail("unterminated container");
}

static locale_t c_locale() {
    static locale_t loc = newlocale(LC_ALL_MASK, "C", (locale_t)0);
    return loc;
}

// float(text) exactly as CPython does: correctly rounded. Fast path (Clinger): <= 15 significant digits and
// |exp10| <= 22 are exact in double arithmetic; everything else goes through strtod.
static bool parse_number(Cur& c, double& out) {
    c.ws();
    const char* s = c.p;
    const char* p = s;
    bool neg = false;
    if (p < c.end && (*p == '-' || *p == '+')) { neg = *p == '-'; ++p; }
    uint64_t mant = 0;
    int digits = 0, exp10 = 0;
    bool any = false, simple = true;
    while (p < c.end && *p >= '0' && *p <= '9') { if (digits < 19) { mant = mant * 10 + (uint64_t)(*p - '0'); if (mant) ++digits; } else simple = false; ++p; any = true; }
    if (p < c.end && *p == '.') {
        ++p;
        while (p < c.end && *p >= '0' && *p <= '9') { if (digits < 19) { mant = mant * 10 + (uint64_t)(*p - '0'); if (mant) ++digits; --exp10; } else simple = false; ++p; any = true; }
    }
    if (p < c.end && (*p == 'e' || *p == 'E')) {
        ++p;
        bool eneg = false;
        if (p < c.end && (*p == '-' || *p == '+')) { eneg = *p == '-'; ++p; }
        int ev = 0;
        while (p < c.end && *p >= '0' && *p <= '9') { if (ev < 10000) ev = ev * 10 + (*p - '0'); ++p; }
        exp10 += eneg ? -ev : ev;
    }
    if (!any) {
        // not a digit string: the literals Python's json module accepts in a number slot. The token is delimited inside
        // [json, json + len) - the span need not be NUL-terminated - and compared explicitly (no strtod on caller memory).
        const char* q = p;
        while (q < c.end && *q != ',' && *q != ']' && *q != '}' && *q != ' ' && *q != '\n' && *q != '\t' && *q != '\r') ++q;
        const size_t n = (size_t)(q - p);
        auto is = [&](const char* lit) { return n == strlen(lit) && memcmp(p, lit, n) == 0; };
        if (is("true")) out = 1.0;                    // float(True): what the reference's tensor / numpy stores make of it
        else if (is("false")) out = 0.0;
        else if (is("NaN")) out = NAN;
        else if (is("Infinity")) out = neg ? -INFINITY : INFINITY;
        else return c.fail("a joint entry is not a number (null or a malformed literal)");
        c.p = q;
        return true;
    }
    static const double p10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16,
                                   1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    bool done = false;
    if (simple && digits <= 15 && exp10 >= -22 && exp10 <= 22) {
        double v = (double)mant;
        v = exp10 < 0 ? v / p10[-exp10] : v * p10[exp10];
        out = neg ? -v : v;
        done = true;
    }
#if defined(__x86_64__) && defined(__SIZEOF_LONG_DOUBLE__) && (__LDBL_MANT_DIG__ == 64)
    // 16-19 significant digits (what repr(float) produces): the mantissa is exact in an x87 long double and so is
    // 10^|e| for |e| <= 27, so one 64-bit-mantissa division/multiplication is within 2^-64 relative of the true
    // value; rounding that to double is the correctly rounded answer unless the 11 dropped bits sit within 1 unit
    // of the half-way pattern - then (about 3 in 2048 numbers) strtod decides.
    if (!done && simple && mant != 0 && exp10 >= -27 && exp10 <= 27) {
        static const long double lp10[28] = {1e0L, 1e1L, 1e2L, 1e3L, 1e4L, 1e5L, 1e6L, 1e7L, 1e8L, 1e9L, 1e10L, 1e11L, 1e12L, 1e13L,
                                              1e14L, 1e15L, 1e16L, 1e17L, 1e18L, 1e19L, 1e20L, 1e21L, 1e22L, 1e23L, 1e24L, 1e25L,
                                              1e26L, 1e27L};
        long double r = (long double)mant;
        r = exp10 < 0 ? r / lp10[-exp10] : r * lp10[exp10];
        uint64_t m64;
        memcpy(&m64, &r, sizeof(m64));                    // the explicit 64-bit significand of the x87 format
        const unsigned dropped = (unsigned)(m64 & 0x7FFu);
        const double v = (double)r;
        if ((dropped < 0x3FFu || dropped > 0x401u) && v >= 2.2250738585072014e-308 && v < 1.7976931348623157e308) {
            out = neg ? -v : v;
            done = true;
        }
    }
#endif
    if (!done) {
        char buf[64];
        const size_t n = (size_t)(p - s);
        // bounded NUL-terminated copy, parsed in the "C" locale whatever LC_NUMERIC the host application set (Qt viewers
        // switch to the user's locale; a decimal comma would truncate every number here)
        if (n < sizeof(buf)) { memcpy(buf, s, n); buf[n] = 0; out = strtod_l(buf, nullptr, c_locale()); }
        else { std::string t(s, n); out = strtod_l(t.c_str(), nullptr, c_locale()); }
    }
    c.p = p;
    return true;
}

// one skeleton object; returns false on malformed input
static bool parse_skeleton(Cur& c, Skel& sk) {
    memset(&sk, 0, sizeof(Skel));
    if (!c.eat('{')) return c.fail("expected a skeleton object");
    if (c.eat('}')) return true;
    while (true) {
        const char *kb, *ke; bool esc;
        if (!scan_string(c, kb, ke, esc)) return false;
        if (!c.eat(':')) return c.fail("expected ':' in a skeleton");
        bool is_joint = ke > kb;
        int j = 0;
        for (const char* q = kb; q < ke; ++q) { if (*q < '0' || *q > '9') { is_joint = false; break; } j = j * 10 + (*q - '0'); if (j > 1000) break; }
        if (!is_joint) {                      // "ID" (graph_generator.py:483) or anything non-numeric
            if (!(ke - kb == 2 && kb[0] == 'I' && kb[1] == 'D')) return c.fail("skeleton key is neither a joint id nor \"ID\"");
            if (!skip_value(c)) return false;
        } else {
            if (j >= B200POSE_N_JOINTS) return c.fail("joint id out of range (COCO-18 expected)");
            if (!c.eat('[')) return c.fail("expected a joint array");
            double v[5] = {0, 0, 0, 0, 0};
            int n = 0;
            if (!c.peek(']')) {
                while (true) {
                    double x;
                    if (!parse_number(c, x)) return false;
                    if (n < 5) v[n] = x;
                    ++n;
                    if (c.eat(',')) continue;
                    break;
                }
            }
            if (!c.eat(']')) return c.fail("expected ']' after a joint");
            if (n < 5) return c.fail("a joint needs 5 numbers [joint, x, y, valid, prob]");
            sk.xy[j][0] = v[1]; sk.xy[j][1] = v[2];
            sk.vp[j][0] = (float)v[3]; sk.vp[j][1] = (float)v[4];
            sk.mask |= 1u << j;
        }
        if (c.eat(',')) continue;
        if (c.eat('}')) return true;
        return c.fail("expected ',' or '}' in a skeleton");
    }
}

static bool parse_skeleton_list(Cur& c, int32_t cam, FrameOut& out, int& n_kept) {
    n_kept = 0;
    if (!c.eat('[')) return c.fail("expected a list of skeletons");
    if (c.eat(']')) return true;
    int idx = 0;
    while (true) {
        Skel sk;
        if (!parse_skeleton(c, sk)) return false;
        if (sk.mask) { sk.cam = cam; sk.index = idx; out.sk.push_back(sk); ++n_kept; }
        ++idx;
        if (c.eat(',')) continue;
        if (c.eat(']')) return true;
        return c.fail("expected ',' or ']' in a skeleton list");
    }
}

static void parse_frame(const char* b, const char* e, const PackCfg& cfg, FrameOut& out) {
    Cur c{b, e, &out.error};
    std::string key, inner;
    out.sk.reserve(32);
    std::vector<int> sizes;
    if (!c.eat('{')) { c.fail("a frame must be an object {camera: [...]}"); return; }
    if (c.eat('}')) return;
    while (true) {
        const char *kb, *ke; bool esc;
        if (!scan_string(c, kb, ke, esc)) return;
        if (esc) unescape(kb, ke, key); else key.assign(kb, ke);
        if (!c.eat(':')) { c.fail("expected ':' after a camera name"); return; }
        int32_t cam = -1;
        for (size_t i = 0; i < cfg.names.size(); ++i) if (cfg.names[i] == key) { cam = cfg.cam_index[i]; break; }
        if (cam < 0) {
            if (!skip_value(c)) return;
        } else {
            if (!c.eat('[')) { c.fail("a camera entry must be a list [skeletons, ...]"); return; }
            int kept = 0;
            if (c.peek('"')) {                // the skeleton list as JSON text
                const char *sb, *se; bool sesc;
                if (!scan_string(c, sb, se, sesc)) return;
                if (sesc) {
                    unescape(sb, se, inner);
                    Cur ci{inner.data(), inner.data() + inner.size(), &out.error};
                    if (!parse_skeleton_list(ci, cam, out, kept)) return;
                } else {
                    Cur ci{sb, se, &out.error};
                    if (!parse_skeleton_list(ci, cam, out, kept)) return;
                }
            } else if (c.peek('[')) {
                if (!parse_skeleton_list(c, cam, out, kept)) return;
            } else if (!c.peek(']')) {
                c.fail("element 0 of a camera entry must be the skeleton list");
                return;
            }
            while (c.eat(',')) if (!skip_value(c)) return;     // timestamp, 'no_image', ground truth, ...
            if (!c.eat(']')) { c.fail("expected ']' after a camera entry"); return; }
            if (kept) sizes.push_back(kept);
        }
        if (c.eat(',')) continue;
        if (c.eat('}')) break;
        c.fail("expected ',' or '}' in a frame");
        return;
    }
    int64_t s = 0, q = 0;
    for (int n : sizes) { s += n; q += (int64_t)n * n; }
    out.enodes = (s * s - q) / 2;             // M = sum_{i<j} n_i n_j (graph_generator.py:854-864)
}

struct Packed {
    std::vector<FrameOut> frames;
    std::vector<int32_t> head_off, node_off;
    int32_t max_heads = 0, max_enodes = 0;
};

}  // namespace b200pose

using namespace b200pose;

extern "C" __attribute__((visibility("default"))) int b200pose_pack_json(const char* json, int64_t len, int32_t n_cams,
                                  const char* const* cam_names, const int32_t* cam_index, int32_t n_threads, void** out_handle)
{
    B2_CHECK_ARG(json && len >= 0 && out_handle && (n_cams == 0 || (cam_names && cam_index)), "pack_json: null argument");
    PackCfg cfg;
    for (int i = 0; i < n_cams; ++i) { cfg.names.emplace_back(cam_names[i]); cfg.cam_index.push_back(cam_index[i]); }
    // The boundary scan (one thread: a frame ends where the top-level value ends) and the frame parsing (worker threads) run
    // as a pipeline: the scanner hands every frame span to a queue as soon as it has it, so the serial pass - a third of the
    // wall time at 8 threads when it ran first - hides behind the parsing. The caller's thread scans, then parses too.
    std::string err;
    Cur c{json, json + len, &err};
    struct Item { size_t f; const char* b; const char* e; };
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Item> queue;
    bool scan_done = false;
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    std::vector<std::vector<std::pair<size_t, FrameOut>>> parsed((size_t)nt);
    auto worker = [&](int t) {
        while (true) {
            Item it;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return !queue.empty() || scan_done; });
                if (queue.empty()) return;
                it = queue.front();
                queue.pop_front();
            }
            parsed[(size_t)t].emplace_back(it.f, FrameOut());
            parse_frame(it.b, it.e, cfg, parsed[(size_t)t].back().second);
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(worker, t);
    size_t B = 0;
    std::string scan_error;
    auto push = [&](const char* b, const char* e) {
        { std::lock_guard<std::mutex> lk(mu); queue.push_back(Item{B, b, e}); }
        cv.notify_one();
        ++B;
    };
    c.ws();
    if (c.p < c.end && *c.p == '[') {
        ++c.p;
        if (!c.eat(']')) {
            while (true) {
                c.ws();
                const char* b = c.p;
                if (!skip_value(c)) { scan_error = err + " (frame " + std::to_string(B) + ")"; break; }
                push(b, c.p);
                if (c.eat(',')) continue;
                if (c.eat(']')) break;
                scan_error = "expected ',' or ']' after frame " + std::to_string(B - 1);
                break;
            }
        }
    } else if (c.p < c.end && *c.p == '{') {
        const char* b = c.p;
        if (!skip_value(c)) scan_error = err;
        else push(b, c.p);
    } else {
        scan_error = "input must be a list of frames or one frame object";
    }
    { std::lock_guard<std::mutex> lk(mu); scan_done = true; }
    cv.notify_all();
    worker(0);
    for (auto& x : th) x.join();
    if (!scan_error.empty()) { set_error("pack_json: %s", scan_error.c_str()); return B200POSE_E_INVALID; }
    Packed* P = new Packed();
    P->frames.resize(B);
    for (auto& list : parsed)
        for (auto& fo : list) P->frames[fo.first] = std::move(fo.second);
    P->head_off.assign(B + 1, 0);
    P->node_off.assign(B + 1, 0);
    for (size_t f = 0; f < B; ++f) {
        const FrameOut& fo = P->frames[f];
        if (!fo.error.empty()) {
            set_error("pack_json: frame %zu: %s", f, fo.error.c_str());
            delete P;
            return B200POSE_E_INVALID;
        }
        const int64_t H = (int64_t)fo.sk.size();
        const int64_t h1 = P->head_off[f] + H, n1 = P->node_off[f] + H + fo.enodes;
        if (h1 > 2147483647LL || n1 > 2147483647LL) { set_error("pack_json: batch too large for 32-bit offsets"); delete P; return B200POSE_E_UNSUPPORTED; }
        P->head_off[f + 1] = (int32_t)h1;
        P->node_off[f + 1] = (int32_t)n1;
        if (H > P->max_heads) P->max_heads = (int32_t)H;
        if (fo.enodes > P->max_enodes) P->max_enodes = (int32_t)fo.enodes;
    }
    *out_handle = P;
    return B200POSE_OK;
}

extern "C" __attribute__((visibility("default"))) int b200pose_packed_sizes(const void* handle, int32_t* n_frames, int32_t* n_heads,
                                     int32_t* n_nodes, int32_t* max_heads, int32_t* max_enodes)
{
    B2_CHECK_ARG(handle, "packed_sizes: null handle");
    const Packed* P = static_cast<const Packed*>(handle);
    if (n_frames) *n_frames = (int32_t)P->frames.size();
    if (n_heads) *n_heads = P->head_off.back();
    if (n_nodes) *n_nodes = P->node_off.back();
    if (max_heads) *max_heads = P->max_heads;
    if (max_enodes) *max_enodes = P->max_enodes;
    return B200POSE_OK;
}

extern "C" __attribute__((visibility("default"))) int b200pose_packed_copy(const void* handle, double* sk_xy_host, float* sk_vp_host,
                                    uint32_t* sk_mask_host, int32_t* sk_cam_host, int32_t* head_off_host, int32_t* node_off_host,
                                    int32_t* skeleton_index_host, int32_t n_threads)
{
    B2_CHECK_ARG(handle && sk_xy_host && sk_vp_host && sk_mask_host && sk_cam_host && head_off_host && node_off_host,
                 "packed_copy: null pointer");
    const Packed* P = static_cast<const Packed*>(handle);
    const size_t B = P->frames.size();
    memcpy(head_off_host, P->head_off.data(), (B + 1) * sizeof(int32_t));
    memcpy(node_off_host, P->node_off.data(), (B + 1) * sizeof(int32_t));
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if ((size_t)nt > B) nt = B ? (int)B : 1;
    auto work = [&](int t) {
        for (size_t f = (size_t)t; f < B; f += (size_t)nt) {
            size_t h = (size_t)P->head_off[f];
            for (const Skel& s : P->frames[f].sk) {
                memcpy(sk_xy_host + h * 36, s.xy, sizeof(s.xy));
                memcpy(sk_vp_host + h * 36, s.vp, sizeof(s.vp));
                sk_mask_host[h] = s.mask;
                sk_cam_host[h] = s.cam;
                if (skeleton_index_host) skeleton_index_host[h] = s.index;
                ++h;
            }
        }
    };
    if (nt == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back(work, t);
        for (auto& x : th) x.join();
    }
    return B200POSE_OK;
}

extern "C" __attribute__((visibility("default"))) void b200pose_packed_free(void* handle)
{
    delete static_cast<Packed*>(handle);
}
