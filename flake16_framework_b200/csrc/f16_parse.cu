// Host-only: tests.json -> arrays, the first step of `load_feat_lab_proj` (experiment.py:410-421).
//
// The reference runs `json.load` and a Python loop over every test for each of its 216 configs; the
// grid engine parses once, and at 8 GPUs that one parse (0.46 s of Python for the 40 MB file of
// 100 000 tests) is an eighth of an end-to-end pass.  This is a single-pass scanner for exactly the
// wire format `write_tests` emits (experiment.py:376-407: {project: {test id: [req_runs, label,
// f0 .. f15]}}, any whitespace): numbers go through strtod (correctly rounded, i.e. the value
// Python's float() gives; plain integers take a fast exact path), test ids are skipped (escape
// aware), project names are copied.  Anything else - escapes in a project name, a test whose list
// has another length, non-numeric items - makes it return F16_ERR_INVALID and the caller falls back
// to the json module; the two paths are compared element for element in tests/test_host_cpu.py.
#include "f16_common.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

extern "C" void f16_set_error(const char* fmt, ...);

struct f16_tests {
    std::vector<double> values;        // [n][n_cols]
    std::vector<int32_t> proj;         // [n] project index
    std::vector<std::string> names;    // project names in file order
    int n_cols;
};

namespace {
struct Scan {
    const char* p;
    const char* end;
    bool ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) p++; return p < end; }
    bool eat(char ch) { if (!ws() || *p != ch) return false; p++; return true; }
    bool peek(char ch) { return ws() && *p == ch; }
    // string without copying; *escaped = true if it holds a backslash escape
    bool str(const char** b, const char** e, bool* escaped) {
        if (!ws() || *p != '"') return false;
        p++;
        *b = p; *escaped = false;
        while (p < end && *p != '"') {
            if (*p == '\\') { *escaped = true; p++; if (p >= end) return false; }
            p++;
        }
        if (p >= end) return false;
        *e = p;
        p++;
        return true;
    }
    bool number(double* out) {
        if (!ws()) return false;
        const char* q = p;
        bool neg = false;
        if (q < end && *q == '-') { neg = true; q++; }
        if (q >= end || *q < '0' || *q > '9') return false;            // JSON numbers start with a digit (no NaN / Infinity here)
        unsigned long long iv = 0;
        int digits = 0;
        const char* d = q;
        while (d < end && *d >= '0' && *d <= '9' && digits < 18) { iv = iv * 10 + (unsigned)(*d - '0'); d++; digits++; }
        if (d < end && (*d == '.' || *d == 'e' || *d == 'E' || (*d >= '0' && *d <= '9'))) {
            char* stop = nullptr;
            *out = strtod(p, &stop);                                     // the buffer is NUL terminated
            if (stop == p) return false;
            p = stop;
            return true;
        }
        if (iv >= (1ull << 53)) {                                       // beyond exact doubles: let strtod round it
            char* stop = nullptr;
            *out = strtod(p, &stop);
            p = stop;
            return true;
        }
        *out = (neg && iv != 0) ? -(double)iv : (double)iv;            // JSON "-0" is the integer 0
        p = d;
        return true;
    }
};
}  // namespace

extern "C" int f16_tests_parse(const char* path, f16_tests** out) {
    if (!path || !out) { f16_set_error("f16_tests_parse: bad arguments"); return F16_ERR_INVALID; }
    FILE* fd = fopen(path, "rb");
    if (!fd) { f16_set_error("f16_tests_parse: cannot open %s", path); return F16_ERR_INVALID; }
    fseek(fd, 0, SEEK_END);
    long size = ftell(fd);
    fseek(fd, 0, SEEK_SET);
    std::vector<char> buf((size_t)size + 1);
    if (size > 0 && fread(buf.data(), 1, (size_t)size, fd) != (size_t)size) { fclose(fd); f16_set_error("f16_tests_parse: short read"); return F16_ERR_INVALID; }
    fclose(fd);
    buf[(size_t)size] = 0;
    f16_tests* T = new f16_tests();
    T->n_cols = -1;
    Scan s{buf.data(), buf.data() + size};
    auto fail = [&](const char* why) {
        f16_set_error("f16_tests_parse: unsupported input (%s) at byte %ld", why, (long)(s.p - buf.data()));
        delete T;
        return F16_ERR_INVALID;
    };
    if (!s.eat('{')) return fail("top-level object");
    if (!s.peek('}')) {
        do {
            const char *b, *e; bool esc;
            if (!s.str(&b, &e, &esc)) return fail("project name");
            if (esc) return fail("escape in a project name");
            const int pi = (int)T->names.size();
            T->names.emplace_back(b, e);
            if (!s.eat(':') || !s.eat('{')) return fail("project object");
            if (!s.peek('}')) {
                do {
                    if (!s.str(&b, &e, &esc)) return fail("test id");
                    if (!s.eat(':') || !s.eat('[')) return fail("test list");
                    int cols = 0;
                    if (!s.peek(']')) {
                        do {
                            double v;
                            if (!s.number(&v)) return fail("number");
                            T->values.push_back(v);
                            cols++;
                        } while (s.eat(','));
                    }
                    if (!s.eat(']')) return fail("end of test list");
                    if (T->n_cols < 0) T->n_cols = cols;
                    if (cols != T->n_cols || cols < 3) return fail("ragged test lists");
                    T->proj.push_back(pi);
                } while (s.eat(','));
            }
            if (!s.eat('}')) return fail("end of project object");
        } while (s.eat(','));
    }
    if (!s.eat('}')) return fail("end of top-level object");
    if (s.ws()) return fail("trailing data");
    if (T->n_cols < 0) T->n_cols = 0;
    *out = T;
    return F16_OK;
}

extern "C" int64_t f16_tests_rows(const f16_tests* T) { return T ? (int64_t)T->proj.size() : 0; }
extern "C" int32_t f16_tests_cols(const f16_tests* T) { return T ? T->n_cols : 0; }
extern "C" int32_t f16_tests_projects(const f16_tests* T) { return T ? (int32_t)T->names.size() : 0; }
extern "C" int64_t f16_tests_names_bytes(const f16_tests* T) {
    int64_t b = 0;
    if (T) for (const auto& n : T->names) b += (int64_t)n.size() + 1;
    return b;
}
// values_host: float64 [rows][cols]; proj_host: int32 [rows]; names_host: the project names, NUL separated
extern "C" int f16_tests_copy(const f16_tests* T, double* values_host, int32_t* proj_host, char* names_host) {
    if (!T || !values_host || !proj_host || !names_host) { f16_set_error("f16_tests_copy: bad arguments"); return F16_ERR_INVALID; }
    memcpy(values_host, T->values.data(), sizeof(double) * T->values.size());
    memcpy(proj_host, T->proj.data(), sizeof(int32_t) * T->proj.size());
    char* w = names_host;
    for (const auto& n : T->names) { memcpy(w, n.data(), n.size()); w += n.size(); *w++ = 0; }
    return F16_OK;
}
extern "C" void f16_tests_free(f16_tests* T) { delete T; }
