// See obj_reader.h.
#include "obj_reader.h"

#include <algorithm>
namespace dgb { unsigned host_threads(); }   // host_threads.cpp: affinity mask capped by the cgroup CPU quota
#include <charconv>
#include <climits>
#include <cstdio>
#include <cstring>
#include <thread>

namespace dgb {
namespace {

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }   // within one line ('\n' ends it)

// operator>>(double): skip blanks, then the longest valid decimal floating literal; an optional '+' is accepted by the stream
// (from_chars does not take it), "inf"/"nan" are not.  Returns false when the extraction would set failbit.
inline bool take_double(const char*& p, const char* end, double& v)
{
    while (p < end && is_space(*p)) p++;
    const char* q = p;
    if (q < end && *q == '+') q++;
    const char* d = (q < end && *q == '-') ? q + 1 : q;
    if (d >= end || !((*d >= '0' && *d <= '9') || *d == '.')) return false;
    const auto r = std::from_chars(q, end, v, std::chars_format::general);
    if (r.ec == std::errc::invalid_argument) return false;
    if (r.ec == std::errc::result_out_of_range) {          // the stream stores +-HUGE_VAL / 0 and sets failbit; keep the value, fail
        p = r.ptr; return false;
    }
    p = r.ptr;
    // num_get accumulates "1e" / "1.5e+" as part of the number and then fails the conversion (value 0, failbit): from_chars stopped in
    // front of the malformed exponent instead
    if (p < end && (*p == 'e' || *p == 'E')) { v = 0.0; return false; }
    return true;
}

struct Chunk {
    std::vector<double> v;
    std::vector<uint32_t> f;
    std::string err;
    long long err_line = -1;          // line index inside the chunk
    long long lines = 0;
};

void parse_chunk(const char* b, const char* e, Chunk& c)
{
    const char* p = b;
    while (p < e) {
        const char* nl = static_cast<const char*>(memchr(p, '\n', (size_t)(e - p)));
        const char* le = nl ? nl : e;
        if (le - p >= 2 && p[0] == 'v' && p[1] == ' ') {                       // line.substr(0, 2) == "v "
            const char* q = p + 2;
            double xyz[3] = {0.0, 0.0, 0.0};
            for (int k = 0; k < 3; k++) if (!take_double(q, le, xyz[k])) break;    // a failed extraction fails the following ones too
            c.v.push_back(xyz[0]); c.v.push_back(xyz[1]); c.v.push_back(xyz[2]);
        } else if (le - p >= 2 && p[0] == 'f' && p[1] == ' ') {                // "f a/.. b/.. c/.."
            const char* q = p + 2;
            for (int k = 0; k < 3; k++) {
                while (q < le && is_space(*q)) q++;
                const char* t = q;
                while (q < le && !is_space(*q)) q++;                           // s >> buf
                const char* te = static_cast<const char*>(memchr(t, '/', (size_t)(q - t)));
                if (!te) te = q;                                               // buf.substr(0, buf.find_first_of('/'))
                // std::stoi: optional sign, digits; throws when there is no digit or the value leaves int
                const char* d = t;
                bool two_signs = false;
                if (d < te && *d == '+') { d++; two_signs = (d < te && (*d == '-' || *d == '+')); }     // strtol takes ONE sign: "+-5" throws in std::stoi
                long long val = 0;
                const auto r = std::from_chars(d, te, val, 10);
                if (two_signs || r.ec != std::errc() || val < INT_MIN || val > INT_MAX) {
                    if (c.err_line < 0) { c.err_line = c.lines; c.err = "face index is not an integer (std::stoi would throw)"; }
                    val = 1;
                }
                c.f.push_back(static_cast<uint32_t>(static_cast<int>(val) - 1));
            }
        }
        c.lines++;
        p = nl ? nl + 1 : e;
    }
}

}  // namespace

bool read_obj(const char* path, ObjData& out, std::string& err)
{
    FILE* fp = std::fopen(path, "rb");
    if (!fp) { err = std::string("Cannot open ") + path; return false; }
    std::fseek(fp, 0, SEEK_END);
    const long sz = std::ftell(fp);
    std::fseek(fp, 0, SEEK_SET);
    std::vector<char> buf((size_t)(sz > 0 ? sz : 0));
    const size_t got = buf.empty() ? 0 : std::fread(buf.data(), 1, buf.size(), fp);
    std::fclose(fp);
    if (got != buf.size()) { err = std::string("short read on ") + path; return false; }

    unsigned nt = dgb::host_threads();
    if (nt == 0) nt = 1;
    if (nt > 32) nt = 32;
    if (buf.size() < (1u << 20)) nt = 1;
    // chunk boundaries moved forward to the next line start
    std::vector<size_t> cut(nt + 1, buf.size());
    cut[0] = 0;
    for (unsigned k = 1; k < nt; k++) {
        size_t p = std::max(cut[k - 1], buf.size() * k / nt);
        const char* nl = p < buf.size() ? static_cast<const char*>(memchr(buf.data() + p, '\n', buf.size() - p)) : nullptr;
        cut[k] = nl ? (size_t)(nl - buf.data()) + 1 : buf.size();
    }
    std::vector<Chunk> chunks(nt);
    if (nt == 1) {
        parse_chunk(buf.data(), buf.data() + buf.size(), chunks[0]);
    } else {
        std::vector<std::thread> th;
        for (unsigned k = 0; k < nt; k++)
            th.emplace_back([&, k]() { parse_chunk(buf.data() + cut[k], buf.data() + cut[k + 1], chunks[k]); });
        for (auto& t : th) t.join();
    }
    long long line0 = 0;
    for (unsigned k = 0; k < nt; k++) {
        if (chunks[k].err_line >= 0) { err = std::string(path) + ":" + std::to_string(line0 + chunks[k].err_line + 1) + ": " + chunks[k].err; return false; }
        line0 += chunks[k].lines;
    }
    size_t nv = 0, nf = 0;
    for (auto& c : chunks) { nv += c.v.size(); nf += c.f.size(); }
    out.vertices.resize(nv); out.faces.resize(nf);
    size_t ov = 0, of = 0;
    for (auto& c : chunks) {
        if (!c.v.empty()) std::memcpy(out.vertices.data() + ov, c.v.data(), c.v.size() * sizeof(double));
        if (!c.f.empty()) std::memcpy(out.faces.data() + of, c.f.data(), c.f.size() * sizeof(uint32_t));
        ov += c.v.size(); of += c.f.size();
    }
    return true;
}

}  // namespace dgb
