/* ini.cpp -- INI / LUT readers restating Core/INIReader.cpp:31-106,221-256 and Core/Curve.cpp:94-203. */
#include "pd_host.h"
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <sys/stat.h>

namespace pdh {

static const std::string kEmpty;

bool file_exists(const std::string& path) { struct stat st; return stat(path.c_str(), &st) == 0 && S_ISREG(st.st_mode); }

std::vector<std::string> split(const std::string& s, const std::string& d) {
    std::vector<std::string> out; size_t pos = 0;
    for (;;) {
        size_t e = s.find(d, pos);
        if (e == std::string::npos) { out.push_back(s.substr(pos)); break; }
        out.push_back(s.substr(pos, e - pos)); pos = e + d.size();
    }
    return out;
}

float stof_ref(const std::string& s) {
    const char* b = s.c_str(); char* e = nullptr;
    float v = strtof(b, &e);
    if (e == b) throw Error("stof: no conversion for '" + s + "'");
    return v;
}
int stoi_ref(const std::string& s) {
    const char* b = s.c_str(); char* e = nullptr;
    long v = strtol(b, &e, 10);
    if (e == b) throw Error("stoi: no conversion for '" + s + "'");
    return (int)v;
}

static bool read_lines(const std::string& path, std::vector<std::string>& lines) {
    std::ifstream fs(path, std::ios::binary);
    if (!fs.is_open()) return false;
    std::string line;
    while (std::getline(fs, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();   /* text-mode CRLF translation */
        lines.push_back(line);
    }
    return true;
}

bool Ini::load(const std::string& path) {
    ready = false; filename = path; sections.clear();
    std::vector<std::string> lines;
    if (!read_lines(path, lines)) return false;
    std::string secName; std::map<std::string, std::string>* sec = nullptr;
    const size_t npos = std::string::npos;
    for (const std::string& line : lines) {
        const size_t comm = line.find(';');
        size_t s1 = line.find('[');
        if (s1 != npos) {
            s1++;
            const size_t s2 = line.find(']');
            if (s2 != npos && s2 > s1 && s2 < comm) {
                secName = line.substr(s1, s2 - s1);
                sec = &sections.insert({secName, {}}).first->second;
            }
        } else if (!secName.empty()) {
            const size_t sp = line.find('=');
            if (sp != npos && sp > 0 && sp < comm) {
                size_t eol = line.find_first_of(";\n", sp);
                if (eol != npos) eol -= sp;
                sec->insert({line.substr(0, sp), line.substr(sp + 1, eol)});
            }
        }
    }
    ready = true;
    return true;
}
bool Ini::hasKey(const std::string& s, const std::string& k) const {
    auto is = sections.find(s); if (is == sections.end()) return false;
    return is->second.find(k) != is->second.end();
}
const std::string& Ini::getString(const std::string& s, const std::string& k) const {
    auto is = sections.find(s); if (is == sections.end()) return kEmpty;
    auto ik = is->second.find(k); if (ik == is->second.end()) return kEmpty;
    return ik->second;
}
int Ini::getInt(const std::string& s, const std::string& k) const { const std::string& v = getString(s, k); return v.empty() ? 0 : stoi_ref(v); }
float Ini::getFloat(const std::string& s, const std::string& k) const { const std::string& v = getString(s, k); return v.empty() ? 0.0f : stof_ref(v); }
void Ini::getFloat3(const std::string& s, const std::string& k, float* o) const {
    o[0] = o[1] = o[2] = 0;
    auto v = split(getString(s, k), ",");
    if (v.size() == 3) { o[0] = stof_ref(v[0]); o[1] = stof_ref(v[1]); o[2] = stof_ref(v[2]); }
}
bool Ini::tryGetInt(const std::string& s, const std::string& k, int& out) const { if (!hasKey(s, k)) return false; out = getInt(s, k); return true; }
bool Ini::tryGetFloat(const std::string& s, const std::string& k, float& out) const { if (!hasKey(s, k)) return false; out = getFloat(s, k); return true; }
bool Ini::tryGetString(const std::string& s, const std::string& k, std::string& out) const { if (!hasKey(s, k)) return false; out = getString(s, k); return true; }

void curve_add(PdCurve& c, float ref, float val) {
    if (c.n >= PD_CURVE_MAX) throw Error("curve has more than PD_CURVE_MAX points");
    c.ref[c.n] = ref; c.val[c.n] = val; c.n++;
}
float curve_value(const PdCurve& c, float ref) {
    if (c.n <= 0) return 0.0f;
    if (ref <= c.ref[0]) return c.val[0];
    for (int id = 1; id < c.n; ++id)
        if (ref <= c.ref[id]) return (((c.val[id] - c.val[id - 1]) * (ref - c.ref[id - 1])) / (c.ref[id] - c.ref[id - 1])) + c.val[id - 1];
    return c.val[c.n - 1];
}
bool load_curve(const std::string& path, PdCurve& out) {
    memset(&out, 0, sizeof(out));
    std::vector<std::string> lines;
    if (!read_lines(path, lines)) return false;
    for (std::string line : lines) {
        const size_t comm = line.find(';');
        if (comm != std::string::npos) line = line.substr(0, comm);
        if (!line.empty()) { auto kv = split(line, "|"); if (kv.size() == 2) curve_add(out, stof_ref(kv[0]), stof_ref(kv[1])); }
    }
    return out.n > 0;
}
bool parse_inline_curve(const std::string& str, PdCurve& out) {
    memset(&out, 0, sizeof(out));
    size_t p1 = str.find("("), p2 = str.find(")");
    if (p1 != std::string::npos && p2 != std::string::npos && p1 < p2) {
        p1++;
        for (auto& pair : split(str.substr(p1, p2 - p1), "|"))
            if (!pair.empty()) { auto kv = split(pair, "="); if (kv.size() == 2) curve_add(out, stof_ref(kv[0]), stof_ref(kv[1])); }
    }
    return out.n > 0;
}
PdCurve Ini::getCurve(const std::string& s, const std::string& k) const {
    PdCurve c; memset(&c, 0, sizeof(c));
    const std::string& kv = getString(s, k);
    if (kv.find(".lut") != std::string::npos) {
        size_t slash = filename.find_last_of("/\\");
        load_curve((slash == std::string::npos ? std::string() : filename.substr(0, slash + 1)) + kv, c);
    } else parse_inline_curve(kv, c);
    return c;
}

} // namespace pdh
