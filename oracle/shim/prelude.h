/*
 * oracle/shim/prelude.h -- TEST INFRASTRUCTURE (oracle build only, never linked into the product).
 *
 * Forced-include prelude that lets g++ compile the reference's MSVC-only sources
 * (src/ProjectD/{Car,Sim,Core}) unmodified, where they lie under /root/reference.
 * It supplies the headers MSVC pulls in implicitly and one MSVC extension the sources rely on:
 * std::wifstream constructed from a std::wstring (Core/Curve.cpp:134, Core/INIReader.cpp:50,
 * Car/SetupManager.cpp:414).
 */
#pragma once
#ifdef __cplusplus
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cwchar>
#include <fstream>
#include <locale>
#include <memory>
#include <string>
#include <vector>

using std::isfinite;

namespace std {
/* wide-path file stream: narrows the path (ASCII paths only) and strips nothing else */
class pd_wifstream : public basic_ifstream<wchar_t> {
public:
    pd_wifstream() {}
    explicit pd_wifstream(const std::wstring& p) { open_w(p); }
    explicit pd_wifstream(const char* p) : basic_ifstream<wchar_t>(p) {}
    void open_w(const std::wstring& p) {
        std::string n; n.reserve(p.size());
        for (wchar_t c : p) n.push_back((char)c);
        this->open(n.c_str());
    }
};
}
#define wifstream pd_wifstream
#endif
