/*
 * oracle/shim/posix_core.cpp -- TEST INFRASTRUCTURE (oracle build only).
 *
 * POSIX implementations of the Windows-only plumbing the reference's Core layer declares:
 *   Core/OS.h (OS.cpp is `#error` off Windows, Core/OS.cpp:246-250),
 *   Core/Diag.h log functions (Core/Diag.cpp uses _wfopen_s),
 *   Core/String.h (MSVC wide printf treats %s as a wide string; glibc does not),
 *   Core/SharedMemory.h (Win32 file mapping; interop is disabled, cfg/sim.ini:10-13).
 * None of this is on the arithmetic path; it only lets the unmodified Car/Sim/Core sources link.
 */
#include "Core/OS.h"
#include "Core/Diag.h"
#include "Core/String.h"
#include "Core/SharedMemory.h"
#include <cstdarg>
#include <sys/stat.h>
#include <unistd.h>
#include <ctime>

namespace D {

static bool g_logQuiet = true;
static std::wstring g_logFile;

static std::string narrow(const std::wstring& w) { std::string s; for (wchar_t c : w) s.push_back((char)c); return s; }
static std::wstring widen(const std::string& s) { std::wstring w; for (char c : s) w.push_back((wchar_t)(unsigned char)c); return w; }

/* ---- String.h ---- */
std::string stra(const wchar_t* str) { return narrow(std::wstring(str)); }
std::string stra(const std::wstring& str) { return narrow(str); }
std::wstring strw(const char* str) { return widen(std::string(str)); }
std::wstring strw(const std::string& str) { return widen(str); }

std::string strafv(const char* format, va_list args) {
    char buf[1024]; int n = vsnprintf(buf, sizeof(buf), format, args);
    return n > 0 ? std::string(buf, (size_t)std::min(n, 1023)) : std::string();
}
std::string straf(const char* format, ...) { va_list a; va_start(a, format); auto s = strafv(format, a); va_end(a); return s; }

/* MSVC wide printf: %s = wide string, %S = narrow string.  glibc: %ls / %s. */
static std::wstring msvc_to_glibc_format(const wchar_t* f) {
    std::wstring o;
    for (; *f; ++f) {
        if (*f != L'%') { o.push_back(*f); continue; }
        o.push_back(*f++);
        if (*f == L'%') { o.push_back(*f); continue; }
        while (*f && wcschr(L"-+ #0123456789.*", *f)) o.push_back(*f++);
        if (*f == L's') { o += L"ls"; }
        else if (*f == L'S') { o += L"s"; }
        else if (*f) { o.push_back(*f); }
        else break;
    }
    return o;
}
std::wstring strwfv(const wchar_t* format, va_list args) {
    wchar_t buf[1024];
    std::wstring f = msvc_to_glibc_format(format);
    int n = vswprintf(buf, 1024, f.c_str(), args);
    return n > 0 ? std::wstring(buf, (size_t)n) : std::wstring();
}
std::wstring strwf(const wchar_t* format, ...) { va_list a; va_start(a, format); auto s = strwfv(format, a); va_end(a); return s; }

template<typename T> static std::vector<T> split_t(const T& s, const T& d) {
    std::vector<T> out; size_t pos = 0;
    while (true) {
        size_t e = s.find(d, pos);
        if (e == T::npos) { out.push_back(s.substr(pos)); return out; }
        out.push_back(s.substr(pos, e - pos)); pos = e + d.size();
    }
}
std::vector<std::wstring> split(const std::wstring& s, const std::wstring& delim) { return split_t(s, delim); }
std::vector<std::string> split(const std::string& s, const std::string& delim) { return split_t(s, delim); }
void replace(std::wstring& s, wchar_t from, wchar_t to) { for (auto& c : s) if (c == from) c = to; }
void replace(std::wstring& s, const std::wstring& from, const std::wstring& to) {
    size_t p = 0; while ((p = s.find(from, p)) != std::wstring::npos) { s.replace(p, from.size(), to); p += to.size(); }
}
bool ends_with(const std::wstring& s, wchar_t ch) { return !s.empty() && s.back() == ch; }

/* ---- Diag.h ---- */
void log_set_file(const wchar_t* filename, bool overwrite) { g_logFile = filename; if (overwrite) log_clear_file(); }
void log_clear_file() { if (!g_logFile.empty()) { FILE* f = fopen(narrow(g_logFile).c_str(), "wt"); if (f) fclose(f); } }
void log_printf(const wchar_t* format, ...) {
    if (g_logQuiet && g_logFile.empty()) return;
    va_list a; va_start(a, format); auto s = strwfv(format, a); va_end(a);
    if (s.empty()) return;
    if (!g_logQuiet) { fputs(narrow(s).c_str(), stderr); fputc('\n', stderr); }
    if (!g_logFile.empty()) { FILE* f = fopen(narrow(g_logFile).c_str(), "at"); if (f) { fputs(narrow(s).c_str(), f); fputc('\n', f); fclose(f); } }
}
void trace_warn(const wchar_t* msg, const char* file, int line) { log_printf(L"%s [FILE: %S LINE: %d]", msg, file, line); }
void trace_error(const wchar_t* msg, const char* file, int line) {
    fprintf(stderr, "[oracle] %s [FILE: %s LINE: %d]\n", narrow(msg).c_str(), file, line);
}
extern "C" void pdref_set_log_quiet(int q) { g_logQuiet = q != 0; }

/* ---- OS.h ---- */
void osTraceDebug(const wchar_t*) {}
unsigned int osGetCurrentProcessId() { return (unsigned)getpid(); }
unsigned int osGetCurrentThreadId() { return 1; } /* the oracle harness is single-threaded per simulator */
unsigned int osGetCurrentTicks() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (unsigned)(t.tv_sec * 1000 + t.tv_nsec / 1000000); }
void* osLoadLibraryA(const char*) { return nullptr; }
void* osLoadLibraryW(const wchar_t*) { return nullptr; }
void* osGetProcAddress(void*, const char*) { return nullptr; }
std::wstring osGetModuleFullPath() { return L""; }
std::wstring osGetCurrentDir() { char b[4096]; return getcwd(b, sizeof(b)) ? widen(b) : L""; }
void osSetCurrentDir(const std::wstring& p) { if (chdir(narrow(p).c_str())) {} }
std::wstring osCanonicPath(const std::wstring& p) { return p; }
std::wstring osCombinePath(const std::wstring& a, const std::wstring& b) { return (a.empty() || a.back() == L'/') ? a + b : a + L"/" + b; }
std::wstring osGetDirPath(const std::wstring& p) { size_t k = p.find_last_of(L"/\\"); return k == std::wstring::npos ? L"" : p.substr(0, k + 1); }
std::wstring osGetFileName(const std::wstring& p) { size_t k = p.find_last_of(L"/\\"); return k == std::wstring::npos ? p : p.substr(k + 1); }
bool osFileExists(const std::wstring& p) { struct stat st; return stat(narrow(p).c_str(), &st) == 0 && S_ISREG(st.st_mode); }
bool osDirExists(const std::wstring& p) { struct stat st; return stat(narrow(p).c_str(), &st) == 0 && S_ISDIR(st.st_mode); }
void osEnsureDirExists(const std::wstring&) {}
void osCreateDirectoryTree(const std::wstring&) {}
void* osFindWindow(const wchar_t*, const wchar_t*) { return nullptr; }
void* osFindProcessWindow(unsigned int) { return nullptr; }

bool FileHandle::open(const wchar_t* filename, const wchar_t* mode) {
    close();
    fd = fopen(narrow(filename).c_str(), narrow(mode).c_str());
    return fd != nullptr;
}
void FileHandle::close() { if (fd) { fclose(fd); fd = nullptr; } }
size_t FileHandle::size() const {
    if (!fd) return 0;
    long cur = ftell(fd); fseek(fd, 0, SEEK_END); long n = ftell(fd); fseek(fd, cur, SEEK_SET); return (size_t)n;
}

/* ---- SharedMemory.h (interop disabled) ---- */
SharedMemory::SharedMemory() {}
SharedMemory::~SharedMemory() {}
void SharedMemory::allocate(const wchar_t*, size_t) {}
void SharedMemory::open(const wchar_t*, size_t) {}
void SharedMemory::close() {}

}
