#pragma once
// Stand-in for <boost/filesystem.hpp>: only the path value type the reference's headers name in declarations.
#include <string>
#include <ostream>
#include <fstream>
#include <vector>
#include <dirent.h>
#include <cstdio>
#include <sys/stat.h>
#include <unistd.h>
namespace boost { namespace filesystem {
class path {
    std::string s;
public:
    path() {}
    path(const char *c) : s(c) {}
    path(const std::string &c) : s(c) {}
    const std::string &string() const { return s; }
    const char *c_str() const { return s.c_str(); }
    bool empty() const { return s.empty(); }
    path operator/(const path &o) const { return path(s.empty() ? o.s : s + "/" + o.s); }
    path &operator/=(const path &o) { *this = *this / o; return *this; }
    bool operator==(const path &o) const { return s == o.s; }
    bool operator!=(const path &o) const { return s != o.s; }
    bool operator<(const path &o) const { return s < o.s; }
    path filename() const { size_t i = s.rfind('/'); return path(i == std::string::npos ? s : s.substr(i + 1)); }
    path parent_path() const { size_t i = s.rfind('/'); return path(i == std::string::npos ? std::string() : s.substr(0, i)); }
    path extension() const { std::string f = filename().s; size_t i = f.rfind('.'); return path(i == std::string::npos ? std::string() : f.substr(i)); }
    path stem() const { std::string f = filename().s; size_t i = f.rfind('.'); return path(i == std::string::npos ? f : f.substr(0, i)); }
    path &replace_extension(const path &e = path()) { std::string f = s; size_t i = f.rfind('.'); if (i != std::string::npos && f.find('/', i) == std::string::npos) f = f.substr(0, i); s = f + e.s; return *this; }
    bool is_absolute() const { return !s.empty() && s[0] == '/'; }
    bool is_complete() const { return is_absolute(); }
};
inline std::ostream &operator<<(std::ostream &o, const path &p) { return o << p.string(); }
inline bool exists(const path &p) { struct stat st; return ::stat(p.c_str(), &st) == 0; }
inline bool is_directory(const path &p) { struct stat st; return ::stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode); }
inline bool is_regular_file(const path &p) { struct stat st; return ::stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode); }
inline unsigned long file_size(const path &p) { struct stat st; return ::stat(p.c_str(), &st) == 0 ? (unsigned long) st.st_size : 0; }
inline bool remove(const path &p) { return ::remove(p.c_str()) == 0; }
inline bool create_directory(const path &p) { return ::mkdir(p.c_str(), 0777) == 0; }
inline void rename(const path &a, const path &b) { ::rename(a.c_str(), b.c_str()); }
inline path current_path() { char buf[4096]; return path(::getcwd(buf, sizeof(buf)) ? buf : "."); }
inline path absolute(const path &p) { return p.is_absolute() ? p : current_path() / p; }
inline path complete(const path &p) { return absolute(p); }
inline path canonical(const path &p) { return absolute(p); }
inline void resize_file(const path &p, unsigned long n) { if (::truncate(p.c_str(), (off_t) n) != 0) {} }
class directory_entry { path m_p; public: directory_entry() {} explicit directory_entry(const path &p) : m_p(p) {} const filesystem::path &path() const { return m_p; } };
class directory_iterator {
    std::vector<directory_entry> m_e; size_t m_i;
public:
    directory_iterator() : m_i(0) {}
    explicit directory_iterator(const path &p) : m_i(0) {
        if (DIR *d = ::opendir(p.c_str())) { while (dirent *e = ::readdir(d)) { std::string n = e->d_name; if (n != "." && n != "..") m_e.push_back(directory_entry(p / path(n))); } ::closedir(d); }
    }
    bool operator!=(const directory_iterator &o) const { return (m_i < m_e.size()) != (o.m_i < o.m_e.size()); }
    bool operator==(const directory_iterator &o) const { return !(*this != o); }
    directory_iterator &operator++() { ++m_i; return *this; }
    const directory_entry &operator*() const { return m_e[m_i]; }
    const directory_entry *operator->() const { return &m_e[m_i]; }
};
} }
