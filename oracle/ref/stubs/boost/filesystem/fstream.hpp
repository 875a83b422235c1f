#pragma once
// Stand-in for <boost/filesystem/fstream.hpp>.
#include <fstream>
#include <boost/filesystem.hpp>
namespace boost { namespace filesystem {
class ifstream : public std::ifstream { public: ifstream() {} explicit ifstream(const path &p, std::ios_base::openmode m = std::ios_base::in) : std::ifstream(p.c_str(), m) {} void open(const path &p, std::ios_base::openmode m = std::ios_base::in) { std::ifstream::open(p.c_str(), m); } };
class ofstream : public std::ofstream { public: ofstream() {} explicit ofstream(const path &p, std::ios_base::openmode m = std::ios_base::out) : std::ofstream(p.c_str(), m) {} void open(const path &p, std::ios_base::openmode m = std::ios_base::out) { std::ofstream::open(p.c_str(), m); } };
} }
