#pragma once
#include <string>
#include <algorithm>
#include <cctype>
namespace boost { inline std::string to_lower_copy(std::string s) { for (auto &c : s) c = (char) std::tolower((unsigned char) c); return s; } inline bool starts_with(const std::string &s, const std::string &p) { return s.size() >= p.size() && s.compare(0, p.size(), p) == 0; }
inline bool ends_with(const std::string &s, const std::string &p) { return s.size() >= p.size() && s.compare(s.size() - p.size(), p.size(), p) == 0; }
inline void to_lower(std::string &s) { for (auto &c : s) c = (char) std::tolower((unsigned char) c); }  namespace algorithm { using boost::starts_with; using boost::ends_with; using boost::to_lower; using boost::to_lower_copy; } }
