#pragma once
#include <string>
#include <algorithm>
#include <cctype>
namespace boost { inline std::string to_lower_copy(std::string s) { for (auto &c : s) c = (char) std::tolower((unsigned char) c); return s; } inline void to_lower(std::string &s) { for (auto &c : s) c = (char) std::tolower((unsigned char) c); } }
