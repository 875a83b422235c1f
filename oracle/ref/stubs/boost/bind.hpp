#pragma once
// Stand-in for <boost/bind.hpp> (global placeholders _1.._4 as Boost.Bind provides them).
#include <functional>
namespace boost { using std::bind; using std::ref; using std::cref; }
using namespace std::placeholders;
