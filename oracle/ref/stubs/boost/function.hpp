#pragma once
// Stand-in for <boost/function.hpp>: the reference only uses boost::function as a callable holder.
#include <functional>
namespace boost { template <class S> using function = std::function<S>; }
