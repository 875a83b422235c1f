#pragma once
#include <boost/thread/all.hpp>
