#pragma once
#define BOOST_VERSION 106000
