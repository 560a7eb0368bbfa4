// Stand-in for <boost/version.hpp> (Boost is not installed here).  Test infrastructure only.
#pragma once
#define BOOST_VERSION 106500
