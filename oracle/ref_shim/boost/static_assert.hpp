// Stand-in for <boost/static_assert.hpp> (Boost is not installed here); lets the reference's own sources
// compile unmodified into oracle/_ref.  Test infrastructure only.
#pragma once
#define BOOST_STATIC_ASSERT(x) static_assert(x, #x)
