// Stand-in for <boost/filesystem.hpp> (Boost is not installed here): std::filesystem when the translation unit is
// compiled as C++17 (the engine sources that really touch files), otherwise a bare path type — lib/FileUtil.h only
// *declares* functions over fs::path and the C++11 translation units never call them.  Test infrastructure only.
#pragma once
#include <string>
#include <vector>
#if __cplusplus >= 201703L
#include <filesystem>
namespace boost { namespace filesystem = std::filesystem; }
#else
namespace boost { namespace filesystem { class path { public: path() {} path(const std::string&) {} }; } }
#endif
